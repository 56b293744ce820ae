// Optimal-ate pairing on BLS12-381 for the Groth16 verifier and point
// (de)compression in the zcash / ark-serialize wire format.  Host + device (the portable path of field.cuh runs the
// same code on the CPU: tests/host/hostcheck.cpp diffs it against the Python oracle without a GPU).
//
// No reference counterpart: the reference verifies a SHA-256 mock (shielder/mocked_zk/src/relations.rs:127-155;
// call sites shielder/contract/lib.rs:56,74).  What this must agree with is the pairing equation the prover's
// output satisfies, checked independently by oracle/pyref/bls12_381.py (a naive pairing over
// Fq[w]/(w^12 - 2 w^6 + 2) with affine lines -- a different algorithm on a different representation).
//
// Tower: Fq2 = Fq[u]/(u^2 + 1), xi = 1 + u, Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v).
// Miller loop: G2 in homogeneous projective coordinates (Costello-Lange-Naehrig doubling/addition), M-type twist
// (y'^2 = x'^3 + 4 xi, untwist (x', y') -> (x'/w^2, y'/w^3)).  The line through T with slope lambda' evaluated at
// P, times w^3 and a factor from Fq2 (both vanish in the final exponentiation), is
//     l = c0 + (c1 xP) w^2 + (c2 yP) w^3 = c0 + (c1 xP) v + (c2 yP) v w      -- slots 0, 1, 4 of the tower,
// with (c0, c1, c2) = (3b'Z^2 - Y^2, 3X^2, -2YZ) for a tangent and (theta xQ - lambda yQ, -theta, lambda) for a chord.
// The loop runs over |x| = 0xd201000000010000 and conjugates at the end (x < 0).
// Final exponentiation: easy part f^((p^6-1)(p^2+1)), hard part via (x-1)^2 (x+p)(x^2+p^2-1) + 3 =
// 3 (p^4-p^2+1)/r (Hayashida-Hayasaka-Teruya; identity checked numerically in tests/test_oracle_pairing.py),
// so GT values are e(P,Q)^3 of the textbook reduced pairing -- the same convention as arkworks [recall]; the
// verification equation is unaffected (gcd(3, r) = 1).
#pragma once
#include "ec.cuh"
#include "pairing_consts.cuh"
#include "glv.cuh"

namespace b200zk {

// ------------------------------------------------------------------ Fq2 helpers
HD Fq2 fq2_mul_by_xi(const Fq2& a) { return Fq2{fp_sub(a.c0, a.c1), fp_add(a.c0, a.c1)}; }
HD Fq2 fq2_conj(const Fq2& a) { return Fq2{a.c0, fp_neg(a.c1)}; }
HD Fq2 fq2_mul_fq(const Fq2& a, const Fq& k) { return Fq2{fp_mul(a.c0, k), fp_mul(a.c1, k)}; }

// ------------------------------------------------------------------ Fq6
struct alignas(16) Fq6 {
    Fq2 c0, c1, c2;
    HD static Fq6 zero() { return Fq6{Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    HD static Fq6 one() { return Fq6{Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    HD bool operator==(const Fq6& b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
};
HD Fq6 fq6_add(const Fq6& a, const Fq6& b) { return Fq6{fp_add(a.c0, b.c0), fp_add(a.c1, b.c1), fp_add(a.c2, b.c2)}; }
HD Fq6 fq6_sub(const Fq6& a, const Fq6& b) { return Fq6{fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1), fp_sub(a.c2, b.c2)}; }
HD Fq6 fq6_neg(const Fq6& a) { return Fq6{fp_neg(a.c0), fp_neg(a.c1), fp_neg(a.c2)}; }
HD Fq6 fq6_mul_by_v(const Fq6& a) { return Fq6{fq2_mul_by_xi(a.c2), a.c0, a.c1}; }

HD_NOINLINE Fq6 fq6_mul(const Fq6& a, const Fq6& b) {  // 6 Fq2 products
    const Fq2 t0 = fp_mul(a.c0, b.c0), t1 = fp_mul(a.c1, b.c1), t2 = fp_mul(a.c2, b.c2);
    Fq6 r;
    r.c0 = fp_add(t0, fq2_mul_by_xi(fp_sub(fp_sub(fp_mul(fp_add(a.c1, a.c2), fp_add(b.c1, b.c2)), t1), t2)));
    r.c1 = fp_add(fp_sub(fp_sub(fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1)), t0), t1), fq2_mul_by_xi(t2));
    r.c2 = fp_add(fp_sub(fp_sub(fp_mul(fp_add(a.c0, a.c2), fp_add(b.c0, b.c2)), t0), t2), t1);
    return r;
}

HD_NOINLINE Fq6 fq6_inv(const Fq6& a) {
    const Fq2 t0 = fp_sub(fp_sqr(a.c0), fq2_mul_by_xi(fp_mul(a.c1, a.c2)));
    const Fq2 t1 = fp_sub(fq2_mul_by_xi(fp_sqr(a.c2)), fp_mul(a.c0, a.c1));
    const Fq2 t2 = fp_sub(fp_sqr(a.c1), fp_mul(a.c0, a.c2));
    const Fq2 d = fp_add(fp_mul(a.c0, t0), fq2_mul_by_xi(fp_add(fp_mul(a.c2, t1), fp_mul(a.c1, t2))));
    const Fq2 di = fp_inv(d);
    return Fq6{fp_mul(t0, di), fp_mul(t1, di), fp_mul(t2, di)};
}

// ------------------------------------------------------------------ Fq12
struct alignas(16) Fq12 {
    Fq6 c0, c1;
    HD static Fq12 one() { return Fq12{Fq6::one(), Fq6::zero()}; }
    HD bool operator==(const Fq12& b) const { return c0 == b.c0 && c1 == b.c1; }
};
HD Fq12 fq12_conj(const Fq12& a) { return Fq12{a.c0, fq6_neg(a.c1)}; }

HD_NOINLINE Fq12 fq12_mul(const Fq12& a, const Fq12& b) {  // 3 Fq6 products
    const Fq6 t0 = fq6_mul(a.c0, b.c0), t1 = fq6_mul(a.c1, b.c1);
    Fq12 r;
    r.c1 = fq6_sub(fq6_sub(fq6_mul(fq6_add(a.c0, a.c1), fq6_add(b.c0, b.c1)), t0), t1);
    r.c0 = fq6_add(t0, fq6_mul_by_v(t1));
    return r;
}

HD_NOINLINE Fq12 fq12_sqr(const Fq12& a) {  // complex squaring, 2 Fq6 products
    const Fq6 t = fq6_mul(a.c0, a.c1);
    Fq12 r;
    r.c0 = fq6_sub(fq6_sub(fq6_mul(fq6_add(a.c0, a.c1), fq6_add(a.c0, fq6_mul_by_v(a.c1))), t), fq6_mul_by_v(t));
    r.c1 = fq6_add(t, t);
    return r;
}

HD_NOINLINE Fq12 fq12_inv(const Fq12& a) {
    const Fq6 t = fq6_inv(fq6_sub(fq6_mul(a.c0, a.c0), fq6_mul_by_v(fq6_mul(a.c1, a.c1))));
    return Fq12{fq6_mul(a.c0, t), fq6_neg(fq6_mul(a.c1, t))};
}

// f^p: with f = sum_i c_i w^i (c0 = w^0, w^2, w^4; c1 = w^1, w^3, w^5), (c w^i)^p = conj(c) xi^(i(p-1)/6) w^i
HD_NOINLINE Fq12 fq12_frob(const Fq12& a) {
    using namespace pairing_consts;
    Fq12 r;
    r.c0.c0 = fq2_conj(a.c0.c0);
    r.c0.c1 = fp_mul(fq2_conj(a.c0.c1), frob_gamma(2));
    r.c0.c2 = fp_mul(fq2_conj(a.c0.c2), frob_gamma(4));
    r.c1.c0 = fp_mul(fq2_conj(a.c1.c0), frob_gamma(1));
    r.c1.c1 = fp_mul(fq2_conj(a.c1.c1), frob_gamma(3));
    r.c1.c2 = fp_mul(fq2_conj(a.c1.c2), frob_gamma(5));
    return r;
}

// f * (c0 + c1 v + c4 v w): the sparse product of a Miller line (13 Fq2 products instead of 18)
HD_NOINLINE Fq12 fq12_mul_by_014(const Fq12& f, const Fq2& c0, const Fq2& c1, const Fq2& c4) {
    // aa = f.c0 * (c0, c1, 0)
    Fq6 aa, bb, ee;
    {
        const Fq6& a = f.c0;
        const Fq2 t0 = fp_mul(a.c0, c0), t1 = fp_mul(a.c1, c1);
        aa.c0 = fp_add(t0, fq2_mul_by_xi(fp_sub(fp_mul(fp_add(a.c1, a.c2), c1), t1)));
        aa.c1 = fp_sub(fp_sub(fp_mul(fp_add(a.c0, a.c1), fp_add(c0, c1)), t0), t1);
        aa.c2 = fp_add(fp_sub(fp_mul(fp_add(a.c0, a.c2), c0), t0), t1);
    }
    {  // bb = f.c1 * (0, c4, 0)
        const Fq6& a = f.c1;
        bb.c0 = fq2_mul_by_xi(fp_mul(a.c2, c4));
        bb.c1 = fp_mul(a.c0, c4);
        bb.c2 = fp_mul(a.c1, c4);
    }
    {  // ee = (f.c0 + f.c1) * (c0, c1 + c4, 0)
        const Fq6 a = fq6_add(f.c0, f.c1);
        const Fq2 o = fp_add(c1, c4);
        const Fq2 t0 = fp_mul(a.c0, c0), t1 = fp_mul(a.c1, o);
        ee.c0 = fp_add(t0, fq2_mul_by_xi(fp_sub(fp_mul(fp_add(a.c1, a.c2), o), t1)));
        ee.c1 = fp_sub(fp_sub(fp_mul(fp_add(a.c0, a.c1), fp_add(c0, o)), t0), t1);
        ee.c2 = fp_add(fp_sub(fp_mul(fp_add(a.c0, a.c2), c0), t0), t1);
    }
    Fq12 r;
    r.c1 = fq6_sub(fq6_sub(ee, aa), bb);
    r.c0 = fq6_add(aa, fq6_mul_by_v(bb));
    return r;
}

// ------------------------------------------------------------------ Miller loop
constexpr uint64_t BLS_X_ABS = 0xd201000000010000ull;  // the curve parameter is -BLS_X_ABS

struct G2Hom {
    Fq2 x, y, z;
};

// T <- 2T; tangent line coefficients (c0, c1, c2)
HD_NOINLINE void miller_double(G2Hom& r, Fq2& l0, Fq2& l1, Fq2& l2) {
    const Fq half = pairing_consts::two_inv();
    const Fq2 a = fq2_mul_fq(fp_mul(r.x, r.y), half);
    const Fq2 b = fp_sqr(r.y);
    const Fq2 c = fp_sqr(r.z);
    const Fq2 c3 = fp_add(fp_dbl(c), c);
    // e = b' * 3c with b' = 4 xi
    const Fq2 e = fq2_mul_by_xi(fp_dbl(fp_dbl(c3)));
    const Fq2 f = fp_add(fp_dbl(e), e);
    const Fq2 g = fq2_mul_fq(fp_add(b, f), half);
    const Fq2 h = fp_sub(fp_sqr(fp_add(r.y, r.z)), fp_add(b, c));
    const Fq2 i = fp_sub(e, b);
    const Fq2 j = fp_sqr(r.x);
    const Fq2 e2 = fp_sqr(e);
    r.x = fp_mul(a, fp_sub(b, f));
    r.y = fp_sub(fp_sqr(g), fp_add(fp_dbl(e2), e2));
    r.z = fp_mul(b, h);
    l0 = i;
    l1 = fp_add(fp_dbl(j), j);
    l2 = fp_neg(h);
}

// T <- T + Q (Q affine); line through T and Q
HD_NOINLINE void miller_add(G2Hom& r, const Affine<Fq2>& q, Fq2& l0, Fq2& l1, Fq2& l2) {
    const Fq2 theta = fp_sub(r.y, fp_mul(q.y, r.z));
    const Fq2 lambda = fp_sub(r.x, fp_mul(q.x, r.z));
    const Fq2 c = fp_sqr(theta);
    const Fq2 d = fp_sqr(lambda);
    const Fq2 e = fp_mul(lambda, d);
    const Fq2 f = fp_mul(r.z, c);
    const Fq2 g = fp_mul(r.x, d);
    const Fq2 h = fp_sub(fp_add(e, f), fp_dbl(g));
    r.x = fp_mul(lambda, h);
    r.y = fp_sub(fp_mul(theta, fp_sub(g, h)), fp_mul(e, r.y));
    r.z = fp_mul(r.z, e);
    l0 = fp_sub(fp_mul(theta, q.x), fp_mul(lambda, q.y));
    l1 = fp_neg(theta);
    l2 = lambda;
}

// f_{|x|,Q}(P), conjugated (x < 0).  Either argument at infinity -> 1.
HD_NOINLINE Fq12 miller_loop(const Affine<Fq>& p, const Affine<Fq2>& q) {
    Fq12 f = Fq12::one();
    if (p.is_inf() || q.is_inf()) return f;
    G2Hom t{q.x, q.y, Fq2::one()};
    Fq2 l0, l1, l2;
    for (int bit = 62; bit >= 0; bit--) {  // bit 63 is the leading one
        f = fq12_sqr(f);
        miller_double(t, l0, l1, l2);
        f = fq12_mul_by_014(f, l0, fq2_mul_fq(l1, p.x), fq2_mul_fq(l2, p.y));
        if ((BLS_X_ABS >> bit) & 1) {
            miller_add(t, q, l0, l1, l2);
            f = fq12_mul_by_014(f, l0, fq2_mul_fq(l1, p.x), fq2_mul_fq(l2, p.y));
        }
    }
    return fq12_conj(f);
}

// ------------------------------------------------------------------ final exponentiation
// Squaring in the cyclotomic subgroup (Granger-Scott; the form ark-ff 0.4 uses in Fp12::cyclotomic_square
// [recall]): with a = (z0, z1), b = (z2, z3), c = (z4, z5) in Fq4 = Fq2[y]/(y^2 - xi), three Fq4 squarings of two
// Fq2 products each -- 18 Fq products instead of the 36 of fq12_sqr.  Valid after the easy part only.
HD_NOINLINE Fq12 fq12_cyclotomic_sqr(const Fq12& f) {
    const Fq2 &z0 = f.c0.c0, &z4 = f.c0.c1, &z3 = f.c0.c2, &z2 = f.c1.c0, &z1 = f.c1.c1, &z5 = f.c1.c2;
    Fq2 tmp = fp_mul(z0, z1);
    const Fq2 t0 = fp_sub(fp_sub(fp_mul(fp_add(z0, z1), fp_add(fq2_mul_by_xi(z1), z0)), tmp), fq2_mul_by_xi(tmp));
    const Fq2 t1 = fp_dbl(tmp);
    tmp = fp_mul(z2, z3);
    const Fq2 t2 = fp_sub(fp_sub(fp_mul(fp_add(z2, z3), fp_add(fq2_mul_by_xi(z3), z2)), tmp), fq2_mul_by_xi(tmp));
    const Fq2 t3 = fp_dbl(tmp);
    tmp = fp_mul(z4, z5);
    const Fq2 t4 = fp_sub(fp_sub(fp_mul(fp_add(z4, z5), fp_add(fq2_mul_by_xi(z5), z4)), tmp), fq2_mul_by_xi(tmp));
    const Fq2 t5 = fp_dbl(tmp);
    Fq12 r;
    r.c0.c0 = fp_add(fp_dbl(fp_sub(t0, z0)), t0);        // 3 t0 - 2 z0
    r.c1.c1 = fp_add(fp_dbl(fp_add(t1, z1)), t1);        // 3 t1 + 2 z1
    tmp = fq2_mul_by_xi(t5);
    r.c1.c0 = fp_add(fp_dbl(fp_add(tmp, z2)), tmp);      // 3 xi t5 + 2 z2
    r.c0.c2 = fp_add(fp_dbl(fp_sub(t4, z3)), t4);        // 3 t4 - 2 z3
    r.c0.c1 = fp_add(fp_dbl(fp_sub(t2, z4)), t2);        // 3 t2 - 2 z4
    r.c1.c2 = fp_add(fp_dbl(fp_add(t3, z5)), t3);        // 3 t3 + 2 z5
    return r;
}

// f^x for a cyclotomic f (x = -BLS_X_ABS: power by |x|, then conjugate = inverse)
HD_NOINLINE Fq12 fq12_exp_by_x(const Fq12& f) {
    Fq12 r = f;
    for (int bit = 62; bit >= 0; bit--) {
        r = fq12_cyclotomic_sqr(r);
        if ((BLS_X_ABS >> bit) & 1) r = fq12_mul(r, f);
    }
    return fq12_conj(r);
}

HD_NOINLINE Fq12 final_exponentiation(const Fq12& f) {
    // easy part: f^((p^6 - 1)(p^2 + 1))
    Fq12 r = fq12_mul(fq12_conj(f), fq12_inv(f));
    r = fq12_mul(fq12_frob(fq12_frob(r)), r);
    // hard part: r^((x-1)^2 (x+p) (x^2+p^2-1) + 3)
    Fq12 y0 = fq12_cyclotomic_sqr(r);
    Fq12 y1 = fq12_exp_by_x(r);
    Fq12 y2 = fq12_conj(r);
    y1 = fq12_mul(y1, y2);                 // r^(x-1)
    y2 = fq12_exp_by_x(y1);
    y1 = fq12_conj(y1);
    y1 = fq12_mul(y1, y2);                 // r^((x-1)^2)
    y2 = fq12_exp_by_x(y1);
    y1 = fq12_frob(y1);
    y1 = fq12_mul(y1, y2);                 // r^((x-1)^2 (x+p))
    r = fq12_mul(r, y0);                   // r^3
    y0 = fq12_exp_by_x(y1);
    y2 = fq12_exp_by_x(y0);
    y0 = fq12_frob(fq12_frob(y1));
    y1 = fq12_conj(y1);
    y1 = fq12_mul(y1, y2);
    y1 = fq12_mul(y1, y0);                 // ^(x^2 + p^2 - 1)
    return fq12_mul(r, y1);
}

// ------------------------------------------------------------------ wire format (zcash / ark-serialize compressed)
// 48-byte big-endian x with flags in the top three bits of byte 0: 0x80 compressed, 0x40 infinity, 0x20 "y is the
// lexicographically larger root"; G2 = x.c1 || x.c0 (SURVEY Appendix B [recall]; oracle/pyref/bls12_381.py g1_compress).

// canonical (non-Montgomery) limbs > (p - 1) / 2 ?
HD bool fq_lex_largest(const Fq& y_mont) {
    const Fq c = fp_from_mont(y_mont);
    // (p-1)/2 limbs
    uint32_t h[12];
    uint32_t carry = 0;
    for (int i = 11; i >= 0; i--) {
        const uint32_t m = FqCfg::mod(i);
        h[i] = (m >> 1) | (carry << 31);
        carry = m & 1;
    }
    return limbs_gt<12>(c.v, h);
}
HD bool fq2_lex_largest(const Fq2& y) { return y.c1.is_zero() ? fq_lex_largest(y.c0) : fq_lex_largest(y.c1); }

// 48 big-endian bytes (flag bits already masked by the caller) -> Montgomery Fq; false if the value is >= p
HD bool fq_from_be(const uint8_t* b, uint8_t mask0, Fq& out) {
    Fq c;
    for (int i = 0; i < 12; i++) {
        const uint8_t* q = b + 44 - 4 * i;
        uint32_t b0 = q[0];
        if (i == 11) b0 &= mask0;
        c.v[i] = (b0 << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    uint32_t m[12];
    for (int i = 0; i < 12; i++) m[i] = FqCfg::mod(i);
    if (!limbs_gt<12>(m, c.v)) return false;
    out = fp_to_mont(c);
    return true;
}
HD void fq_to_be(const Fq& a_mont, uint8_t* b) {
    const Fq c = fp_from_mont(a_mont);
    for (int i = 0; i < 12; i++) {
        uint8_t* q = b + 44 - 4 * i;
        q[0] = (uint8_t)(c.v[i] >> 24);
        q[1] = (uint8_t)(c.v[i] >> 16);
        q[2] = (uint8_t)(c.v[i] >> 8);
        q[3] = (uint8_t)c.v[i];
    }
}

HD_NOINLINE bool fq_sqrt(const Fq& a, Fq& out) {  // p = 3 mod 4
    uint32_t e[12];
    for (int i = 0; i < 12; i++) e[i] = pairing_consts::exp_p_plus_1_div_4(i);
    const Fq s = fp_pow(a, e, 12);
    out = s;
    return fp_sqr(s) == a;
}

// the norm method of oracle/pyref/bls12_381.py fq2_sqrt
HD_NOINLINE bool fq2_sqrt(const Fq2& a, Fq2& out) {
    if (a.is_zero()) {
        out = Fq2::zero();
        return true;
    }
    Fq s;
    if (a.c1.is_zero()) {
        if (fq_sqrt(a.c0, s)) {
            out = Fq2{s, Fq::zero()};
            return true;
        }
        if (fq_sqrt(fp_neg(a.c0), s)) {
            out = Fq2{Fq::zero(), s};
            return true;
        }
        return false;
    }
    Fq n;
    if (!fq_sqrt(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)), n)) return false;
    const Fq half = pairing_consts::two_inv();
    for (int k = 0; k < 2; k++) {
        const Fq nn = k == 0 ? n : fp_neg(n);
        Fq x0;
        if (!fq_sqrt(fp_mul(fp_add(a.c0, nn), half), x0) || x0.is_zero()) continue;
        const Fq x1 = fp_mul(a.c1, fp_inv(fp_dbl(x0)));
        const Fq2 cand{x0, x1};
        if (fp_sqr(cand) == a) {
            out = cand;
            return true;
        }
    }
    return false;
}

// status of a decoded point
enum { POINT_OK = 0, POINT_BAD_ENCODING = 1, POINT_NOT_ON_CURVE = 2, POINT_NOT_IN_SUBGROUP = 3 };

HD Fq g1_b() {  // 4
    const Fq one = Fq::one();
    return fp_dbl(fp_dbl(one));
}
HD Fq2 g2_b() { return Fq2{g1_b(), g1_b()}; }  // 4 (1 + u)

HD_NOINLINE int g1_decompress(const uint8_t* b, Affine<Fq>& out) {
    if (!(b[0] & 0x80)) return POINT_BAD_ENCODING;
    if (b[0] & 0x40) {
        bool rest = (b[0] & 0x3f) == 0;
        for (int i = 1; i < 48; i++) rest = rest && b[i] == 0;
        out = Affine<Fq>::inf();
        return rest ? POINT_OK : POINT_BAD_ENCODING;
    }
    Fq x, y;
    if (!fq_from_be(b, 0x1f, x)) return POINT_BAD_ENCODING;
    if (!fq_sqrt(fp_add(fp_mul(fp_sqr(x), x), g1_b()), y)) return POINT_NOT_ON_CURVE;
    if (fq_lex_largest(y) != ((b[0] & 0x20) != 0)) y = fp_neg(y);
    out = Affine<Fq>{x, y};
    return POINT_OK;
}

HD_NOINLINE int g2_decompress(const uint8_t* b, Affine<Fq2>& out) {
    if (!(b[0] & 0x80)) return POINT_BAD_ENCODING;
    if (b[0] & 0x40) {
        bool rest = (b[0] & 0x3f) == 0;
        for (int i = 1; i < 96; i++) rest = rest && b[i] == 0;
        out = Affine<Fq2>::inf();
        return rest ? POINT_OK : POINT_BAD_ENCODING;
    }
    Fq2 x, y;
    if (!fq_from_be(b, 0x1f, x.c1) || !fq_from_be(b + 48, 0xff, x.c0)) return POINT_BAD_ENCODING;
    if (!fq2_sqrt(fp_add(fp_mul(fp_sqr(x), x), g2_b()), y)) return POINT_NOT_ON_CURVE;
    if (fq2_lex_largest(y) != ((b[0] & 0x20) != 0)) y = fp_neg(y);
    out = Affine<Fq2>{x, y};
    return POINT_OK;
}

HD_NOINLINE void g1_compress(const Affine<Fq>& p, uint8_t* b) {
    if (p.is_inf()) {
        for (int i = 0; i < 48; i++) b[i] = 0;
        b[0] = 0xc0;
        return;
    }
    fq_to_be(p.x, b);
    b[0] |= 0x80 | (fq_lex_largest(p.y) ? 0x20 : 0);
}
HD_NOINLINE void g2_compress(const Affine<Fq2>& p, uint8_t* b) {
    if (p.is_inf()) {
        for (int i = 0; i < 96; i++) b[i] = 0;
        b[0] = 0xc0;
        return;
    }
    fq_to_be(p.x.c1, b);
    fq_to_be(p.x.c0, b + 48);
    b[0] |= 0x80 | (fq2_lex_largest(p.y) ? 0x20 : 0);
}

// [r]P == O: the plain subgroup test (r = the Fr modulus), kept as the independent check of the fast tests below
template <class F>
HD_NOINLINE bool ec_in_subgroup_plain(const Affine<F>& p) {
    if (p.is_inf()) return true;
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = FrCfg::mod(i);
    return ec_mul_scalar(XYZZ<F>::from_affine(p), r).is_inf();
}

// same point? (XYZZ representatives are not unique)
template <class F>
HD bool ec_same_point(const XYZZ<F>& a, const XYZZ<F>& b) {
    if (a.is_inf() || b.is_inf()) return a.is_inf() && b.is_inf();
    return fp_mul(a.x, b.zz) == fp_mul(b.x, a.zz) && fp_mul(a.y, b.zzz) == fp_mul(b.y, a.zzz);
}

template <class F>
HD XYZZ<F> ec_mul_by_x_abs(const XYZZ<F>& p) {  // |z| * p, z = the curve parameter (64 bits, weight 6)
    const uint32_t k[2] = {(uint32_t)BLS_X_ABS, (uint32_t)(BLS_X_ABS >> 32)};
    return ec_mul_scalar(p, k, 2);
}

// Subgroup membership through the endomorphisms (Bowe 2019, Scott 2021): two / one multiplications by the 64-bit
// curve parameter instead of one by the 255-bit r.
//   G1: phi(P) = (beta x, y) has eigenvalue lambda = z^2 - 1 on G1 and lambda^2 + lambda + 1 = r as integers, so
//       phi(P) == [z^2 - 1] P  implies  0 = (phi^2 + phi + 1) P = [r] P, and conversely holds on G1.
//   G2: psi = twist o Frobenius o untwist has eigenvalue z on G2; psi(P) == [z] P characterises G2 (Scott,
//       "A note on group membership tests for G1, G2 and GT on BLS pairing-friendly curves", section 4).
// Both are pinned against the plain test, on subgroup points and on curve points outside it (tests/test_host.py).
HD_NOINLINE bool ec_in_subgroup(const Affine<Fq>& p) {
    if (p.is_inf()) return true;
    const XYZZ<Fq> P = XYZZ<Fq>::from_affine(p);
    XYZZ<Fq> q = ec_mul_by_x_abs(ec_mul_by_x_abs(P));   // [z^2] P (the two signs cancel)
    ec_madd(q, p, true);                                // [z^2 - 1] P
    return ec_same_point(q, XYZZ<Fq>::from_affine(Affine<Fq>{glv_phi_x(p.x), p.y}));
}
HD_NOINLINE bool ec_in_subgroup(const Affine<Fq2>& p) {
    if (p.is_inf()) return true;
    using namespace pairing_consts;
    const XYZZ<Fq2> q = ec_neg(ec_mul_by_x_abs(XYZZ<Fq2>::from_affine(p)));   // [z] P, z < 0
    const Affine<Fq2> psi{fp_mul(fq2_conj(p.x), psi_cx()), fp_mul(fq2_conj(p.y), psi_cy())};
    return ec_same_point(q, XYZZ<Fq2>::from_affine(psi));
}

template <class F>
HD bool ec_on_curve(const Affine<F>& p, const F& b) {
    if (p.is_inf()) return true;
    return fp_sqr(p.y) == fp_add(fp_mul(fp_sqr(p.x), p.x), b);
}

}  // namespace b200zk
