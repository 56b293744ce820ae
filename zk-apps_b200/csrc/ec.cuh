// Short-Weierstrass (a = 0) group law for BLS12-381 G1 (over Fq) and G2 (over Fq2), written
// once over the coordinate field F.  Accumulators use XYZZ coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): mixed addition 8M+2S, doubling 6M+4S(2M+... ) -- cheaper
// than Jacobian mixed addition (7M+4S) for the bucket sums that dominate an MSM.
// No reference counterpart (VariableBaseMSM is absent from
// /root/reference); results are compared with the oracle in affine form, which is unique.
#pragma once
#include "field.cuh"

namespace b200zk {

template <class F>
struct alignas(16) Affine {
    F x, y;  // (0, 0) encodes the point at infinity (not on either curve since b != 0)
    HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    HD static Affine inf() { return Affine{F::zero(), F::zero()}; }
};

template <class F>
struct alignas(16) XYZZ {
    F x, y, zz, zzz;  // zz == 0 <=> infinity
    HD bool is_inf() const { return zz.is_zero(); }
    HD static XYZZ inf() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
    HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }
};

// Rare-path helpers are kept out of line: they sit on branches that random inputs never take,
// and inlining them into every addition multiplies code size (and ptxas time) for nothing.

// 2 * (affine p)
template <class F>
HD_NOINLINE XYZZ<F> ec_dbl_affine(const Affine<F>& p) {
    if (p.is_inf()) return XYZZ<F>::inf();
    F u = fp_dbl(p.y);
    F v = fp_sqr(u);
    F w = fp_mul(u, v);
    F s = fp_mul(p.x, v);
    F x2 = fp_sqr(p.x);
    F m = fp_add(fp_dbl(x2), x2);
    XYZZ<F> r;
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

template <class F>
HD XYZZ<F> ec_dbl(const XYZZ<F>& p) {
    if (p.is_inf()) return p;
    F u = fp_dbl(p.y);
    F v = fp_sqr(u);
    F w = fp_mul(u, v);
    F s = fp_mul(p.x, v);
    F x2 = fp_sqr(p.x);
    F m = fp_add(fp_dbl(x2), x2);
    XYZZ<F> r;
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = fp_mul(v, p.zz);
    r.zzz = fp_mul(w, p.zzz);
    return r;
}

// How the field products inside a group operation are issued.  InlineOps expands every product in place
// (fastest ptxas schedule per product, but a G1 mixed addition is then ~4,300 instructions = 68 KB of
// straight-line code, twice the 32 KB L1.5 instruction cache: ncu shows `no_instruction` stalls on
// msm_accumulate).  CallOps issues real calls to ONE copy of the Fq product / square with the operands
// passed BY VALUE -- the device ABI keeps them in registers (no local-memory traffic, checked in SASS:
// 0 LDL/STL) -- so the hot loop shrinks to ~1,800 instructions and stays cache-resident.
struct InlineOps {
    template <class F> static HD F mul(const F& a, const F& b) { return fp_mul(a, b); }
    template <class F> static HD F sqr(const F& a) { return fp_sqr(a); }
};
#if defined(__CUDACC__)
__device__ __noinline__ inline Fq fq_mul_call(Fq a, Fq b) { return fp_mul(a, b); }
__device__ __noinline__ inline Fq fq_sqr_call(Fq a) { return fp_sqr(a); }
struct CallOps {
    static DEV Fq mul(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
    static DEV Fq sqr(const Fq& a) { return fq_sqr_call(a); }
    static DEV Fq2 mul(const Fq2& a, const Fq2& b) {  // Karatsuba around three calls; the sums stay inline
        Fq t0 = fq_mul_call(a.c0, b.c0);
        Fq t1 = fq_mul_call(a.c1, b.c1);
        Fq t2 = fq_mul_call(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
        return Fq2{fp_sub(t0, t1), fp_sub(fp_sub(t2, t0), t1)};
    }
    static DEV Fq2 sqr(const Fq2& a) {
        Fq t0 = fq_mul_call(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1));
        Fq t1 = fq_mul_call(a.c0, a.c1);
        return Fq2{t0, fp_dbl(t1)};
    }
};
// Fq2 products as calls of their own (by value: 48 registers in, 24 out) around the Fq calls: smaller still
__device__ __noinline__ inline Fq2 fq2_mul_call(Fq2 a, Fq2 b) { return CallOps::mul(a, b); }
__device__ __noinline__ inline Fq2 fq2_sqr_call(Fq2 a) { return CallOps::sqr(a); }
struct NestedCallOps {
    static DEV Fq mul(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
    static DEV Fq sqr(const Fq& a) { return fq_sqr_call(a); }
    static DEV Fq2 mul(const Fq2& a, const Fq2& b) { return fq2_mul_call(a, b); }
    static DEV Fq2 sqr(const Fq2& a) { return fq2_sqr_call(a); }
};
#endif

// acc += q  (q affine, optionally negated).  Handles every special case.
template <class F, class O = InlineOps>
HD void ec_madd(XYZZ<F>& acc, const Affine<F>& q_in, bool negate = false) {
    if (q_in.is_inf()) return;
    Affine<F> q = q_in;
    if (negate) q.y = fp_neg(q.y);
    if (acc.is_inf()) {
        acc = XYZZ<F>{q.x, q.y, F::one(), F::one()};
        return;
    }
    F u2 = O::mul(q.x, acc.zz);
    F s2 = O::mul(q.y, acc.zzz);
    F p = fp_sub(u2, acc.x);
    F r = fp_sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = ec_dbl_affine(q);
        else acc = XYZZ<F>::inf();
        return;
    }
    F pp = O::sqr(p);
    F ppp = O::mul(p, pp);
    F qq = O::mul(acc.x, pp);
    F x3 = fp_sub(fp_sub(O::sqr(r), ppp), fp_dbl(qq));
    F y3 = fp_sub(O::mul(r, fp_sub(qq, x3)), O::mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = O::mul(acc.zz, pp);
    acc.zzz = O::mul(acc.zzz, ppp);
}

// acc += b  (both XYZZ)
template <class F, class O = InlineOps>
HD void ec_add(XYZZ<F>& acc, const XYZZ<F>& b) {
    if (b.is_inf()) return;
    if (acc.is_inf()) {
        acc = b;
        return;
    }
    F u1 = O::mul(acc.x, b.zz);
    F u2 = O::mul(b.x, acc.zz);
    F s1 = O::mul(acc.y, b.zzz);
    F s2 = O::mul(b.y, acc.zzz);
    F p = fp_sub(u2, u1);
    F r = fp_sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) acc = ec_dbl(acc);
        else acc = XYZZ<F>::inf();
        return;
    }
    F pp = O::sqr(p);
    F ppp = O::mul(p, pp);
    F qq = O::mul(u1, pp);
    F x3 = fp_sub(fp_sub(O::sqr(r), ppp), fp_dbl(qq));
    F y3 = fp_sub(O::mul(r, fp_sub(qq, x3)), O::mul(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = O::mul(O::mul(acc.zz, b.zz), pp);
    acc.zzz = O::mul(O::mul(acc.zzz, b.zzz), ppp);
}

template <class F>
HD XYZZ<F> ec_neg(const XYZZ<F>& p) {
    XYZZ<F> r = p;
    r.y = fp_neg(p.y);
    return r;
}

// k * p for a small public multiplier (double-and-add, MSB first)
template <class F>
HD XYZZ<F> ec_mul_small(const XYZZ<F>& p, uint32_t k) {
    XYZZ<F> r = XYZZ<F>::inf();
    for (int bit = 31; bit >= 0; bit--) {
        r = ec_dbl(r);
        if ((k >> bit) & 1) ec_add(r, p);
    }
    return r;
}

// k * p for a canonical 256-bit little-endian scalar
template <class F>
HD XYZZ<F> ec_mul_scalar(const XYZZ<F>& p, const uint32_t* k, int nlimbs = 8) {
    XYZZ<F> r = XYZZ<F>::inf();
    for (int i = nlimbs - 1; i >= 0; i--)
        for (int bit = 31; bit >= 0; bit--) {
            r = ec_dbl(r);
            if ((k[i] >> bit) & 1) ec_add(r, p);
        }
    return r;
}

template <class F>
HD_NOINLINE Affine<F> ec_to_affine(const XYZZ<F>& p) {
    if (p.is_inf()) return Affine<F>::inf();
    // 1/ZZZ gives both: 1/ZZ = ZZZ^-1 ... use two inversions' worth via one: zi = (zz*zzz)^-1
    F zi = fp_inv(fp_mul(p.zz, p.zzz));
    F zz_inv = fp_mul(zi, p.zzz);
    F zzz_inv = fp_mul(zi, p.zz);
    return Affine<F>{fp_mul(p.x, zz_inv), fp_mul(p.y, zzz_inv)};
}

using G1Affine = Affine<Fq>;
using G2Affine = Affine<Fq2>;
using G1XYZZ = XYZZ<Fq>;
using G2XYZZ = XYZZ<Fq2>;

}  // namespace b200zk
