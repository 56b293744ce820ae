// Groth16 verification, cut into the pieces the batched verifier kernels (verify.cu) run one thread each
// (SURVEY.md section 8f rank 1).  Host + device: tests/host/hostcheck.cpp runs the same code on the CPU and
// tests/test_oracle_verify.py diffs its verdicts against oracle/pyref/groth16.py verify_with_vk.
//
// No reference counterpart: the contract calls a SHA-256 mock where a verifier would sit
// (shielder/contract/lib.rs:56,74 -> shielder/mocked_zk/src/relations.rs:127-155).  The equation is the one
// arkworks' Groth16::verify_proof_with_prepared_inputs evaluates [recall]:
//     e(A, B) * e(L, -gamma) * e(C, -delta) == e(alpha, beta),   L = gamma_abc[0] + sum_i x_i gamma_abc[i + 1]
// with e(alpha, beta) and the negated G2 elements prepared once per key (arkworks' PreparedVerifyingKey).
#pragma once
#include "pairing.cuh"

namespace b200zk {

// per-proof verdicts (include/b200zk.h: b200zk_proof_status)
enum {
    PROOF_ACCEPTED = 0,
    PROOF_REJECTED = 1,          // well-formed, pairing equation does not hold
    PROOF_BAD_ENCODING = 2,      // flag bits / x >= p
    PROOF_NOT_ON_CURVE = 3,
    PROOF_NOT_IN_SUBGROUP = 4,
    PROOF_BAD_INPUT = 5,         // a public input is not a reduced Fr element
};

struct alignas(16) PreparedVk {
    Affine<Fq> alpha_g1;
    Affine<Fq2> beta_g2, gamma_g2_neg, delta_g2_neg;
    Fq12 alpha_beta;             // final_exponentiation(miller_loop(alpha, beta))
};

template <class F>
HD Affine<F> affine_neg(const Affine<F>& p) { return Affine<F>{p.x, fp_neg(p.y)}; }

HD int point_status_to_proof(int st) {
    return st == POINT_OK ? PROOF_ACCEPTED
           : st == POINT_BAD_ENCODING ? PROOF_BAD_ENCODING
           : st == POINT_NOT_ON_CURVE ? PROOF_NOT_ON_CURVE : PROOF_NOT_IN_SUBGROUP;
}

// element `which` (0 = A, 1 = B, 2 = C) of a 192-byte compressed proof A(48) | B(96) | C(48)
HD_NOINLINE int decode_proof_g1(const uint8_t* proof, int which, bool check_subgroup, Affine<Fq>& out) {
    const int st = g1_decompress(proof + (which == 0 ? 0 : 144), out);
    if (st != POINT_OK) return point_status_to_proof(st);
    if (check_subgroup && !ec_in_subgroup(out)) return PROOF_NOT_IN_SUBGROUP;
    return PROOF_ACCEPTED;
}
HD_NOINLINE int decode_proof_g2(const uint8_t* proof, bool check_subgroup, Affine<Fq2>& out) {
    const int st = g2_decompress(proof + 48, out);
    if (st != POINT_OK) return point_status_to_proof(st);
    if (check_subgroup && !ec_in_subgroup(out)) return PROOF_NOT_IN_SUBGROUP;
    return PROOF_ACCEPTED;
}

// x * base for a Montgomery-form public input; false if x is not reduced
HD_NOINLINE bool input_term(const Affine<Fq>& base, const Fr& x_mont, XYZZ<Fq>& out) {
    uint32_t m[8];
    for (int i = 0; i < 8; i++) m[i] = FrCfg::mod(i);
    if (!limbs_gt<8>(m, x_mont.v)) return false;
    const Fr x = fp_from_mont(x_mont);
    out = ec_mul_scalar(XYZZ<Fq>::from_affine(base), x.v);
    return true;
}

// the three Miller loops of one proof; `pair` selects which one this thread evaluates
HD_NOINLINE Fq12 verify_miller(const PreparedVk& vk, int pair, const Affine<Fq>& a, const Affine<Fq2>& b,
                               const Affine<Fq>& l, const Affine<Fq>& c) {
    if (pair == 0) return miller_loop(a, b);
    if (pair == 1) return miller_loop(l, vk.gamma_g2_neg);
    return miller_loop(c, vk.delta_g2_neg);
}

HD_NOINLINE bool verify_final(const PreparedVk& vk, const Fq12& f0, const Fq12& f1, const Fq12& f2) {
    return final_exponentiation(fq12_mul(fq12_mul(f0, f1), f2)) == vk.alpha_beta;
}

HD_NOINLINE void prepare_vk(const Affine<Fq>& alpha_g1, const Affine<Fq2>& beta_g2, const Affine<Fq2>& gamma_g2,
                            const Affine<Fq2>& delta_g2, PreparedVk& out) {
    out.alpha_g1 = alpha_g1;
    out.beta_g2 = beta_g2;
    out.gamma_g2_neg = affine_neg(gamma_g2);
    out.delta_g2_neg = affine_neg(delta_g2);
    out.alpha_beta = final_exponentiation(miller_loop(alpha_g1, beta_g2));
}

// The whole check for one proof, sequentially (host path / small batches): the composition the kernels split up.
// public_inputs: num_public Montgomery Fr (the leading 1 is implicit), gamma_abc: num_public + 1 points.
HD_NOINLINE int verify_one(const PreparedVk& vk, const Affine<Fq>* gamma_abc, uint32_t num_public, const uint8_t* proof,
                           const Fr* public_inputs, bool check_subgroup) {
    Affine<Fq> a, c;
    Affine<Fq2> b;
    int st = decode_proof_g1(proof, 0, check_subgroup, a);
    if (st != PROOF_ACCEPTED) return st;
    st = decode_proof_g2(proof, check_subgroup, b);
    if (st != PROOF_ACCEPTED) return st;
    st = decode_proof_g1(proof, 2, check_subgroup, c);
    if (st != PROOF_ACCEPTED) return st;
    XYZZ<Fq> acc = XYZZ<Fq>::from_affine(gamma_abc[0]);
    for (uint32_t i = 0; i < num_public; i++) {
        XYZZ<Fq> t;
        if (!input_term(gamma_abc[i + 1], public_inputs[i], t)) return PROOF_BAD_INPUT;
        ec_add(acc, t);
    }
    const Affine<Fq> l = ec_to_affine(acc);
    const Fq12 f0 = verify_miller(vk, 0, a, b, l, c), f1 = verify_miller(vk, 1, a, b, l, c),
               f2 = verify_miller(vk, 2, a, b, l, c);
    return verify_final(vk, f0, f1, f2) ? PROOF_ACCEPTED : PROOF_REJECTED;
}

}  // namespace b200zk
