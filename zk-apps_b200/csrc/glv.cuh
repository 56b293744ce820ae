// GLV decomposition for G1 and G2 of BLS12-381 (used by plain-bases MSMs, msm.cu).
//
// With z = -0xd201000000010000, lambda = z^2 - 1 satisfies lambda^2 + lambda + 1 = r EXACTLY (not only mod r),
// and phi(x, y) = (beta * x, y), beta a primitive cube root of unity in Fq, acts on G1 as multiplication by
// lambda (on G2, over the twist, the root with that eigenvalue is beta^2).  So the integer identity  k = k1 + k2 * lambda,  k2 = floor(k / lambda),  k1 = k mod lambda  gives
// k * P = k1 * P + k2 * phi(P)  with two NON-NEGATIVE half-length scalars (k1 < lambda < 2^128,
// k2 <= floor((2^256 - 1) / lambda) < 2^129): half the windows, half the bucket sets and half the doublings of
// the Horner step for twice the (cheap) bucket entries per window.  Constants derived and checked with the
// Python oracle (tests/test_host.py::test_glv_split); nothing comparable exists in the reference tree.
#pragma once
#include <cstdint>

#include "field.cuh"

namespace b200zk {

HD constexpr uint32_t glv_lambda(int i) {  // 0xac45a4010001a40200000000ffffffff
    constexpr uint32_t L[5] = {0xffffffffu, 0x00000000u, 0x0001a402u, 0xac45a401u, 0u};
    return L[i];
}
HD constexpr uint32_t glv_recip(int i) {   // floor(2^256 / lambda), 129 bits
    constexpr uint32_t M[5] = {0xf6cfee30u, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x00000001u};
    return M[i];
}
HD Fq glv_beta() {                  // beta * 2^384 mod p: phi(G) = lambda * G with this root
    Fq b = {{0x8671f071u, 0xcd03c9e4u, 0x1fcda5d2u, 0x5dab2246u, 0xd3851b95u, 0x587042afu, 0x01bacb9eu,
             0x8eb60ebeu, 0x83d050d2u, 0x03f97d6eu, 0x54638741u, 0x18f02065u}};
    return b;
}

HD Fq glv_beta_g2() {               // beta^2: on the twist E'(Fq2) it is (beta^2 x, y) that acts on G2 as lambda
    Fq b = {{0x798a64e8u, 0x30f1361bu, 0x7ece5a2au, 0xf3b8ddabu, 0xc61577f7u, 0x16a8ca3au, 0x74fd029bu,
             0xc26a2ff8u, 0x60701c6eu, 0x3636b766u, 0x241b6160u, 0x051ba4abu}};
    return b;
}
// x-coordinate of phi(P) = lambda * P
HD Fq glv_phi_x(const Fq& x) { return fp_mul(x, glv_beta()); }
HD Fq2 glv_phi_x(const Fq2& x) {
    const Fq b = glv_beta_g2();
    return Fq2{fp_mul(x.c0, b), fp_mul(x.c1, b)};
}

constexpr int GLV_LIMBS = 5;   // limbs of a half scalar
constexpr int GLV_BITS = 130;  // k1 < 2^128, k2 < 2^129, + 1 for the signed-digit carry

// k: any 256-bit integer, little-endian limbs.  k1, k2: GLV_LIMBS limbs each.
HD void glv_split(const uint32_t* k, uint32_t* k1, uint32_t* k2) {
    // qhat = floor(k * M / 2^256) with M = floor(2^256 / lambda):  q - 2 <= qhat <= q
    uint32_t prod[13];
#pragma unroll
    for (int i = 0; i < 13; i++) prod[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const uint64_t t = (uint64_t)k[i] * glv_recip(j) + prod[i + j] + carry;
            prod[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        prod[i + 5] = (uint32_t)carry;
    }
    uint32_t q[GLV_LIMBS];
#pragma unroll
    for (int i = 0; i < GLV_LIMBS; i++) q[i] = prod[8 + i];
    // d = k - q * lambda  (0 <= d < 3 * lambda, so 5 limbs are enough; computed mod 2^160)
    uint32_t t[GLV_LIMBS];
#pragma unroll
    for (int i = 0; i < GLV_LIMBS; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < GLV_LIMBS; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j + i < GLV_LIMBS; j++) {
            const uint64_t v = (uint64_t)q[i] * glv_lambda(j) + t[i + j] + carry;
            t[i + j] = (uint32_t)v;
            carry = v >> 32;
        }
    }
    uint32_t d[GLV_LIMBS];
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < GLV_LIMBS; i++) {
        const uint64_t v = (uint64_t)k[i] - t[i] - br;
        d[i] = (uint32_t)v;
        br = (v >> 63) & 1;
    }
    // at most two corrections
    for (int it = 0; it < 2; it++) {
        uint32_t e[GLV_LIMBS];
        uint64_t b2 = 0;
#pragma unroll
        for (int i = 0; i < GLV_LIMBS; i++) {
            const uint64_t v = (uint64_t)d[i] - glv_lambda(i) - b2;
            e[i] = (uint32_t)v;
            b2 = (v >> 63) & 1;
        }
        if (b2) break;  // d < lambda
        uint64_t c = 1;
#pragma unroll
        for (int i = 0; i < GLV_LIMBS; i++) {
            d[i] = e[i];
            c += q[i];
            q[i] = (uint32_t)c;
            c >>= 32;
        }
    }
#pragma unroll
    for (int i = 0; i < GLV_LIMBS; i++) {
        k1[i] = d[i];
        k2[i] = q[i];
    }
}

}  // namespace b200zk
