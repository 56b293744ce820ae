// Fq in REDUCED RADIX: 14 limbs of 28 bits, Montgomery with R' = 2^392, products summed in 64-bit columns WITHOUT
// carry chains.
//
// Why a second representation of the same field.  Measured on B200 (tools/measure_peaks.py, profiles/r02_int_peaks.json):
//     IMAD.WIDE.U32   (32x32+64, no carry)                       17.6 T/s  = 60 / clk / SM, the rate of the 32-bit IMAD
//     IMAD.WIDE.U32.X (the same with carry in/out: the mad.lo.cc / madc.hi.cc rows of csrc/field_asm.cuh)  8.4 T/s
// The carry-chained form issues at HALF rate, and two independent product chains per thread do not help
// (30.3 G Fq products/s either way): the 32-bit-limb Montgomery product of field.cuh is bound by the issue rate of
// IMAD.WIDE.X, i.e. it can never use more than half of the multiplier.  With 28-bit limbs a column of up to 28 partial
// products (each < 2^58 even for operands that carry one unpropagated addition) fits a 64-bit accumulator, so every
// multiply-add is a plain IMAD.WIDE: 2 * 14^2 + 14 = 406 of them per product instead of 300 carry-chained ones, at
// twice the rate (406 * 2 = 812 pipe cycles against 300 * 4.1 = 1,230).  Carries are propagated once per product
// (shift / mask / add on the otherwise idle ALU pipe).
//
// Representation: value = sum l[i] * 2^(28 i).  "Tight" = every limb < 2^28; "loose" = limbs < 2^30 (a tight value plus or
// minus a couple of others, no carry propagated).  Values are residues mod p kept in [0, 2^385) -- the 11 spare bits of
// R' make Montgomery outputs contract: for inputs below 16 p the output is below 1.13 p without any conditional
// subtraction.  Montgomery form here is a * 2^392 mod p, NOT the a * 2^384 of field.cuh / arkworks: fq28_from_fq /
// fq28_to_fq convert (one product each), and nothing in this form ever crosses the C ABI.
//
// No reference counterpart (the reference has no field arithmetic).  Pinned on the host
// against Python integers (tests/test_host.py::test_fq28_arithmetic) and on the device through the MSM parity tests.
#pragma once
#include "field.cuh"

namespace b200zk {
namespace r28 {

constexpr int NL = 14, LB = 28;
constexpr uint32_t LM = (1u << LB) - 1;

struct Fq28 {
    uint32_t l[NL];
};

// bits [28 i, 28 i + 28) of the little-endian 32-bit word array w[12]
HD constexpr uint32_t limb_of(const uint32_t* w, int i) {
    const int bit = LB * i, k = bit >> 5, s = bit & 31;
    const uint64_t lo = k < 12 ? w[k] : 0, hi = k + 1 < 12 ? w[k + 1] : 0;
    return (uint32_t)(((lo | (hi << 32)) >> s) & LM);
}
HD constexpr uint32_t p28(int i) {
    constexpr uint32_t m[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return limb_of(m, i);
}
// -p^-1 mod 2^28 (Newton on the low word of p; p is odd)
HD constexpr uint32_t neg_inv28() {
    uint32_t p0 = 0xffffaaabu, x = 1;
    for (int i = 0; i < 6; i++) x = x * (2u - p0 * x);  // p0 * x == 1 mod 2^32
    return (0u - x) & LM;
}
constexpr uint32_t INV28 = neg_inv28();

// t += a * b, one IMAD.WIDE.U32 (spelled in PTX on the device so that ptxas sees exactly one multiply-add per column
// entry; the C form made nvcc add separate high-word corrections)
HD void mac(uint64_t& t, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a), "r"(b));
#else
    t += (uint64_t)a * b;
#endif
}

// Montgomery product a * b / 2^392 mod p.  Inputs loose (limbs < 2^30, value < 16 p); output tight, value < 1.13 p.
HD Fq28 mul(const Fq28& a, const Fq28& b) {
    uint64_t T[NL + 1];
#pragma unroll
    for (int j = 0; j <= NL; j++) T[j] = 0;
    uint32_t pl[NL];
#pragma unroll
    for (int j = 0; j < NL; j++) pl[j] = p28(j);
#pragma unroll
    for (int i = 0; i < NL; i++) {
        const uint32_t bi = b.l[i];
#pragma unroll
        for (int j = 0; j < NL; j++) mac(T[j], a.l[j], bi);
        const uint32_t m = ((uint32_t)T[0] * INV28) & LM;
#pragma unroll
        for (int j = 0; j < NL; j++) mac(T[j], m, pl[j]);
        const uint64_t carry = T[0] >> LB;  // the low 28 bits of T[0] are zero now
#pragma unroll
        for (int j = 0; j < NL; j++) T[j] = T[j + 1];
        T[NL] = 0;
        T[0] += carry;
    }
    Fq28 r;
#pragma unroll
    for (int j = 0; j < NL; j++) {
        r.l[j] = (uint32_t)T[j] & LM;
        if (j + 1 < NL) T[j + 1] += T[j] >> LB;
    }
    return r;
}

}  // namespace r28
}  // namespace b200zk
