// EXPERIMENT (lever 1) -- Fq Montgomery product on the FP64 pipe.
//
// The bucket accumulation is bound by the 32-bit integer multiplier (IMAD.WIDE holds the FMA-heavy pipe four
// cycles per warp), while the FP64 pipe issues as many DFMA per second as the integer pipe issues IMAD and sits
// idle.  This header computes EXACTLY the same Montgomery product as fp_mul<FqCfg> (same R = 2^384, same bytes)
// with double-precision FMAs, after Emmart-Zheng-Weems: operands in 8 limbs of 48 bits held as doubles; for two
// limbs a, b < 2^48
//     hi = fma_rz(a, b, 2^100)                 -> mantissa field = floor(a*b / 2^48)   (ulp(2^100) = 2^48, truncation)
//     lo = fma_rz(a, b, (2^100 + 2^52) - hi)   -> mantissa field = a*b mod 2^48        (exact: the sum is < 2^53)
// so both halves of the 96-bit product drop out of the raw bit patterns and are accumulated with INTEGER adds
// (the exponent bits are cancelled by one compile-time constant per column and round).  Reduction is word-serial
// CIOS over the 48-bit limbs: q = t0 * (-p^-1) mod 2^48 with the same three operations, then t += q * p.
// Per product: 8 * (16 limb products * 3 + 5) = 424 FP64-pipe operations + 16 for the operand conversion,
// against 300 IMAD.WIDE; the price is ~550 integer add/shift instructions on the ALU pipe.
//
// Status: bit-exact against fp_mul on the host (tests/test_host.py::test_fq_mul_dfma, rounding mode set with
// fesetround) and timed on the GPU by b200zk_dbg_int_peak(kind 5).  Not used by any product kernel yet.
#pragma once
#include "field.cuh"
#if !defined(__CUDA_ARCH__)
#include <cfenv>
#include <cmath>
#include <cstring>
#endif

namespace b200zk {
namespace dfma {

constexpr int L = 8;  // 48-bit limbs of a 384-bit operand
constexpr uint64_t MASK48 = (1ull << 48) - 1;
constexpr uint64_t E52 = 0x4330000000000000ull;   // bit pattern of 2^52
constexpr uint64_t E100 = 0x4630000000000000ull;  // bit pattern of 2^100

HD double bits_to_double(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)x);
#else
    double d;
    memcpy(&d, &x, 8);
    return d;
#endif
}
HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t x;
    memcpy(&x, &d, 8);
    return x;
#endif
}
// fused multiply-add rounded toward zero.  Host: the caller holds FE_TOWARDZERO (RoundTowardZero below).
HD double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
#if !defined(__CUDA_ARCH__)
struct RoundTowardZero {
    int saved;
    RoundTowardZero() : saved(fegetround()) { fesetround(FE_TOWARDZERO); }
    ~RoundTowardZero() { fesetround(saved); }
};
#endif

// limb k of a 12 x u32 little-endian integer: bits [48k, 48k + 48)
HD uint64_t limb48(const uint32_t* w, int k) {
    const int m = (3 * k) >> 1;  // first 32-bit word
    if ((k & 1) == 0) return (uint64_t)w[m] | ((uint64_t)(w[m + 1] & 0xffffu) << 32);
    return (uint64_t)(w[m] >> 16) | ((uint64_t)w[m + 1] << 16);
}
HD constexpr uint64_t p48(int k) {  // 48-bit limbs of p
    const int m = (3 * k) >> 1;
    return (k & 1) == 0 ? ((uint64_t)FqCfg::mod(m) | ((uint64_t)(FqCfg::mod(m + 1) & 0xffffu) << 32))
                        : ((uint64_t)(FqCfg::mod(m) >> 16) | ((uint64_t)FqCfg::mod(m + 1) << 16));
}
HD constexpr uint64_t neg_p_inv48() {  // -p^-1 mod 2^48 by Newton iteration (p odd)
    uint64_t p0 = p48(0), x = 1;
    for (int i = 0; i < 6; i++) x = (x * (2 - p0 * x)) & MASK48;
    return (0 - x) & MASK48;
}
// an integer < 2^52 as a double, exactly (OR into the mantissa of 2^52, subtract 2^52)
HD double u52_to_double(uint64_t x) { return bits_to_double(E52 | x) - 4503599627370496.0; }

// raw bit patterns of the two halves of a*b (a, b integers < 2^48 as doubles):
//   hi_bits = E100 + floor(ab / 2^48),  lo_bits = E52 + (ab mod 2^48)
HD void mul_halves(double a, double b, uint64_t& hi_bits, uint64_t& lo_bits) {
    const double c1 = 1267650600228229401496703205376.0;                          // 2^100
    const double c2 = 1267650600228229401496703205376.0 + 4503599627370496.0;     // 2^100 + 2^52 (exact)
    const double hi = fma_rz(a, b, c1);
    const double lo = fma_rz(a, b, c2 - hi);
    hi_bits = double_to_bits(hi);
    lo_bits = double_to_bits(lo);
}

// a * b * 2^-384 mod p, fully reduced: the same bytes as fp_mul(a, b)
HD Fq fq_mul_dfma(const Fq& a, const Fq& b) {
#if !defined(__CUDA_ARCH__)
    RoundTowardZero guard;
#endif
    double ad[L], bd[L], pd[L];
#pragma unroll
    for (int k = 0; k < L; k++) {
        ad[k] = u52_to_double(limb48(a.v, k));
        bd[k] = u52_to_double(limb48(b.v, k));
        pd[k] = u52_to_double(p48(k));
    }
    const double ninv = u52_to_double(neg_p_inv48());
    // column accumulators (carry-save: up to 2^53 each, never normalised inside the loop)
    uint64_t col[L + 1];
#pragma unroll
    for (int k = 0; k <= L; k++) col[k] = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        // exponent bits of this round's terms: column j gets lo(a_i b_j), lo(q p_j), hi(a_i b_{j-1}), hi(q p_{j-1})
        col[0] -= 2 * E52;
#pragma unroll
        for (int j = 1; j < L; j++) col[j] -= 2 * E52 + 2 * E100;
        col[L] -= 2 * E100;
#pragma unroll
        for (int j = 0; j < L; j++) {
            uint64_t h, l;
            mul_halves(ad[i], bd[j], h, l);
            col[j] += l;
            col[j + 1] += h;
        }
        // q = (t mod 2^48) * (-p^-1) mod 2^48.  Column 0 still misses lo(q p_0), i.e. it is short by one E52:
        // its low 48 bits are already the true ones because E52 = 0 mod 2^48.
        const double t0 = u52_to_double(col[0] & MASK48);
        uint64_t qh, ql;
        mul_halves(t0, ninv, qh, ql);
        const double q = bits_to_double(ql) - 4503599627370496.0;  // ql = E52 + q: exact
#pragma unroll
        for (int j = 0; j < L; j++) {
            uint64_t h, l;
            mul_halves(q, pd[j], h, l);
            col[j] += l;
            col[j + 1] += h;
        }
        // column 0 is now 0 mod 2^48: pass its carry up and drop it
        col[1] += col[0] >> 48;
#pragma unroll
        for (int j = 0; j < L; j++) col[j] = col[j + 1];
        col[L] = 0;
    }
    // normalise to 48-bit limbs, repack as 12 x u32, one conditional subtraction (t < 2p)
    uint64_t carry = 0, lim[L];
#pragma unroll
    for (int k = 0; k < L; k++) {
        const uint64_t v = col[k] + carry;
        lim[k] = v & MASK48;
        carry = v >> 48;
    }
    Fq r;
#pragma unroll
    for (int k = 0; k < L; k += 2) {  // two limbs = 96 bits = three words
        const int m = (3 * k) >> 1;
        r.v[m] = (uint32_t)lim[k];
        r.v[m + 1] = (uint32_t)(lim[k] >> 32) | ((uint32_t)lim[k + 1] << 16);
        r.v[m + 2] = (uint32_t)(lim[k + 1] >> 16);
    }
    fp_reduce_once(r);
    return r;
}

}  // namespace dfma
}  // namespace b200zk
