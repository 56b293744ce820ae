// K1 -- BLS12-381 Fr / Fq / Fq2 Montgomery arithmetic for sm_100a.
//
// No reference counterpart exists: /root/reference has no field arithmetic (SURVEY.md
// section 8 row a10; it lives in lockfile-only halo2curves / ark-ff,
// shielder/contract/Cargo.lock:224-225).  Representation = what ark-ff 0.4 stores:
// little-endian limbs of a*R mod p with R = 2^(32*N) (N = 8 for Fr, 12 for Fq), so the
// bit pattern equals arkworks' 64-bit-limb BigInt<4>/BigInt<6>.
//
// Device path: 32-bit mad.lo.cc / madc.hi.cc carry chains.  Each lo/hi pair on an
// aligned register pair is one IMAD.WIDE.U32(.X) in SASS.  The running value is kept
// split in two arrays X (register pairs at even limb positions) and Y (pairs at odd
// positions), V = X + 2^32 * Y, so every 32x32 product lands on an aligned pair and no
// carry chain is ever broken.  One b-limb per iteration: add a*b_i, add m*p with
// m = X[0] * (-p^-1), shift right by one limb -- the shift swaps the roles of the two
// arrays (pure register renaming); the single limb that changes parity (old X[1]) is
// folded into new X[0] with one add.cc whose carry enters the next Y chain.
// Host path (also used by the host-side constant generation): portable CIOS in C++.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
// out-of-line on purpose (code size / ptxas time); `inline` only for linkage
#define HD_NOINLINE __host__ __device__ __noinline__ inline
#else
#define HD inline
#define DEV inline
#define HD_NOINLINE inline
#endif

namespace b200zk {

// ------------------------------------------------------------------ field parameters
struct FrCfg {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;  // -r^-1 mod 2^32
    HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return m[i];
    }
    HD static constexpr uint32_t one(int i) {  // R mod r
        constexpr uint32_t m[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return m[i];
    }
    HD static constexpr uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t m[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return m[i];
    }
};

struct FqCfg {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;  // -p^-1 mod 2^32
    HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu,
                                    0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u,
                                    0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return m[i];
    }
    HD static constexpr uint32_t one(int i) {  // R mod p
        constexpr uint32_t m[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu,
                                    0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u,
                                    0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return m[i];
    }
    HD static constexpr uint32_t p2(int i) {  // p^2, 24 limbs (keeps lazy Fq2 differences non-negative)
        constexpr uint32_t m[24] = {0x1c718e39u, 0x26aa0000u, 0x76382eabu, 0x7ced6b1du, 0x62113cfdu, 0x162c3383u, 0x3e71b743u, 0x66bf91edu, 0x7091a049u, 0x292e85a8u, 0x86185c7bu, 0x1d68619cu, 0x0978ef01u, 0xf5314933u, 0x16ddca6eu, 0x50a62cfdu, 0x349e8bd0u, 0x66e59e49u, 0x0e7046b4u, 0xe2dc90e5u, 0xa22f25e9u, 0x4bd278eau, 0xb8c35fc7u, 0x02a437a4u};
        return m[i];
    }
    HD static constexpr uint32_t r2(int i) {  // R^2 mod p
        constexpr uint32_t m[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u,
                                    0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u,
                                    0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return m[i];
    }
};

// ------------------------------------------------------------------ PTX carry chains (generated)
}  // namespace b200zk
#include "field_asm.cuh"
namespace b200zk {

// ------------------------------------------------------------------ generic prime field element
template <class C>
struct alignas(16) Fp {
    static constexpr int N = C::N;
    using Cfg = C;
    uint32_t v[N];

    HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = C::one(i);
        return r;
    }
    HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = C::r2(i);
        return r;
    }
    HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= v[i];
        return o == 0;
    }
    HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = a - p if a >= p else a   (a < 2p)
template <class C>
HD void fp_reduce_once(Fp<C>& a) {
    constexpr int N = C::N;
    uint32_t t[N];
#if defined(__CUDA_ARCH__)
    uint32_t pm[N];
#pragma unroll
    for (int i = 0; i < N; i++) pm[i] = C::mod(i);
    uint32_t borrow = ptx::sub_n<N>(t, a.v, pm);  // all-ones if a < p
#pragma unroll
    for (int i = 0; i < N; i++) a.v[i] = borrow ? a.v[i] : t[i];
#else
    uint64_t br = 0;
    for (int i = 0; i < N; i++) {
        uint64_t d = (uint64_t)a.v[i] - C::mod(i) - br;
        t[i] = (uint32_t)d;
        br = (d >> 63) & 1;
    }
    if (!br)
        for (int i = 0; i < N; i++) a.v[i] = t[i];
#endif
}

template <class C>
HD Fp<C> fp_add(const Fp<C>& a, const Fp<C>& b) {
    constexpr int N = C::N;
    Fp<C> r;
#if defined(__CUDA_ARCH__)
    ptx::add_n<N>(r.v, a.v, b.v);  // 2p < 2^(32N): no carry out
#else
    uint64_t c = 0;
    for (int i = 0; i < N; i++) {
        uint64_t s = (uint64_t)a.v[i] + b.v[i] + c;
        r.v[i] = (uint32_t)s;
        c = s >> 32;
    }
#endif
    fp_reduce_once(r);
    return r;
}

template <class C>
HD Fp<C> fp_sub(const Fp<C>& a, const Fp<C>& b) {
    constexpr int N = C::N;
    Fp<C> r;
#if defined(__CUDA_ARCH__)
    uint32_t t[N], pm[N];
    uint32_t borrow = ptx::sub_n<N>(t, a.v, b.v);  // all-ones if a < b
#pragma unroll
    for (int i = 0; i < N; i++) pm[i] = C::mod(i) & borrow;
    ptx::add_n<N>(r.v, t, pm);
#else
    uint64_t br = 0;
    for (int i = 0; i < N; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        r.v[i] = (uint32_t)d;
        br = (d >> 63) & 1;
    }
    if (br) {
        uint64_t c = 0;
        for (int i = 0; i < N; i++) {
            uint64_t s = (uint64_t)r.v[i] + C::mod(i) + c;
            r.v[i] = (uint32_t)s;
            c = s >> 32;
        }
    }
#endif
    return r;
}

template <class C>
HD Fp<C> fp_neg(const Fp<C>& a) {
    return a.is_zero() ? a : fp_sub(Fp<C>::zero(), a);
}

template <class C>
HD Fp<C> fp_dbl(const Fp<C>& a) { return fp_add(a, a); }

// Montgomery product a*b/R mod p, fully reduced.  Inputs < p.
template <class C>
HD Fp<C> fp_mul(const Fp<C>& a, const Fp<C>& b) {
    constexpr int N = C::N;
    Fp<C> r;
#if defined(__CUDA_ARCH__)
    // V = X + 2^32 * Y.  X: limbs at positions 0..N, Y: limbs at positions 1..N.
    uint32_t X[N + 1], Y[N], pm[N];
#pragma unroll
    for (int k = 0; k <= N; k++) X[k] = 0;
#pragma unroll
    for (int k = 0; k < N; k++) Y[k] = 0;
#pragma unroll
    for (int k = 0; k < N; k++) pm[k] = C::mod(k);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint32_t bi = b.v[i];
        if (i == 0) {
            ptx::row_odd<N>(Y, a.v, bi);
        } else {
            // shift right one limb: X' = Y (+ old X[1] at position 0), Y' = X >> 64;
            // the carry of Y[0] + X[1] enters the Y' chain (position 1) inside the same asm block
            uint32_t nx[N + 1], ny[N];
#pragma unroll
            for (int k = 1; k < N; k++) nx[k] = Y[k];
            nx[N] = 0;
#pragma unroll
            for (int k = 0; k < N - 1; k++) ny[k] = X[k + 2];
            ny[N - 1] = 0;
            ptx::row_odd_cin<N>(nx[0], Y[0], X[1], ny, a.v, bi);
#pragma unroll
            for (int k = 0; k <= N; k++) X[k] = nx[k];
#pragma unroll
            for (int k = 0; k < N; k++) Y[k] = ny[k];
        }
        ptx::row_even<N>(X, a.v, bi);
        // m = X[0] * (-p^-1) mod 2^32;  V += m * p  (makes X[0] == 0)
        const uint32_t m = X[0] * C::INV;
        ptx::row_even<N>(X, pm, m);
        ptx::row_odd<N>(Y, pm, m);
    }
    // V / 2^32 = Y + (X >> 32)
    ptx::add_n<N>(r.v, Y, X + 1);
#else
    uint32_t t[N + 2];
    for (int k = 0; k < N + 2; k++) t[k] = 0;
    for (int i = 0; i < N; i++) {
        uint64_t c = 0;
        for (int j = 0; j < N; j++) {
            uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + c;
            t[j] = (uint32_t)s;
            c = s >> 32;
        }
        uint64_t s = (uint64_t)t[N] + c;
        t[N] = (uint32_t)s;
        t[N + 1] = (uint32_t)(s >> 32);
        uint32_t m = t[0] * C::INV;
        s = (uint64_t)m * C::mod(0) + t[0];
        c = s >> 32;
        for (int j = 1; j < N; j++) {
            s = (uint64_t)m * C::mod(j) + t[j] + c;
            t[j - 1] = (uint32_t)s;
            c = s >> 32;
        }
        s = (uint64_t)t[N] + c;
        t[N - 1] = (uint32_t)s;
        t[N] = t[N + 1] + (uint32_t)(s >> 32);
    }
    for (int k = 0; k < N; k++) r.v[k] = t[k];
#endif
    fp_reduce_once(r);
    return r;
}

#if defined(__CUDA_ARCH__)
// Montgomery reduction of a 2N-limb value T < p * 2^(32N) (word by word, same X/Y split as fp_mul):
// T / 2^(32N) mod p, fully reduced.  The upper limbs of T enter the sliding window one per step.
template <class C>
DEV Fp<C> fp_redc(const uint32_t* T) {
    constexpr int N = C::N;
    Fp<C> r;
    uint32_t X[N + 1], Y[N], pm[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        X[k] = T[k];
        Y[k] = 0;
        pm[k] = C::mod(k);
    }
    X[N] = 0;
    {
        const uint32_t m = X[0] * C::INV;
        ptx::row_even<N>(X, pm, m);
        ptx::row_odd<N>(Y, pm, m);
    }
#pragma unroll
    for (int i = 1; i < N; i++) {
        uint32_t nx[N + 1], ny[N], m;
#pragma unroll
        for (int k = 1; k < N - 1; k++) nx[k] = Y[k];
#pragma unroll
        for (int k = 0; k < N - 1; k++) ny[k] = X[k + 2];
        ny[N - 1] = 0;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=&r"(nx[N - 1]), "=&r"(nx[N]) : "r"(Y[N - 1]), "r"(T[N - 1 + i]));
        ptx::redc_step<N>(nx[0], m, Y[0], X[1], ny, pm, C::INV);
#pragma unroll
        for (int k = 0; k <= N; k++) X[k] = nx[k];
#pragma unroll
        for (int k = 0; k < N; k++) Y[k] = ny[k];
        ptx::row_even<N>(X, pm, m);
    }
    ptx::add_n<N>(r.v, Y, X + 1);
    r.v[N - 1] += T[2 * N - 1];
    fp_reduce_once(r);
    return r;
}
#endif

// a^2: on the device the N(N+1)/2 distinct limb products are formed once (cross terms doubled), then one
// Montgomery reduction: 78 + 156 wide multiply-adds for Fq instead of 300.
template <class C>
HD Fp<C> fp_sqr(const Fp<C>& a) {
#if defined(__CUDA_ARCH__)
    uint32_t T[2 * C::N];
    ptx::sqr_wide<C::N>(T, a.v);
    return fp_redc<C>(T);
#else
    return fp_mul(a, a);
#endif
}

template <class C>
HD Fp<C> fp_to_mont(const Fp<C>& a) { return fp_mul(a, Fp<C>::r2()); }

template <class C>
HD Fp<C> fp_from_mont(const Fp<C>& a) {
    Fp<C> o = Fp<C>::zero();
    o.v[0] = 1;
    return fp_mul(a, o);
}

// a^e for a little-endian limb exponent (not constant time; exponents here are public).
template <class C>
HD_NOINLINE Fp<C> fp_pow(const Fp<C>& a, const uint32_t* e, int nlimbs) {
    Fp<C> r = Fp<C>::one();
    bool started = false;
    for (int i = nlimbs - 1; i >= 0; i--) {
        for (int bit = 31; bit >= 0; bit--) {
            if (started) r = fp_sqr(r);
            if ((e[i] >> bit) & 1) {
                r = started ? fp_mul(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}

// a < b on the canonical (non-Montgomery) integer values given as limbs
template <int N>
HD bool limbs_gt(const uint32_t* a, const uint32_t* b) {
    for (int i = N - 1; i >= 0; i--) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return false;
}

// a^(p-2); returns 0 for a == 0.  Kept as the independent check of fp_inv (tests) -- ~570 products.
template <class C>
HD Fp<C> fp_inv_fermat(const Fp<C>& a) {
    uint32_t e[C::N];
    uint32_t borrow = 2;
    for (int i = 0; i < C::N; i++) {  // p - 2
        uint32_t m = C::mod(i);
        e[i] = m - borrow;
        borrow = (m < borrow) ? 1 : 0;
    }
    return fp_pow(a, e, C::N);
}

// fp_inv = the safegcd inversion below (fp_inv_safegcd: a quarter of the instructions of the bit-by-bit Euclid, so the
// to-affine tails of every MSM, the proof assembly and the final exponentiation wait a quarter as long); the binary
// Euclid is kept as fp_inv_binary, the independent implementation the host tests compare it with.
template <class C> HD Fp<C> fp_inv_safegcd(const Fp<C>& a);
template <class C> HD Fp<C> fp_inv(const Fp<C>& a) { return fp_inv_safegcd(a); }

// Inversion by the binary extended Euclidean algorithm (Guide to ECC, Alg. 2.22): shifts, additions and
// subtractions only -- roughly a tenth of the instructions of a^(p-2) and none of them on the multiply pipe,
// which matters because every inversion here sits on a latency-bound tail (to-affine at the end of an MSM,
// proof assembly, final exponentiation).  Not constant time: the operands are public (curve points, pairing
// values).  Invariants: x1 * a = u, x2 * a = v (mod p) on the plain integers; for the Montgomery word aR this
// yields (aR)^-1 = a^-1 R^-1, and two products by R^2 return a^-1 R.  Returns 0 for a == 0.
template <class C>
HD Fp<C> fp_inv_binary(const Fp<C>& a) {
    constexpr int N = C::N;
    uint32_t u[N], v[N];
    Fp<C> x1 = Fp<C>::zero(), x2 = Fp<C>::zero();
    x1.v[0] = 1;
#pragma unroll
    for (int i = 0; i < N; i++) {
        u[i] = a.v[i];
        v[i] = C::mod(i);
    }
    // Total on every N-limb word: a caller-supplied, unreduced word (u >= p; u == p in particular) would make
    // the halving loop below spin on u = 0, so reduce first (at most 2^(32N) / p < 10 subtractions) and map
    // every multiple of p to 0 like a == 0.
    for (;;) {
        uint32_t t[N];
        uint64_t br = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint64_t d = (uint64_t)u[i] - v[i] - br;
            t[i] = (uint32_t)d;
            br = (d >> 63) & 1;
        }
        if (br) break;  // u < p
#pragma unroll
        for (int i = 0; i < N; i++) u[i] = t[i];
    }
    {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= u[i];
        if (o == 0) return Fp<C>::zero();
    }
    auto is_one = [](const uint32_t* w) {
        uint32_t o = w[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; i++) o |= w[i];
        return o == 0;
    };
    auto shr1 = [](uint32_t* w, uint32_t top) {  // w = (top:w) >> 1
#pragma unroll
        for (int i = 0; i < N - 1; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
        w[N - 1] = (w[N - 1] >> 1) | (top << 31);
    };
    auto halve = [&](Fp<C>& x) {  // x / 2 mod p
        uint32_t carry = 0;
        if (x.v[0] & 1u) {  // x + p < 2^(32N) for both fields; keep the carry anyway
            uint64_t c = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                c += (uint64_t)x.v[i] + C::mod(i);
                x.v[i] = (uint32_t)c;
                c >>= 32;
            }
            carry = (uint32_t)c;
        }
        shr1(x.v, carry);
    };
    auto sub_to = [](uint32_t* t, const uint32_t* w, const uint32_t* z) {  // t = w - z, returns the borrow
        uint64_t br = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint64_t d = (uint64_t)w[i] - z[i] - br;
            t[i] = (uint32_t)d;
            br = (d >> 63) & 1;
        }
        return (uint32_t)br;
    };
    while (!is_one(u) && !is_one(v)) {
        while (!(u[0] & 1u)) {
            shr1(u, 0);
            halve(x1);
        }
        while (!(v[0] & 1u)) {
            shr1(v, 0);
            halve(x2);
        }
        uint32_t t[N];
        if (sub_to(t, u, v)) {  // u < v
            sub_to(v, v, u);
            x2 = fp_sub(x2, x1);
        } else {
#pragma unroll
            for (int i = 0; i < N; i++) u[i] = t[i];
            x1 = fp_sub(x1, x2);
        }
    }
    const Fp<C> r = is_one(u) ? x1 : x2;
    return fp_mul(fp_mul(r, Fp<C>::r2()), Fp<C>::r2());
}

// GROUNDWORK (not used by a product kernel yet): the same inversion with ONE instruction stream for every input --
// each step is  "if u is odd: (swap so that u >= v), u -= v, x1 -= x2;  then u /= 2, x1 /= 2"  done with selects, for a
// fixed 2 * bits(p) steps (every step removes at least one bit from u or v; v stays odd; at the end u = 0, v = 1 and
// x2 * a = v).  In a warp the data-dependent branches of fp_inv serialise (measured: one inversion costs about as much
// as 30 affine additions); this form trades ~2x the instructions per lane for no divergence.
// Pinned against fp_inv on the host (tests/test_host.py::test_host_field_ops, op 8).
template <class C>
HD Fp<C> fp_inv_uniform(const Fp<C>& a) {
    constexpr int N = C::N;
    uint32_t u[N], v[N];
    Fp<C> x1 = Fp<C>::zero(), x2 = Fp<C>::zero();
    x1.v[0] = 1;
    int bits = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        u[i] = a.v[i];
        v[i] = C::mod(i);
    }
    {  // bit length of p
        int top = N - 1;
        while (top > 0 && C::mod(top) == 0) top--;
        uint32_t w = C::mod(top);
        bits = 32 * top;
        while (w) {
            bits++;
            w >>= 1;
        }
    }
    for (int step = 0; step < 2 * bits; step++) {
        const uint32_t odd = 0u - (u[0] & 1u);  // all-ones if u is odd
        // t = u - v, borrow <=> u < v
        uint32_t t[N];
        uint64_t br = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint64_t d = (uint64_t)u[i] - v[i] - br;
            t[i] = (uint32_t)d;
            br = (d >> 63) & 1;
        }
        const uint32_t lt = 0u - (uint32_t)br;
        const uint32_t swap = odd & lt;      // u odd and u < v: (u, v) = (v - u... handled as v' = u, u' = v - u)
        // new u = odd ? |u - v| : u ;  new v = swap ? u : v
        uint64_t c = swap & 1u;              // two's complement negate of t when swapping
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t ti = t[i] ^ swap;
            c += ti;
            const uint32_t neg_or_t = (uint32_t)c;  // swap ? -(u - v) = v - u : u - v
            c >>= 32;
            const uint32_t old_u = u[i];
            u[i] = (odd & neg_or_t) | (~odd & old_u);
            v[i] = (swap & old_u) | (~swap & v[i]);
        }
        // x: swap first, then x1 -= x2 when u was odd
        Fp<C> nx1, nx2;
#pragma unroll
        for (int i = 0; i < N; i++) {
            nx1.v[i] = (swap & x2.v[i]) | (~swap & x1.v[i]);
            nx2.v[i] = (swap & x1.v[i]) | (~swap & x2.v[i]);
        }
        const Fp<C> diff = fp_sub(nx1, nx2);
#pragma unroll
        for (int i = 0; i < N; i++) x1.v[i] = (odd & diff.v[i]) | (~odd & nx1.v[i]);
        x2 = nx2;
        // halve u (now even) and x1 (mod p)
#pragma unroll
        for (int i = 0; i < N - 1; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
        u[N - 1] >>= 1;
        const uint32_t xodd = 0u - (x1.v[0] & 1u);
        uint64_t cc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            cc += (uint64_t)x1.v[i] + (C::mod(i) & xodd);
            x1.v[i] = (uint32_t)cc;
            cc >>= 32;
        }
#pragma unroll
        for (int i = 0; i < N - 1; i++) x1.v[i] = (x1.v[i] >> 1) | (x1.v[i + 1] << 31);
        x1.v[N - 1] = (x1.v[N - 1] >> 1) | ((uint32_t)cc << 31);
    }
    // a == 0: u stays 0, v = p, x2 = 0 -> returns 0 like fp_inv
    return fp_mul(fp_mul(x2, Fp<C>::r2()), Fp<C>::r2());
}

// Inversion by "safegcd" division steps (Bernstein-Yang 2019) on 30-bit signed limbs, 30 steps at a time: the steps of
// one batch run on the low words of f and g only and yield a 2x2 transition matrix t with |entries| <= 2^30; then
// [f, g] <- t [f, g] / 2^30 (exact) and [d, e] <- t [d, e] / 2^30 (mod p) are 4 + 6 multiply-accumulate rows over the
// limbs.  ~25 batches for a 381-bit modulus, ~18 k instructions with no long carry chain -- a quarter of the bit-by-bit
// Euclid of fp_inv and far less dependent -- which is what makes one inversion per warp cheap enough for the batched-
// affine additions (msm_affine.cuh).  Variable time (counts trailing zeros, stops when g = 0): operands are public, and
// the caller runs it on warp-uniform data, so nothing diverges.  Invariants: d * a = f, e * a = g (mod p); at the end
// g = 0, f = +-1, so +-d = a^-1.  Total on every N-limb word like fp_inv (reduces first; multiples of p give 0).
// Pinned against pow(x, -1, p) on the host (tests/test_host.py::test_host_field_ops, op 9) and against fp_inv.
template <class C>
HD Fp<C> fp_inv_safegcd(const Fp<C>& a) {
    constexpr int N = C::N;
    constexpr int L = (32 * N + 2 + 29) / 30;  // 30-bit limbs: 13 for Fq, 9 for Fr (room for the range (-2p, p))
    constexpr int32_t M30 = 0x3fffffff;
    uint32_t w[N];
#pragma unroll
    for (int i = 0; i < N; i++) w[i] = a.v[i];
    for (;;) {  // canonicalise: at most 2^(32N) / p < 10 subtractions
        uint32_t t[N];
        uint64_t br = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint64_t dd = (uint64_t)w[i] - C::mod(i) - br;
            t[i] = (uint32_t)dd;
            br = (dd >> 63) & 1;
        }
        if (br) break;
#pragma unroll
        for (int i = 0; i < N; i++) w[i] = t[i];
    }
    int32_t f[L], g[L], d[L], e[L], pm[L];
#pragma unroll
    for (int i = 0; i < L; i++) {
        const int bit = 30 * i, limb = bit >> 5, sh = bit & 31;
        uint64_t vg = limb < N ? w[limb] : 0u, vp = limb < N ? C::mod(limb) : 0u;
        if (limb + 1 < N) {
            vg |= (uint64_t)w[limb + 1] << 32;
            vp |= (uint64_t)C::mod(limb + 1) << 32;
        }
        g[i] = (int32_t)((vg >> sh) & M30);
        pm[i] = (int32_t)((vp >> sh) & M30);
        f[i] = pm[i];
        d[i] = 0;
        e[i] = 0;
    }
    e[0] = 1;
    uint32_t pinv = C::mod(0);  // p^-1 mod 2^32 by Newton (p odd: p * p = 1 mod 8)
#pragma unroll
    for (int i = 0; i < 4; i++) pinv *= 2u - C::mod(0) * pinv;
    int32_t eta = -1;
    for (int it = 0; it < 48; it++) {  // ends on g = 0: <= ceil((49 * 384 + 57) / 17 / 30) = 38 batches
        {
            uint32_t o = 0;
#pragma unroll
            for (int i = 0; i < L; i++) o |= (uint32_t)g[i];
            if (o == 0) break;
        }
        // ---- 30 division steps on the low words -> t = [u v; q r]
        uint32_t u = 1, v = 0, q = 0, r = 1, ff = (uint32_t)f[0], gg = (uint32_t)g[0];
        int i = 30;
        for (;;) {
            const uint32_t sentinel = gg | (0xffffffffu << i);
#if defined(__CUDA_ARCH__)
            const int zeros = __ffs((int)sentinel) - 1;
#else
            const int zeros = __builtin_ctz(sentinel);
#endif
            gg >>= zeros;
            u <<= zeros;
            v <<= zeros;
            eta -= zeros;
            i -= zeros;
            if (i == 0) break;
            uint32_t wm;
            if (eta < 0) {  // swap: (f, g) <- (g, -f); cancel up to 6 low bits of g
                eta = -eta;
                uint32_t tmp = ff; ff = gg; gg = 0u - tmp;
                tmp = u; u = q; q = 0u - tmp;
                tmp = v; v = r; r = 0u - tmp;
                const int limit = (eta + 1) > i ? i : (eta + 1);
                const uint32_t m = (0xffffffffu >> (32 - limit)) & 63u;
                wm = (ff * gg * (ff * ff - 2u)) & m;      // -g / f mod 2^6
            } else {        // cancel up to 4 low bits
                const int limit = (eta + 1) > i ? i : (eta + 1);
                const uint32_t m = (0xffffffffu >> (32 - limit)) & 15u;
                wm = ff + (((ff + 1u) & 4u) << 1);         // 1 / f mod 2^4
                wm = ((0u - wm) * gg) & m;
            }
            gg += ff * wm;
            q += u * wm;
            r += v * wm;
        }
        const int32_t tu = (int32_t)u, tv = (int32_t)v, tq = (int32_t)q, tr = (int32_t)r;
        // ---- [d, e] <- t [d, e] / 2^30 mod p, kept in (-2p, p)
        {
            const int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
            int32_t md = (tu & sd) + (tv & se), me = (tq & sd) + (tr & se);
            int64_t cd = (int64_t)tu * d[0] + (int64_t)tv * e[0];
            int64_t ce = (int64_t)tq * d[0] + (int64_t)tr * e[0];
            md -= (int32_t)((pinv * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
            me -= (int32_t)((pinv * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
            cd += (int64_t)pm[0] * md;
            ce += (int64_t)pm[0] * me;
            cd >>= 30;
            ce >>= 30;
#pragma unroll
            for (int k = 1; k < L; k++) {
                const int32_t dk = d[k], ek = e[k];
                cd += (int64_t)tu * dk + (int64_t)tv * ek + (int64_t)pm[k] * md;
                ce += (int64_t)tq * dk + (int64_t)tr * ek + (int64_t)pm[k] * me;
                d[k - 1] = (int32_t)cd & M30;
                e[k - 1] = (int32_t)ce & M30;
                cd >>= 30;
                ce >>= 30;
            }
            d[L - 1] = (int32_t)cd;
            e[L - 1] = (int32_t)ce;
        }
        // ---- [f, g] <- t [f, g] / 2^30 (exact)
        {
            int64_t cf = (int64_t)tu * f[0] + (int64_t)tv * g[0];
            int64_t cg = (int64_t)tq * f[0] + (int64_t)tr * g[0];
            cf >>= 30;
            cg >>= 30;
#pragma unroll
            for (int k = 1; k < L; k++) {
                const int32_t fk = f[k], gk = g[k];
                cf += (int64_t)tu * fk + (int64_t)tv * gk;
                cg += (int64_t)tq * fk + (int64_t)tr * gk;
                f[k - 1] = (int32_t)cf & M30;
                g[k - 1] = (int32_t)cg & M30;
                cf >>= 30;
                cg >>= 30;
            }
            f[L - 1] = (int32_t)cf;
            g[L - 1] = (int32_t)cg;
        }
    }
    // ---- normalise d to [0, p), negated when f = -1
    {
        const int32_t sign = f[L - 1] >> 31;
        int32_t cond_add = d[L - 1] >> 31;
#pragma unroll
        for (int k = 0; k < L; k++) {
            d[k] += pm[k] & cond_add;
            d[k] = (d[k] ^ sign) - sign;
        }
#pragma unroll
        for (int k = 0; k < L - 1; k++) {
            d[k + 1] += d[k] >> 30;
            d[k] &= M30;
        }
        cond_add = d[L - 1] >> 31;
#pragma unroll
        for (int k = 0; k < L; k++) d[k] += pm[k] & cond_add;
#pragma unroll
        for (int k = 0; k < L - 1; k++) {
            d[k + 1] += d[k] >> 30;
            d[k] &= M30;
        }
    }
    Fp<C> res;
#pragma unroll
    for (int i = 0; i < N; i++) {  // 30-bit limbs -> 32-bit words
        const int bit = 32 * i, k = bit / 30, sh = bit % 30;
        uint64_t acc = (uint64_t)(uint32_t)d[k] >> sh;
        if (k + 1 < L) acc |= (uint64_t)(uint32_t)d[k + 1] << (30 - sh);
        if (k + 2 < L) acc |= (uint64_t)(uint32_t)d[k + 2] << (60 - sh);
        res.v[i] = (uint32_t)acc;
    }
    // (aR)^-1 = a^-1 R^-1; two Montgomery products by R^2 give a^-1 R
    return fp_mul(fp_mul(res, Fp<C>::r2()), Fp<C>::r2());
}

using Fr = Fp<FrCfg>;
using Fq = Fp<FqCfg>;

// ------------------------------------------------------------------ Fq2 = Fq[u]/(u^2 + 1)
struct alignas(16) Fq2 {
    Fq c0, c1;
    HD static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    HD static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    HD bool operator!=(const Fq2& b) const { return !(*this == b); }
};

HD Fq2 fp_add(const Fq2& a, const Fq2& b) { return Fq2{fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
HD Fq2 fp_sub(const Fq2& a, const Fq2& b) { return Fq2{fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
HD Fq2 fp_neg(const Fq2& a) { return Fq2{fp_neg(a.c0), fp_neg(a.c1)}; }
HD Fq2 fp_dbl(const Fq2& a) { return Fq2{fp_dbl(a.c0), fp_dbl(a.c1)}; }
// Fq2 products are out of line: a G2 accumulator does not fit the register file anyway, the
// operands travel through L1-resident local memory (~2% of the product's issue time), and the
// code shrinks ~10x (I-cache, ptxas time).
// B200ZK_FQ2_LAZY (off): measured on B200 the lazy-reduction product below (744 instead of 900 wide
// multiply-adds, bit-exact, tests green) makes the G2 bucket accumulation 7 % SLOWER (13.3 -> 14.2 ms per step):
// the 24-limb add/sub carry chains are pure latency at 8 warps/SM and the bigger callee frame costs the caller
// more spills.  Kept for a kernel with more resident warps.
HD_NOINLINE Fq2 fp_mul(const Fq2& a, const Fq2& b) {  // Karatsuba, 3 Fq mul
#if defined(__CUDA_ARCH__) && defined(B200ZK_FQ2_LAZY)
    // lazy reduction: three unreduced 24-limb products, the Karatsuba sums on the wide values, two Montgomery
    // reductions instead of three (3*144 + 2*156 = 744 wide multiply-adds instead of 900).
    // c1 = (a0+a1)(b0+b1) - a0 b0 - a1 b1 in [0, 2p^2);  c0 = a0 b0 - a1 b1 + p^2 in (0, 2p^2);  2p^2 < p * 2^384.
    constexpr int N = FqCfg::N;
    uint32_t t0[2 * N], t1[2 * N], t2[2 * N], sa[N], sb[N], pp[2 * N];
    ptx::add_n<N>(sa, a.c0.v, a.c1.v);  // < 2p < 2^382: no reduction needed
    ptx::add_n<N>(sb, b.c0.v, b.c1.v);
    ptx::mul_wide<N>(t2, sa, sb);
    ptx::mul_wide<N>(t0, a.c0.v, b.c0.v);
    ptx::mul_wide<N>(t1, a.c1.v, b.c1.v);
    ptx::sub_wide<N>(t2, t2, t0);
    ptx::sub_wide<N>(t2, t2, t1);
#pragma unroll
    for (int k = 0; k < 2 * N; k++) pp[k] = FqCfg::p2(k);
    ptx::add_wide<N>(t0, t0, pp);
    ptx::sub_wide<N>(t0, t0, t1);
    return Fq2{fp_redc<FqCfg>(t0), fp_redc<FqCfg>(t2)};
#else
    Fq t0 = fp_mul(a.c0, b.c0);
    Fq t1 = fp_mul(a.c1, b.c1);
    Fq t2 = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
    return Fq2{fp_sub(t0, t1), fp_sub(fp_sub(t2, t0), t1)};
#endif
}
HD_NOINLINE Fq2 fp_sqr(const Fq2& a) {  // (a0+a1)(a0-a1) + 2 a0 a1 u
    Fq t0 = fp_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1));
    Fq t1 = fp_mul(a.c0, a.c1);
    return Fq2{t0, fp_dbl(t1)};
}
HD Fq2 fp_inv(const Fq2& a) {
    Fq d = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
    return Fq2{fp_mul(a.c0, d), fp_neg(fp_mul(a.c1, d))};
}

}  // namespace b200zk
