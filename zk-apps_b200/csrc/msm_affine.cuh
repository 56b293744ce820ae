// K4/K5 -- batched-affine pairwise summation ("affine levels") for MSMs over full digit tables.
//
// With the digit tables resident (msm.cu, precompute level 2) an MSM is a plain sum of affine table rows, one per
// non-zero signed digit.  The XYZZ running sum of msm_accumulate pays 8M + 2S per row.  Here the rows are summed
// PAIRWISE, level by level -- level 0 adds rows (2j, 2j+1) of the entry list, level 1 adds the results pairwise, ... --
// so every addition of a level is independent of every other and all of them can share field inversions
// (Montgomery's trick): affine + affine = 5M + 1S.
//
//   forward   d_i = x_Q - x_P of the lane's i-th pair;  pre[i] = d_0 d_1 ... d_i              (1 M per pair)
//   warp      prefix / suffix product scans over the 32 lane totals with shuffles (12 M per lane), ONE inversion of
//             the warp total -- every lane runs it on identical data, so the branchy binary Euclid of fp_inv does not
//             diverge -- and each lane leaves with 1 / (its own total)
//   backward  1/d_i = inv * pre[i-1];  inv *= d_i;  lambda = (y_Q - y_P) / d_i;
//             x3 = lambda^2 - x_P - x_Q;  y3 = lambda (x_P - x3) - y_P                        (4 M + 1 S per pair)
//
// The inversion (shifts and subtractions, ALU pipe only) and the scans are amortised over 32 * B additions and leave
// the multiplier -- the pipe that bounds everything else on this path -- to the other warps.
// K levels take 1 - 2^-K of the additions; what is left (total / 2^K points) goes through msm_accumulate unchanged.
// Proofs never mix: the entry list of every proof is padded to a multiple of 2^K with "infinity" entries
// (table_pad_offsets / table_entries in msm.cu), so pairs never straddle a proof boundary at any level.
//
// Special cases are resolved per pair without breaking the product chain (they contribute the factor 1, or 2y for a
// doubling): inf + Q, P + inf, P + (-P) = inf, P + P.  They are rare (padding, repeated bases) and sit on a divergent
// slow path.  Replaces nothing in /root/reference (VariableBaseMSM is absent there);
// the result is a group element, so the affine bytes are those of any other correct schedule.
#pragma once
#include "ec.cuh"

namespace b200zk {

constexpr uint32_t AFF_PAD_ENTRY = 0xffffffffu;  // entry-list padding: the point at infinity (index field all ones)
constexpr uint32_t AFF_B_MIN = 8, AFF_B_MAX = 1024;  // additions per lane that share one inversion per warp

#if defined(__CUDACC__)

__device__ __noinline__ inline Fq fq_inv_call(Fq a) { return fp_inv_safegcd(a); }
template <class F> DEV F aff_inv(const F& a);
template <> DEV Fq aff_inv<Fq>(const Fq& a) { return fq_inv_call(a); }
template <> DEV Fq2 aff_inv<Fq2>(const Fq2& a) {
    const Fq d = fq_inv_call(fp_add(CallOps::sqr(a.c0), CallOps::sqr(a.c1)));
    return Fq2{CallOps::mul(a.c0, d), fp_neg(CallOps::mul(a.c1, d))};
}

template <class F>
DEV F aff_shfl(const F& a, int src_lane) {
    F r;
    constexpr int W = sizeof(F) / 4;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&a);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < W; i++) d[i] = __shfl_sync(0xffffffffu, s[i], src_lane);
    return r;
}

// every lane enters with t != 0 and leaves with 1 / t: two scans, one inversion per warp
template <class F>
DEV F aff_warp_invert(const F& t, uint32_t lane) {
    F pre = t, suf = t;
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) {
        const F y = aff_shfl(pre, ((int)lane - o) & 31);  // out-of-range lanes read some lane, result unused
        const F z = aff_shfl(suf, ((int)lane + o) & 31);
        const F py = CallOps::mul(pre, y), sz = CallOps::mul(suf, z);
        if ((int)lane >= o) pre = py;
        if ((int)lane + o < 32) suf = sz;
    }
    const F total_inv = aff_inv(aff_shfl(pre, 31));
    F left = aff_shfl(pre, ((int)lane - 1) & 31), right = aff_shfl(suf, ((int)lane + 1) & 31);
    if (lane == 0) left = F::one();
    if (lane == 31) right = F::one();
    return CallOps::mul(CallOps::mul(total_inv, left), right);
}

enum : uint32_t { AFF_ADD = 0, AFF_DBL = 1, AFF_COPY_P = 2, AFF_COPY_Q = 3, AFF_INF = 4 };

// full classification of a pair whose fast test failed (x_P = 0, x_Q = 0 or x_P = x_Q); d = the pair's factor
template <class F>
__device__ __noinline__ uint32_t aff_classify_slow(const Affine<F>& p, const Affine<F>& q, F& d) {
    d = F::one();
    if (p.is_inf()) return q.is_inf() ? AFF_INF : AFF_COPY_Q;
    if (q.is_inf()) return AFF_COPY_P;
    if (p.x == q.x) {
        if (p.y == q.y && !p.y.is_zero()) {
            d = fp_dbl(p.y);
            return AFF_DBL;
        }
        return AFF_INF;  // opposite points
    }
    d = fp_sub(q.x, p.x);
    return AFF_ADD;
}

// Shared memory of one CTA of msm_affine_level, in bytes (dynamic: the G2 build needs more than 48 KB)
template <class F>
constexpr size_t aff_smem_bytes(bool first) {
    constexpr size_t Q = sizeof(Affine<F>) / 16;
    return (2 * Q + Q / 2) * 16 * 128 + (first ? 8 * 8 * 128 : 0);
}

// One level: out[j] = in[2j] + in[2j+1] for j < n_pairs = (total[0] >> shift) / 2; total[1 + shift] = B of the level.  FIRST: in[k] = +-table[entries[k]]
// (bit 31 = negate, AFF_PAD_ENTRY = infinity).  Every lane does up to B additions; consecutive lanes touch consecutive
// pairs, so the loads / stores of the upper levels and of the entry list are contiguous per warp.
//
// Everything the two loops consume arrives through cp.async (LDGSTS) into shared memory, issued iterations ahead:
// the table rows, the entries that address them, and -- on the way back -- the prefix products the forward pass left
// in `pre_g` ([row][quad][lane], global scratch).  A plain load would be waited for at the next CALL (the field products
// are calls, ec.cuh), i.e. its whole latency exposed once per pair; an asynchronous copy is only waited for when its
// data is read.  One commit group per iteration:
//   forward   G_i = { x_P, x_Q of pair i + 2;  entries of pair i + 4 },        wait_group 1 at the top of iteration i
//   backward  G_i = { P, Q of pair i - 1;  pre[i - 2];  entries of pair i - 3 }, wait_group 0
template <class F, bool FIRST, int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) msm_affine_level(const Affine<F>* __restrict__ in,
                                                                    const Affine<F>* __restrict__ in2, uint32_t n_split,
                                                                    const uint32_t* __restrict__ entries,
                                                                    const uint32_t* __restrict__ total, uint32_t shift,
                                                                    Affine<F>* __restrict__ out, F* __restrict__ pre_g) {
    constexpr int Q = sizeof(Affine<F>) / 16;  // 16-byte quads per point
    constexpr int QX = Q / 2;                  // ... per coordinate
    // staging, [quad][thread] so that a warp's LDS.128 / LDGSTS.128 are conflict-free.  Forward pass: two slots of
    // (x_P, x_Q) = 4 QX quads; backward pass: one slot of (P, Q, pre) = 2 Q + QX quads.  Then the entry ring.
    extern __shared__ uint4 aff_smem[];
    uint4* my = aff_smem + threadIdx.x;
    uint2* ering = reinterpret_cast<uint2*>(aff_smem + (2 * Q + QX) * 128) + threadIdx.x;  // [slot & 7][thread]

    const uint32_t n_pairs = (total[0] >> shift) >> 1;
    const uint32_t B = min(total[1 + shift], AFF_B_MAX);  // plan of this level (table_pad_offsets)
    const uint32_t warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    // W warps share the level; in its i-th iteration warp w takes the 32 pairs of row i W + w, so one iteration of
    // all warps together sweeps one contiguous stretch of the arrays (entries, upper-level points, prefix products,
    // results) instead of W streams a warp's whole share apart
    const uint32_t W = (n_pairs + 32u * B - 1) / (32u * B);
    if (warp >= W) return;  // whole warp
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rows = lane < n_pairs ? (n_pairs - lane + 31) / 32 : 0;  // rows r with pair 32 r + lane in range
    const int m = rows > warp ? (int)min(B, (rows - warp + W - 1) / W) : 0;
    auto pair_of = [&](int i) { return ((uint32_t)i * W + warp) * 32u + lane; };
    // prefix product i of this lane: quad q at pre_q[((i W + w) QX + q) 32 + lane] -- a warp's 16-byte stores coalesce
    uint4* pre_q = reinterpret_cast<uint4*>(pre_g) + lane;
    auto pre_quad = [&](int i, int q) { return pre_q + (((size_t)i * W + warp) * (sizeof(F) / 16) + q) * 32; };

    auto cp16 = [](uint4* dst, const uint4* src) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
    };
    auto e_fetch = [&](int i) {
        if constexpr (FIRST) {
            if (i >= 0 && i < m)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(ering + (i & 7) * 128)),
                             "l"(reinterpret_cast<const uint2*>(entries) + pair_of(i)));
        }
    };
    auto e_get = [&](int i) -> uint2 {
        if constexpr (FIRST) return ering[(i & 7) * 128];
        else return make_uint2(0, 0);
    };
    auto point_ptr = [&](int i, int which, uint32_t e) -> const Affine<F>* {
        if constexpr (FIRST) {  // entries below n_split index `in`, the others `in2` (the phi half of a GLV base table)
            const uint32_t idx = e & 0x7fffffffu;
            return idx < n_split ? in + idx : in2 + (idx - n_split);
        } else return in + 2 * (size_t)pair_of(i) + which;
    };
    // cp.async nq quads of the point into staging quads dst0...; a padding entry stages zeros
    auto stage_point = [&](const Affine<F>* p, bool pad, int nq, int dst0) {
        const uint4* src = reinterpret_cast<const uint4*>(p);
        for (int q = 0; q < nq; q++) {
            uint4* d = my + (dst0 + q) * 128;
            if (FIRST && pad) *d = make_uint4(0, 0, 0, 0);
            else cp16(d, src + q);
        }
    };
    auto read_coord = [&](int quad0) {
        F r;
        uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
        for (int q = 0; q < QX; q++) d[q] = my[(quad0 + q) * 128];
        return r;
    };
    auto load_full = [&](int i, int which, uint32_t e) {  // slow path: straight from global memory, sign applied
        Affine<F> r = Affine<F>::inf();
        if (FIRST && e == AFF_PAD_ENTRY) return r;
        const uint4* src = reinterpret_cast<const uint4*>(point_ptr(i, which, e));
        uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
        for (int q = 0; q < Q; q++) d[q] = __ldg(src + q);
        if (FIRST && (e >> 31)) r.y = fp_neg(r.y);
        return r;
    };
    auto commit = [] { asm volatile("cp.async.commit_group;"); };

    // ---------------------------------------------------------------- forward: running product of the denominators
    F acc = F::one();
    {
        auto issue_x = [&](int i) {  // x_P, x_Q of pair i -> slot i & 1 (entries of pair i already in the ring)
            if (i < m) {
                const uint2 e = e_get(i);
                const int slot = (i & 1) * 2 * QX;
                stage_point(point_ptr(i, 0, e.x), e.x == AFF_PAD_ENTRY, QX, slot);
                stage_point(point_ptr(i, 1, e.y), e.y == AFF_PAD_ENTRY, QX, slot + QX);
            }
        };
        e_fetch(0); e_fetch(1); e_fetch(2); e_fetch(3);
        commit();
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        issue_x(0);
        commit();
        issue_x(1);
        commit();
#pragma unroll 1
        for (int i = 0; i < m; i++) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            const int slot = (i & 1) * 2 * QX;
            const F xp = read_coord(slot), xq = read_coord(slot + QX);
            const uint2 ecur = e_get(i);
            issue_x(i + 2);
            e_fetch(i + 4);
            commit();
            F d = fp_sub(xq, xp);
            if (xp.is_zero() || xq.is_zero() || d.is_zero()) {
                const Affine<F> p = load_full(i, 0, ecur.x), q = load_full(i, 1, ecur.y);
                F ds;  // the callee takes references: keep what it touches local to this branch, or `d` lives on the stack
                aff_classify_slow(p, q, ds);
                d = ds;
            }
            acc = CallOps::mul(acc, d);
            const uint4* src = reinterpret_cast<const uint4*>(&acc);
#pragma unroll
            for (int qq = 0; qq < QX; qq++) *pre_quad(i, qq) = src[qq];
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // ---------------------------------------------------------------- one inversion per warp
    F inv = aff_warp_invert(acc, lane);
    // ---------------------------------------------------------------- backward: the additions
    {
        auto issue_full = [&](int i) {  // P, Q of pair i and pre[i - 1] -> the slot
            if (i >= 0) {
                const uint2 e = e_get(i);
                stage_point(point_ptr(i, 0, e.x), e.x == AFF_PAD_ENTRY, Q, 0);
                stage_point(point_ptr(i, 1, e.y), e.y == AFF_PAD_ENTRY, Q, Q);
                if (i > 0) {
#pragma unroll
                    for (int qq = 0; qq < QX; qq++) cp16(my + (2 * Q + qq) * 128, pre_quad(i - 1, qq));
                }
            }
        };
        e_fetch(m - 1); e_fetch(m - 2);
        commit();
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        issue_full(m - 1);
        e_fetch(m - 3);
        commit();
#pragma unroll 1
        for (int i = m - 1; i >= 0; i--) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            Affine<F> p, q;
            p.x = read_coord(0);
            p.y = read_coord(QX);
            q.x = read_coord(Q);
            q.y = read_coord(Q + QX);
            const F pre_prev = read_coord(2 * Q);  // garbage for i = 0, unused
            const uint2 ecur = e_get(i);
            if constexpr (FIRST) {  // padding stays (0, 0): -0 = 0
                if (ecur.x >> 31) p.y = fp_neg(p.y);
                if (ecur.y >> 31) q.y = fp_neg(q.y);
            }
            issue_full(i - 1);
            e_fetch(i - 3);
            commit();
            F d = fp_sub(q.x, p.x);
            F num = fp_sub(q.y, p.y);
            uint32_t kind = AFF_ADD;
            if (p.x.is_zero() || q.x.is_zero() || d.is_zero()) {
                const Affine<F> pc = p, qc = q;  // copies for the by-reference callee: p, q, d stay in registers on the fast path
                F ds;
                kind = aff_classify_slow(pc, qc, ds);
                d = ds;
                if (kind == AFF_DBL) {
                    const F xx = CallOps::sqr(p.x);
                    num = fp_add(fp_dbl(xx), xx);
                }
            }
            const F dinv = i ? CallOps::mul(inv, pre_prev) : inv;
            inv = CallOps::mul(inv, d);
            Affine<F> r;
            if (kind <= AFF_DBL) {
                const F lam = CallOps::mul(num, dinv);
                r.x = fp_sub(fp_sub(CallOps::sqr(lam), p.x), q.x);
                r.y = fp_sub(CallOps::mul(lam, fp_sub(p.x, r.x)), p.y);
            } else {
                r = kind == AFF_COPY_P ? p : (kind == AFF_COPY_Q ? q : Affine<F>::inf());
            }
            uint4* dst = reinterpret_cast<uint4*>(out + pair_of(i));
            const uint4* src = reinterpret_cast<const uint4*>(&r);
#pragma unroll
            for (int qq = 0; qq < Q; qq++) dst[qq] = src[qq];
        }
    }
}

#endif  // __CUDACC__

}  // namespace b200zk
