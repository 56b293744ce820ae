// K4/K5 -- bucketed (Pippenger) multi-scalar multiplication over G1 (Fq) and G2 (Fq2).
//
// Replaces ark_ec::VariableBaseMSM::msm_bigint (ark-ec 0.4.2 [recall]); absent from
// /root/reference.  The result is a group element, so any
// correct schedule yields the same affine bytes as arkworks' (App. B).
//
// Pipeline (all kernels below are ours; no CUB / thrust):
//   1. msm_count:   signed c-bit digit decomposition of every scalar (digits in
//                   [-2^(c-1), 2^(c-1)], so only 2^(c-1) buckets per window), histogram of
//                   bucket keys  key = (msm_in_batch * windows + window) * 2^(c-1) + |d|-1
//   2. scan_*:      exclusive prefix sum of the histogram (bucket start offsets)
//   3. msm_scatter: counting-sort scatter of (point index | sign) into bucket order
//   4. msm_accumulate: one thread per fixed run of 2^log_tl sorted entries (independent of bucket
//                   boundaries, so all lanes do equal work), mixed additions (XYZZ += affine,
//                   8M+2S), 128-bit loads, next point prefetched while adding; msm_fold_* sums the
//                   partials of buckets that span several runs
//   5. msm_chunk_reduce + msm_sum: weighted bucket sum  sum_j (j+1) B_j  per window by running
//                   sums over 32-bucket chunks (+ one small scalar multiple per chunk), then a
//                   log-depth tree of plain sums
//   6. msm_finish:  Horner over windows (c doublings each) and conversion to affine.
// With `precompute` bases (proving-key queries are fixed) the table holds 2^(c*w) * P_i for every
// window, all windows share one bucket set, and step 6 has nothing to combine.
// A batch of MSMs over the same bases (one per proof) is a single pass of this pipeline: the
// msm index is just the high part of the bucket key.
#include <algorithm>
#include <cstring>
#include <type_traits>

#include "glv.cuh"
#include "msm_affine.cuh"
#include "types.cuh"

using namespace b200zk;

namespace {

struct Plan {
    uint32_t c, windows, nb;      // nb = buckets per window = 2^(c-1)
    uint32_t weff;                // bucket sets per msm: windows (of this part), or 1 when precomputed
    bool precomp;
    uint32_t w0, w1;              // windows [w0, w1) this pass of the pipeline handles (all of them unless split)
    bool glv;                     // plain bases: every scalar is k1 + k2*lambda, entries (k1, P_i) and (k2, phi(P_i))
    bool table = false;           // precompute level 2: full digit table, "bucket" = the whole MSM of one proof
    bool tglv = false;            // ... over GLV half scalars: the `batch` of a pass counts (proof, half) pairs, half 0 = k1
                                  // (sum over the table rows as they are), half 1 = k2 (the same rows, phi applied to the sum)
    uint32_t mult = 0;            // table entries per (window, base) = 2^(c-1)
    size_t n_table = 0;           // bases per window in a precomputed table (= bases of the handle, >= scalars of a call)
};

// Window size from an operation-count model (Fq multiplications): n*W mixed additions (10 each)
// plus ~4 full additions (14 each) per bucket for the reduction; without precomputed window
// multiples every window has its own bucket set.
uint32_t pick_window(size_t n, bool precomp, bool glv) {
    if (const char* e = getenv(precomp ? "B200ZK_MSM_C_PRE" : "B200ZK_MSM_C")) {
        int v = atoi(e);
        if (v >= 2 && v <= 22) return (uint32_t)v;
    }
    // GLV half scalars over plain bases: measured (profiles/r02_msm_window_sweep.log, G1 2^16...2^26, G2 2^18).  The
    // operation-count model below does not see that the top window of a 128-bit half scalar is nearly empty at c = 16
    // (bit 128 up) and badly skewed at c = 17, 18 (a few hundred huge buckets), nor what the affine levels save.
    if (glv && !precomp) return n < (1u << 21) ? 15u : n < (1u << 25) ? 16u : 19u;
    const uint32_t bits = glv ? GLV_BITS : 256;
    double best = 1e300;
    uint32_t best_c = 3;
    for (uint32_t c = 3; c <= 20; c++) {
        const double W = (bits + c - 1) / c;
        const double madds = (double)n * (glv ? 2.0 : 1.0) * W * 10.0, per_set = (double)(1u << (c - 1)) * 56.0;
        const double cost = madds + (precomp ? per_set : per_set * W);
        if (cost < best) {
            best = cost;
            best_c = c;
        }
    }
    return best_c;
}

Plan make_plan(size_t n, bool precomp, uint32_t c_fixed, bool glv = false, bool table_glv = false) {
    Plan p;
    p.glv = glv && !precomp;
    p.tglv = table_glv;
    p.c = c_fixed ? c_fixed : pick_window(n, precomp, p.glv);
    p.windows = (((p.glv || p.tglv) ? GLV_BITS : 256) + p.c - 1) / p.c;
    p.nb = 1u << (p.c - 1);
    p.precomp = precomp;
    p.weff = precomp ? 1 : p.windows;
    p.w0 = 0;
    p.w1 = p.windows;
    p.n_table = n;
    return p;
}

// signed digit `w` of canonical scalar k; carry threads through successive calls (w ascending)
__device__ __forceinline__ int32_t next_digit(const uint32_t* k, uint32_t w, uint32_t c, uint32_t& carry) {
    const uint32_t bit = w * c;
    uint32_t raw = 0;
    if (bit < 256) {
        const uint32_t limb = bit >> 5, sh = bit & 31;
        raw = k[limb] >> sh;
        if (sh + c > 32 && limb + 1 < 8) raw |= k[limb + 1] << (32 - sh);
        raw &= (1u << c) - 1;
    }
    raw += carry;
    if (raw > (1u << (c - 1))) {
        carry = 1;
        return (int32_t)raw - (int32_t)(1u << c);
    }
    carry = 0;
    return (int32_t)raw;
}

__device__ __forceinline__ void load_scalar(const uint32_t* scalars, size_t idx, bool mont, uint32_t* k) {
    const uint4* q = reinterpret_cast<const uint4*>(scalars + idx * 8);
    uint4 a = q[0], b = q[1];
    Fr s;
    s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.v[3] = a.w;
    s.v[4] = b.x; s.v[5] = b.y; s.v[6] = b.z; s.v[7] = b.w;
    if (mont) s = fp_from_mont(s);
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = s.v[i];
}

// Calls f(key, entry) for every non-zero signed digit of scalar (b, i) that falls into the windows of this pass.
// k: 8 limbs (upper ones zero for a GLV half scalar); pidx0: index of the base in window 0's table.
template <class Fn>
__device__ __forceinline__ void for_each_digit(const uint32_t* k, const Plan& pl, uint32_t b, size_t pidx0, size_t n, Fn f) {
    uint32_t carry = 0;
    for (uint32_t w = 0; w < pl.windows; w++) {
        int32_t d = next_digit(k, w, pl.c, carry);
        if (d == 0 || w < pl.w0 || w >= pl.w1) continue;
        uint32_t bucket = (uint32_t)(d < 0 ? -d : d) - 1;
        uint32_t key = (b * pl.weff + (pl.precomp ? 0 : w - pl.w0)) * pl.nb + bucket;
        uint32_t pidx = (uint32_t)(pl.precomp ? (size_t)w * pl.n_table + pidx0 : pidx0);
        f(key, pidx | (d < 0 ? 0x80000000u : 0u));
    }
}

template <class Fn>
__device__ __forceinline__ void for_each_entry(const uint32_t* scalars, size_t n, size_t stride, size_t batch, int mont,
                                               const Plan& pl, const uint8_t* __restrict__ skip, Fn f) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    const size_t b = t / n, i = t % n;
    if (skip && skip[i]) return;
    uint32_t k[8];
    load_scalar(scalars, b * stride + i, mont != 0, k);
    if (pl.glv) {
        uint32_t k1[8], k2[8];
#pragma unroll
        for (int j = GLV_LIMBS; j < 8; j++) k1[j] = k2[j] = 0;
        glv_split(k, k1, k2);
        for_each_digit(k1, pl, (uint32_t)b, i, n, f);
        for_each_digit(k2, pl, (uint32_t)b, n + i, n, f);   // phi(P_i) lives at index n + i
    } else {
        for_each_digit(k, pl, (uint32_t)b, i, n, f);
    }
}

// scalars: batch rows of n scalars, row stride `stride` elements
__global__ void msm_count(const uint32_t* scalars, size_t n, size_t stride, size_t batch, int mont, Plan pl,
                          const uint8_t* __restrict__ skip, uint32_t* counts) {
    for_each_entry(scalars, n, stride, batch, mont, pl, skip, [&](uint32_t key, uint32_t) { atomicAdd(&counts[key], 1u); });
}

__global__ void msm_scatter(const uint32_t* scalars, size_t n, size_t stride, size_t batch, int mont, Plan pl,
                            const uint8_t* __restrict__ skip, uint32_t* cursor, uint32_t* sorted) {
    for_each_entry(scalars, n, stride, batch, mont, pl, skip,
                   [&](uint32_t key, uint32_t entry) { sorted[atomicAdd(&cursor[key], 1u)] = entry; });
}

// second half of a GLV base table: phi(P) = (beta * x, y)   (beta^2 on G2)
template <class F>
__global__ void glv_phi_kernel(const Affine<F>* __restrict__ in, size_t n, Affine<F>* __restrict__ out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = in[i];
    p.x = glv_phi_x(p.x);
    out[i] = p;
}

// ---------------------------------------------------------------- exclusive scan (3 kernels)
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    uint32_t prefix = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return prefix + x - v;
}

__global__ void scan_block_sums(const uint32_t* in, size_t n, uint32_t* block_sums) {
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) s += in[base + i];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of m values, running carry across chunks
__global__ void scan_single(uint32_t* data, size_t m) {
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (size_t base = 0; base < m; base += SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < m ? data[i] : 0;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total);
        uint32_t c = carry_s;
        if (i < m) data[i] = ex + c;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = c + total;
        __syncthreads();
    }
}

// out[i] = exclusive prefix of in; out[n] = total (written by the last block)
__global__ void scan_apply(const uint32_t* in, size_t n, const uint32_t* block_offsets, uint32_t* out) {
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total) + block_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) out[n] = ex;
}

// ---------------------------------------------------------------- point I/O
template <class F>
__device__ __forceinline__ Affine<F> load_affine(const Affine<F>* p) {
    Affine<F> r;
    constexpr int Q = sizeof(Affine<F>) / 16;
    const uint4* src = reinterpret_cast<const uint4*>(p);
    uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < Q; i++) dst[i] = __ldg(src + i);
    return r;
}

// ---------------------------------------------------------------- bucket accumulation
// The bucket-sorted entry array is cut into fixed runs of L = 2^log_tl entries, one thread per run,
// whatever the bucket boundaries are: every lane of a warp executes exactly L mixed additions
// (ncu on the one-thread-per-bucket version: 16-23 of 32 lanes active, because bucket sizes are
// Poisson distributed and the top window / 0-1-heavy witnesses create huge buckets).  When a run
// crosses a bucket boundary the running sum is flushed.  A bucket covered by a single run is
// written straight to `buckets`; a bucket split over several runs gets one partial per run
// (slot = seg_off[bucket] + run - first_run_of_bucket) folded by msm_fold_small / msm_fold_big.
constexpr uint32_t FOLD_SMALL_MAX = 8;

// The run length is fixed ON THE DEVICE from the real entry count (the host only knows an upper bound and never
// waits for the count): every thread does the same number of mixed additions, so a launch is a sequence of waves of
// equal-duration CTAs and a partially filled last wave is pure loss (15.0 M entries in 64-entry runs = 4.4 waves of
// 444 resident CTAs: 12 % of the launch idle).  L is chosen so that the runs fill a whole number of waves: the
// nearest whole number of waves at the host's guess L0, then L = ceil(total / (lanes per wave * waves)), kept within
// [L0 / 2, 2 L0].
struct RunPlan {
    uint32_t L;       // entries per run
    uint32_t n_runs;  // ceil(total / L)
};

__global__ void msm_plan_runs(const uint32_t* __restrict__ offsets, uint32_t n_keys, uint32_t L0, uint32_t lanes_per_wave,
                              RunPlan* __restrict__ plan) {
    const uint32_t total = offsets[n_keys];
    uint32_t L = L0;
    if (total && lanes_per_wave) {
        const uint64_t per_wave = (uint64_t)lanes_per_wave * L0;
        uint64_t waves = (total + per_wave / 2) / per_wave;
        if (waves == 0) waves = 1;
        const uint64_t lanes = lanes_per_wave * waves;
        L = (uint32_t)((total + lanes - 1) / lanes);
        const uint32_t lo = L0 > 1 ? L0 / 2 : 1, hi = 2 * L0;
        L = L < lo ? lo : (L > hi ? hi : L);
    }
    plan->L = L;
    plan->n_runs = total ? (total + L - 1) / L : 0;
}

// number of runs each bucket intersects (0 for an empty bucket)
__global__ void msm_seg_counts(const uint32_t* __restrict__ offsets, uint32_t n_keys, const RunPlan* __restrict__ plan,
                               uint32_t* __restrict__ scount) {
    uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= n_keys) return;
    const uint32_t L = plan->L;
    const uint32_t lo = offsets[key], hi = offsets[key + 1];
    scount[key] = hi > lo ? (hi - 1) / L - lo / L + 1 : 0;
}

template <class F>
__device__ __forceinline__ void flush_bucket(uint32_t key, uint32_t run, uint32_t L, const XYZZ<F>& acc,
                                             const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ soff,
                                             XYZZ<F>* __restrict__ buckets, XYZZ<F>* __restrict__ partials) {
    const uint32_t s0 = soff[key], ns = soff[key + 1] - s0;
    if (ns == 1) buckets[key] = acc;
    else partials[s0 + (run - offsets[key] / L)] = acc;
}

template <class F, int MIN_BLOCKS, class O>
__global__ void __launch_bounds__(128, MIN_BLOCKS) msm_accumulate(const Affine<F>* __restrict__ bases,
                                                      const Affine<F>* __restrict__ bases2, uint32_t n_split,
                                                      const uint32_t* __restrict__ offsets,
                                                      const uint32_t* __restrict__ sorted,
                                                      const uint32_t* __restrict__ soff, uint32_t n_keys,
                                                      const RunPlan* __restrict__ plan, XYZZ<F>* __restrict__ buckets,
                                                      XYZZ<F>* __restrict__ partials) {
    // the grid is sized for the worst case; the real entry count and the run length derived from it live in
    // device memory so the host never has to wait for them
    const uint32_t run = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t L = plan->L;
    if (run >= plan->n_runs) return;
    const uint32_t total = offsets[n_keys];
    uint32_t k = run * L;
    const uint32_t end = min(k + L, total);
    // bucket of the first entry: the last key with offsets[key] <= k (skips empty buckets)
    uint32_t lo = 0, hi = n_keys;  // invariant: offsets[lo] <= k < offsets[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= k) lo = mid;
        else hi = mid;
    }
    uint32_t key = lo;
    uint32_t bend = offsets[key + 1];
    XYZZ<F> acc = XYZZ<F>::inf();
    // entries below n_split index `bases`, the others `bases2` (the phi half of a GLV table)
    auto base_ptr = [&](uint32_t e) {
        const uint32_t idx = e & 0x7fffffffu;
        return idx < n_split ? bases + idx : bases2 + (idx - n_split);
    };
    // How the NEXT point is fetched while the current addition runs.
    //  * fully inlined G1 build: into registers (the load stays in flight across the straight-line addition).
    //  * every build whose field products are real calls (CallOps) and G2: a load in flight cannot cross a CALL --
    //    its scoreboard is waited for first, which exposed the whole DRAM latency of a table gather once per addition
    //    (ncu, full digit tables: long_scoreboard 0.59 per issue) -- and G2 has no registers for a second point
    //    anyway.  These stage the point through SHARED MEMORY with cp.async (LDGSTS): issued before the addition,
    //    waited for after it; it needs no destination registers, survives the calls, and unlike prefetch.global it is
    //    never dropped on a TLB miss (the digit tables span > 100 GB).  One 16-byte slot column per thread,
    //    [quad][thread], so the LDS.128 / LDGSTS.128 of a warp are conflict-free.
    constexpr bool PREFETCH_TO_REGS = sizeof(F) == sizeof(Fq) && MIN_BLOCKS <= 3 && std::is_same<O, InlineOps>::value;
    constexpr int Q = sizeof(Affine<F>) / 16;
    __shared__ uint4 stage[PREFETCH_TO_REGS ? 1 : Q * 128];
    uint4* my_stage = stage + (PREFETCH_TO_REGS ? 0 : threadIdx.x);
    auto stage_issue = [&](const Affine<F>* p) {
        const uint4* src = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(my_stage + q * 128);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + q));
        }
        asm volatile("cp.async.commit_group;");
    };
    auto stage_take = [&]() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        Affine<F> r;
        uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
        for (int q = 0; q < Q; q++) dst[q] = my_stage[q * 128];
        return r;
    };
    // sorted == nullptr: the identity list (the points of an affine level are summed in array order)
    uint32_t v = sorted ? sorted[k] : k;
    Affine<F> cur;
    if constexpr (PREFETCH_TO_REGS) {
        cur = load_affine(base_ptr(v));
    } else {
        stage_issue(base_ptr(v));
        cur = stage_take();
    }
    bool neg = (v >> 31) != 0;
    for (;;) {
        const uint32_t kn = k + 1;
        Affine<F> nxt;
        uint32_t vn = 0;
        if (kn < end) {
            vn = sorted ? sorted[kn] : kn;
            if constexpr (PREFETCH_TO_REGS) nxt = load_affine(base_ptr(vn));  // in flight during the addition
            else stage_issue(base_ptr(vn));
        }
        ec_madd<F, O>(acc, cur, neg);
        if (kn >= end) break;
        if (kn == bend) {  // next entry belongs to a later bucket
            flush_bucket(key, run, L, acc, offsets, soff, buckets, partials);
            acc = XYZZ<F>::inf();
            do {
                key++;
                bend = offsets[key + 1];
            } while (bend == kn);
        }
        if constexpr (PREFETCH_TO_REGS) cur = nxt;
        else cur = stage_take();
        neg = (vn >> 31) != 0;
        k = kn;
    }
    flush_bucket(key, run, L, acc, offsets, soff, buckets, partials);
}

// thread per bucket: empty -> infinity, <= FOLD_SMALL_MAX partials -> summed here, more -> queued
template <class F>
__global__ void __launch_bounds__(64) msm_fold_small(const uint32_t* __restrict__ toff, uint32_t n_keys,
                                                     const XYZZ<F>* __restrict__ partials,
                                                     XYZZ<F>* __restrict__ buckets, uint32_t* __restrict__ big_list,
                                                     uint32_t* __restrict__ big_count) {
    uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= n_keys) return;
    const uint32_t lo = toff[key], nt = toff[key + 1] - lo;
    if (nt == 1) return;
    if (nt == 0) {
        buckets[key] = XYZZ<F>::inf();
        return;
    }
    if (nt > FOLD_SMALL_MAX) {
        big_list[atomicAdd(big_count, 1u)] = key;
        return;
    }
    XYZZ<F> acc = partials[lo];
    for (uint32_t t = 1; t < nt; t++) {
        XYZZ<F> b = partials[lo + t];
        ec_add<F, CallOps>(acc, b);
    }
    buckets[key] = acc;
}

// one warp per queued bucket: lanes stride over the partials, then a shared-memory tree
template <class F>
__global__ void __launch_bounds__(32) msm_fold_big(const uint32_t* __restrict__ toff,
                                                   const XYZZ<F>* __restrict__ partials,
                                                   XYZZ<F>* __restrict__ buckets,
                                                   const uint32_t* __restrict__ big_list,
                                                   const uint32_t* __restrict__ big_count) {
    __shared__ XYZZ<F> sh[32];
    for (uint32_t item = blockIdx.x; item < *big_count; item += gridDim.x) {
        const uint32_t key = big_list[item];
        const uint32_t lo = toff[key], nt = toff[key + 1] - lo;
        const uint32_t lane = threadIdx.x;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t t = lane; t < nt; t += 32) {
            XYZZ<F> b = partials[lo + t];
            ec_add<F, CallOps>(acc, b);
        }
        sh[lane] = acc;
        __syncwarp();
        for (uint32_t step = 16; step >= 1; step >>= 1) {
            if (lane < step) {
                XYZZ<F> b = sh[lane + step];
                ec_add<F, CallOps>(acc, b);
                sh[lane] = acc;
            }
            __syncwarp();
        }
        if (lane == 0) buckets[key] = acc;
        __syncwarp();
    }
}

// ---------------------------------------------------------------- bucket reduction
// Weighted bucket sum  W = sum_j (j+1) B_j  of one bucket set, organised for DEPTH, not only for work: this
// phase has little arithmetic (2-3 additions per bucket) but sits at the tail of every MSM, where nothing
// is left to overlap with.  One CTA per segment of up to RED_SEG buckets:
//   thread t owns I consecutive buckets: running sums give R_t = sum B and w_t = sum (i+1) B      (2I adds deep)
//   Hillis-Steele suffix scan of R_t over the CTA in shared memory: Inc_t = sum_{u >= t} R_u      (log T deep)
//   v_t = w_t + I * Inc_t (t >= 1), tree sum of v_t                                              (log T deep)
// since  sum_j (j+1) B_j = sum_t [ w_t + t I R_t ]  and  sum_t t R_t = sum_{t >= 1} Inc_t.
// ~36 additions deep for 2048 buckets instead of ~95 for the chunked running-sum version it replaces.
constexpr uint32_t RED_THREADS = 256, RED_SEG = 2048;

// all threads of the CTA call this; results valid in thread 0
template <class F>
__device__ void block_weighted_sum(const XYZZ<F>* __restrict__ items, uint32_t m, XYZZ<F>* sh, XYZZ<F>& W_out,
                                   XYZZ<F>& R_out) {
    const uint32_t t = threadIdx.x, T = blockDim.x;
    uint32_t log_i = 0;
    while ((T << log_i) < m) log_i++;
    const uint32_t I = 1u << log_i;
    XYZZ<F> run = XYZZ<F>::inf(), w = XYZZ<F>::inf();
    if (t * I < m) {
        const XYZZ<F>* B = items + (size_t)t * I;
        for (int i = (int)I - 1; i >= 0; i--) {
            XYZZ<F> b = B[i];
            ec_add<F, CallOps>(run, b);
            ec_add<F, CallOps>(w, run);
        }
    }
    // inclusive suffix scan of `run`
    XYZZ<F> x = run;
    sh[t] = x;
    __syncthreads();
    const uint32_t active = (m + I - 1) >> log_i;  // threads holding data
    for (uint32_t d = 1; d < active; d <<= 1) {
        const bool has = t + d < active;
        XYZZ<F> other;
        if (has) other = sh[t + d];
        __syncthreads();
        if (has) {
            ec_add<F, CallOps>(x, other);
            sh[t] = x;
        }
        __syncthreads();
    }
    XYZZ<F> v = w;
    if (t >= 1 && t < active) {
        XYZZ<F> s = x;
        for (uint32_t d = 0; d < log_i; d++) s = ec_dbl(s);
        ec_add<F, CallOps>(v, s);
    }
    if (t == 0) R_out = x;
    __syncthreads();
    sh[t] = v;
    __syncthreads();
    for (uint32_t step = T >> 1; step >= 1; step >>= 1) {
        if (t < step && t + step < active) {
            XYZZ<F> b = sh[t + step];
            ec_add<F, CallOps>(v, b);
            sh[t] = v;
        }
        __syncthreads();
    }
    if (t == 0) W_out = v;
}

// The same two sums over a short segment, laid out for THROUGHPUT instead of depth: one thread per segment of
// RED_CHUNK consecutive buckets and nothing but the two running sums -- 2 additions per bucket, the minimum, no
// scan and no shared memory.  A batch of proofs has hundreds of bucket sets to reduce at once (128 proofs x 5 MSMs
// x 2,048 buckets), so there is parallelism to spare, and the reduction usually runs beside the bucket accumulation
// of another MSM: what matters there is how few SM cycles it takes away, not its own latency.  (The 256-thread
// scan version below does 2.2x the additions and holds half an SM per CTA for milliseconds; measured
// 8.0 ms per step serialised.)  The per-set tail (msm_seg_combine over nb / RED_CHUNK partial sums) stays log-depth.
constexpr uint32_t RED_CHUNK = 16;
// The fold of the accumulation's partial sums is fused in: a bucket written whole by one run (or folded by
// msm_fold_big because it is huge) is read from `buckets`, an empty one is the identity, and a bucket split over a
// few runs is summed from its partials right here -- one launch and one pass over the buckets less per MSM.
template <class F>
__global__ void __launch_bounds__(128) msm_chunk_reduce(const XYZZ<F>* __restrict__ buckets,
                                                        const XYZZ<F>* __restrict__ partials,
                                                        const uint32_t* __restrict__ toff, uint32_t n_chunks,
                                                        XYZZ<F>* __restrict__ W, XYZZ<F>* __restrict__ R) {
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= n_chunks) return;
    const uint32_t key0 = chunk * RED_CHUNK;
    XYZZ<F> run = XYZZ<F>::inf(), w = XYZZ<F>::inf();
#pragma unroll 1
    for (int i = (int)RED_CHUNK - 1; i >= 0; i--) {
        const uint32_t key = key0 + (uint32_t)i;
        const uint32_t lo = toff[key], nt = toff[key + 1] - lo;
        if (nt == 1 || nt > FOLD_SMALL_MAX) {
            const XYZZ<F> b = buckets[key];
            ec_add<F, CallOps>(run, b);
        } else {
            for (uint32_t t = 0; t < nt; t++) {
                const XYZZ<F> b = partials[lo + t];
                ec_add<F, CallOps>(run, b);
            }
        }
        ec_add<F, CallOps>(w, run);
    }
    W[chunk] = w;
    R[chunk] = run;
}

// thread per bucket: queue the buckets with more than FOLD_SMALL_MAX partials for msm_fold_big (no arithmetic here)
__global__ void msm_fold_classify(const uint32_t* __restrict__ toff, uint32_t n_keys, uint32_t* __restrict__ big_list,
                                  uint32_t* __restrict__ big_count) {
    uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= n_keys) return;
    if (toff[key + 1] - toff[key] > FOLD_SMALL_MAX) big_list[atomicAdd(big_count, 1u)] = key;
}

// segment g of set s covers buckets [g*seg, (g+1)*seg): W[s*segs+g] = sum_i (i+1) B_i, R[...] = sum_i B_i
template <class F>
__global__ void __launch_bounds__(RED_THREADS) msm_seg_reduce(const XYZZ<F>* __restrict__ buckets, uint32_t seg,
                                                              XYZZ<F>* __restrict__ W, XYZZ<F>* __restrict__ R) {
    extern __shared__ uint4 red_smem[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(red_smem);
    XYZZ<F> w, r;
    block_weighted_sum<F>(buckets + (size_t)blockIdx.x * seg, seg, sh, w, r);
    if (threadIdx.x == 0) {
        W[blockIdx.x] = w;
        R[blockIdx.x] = r;
    }
}

// one CTA per set: out = sum_g W_g + seg * sum_{g >= 1} g R_g   (segs <= RED_THREADS)
template <class F>
__global__ void __launch_bounds__(RED_THREADS) msm_seg_combine(const XYZZ<F>* __restrict__ W, const XYZZ<F>* __restrict__ R,
                                                               uint32_t segs, uint32_t log_seg, XYZZ<F>* __restrict__ out) {
    extern __shared__ uint4 red_smem[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(red_smem);
    const uint32_t t = threadIdx.x;
    const XYZZ<F>* Ws = W + (size_t)blockIdx.x * segs;
    const XYZZ<F>* Rs = R + (size_t)blockIdx.x * segs;
    XYZZ<F> hi, unused;
    block_weighted_sum<F>(Rs + 1, segs - 1, sh, hi, unused);  // sum_{g>=1} g R_g (weights 1.. over R_1..)
    __syncthreads();
    XYZZ<F> v = t < segs ? Ws[t] : XYZZ<F>::inf();
    sh[t] = v;
    __syncthreads();
    for (uint32_t step = blockDim.x >> 1; step >= 1; step >>= 1) {
        if (t < step) {
            XYZZ<F> b = sh[t + step];
            ec_add<F, CallOps>(v, b);
            sh[t] = v;
        }
        __syncthreads();
    }
    if (t == 0) {
        for (uint32_t d = 0; d < log_seg; d++) hi = ec_dbl(hi);
        ec_add<F, CallOps>(v, hi);
        out[blockIdx.x] = v;
    }
}

// window sums [batch][weff] -> affine result per msm
template <class F>
__global__ void __launch_bounds__(32) msm_finish(const XYZZ<F>* __restrict__ wsums, uint32_t batch, uint32_t weff,
                                                 uint32_t c, Affine<F>* __restrict__ out) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    XYZZ<F> acc = wsums[(size_t)b * weff + (weff - 1)];
    for (int w = (int)weff - 2; w >= 0; w--) {
        for (uint32_t d = 0; d < c; d++) acc = ec_dbl(acc);
        XYZZ<F> s = wsums[(size_t)b * weff + w];
        ec_add<F, CallOps>(acc, s);
    }
    out[b] = ec_to_affine(acc);
}

// digit tables over GLV half scalars: sums[2 b] = sum of the k1 rows, sums[2 b + 1] = sum of the k2 rows of proof b;
// phi is a homomorphism, so the k2 part of the MSM is phi of that sum: out[b] = sums[2 b] + phi(sums[2 b + 1])
template <class F>
__global__ void __launch_bounds__(32) msm_finish_glv(const XYZZ<F>* __restrict__ sums, uint32_t batch, Affine<F>* __restrict__ out) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    XYZZ<F> acc = sums[2 * (size_t)b];
    XYZZ<F> s2 = sums[2 * (size_t)b + 1];
    s2.x = glv_phi_x(s2.x);  // x = X / ZZ: scaling X scales x
    ec_add<F, CallOps>(acc, s2);
    out[b] = ec_to_affine(acc);
}

// table[w*n + i] = 2^c * table[(w-1)*n + i]
template <class F>
__global__ void __launch_bounds__(64) msm_precompute_step(const Affine<F>* __restrict__ prev, size_t n, uint32_t c,
                                                          Affine<F>* __restrict__ next) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> p = XYZZ<F>::from_affine(load_affine(prev + i));
    for (uint32_t d = 0; d < c; d++) p = ec_dbl(p);
    next[i] = ec_to_affine(p);
}

template <class F>
__global__ void mark_infinity(const Affine<F>* __restrict__ pts, size_t n, uint8_t* __restrict__ flags,
                              unsigned long long* __restrict__ count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool inf = load_affine(pts + i).is_inf();
    flags[i] = inf ? 1 : 0;
    if (inf) atomicAdd(count, 1ull);
}

template <class F>
__global__ void apply_inf_flags(Affine<F>* pts, const uint8_t* flags, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n && flags[i]) pts[i] = Affine<F>::inf();
}

// out = sum of n affine points: lanes stride over the points, then a shared-memory tree (one warp).
// Used to combine the per-GPU partial results of a sharded MSM.
template <class F>
__global__ void __launch_bounds__(32) points_sum_kernel(const Affine<F>* __restrict__ pts, uint32_t n,
                                                        Affine<F>* __restrict__ out) {
    __shared__ XYZZ<F> sh[32];
    const uint32_t lane = threadIdx.x;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = lane; i < n; i += 32) ec_madd<F, CallOps>(acc, load_affine(pts + i));
    sh[lane] = acc;
    __syncwarp();
    for (uint32_t step = 16; step >= 1; step >>= 1) {
        if (lane < step) {
            XYZZ<F> b = sh[lane + step];
            ec_add<F, CallOps>(acc, b);
            sh[lane] = acc;
        }
        __syncwarp();
    }
    if (lane == 0) *out = ec_to_affine(acc);
}

// ---------------------------------------------------------------- full digit tables (precompute level 2)
// With T[(i W + w) mult + m] = (m + 1) 2^(c w) P_i resident, sum_i k_i P_i = sum over the non-zero signed digits
// d of every scalar of  sign(d) * T[w, i, |d| - 1]: one mixed addition per digit, the same count as the bucket
// method's accumulation, and NOTHING else -- no sort (the entries are taken in scalar order), no fold of split
// buckets, no bucket reduction, no window combine.  The entry list of a batch is laid out proof after proof and cut
// into equal runs exactly like the bucket-sorted list (msm_accumulate is reused unchanged, with "bucket" = proof);
// the runs of a proof are then added by one CTA (msm_sum_partials).

// number of non-zero digits of scalar (b, i)
// scalar (b, i) of a table pass; tglv: b counts (proof, half) pairs and k is that GLV half (upper limbs zero)
__device__ __forceinline__ void load_table_scalar(const uint32_t* scalars, size_t b, size_t i, size_t stride, bool mont, bool tglv,
                                                  uint32_t* k) {
    if (!tglv) {
        load_scalar(scalars, b * stride + i, mont, k);
        return;
    }
    uint32_t full[8], k1[GLV_LIMBS], k2[GLV_LIMBS];
    load_scalar(scalars, (b >> 1) * stride + i, mont, full);
    glv_split(full, k1, k2);
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = j < GLV_LIMBS ? ((b & 1) ? k2[j] : k1[j]) : 0u;
}

__global__ void table_count(const uint32_t* scalars, size_t n, size_t stride, size_t batch, int mont, uint32_t c,
                            uint32_t windows, int tglv, const uint8_t* __restrict__ skip, uint32_t* __restrict__ counts) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    const size_t b = t / n, i = t % n;
    uint32_t cnt = 0;
    if (!(skip && skip[i])) {
        uint32_t k[8];
        load_table_scalar(scalars, b, i, stride, mont != 0, tglv != 0, k);
        uint32_t carry = 0;
        for (uint32_t w = 0; w < windows; w++) cnt += next_digit(k, w, c, carry) != 0 ? 1u : 0u;
    }
    counts[t] = cnt;
}

// entries of scalar (b, i) at off[b n + i]...: table index | sign
// pstart (optional): padded first entry of every proof (table_pad_offsets); the list of proof b then lives at
// pstart[b]... and the thread of its last scalar fills the gap up to pstart[b + 1] with AFF_PAD_ENTRY
__global__ void table_entries(const uint32_t* scalars, size_t n, size_t stride, size_t batch, int mont, uint32_t c,
                              uint32_t windows, uint32_t mult, int tglv, const uint8_t* __restrict__ skip,
                              const uint32_t* __restrict__ off, const uint32_t* __restrict__ pstart,
                              uint32_t* __restrict__ entries) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    const size_t b = t / n, i = t % n;
    uint32_t pos = off[t];
    if (pstart) {
        const uint32_t first = off[b * n];
        pos = pos - first + pstart[b];
        if (i == n - 1)
            for (uint32_t q = pstart[b] + (off[(b + 1) * n] - first); q < pstart[b + 1]; q++) entries[q] = AFF_PAD_ENTRY;
    }
    if (skip && skip[i]) return;
    uint32_t k[8];
    load_table_scalar(scalars, b, i, stride, mont != 0, tglv != 0, k);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < windows; w++) {
        const int32_t d = next_digit(k, w, c, carry);
        if (d == 0) continue;
        const uint32_t m = (uint32_t)(d < 0 ? -d : d) - 1;
        entries[pos++] = (uint32_t)((i * windows + w) * mult + m) | (d < 0 ? 0x80000000u : 0u);
    }
}

// poff[b] = first entry of proof b, poff[batch] = total
__global__ void table_proof_offsets(const uint32_t* __restrict__ off, size_t n, size_t batch, uint32_t* __restrict__ poff) {
    size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (b <= batch) poff[b] = off[b * n];
}

// Affine levels (msm_affine.cuh): every proof's entry list is padded to a multiple of 2^levels entries.
// pstart[b] = padded first entry of proof b (pstart[batch] = padded total), poff_out[b] = pstart[b] >> levels = first
// point of proof b after the last level.  One CTA, running carry over chunks of SCAN_THREADS proofs.
// Also the plan of the levels: B[lv] = additions per lane of level lv (pstart[batch + 1 + lv]), chosen from the real
// entry count so that the CTAs of a level fill a whole number of waves (every lane does the same work, so a partly
// filled last wave is pure loss): the nearest whole number of waves at the host's target B0, then
// B = ceil(pairs / (lanes per wave * waves)), at least b_min (below that the shared inversion is no longer amortised).
__global__ void table_pad_offsets(const uint32_t* __restrict__ off, size_t n, size_t batch, uint32_t levels, uint32_t B0,
                                  uint32_t b_min, uint32_t lanes_per_wave, uint32_t* __restrict__ pstart,
                                  uint32_t* __restrict__ poff_out) {
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const uint32_t mask = (1u << levels) - 1;
    for (size_t base = 0; base < batch; base += SCAN_THREADS) {
        const size_t b = base + threadIdx.x;
        const uint32_t len = b < batch ? (off[(b + 1) * n] - off[b * n] + mask) & ~mask : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(len, &total);
        const uint32_t c = carry_s;
        if (b < batch) {
            pstart[b] = ex + c;
            poff_out[b] = (ex + c) >> levels;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pstart[batch] = carry_s;
        poff_out[batch] = carry_s >> levels;
        for (uint32_t lv = 0; lv < levels; lv++) {
            const uint64_t pairs = carry_s >> (lv + 1), per_wave = (uint64_t)lanes_per_wave * B0;
            uint32_t B = B0;
            if (lanes_per_wave) {
                uint64_t waves = (pairs + per_wave / 2) / per_wave;
                if (waves == 0) waves = 1;
                const uint64_t lanes = lanes_per_wave * waves;
                B = (uint32_t)((pairs + lanes - 1) / lanes);
            }
            pstart[batch + 1 + lv] = B < b_min ? b_min : B;
        }
    }
}

// ---- affine levels over the bucket-sorted entry list of plain bases: every bucket is padded to a multiple of 2^levels
// entries (the sorted array is pre-filled with AFF_PAD_ENTRY), so pairs never straddle a bucket at any level
__global__ void pad_counts(uint32_t* __restrict__ counts, uint32_t n_keys, uint32_t levels) {
    const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t mask = (1u << levels) - 1;
    if (key < n_keys) counts[key] = (counts[key] + mask) & ~mask;
}
// plan[0] = padded entry count, plan[1 + lv] = additions per lane of level lv (whole waves, as in table_pad_offsets)
__global__ void bucket_affine_plan(const uint32_t* __restrict__ offsets, uint32_t n_keys, uint32_t levels, uint32_t B0,
                                   uint32_t b_min, uint32_t lanes_per_wave, uint32_t* __restrict__ plan) {
    const uint32_t total = offsets[n_keys];
    plan[0] = total;
    for (uint32_t lv = 0; lv < levels; lv++) {
        const uint64_t pairs = total >> (lv + 1), per_wave = (uint64_t)lanes_per_wave * B0;
        uint32_t B = B0;
        if (lanes_per_wave) {
            uint64_t waves = (pairs + per_wave / 2) / per_wave;
            if (waves == 0) waves = 1;
            const uint64_t lanes = lanes_per_wave * waves;
            B = (uint32_t)((pairs + lanes - 1) / lanes);
        }
        plan[1 + lv] = B < b_min ? b_min : B;
    }
}
__global__ void shift_offsets(uint32_t* __restrict__ offsets, uint32_t count, uint32_t levels) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) offsets[i] >>= levels;
}

// one CTA per proof: sum of the partial sums its runs left (or the single whole sum), log-depth
template <class F>
__global__ void __launch_bounds__(128) msm_sum_partials(const uint32_t* __restrict__ toff, const XYZZ<F>* __restrict__ partials,
                                                        const XYZZ<F>* __restrict__ whole, XYZZ<F>* __restrict__ out) {
    extern __shared__ uint4 red_smem[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(red_smem);
    const uint32_t key = blockIdx.x, t = threadIdx.x;
    const uint32_t lo = toff[key], nt = toff[key + 1] - lo;
    XYZZ<F> acc = XYZZ<F>::inf();
    if (nt == 1) {
        if (t == 0) acc = whole[key];
    } else {
        for (uint32_t j = t; j < nt; j += blockDim.x) {
            const XYZZ<F> b = partials[lo + j];
            ec_add<F, CallOps>(acc, b);
        }
    }
    sh[t] = acc;
    __syncthreads();
    for (uint32_t step = blockDim.x >> 1; step >= 1; step >>= 1) {
        if (t < step) {
            const XYZZ<F> b = sh[t + step];
            ec_add<F, CallOps>(acc, b);
            sh[t] = acc;
        }
        __syncthreads();
    }
    if (t == 0) out[key] = acc;
}

// table build, step 1: table row r = i * windows + w (base-major: the ~22 rows one scalar touches are adjacent, 4 MB
// apart at most, so consecutive entries of a run mostly share a 2 MB page -- with the window-major order of Q they
// were n rows = 1 GB apart and every gather was a TLB miss) <- the first `mult` multiples of Q[w * n + i], as XYZZ
template <class F>
__global__ void __launch_bounds__(64) table_multiples(const Affine<F>* __restrict__ Q, size_t n, uint32_t windows, size_t r0,
                                                      size_t rows, uint32_t mult, XYZZ<F>* __restrict__ out) {
    size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const size_t i = (r0 + r) / windows, w = (r0 + r) % windows;
    const Affine<F> q = load_affine(Q + w * n + i);
    XYZZ<F> acc = XYZZ<F>::inf();
    XYZZ<F>* dst = out + r * mult;
    for (uint32_t m = 0; m < mult; m++) {
        ec_madd<F, CallOps>(acc, q);
        dst[m] = acc;
    }
}

// table build, step 2: XYZZ -> affine, TA_CHUNK points per thread sharing one inversion (Montgomery's trick)
constexpr int TA_CHUNK = 16;
template <class F>
__global__ void __launch_bounds__(64) table_to_affine(const XYZZ<F>* __restrict__ in, size_t count, Affine<F>* __restrict__ out) {
    size_t chunk = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t base = chunk * TA_CHUNK;
    if (base >= count) return;
    const int m = (int)(count - base < (size_t)TA_CHUNK ? count - base : (size_t)TA_CHUNK);
    F pref[TA_CHUNK];
    F run = F::one();
    for (int k = 0; k < m; k++) {      // running product of the denominators zz * zzz (1 for a point at infinity)
        const XYZZ<F> p = in[base + k];
        pref[k] = run;
        if (!p.is_inf()) run = fp_mul(run, fp_mul(p.zz, p.zzz));
    }
    F inv = fp_inv(run);
    for (int k = m - 1; k >= 0; k--) {
        const XYZZ<F> p = in[base + k];
        if (p.is_inf()) {
            out[base + k] = Affine<F>::inf();
            continue;
        }
        const F zi = fp_mul(inv, pref[k]);                 // 1 / (zz zzz) of point k
        inv = fp_mul(inv, fp_mul(p.zz, p.zzz));
        out[base + k] = Affine<F>{fp_mul(p.x, fp_mul(zi, p.zzz)), fp_mul(p.y, fp_mul(zi, p.zz))};
    }
}

int exclusive_scan(b200zk_ctx* ctx, cudaStream_t st, int slot, const uint32_t* d_in, size_t n,
                   uint32_t* d_out /* n+1 */) {
    const unsigned blocks = div_up(n, SCAN_BLOCK);
    void* bs;
    B200ZK_TRY(scratch(ctx, "msm_scan_blocks", (size_t)blocks * 4 + 16, &bs, slot));
    scan_block_sums<<<blocks, SCAN_THREADS, 0, st>>>(d_in, n, (uint32_t*)bs);
    B200ZK_TRY(check_launch(ctx, "scan_block_sums"));
    scan_single<<<1, SCAN_THREADS, 0, st>>>((uint32_t*)bs, blocks);
    B200ZK_TRY(check_launch(ctx, "scan_single"));
    scan_apply<<<blocks, SCAN_THREADS, 0, st>>>(d_in, n, (const uint32_t*)bs, d_out);
    return check_launch(ctx, "scan_apply");
}

}  // namespace

namespace b200zk {

// One pass of the pipeline (sort -> accumulate -> fold -> bucket reduction) over the windows [pl.w0, pl.w1) on the
// stream of `slot`, with that slot's scratch buffers.  *sums_out: device array [batch][pl.weff] of window sums.
template <class F>
int msm_part(b200zk_ctx* ctx, const b200zk_bases* h, const uint32_t* d_scalars, size_t n, size_t stride, size_t batch,
             bool mont, const Plan& pl, int slot, const Affine<F>* d_phi, const XYZZ<F>** sums_out) {
    const cudaStream_t st = slot_stream(ctx, slot);
    if (!ctx->concurrency) slot = 0;
    const uint32_t part_windows = pl.w1 - pl.w0;
    const uint64_t n_keys64 = (uint64_t)batch * pl.weff * pl.nb;
    const uint64_t max_entries = (uint64_t)batch * n * part_windows * (pl.glv ? 2 : 1);
    if (n_keys64 >= (1ull << 31) || max_entries >= (1ull << 32) || (uint64_t)h->n * pl.windows >= (1ull << 31) ||
        (pl.glv && 2 * (uint64_t)n >= (1ull << 31)))
        return fail(ctx, B200ZK_ERR_BAD_LEN, "MSM too large for 32-bit bucket keys; split the batch");
    if (pl.table && ((uint64_t)h->n * pl.windows * pl.mult >= (1ull << 31) || (uint64_t)batch * h->n + 1 >= (1ull << 32)))
        return fail(ctx, B200ZK_ERR_BAD_LEN, "digit table too large for 31-bit entry indices");
    const uint32_t n_keys = (uint32_t)n_keys64;

    void *d_counts, *d_offsets, *d_cursor, *d_sorted, *d_buckets, *d_red_a, *d_red_b;
    B200ZK_TRY(scratch(ctx, "msm_counts", ((size_t)n_keys + 1) * 4, &d_counts, slot));
    B200ZK_TRY(scratch(ctx, "msm_offsets", ((size_t)n_keys + 1) * 4, &d_offsets, slot));
    B200ZK_TRY(scratch(ctx, "msm_cursor", ((size_t)n_keys + 1) * 4, &d_cursor, slot));
    B200ZK_TRY(scratch(ctx, "msm_sorted", ((size_t)max_entries + (pl.table ? batch * 256 : (size_t)n_keys * 16)) * 4, &d_sorted, slot));   // + affine-level padding
    B200ZK_TRY(scratch(ctx, sizeof(F) == sizeof(Fq) ? "msm_buckets_g1" : "msm_buckets_g2",
                       (size_t)n_keys * sizeof(XYZZ<F>), &d_buckets, slot));

    const size_t total = n * batch;
    // ---- affine levels (msm_affine.cuh): K rounds of pairwise batched-affine additions before the XYZZ running sums
    uint32_t aff_levels = 0;
    if (ctx->msm_affine_levels > 0 &&
        max_entries >= (uint64_t)(pl.table ? ctx->msm_affine_min_entries : ctx->msm_affine_min_entries_buckets)) {
        aff_levels = (uint32_t)ctx->msm_affine_levels;
        static const int buckets_env = getenv("B200ZK_AFFINE_BUCKETS") ? atoi(getenv("B200ZK_AFFINE_BUCKETS")) : 1;
        if (!pl.table && !buckets_env) aff_levels = 0;
        if (!pl.table) {
            // buckets: padding costs (2^levels - 1) / 2 entries per bucket on average -> levels <= log2(mean bucket / 8);
            // and the levels need room (half the entries as points, twice): skipped when the device is that full
            const uint64_t mean = max_entries / std::max<uint64_t>(1, n_keys64);
            aff_levels = std::min(aff_levels, 4u);
            while (aff_levels > 0 && (8ull << aff_levels) > mean) aff_levels--;
            size_t free_b = 0, total_b = 0;
            const uint64_t need = (max_entries + (n_keys64 << aff_levels)) * (sizeof(Affine<F>) * 3 / 4 + sizeof(F) / 2 + 8);
            if (aff_levels && (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < need + (need >> 3))) {
                // unless this slot already holds buffers that large from an earlier call
                auto it = ctx->scratch.find(std::string(sizeof(F) == sizeof(Fq) ? "msm_aff_a_g1" : "msm_aff_a_g2") +
                                            (slot > 0 ? "#" + std::to_string(slot) : ctx->lane ? "@" + std::to_string(ctx->lane) : ""));
                if (it == ctx->scratch.end() || it->second.bytes < (max_entries / 2) * sizeof(Affine<F>)) aff_levels = 0;
            }
        }
    }
    const uint64_t padded_entries = max_entries + (aff_levels ? (pl.table ? (uint64_t)batch : n_keys64) << aff_levels : 0);
    if (padded_entries >= (1ull << 32)) return fail(ctx, B200ZK_ERR_BAD_LEN, "MSM too large for 32-bit entry offsets; split the batch");
    const Affine<F>* acc_bases = (const Affine<F>*)(pl.table ? h->d_table : h->d_points);
    const uint32_t* acc_entries = (const uint32_t*)d_sorted;
    uint32_t acc_n_split = pl.glv ? (uint32_t)n : 0x80000000u;
    // the levels: entries (d_sorted) over `first` / `first2` -> A (half the entries) -> B (a quarter) -> A ...; d_plan[0] =
    // padded entry count, d_plan[1 + lv] = additions per lane of level lv (both fixed on the device)
    // resident CTAs per SM of the level kernels (164 / 255 registers).  Measured: one more CTA per SM (128 / 168
    // registers, 0.4 / 1 KB of spills per thread) is slower, G1 15.5 -> 16.2 ms, G2 6.8 -> 8.9 ms per step
    const int aff_occ = sizeof(F) == sizeof(Fq) ? 3 : 2;
    auto run_levels = [&](const Affine<F>* first, const Affine<F>* first2, uint32_t n_split, const uint32_t* d_plan_aff) -> int {
        ProfScope ps(ctx, sizeof(F) == sizeof(Fq) ? "msm_affine_g1" : "msm_affine_g2", st);
        void *d_a, *d_b, *d_pre;
        B200ZK_TRY(scratch(ctx, sizeof(F) == sizeof(Fq) ? "msm_aff_a_g1" : "msm_aff_a_g2",
                           (size_t)(padded_entries / 2 + 1) * sizeof(Affine<F>), &d_a, slot));
        B200ZK_TRY(scratch(ctx, sizeof(F) == sizeof(Fq) ? "msm_aff_b_g1" : "msm_aff_b_g2",
                           (size_t)(padded_entries / 4 + 1) * sizeof(Affine<F>), &d_b, slot));
        // prefix products of the forward pass, [row][quad][lane]: one field element per pair of the largest level
        // (+ slack: the last rows are addressed whole)
        B200ZK_TRY(scratch(ctx, sizeof(F) == sizeof(Fq) ? "msm_aff_pre_g1" : "msm_aff_pre_g2",
                           ((size_t)(padded_entries / 2) + 32 * (size_t)AFF_B_MAX) * sizeof(F), &d_pre, slot));
        const Affine<F>* src = first;
        for (uint32_t lv = 0; lv < aff_levels; lv++) {
            Affine<F>* dst = (Affine<F>*)((lv & 1) ? d_b : d_a);
            const uint64_t max_pairs = padded_entries >> (lv + 1);
            // the grid covers the worst case of the plan (B >= 2/3 B0 with several waves, one wave otherwise); warps
            // beyond the real count exit at once
            const size_t B0 = std::max<size_t>(AFF_B_MIN, (size_t)ctx->msm_affine_b);
            const size_t wave_warps = (size_t)ctx->sm_count * aff_occ * 4;
            const unsigned grid = div_up(std::max(wave_warps, (size_t)div_up(max_pairs * 3, 64 * B0)) + 1, 4);
            constexpr int OCC = sizeof(F) == sizeof(Fq) ? 3 : 2;
            auto kern = lv == 0 ? msm_affine_level<F, true, OCC> : msm_affine_level<F, false, OCC>;
            const size_t smem = aff_smem_bytes<F>(lv == 0);
            if (smem > 48 * 1024) B200ZK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 128, smem, st>>>(src, first2, n_split, lv == 0 ? (const uint32_t*)d_sorted : nullptr, d_plan_aff, lv, dst,
                                          (F*)d_pre);
            B200ZK_TRY(check_launch(ctx, "msm_affine_level"));
            src = dst;
        }
        acc_bases = src;
        acc_entries = nullptr;
        acc_n_split = 0x80000000u;
        if (ctx->prof_enabled && !ctx->concurrency) {  // work counter: additions done by the affine levels
            uint32_t padded = 0;
            B200ZK_CUDA(ctx, cudaMemcpyAsync(&padded, d_plan_aff, 4, cudaMemcpyDeviceToHost, st));
            B200ZK_CUDA(ctx, cudaStreamSynchronize(st));
            ctx->stats[sizeof(F) == sizeof(Fq) ? "msm_affine_adds_g1" : "msm_affine_adds_g2"] += padded - (padded >> aff_levels);
        }
        return B200ZK_OK;
    };
    if (pl.table) {
        // entries in scalar order, proof after proof: per-scalar digit counts -> scan -> entries; no sort, no atomics
        void *d_cnt, *d_off, *d_pstart = nullptr;
        {
            ProfScope ps(ctx, "msm_sort", st);
            B200ZK_TRY(scratch(ctx, "msm_tab_cnt", (total + 1) * 4, &d_cnt, slot));
            B200ZK_TRY(scratch(ctx, "msm_tab_off", (total + 1) * 4, &d_off, slot));
            if (aff_levels) B200ZK_TRY(scratch(ctx, "msm_tab_pstart", (batch + 1 + 8) * 4, &d_pstart, slot));
            table_count<<<div_up(total, 256), 256, 0, st>>>(d_scalars, n, stride, batch, mont ? 1 : 0, pl.c, pl.windows,
                                                            pl.tglv ? 1 : 0, h->d_skip, (uint32_t*)d_cnt);
            B200ZK_TRY(check_launch(ctx, "table_count"));
            B200ZK_TRY(exclusive_scan(ctx, st, slot, (const uint32_t*)d_cnt, total, (uint32_t*)d_off));
            if (aff_levels) {
                static const int bal_env = getenv("B200ZK_AFFINE_BALANCE") ? atoi(getenv("B200ZK_AFFINE_BALANCE")) : 1;
                const uint32_t aff_lanes = (uint32_t)ctx->sm_count * (uint32_t)aff_occ * 128u;
                table_pad_offsets<<<1, SCAN_THREADS, 0, st>>>((const uint32_t*)d_off, n, batch, aff_levels, (uint32_t)ctx->msm_affine_b,
                                                              AFF_B_MIN, bal_env ? aff_lanes : 0u, (uint32_t*)d_pstart,
                                                              (uint32_t*)d_offsets);
                B200ZK_TRY(check_launch(ctx, "table_pad_offsets"));
            } else {
                table_proof_offsets<<<div_up(batch + 1, 128), 128, 0, st>>>((const uint32_t*)d_off, n, batch, (uint32_t*)d_offsets);
                B200ZK_TRY(check_launch(ctx, "table_proof_offsets"));
            }
            table_entries<<<div_up(total, 256), 256, 0, st>>>(d_scalars, n, stride, batch, mont ? 1 : 0, pl.c, pl.windows, pl.mult,
                                                              pl.tglv ? 1 : 0, h->d_skip, (const uint32_t*)d_off, (const uint32_t*)d_pstart,
                                                              (uint32_t*)d_sorted);
            B200ZK_TRY(check_launch(ctx, "table_entries"));
        }
        if (aff_levels) B200ZK_TRY(run_levels((const Affine<F>*)h->d_table, nullptr, 0x80000000u, (const uint32_t*)d_pstart + batch));
    } else {
        B200ZK_CUDA(ctx, cudaMemsetAsync(d_counts, 0, ((size_t)n_keys + 1) * 4, st));
        ProfScope ps(ctx, "msm_sort", st);
        msm_count<<<div_up(total, 256), 256, 0, st>>>(d_scalars, n, stride, batch, mont ? 1 : 0, pl, h->d_skip,
                                                               (uint32_t*)d_counts);
        B200ZK_TRY(check_launch(ctx, "msm_count"));
        if (aff_levels) {  // bucket starts on multiples of 2^levels, the gaps hold the padding entry
            pad_counts<<<div_up(n_keys, 256), 256, 0, st>>>((uint32_t*)d_counts, n_keys, aff_levels);
            B200ZK_TRY(check_launch(ctx, "pad_counts"));
            B200ZK_CUDA(ctx, cudaMemsetAsync(d_sorted, 0xff, (size_t)padded_entries * 4, st));
        }
        B200ZK_TRY(exclusive_scan(ctx, st, slot, (const uint32_t*)d_counts, n_keys, (uint32_t*)d_offsets));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(d_cursor, d_offsets, (size_t)n_keys * 4, cudaMemcpyDeviceToDevice, st));
        msm_scatter<<<div_up(total, 256), 256, 0, st>>>(d_scalars, n, stride, batch, mont ? 1 : 0, pl, h->d_skip,
                                                                 (uint32_t*)d_cursor, (uint32_t*)d_sorted);
        B200ZK_TRY(check_launch(ctx, "msm_scatter"));
    }
    if (aff_levels && !pl.table) {
        void* d_aplan;
        B200ZK_TRY(scratch(ctx, "msm_aff_plan", 16 * 4, &d_aplan, slot));
        static const int bal_env = getenv("B200ZK_AFFINE_BALANCE") ? atoi(getenv("B200ZK_AFFINE_BALANCE")) : 1;
        const uint32_t aff_lanes = (uint32_t)ctx->sm_count * (uint32_t)aff_occ * 128u;
        bucket_affine_plan<<<1, 1, 0, st>>>((const uint32_t*)d_offsets, n_keys, aff_levels, (uint32_t)ctx->msm_affine_b, AFF_B_MIN,
                                            bal_env ? aff_lanes : 0u, (uint32_t*)d_aplan);
        B200ZK_TRY(check_launch(ctx, "bucket_affine_plan"));
        B200ZK_TRY(run_levels((const Affine<F>*)h->d_points, d_phi, pl.glv ? (uint32_t)n : 0x80000000u, (const uint32_t*)d_aplan));
        shift_offsets<<<div_up((size_t)n_keys + 1, 256), 256, 0, st>>>((uint32_t*)d_offsets, n_keys + 1, aff_levels);
        B200ZK_TRY(check_launch(ctx, "shift_offsets"));
    }
    // runs of L = 2^log_tl entries; smaller L when the problem is small, to keep the SMs busy
    void *d_tcount, *d_toff, *d_partials, *d_big;
    // ... and longer runs when the buckets are large (run ~ half the mean bucket), so that a bucket still
    // spans only ~3 runs: with 64-entry runs a 2^26-point GLV MSM (512 entries per bucket) sent every bucket,
    // 9 partials each, to the warp-per-bucket fold (129 ms instead of 6)
    const uint64_t acc_max_entries = aff_levels ? padded_entries >> aff_levels : max_entries;   // what msm_accumulate sums
    uint32_t log_tl = 6;
    while (log_tl < 10 && (acc_max_entries / n_keys) >= (4ull << log_tl)) log_tl++;
    while (log_tl > 3 && (acc_max_entries >> log_tl) < 65536) log_tl--;
    // full digit tables: every run of a proof leaves one partial for msm_sum_partials (one CTA per proof); a small batch
    // must not be cut into so many runs that summing them becomes the latency of the MSM (one proof in 8-entry runs:
    // 15,000 partials, 1.2 ms on one CTA)
    if (pl.table)
        while (log_tl < 10 && ((acc_max_entries / n_keys) >> log_tl) > 4096) log_tl++;
    const uint32_t L0 = 1u << log_tl, L_min = L0 > 1 ? L0 / 2 : 1;   // msm_plan_runs picks L in [L0 / 2, 2 L0]
    const uint64_t max_runs = acc_max_entries / L_min + 1;
    const uint64_t max_segs = (uint64_t)n_keys + max_runs + 1;
    // ---- which build of the accumulation kernel runs, and how many of its threads one wave holds
    // resident CTAs per SM (register cap 65536 / (128 * blocks)); tunable for experiments
    static const int occ_env = getenv("B200ZK_ACC_BLOCKS") ? atoi(getenv("B200ZK_ACC_BLOCKS")) : 0;
    static const int occ2_env = getenv("B200ZK_ACC_BLOCKS_G2") ? atoi(getenv("B200ZK_ACC_BLOCKS_G2")) : occ_env;
    const int occ_sel = sizeof(F) == sizeof(Fq) ? occ_env : occ2_env;
    // measured (profiles/r02_acc_sweep.log, 128-proof step): with the call-based products G1 gains 8 % from a
    // third resident CTA (24.1 -> 22.35 ms, 0.82 -> 0.885 of the Fq-mul peak; a fourth adds 0.5 % and costs the
    // concurrent tails), G2 is best at 2 CTAs (255 registers, 12.9 ms; 168 registers spill: 13.2 ms)
    const int occ = occ_sel ? occ_sel : (sizeof(F) == sizeof(Fq) ? 3 : 2);
    // CTA width (<= 128): narrower CTAs leave registers for latency-bound CTAs of other MSMs (experiments)
    static const int thr_env = getenv("B200ZK_ACC_THREADS_G2") ? atoi(getenv("B200ZK_ACC_THREADS_G2")) : 0;
    const unsigned acc_threads = (sizeof(F) != sizeof(Fq) && (thr_env == 64 || thr_env == 96)) ? (unsigned)thr_env : 128u;
    static const int balance_env = getenv("B200ZK_ACC_BALANCE") ? atoi(getenv("B200ZK_ACC_BALANCE")) : 1;
    const uint32_t lanes_per_wave = (uint32_t)ctx->sm_count * (uint32_t)std::min(occ, 4) * acc_threads;
    void* d_plan;
    B200ZK_TRY(scratch(ctx, "msm_plan", sizeof(RunPlan), &d_plan, slot));
    B200ZK_TRY(scratch(ctx, "msm_tcount", ((size_t)n_keys + 1) * 4, &d_tcount, slot));
    B200ZK_TRY(scratch(ctx, "msm_toff", ((size_t)n_keys + 1) * 4, &d_toff, slot));
    B200ZK_TRY(scratch(ctx, "msm_big", ((size_t)max_runs / FOLD_SMALL_MAX + 8) * 4, &d_big, slot));
    {
        ProfScope ps(ctx, "msm_tasks", st);
        msm_plan_runs<<<1, 1, 0, st>>>((const uint32_t*)d_offsets, n_keys, L0, balance_env ? lanes_per_wave : 0u, (RunPlan*)d_plan);
        B200ZK_TRY(check_launch(ctx, "msm_plan_runs"));
        msm_seg_counts<<<div_up(n_keys, 256), 256, 0, st>>>((const uint32_t*)d_offsets, n_keys, (const RunPlan*)d_plan,
                                                            (uint32_t*)d_tcount);
        B200ZK_TRY(check_launch(ctx, "msm_seg_counts"));
        B200ZK_TRY(exclusive_scan(ctx, st, slot, (const uint32_t*)d_tcount, n_keys, (uint32_t*)d_toff));
        B200ZK_CUDA(ctx, cudaMemsetAsync(d_big, 0, 4, st));
        if (ctx->prof_enabled && !ctx->concurrency) {  // work counters for the roofline (costs a host sync: serialised profiling runs only)
            uint32_t n_entries = 0;
            B200ZK_CUDA(ctx, cudaMemcpyAsync(&n_entries, (const uint32_t*)d_offsets + n_keys, 4, cudaMemcpyDeviceToHost, st));
            B200ZK_CUDA(ctx, cudaStreamSynchronize(st));
            ctx->stats[sizeof(F) == sizeof(Fq) ? "msm_entries_g1" : "msm_entries_g2"] += n_entries;
            ctx->stats[sizeof(F) == sizeof(Fq) ? "msm_buckets_g1" : "msm_buckets_g2"] += n_keys;
        }
    }
    B200ZK_TRY(scratch(ctx, sizeof(F) == sizeof(Fq) ? "msm_partials_g1" : "msm_partials_g2",
                       ((size_t)max_segs + 1) * sizeof(XYZZ<F>), &d_partials, slot));
    uint32_t* big_count = (uint32_t*)d_big;
    uint32_t* big_list = (uint32_t*)d_big + 1;
    {
        ProfScope ps(ctx, sizeof(F) == sizeof(Fq) ? "msm_accumulate_g1" : "msm_accumulate_g2", st);
        // how the field products are issued (ec.cuh): 0 = expanded in place, 1 = calls to one Fq product / square
        // (operands by value, in registers), 2 = G2 only: Fq2 products as calls around the Fq calls
        static const int var_env = getenv("B200ZK_ACC_VARIANT") ? atoi(getenv("B200ZK_ACC_VARIANT")) : 1;
        static const int var2_env = getenv("B200ZK_ACC_VARIANT_G2") ? atoi(getenv("B200ZK_ACC_VARIANT_G2")) : var_env;
        const int variant = sizeof(F) == sizeof(Fq) ? std::min(var_env, 1) : var2_env;
        auto kern = msm_accumulate<F, 2, CallOps>;
        if (variant == 0) kern = occ >= 3 ? msm_accumulate<F, 3, InlineOps> : msm_accumulate<F, 2, InlineOps>;
        else if (variant == 2) kern = occ >= 3 ? msm_accumulate<F, 3, NestedCallOps> : msm_accumulate<F, 2, NestedCallOps>;
        else kern = occ >= 4 ? msm_accumulate<F, 4, CallOps> : occ == 3 ? msm_accumulate<F, 3, CallOps> : msm_accumulate<F, 2, CallOps>;
        // experiments: a dynamic shared-memory request caps the resident CTAs per SM independently of the register
        // budget the kernel was compiled for (B200ZK_ACC_BLOCKS=3 -> 168 registers, B200ZK_ACC_SMEM_KB=100 -> 2 CTAs)
        static const int smem_kb = getenv("B200ZK_ACC_SMEM_KB") ? atoi(getenv("B200ZK_ACC_SMEM_KB")) : 0;
        if (smem_kb > 0) B200ZK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024));
        kern<<<div_up(max_runs, acc_threads), acc_threads, (size_t)smem_kb * 1024, st>>>(
            acc_bases, d_phi, acc_n_split,
            (const uint32_t*)d_offsets, acc_entries,
            (const uint32_t*)d_toff, n_keys, (const RunPlan*)d_plan, (XYZZ<F>*)d_buckets, (XYZZ<F>*)d_partials);
        B200ZK_TRY(check_launch(ctx, "msm_accumulate"));
    }
    if (pl.table) {  // the runs of a proof are all that is left to add: one CTA per proof
        ProfScope ps(ctx, "msm_reduce", st);
        B200ZK_TRY(scratch(ctx, "msm_red_b", ((size_t)n_keys + 8) * sizeof(XYZZ<F>), &d_red_b, slot));
        const size_t smem = 128 * sizeof(XYZZ<F>);
        if (smem > 48 * 1024) B200ZK_CUDA(ctx, cudaFuncSetAttribute(msm_sum_partials<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        msm_sum_partials<F><<<n_keys, 128, smem, st>>>((const uint32_t*)d_toff, (const XYZZ<F>*)d_partials,
                                                       (const XYZZ<F>*)d_buckets, (XYZZ<F>*)d_red_b);
        B200ZK_TRY(check_launch(ctx, "msm_sum_partials"));
        *sums_out = (const XYZZ<F>*)d_red_b;
        return B200ZK_OK;
    }
    // CTA shape of the scan-shaped reduction: 256 threads x 8 buckets.  Narrower CTAs (64 or 32 threads over 512/256-bucket
    // segments) fit in the registers two resident msm_accumulate CTAs leave free, but measured slower for the proof batch
    // (51.2 vs 50.3 ms per step, profiles/r01_red_sweep.log): B200ZK_RED_THREADS keeps the experiment reproducible.
    static const int red_env = getenv("B200ZK_RED_THREADS") ? atoi(getenv("B200ZK_RED_THREADS")) : 0;
    // small windows (a batch of proofs: c <= 13): throughput layout, RED_CHUNK buckets per thread, fold fused in; large
    // windows (one big MSM, few sets): fold kernel + the scan-shaped CTA per 2,048 buckets
    static const int chunk_env = getenv("B200ZK_RED_CHUNKED") ? atoi(getenv("B200ZK_RED_CHUNKED")) : 1;
    const bool chunked = chunk_env && pl.nb >= RED_CHUNK && pl.nb / RED_CHUNK <= RED_THREADS && red_env == 0;
    {
        ProfScope ps(ctx, "msm_fold", st);
        if (chunked) {
            msm_fold_classify<<<div_up(n_keys, 256), 256, 0, st>>>((const uint32_t*)d_toff, n_keys, big_list, big_count);
            B200ZK_TRY(check_launch(ctx, "msm_fold_classify"));
        } else {
            msm_fold_small<F><<<div_up(n_keys, 64), 64, 0, st>>>((const uint32_t*)d_toff, n_keys,
                                                                           (const XYZZ<F>*)d_partials, (XYZZ<F>*)d_buckets,
                                                                           big_list, big_count);
            B200ZK_TRY(check_launch(ctx, "msm_fold_small"));
        }
        msm_fold_big<F><<<ctx->sm_count * 4, 32, 0, st>>>((const uint32_t*)d_toff, (const XYZZ<F>*)d_partials,
                                                                   (XYZZ<F>*)d_buckets, big_list, big_count);
        B200ZK_TRY(check_launch(ctx, "msm_fold_big"));
    }
    // bucket reduction
    const uint32_t sets = (uint32_t)batch * pl.weff;
    uint32_t red_threads = RED_THREADS;
    if (red_env == 32 || red_env == 64 || red_env == 128 || red_env == 256) red_threads = (uint32_t)red_env;
    const uint32_t seg = chunked ? RED_CHUNK : std::min(pl.nb, red_threads * 8), segs = pl.nb / seg;
    uint32_t log_seg = 0;
    while ((1u << log_seg) < seg) log_seg++;
    if (segs > RED_THREADS) return fail(ctx, B200ZK_ERR_BAD_ARG, "window too large for the bucket reduction");
    uint32_t comb_threads = 32;
    while (comb_threads < segs) comb_threads <<= 1;
    const size_t red_smem = (size_t)RED_THREADS * sizeof(XYZZ<F>);
    B200ZK_TRY(scratch(ctx, "msm_red_a", (size_t)sets * segs * 2 * sizeof(XYZZ<F>), &d_red_a, slot));
    B200ZK_TRY(scratch(ctx, "msm_red_b", ((size_t)sets + 8) * sizeof(XYZZ<F>), &d_red_b, slot));
    {
        ProfScope ps(ctx, "msm_reduce", st);
        if (red_smem > 48 * 1024) {  // per device and per function, so set on every call (a few hundred ns): no
                                     // process-wide "done" flag that a second device or a second thread could trip over
            B200ZK_CUDA(ctx, cudaFuncSetAttribute(msm_seg_reduce<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_smem));
            B200ZK_CUDA(ctx, cudaFuncSetAttribute(msm_seg_combine<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_smem));
        }
        XYZZ<F>* Wseg = (XYZZ<F>*)d_red_a;
        XYZZ<F>* Rseg = Wseg + (size_t)sets * segs;
        if (chunked) {
            msm_chunk_reduce<F><<<div_up((size_t)sets * segs, 128), 128, 0, st>>>(
                (const XYZZ<F>*)d_buckets, (const XYZZ<F>*)d_partials, (const uint32_t*)d_toff, sets * segs, Wseg, Rseg);
            B200ZK_TRY(check_launch(ctx, "msm_chunk_reduce"));
        } else {
            msm_seg_reduce<F><<<sets * segs, red_threads, (size_t)red_threads * sizeof(XYZZ<F>), st>>>(
                (const XYZZ<F>*)d_buckets, seg, Wseg, Rseg);
            B200ZK_TRY(check_launch(ctx, "msm_seg_reduce"));
        }
        XYZZ<F>* src = Wseg;
        if (segs > 1) {
            msm_seg_combine<F><<<sets, comb_threads, (size_t)comb_threads * sizeof(XYZZ<F>), st>>>(
                Wseg, Rseg, segs, log_seg, (XYZZ<F>*)d_red_b);
            B200ZK_TRY(check_launch(ctx, "msm_seg_combine"));
            src = (XYZZ<F>*)d_red_b;
        }
        *sums_out = src;
    }
    return B200ZK_OK;
}

// Device-resident batched MSM.  d_scalars: batch rows (row stride `stride` Fr elements) of n
// scalars, canonical or Montgomery (`mont`).  d_out: batch affine points.
//
// A single MSM over plain (not precomputed) bases can be cut into up to 4 groups of windows, each a full pass
// of the pipeline on its own stream (the caller's first, then auxiliary streams of decreasing priority), the
// window sums meeting in one array for the Horner step.  The idea: the sort of part k+1 (atomics/HBM bound) and
// the fold and bucket reduction of part k (latency bound) run under the bucket accumulation (multiply-pipe
// bound) of another part.  Measured, the overlap is small -- the bucket reduction's 256-thread CTAs take whole
// SMs away from the accumulation -- so it is on by default only from 2^23 points (2^24: 119.7 -> 117.5 ms).
template <class F>
int msm_device(b200zk_ctx* ctx, const b200zk_bases* h, const uint32_t* d_scalars, size_t n, size_t stride,
               size_t batch, bool mont, Affine<F>* d_out, int slot) {
    if (n > h->n) return fail(ctx, B200ZK_ERR_BAD_LEN, "more scalars than bases");
    if (batch == 0) return B200ZK_OK;
    const cudaStream_t st = slot_stream(ctx, slot);
    if (n == 0) {
        B200ZK_CUDA(ctx, cudaMemsetAsync(d_out, 0, batch * sizeof(Affine<F>), st));
        return B200ZK_OK;
    }
    // plain bases: GLV halves the scalar length (glv.cuh); the phi half of the table is rebuilt per call
    // (n products, one pass over the points) so the handle stays a plain array of the caller's bases
    const bool glv = !h->precomputed && ctx->msm_glv && n >= 2;
    Plan pl = make_plan(h->n, h->precomputed != 0, h->precomputed ? h->c : 0, glv, h->d_table && h->table_glv);
    if (h->d_table) {  // full digit table: one "bucket" per MSM of the batch (two with GLV half scalars)
        pl.table = true;
        pl.mult = h->mult;
        pl.weff = 1;
        pl.nb = 1;
    }
    const Affine<F>* d_phi = nullptr;
    if (pl.glv) {
        void* ph;
        B200ZK_TRY(scratch(ctx, sizeof(F) == sizeof(Fq) ? "msm_glv_phi_g1" : "msm_glv_phi_g2", n * sizeof(Affine<F>), &ph,
                           ctx->concurrency ? slot : 0));
        glv_phi_kernel<F><<<div_up(n, 128), 128, 0, st>>>((const Affine<F>*)h->d_points, n, (Affine<F>*)ph);
        B200ZK_TRY(check_launch(ctx, "glv_phi_kernel"));
        d_phi = (const Affine<F>*)ph;
    }
    uint32_t parts = 1;
    if (!pl.precomp && batch == 1 && slot == 0 && ctx->concurrency) {
        if (ctx->msm_parts) parts = (uint32_t)ctx->msm_parts;
        // measured (profiles/r01_msm_parts_sweep.log): +2 % at 2^24, a loss below 2^22 -- and a loss at every size once
        // the affine levels run in front of the accumulation (four small level pipelines instead of one that fills
        // the machine: 2^23 points 60 vs 50 ms, profiles/r02_bench_n2.json), so only without them
        else if (n >= (1u << 23) && ctx->msm_affine_levels == 0) parts = 4;
    }
    parts = std::min(parts, pl.windows);
    const XYZZ<F>* sums = nullptr;
    if (parts == 1) {
        B200ZK_TRY(msm_part<F>(ctx, h, d_scalars, n, stride, pl.tglv ? 2 * batch : batch, mont, pl, slot, d_phi, &sums));
    } else {
        static const int part_slot[4] = {0, 5, 6, 8};   // streams: main, aux[0], aux[1], aux[3] (priority order)
        void* d_ws;
        B200ZK_TRY(scratch(ctx, "msm_wsums", (size_t)pl.windows * sizeof(XYZZ<F>), &d_ws, 0));
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
        uint32_t w = 0;
        for (uint32_t i = 0; i < parts; i++) {
            Plan pp = pl;
            pp.w0 = w;
            pp.w1 = w + (pl.windows - w) / (parts - i);
            pp.weff = pp.w1 - pp.w0;
            w = pp.w1;
            const cudaStream_t ps = slot_stream(ctx, part_slot[i]);
            if (i) B200ZK_CUDA(ctx, cudaStreamWaitEvent(ps, ctx->ev_fork, 0));
            const XYZZ<F>* part_sums = nullptr;
            B200ZK_TRY(msm_part<F>(ctx, h, d_scalars, n, stride, batch, mont, pp, part_slot[i], d_phi, &part_sums));
            B200ZK_CUDA(ctx, cudaMemcpyAsync((XYZZ<F>*)d_ws + pp.w0, part_sums, (size_t)pp.weff * sizeof(XYZZ<F>),
                                             cudaMemcpyDeviceToDevice, ps));
            if (i) {
                B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_join[i], ps));
                B200ZK_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
            }
        }
        sums = (const XYZZ<F>*)d_ws;
    }
    {
        ProfScope ps(ctx, "msm_reduce", st);
        if (pl.tglv) msm_finish_glv<F><<<div_up(batch, 32), 32, 0, st>>>(sums, (uint32_t)batch, d_out);
        else msm_finish<F><<<div_up(batch, 32), 32, 0, st>>>(sums, (uint32_t)batch, pl.weff, pl.c, d_out);
        B200ZK_TRY(check_launch(ctx, "msm_finish"));
    }
    return B200ZK_OK;
}

template int msm_device<Fq>(b200zk_ctx*, const b200zk_bases*, const uint32_t*, size_t, size_t, size_t, bool,
                            Affine<Fq>*, int);
template int msm_device<Fq2>(b200zk_ctx*, const b200zk_bases*, const uint32_t*, size_t, size_t, size_t, bool,
                             Affine<Fq2>*, int);

template <class F>
int bases_build(b200zk_ctx* ctx, b200zk_bases* h, const Affine<F>* d_src, bool src_is_device, const uint8_t* inf_flags,
                size_t n, int precompute) {
    // precompute: 0 = plain bases, 1 = window multiples 2^(c w) P_i, 2 = full digit table (m + 1) 2^(c w) P_i, m < 2^(c-1)
    uint32_t c_table = 0;
    if (precompute >= 2) {  // the window of a full table is set by the memory it may take, not by the operation count
        c_table = sizeof(F) == sizeof(Fq) ? (uint32_t)ctx->table_c_g1 : (uint32_t)ctx->table_c_g2;
        if (c_table < 2 || c_table > 16) return fail(ctx, B200ZK_ERR_BAD_ARG, "table_c_g1 / table_c_g2 must be 2..16");
    }
    // full digit tables are built over GLV half scalars (130 instead of 256 bits of windows: half the table, so a
    // wider window fits) unless msm_glv is off; like the plain GLV path this needs the bases in the order-r subgroup
    const bool table_glv = precompute >= 2 && ctx->msm_glv;
    Plan pl = make_plan(n, precompute != 0, c_table, false, table_glv);
    h->table_glv = table_glv ? 1 : 0;
    h->n = n;
    h->precomputed = precompute ? 1 : 0;
    h->c = pl.c;
    h->windows = pl.windows;
    const size_t copies = precompute ? pl.windows : 1;
    B200ZK_CUDA(ctx, cudaMalloc(&h->d_points, std::max<size_t>(1, n * copies) * sizeof(Affine<F>)));
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(h->d_points, d_src, n * sizeof(Affine<F>),
                                     src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    if (inf_flags) {
        void* df;
        B200ZK_TRY(scratch(ctx, "msm_inf", n, &df));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(df, inf_flags, n, cudaMemcpyHostToDevice, ctx->stream));
        apply_inf_flags<F><<<div_up(n, 256), 256, 0, ctx->stream>>>((Affine<F>*)h->d_points, (const uint8_t*)df, n);
        B200ZK_TRY(check_launch(ctx, "apply_inf_flags"));
    }
    {   // bases at infinity are dropped at sort time instead of costing a (no-op) bucket addition each
        void* dcount;
        B200ZK_TRY(scratch(ctx, "msm_inf_count", 8, &dcount));
        B200ZK_CUDA(ctx, cudaMemsetAsync(dcount, 0, 8, ctx->stream));
        B200ZK_CUDA(ctx, cudaMalloc(&h->d_skip, n));
        mark_infinity<F><<<div_up(n, 256), 256, 0, ctx->stream>>>((const Affine<F>*)h->d_points, n, h->d_skip,
                                                                  (unsigned long long*)dcount);
        B200ZK_TRY(check_launch(ctx, "mark_infinity"));
        unsigned long long cnt = 0;
        B200ZK_CUDA(ctx, cudaMemcpyAsync(&cnt, dcount, 8, cudaMemcpyDeviceToHost, ctx->stream));
        B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        h->n_skip = (size_t)cnt;
        if (cnt == 0) {
            cudaFree(h->d_skip);
            h->d_skip = nullptr;
        }
    }
    if (precompute) {
        Affine<F>* t = (Affine<F>*)h->d_points;
        for (uint32_t w = 1; w < pl.windows; w++) {
            msm_precompute_step<F><<<div_up(n, 64), 64, 0, ctx->stream>>>(t + (size_t)(w - 1) * n, n, pl.c,
                                                                          t + (size_t)w * n);
            B200ZK_TRY(check_launch(ctx, "msm_precompute_step"));
        }
    }
    if (precompute >= 2 && n) {
        // ---- full digit table from the window multiples Q[w][i] just built: row r = (w, i) -> (m + 1) Q[r], m < mult.
        // Built slab by slab through an XYZZ scratch (one inversion per TA_CHUNK points); a one-off cost per key.
        const uint32_t mult = pl.nb;
        const size_t rows = n * (size_t)pl.windows;
        if ((uint64_t)rows * mult >= (1ull << 31)) return fail(ctx, B200ZK_ERR_BAD_LEN, "digit table too large for 31-bit entry indices");
        const size_t table_bytes = rows * mult * sizeof(Affine<F>);
        cudaError_t e = cudaMalloc(&h->d_table, table_bytes);
        if (e != cudaSuccess)
            return fail(ctx, B200ZK_ERR_CUDA, "digit table of " + std::to_string(table_bytes >> 20) + " MiB does not fit: " +
                                                  cudaGetErrorString(e) + " (lower table_c_g1 / table_c_g2)");
        h->mult = mult;
        size_t slab_rows = std::max<size_t>(1, ((size_t)2 << 30) / ((size_t)mult * sizeof(XYZZ<F>)));
        slab_rows = std::min(slab_rows, rows);
        void* d_x;
        B200ZK_CUDA(ctx, cudaMalloc(&d_x, slab_rows * mult * sizeof(XYZZ<F>)));
        const Affine<F>* Q = (const Affine<F>*)h->d_points;
        int rc = B200ZK_OK;
        for (size_t r0 = 0; r0 < rows && rc == B200ZK_OK; r0 += slab_rows) {
            const size_t nr = std::min(slab_rows, rows - r0), cnt = nr * mult;
            table_multiples<F><<<div_up(nr, 64), 64, 0, ctx->stream>>>(Q, n, pl.windows, r0, nr, mult, (XYZZ<F>*)d_x);
            rc = check_launch(ctx, "table_multiples");
            if (rc != B200ZK_OK) break;
            table_to_affine<F><<<div_up(div_up(cnt, TA_CHUNK), 64), 64, 0, ctx->stream>>>((const XYZZ<F>*)d_x, cnt,
                                                                                         (Affine<F>*)h->d_table + r0 * mult);
            rc = check_launch(ctx, "table_to_affine");
        }
        cudaError_t se = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_x);
        if (rc != B200ZK_OK) return rc;
        if (se != cudaSuccess) return fail(ctx, B200ZK_ERR_CUDA, std::string("digit table build: ") + cudaGetErrorString(se));
    }
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host source buffers may go away
    return B200ZK_OK;
}

template int bases_build<Fq>(b200zk_ctx*, b200zk_bases*, const Affine<Fq>*, bool, const uint8_t*, size_t, int);
template int bases_build<Fq2>(b200zk_ctx*, b200zk_bases*, const Affine<Fq2>*, bool, const uint8_t*, size_t, int);

}  // namespace b200zk

namespace {

template <class F>
int msm_host_entry(b200zk_ctx* ctx, int group, const uint8_t* bases, const uint8_t* inf_flags, const uint8_t* scalars,
                   size_t n, uint8_t* out_affine, uint8_t* out_is_inf) {
    if (!ctx || !out_affine || (n && (!bases || !scalars))) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    b200zk_bases h;
    h.group = group;
    int rc = bases_build<F>(ctx, &h, (const Affine<F>*)bases, false, inf_flags, n, 0);
    if (rc == B200ZK_OK) {
        void *ds, *dout;
        rc = scratch(ctx, "msm_scalars", std::max<size_t>(32, n * 32), &ds);
        if (rc == B200ZK_OK) rc = scratch(ctx, "msm_out", sizeof(Affine<F>), &dout);
        if (rc == B200ZK_OK && n)
            if (cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
                rc = fail(ctx, B200ZK_ERR_CUDA, "scalar upload failed");
        if (rc == B200ZK_OK) rc = msm_device<F>(ctx, &h, (const uint32_t*)ds, n, n, 1, false, (Affine<F>*)dout);
        if (rc == B200ZK_OK) {
            Affine<F> r;
            if (cudaMemcpyAsync(&r, dout, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                rc = fail(ctx, B200ZK_ERR_CUDA, std::string("msm failed: ") + cudaGetErrorString(cudaGetLastError()));
            else {
                memcpy(out_affine, &r, sizeof(r));
                if (out_is_inf) *out_is_inf = r.is_inf() ? 1 : 0;
            }
        }
    }
    cudaStreamSynchronize(ctx->stream);
    if (h.d_points) cudaFree(h.d_points);
    if (h.d_skip) cudaFree(h.d_skip);
    if (h.d_table) cudaFree(h.d_table);
    return rc;
}

}  // namespace

extern "C" {

int b200zk_msm_g1(b200zk_ctx* ctx, const uint8_t* bases, const uint8_t* inf_flags, const uint8_t* scalars, size_t n,
                  uint8_t out_affine[96], uint8_t* out_is_inf) {
    return msm_host_entry<Fq>(ctx, 1, bases, inf_flags, scalars, n, out_affine, out_is_inf);
}

int b200zk_msm_g2(b200zk_ctx* ctx, const uint8_t* bases, const uint8_t* inf_flags, const uint8_t* scalars, size_t n,
                  uint8_t out_affine[192], uint8_t* out_is_inf) {
    return msm_host_entry<Fq2>(ctx, 2, bases, inf_flags, scalars, n, out_affine, out_is_inf);
}

int b200zk_bases_upload(b200zk_ctx* ctx, int group, const uint8_t* bases, const uint8_t* inf_flags, size_t n,
                        int precompute, b200zk_bases** out) {
    if (!ctx || !out || (n && !bases)) return B200ZK_ERR_BAD_ARG;
    if (group != 1 && group != 2) return fail(ctx, B200ZK_ERR_BAD_ARG, "group must be 1 or 2");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    b200zk_bases* h = new b200zk_bases();
    h->group = group;
    int rc = group == 1 ? bases_build<Fq>(ctx, h, (const Affine<Fq>*)bases, false, inf_flags, n, precompute)
                        : bases_build<Fq2>(ctx, h, (const Affine<Fq2>*)bases, false, inf_flags, n, precompute);
    if (rc != B200ZK_OK) {
        if (h->d_points) cudaFree(h->d_points);
        if (h->d_skip) cudaFree(h->d_skip);
        if (h->d_table) cudaFree(h->d_table);
        delete h;
        return rc;
    }
    *out = h;
    return B200ZK_OK;
}

int b200zk_bases_from_device(b200zk_ctx* ctx, int group, const void* d_points, size_t n, int precompute,
                             b200zk_bases** out) {
    if (!ctx || !out || (n && !d_points)) return B200ZK_ERR_BAD_ARG;
    if (group != 1 && group != 2) return fail(ctx, B200ZK_ERR_BAD_ARG, "group must be 1 or 2");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    b200zk_bases* h = new b200zk_bases();
    h->group = group;
    int rc = group == 1 ? bases_build<Fq>(ctx, h, (const Affine<Fq>*)d_points, true, nullptr, n, precompute)
                        : bases_build<Fq2>(ctx, h, (const Affine<Fq2>*)d_points, true, nullptr, n, precompute);
    if (rc != B200ZK_OK) {
        if (h->d_points) cudaFree(h->d_points);
        if (h->d_skip) cudaFree(h->d_skip);
        if (h->d_table) cudaFree(h->d_table);
        delete h;
        return rc;
    }
    *out = h;
    return B200ZK_OK;
}

void b200zk_bases_free(b200zk_ctx* ctx, b200zk_bases* h) {
    if (!h) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (h->d_points) cudaFree(h->d_points);
    if (h->d_skip) cudaFree(h->d_skip);
    if (h->d_table) cudaFree(h->d_table);
    delete h;
}

int b200zk_msm_resident(b200zk_ctx* ctx, const b200zk_bases* h, const void* scalars, int scalars_on_device, size_t n,
                        size_t batch, uint8_t* out_affine, uint8_t* out_is_inf) {
    if (!ctx || !h || !out_affine || (n && batch && !scalars)) return B200ZK_ERR_BAD_ARG;
    if (n > h->n) return fail(ctx, B200ZK_ERR_BAD_LEN, "more scalars than bases");
    if (batch == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pt = h->group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    const uint32_t* ds = (const uint32_t*)scalars;
    if (!scalars_on_device) {
        void* d;
        B200ZK_TRY(scratch(ctx, "msm_scalars", std::max<size_t>(32, n * batch * 32), &d));
        if (n) B200ZK_CUDA(ctx, cudaMemcpyAsync(d, scalars, n * batch * 32, cudaMemcpyHostToDevice, ctx->stream));
        ds = (const uint32_t*)d;
    }
    void* dout;
    B200ZK_TRY(scratch(ctx, "msm_out", batch * pt, &dout));
    if (h->group == 1) B200ZK_TRY(msm_device<Fq>(ctx, h, ds, n, n, batch, false, (G1Affine*)dout));
    else B200ZK_TRY(msm_device<Fq2>(ctx, h, ds, n, n, batch, false, (G2Affine*)dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_affine, dout, batch * pt, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_is_inf)
        for (size_t b = 0; b < batch; b++) {
            bool z = true;
            for (size_t i = 0; i < pt; i++) z = z && out_affine[b * pt + i] == 0;
            out_is_inf[b] = z ? 1 : 0;
        }
    return B200ZK_OK;
}

int b200zk_msm_resident_device(b200zk_ctx* ctx, const b200zk_bases* h, const void* scalars, int scalars_on_device,
                               size_t n, size_t batch, void* d_out_affine) {
    if (!ctx || !h || !d_out_affine || (n && batch && !scalars)) return B200ZK_ERR_BAD_ARG;
    if (n > h->n) return fail(ctx, B200ZK_ERR_BAD_LEN, "more scalars than bases");
    if (batch == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t* ds = (const uint32_t*)scalars;
    if (!scalars_on_device) {
        void* d;
        B200ZK_TRY(scratch(ctx, "msm_scalars", std::max<size_t>(32, n * batch * 32), &d));
        if (n) B200ZK_CUDA(ctx, cudaMemcpyAsync(d, scalars, n * batch * 32, cudaMemcpyHostToDevice, ctx->stream));
        ds = (const uint32_t*)d;
    }
    if (h->group == 1) return msm_device<Fq>(ctx, h, ds, n, n, batch, false, (G1Affine*)d_out_affine);
    return msm_device<Fq2>(ctx, h, ds, n, n, batch, false, (G2Affine*)d_out_affine);
}

int b200zk_points_sum_device(b200zk_ctx* ctx, int group, const void* d_points, size_t n, void* d_out_affine) {
    if (!ctx || !d_out_affine || (n && !d_points)) return B200ZK_ERR_BAD_ARG;
    if (group != 1 && group != 2) return fail(ctx, B200ZK_ERR_BAD_ARG, "group must be 1 or 2");
    if (n >= (1ull << 32)) return fail(ctx, B200ZK_ERR_BAD_LEN, "too many points");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (group == 1) points_sum_kernel<Fq><<<1, 32, 0, ctx->stream>>>((const G1Affine*)d_points, (uint32_t)n, (G1Affine*)d_out_affine);
    else points_sum_kernel<Fq2><<<1, 32, 0, ctx->stream>>>((const G2Affine*)d_points, (uint32_t)n, (G2Affine*)d_out_affine);
    return check_launch(ctx, "points_sum_kernel");
}

int b200zk_points_sum(b200zk_ctx* ctx, int group, const uint8_t* points, size_t n, uint8_t* out_affine,
                      uint8_t* out_is_inf) {
    if (!ctx || !out_affine || (n && !points)) return B200ZK_ERR_BAD_ARG;
    if (group != 1 && group != 2) return fail(ctx, B200ZK_ERR_BAD_ARG, "group must be 1 or 2");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pt = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    void *d, *dout;
    B200ZK_TRY(scratch(ctx, "psum_in", std::max<size_t>(pt, n * pt), &d));
    B200ZK_TRY(scratch(ctx, "psum_out", pt, &dout));
    if (n) B200ZK_CUDA(ctx, cudaMemcpyAsync(d, points, n * pt, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(b200zk_points_sum_device(ctx, group, d, n, dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_affine, dout, pt, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_is_inf) {
        bool z = true;
        for (size_t i = 0; i < pt; i++) z = z && out_affine[i] == 0;
        *out_is_inf = z ? 1 : 0;
    }
    return B200ZK_OK;
}

}  // extern "C"
