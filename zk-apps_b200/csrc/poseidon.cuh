// Device-side Poseidon (t = 5, alpha = 5, R_F = 8, R_P = 56; relations/src/lib.rs:17-26) shared by the witness
// generator (relation.cu) and the note tree (merkle.cu): 32-byte vector loads/stores and the warp-wide,
// depth-optimised permutation.
#pragma once
#include "types.cuh"

namespace b200zk {
namespace poseidon_dev {

using host::PoseidonConsts;

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// ---- warp-wide Poseidon for the witness generator -------------------------------------------------
// Witness generation is pure latency (a few hundred warps in flight, nothing to overlap with: every MSM
// waits for z), so the permutation is laid out for depth.  Lane 5*l + j (l, j < 5) holds a copy of state
// word j: every lane applies the S-box to its own copy (redundantly across l), multiplies by M[l][j], and
// the row sums  s'_j = sum_k M[j][k] s_k  are gathered with shuffles from lanes 5*j + k -- the gather also
// transposes, so each lane ends up with its own word again.  4 multiplications deep per round instead of 8.
__device__ __forceinline__ Fr warp_get(const Fr& v, int src) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src);
    return r;
}

static __device__ void poseidon_permute_w(Fr& s, int lane, const PoseidonConsts* pc, Fr* trace) {
    const int half = host::POSEIDON_RF / 2;
    const int j = lane % 5, l = lane < 25 ? lane / 5 : 0;  // lanes 25..31 shadow row 0; nobody reads them
    int sbox = 0;
    for (int rnd = 0; rnd < host::POSEIDON_ROUNDS; rnd++) {
        const bool full = rnd < half || rnd >= half + host::POSEIDON_RP;
        s = fp_add(s, ld_fr(&pc->rc[rnd][j]));
        if (full || j == 0) {
            Fr x2 = fp_sqr(s);
            Fr x4 = fp_sqr(x2);
            Fr x5 = fp_mul(x4, s);
            if (trace && lane < 5) {
                Fr* t = trace + (size_t)(sbox + (full ? j : 0)) * 3;
                st_fr(t, x2);
                st_fr(t + 1, x4);
                st_fr(t + 2, x5);
            }
            s = x5;
        }
        sbox += full ? host::POSEIDON_T : 1;
        const Fr p = fp_mul(ld_fr(&pc->mds[l][j]), s);
        Fr acc = warp_get(p, 5 * j);
#pragma unroll 1
        for (int k = 1; k < host::POSEIDON_T; k++) acc = fp_add(acc, warp_get(p, 5 * j + k));
        s = acc;
    }
}

// hash_fix_len_array, warp-wide; returns the digest in every lane
template <class GetIn>
__device__ Fr poseidon_hash_w(int n_in, GetIn in_of, int lane, const PoseidonConsts* pc, Fr* trace) {
    const int j = lane % 5;
    Fr s = Fr::zero();
    if (j == 0) s = ld_fr(&pc->two64);
    const int chunks = n_in / host::POSEIDON_RATE + 1;
    for (int c = 0; c < chunks; c++) {
        const int lo = c * host::POSEIDON_RATE;
        const int len = lo < n_in ? min(host::POSEIDON_RATE, n_in - lo) : 0;
        if (j >= 1 && j <= len) s = fp_add(s, in_of(lo + j - 1));
        if (len + 1 < host::POSEIDON_T && j == len + 1) s = fp_add(s, Fr::one());
        poseidon_permute_w(s, lane, pc, trace ? trace + (size_t)c * host::POSEIDON_TRACE : nullptr);
    }
    return warp_get(s, 1);
}

}  // namespace poseidon_dev
}  // namespace b200zk
