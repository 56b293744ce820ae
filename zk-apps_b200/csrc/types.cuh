// Opaque handle types of the C ABI and the entry points translation units share.
#pragma once
#include "common.cuh"
#include "ec.cuh"
#include "host_r1cs.hpp"

struct b200zk_bases {
    int group = 0;             // 1 = G1, 2 = G2
    size_t n = 0;              // points per window
    void* d_points = nullptr;  // Affine<F>[n * (precomputed ? windows : 1)]
    int precomputed = 0;
    uint32_t c = 0, windows = 0;
    uint8_t* d_skip = nullptr;  // n flags: base i is the point at infinity (its digits never enter a bucket);
                                // null when no base is (proving-key b queries are ~30 % infinity)
    size_t n_skip = 0;
    // precompute level 2: the FULL digit table T[(i * windows + w) * mult + m] = (m + 1) * 2^(c w) * P_i, m < mult = 2^(c-1),
    // affine, resident in HBM (tens of GB per query: sized for the 180 GB of a B200).  An MSM over it is a plain sum
    // of one table entry per non-zero signed digit: no buckets, no sort, no bucket reduction.
    void* d_table = nullptr;
    uint32_t mult = 0;
    int table_glv = 0;  // the table's windows cover GLV half scalars (GLV_BITS): k1 rows as they are, phi applied to the k2 sum
};

struct b200zk_r1cs {
    b200zk::host::R1CS cs;
};

namespace b200zk {

// ntt.cu
int ntt_device(b200zk_ctx* ctx, Fr* d_data, uint32_t log_n, bool inverse, const Fr* coset_offset, size_t batch);
// msm.cu (explicitly instantiated for Fq and Fq2)
template <class F>
int msm_device(b200zk_ctx* ctx, const b200zk_bases* h, const uint32_t* d_scalars, size_t n, size_t stride, size_t batch,
               bool mont, Affine<F>* d_out, int slot = 0);
template <class F>
int bases_build(b200zk_ctx* ctx, b200zk_bases* h, const Affine<F>* d_src, bool src_is_device, const uint8_t* inf_flags,
                size_t n, int precompute);
// relation.cu
int poseidon_consts_device(b200zk_ctx* ctx, const host::PoseidonConsts** out);
int update_note_witness_device(b200zk_ctx* ctx, int kind, uint32_t H, uint32_t num_vars, const Fr* d_inputs, size_t batch,
                               Fr* d_z, uint32_t* d_status);
int relation_witness_device(b200zk_ctx* ctx, int relation, int kind, uint32_t H, uint32_t num_vars, const Fr* d_inputs,
                            size_t batch, Fr* d_z, uint32_t* d_status);

// verify.cu: zcash / ark-serialize compressed encoding, one thread per point (device buffers)
int points_compress_device(b200zk_ctx* ctx, int group, const void* d_affine, size_t n, uint8_t* d_out);
int points_decompress_device(b200zk_ctx* ctx, int group, const uint8_t* d_in, size_t n, bool check_subgroup, void* d_affine,
                             int32_t* d_status);

}  // namespace b200zk
