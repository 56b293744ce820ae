// The note tree on the GPU: the step before the prover -- producing the
// (path, path_shape, root) of every note that update_note_circuit consumes.
//
// Semantics follow the reference's only executable tree, with the circuit's hash in place of SHA-256:
//   layout / add_leaf / gen_proof / root / roots_log   shielder/contract/merkle.rs:11-22,48-80,82-102,104-106
//       heap layout (root = node 1, leaves at size + idx, size = 2^DEPTH); a node that was never written
//       reads as 0, so an EMPTY SUBTREE IS 0, not H(0,0) (merkle.rs:62-69); gen_proof takes node[id ^ 1]
//       per level (merkle.rs:97-100) and fails once the tree is full (merkle.rs:91-93).
//   hash = Poseidon-2  `hash_fix_len_array(&[left, right])`   shielder/relations/src/merkle_proof.rs:49-57
//   path_shape[i] = "current node is the left child at level i" (is_zero/select, merkle_proof.rs:53-55)
//
// Layout in HBM: one array of 2 * 2^depth Fr (Montgomery), zero-filled at creation -- "missing node = 0"
// is then simply what the array holds.  A node's two children are adjacent (ids 2i, 2i+1), so one level is a
// flat batch of hashes over contiguous 64-byte pairs: one thread per hash, whole Poseidon state in registers,
// round constants and MDS read through the uniform path (every lane reads the same address).
// Appending n leaves touches ids [lo, hi] per level (lo, hi halve each level): n + n/2 + ... hashes instead of
// the contract's n * depth.  The historical roots (one per inserted leaf, merkle.rs:78) are recomputed from the
// finished tree: the root after leaf k is the walk from k with final left siblings and 0 for right siblings.
#include <array>
#include <cstring>
#include <set>

#include "poseidon.cuh"

using namespace b200zk;
using b200zk::host::PoseidonConsts;

struct b200zk_merkle {
    uint32_t depth = 0;
    uint64_t size = 0;           // leaves = 2^depth
    uint64_t next_leaf_idx = 0;
    Fr* d_nodes = nullptr;       // 2 * size
    bool log_roots = false;
    std::set<std::array<uint8_t, 32>> roots_log;
    bool poisoned = false;       // a CUDA call failed in the middle of add_leaves: device nodes and host state disagree
};

namespace {

__device__ __forceinline__ Fr ldg_fr(const Fr* p) { return poseidon_dev::ld_fr(p); }
__device__ __forceinline__ void stg_fr(Fr* p, const Fr& r) { poseidon_dev::st_fr(p, r); }

// One thread, one permutation (t = 5, R_F = 8, R_P = 56, x^5, dense MDS every round).
__device__ __noinline__ void poseidon_permute_thread(Fr (&s)[5], const PoseidonConsts* __restrict__ pc) {
    constexpr int T = host::POSEIDON_T, half = host::POSEIDON_RF / 2;
#pragma unroll 1
    for (int rnd = 0; rnd < host::POSEIDON_ROUNDS; rnd++) {
        const bool full = rnd < half || rnd >= half + host::POSEIDON_RP;
#pragma unroll
        for (int j = 0; j < T; j++) s[j] = fp_add(s[j], ldg_fr(&pc->rc[rnd][j]));
        if (full) {
#pragma unroll
            for (int j = 0; j < T; j++) {
                const Fr x2 = fp_sqr(s[j]);
                s[j] = fp_mul(fp_sqr(x2), s[j]);
            }
        } else {
            const Fr x2 = fp_sqr(s[0]);
            s[0] = fp_mul(fp_sqr(x2), s[0]);
        }
        Fr t[T];
#pragma unroll
        for (int l = 0; l < T; l++) {
            t[l] = fp_mul(ldg_fr(&pc->mds[l][0]), s[0]);
#pragma unroll
            for (int j = 1; j < T; j++) t[l] = fp_add(t[l], fp_mul(ldg_fr(&pc->mds[l][j]), s[j]));
        }
#pragma unroll
        for (int l = 0; l < T; l++) s[l] = t[l];
    }
}

// hash_fix_len_array(&[left, right]): state = [2^64, left, right, 1 (padding), 0], one permutation, out = state[1]
__device__ __forceinline__ Fr poseidon2_thread(const Fr& left, const Fr& right, const PoseidonConsts* __restrict__ pc) {
    Fr s[5];
    s[0] = ldg_fr(&pc->two64);
    s[1] = left;
    s[2] = right;
    s[3] = Fr::one();
    s[4] = Fr::zero();
    poseidon_permute_thread(s, pc);
    return s[1];
}

// nodes[id] = H(nodes[2 id], nodes[2 id + 1]) for id in [first_id, first_id + count)
__global__ void __launch_bounds__(128) merkle_level_kernel(Fr* __restrict__ nodes, uint64_t first_id, uint64_t count,
                                                          const PoseidonConsts* __restrict__ pc) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t id = first_id + i;
    const Fr l = ldg_fr(nodes + 2 * id), r = ldg_fr(nodes + 2 * id + 1);
    stg_fr(nodes + id, poseidon2_thread(l, r, pc));
}

// ---- latency path: one WARP per hash (the depth-optimised permutation of poseidon.cuh, ~8x shorter than the
// one-thread form), for levels too small to fill the machine with threads and for single-leaf appends.
__device__ __forceinline__ Fr poseidon2_warp(const Fr& left, const Fr& right, int lane, const PoseidonConsts* pc) {
    return poseidon_dev::poseidon_hash_w(2, [&](int k) { return k == 0 ? left : right; }, lane, pc, nullptr);
}

// one level, one warp per node
__global__ void __launch_bounds__(128) merkle_level_warp_kernel(Fr* __restrict__ nodes, uint64_t first_id, uint64_t count,
                                                               const PoseidonConsts* __restrict__ pc) {
    const uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= count) return;  // whole warps leave together
    const int lane = threadIdx.x & 31;
    const uint64_t id = first_id + i;
    const Fr h = poseidon2_warp(ldg_fr(nodes + 2 * id), ldg_fr(nodes + 2 * id + 1), lane, pc);
    if (lane == 0) stg_fr(nodes + id, h);
}

// The top of the tree in one launch: levels with at most 32 dirty nodes, one CTA of 32 warps, a barrier per level
// (a chain of tiny launches would cost more than the hashing).
__global__ void __launch_bounds__(1024) merkle_top_kernel(Fr* __restrict__ nodes, uint64_t lo, uint64_t hi,
                                                         const PoseidonConsts* __restrict__ pc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (true) {
        const uint64_t id = lo + warp;
        if (id <= hi) {
            const Fr h = poseidon2_warp(ldg_fr(nodes + 2 * id), ldg_fr(nodes + 2 * id + 1), lane, pc);
            if (lane == 0) stg_fr(nodes + id, h);
        }
        if (lo == 1) break;
        lo >>= 1;
        hi >>= 1;
        __syncthreads();
    }
}

// historical roots, one warp per inserted leaf (small batches)
__global__ void __launch_bounds__(128) merkle_roots_warp_kernel(const Fr* __restrict__ nodes, uint64_t size, uint64_t first,
                                                               uint64_t n, uint32_t depth,
                                                               const PoseidonConsts* __restrict__ pc, Fr* __restrict__ roots) {
    const uint64_t k = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= n) return;
    const int lane = threadIdx.x & 31;
    uint64_t id = size + first + k;
    Fr cur = ldg_fr(nodes + id);
    for (uint32_t lvl = 0; lvl < depth; lvl++) {
        if (id & 1) cur = poseidon2_warp(ldg_fr(nodes + (id ^ 1)), cur, lane, pc);
        else cur = poseidon2_warp(cur, Fr::zero(), lane, pc);
        id >>= 1;
    }
    if (lane == 0) stg_fr(roots + k, cur);
}

// root after the insertion of leaf (first + k), k < n: walk up with the finished left siblings and 0 on the right
__global__ void __launch_bounds__(128) merkle_roots_kernel(const Fr* __restrict__ nodes, uint64_t size, uint64_t first,
                                                          uint64_t n, uint32_t depth, const PoseidonConsts* __restrict__ pc,
                                                          Fr* __restrict__ roots) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t id = size + first + k;
    Fr cur = ldg_fr(nodes + id);
    for (uint32_t lvl = 0; lvl < depth; lvl++) {
        if (id & 1) cur = poseidon2_thread(ldg_fr(nodes + (id ^ 1)), cur, pc);
        else cur = poseidon2_thread(cur, Fr::zero(), pc);
        id >>= 1;
    }
    stg_fr(roots + k, cur);
}

// gen_proof for many leaves: path[k][lvl] = nodes[id ^ 1]; shape = (id even).  Output either as a packed
// (path Fr[n][depth], shape u8[n][depth]) pair, or straight into update-note input rows (shape as Fr 0/1).
__global__ void merkle_paths_kernel(const Fr* __restrict__ nodes, uint64_t size, const uint64_t* __restrict__ leaf_ids,
                                    uint64_t n, uint32_t depth, Fr* __restrict__ path_out, uint64_t path_stride,
                                    uint8_t* __restrict__ shape_u8, Fr* __restrict__ shape_fr, uint64_t shape_stride,
                                    Fr* __restrict__ root_rows) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * depth) return;
    const uint64_t k = t / depth;
    const uint32_t lvl = (uint32_t)(t % depth);
    const uint64_t id = (size + (leaf_ids[k] & (size - 1))) >> lvl;  // ids are < size by contract; masked for memory safety
    if (root_rows && lvl == 0) stg_fr(root_rows + k * shape_stride, ldg_fr(nodes + 1));
    stg_fr(path_out + k * path_stride + lvl, ldg_fr(nodes + (id ^ 1)));
    const bool left = (id & 1) == 0;
    if (shape_u8) shape_u8[k * depth + lvl] = left ? 1 : 0;
    if (shape_fr) stg_fr(shape_fr + k * shape_stride + lvl, left ? Fr::one() : Fr::zero());
}

// Up to this many hashes a launch runs one warp per hash (latency), above it one thread per hash (throughput):
// 148 SMs x 32 resident warps of the warp-wide permutation.
constexpr uint64_t WARP_PATH_MAX = 4096;

int rebuild_levels(b200zk_ctx* ctx, b200zk_merkle* t, uint64_t first, uint64_t n, const PoseidonConsts* pc) {
    uint64_t lo = (t->size + first) >> 1, hi = (t->size + first + n - 1) >> 1;
    ProfScope ps(ctx, "merkle_levels");
    for (uint32_t lvl = 0; lvl < t->depth; lvl++) {
        const uint64_t count = hi - lo + 1;
        if (count <= 32) {
            merkle_top_kernel<<<1, 1024, 0, ctx->stream>>>(t->d_nodes, lo, hi, pc);
            B200ZK_TRY(check_launch(ctx, "merkle_top_kernel"));
            break;
        }
        if (count <= WARP_PATH_MAX) {
            merkle_level_warp_kernel<<<div_up(count, 4), 128, 0, ctx->stream>>>(t->d_nodes, lo, count, pc);
            B200ZK_TRY(check_launch(ctx, "merkle_level_warp_kernel"));
        } else {
            merkle_level_kernel<<<div_up(count, 128), 128, 0, ctx->stream>>>(t->d_nodes, lo, count, pc);
            B200ZK_TRY(check_launch(ctx, "merkle_level_kernel"));
        }
        lo >>= 1;
        hi >>= 1;
    }
    return B200ZK_OK;
}

int check_ids(b200zk_ctx* ctx, const b200zk_merkle* t, const uint64_t* ids, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (ids[i] >= t->size) return fail(ctx, B200ZK_ERR_BAD_ARG, "leaf id outside the tree");
    return B200ZK_OK;
}

}  // namespace

extern "C" {

int b200zk_merkle_new(b200zk_ctx* ctx, uint32_t depth, int log_roots, b200zk_merkle** out) {
    if (!ctx || !out || depth == 0 || depth > 31) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    b200zk_merkle* t = new b200zk_merkle();
    t->depth = depth;
    t->size = 1ull << depth;
    t->log_roots = log_roots != 0;
    const size_t bytes = 2 * t->size * sizeof(Fr);
    cudaError_t e = cudaMalloc(&t->d_nodes, bytes);
    if (e != cudaSuccess) {
        delete t;
        return fail(ctx, B200ZK_ERR_CUDA, std::string("merkle nodes: ") + cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(t->d_nodes, 0, bytes, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(t->d_nodes);
        delete t;
        return fail(ctx, B200ZK_ERR_CUDA, std::string("merkle memset: ") + cudaGetErrorString(e));
    }
    *out = t;
    return B200ZK_OK;
}

void b200zk_merkle_free(b200zk_ctx* ctx, b200zk_merkle* t) {
    if (!t) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (t->d_nodes) cudaFree(t->d_nodes);
    delete t;
}

int b200zk_merkle_info(const b200zk_merkle* t, uint32_t* depth, uint64_t* size, uint64_t* next_leaf_idx) {
    if (!t) return B200ZK_ERR_BAD_ARG;
    if (depth) *depth = t->depth;
    if (size) *size = t->size;
    if (next_leaf_idx) *next_leaf_idx = t->next_leaf_idx;
    return B200ZK_OK;
}

// the part of add_leaves that mutates the tree; any failure in here leaves it inconsistent (caller poisons it)
static int add_leaves_mutate(b200zk_ctx* ctx, b200zk_merkle* t, const void* leaves, int on_device, size_t n,
                             const PoseidonConsts* pc, void* d_roots, uint8_t* roots_dst) {
    const uint64_t first = t->next_leaf_idx;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(t->d_nodes + t->size + first, leaves, n * sizeof(Fr),
                                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(rebuild_levels(ctx, t, first, n, pc));
    if (d_roots) {
        {
            ProfScope ps(ctx, "merkle_roots");
            if (n <= WARP_PATH_MAX)
                merkle_roots_warp_kernel<<<div_up(n, 4), 128, 0, ctx->stream>>>(t->d_nodes, t->size, first, n, t->depth, pc,
                                                                                (Fr*)d_roots);
            else
                merkle_roots_kernel<<<div_up(n, 128), 128, 0, ctx->stream>>>(t->d_nodes, t->size, first, n, t->depth, pc,
                                                                             (Fr*)d_roots);
        }
        B200ZK_TRY(check_launch(ctx, "merkle_roots_kernel"));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(roots_dst, d_roots, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // also: the host buffer `leaves` is the caller's again
    return B200ZK_OK;
}

int b200zk_merkle_add_leaves(b200zk_ctx* ctx, b200zk_merkle* t, const void* leaves, int on_device, size_t n,
                             uint64_t* first_leaf_id, uint8_t* roots_out) {
    if (!ctx || !t || (!leaves && n)) return B200ZK_ERR_BAD_ARG;
    if (t->poisoned) return fail(ctx, B200ZK_ERR_CUDA, "merkle tree unusable: an earlier add_leaves failed part-way");
    if (first_leaf_id) *first_leaf_id = t->next_leaf_idx;
    if (n == 0) return B200ZK_OK;
    if (n > t->size - t->next_leaf_idx)                                   // merkle.rs:49-51
        return fail(ctx, B200ZK_ERR_MERKLE_LIMIT_EXCEEDED, "MerkleTreeLimitExceeded: the leaves do not fit (none was added)");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    // "all or nothing": everything that can fail without the device being lost (constant upload, scratch and
    // host allocations) happens BEFORE the first write to the tree
    const PoseidonConsts* pc;
    B200ZK_TRY(poseidon_consts_device(ctx, &pc));
    const bool want_roots = t->log_roots || roots_out;
    void* d_roots = nullptr;
    std::vector<uint8_t> tmp;
    uint8_t* dst = roots_out;
    if (want_roots) {
        B200ZK_TRY(scratch(ctx, "merkle_roots", n * sizeof(Fr), &d_roots));
        if (!dst) {
            try {
                tmp.resize(n * 32);
            } catch (...) {
                return fail(ctx, B200ZK_ERR_BAD_LEN, "merkle_add_leaves: out of host memory for the roots");
            }
            dst = tmp.data();
        }
    }
    const int rc = add_leaves_mutate(ctx, t, leaves, on_device, n, pc, d_roots, dst);
    if (rc != B200ZK_OK) {  // a launch / copy failed on a tree already written to: the CUDA context is gone or the
        t->poisoned = true; // nodes are half updated; refuse further use instead of serving a wrong root
        return rc;
    }
    t->next_leaf_idx += n;
    if (t->log_roots)
        for (size_t i = 0; i < n; i++) {
            std::array<uint8_t, 32> r;
            memcpy(r.data(), dst + i * 32, 32);
            t->roots_log.insert(r);
        }
    return B200ZK_OK;
}

int b200zk_merkle_root(b200zk_ctx* ctx, const b200zk_merkle* t, uint8_t out[32]) {
    if (!ctx || !t || !out) return B200ZK_ERR_BAD_ARG;
    if (t->poisoned) return fail(ctx, B200ZK_ERR_CUDA, "merkle tree unusable: an earlier add_leaves failed part-way");
    if (t->next_leaf_idx == 0)                                            // merkle.rs:42-46: node 1 was never written
        return fail(ctx, B200ZK_ERR_MERKLE_NON_EXISTING_NODE, "MerkleTreeNonExistingNode: the tree is empty");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, t->d_nodes + 1, 32, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_merkle_node(b200zk_ctx* ctx, const b200zk_merkle* t, uint64_t id, uint8_t out[32]) {
    if (!ctx || !t || !out || id == 0 || id >= 2 * t->size) return B200ZK_ERR_BAD_ARG;
    if (t->poisoned) return fail(ctx, B200ZK_ERR_CUDA, "merkle tree unusable: an earlier add_leaves failed part-way");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, t->d_nodes + id, 32, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_merkle_is_historical_root(b200zk_ctx* ctx, const b200zk_merkle* t, const uint8_t root[32], int* out) {
    if (!ctx || !t || !root || !out) return B200ZK_ERR_BAD_ARG;
    if (!t->log_roots) return fail(ctx, B200ZK_ERR_BAD_ARG, "the tree was created without log_roots");
    std::array<uint8_t, 32> r;
    memcpy(r.data(), root, 32);
    *out = t->roots_log.count(r) ? 1 : 0;
    return B200ZK_OK;
}

int b200zk_merkle_gen_proofs(b200zk_ctx* ctx, const b200zk_merkle* t, const uint64_t* leaf_ids, size_t n,
                             uint8_t* path_out, uint8_t* shape_out) {
    if (!ctx || !t || (!leaf_ids && n) || (!path_out && n)) return B200ZK_ERR_BAD_ARG;
    if (t->next_leaf_idx == t->size)                                      // merkle.rs:91-93
        return fail(ctx, B200ZK_ERR_MERKLE_PROOF_GEN_FAIL, "MerkleTreeProofGenFail: the tree is full");
    if (n == 0) return B200ZK_OK;
    B200ZK_TRY(check_ids(ctx, t, leaf_ids, n));
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    void *d_ids, *d_path, *d_shape;
    B200ZK_TRY(scratch(ctx, "merkle_ids", n * 8, &d_ids));
    B200ZK_TRY(scratch(ctx, "merkle_path", n * t->depth * sizeof(Fr), &d_path));
    B200ZK_TRY(scratch(ctx, "merkle_shape", n * t->depth, &d_shape));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(d_ids, leaf_ids, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    merkle_paths_kernel<<<div_up(n * t->depth, 256), 256, 0, ctx->stream>>>(t->d_nodes, t->size, (const uint64_t*)d_ids, n,
                                                                            t->depth, (Fr*)d_path, t->depth,
                                                                            (uint8_t*)d_shape, nullptr, 0, nullptr);
    B200ZK_TRY(check_launch(ctx, "merkle_paths_kernel"));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(path_out, d_path, n * t->depth * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    if (shape_out) B200ZK_CUDA(ctx, cudaMemcpyAsync(shape_out, d_shape, n * t->depth, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_merkle_fill_update_note_inputs_device(b200zk_ctx* ctx, const b200zk_merkle* t, const void* d_leaf_ids, size_t n,
                                                 void* d_inputs) {
    if (!ctx || !t || (!d_leaf_ids && n) || (!d_inputs && n)) return B200ZK_ERR_BAD_ARG;
    if (t->next_leaf_idx == t->size)
        return fail(ctx, B200ZK_ERR_MERKLE_PROOF_GEN_FAIL, "MerkleTreeProofGenFail: the tree is full");
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    // input row = 18 + 2H Fr: ... | path_shape[H] at 13 | path[H] at 13 + H | ...   (include/b200zk.h, K6)
    const uint64_t stride = 18 + 2 * (uint64_t)t->depth;
    Fr* rows = (Fr*)d_inputs;
    merkle_paths_kernel<<<div_up(n * t->depth, 256), 256, 0, ctx->stream>>>(
        t->d_nodes, t->size, (const uint64_t*)d_leaf_ids, n, t->depth, rows + 13 + t->depth, stride, nullptr, rows + 13, stride,
        rows + 4);  // slot 4 = merkle_root: the current root
    return check_launch(ctx, "merkle_paths_kernel");
}

}  // extern "C"
