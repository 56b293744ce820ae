// K6 -- batched Poseidon (t=5, alpha=5, R_F=8, R_P=56) and the shielder "update note" witness
// generator, plus the host-side R1CS export of the same relation.
//
// Spec: /root/reference/shielder/relations/src/relations/update_note.rs:106-149 (statement),
// merkle_proof.rs:38-61 (path walk), lib.rs:17-26 (Poseidon parameters); the permutation itself
// lives in the un-vendored halo2-base 0.4.1 (PARITY UNPINNED, see oracle/pyref/poseidon.py).
//
// Mapping.  Witness generator: two warps per proof (the Merkle-path chain on one, the note/account
// hashes and the balance update on the other), each running a warp-wide Poseidon laid out for depth
// (25 lanes = 5x5 MDS products, see poseidon_permute_w).  Batched hashing (poseidon_hash_batch): one
// hash per 8-lane group, 5 lanes hold the state, laid out for throughput.
// Every S-box writes its (x^2, x^4, x^5) straight into the proof's assignment vector z, in the
// order host_r1cs.hpp allocates them, so the R1CS witness is a by-product of hashing.
#include <cstring>
#include <memory>

#include "poseidon.cuh"

using namespace b200zk;
using namespace b200zk::poseidon_dev;
using b200zk::host::PoseidonConsts;

namespace {

constexpr int GROUP = 8;  // lanes per proof

// word `src` (0..7) of this lane's group
__device__ __forceinline__ Fr group_bcast(const Fr& v, int src) {
    const int lane = threadIdx.x & 31;
    const int from = (lane & ~(GROUP - 1)) + src;
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], from);
    return r;
}

// Cooperative permutation: lane l (< 5) holds state word l.  trace (may be null) receives the
// 96 * 3 S-box witnesses.  All 32 lanes of the warp must call this together.
__device__ void poseidon_permute(Fr& s, int l, const PoseidonConsts* pc, Fr* trace) {
    const int half = host::POSEIDON_RF / 2;
    int sbox = 0;
    for (int rnd = 0; rnd < host::POSEIDON_ROUNDS; rnd++) {
        const bool full = rnd < half || rnd >= half + host::POSEIDON_RP;
        if (l < host::POSEIDON_T) {
            s = fp_add(s, ld_fr(&pc->rc[rnd][l]));
            if (full || l == 0) {
                Fr x2 = fp_sqr(s);
                Fr x4 = fp_sqr(x2);
                Fr x5 = fp_mul(x4, s);
                if (trace) {
                    Fr* t = trace + (size_t)(sbox + (full ? l : 0)) * 3;
                    st_fr(t, x2);
                    st_fr(t + 1, x4);
                    st_fr(t + 2, x5);
                }
                s = x5;
            }
        }
        sbox += full ? host::POSEIDON_T : 1;
        Fr acc = Fr::zero();
#pragma unroll 1
        for (int j = 0; j < host::POSEIDON_T; j++) {
            Fr sj = group_bcast(s, j);
            if (l < host::POSEIDON_T) acc = fp_add(acc, fp_mul(ld_fr(&pc->mds[l][j]), sj));
        }
        s = acc;
    }
}

// hash_fix_len_array of `n_in` inputs (n_in in {2, 4} here, any n works): lane l in 1..4 supplies
// input (chunk*4 + l - 1) through `in_of(lane_input_index)`.  Returns the digest in every lane.
template <class GetIn>
__device__ Fr poseidon_hash(int n_in, GetIn in_of, int l, const PoseidonConsts* pc, Fr* trace) {
    Fr s = Fr::zero();
    if (l == 0) s = ld_fr(&pc->two64);
    const int chunks = n_in / host::POSEIDON_RATE + 1;
    for (int c = 0; c < chunks; c++) {
        const int lo = c * host::POSEIDON_RATE;
        const int len = lo < n_in ? min(host::POSEIDON_RATE, n_in - lo) : 0;
        if (l >= 1 && l <= len) s = fp_add(s, in_of(lo + l - 1));
        if (len + 1 < host::POSEIDON_T && l == len + 1) s = fp_add(s, Fr::one());
        poseidon_permute(s, l, pc, trace ? trace + (size_t)c * host::POSEIDON_TRACE : nullptr);
    }
    return group_bcast(s, 1);
}

__device__ __forceinline__ bool fits_bits128(const Fr& canon) {
    return (canon.v[4] | canon.v[5] | canon.v[6] | canon.v[7]) == 0;
}

// every input word of an instance must be a field element (< r): ark's Fr cannot hold anything else, and the
// inversions below must never see an unreduced word.  All threads of the CTA call this.
__device__ __forceinline__ bool inputs_canonical(const Fr* in, uint32_t n_in) {
    bool canon = true;
    for (uint32_t i = threadIdx.x; i < n_in; i += blockDim.x) {
        const Fr v = ld_fr(in + i);
        bool lt = false;  // v < r, most significant limb first
        for (int k = 7; k >= 0; k--) {
            if (v.v[k] != FrCfg::mod(k)) {
                lt = v.v[k] < FrCfg::mod(k);
                break;
            }
        }
        canon = canon && lt;
    }
    return __syncthreads_and(canon ? 1 : 0) != 0;
}

// Witness block of g_update_account (host_r1cs.hpp), warp-wide: H(old_account) == old_hash, the bits of `amount`,
// per token slot (inverse-or-0, is_equal, delta, bits of the new balance), H(new_account) == new_hash.
// acc = (token0, balance0, token1, balance1); zb = first witness of the block.  Returns "all checks hold".
__device__ bool account_update_witness(const Fr* __restrict__ acc, const Fr& amount, const Fr& token, const Fr& old_hash,
                                       const Fr& new_hash, int kind, int lane, const PoseidonConsts* pc, Fr* zb) {
    const uint32_t T = host::POSEIDON_TRACE;
    Fr* z_old = zb;
    Fr* z_bits = zb + 2 * T;
    Fr* z_upd = z_bits + host::BALANCE_BITS;
    Fr* z_new = z_upd + 2 * (3 + host::BALANCE_BITS);
    const bool wr = lane == 0;
    bool ok = true;
    // ---- H(old_account) == old_hash                                     update_account.rs:79-85
    Fr h_acc = poseidon_hash_w(4, [&](int i) { return ld_fr(acc + i); }, lane, pc, z_old);
    ok = ok && (h_acc == old_hash);
    // ---- account update                                                  account.rs:36-79 (mock)
    {
        const Fr canon = fp_from_mont(amount);
        ok = ok && fits_bits128(canon);
        for (int b = lane; b < host::BALANCE_BITS; b += 32)
            st_fr(z_bits + b, ((canon.v[b >> 5] >> (b & 31)) & 1) ? Fr::one() : Fr::zero());
    }
    Fr new_bal[2];
    int n_match = 0;
    for (int i = 0; i < 2; i++) {
        Fr* zu = z_upd + (size_t)i * (3 + host::BALANCE_BITS);
        const Fr tok = ld_fr(acc + 2 * i), bal = ld_fr(acc + 2 * i + 1);
        const Fr diff = fp_sub(tok, token);
        const bool eq = diff.is_zero();
        n_match += eq ? 1 : 0;
        const Fr delta = eq ? amount : Fr::zero();
        if (wr) {
            st_fr(zu, eq ? Fr::zero() : fp_inv(diff));
            st_fr(zu + 1, eq ? Fr::one() : Fr::zero());
            st_fr(zu + 2, delta);
        }
        new_bal[i] = kind == host::KIND_DEPOSIT ? fp_add(bal, delta) : fp_sub(bal, delta);
        const Fr canon = fp_from_mont(new_bal[i]);
        ok = ok && fits_bits128(canon);                                // checked_add / checked_sub
        for (int b = lane; b < host::BALANCE_BITS; b += 32)
            st_fr(zu + 3 + b, ((canon.v[b >> 5] >> (b & 31)) & 1) ? Fr::one() : Fr::zero());
    }
    ok = ok && n_match == 1;
    // ---- H(new_account) == new_hash                                      update_account.rs:88-94
    Fr h_nacc = poseidon_hash_w(4, [&](int i) { return (i & 1) ? new_bal[i >> 1] : ld_fr(acc + i); }, lane, pc, z_new);
    ok = ok && (h_nacc == new_hash);
    return ok;
}

// inputs per proof, each a Montgomery Fr, in UpdateNoteInput::new argument order:
//   op_pub (amount, token, user) | new_note_hash | merkle_root | new_note[4] | old_note[4] |
//   path_shape[H] | path[H] | op_priv.user | old_account (token0, balance0, token1, balance1)
// One CTA of two warps per proof, split along the data dependencies of the statement:
//   warp 0: loaded witnesses, H(old_note), the Merkle path (the long chain: 2 + H permutations)
//   warp 1: H(new_note), H(old_account), the balance update and its range bits, H(new_account) (6 permutations)
__global__ void __launch_bounds__(64) update_note_witness_kernel(const Fr* __restrict__ inputs, uint32_t n_proofs,
                                                                uint32_t H, int kind, uint32_t num_vars,
                                                                const PoseidonConsts* __restrict__ pc,
                                                                Fr* __restrict__ z_all, uint32_t* __restrict__ status) {
    __shared__ uint32_t ok_sh[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t proof = blockIdx.x;
    if (proof >= n_proofs) return;
    const uint32_t n_in = 18 + 2 * H;
    const Fr* in = inputs + (size_t)proof * n_in;
    Fr* z = z_all + (size_t)proof * num_vars;
    // input slots
    const int I_AMOUNT = 0, I_TOKEN = 1, I_USER = 2, I_NNH = 3, I_ROOT = 4, I_NEW = 5, I_OLD = 9;
    const int I_SHAPE = 13, I_PATH = 13 + H, I_PRIV = 13 + 2 * H, I_ACC = 14 + 2 * H;
    // assignment layout (the allocation order of host_r1cs.hpp)
    const uint32_t T = host::POSEIDON_TRACE;
    const uint32_t n_copy = 4 + 3 + 2 * H + 1 + 4;
    const uint32_t O_COPY = 7, O_NEW_HASH = O_COPY + n_copy, O_OLD_HASH = O_NEW_HASH + 2 * T, O_PATH = O_OLD_HASH + 2 * T;
    const uint32_t O_ACC_HASH = O_PATH + H * (4 + T), O_AMOUNT_BITS = O_ACC_HASH + 2 * T;
    const uint32_t O_UPDATE = O_AMOUNT_BITS + host::BALANCE_BITS, O_NACC_HASH = O_UPDATE + 2 * (3 + host::BALANCE_BITS);
    const uint32_t O_END = O_NACC_HASH + 2 * T;
    const bool wr = lane == 0;  // the lane that writes scalar witnesses
    bool ok = true;
    // Input rows come from the caller unchecked: an instance holding a word >= r is reported as unsatisfied and nothing
    // below (inversions included) ever sees it.
    if (!inputs_canonical(in, n_in)) {
        if (threadIdx.x == 0) status[proof] = 1u | (O_END == num_vars ? 0u : 2u);
        return;
    }
    if (warp == 0) {
        // ---- z[0] = 1, instance variables, loaded witnesses
        if (wr) {
            st_fr(z + 0, Fr::one());
            st_fr(z + 1, ld_fr(in + I_AMOUNT));
            st_fr(z + 2, ld_fr(in + I_TOKEN));
            st_fr(z + 3, ld_fr(in + I_USER));
            st_fr(z + 4, ld_fr(in + I_NNH));
            st_fr(z + 5, ld_fr(in + I_ROOT));
            st_fr(z + 6, ld_fr(in + I_OLD + 2));  // old_note.nullifier
        }
        // new_note[4], old zk_id, old trapdoor, old account_hash, path_shape[H], path[H], op_priv, account[4]
        for (uint32_t i = lane; i < n_copy; i += 32) {
            uint32_t src;
            if (i < 4) src = I_NEW + i;
            else if (i < 7) src = I_OLD + (i == 4 ? 0 : i == 5 ? 1 : 3);
            else src = I_SHAPE + (i - 7);  // shape, path, op_priv, account are contiguous in the input
            st_fr(z + O_COPY + i, ld_fr(in + src));
        }
        // ---- old_note_hash                                               update_note.rs:131
        Fr current = poseidon_hash_w(4, [&](int i) { return ld_fr(in + I_OLD + i); }, lane, pc, z + O_OLD_HASH);
        // ---- Merkle path                                                 merkle_proof.rs:49-57
        for (uint32_t i = 0; i < H; i++) {
            Fr* zl = z + O_PATH + (size_t)i * (4 + T);
            const Fr shape = ld_fr(in + I_SHAPE + i);
            const Fr sib = ld_fr(in + I_PATH + i);
            const bool sel = shape.is_zero();  // selector = is_zero(shape)
            // left = select(sibling, current, selector) ; right = select(current, sibling, selector)
            const Fr t1 = sel ? fp_sub(sib, current) : Fr::zero();
            const Fr t2 = sel ? fp_sub(current, sib) : Fr::zero();
            if (wr) {
                Fr inv = sel ? Fr::zero() : (shape == Fr::one() ? Fr::one() : fp_inv(shape));
                st_fr(zl, inv);
                st_fr(zl + 1, sel ? Fr::one() : Fr::zero());
                st_fr(zl + 2, t1);
                st_fr(zl + 3, t2);
            }
            const Fr left = fp_add(t1, current), right = fp_add(t2, sib);
            current = poseidon_hash_w(2, [&](int k) { return k == 0 ? left : right; }, lane, pc, zl + 4);
        }
        ok = ok && (current == ld_fr(in + I_ROOT));                        // merkle_proof.rs:59-60
        ok = ok && (ld_fr(in + I_USER) == ld_fr(in + I_PRIV));             // combine(): ops.rs:47-62
    } else {
        // ---- H(new_note) == new_note_hash                               update_note.rs:129
        Fr h_new = poseidon_hash_w(4, [&](int i) { return ld_fr(in + I_NEW + i); }, lane, pc, z + O_NEW_HASH);
        ok = ok && (h_new == ld_fr(in + I_NNH));
        // ---- update_account_circuit                                      update_note.rs:141-148
        const bool acc_ok = account_update_witness(in + I_ACC, ld_fr(in + I_AMOUNT), ld_fr(in + I_TOKEN), ld_fr(in + I_OLD + 3),
                                                   ld_fr(in + I_NEW + 3), kind, lane, pc, z + O_ACC_HASH);
        ok = ok && acc_ok;
    }
    if (wr) ok_sh[warp] = ok ? 1u : 0u;
    __syncthreads();
    if (threadIdx.x == 0) status[proof] = ((ok_sh[0] & ok_sh[1]) ? 0u : 1u) | (O_END == num_vars ? 0u : 2u);
}

// update_account_circuit as a relation of its own (update_account.rs:68-95): one warp per instance.
// input row (9 Fr, UpdateAccountInput::new argument order :37-42): old_account_hash | new_account_hash |
// operation (amount, token, user) | old_account (token0, balance0, token1, balance1)
__global__ void __launch_bounds__(32) update_account_witness_kernel(const Fr* __restrict__ inputs, uint32_t n, int kind,
                                                                   uint32_t num_vars, const PoseidonConsts* __restrict__ pc,
                                                                   Fr* __restrict__ z_all, uint32_t* __restrict__ status) {
    const uint32_t inst = blockIdx.x;
    if (inst >= n) return;
    const int lane = threadIdx.x;
    const Fr* in = inputs + (size_t)inst * 9;
    Fr* z = z_all + (size_t)inst * num_vars;
    const uint32_t T = host::POSEIDON_TRACE;
    const uint32_t O_END = 10 + 2 * T + host::BALANCE_BITS + 2 * (3 + host::BALANCE_BITS) + 2 * T;
    if (!inputs_canonical(in, 9)) {
        if (lane == 0) status[inst] = 1u | (O_END == num_vars ? 0u : 2u);
        return;
    }
    if (lane == 0) st_fr(z, Fr::one());
    if (lane < 9) st_fr(z + 1 + lane, ld_fr(in + lane));   // 5 instance variables, then the 4 account witnesses
    const bool ok = account_update_witness(in + 5, ld_fr(in + 2), ld_fr(in + 3), ld_fr(in + 0), ld_fr(in + 1), kind, lane, pc, z + 10);
    if (lane == 0) status[inst] = (ok ? 0u : 1u) | (O_END == num_vars ? 0u : 2u);
}

// n_hashes independent hash_fix_len_array calls of the same arity (Merkle tree levels, note hashes)
__global__ void __launch_bounds__(32) poseidon_hash_batch_kernel(const Fr* __restrict__ in, size_t n_hashes,
                                                                uint32_t arity,
                                                                const PoseidonConsts* __restrict__ pc,
                                                                Fr* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int l = lane & (GROUP - 1);
    const size_t raw = (size_t)blockIdx.x * (32 / GROUP) + (lane / GROUP);
    const bool active = raw < n_hashes;
    const size_t idx = active ? raw : n_hashes - 1;
    const Fr* src = in + idx * arity;
    Fr h = poseidon_hash((int)arity, [&](int i) { return ld_fr(src + i); }, l, pc, nullptr);
    if (active && l == 0) st_fr(out + idx, h);
}

}  // namespace

namespace b200zk {

int poseidon_consts_device(b200zk_ctx* ctx, const PoseidonConsts** out) {
    if (!ctx->poseidon_consts) {
        const PoseidonConsts& pc = host::poseidon_consts();
        B200ZK_CUDA(ctx, cudaMalloc(&ctx->poseidon_consts, sizeof(PoseidonConsts)));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(ctx->poseidon_consts, &pc, sizeof(pc), cudaMemcpyHostToDevice, ctx->stream));
        B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out = (const PoseidonConsts*)ctx->poseidon_consts;
    return B200ZK_OK;
}

// d_inputs: batch * (18 + 2H) Fr; d_z: batch * num_vars Fr; d_status: batch u32
int update_note_witness_device(b200zk_ctx* ctx, int kind, uint32_t H, uint32_t num_vars, const Fr* d_inputs,
                               size_t batch, Fr* d_z, uint32_t* d_status) {
    const PoseidonConsts* pc;
    B200ZK_TRY(poseidon_consts_device(ctx, &pc));
    {
        ProfScope ps(ctx, "witness");
        update_note_witness_kernel<<<(unsigned)batch, 64, 0, ctx->stream>>>(d_inputs, (uint32_t)batch, H, kind, num_vars, pc,
                                                                            d_z, d_status);
    }
    return check_launch(ctx, "update_note_witness_kernel");
}

// witness generation of either relation; d_inputs: batch * inputs_per_instance Fr
int relation_witness_device(b200zk_ctx* ctx, int relation, int kind, uint32_t H, uint32_t num_vars, const Fr* d_inputs,
                            size_t batch, Fr* d_z, uint32_t* d_status) {
    if (relation == host::RELATION_UPDATE_NOTE) return update_note_witness_device(ctx, kind, H, num_vars, d_inputs, batch, d_z, d_status);
    const PoseidonConsts* pc;
    B200ZK_TRY(poseidon_consts_device(ctx, &pc));
    {
        ProfScope ps(ctx, "witness");
        update_account_witness_kernel<<<(unsigned)batch, 32, 0, ctx->stream>>>(d_inputs, (uint32_t)batch, kind, num_vars, pc, d_z,
                                                                               d_status);
    }
    return check_launch(ctx, "update_account_witness_kernel");
}

}  // namespace b200zk

extern "C" {

int b200zk_poseidon_constants(uint8_t* round_constants, uint8_t* mds) {
    const PoseidonConsts& pc = host::poseidon_consts();
    if (round_constants) memcpy(round_constants, pc.rc, sizeof(pc.rc));
    if (mds) memcpy(mds, pc.mds, sizeof(pc.mds));
    return B200ZK_OK;
}

int b200zk_update_note_r1cs(int kind, uint32_t tree_height, b200zk_r1cs** out) {
    if (!out || (kind != 0 && kind != 1) || tree_height == 0 || tree_height > 64) return B200ZK_ERR_BAD_ARG;
    b200zk_r1cs* r = new b200zk_r1cs();
    r->cs = host::synthesize_update_note(kind, tree_height);
    *out = r;
    return B200ZK_OK;
}

int b200zk_update_account_r1cs(int kind, b200zk_r1cs** out) {
    if (!out || (kind != 0 && kind != 1)) return B200ZK_ERR_BAD_ARG;
    b200zk_r1cs* r = new b200zk_r1cs();
    r->cs = host::synthesize_update_account(kind);
    *out = r;
    return B200ZK_OK;
}

void b200zk_r1cs_free(b200zk_r1cs* r) { delete r; }

int b200zk_r1cs_shape(const b200zk_r1cs* r, uint64_t* num_constraints, uint64_t* num_inputs, uint64_t* num_aux,
                      uint64_t nnz[3]) {
    if (!r) return B200ZK_ERR_BAD_ARG;
    if (num_constraints) *num_constraints = r->cs.A.size();
    if (num_inputs) *num_inputs = r->cs.num_inputs;
    if (num_aux) *num_aux = r->cs.num_aux;
    if (nnz) {
        const std::vector<host::LC>* M[3] = {&r->cs.A, &r->cs.B, &r->cs.C};
        for (int m = 0; m < 3; m++) {
            uint64_t c = 0;
            for (auto& row : *M[m]) c += row.t.size();
            nnz[m] = c;
        }
    }
    return B200ZK_OK;
}

int b200zk_r1cs_matrix(const b200zk_r1cs* r, int which, uint64_t* row_ptr, uint32_t* cols, uint8_t* vals) {
    if (!r || which < 0 || which > 2 || !row_ptr || !cols || !vals) return B200ZK_ERR_BAD_ARG;
    const std::vector<host::LC>& M = which == 0 ? r->cs.A : which == 1 ? r->cs.B : r->cs.C;
    uint64_t k = 0;
    for (size_t i = 0; i < M.size(); i++) {
        row_ptr[i] = k;
        for (auto& e : M[i].t) {
            cols[k] = e.first;
            memcpy(vals + k * 32, &e.second, 32);
            k++;
        }
    }
    row_ptr[M.size()] = k;
    return B200ZK_OK;
}

int b200zk_poseidon_hash_batch(b200zk_ctx* ctx, const uint8_t* inputs, size_t n_hashes, uint32_t arity, uint8_t* out) {
    if (!ctx || !inputs || !out || arity == 0 || arity > 64) return B200ZK_ERR_BAD_ARG;
    if (n_hashes == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const PoseidonConsts* pc;
    B200ZK_TRY(poseidon_consts_device(ctx, &pc));
    void *din, *dout;
    B200ZK_TRY(scratch(ctx, "pos_in", n_hashes * arity * 32, &din));
    B200ZK_TRY(scratch(ctx, "pos_out", n_hashes * 32, &dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(din, inputs, n_hashes * arity * 32, cudaMemcpyHostToDevice, ctx->stream));
    poseidon_hash_batch_kernel<<<div_up(n_hashes, 32 / GROUP), 32, 0, ctx->stream>>>((const Fr*)din, n_hashes, arity, pc,
                                                                                    (Fr*)dout);
    B200ZK_TRY(check_launch(ctx, "poseidon_hash_batch_kernel"));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, dout, n_hashes * 32, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

static int witness_batch_impl(b200zk_ctx* ctx, const b200zk_r1cs* r, int relation, const uint8_t* inputs, size_t batch,
                              uint8_t* out_assignments, void* d_out_assignments, uint8_t* out_status);

int b200zk_update_note_witness_batch(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t* inputs, size_t batch,
                                     uint8_t* out_assignments, void* d_out_assignments, uint8_t* out_status) {
    return witness_batch_impl(ctx, r, host::RELATION_UPDATE_NOTE, inputs, batch, out_assignments, d_out_assignments, out_status);
}

int b200zk_update_account_witness_batch(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t* inputs, size_t batch,
                                        uint8_t* out_assignments, void* d_out_assignments, uint8_t* out_status) {
    return witness_batch_impl(ctx, r, host::RELATION_UPDATE_ACCOUNT, inputs, batch, out_assignments, d_out_assignments, out_status);
}

static int witness_batch_impl(b200zk_ctx* ctx, const b200zk_r1cs* r, int relation, const uint8_t* inputs, size_t batch,
                              uint8_t* out_assignments, void* d_out_assignments, uint8_t* out_status) {
    if (!ctx || !r || !inputs) return B200ZK_ERR_BAD_ARG;
    if (r->cs.relation != relation) return fail(ctx, B200ZK_ERR_BAD_ARG, "the R1CS handle belongs to the other relation");
    if (batch == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t H = r->cs.tree_height, nv = r->cs.num_variables();
    const size_t in_bytes = batch * (size_t)r->cs.inputs_per_instance * 32, z_bytes = batch * (size_t)nv * 32;
    void *din, *dz = d_out_assignments, *dst;
    B200ZK_TRY(scratch(ctx, "wit_in", in_bytes, &din));
    if (!dz) B200ZK_TRY(scratch(ctx, "wit_z", z_bytes, &dz));
    B200ZK_TRY(scratch(ctx, "wit_status", batch * 4, &dst));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(din, inputs, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(relation_witness_device(ctx, relation, r->cs.kind, H, nv, (const Fr*)din, batch, (Fr*)dz, (uint32_t*)dst));
    std::vector<uint32_t> st(batch);
    B200ZK_CUDA(ctx, cudaMemcpyAsync(st.data(), dst, batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_assignments)
        B200ZK_CUDA(ctx, cudaMemcpyAsync(out_assignments, dz, z_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = B200ZK_OK;
    for (size_t i = 0; i < batch; i++) {
        if (out_status) out_status[i] = (uint8_t)st[i];
        if (st[i] & 2) return fail(ctx, B200ZK_ERR_BAD_ARG, "witness layout does not match the R1CS (internal error)");
        if (st[i] & 1) rc = B200ZK_ERR_UNSATISFIED;
    }
    if (rc != B200ZK_OK) fail(ctx, rc, "a witness does not satisfy the relation (see out_status)");
    return rc;
}

}  // extern "C"
