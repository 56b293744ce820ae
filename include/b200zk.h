/* b200zk.h -- C ABI of the B200-native Groth16 / BLS12-381 prover backend.
 *
 * This is the drop-in boundary a Rust `-sys` crate (or cgo/ctypes) binds.  The reference
 * tree (/root/reference, Cardinal-Cryptography/zk-apps @967d180) has NO prover-level API:
 * its only first-party interface on this path is the circuit builder
 *   update_note_circuit(ctx, UpdateNoteInput, make_public)
 *     -- shielder/relations/src/relations/update_note.rs:106-149
 * plus the Account/Operation traits (relations/src/account.rs:8-21, operation.rs:3-23).
 * The entry points below are therefore shaped after the arkworks 0.4 call sites
 * BASELINE.json names ([recall]): each function cites the arkworks
 * item it replaces and, where one exists, the reference file that defines the semantics.
 *
 * Conventions (all functions):
 *   - return 0 (B200ZK_OK) or a negative b200zk_status; b200zk_last_error(ctx) has the text;
 *   - the caller owns every buffer; the library never frees caller memory and keeps no host
 *     pointer after return; no exceptions cross the boundary;
 *   - a ctx is single-threaded and bound to one GPU (one stream); use one ctx per thread;
 *   - field elements are little-endian limbs in Montgomery form (a*R mod p, R = 2^256 for Fr,
 *     2^384 for Fq) = the in-memory words of ark-ff 0.4 `Fp<MontBackend>`; MSM scalars are
 *     canonical 256-bit little-endian integers = arkworks `BigInt<4>` (`into_bigint()`);
 *   - G1 affine = x||y (96 B), G2 affine = x.c0||x.c1||y.c0||y.c1 (192 B); the point at
 *     infinity is the all-zero encoding on input and output (plus an explicit flag where
 *     stated); outputs are affine and fully reduced so byte comparison is meaningful;
 *   - there is no CPU fallback: without a CUDA device b200zk_init fails.
 */
#ifndef B200ZK_H
#define B200ZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200zk_ctx b200zk_ctx;
typedef struct b200zk_bases b200zk_bases;   /* device-resident MSM bases (proving-key query) */
typedef struct b200zk_pk b200zk_pk;         /* device-resident Groth16 proving key + R1CS matrices */
typedef struct b200zk_vk b200zk_vk;         /* device-resident prepared verifying key */

typedef enum {
    B200ZK_OK = 0,
    B200ZK_ERR_BAD_ARG = -1,
    B200ZK_ERR_BAD_LEN = -2,          /* ark `msm` returns Err(len) on length mismatch */
    B200ZK_ERR_DOMAIN_TOO_LARGE = -3, /* ark `Radix2EvaluationDomain::new` returns None if log2 > 32 */
    B200ZK_ERR_CUDA = -4,
    B200ZK_ERR_NO_DEVICE = -5,
    B200ZK_ERR_UNSATISFIED = -6,      /* witness does not satisfy the relation (ark: SynthesisError::Unsatisfiable) */
    B200ZK_ERR_NOT_IMPLEMENTED = -7,
    B200ZK_ERR_MERKLE_LIMIT_EXCEEDED = -8,    /* ShielderError::MerkleTreeLimitExceeded   (contract/merkle.rs:49-51) */
    B200ZK_ERR_MERKLE_PROOF_GEN_FAIL = -9,    /* ShielderError::MerkleTreeProofGenFail    (contract/merkle.rs:91-93) */
    B200ZK_ERR_MERKLE_NON_EXISTING_NODE = -10,/* ShielderError::MerkleTreeNonExistingNode (contract/merkle.rs:42-46) */
    B200ZK_ERR_BAD_ENCODING = -11,            /* serialized key: truncated / invalid point (ark SerializationError::InvalidData) */
    B200ZK_ERR_NCCL = -12                     /* NCCL missing at run time, or a communicator / collective call failed */
} b200zk_status;

enum { B200ZK_FIELD_FR = 0, B200ZK_FIELD_FQ = 1, B200ZK_FIELD_FQ2 = 2 };
enum { B200ZK_OP_ADD = 0, B200ZK_OP_SUB = 1, B200ZK_OP_MUL = 2, B200ZK_OP_SQR = 3, B200ZK_OP_INV = 4,
       B200ZK_OP_TO_MONT = 5, B200ZK_OP_FROM_MONT = 6,
       B200ZK_OP_MUL_DFMA = 7 /* experiment: Fq product on the FP64 pipe (csrc/field_dfma.cuh); Fq only */ };

/* ---- context -------------------------------------------------------------------------
 * One ctx per GPU.  The deployment model is one host thread (or process) per GPU, each with its own ctx
 * (b200zk_init(device)); GPUs cooperate through a communicator attached to the ctx (b200zk_comm_init below).
 * b200zk_init_multi is the single-process form SURVEY.md section 8b sketched as `b200zk_init(n_gpus)`: it creates
 * the contexts of devices 0..n_gpus-1 and one NCCL communicator over them (ncclCommInitAll). */
int b200zk_init(int device, b200zk_ctx** out);
int b200zk_init_multi(int n_gpus, b200zk_ctx** out /* n_gpus entries */);
void b200zk_destroy(b200zk_ctx* ctx);
const char* b200zk_last_error(b200zk_ctx* ctx);
int b200zk_sync(b200zk_ctx* ctx);
/* options: "concurrency" (default 1): run the independent MSMs of a proof batch on auxiliary
 * streams; 0 serialises everything on the ctx stream (used for per-kernel profiling).
 * "msm_parts" (default 0 = automatic: 4 from 2^23 points, else 1): number of window groups a single
 * MSM over plain (not precomputed) bases is cut into, each a pass of the pipeline on its own stream.
 * "table_c_g1" (default 13) / "table_c_g2" (default 13): window of full digit tables (precompute level 2) built after
 * the call.
 * "msm_affine_levels" (default 4; 0 = off): rounds of pairwise batched-affine additions (5M + 1S each, one shared
 * inversion per warp) in front of the running sums of an MSM -- over digit tables for batches with at least
 * "msm_affine_min_entries" table entries (default 2^22; 0 = always), over plain bases (bucket method) from 2^26 entries;
 * "msm_affine_b" (default 96): target additions per lane that share one inversion.  Same result bytes with any setting.
 * "msm_glv" (default 1): MSMs (G1 and G2) over plain bases split every scalar as k1 + k2*lambda (two non-negative
 * 128/129-bit halves, phi(x, y) = (beta x, y)), and full digit tables built while it is on cover the half scalars
 * (130 instead of 256 bits of windows; the k2 rows are summed as they are and phi is applied once to their sum);
 * 0 keeps full-length scalars.  Same result bytes either way FOR BASES
 * IN THE ORDER-r SUBGROUP (k*P = k1*P + k2*phi(P) needs phi(P) = lambda*P, which holds on G1/G2 only: for the
 * order-3 point (0, 2) the two paths differ).  Every proving-key query and every ark `G1Affine`/`G2Affine` that
 * went through checked deserialisation is in the subgroup; for bases of unknown origin call
 * b200zk_points_validate(check_subgroup = 1) first or set msm_glv = 0. */
int b200zk_set_option(b200zk_ctx* ctx, const char* name, int value);
/* raw device memory for callers that keep operands resident in HBM */
int b200zk_dev_alloc(b200zk_ctx* ctx, size_t bytes, void** dptr);
int b200zk_dev_free(b200zk_ctx* ctx, void* dptr);
int b200zk_dev_upload(b200zk_ctx* ctx, void* dptr, const void* host, size_t bytes);
int b200zk_dev_download(b200zk_ctx* ctx, void* host, const void* dptr, size_t bytes);
/* the CUDA stream of the ctx (cudaStream_t), so callers can record their own events on it */
void* b200zk_stream(b200zk_ctx* ctx);

/* ---- profiling hooks (bench.py roofline leg) ----------------------------------------------
 * With profiling enabled every launch of a tracked kernel family is bracketed by CUDA events on
 * the ctx stream.  b200zk_prof_get syncs and returns the summed device milliseconds and launch
 * count of `name` since the last reset.  b200zk_launch_count = all kernel launches of the ctx. */
int b200zk_prof_enable(b200zk_ctx* ctx, int on);
int b200zk_prof_reset(b200zk_ctx* ctx);
int b200zk_prof_get(b200zk_ctx* ctx, const char* name, double* ms, long* launches);
int b200zk_prof_names(b200zk_ctx* ctx, char* buf, size_t buflen);   /* comma-separated */
/* one line "name start_ms end_ms" per bracket since the last reset, offsets from the first bracket: shows
 * how the streams of a proof batch overlap (valid when all brackets were resolved by one call). */
int b200zk_prof_timeline(b200zk_ctx* ctx, char* buf, size_t buflen);
long b200zk_launch_count(b200zk_ctx* ctx);
/* work counters accumulated by the kernels' host wrappers, for roofline arithmetic:
 * "msm_entries_g1"/"msm_entries_g2" = mixed additions issued by msm_accumulate, "msm_buckets_g1/2". */
int b200zk_stat_get(b200zk_ctx* ctx, const char* name, double* value);
int b200zk_stat_reset(b200zk_ctx* ctx);

/* ---- K1: field arithmetic test entry points ----------------------------------------------
 * Element-wise out[i] = a[i] op b[i] on the GPU.  Replaces nothing in the reference (row a10);
 * exists so the Montgomery kernels can be diffed against the oracle byte for byte. */
int b200zk_dbg_field_op(b200zk_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b,
                        uint8_t* out, size_t n);
/* Integer-pipe micro-benchmarks: kind 0 = 32-bit IMAD, 1 = IMAD.WIDE.U32 (mad.wide), 2 = Fr
 * Montgomery mul, 3 = Fq Montgomery mul, 4 = DFMA, 5 = Fq Montgomery mul on the FP64 pipe (experiment),
 * 6 = one integer and one FP64 product chain per thread, 7 = integer products in the even warps and FP64
 * products in the odd warps of every CTA, 8 = carry-chained wide multiply-adds (the rows of a Montgomery product,
 * four independent accumulators per thread), 9 = two independent Fq product chains per thread.  Returns operations (IMADs, or field muls) per
 * second sustained over all SMs -- the measured denominator of the integer roofline. */
int b200zk_dbg_int_peak(b200zk_ctx* ctx, int kind, double* ops_per_sec);
/* Groundwork (csrc/ec_batch_affine.cuh): out[i] = p[i] + q[i] for n pairs of affine points (host buffers, FFI
 * layout), one thread per `chunk` (1..64) consecutive pairs sharing one inversion.  ms_batch / ms_xyzz (may be
 * NULL): best-of-3 device time of that kernel and of the same additions done as XYZZ mixed additions. */
int b200zk_dbg_batch_add_affine(b200zk_ctx* ctx, int group, const uint8_t* p, const uint8_t* q, size_t n, int chunk,
                                uint8_t* out, double* ms_batch, double* ms_xyzz);
/* k_i * G for canonical scalars (fixed-base, used to make synthetic bases on the GPU):
 * group 1 -> 96 B G1 affine each, group 2 -> 192 B G2 affine each; host buffers. */
int b200zk_fixed_base_mul(b200zk_ctx* ctx, int group, const uint8_t* scalars, size_t n, uint8_t* out_points);
/* same, device buffers (d_scalars n*32 B canonical, d_out n*96/192 B) */
int b200zk_fixed_base_mul_device(b200zk_ctx* ctx, int group, const void* d_scalars, size_t n, void* d_out);

/* ---- K2/K3: EvaluationDomain ------------------------------------------------------------
 * Replaces ark_poly::Radix2EvaluationDomain::{fft_in_place, ifft_in_place} and the coset forms
 * reached through get_coset(offset) (ark-poly 0.4.2 [recall]; absent from the reference, row a9).
 * data: batch * 2^log_n Fr elements (Montgomery), natural order in and out, transformed in place.
 *   forward:  X[k] = sum_j x[j] * offset^j * w^(jk)
 *   inverse:  x[j] = offset^-j * n^-1 * sum_k X[k] * w^(-jk)
 * coset_offset: 32 B Montgomery Fr, or NULL for offset 1.  log_n > 32 -> DOMAIN_TOO_LARGE
 * (log_n above what fits device memory -> CUDA error).  The caller zero-pads to 2^log_n. */
int b200zk_ntt_fr(b200zk_ctx* ctx, uint8_t* data, uint32_t log_n, int inverse,
                  const uint8_t* coset_offset, size_t batch);
int b200zk_ntt_fr_device(b200zk_ctx* ctx, void* d_data, uint32_t log_n, int inverse,
                         const uint8_t* coset_offset, size_t batch);

/* Building blocks of the multi-GPU four-step transform (zk-apps_b200/sharded.py
 * ShardedNTT): n = n1 * n2 over G GPUs, rank g holds local[c][j1] = x[j1 * n2 + g * C + c], C = n2 / G.
 *   1. b200zk_ntt_fr_device(local, log n1, batch = C)           column transforms
 *   2. b200zk_ntt_twiddle_transpose_device                       out[k][c] = in[c][k] * w_n^(+-(row0 + c) * k),
 *      in = rows x cols Fr, out = cols x rows (chunk-major by destination rank), row0 = g * C
 *   3. all-to-all of (n1 / G) x C chunks (NCCL), b200zk_copy2d_device to interleave them into rows
 *   4. b200zk_ntt_fr_device(rows, log n2, batch = n1 / G)        row transforms
 * d_in and d_out must not alias. */
int b200zk_ntt_twiddle_transpose_device(b200zk_ctx* ctx, const void* d_in, void* d_out, uint32_t log_n,
                                        uint64_t rows, uint64_t cols, uint64_t row0, int inverse);
/* strided device-to-device copy on the ctx stream (cudaMemcpy2DAsync): `height` rows of `width` bytes */
int b200zk_copy2d_device(b200zk_ctx* ctx, void* d_dst, size_t dpitch, const void* d_src, size_t spitch,
                         size_t width, size_t height);

/* ---- K4/K5: VariableBaseMSM ------------------------------------------------------------
 * Replaces ark_ec::VariableBaseMSM::msm_bigint for G1Projective / G2Projective (ark-ec 0.4.2
 * [recall]; absent from the reference, rows a7/a8).  sum_i scalars[i] * bases[i], result affine.
 * inf_flags: n bytes (non-zero = base is the point at infinity) or NULL.
 * b200zk_msm_* take host buffers; *_resident take bases uploaded once (optionally with the
 * per-window multiples 2^(c*w) * P precomputed, which removes the window combine).
 * PRECONDITIONS (not checked on this path, as in ark's msm_bigint): scalars are canonical, i.e. < r -- what
 * `Fr::into_bigint()` yields; a 256-bit word >= 2^255 can lose the carry out of the top signed window -- and the
 * bases are reduced affine points of the order-r subgroup (see "msm_glv" above; b200zk_points_validate checks). */
int b200zk_msm_g1(b200zk_ctx* ctx, const uint8_t* bases, const uint8_t* inf_flags,
                  const uint8_t* scalars, size_t n, uint8_t out_affine[96], uint8_t* out_is_inf);
int b200zk_msm_g2(b200zk_ctx* ctx, const uint8_t* bases, const uint8_t* inf_flags,
                  const uint8_t* scalars, size_t n, uint8_t out_affine[192], uint8_t* out_is_inf);
/* group: 1 = G1, 2 = G2.  precompute: 0 = plain bases (GLV at MSM time); 1 = also the window multiples 2^(c w) P_i
 * (memory x windows; all windows then share one bucket set and nothing is left to combine); 2 = the FULL DIGIT TABLE
 * (m + 1) 2^(c w) P_i for every m < 2^(c-1), resident in HBM (n * ceil(130/c) * 2^(c-1) points over GLV half scalars,
 * ceil(256/c) windows with msm_glv = 0: 22 GB for a G1 query of 5,653 points at c = 13 -- sized for the 180 GB of a
 * B200; window from the options "table_c_g1" / "table_c_g2" (default 13) at the time of this call).  An MSM over a
 * full table is one addition per non-zero signed digit and nothing else: no buckets, no sort, no bucket reduction.
 * Same result bytes at every level. */
int b200zk_bases_upload(b200zk_ctx* ctx, int group, const uint8_t* bases, const uint8_t* inf_flags,
                        size_t n, int precompute, b200zk_bases** out);
/* wrap points already in device memory (affine, library layout); the library copies them */
int b200zk_bases_from_device(b200zk_ctx* ctx, int group, const void* d_points, size_t n,
                             int precompute, b200zk_bases** out);
void b200zk_bases_free(b200zk_ctx* ctx, b200zk_bases* h);
/* batch MSMs sharing the bases: scalars = batch * n * 32 B (row b starts at b*n*32), out = batch
 * affine points.  scalars_on_device != 0 means `scalars` is a device pointer. */
int b200zk_msm_resident(b200zk_ctx* ctx, const b200zk_bases* h, const void* scalars,
                        int scalars_on_device, size_t n, size_t batch, uint8_t* out_affine,
                        uint8_t* out_is_inf);

/* ---- multi-GPU (csrc/comm.cu) ---------------------------------------------------
 * NCCL is loaded at run time (dlopen "libnccl.so.2", or the path in $B200ZK_NCCL_LIB): the library has no link-time
 * dependency on it, and without it only the entry points of this section fail (B200ZK_ERR_NCCL).
 *   b200zk_comm_unique_id: rank 0 obtains the 128-byte ncclUniqueId; the host distributes it to the other ranks by
 *     its own means (a file, MPI, torch.distributed ...).
 *   b200zk_comm_init(ctx, id, rank, world): collective over all ranks (ncclCommInitRank); world == 1 needs no id
 *     and no NCCL.  b200zk_comm_destroy releases it (b200zk_destroy does so too).
 *   b200zk_msm_sharded[_device]: VariableBaseMSM over world GPUs, split by contiguous point range: h_local /
 *     scalars_local / n_local are THIS rank's slice (any sizes, 0 allowed).  Local bucket pipeline -> the affine
 *     partial is written by the last kernel straight into this rank's slot of the all_gather buffer -> in-place
 *     ncclAllGather of world * 96 B (G1) / 192 B (G2) on the ctx stream -> every rank adds the partials (curve
 *     addition is not an NCCL reduction op).  Every rank gets the same affine result.  The _device form leaves it
 *     in device memory and does not synchronise the host: kernels, collective and sum are one stream of work.
 *   b200zk_msm_sharded_multi: the same for the contexts of b200zk_init_multi, driven by ONE host thread (the
 *     collectives of all ranks are issued inside one NCCL group).
 *   b200zk_ntt_sharded_device: radix-2 Fr NTT of size 2^log_n = n1 * n2 (n1 = 2^log_n1) as a four-step transform
 *     with ONE exchange.  Layout L(a, b) of a length-a*b vector v over G ranks: rank g holds
 *     local[c][j] = v[j*b + g*(b/G) + c], c < b/G, j < a.  Input: this rank's part in L(n1, n2), n/G elements,
 *     transformed in place; output: the transform in L(n2, n1) -- so the opposite direction with log_n1' = log_n -
 *     log_n1 consumes it directly and returns to L(n1, n2).  G must be a power of two <= min(n1, n2).  Asynchronous
 *     on the ctx stream.  NTTs up to 2^22 are faster on one GPU: replicas only.
 * The building blocks stay public: b200zk_msm_resident_device (partial left in device memory, no host sync) and
 * b200zk_points_sum[_device]. */
#define B200ZK_COMM_ID_BYTES 128
int b200zk_comm_unique_id(uint8_t id_out[128]);
int b200zk_comm_init(b200zk_ctx* ctx, const uint8_t id[128], int rank, int world);
int b200zk_comm_destroy(b200zk_ctx* ctx);
int b200zk_comm_info(const b200zk_ctx* ctx, int* rank, int* world);
int b200zk_msm_sharded(b200zk_ctx* ctx, const b200zk_bases* h_local, const void* scalars_local, int scalars_on_device,
                       size_t n_local, uint8_t* out_affine, uint8_t* out_is_inf);
int b200zk_msm_sharded_device(b200zk_ctx* ctx, const b200zk_bases* h_local, const void* scalars_local,
                              int scalars_on_device, size_t n_local, void* d_out_affine);
int b200zk_msm_sharded_multi(b200zk_ctx** ctxs, int n_gpus, const b200zk_bases** h_local,
                             const void** scalars_local, int scalars_on_device, const size_t* n_local,
                             uint8_t* out_affine, uint8_t* out_is_inf);
int b200zk_ntt_sharded_device(b200zk_ctx* ctx, void* d_local, uint32_t log_n, uint32_t log_n1, int inverse);
int b200zk_msm_resident_device(b200zk_ctx* ctx, const b200zk_bases* h, const void* scalars, int scalars_on_device,
                               size_t n, size_t batch, void* d_out_affine);
int b200zk_points_sum(b200zk_ctx* ctx, int group, const uint8_t* points, size_t n, uint8_t* out_affine,
                      uint8_t* out_is_inf);
int b200zk_points_sum_device(b200zk_ctx* ctx, int group, const void* d_points, size_t n, void* d_out_affine);

/* ---- the shielder relation: ConstraintSynthesizer side (host only, no GPU needed) ---------------
 * R1CS of update_note_circuit (shielder/relations/src/relations/update_note.rs:106-149) with the
 * concrete Account/Operation of the mock (shielder/mocked_zk/src/account.rs, ops.rs); see
 * zk-apps_b200/csrc/host_r1cs.hpp.  kind: 0 = deposit, 1 = withdraw.  Variable numbering is
 * arkworks': z = [1, instance (amount, token, user, new_note_hash, merkle_root, old_nullifier),
 * witness...].  Matrices come out as CSR, coefficients 32 B Montgomery Fr. */
typedef struct b200zk_r1cs b200zk_r1cs;
int b200zk_update_note_r1cs(int kind, uint32_t tree_height, b200zk_r1cs** out);
/* R1CS of update_account_circuit as a relation of its own (shielder/relations/src/relations/update_account.rs:68-95;
 * UpdateAccountInput :18-30), same concrete Account/Operation.  z = [1, instance (old_account_hash,
 * new_account_hash, amount, token, user -- the struct's "public inputs" in field order), witness (old_account:
 * token0, balance0, token1, balance1, then the gadget witnesses)].  It is the sub-circuit update_note_circuit calls
 * last (update_note.rs:141-148); both are built from one gadget so they cannot drift. */
int b200zk_update_account_r1cs(int kind, b200zk_r1cs** out);
void b200zk_r1cs_free(b200zk_r1cs* r);
int b200zk_r1cs_shape(const b200zk_r1cs* r, uint64_t* num_constraints, uint64_t* num_inputs /* incl. ONE */,
                      uint64_t* num_aux, uint64_t nnz[3]);
int b200zk_r1cs_matrix(const b200zk_r1cs* r, int which /*0=A,1=B,2=C*/, uint64_t* row_ptr /* nc+1 */,
                       uint32_t* cols, uint8_t* vals);
/* Poseidon parameters of relations/src/lib.rs:17-26 (T=5, R_F=8, R_P=56) generated on the host by
 * the Grain LFSR: round constants 64*5*32 B, MDS 25*32 B, Montgomery Fr. */
int b200zk_poseidon_constants(uint8_t* round_constants, uint8_t* mds);

/* ---- K6: batched Poseidon and witness generation ------------------------------------------
 * b200zk_poseidon_hash_batch: n_hashes independent PoseidonHasher::hash_fix_len_array calls of
 * `arity` inputs each (halo2-base 0.4.1 semantics [recall]); inputs n_hashes*arity*32 B, out n*32 B.
 * b200zk_update_note_witness_batch: full R1CS assignment z of each instance.  inputs = batch rows of
 * (18 + 2*H) Fr in UpdateNoteInput::new argument order (update_note.rs:47-57):
 *   amount, token, user | new_note_hash | merkle_root | new_note (zk_id, trapdoor, nullifier,
 *   account_hash) | old_note (same 4) | path_shape[H] (0/1) | path[H] | op_priv.user |
 *   old_account (token0, balance0, token1, balance1).
 * out_assignments (host, may be NULL) / d_out_assignments (device, may be NULL): batch*num_vars*32 B.
 * out_status[b] (may be NULL): 0 = satisfied, 1 = unsatisfied.  Returns B200ZK_ERR_UNSATISFIED if
 * any instance is unsatisfied (assignments are still written). */
int b200zk_poseidon_hash_batch(b200zk_ctx* ctx, const uint8_t* inputs, size_t n_hashes, uint32_t arity, uint8_t* out);
int b200zk_update_note_witness_batch(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t* inputs, size_t batch,
                                     uint8_t* out_assignments, void* d_out_assignments, uint8_t* out_status);
/* same for the update-account relation: inputs = batch rows of 9 Fr in UpdateAccountInput::new argument order
 * (update_account.rs:37-42): old_account_hash | new_account_hash | operation (amount, token, user) |
 * old_account (token0, balance0, token1, balance1).  Any input word >= r marks the instance unsatisfied. */
int b200zk_update_account_witness_batch(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t* inputs, size_t batch,
                                        uint8_t* out_assignments, void* d_out_assignments, uint8_t* out_status);

/* ---- the note tree (csrc/merkle.cu) -----------------------------------
 * Replaces the reference's MerkleTree<DEPTH> (shielder/contract/merkle.rs:11-106) with the circuit's hash:
 * node = Poseidon hash_fix_len_array(&[left, right]) (relations/src/merkle_proof.rs:49-57) instead of SHA-256
 * (merkle.rs:24-28).  Same indexing and the same observable behaviour: heap layout, root = node 1, leaves at
 * size + idx, a node never written reads as 0 (an empty subtree is 0, not H(0,0)); add_leaf on a full tree ->
 * MERKLE_LIMIT_EXCEEDED; root of an empty tree -> MERKLE_NON_EXISTING_NODE; gen_proof on a FULL tree ->
 * MERKLE_PROOF_GEN_FAIL (merkle.rs:91-93, kept as is).  Leaves, nodes and roots are 32 B Montgomery Fr.
 *   b200zk_merkle_add_leaves = n x add_leaf (merkle.rs:48-80) in one call: n + n/2 + ... hashes.  All or nothing:
 *     if the n leaves do not fit none is added.  first_leaf_id (may be NULL) = id of leaves[0].  roots_out (may be
 *     NULL): n x 32 B, the root after each single insertion, i.e. what the contract pushes into roots_log.
 *     With log_roots != 0 at creation those n roots are also kept (host side) for is_historical_root.
 *   b200zk_merkle_gen_proofs = gen_proof (merkle.rs:89-102) for n leaves: path_out n x depth x 32 B (sibling per
 *     level, leaf level first), shape_out (may be NULL) n x depth bytes = MerkleProof::path_shape
 *     (merkle_proof.rs:11-14): 1 when the node on the path is the left child at that level.
 *   b200zk_merkle_fill_update_note_inputs_device: writes path_shape[H] (as Fr 0/1), path[H] and merkle_root into
 *     n device-resident input rows of b200zk_update_note_prove_batch_device (row = 18 + 2H Fr, H = depth), so a
 *     batch goes tree -> witness -> proof without touching the host.  d_leaf_ids: n x u64, each < 2^depth.
 *     Asynchronous on the ctx stream. */
typedef struct b200zk_merkle b200zk_merkle;
int b200zk_merkle_new(b200zk_ctx* ctx, uint32_t depth /* 1..31 */, int log_roots, b200zk_merkle** out);
void b200zk_merkle_free(b200zk_ctx* ctx, b200zk_merkle* t);
int b200zk_merkle_info(const b200zk_merkle* t, uint32_t* depth, uint64_t* size, uint64_t* next_leaf_idx);
int b200zk_merkle_add_leaves(b200zk_ctx* ctx, b200zk_merkle* t, const void* leaves, int leaves_on_device, size_t n,
                             uint64_t* first_leaf_id, uint8_t* roots_out);
int b200zk_merkle_root(b200zk_ctx* ctx, const b200zk_merkle* t, uint8_t out[32]);
int b200zk_merkle_node(b200zk_ctx* ctx, const b200zk_merkle* t, uint64_t id, uint8_t out[32]);
int b200zk_merkle_is_historical_root(b200zk_ctx* ctx, const b200zk_merkle* t, const uint8_t root[32], int* out);
int b200zk_merkle_gen_proofs(b200zk_ctx* ctx, const b200zk_merkle* t, const uint64_t* leaf_ids, size_t n,
                             uint8_t* path_out, uint8_t* shape_out);
int b200zk_merkle_fill_update_note_inputs_device(b200zk_ctx* ctx, const b200zk_merkle* t, const void* d_leaf_ids,
                                                 size_t n, void* d_inputs);

/* ---- Groth16 ------------------------------------------------------------------------------
 * b200zk_pk_upload: a proving key produced elsewhere (ark_groth16::ProvingKey fields, affine FFI
 * layout; query lengths: a, b_g1, b_g2 = num_variables, l = num_aux, h = domain_size - 1).
 * b200zk_groth16_setup: key generation from explicit toxic waste (alpha, beta, gamma, delta, tau:
 * 5 x 32 B canonical LE) -- ark_groth16::generate_parameters_with_qap [recall]; fixed-base
 * multiplications run on the GPU.  vk_out (may be NULL): alpha_g1 (96) | beta_g2 (192) |
 * gamma_g2 (192) | delta_g2 (192) | gamma_abc_g1 (num_inputs * 96).
 * precompute: 0 / 1 / 2 as in b200zk_bases_upload, applied to every query (2 = full digit tables: ~143 GB (133 GiB) for the
 * withdraw key at the default windows, built once in a few seconds).
 * b200zk_groth16_prove_batch = create_proof_with_reduction(circuit, pk, r, s) for `batch` full
 * assignments (batch*num_vars*32 B Montgomery, host or device); r, s: batch*32 B canonical LE.
 * proofs_out: batch * 192 B = compressed A (48) | B (96) | C (48), ark-serialize / zcash format.
 * points_out (may be NULL): batch * 384 B = affine A (96) | B (192) | C (96), FFI layout.
 * b200zk_update_note_prove_batch: witness generation + proving in one call (the user-facing path). */
int b200zk_pk_upload(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t* alpha_g1, const uint8_t* beta_g1,
                     const uint8_t* beta_g2, const uint8_t* delta_g1, const uint8_t* delta_g2, const uint8_t* a_query,
                     const uint8_t* b_g1_query, const uint8_t* b_g2_query, const uint8_t* l_query,
                     const uint8_t* h_query, int precompute, b200zk_pk** out);
int b200zk_groth16_setup(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t toxic[160], int precompute,
                         b200zk_pk** out, uint8_t* vk_out);
void b200zk_pk_free(b200zk_ctx* ctx, b200zk_pk* pk);
/* which: 0 = a_query, 1 = b_g1_query, 2 = b_g2_query, 3 = l_query, 4 = h_query (first window only) */
int b200zk_pk_export_query(b200zk_ctx* ctx, const b200zk_pk* pk, int which, uint8_t* out, size_t* count);
int b200zk_groth16_prove_batch(b200zk_ctx* ctx, const b200zk_pk* pk, const void* assignments, int on_device,
                               size_t batch, const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                               uint8_t* points_out);
int b200zk_update_note_prove_batch(b200zk_ctx* ctx, const b200zk_pk* pk, const uint8_t* inputs, size_t batch,
                                   const uint8_t* r, const uint8_t* s, uint8_t* proofs_out, uint8_t* out_status);
/* witness generation + proving for a key made from b200zk_update_account_r1cs (rows of 9 Fr, see above) */
int b200zk_update_account_prove_batch(b200zk_ctx* ctx, const b200zk_pk* pk, const uint8_t* inputs, size_t batch,
                                      const uint8_t* r, const uint8_t* s, uint8_t* proofs_out, uint8_t* out_status);
/* Asynchronous form: up to TWO batches in flight per ctx.  _submit stages the host buffers (inputs unless
 * inputs_on_device, r, s) in pinned memory of the library, enqueues witness generation and proving, and returns a
 * ticket without waiting for the GPU: the caller's input buffers are free again on return.  b200zk_prove_wait(ticket)
 * blocks until that batch is done and THEN fills proofs_out (batch * 192 B) and out_status (may be NULL) given to
 * _submit -- those two must stay valid until the wait; this is the one place where the library keeps host pointers
 * across calls.  An unsatisfied witness is reported by the wait (B200ZK_ERR_UNSATISFIED, nothing written to
 * proofs_out).  Submitting batch k+1 before waiting for batch k lets the head of one batch (input copy, witness
 * generation, sorting) and the tail of the other (bucket reduction, assembly) run under bucket accumulation: each
 * batch has its own main stream and scratch buffers, the MSM streams are shared.  The synchronous calls above are
 * submit + wait.  Tickets must be awaited in the order they were issued before a third submit. */
int b200zk_update_note_prove_submit(b200zk_ctx* ctx, const b200zk_pk* pk, const void* inputs, int inputs_on_device,
                                    size_t batch, const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                                    uint8_t* out_status, uint64_t* ticket);
int b200zk_update_account_prove_submit(b200zk_ctx* ctx, const b200zk_pk* pk, const void* inputs, int inputs_on_device,
                                       size_t batch, const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                                       uint8_t* out_status, uint64_t* ticket);
int b200zk_prove_wait(b200zk_ctx* ctx, uint64_t ticket);
/* same with the instance inputs already resident in device memory (r, s, proofs stay host buffers) */
int b200zk_update_note_prove_batch_device(b200zk_ctx* ctx, const b200zk_pk* pk, const void* d_inputs, size_t batch,
                                          const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                                          uint8_t* out_status);

/* ---- Groth16 verifier and wire formats ------------------------
 * Replaces the mock the contract calls where a verifier would sit: ZkProof::verify_creation /
 * verify_update (shielder/mocked_zk/src/relations.rs:127-155; call sites shielder/contract/lib.rs:56,74)
 * with ark_groth16::Groth16::verify_proof over a PreparedVerifyingKey [recall]:
 *     e(A, B) * e(L, -gamma) * e(C, -delta) == e(alpha, beta),  L = gamma_abc[0] + sum x_i gamma_abc[i+1].
 * b200zk_vk_upload: vk in b200zk_groth16_setup's vk_out layout (alpha_g1 96 | beta_g2 192 | gamma_g2 192 |
 *   delta_g2 192 | gamma_abc_g1 num_inputs*96, num_inputs counts the constant ONE); e(alpha, beta) is
 *   evaluated once here (process_vk).
 * b200zk_groth16_verify_batch: proofs = batch * 192 B compressed A|B|C; public_inputs = batch *
 *   (num_inputs-1) * 32 B Montgomery Fr (ark `&[Fr]`, the ONE is implicit); both host or both device.
 *   status_out (host, one per proof) receives a b200zk_proof_status: never an aggregate verdict.
 *   check_subgroup != 0 also tests subgroup membership of A, B, C (ark Validate::Yes on deserialisation): the
 *   endomorphism tests phi(P) == [z^2 - 1]P (G1) and psi(P) == [z]P (G2), equivalent to [r]P = O.
 * b200zk_groth16_verify_aggregate: ONE verdict for the whole batch from a random linear combination
 *   (prod e(r_i A_i, B_i) = e(alpha,beta)^(sum r_i) e(sum r_i L_i, gamma) e(sum r_i C_i, delta)): batch + 3
 *   Miller loops and one final exponentiation instead of 3 * batch and batch.  coeffs = batch * 16 B
 *   little-endian 128-bit randomisers chosen by the CALLER (soundness error 2^-128 over their choice); always a
 *   host buffer; a zero coefficient would exclude its proof from the check and is rejected (BAD_ARG).
 *   *all_valid = 1 iff every proof decodes and the combined equation holds.
 * b200zk_points_compress / _decompress: the zcash / ark-serialize compressed encoding (48 B G1, 96 B G2:
 *   big-endian x, flag bits 0x80 compressed, 0x40 infinity, 0x20 y lexicographically largest; G2 = x.c1|x.c0),
 *   one thread per point; decompress status per point: 0 ok, 1 bad encoding, 2 not on curve, 3 not in subgroup.
 * b200zk_vk_serialize / _deserialize: ark CanonicalSerialize (compressed) of VerifyingKey [recall]:
 *   alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | u64 LE len | gamma_abc_g1.   *len = 336 + 8 + 48*num_inputs.
 * b200zk_pk_serialize / _deserialize: ProvingKey [recall]: vk | beta_g1 | delta_g1 | a_query | b_g1_query |
 *   b_g2_query | h_query | l_query, every Vec as u64 LE length + compressed points.  Call _serialize with
 *   out == NULL to get *len.  _deserialize validates encodings always and subgroup membership on request. */
typedef enum {
    B200ZK_PROOF_ACCEPTED = 0,
    B200ZK_PROOF_REJECTED = 1,
    B200ZK_PROOF_BAD_ENCODING = 2,
    B200ZK_PROOF_NOT_ON_CURVE = 3,
    B200ZK_PROOF_NOT_IN_SUBGROUP = 4,
    B200ZK_PROOF_BAD_INPUT = 5
} b200zk_proof_status;
int b200zk_vk_upload(b200zk_ctx* ctx, const uint8_t* vk, uint32_t num_inputs, b200zk_vk** out);
void b200zk_vk_free(b200zk_ctx* ctx, b200zk_vk* vk);
int b200zk_vk_num_inputs(const b200zk_vk* vk, uint32_t* num_inputs);
int b200zk_vk_export(const b200zk_vk* vk, uint8_t* out /* 672 + 96 * num_inputs */);
int b200zk_groth16_verify_batch(b200zk_ctx* ctx, const b200zk_vk* vk, const void* proofs, const void* public_inputs,
                                int on_device, size_t batch, int check_subgroup, int32_t* status_out);
int b200zk_groth16_verify_aggregate(b200zk_ctx* ctx, const b200zk_vk* vk, const void* proofs, const void* public_inputs,
                                    int on_device, size_t batch, const uint8_t* coeffs, int check_subgroup,
                                    int* all_valid);
int b200zk_points_compress(b200zk_ctx* ctx, int group, const uint8_t* affine, size_t n, uint8_t* out);
int b200zk_points_decompress(b200zk_ctx* ctx, int group, const uint8_t* in, size_t n, int check_subgroup,
                             uint8_t* out_affine, int32_t* status);
/* per-point status (same codes as _decompress) of n affine points in the FFI layout: every coordinate word < p,
 * on the curve, and -- with check_subgroup -- in the order-r subgroup (the precondition of the GLV MSM path) */
int b200zk_points_validate(b200zk_ctx* ctx, int group, const uint8_t* affine, size_t n, int check_subgroup,
                           int32_t* status);
int b200zk_vk_serialize(b200zk_ctx* ctx, const b200zk_vk* vk, uint8_t* out, size_t* len);
int b200zk_vk_deserialize(b200zk_ctx* ctx, const uint8_t* in, size_t len, int check_subgroup, b200zk_vk** out);
int b200zk_pk_serialize(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_vk* vk, uint8_t* out, size_t* len);
int b200zk_pk_deserialize(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t* in, size_t len, int check_subgroup,
                          int precompute, b200zk_pk** pk_out, b200zk_vk** vk_out);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_H */
