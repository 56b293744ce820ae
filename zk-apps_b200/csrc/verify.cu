// Batched Groth16 verifier and the compressed point wire format on the GPU.
//
// What it replaces: the mock the contract calls where a verifier would sit -- ZkProof::verify_creation /
// verify_update (shielder/mocked_zk/src/relations.rs:127-155; call sites shielder/contract/lib.rs:56,74) -- with the
// check a real deployment needs: arkworks' Groth16::verify_proof over (PreparedVerifyingKey, proof, public inputs)
// [recall; not in the tree].  b200zk_groth16_verify_batch gives per-proof verdicts (one malformed or forged proof
// must not hide among valid ones); b200zk_groth16_verify_aggregate gives one verdict for the batch from a random
// linear combination at a third of the pairing work.
//
// Work split (every piece is a long dependent chain of Fq products, so the unit of parallelism is the thread and
// the kernels are latency-bound until the batch reaches a few thousand proofs):
//   verify_decode   batch x 3 threads   A, B, C from the 192-byte zcash-style encoding (sqrt, sign, optional [r]P = O)
//   verify_inputs   batch x num_public  x_i * gamma_abc[i+1] (double-and-add over the 255-bit input)
//   verify_miller   batch x 3 threads   L = gamma_abc[0] + sum terms; the three Miller loops, one per thread
//   verify_final    batch threads       product, final exponentiation, comparison with the prepared e(alpha, beta)
// Threads are indexed element-major (t = element * batch + proof) so a warp runs ONE code path (G1 vs G2 decode,
// one Miller loop) without divergence.
#include <algorithm>
#include <cstring>
#include <memory>

#include "types.cuh"
#include "verify.cuh"

using namespace b200zk;

struct b200zk_vk {
    uint32_t num_inputs = 0;            // incl. the constant ONE: gamma_abc has num_inputs points
    PreparedVk* d_vk = nullptr;
    Affine<Fq>* d_gamma_abc = nullptr;
    std::vector<uint8_t> raw;           // the uncompressed key as uploaded (b200zk_groth16_setup's vk_out layout)
};

namespace {

constexpr int VT = 32;  // threads per block: one warp, so small batches still spread over many SMs

__global__ void __launch_bounds__(VT) verify_prepare_kernel(const Affine<Fq>* alpha, const Affine<Fq2>* g2s, PreparedVk* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) prepare_vk(*alpha, g2s[0], g2s[1], g2s[2], *out);
}

__global__ void __launch_bounds__(VT) verify_decode_kernel(const uint8_t* __restrict__ proofs, uint32_t batch, int check_subgroup,
                                                           Affine<Fq>* __restrict__ a, Affine<Fq2>* __restrict__ b,
                                                           Affine<Fq>* __restrict__ c, int32_t* __restrict__ st3) {
    const uint32_t t = blockIdx.x * VT + threadIdx.x;
    if (t >= 3 * batch) return;
    const uint32_t which = t / batch, i = t % batch;
    const uint8_t* proof = proofs + (size_t)i * 192;
    int st;
    if (which == 1) {
        Affine<Fq2> q = Affine<Fq2>::inf();
        st = decode_proof_g2(proof, check_subgroup != 0, q);
        b[i] = q;
    } else {
        Affine<Fq> p = Affine<Fq>::inf();
        st = decode_proof_g1(proof, which, check_subgroup != 0, p);
        (which == 0 ? a : c)[i] = p;
    }
    st3[(size_t)i * 3 + which] = st;
}

__global__ void __launch_bounds__(VT) verify_inputs_kernel(const Affine<Fq>* __restrict__ gamma_abc, const Fr* __restrict__ inputs,
                                                           uint32_t batch, uint32_t num_public, XYZZ<Fq>* __restrict__ terms,
                                                           int32_t* __restrict__ bad_input) {
    const uint32_t t = blockIdx.x * VT + threadIdx.x;
    if (t >= batch * num_public) return;
    const uint32_t j = t / batch, i = t % batch;  // input-major: neighbouring lanes share the base
    XYZZ<Fq> r = XYZZ<Fq>::inf();
    if (!input_term(gamma_abc[j + 1], inputs[(size_t)i * num_public + j], r)) atomicOr(&bad_input[i], 1);
    terms[(size_t)i * num_public + j] = r;
}

__device__ __forceinline__ int first_failure(const int32_t* st3, const int32_t* bad_input, uint32_t i) {
    for (int k = 0; k < 3; k++)
        if (st3[(size_t)i * 3 + k] != PROOF_ACCEPTED) return st3[(size_t)i * 3 + k];
    return bad_input[i] ? PROOF_BAD_INPUT : PROOF_ACCEPTED;
}

__global__ void __launch_bounds__(VT) verify_miller_kernel(const PreparedVk* __restrict__ vk, const Affine<Fq>* __restrict__ gamma_abc,
                                                           const Affine<Fq>* __restrict__ a, const Affine<Fq2>* __restrict__ b,
                                                           const Affine<Fq>* __restrict__ c, const XYZZ<Fq>* __restrict__ terms,
                                                           const int32_t* __restrict__ st3, const int32_t* __restrict__ bad_input,
                                                           uint32_t batch, uint32_t num_public, Fq12* __restrict__ f) {
    const uint32_t t = blockIdx.x * VT + threadIdx.x;
    if (t >= 3 * batch) return;
    const uint32_t pair = t / batch, i = t % batch;
    Affine<Fq> p = Affine<Fq>::inf();
    Affine<Fq2> q = Affine<Fq2>::inf();
    if (first_failure(st3, bad_input, i) == PROOF_ACCEPTED) {
        if (pair == 0) {
            p = a[i];
            q = b[i];
        } else if (pair == 1) {
            XYZZ<Fq> acc = XYZZ<Fq>::from_affine(gamma_abc[0]);
            for (uint32_t j = 0; j < num_public; j++) ec_add(acc, terms[(size_t)i * num_public + j]);
            p = ec_to_affine(acc);
            q = vk->gamma_g2_neg;
        } else {
            p = c[i];
            q = vk->delta_g2_neg;
        }
    }
    f[(size_t)pair * batch + i] = miller_loop(p, q);  // one call site: the warp stays converged
}

__global__ void __launch_bounds__(VT) verify_final_kernel(const PreparedVk* __restrict__ vk, const Fq12* __restrict__ f,
                                                          const int32_t* __restrict__ st3, const int32_t* __restrict__ bad_input,
                                                          uint32_t batch, int32_t* __restrict__ status) {
    const uint32_t i = blockIdx.x * VT + threadIdx.x;
    if (i >= batch) return;
    int st = first_failure(st3, bad_input, i);
    const bool ok = verify_final(*vk, f[i], f[(size_t)batch + i], f[(size_t)2 * batch + i]);
    if (st == PROOF_ACCEPTED) st = ok ? PROOF_ACCEPTED : PROOF_REJECTED;
    status[i] = st;
}

// ---- aggregate verification (random linear combination)
// flags[0] = some proof failed to decode / has a bad input
__global__ void __launch_bounds__(VT) agg_scale_kernel(const Affine<Fq>* __restrict__ a, const Affine<Fq>* __restrict__ c,
                                                       const uint8_t* __restrict__ coeffs, const int32_t* __restrict__ st3,
                                                       uint32_t batch, Affine<Fq>* __restrict__ ra, XYZZ<Fq>* __restrict__ rc,
                                                       int32_t* __restrict__ flags) {
    const uint32_t t = blockIdx.x * VT + threadIdx.x;
    if (t >= 2 * batch) return;
    const uint32_t which = t / batch, i = t % batch;
    bool bad = false;
    for (int k = 0; k < 3; k++) bad = bad || st3[(size_t)i * 3 + k] != PROOF_ACCEPTED;
    if (bad && which == 0) atomicOr(&flags[0], 1);
    uint32_t k[4];
    for (int w = 0; w < 4; w++) {
        const uint8_t* q = coeffs + (size_t)i * 16 + 4 * w;
        k[w] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    }
    const Affine<Fq> p = bad ? Affine<Fq>::inf() : (which == 0 ? a : c)[i];
    const XYZZ<Fq> r = ec_mul_scalar(XYZZ<Fq>::from_affine(p), k, 4);
    if (which == 0) ra[i] = ec_to_affine(r);
    else rc[i] = r;
}

// block j < num_public: s_j = sum_i r_i x_ij; block num_public: s = sum_i r_i.  Then term_j = s_j * gamma_abc[j+1]
// (resp. s * gamma_abc[0] and, from the same s, -s * alpha).  One block per scalar, tree reduction in shared memory.
__global__ void __launch_bounds__(128) agg_inputs_kernel(const PreparedVk* __restrict__ vk, const Affine<Fq>* __restrict__ gamma_abc,
                                                         const Fr* __restrict__ inputs, const uint8_t* __restrict__ coeffs,
                                                         uint32_t batch, uint32_t num_public, XYZZ<Fq>* __restrict__ terms,
                                                         Affine<Fq>* __restrict__ neg_s_alpha, int32_t* __restrict__ flags) {
    __shared__ Fr sh[128];
    const uint32_t j = blockIdx.x;
    uint32_t m[8];
    for (int w = 0; w < 8; w++) m[w] = FrCfg::mod(w);
    Fr acc = Fr::zero();
    bool bad = false;
    for (uint32_t i = threadIdx.x; i < batch; i += blockDim.x) {
        Fr r = Fr::zero();
        for (int w = 0; w < 4; w++) {
            const uint8_t* q = coeffs + (size_t)i * 16 + 4 * w;
            r.v[w] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
        }
        r = fp_to_mont(r);
        if (j < num_public) {
            const Fr x = inputs[(size_t)i * num_public + j];
            if (!limbs_gt<8>(m, x.v)) bad = true;
            else acc = fp_add(acc, fp_mul(r, x));
        } else {
            acc = fp_add(acc, r);
        }
    }
    if (bad) atomicOr(&flags[0], 1);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = fp_add(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x >= 2 || (threadIdx.x == 1 && j < num_public)) return;
    const Fr sc = fp_from_mont(sh[0]);
    if (threadIdx.x == 0) {
        const Affine<Fq> base = gamma_abc[j < num_public ? j + 1 : 0];
        terms[j] = ec_mul_scalar(XYZZ<Fq>::from_affine(base), sc.v);
    } else {
        *neg_s_alpha = affine_neg(ec_to_affine(ec_mul_scalar(XYZZ<Fq>::from_affine(vk->alpha_g1), sc.v)));
    }
}

// strided partial sums of XYZZ points: out[block] = sum of in[block*VT + lane + k*stride]
__global__ void __launch_bounds__(VT) xyzz_sum_kernel(const XYZZ<Fq>* __restrict__ in, uint32_t n, XYZZ<Fq>* __restrict__ out) {
    __shared__ XYZZ<Fq> sh[VT];
    const uint32_t stride = gridDim.x * VT;
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < n; i += stride) ec_add(acc, in[i]);
    sh[threadIdx.x] = acc;
    __syncwarp();
    for (uint32_t s = VT / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            XYZZ<Fq> x = sh[threadIdx.x];
            ec_add(x, sh[threadIdx.x + s]);
            sh[threadIdx.x] = x;
        }
        __syncwarp();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// Miller loops of the combined equation: t < batch: (r_i A_i, B_i); then (-s alpha, beta), (L, -gamma), (C, -delta)
__global__ void __launch_bounds__(VT) agg_miller_kernel(const PreparedVk* __restrict__ vk, const Affine<Fq>* __restrict__ ra,
                                                        const Affine<Fq2>* __restrict__ b, const Affine<Fq>* __restrict__ neg_s_alpha,
                                                        const XYZZ<Fq>* __restrict__ l_terms, uint32_t n_terms,
                                                        const XYZZ<Fq>* __restrict__ c_sum, uint32_t batch, Fq12* __restrict__ f) {
    const uint32_t t = blockIdx.x * VT + threadIdx.x;
    if (t >= batch + 3) return;
    Affine<Fq> p;
    Affine<Fq2> q;
    if (t < batch) {
        p = ra[t];
        q = b[t];
    } else if (t == batch) {
        p = *neg_s_alpha;
        q = vk->beta_g2;
    } else if (t == batch + 1) {
        XYZZ<Fq> acc = XYZZ<Fq>::inf();
        for (uint32_t j = 0; j < n_terms; j++) ec_add(acc, l_terms[j]);
        p = ec_to_affine(acc);
        q = vk->gamma_g2_neg;
    } else {
        p = ec_to_affine(*c_sum);
        q = vk->delta_g2_neg;
    }
    f[t] = miller_loop(p, q);
}

__global__ void __launch_bounds__(VT) fq12_product_kernel(const Fq12* __restrict__ in, uint32_t n, Fq12* __restrict__ out) {
    __shared__ Fq12 sh[VT];
    const uint32_t stride = gridDim.x * VT;
    Fq12 acc = Fq12::one();
    for (uint32_t i = blockIdx.x * VT + threadIdx.x; i < n; i += stride) acc = fq12_mul(acc, in[i]);
    sh[threadIdx.x] = acc;
    __syncwarp();
    for (uint32_t s = VT / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = fq12_mul(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncwarp();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(VT) agg_final_kernel(const Fq12* __restrict__ f, const int32_t* __restrict__ flags,
                                                       int32_t* __restrict__ verdict) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const bool ok = final_exponentiation(*f) == Fq12::one();
    *verdict = (ok && flags[0] == 0) ? 1 : 0;
}

// ---- wire format: one thread per point
template <class F>
__global__ void points_compress_kernel(const Affine<F>* __restrict__ in, size_t n, uint8_t* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int W = sizeof(F);  // 48 / 96
    uint8_t buf[W];
    const Affine<F> p = in[i];
    if constexpr (W == 48) g1_compress(p, buf); else g2_compress(p, buf);
    for (int k = 0; k < W; k++) out[i * W + k] = buf[k];
}

template <class F>
__global__ void __launch_bounds__(64) points_decompress_kernel(const uint8_t* __restrict__ in, size_t n, int check_subgroup,
                                                               Affine<F>* __restrict__ out, int32_t* __restrict__ status) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int W = sizeof(F);
    uint8_t buf[W];
    for (int k = 0; k < W; k++) buf[k] = in[i * W + k];
    Affine<F> p = Affine<F>::inf();
    int st;
    if constexpr (W == 48) st = g1_decompress(buf, p); else st = g2_decompress(buf, p);
    if (st == POINT_OK && check_subgroup && !ec_in_subgroup(p)) st = POINT_NOT_IN_SUBGROUP;
    out[i] = p;
    status[i] = st;
}

// affine points as an MSM would take them: coordinates reduced, on the curve, in the order-r subgroup
template <class F>
__global__ void __launch_bounds__(64) points_validate_kernel(const Affine<F>* __restrict__ in, size_t n, int check_subgroup,
                                                             int32_t* __restrict__ status) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine<F> p = in[i];
    int st = POINT_OK;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&p);
    for (int k = 0; k < (int)(sizeof(Affine<F>) / sizeof(Fq)); k++) {  // every Fq word < p
        bool lt = false;
        for (int j = 11; j >= 0; j--) {
            const uint32_t a = w[k * 12 + j], m = FqCfg::mod(j);
            if (a != m) {
                lt = a < m;
                break;
            }
        }
        if (!lt) st = POINT_BAD_ENCODING;
    }
    if (st == POINT_OK) {
        bool on;
        if constexpr (sizeof(F) == sizeof(Fq)) on = ec_on_curve(p, g1_b());
        else on = ec_on_curve(p, g2_b());
        if (!on) st = POINT_NOT_ON_CURVE;
        else if (check_subgroup && !ec_in_subgroup(p)) st = POINT_NOT_IN_SUBGROUP;
    }
    status[i] = st;
}

void free_vk(b200zk_vk* vk) {
    if (!vk) return;
    if (vk->d_vk) cudaFree(vk->d_vk);
    if (vk->d_gamma_abc) cudaFree(vk->d_gamma_abc);
    delete vk;
}

}  // namespace

namespace b200zk {

int points_compress_device(b200zk_ctx* ctx, int group, const void* d_affine, size_t n, uint8_t* d_out) {
    if (n == 0) return B200ZK_OK;
    ProfScope p(ctx, "wire_format");
    if (group == 1)
        points_compress_kernel<Fq><<<div_up(n, 128), 128, 0, ctx->stream>>>((const Affine<Fq>*)d_affine, n, d_out);
    else
        points_compress_kernel<Fq2><<<div_up(n, 128), 128, 0, ctx->stream>>>((const Affine<Fq2>*)d_affine, n, d_out);
    return check_launch(ctx, "points_compress");
}

int points_decompress_device(b200zk_ctx* ctx, int group, const uint8_t* d_in, size_t n, bool check_subgroup, void* d_affine,
                             int32_t* d_status) {
    if (n == 0) return B200ZK_OK;
    ProfScope p(ctx, "wire_format");
    if (group == 1)
        points_decompress_kernel<Fq><<<div_up(n, 64), 64, 0, ctx->stream>>>(d_in, n, check_subgroup, (Affine<Fq>*)d_affine, d_status);
    else
        points_decompress_kernel<Fq2><<<div_up(n, 64), 64, 0, ctx->stream>>>(d_in, n, check_subgroup, (Affine<Fq2>*)d_affine, d_status);
    return check_launch(ctx, "points_decompress");
}

}  // namespace b200zk

extern "C" {

int b200zk_vk_upload(b200zk_ctx* ctx, const uint8_t* vk_bytes, uint32_t num_inputs, b200zk_vk** out) {
    if (!ctx || !vk_bytes || !out || num_inputs == 0) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<b200zk_vk, void (*)(b200zk_vk*)> vk(new b200zk_vk(), free_vk);
    vk->num_inputs = num_inputs;
    const size_t len = 672 + (size_t)num_inputs * 96;
    vk->raw.assign(vk_bytes, vk_bytes + len);
    B200ZK_CUDA(ctx, cudaMalloc(&vk->d_vk, sizeof(PreparedVk)));
    B200ZK_CUDA(ctx, cudaMalloc(&vk->d_gamma_abc, (size_t)num_inputs * 96));
    void* d_raw;
    B200ZK_TRY(scratch(ctx, "vk_raw", 672, &d_raw));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(d_raw, vk_bytes, 672, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(vk->d_gamma_abc, vk_bytes + 672, (size_t)num_inputs * 96, cudaMemcpyHostToDevice, ctx->stream));
    verify_prepare_kernel<<<1, VT, 0, ctx->stream>>>((const Affine<Fq>*)d_raw, (const Affine<Fq2>*)((const uint8_t*)d_raw + 96),
                                                     vk->d_vk);
    B200ZK_TRY(check_launch(ctx, "verify_prepare"));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = vk.release();
    return B200ZK_OK;
}

void b200zk_vk_free(b200zk_ctx* ctx, b200zk_vk* vk) {
    if (!vk) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    free_vk(vk);
}

int b200zk_vk_num_inputs(const b200zk_vk* vk, uint32_t* num_inputs) {
    if (!vk || !num_inputs) return B200ZK_ERR_BAD_ARG;
    *num_inputs = vk->num_inputs;
    return B200ZK_OK;
}

int b200zk_vk_export(const b200zk_vk* vk, uint8_t* out) {
    if (!vk || !out) return B200ZK_ERR_BAD_ARG;
    memcpy(out, vk->raw.data(), vk->raw.size());
    return B200ZK_OK;
}

int b200zk_groth16_verify_batch(b200zk_ctx* ctx, const b200zk_vk* vk, const void* proofs, const void* public_inputs,
                                int on_device, size_t batch, int check_subgroup, int32_t* status_out) {
    if (!ctx || !vk || (batch && (!status_out || !proofs || (vk->num_inputs > 1 && !public_inputs)))) return B200ZK_ERR_BAD_ARG;
    if (batch == 0) return B200ZK_OK;
    if (batch > (1u << 24)) return fail(ctx, B200ZK_ERR_BAD_LEN, "verify batch too large");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t nb = (uint32_t)batch, np = vk->num_inputs - 1;
    cudaStream_t st = ctx->stream;
    const uint8_t* d_proofs = (const uint8_t*)proofs;
    const Fr* d_inputs = (const Fr*)public_inputs;
    if (!on_device) {
        void *dp, *di;
        B200ZK_TRY(scratch(ctx, "verify_proofs", batch * 192, &dp));
        B200ZK_TRY(scratch(ctx, "verify_inputs", batch * (size_t)(np ? np : 1) * 32, &di));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(dp, proofs, batch * 192, cudaMemcpyHostToDevice, st));
        if (np) B200ZK_CUDA(ctx, cudaMemcpyAsync(di, public_inputs, batch * (size_t)np * 32, cudaMemcpyHostToDevice, st));
        d_proofs = (const uint8_t*)dp;
        d_inputs = (const Fr*)di;
    }
    void *da, *db, *dc, *dterms, *dflags, *df, *dstatus;
    B200ZK_TRY(scratch(ctx, "verify_a", batch * sizeof(Affine<Fq>), &da));
    B200ZK_TRY(scratch(ctx, "verify_b", batch * sizeof(Affine<Fq2>), &db));
    B200ZK_TRY(scratch(ctx, "verify_c", batch * sizeof(Affine<Fq>), &dc));
    B200ZK_TRY(scratch(ctx, "verify_terms", batch * (size_t)(np ? np : 1) * sizeof(XYZZ<Fq>), &dterms));
    B200ZK_TRY(scratch(ctx, "verify_flags", batch * 4 * sizeof(int32_t), &dflags));  // st3[batch][3] then bad_input[batch]
    B200ZK_TRY(scratch(ctx, "verify_f", batch * 3 * sizeof(Fq12), &df));
    B200ZK_TRY(scratch(ctx, "verify_status", batch * sizeof(int32_t), &dstatus));
    int32_t* st3 = (int32_t*)dflags;
    int32_t* bad = st3 + batch * 3;
    B200ZK_CUDA(ctx, cudaMemsetAsync(bad, 0, batch * sizeof(int32_t), st));
    {
        ProfScope p(ctx, "verify_decode");
        verify_decode_kernel<<<div_up(3 * batch, VT), VT, 0, st>>>(d_proofs, nb, check_subgroup, (Affine<Fq>*)da, (Affine<Fq2>*)db,
                                                                   (Affine<Fq>*)dc, st3);
        B200ZK_TRY(check_launch(ctx, "verify_decode"));
    }
    if (np) {
        ProfScope p(ctx, "verify_inputs");
        verify_inputs_kernel<<<div_up(batch * np, VT), VT, 0, st>>>(vk->d_gamma_abc, d_inputs, nb, np, (XYZZ<Fq>*)dterms, bad);
        B200ZK_TRY(check_launch(ctx, "verify_inputs"));
    }
    {
        ProfScope p(ctx, "verify_miller");
        verify_miller_kernel<<<div_up(3 * batch, VT), VT, 0, st>>>(vk->d_vk, vk->d_gamma_abc, (const Affine<Fq>*)da,
                                                                   (const Affine<Fq2>*)db, (const Affine<Fq>*)dc,
                                                                   (const XYZZ<Fq>*)dterms, st3, bad, nb, np, (Fq12*)df);
        B200ZK_TRY(check_launch(ctx, "verify_miller"));
    }
    {
        ProfScope p(ctx, "verify_final");
        verify_final_kernel<<<div_up(batch, VT), VT, 0, st>>>(vk->d_vk, (const Fq12*)df, st3, bad, nb, (int32_t*)dstatus);
        B200ZK_TRY(check_launch(ctx, "verify_final"));
    }
    B200ZK_CUDA(ctx, cudaMemcpyAsync(status_out, dstatus, batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(st));
    return B200ZK_OK;
}

int b200zk_groth16_verify_aggregate(b200zk_ctx* ctx, const b200zk_vk* vk, const void* proofs, const void* public_inputs,
                                    int on_device, size_t batch, const uint8_t* coeffs, int check_subgroup, int* all_valid) {
    if (!ctx || !vk || !all_valid || (batch && (!proofs || !coeffs || (vk->num_inputs > 1 && !public_inputs))))
        return B200ZK_ERR_BAD_ARG;
    *all_valid = 0;
    if (batch == 0) {
        *all_valid = 1;
        return B200ZK_OK;
    }
    if (batch > (1u << 24)) return fail(ctx, B200ZK_ERR_BAD_LEN, "verify batch too large");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t nb = (uint32_t)batch, np = vk->num_inputs - 1;
    cudaStream_t st = ctx->stream;
    const uint8_t* d_proofs = (const uint8_t*)proofs;
    const Fr* d_inputs = (const Fr*)public_inputs;
    if (!on_device) {
        void *dp, *di;
        B200ZK_TRY(scratch(ctx, "verify_proofs", batch * 192, &dp));
        B200ZK_TRY(scratch(ctx, "verify_inputs", batch * (size_t)(np ? np : 1) * 32, &di));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(dp, proofs, batch * 192, cudaMemcpyHostToDevice, st));
        if (np) B200ZK_CUDA(ctx, cudaMemcpyAsync(di, public_inputs, batch * (size_t)np * 32, cudaMemcpyHostToDevice, st));
        d_proofs = (const uint8_t*)dp;
        d_inputs = (const Fr*)di;
    }
    const uint32_t sum_blocks = std::min<uint32_t>(div_up(batch, VT), 64);
    void *da, *db, *dc, *dra, *drc, *dterms, *dflags, *df, *dco, *dmisc;
    B200ZK_TRY(scratch(ctx, "verify_a", batch * sizeof(Affine<Fq>), &da));
    B200ZK_TRY(scratch(ctx, "verify_b", batch * sizeof(Affine<Fq2>), &db));
    B200ZK_TRY(scratch(ctx, "verify_c", batch * sizeof(Affine<Fq>), &dc));
    B200ZK_TRY(scratch(ctx, "verify_ra", batch * sizeof(Affine<Fq>), &dra));
    B200ZK_TRY(scratch(ctx, "verify_rc", (batch + sum_blocks + 1) * sizeof(XYZZ<Fq>), &drc));
    B200ZK_TRY(scratch(ctx, "verify_terms", std::max<size_t>(batch * (size_t)(np ? np : 1), np + 1) * sizeof(XYZZ<Fq>), &dterms));
    B200ZK_TRY(scratch(ctx, "verify_flags", (batch * 4 + 4) * sizeof(int32_t), &dflags));
    B200ZK_TRY(scratch(ctx, "verify_f", (batch * 3 + 64 + 4) * sizeof(Fq12), &df));
    B200ZK_TRY(scratch(ctx, "verify_coeffs", batch * 16, &dco));
    B200ZK_TRY(scratch(ctx, "verify_misc", sizeof(Affine<Fq>), &dmisc));
    int32_t* st3 = (int32_t*)dflags;
    int32_t* flags = st3 + batch * 4;  // [0] failure seen, [1] verdict
    B200ZK_CUDA(ctx, cudaMemsetAsync(flags, 0, 4 * sizeof(int32_t), st));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dco, coeffs, batch * 16, cudaMemcpyHostToDevice, st));
    XYZZ<Fq>* rc = (XYZZ<Fq>*)drc;
    XYZZ<Fq>* rc_part = rc + batch;
    XYZZ<Fq>* rc_sum = rc_part + sum_blocks;
    Fq12* f = (Fq12*)df;
    {
        ProfScope p(ctx, "verify_decode");
        verify_decode_kernel<<<div_up(3 * batch, VT), VT, 0, st>>>(d_proofs, nb, check_subgroup, (Affine<Fq>*)da, (Affine<Fq2>*)db,
                                                                   (Affine<Fq>*)dc, st3);
        B200ZK_TRY(check_launch(ctx, "verify_decode"));
    }
    {
        ProfScope p(ctx, "verify_inputs");
        agg_scale_kernel<<<div_up(2 * batch, VT), VT, 0, st>>>((const Affine<Fq>*)da, (const Affine<Fq>*)dc, (const uint8_t*)dco, st3,
                                                               nb, (Affine<Fq>*)dra, rc, flags);
        B200ZK_TRY(check_launch(ctx, "agg_scale"));
        agg_inputs_kernel<<<np + 1, 128, 0, st>>>(vk->d_vk, vk->d_gamma_abc, d_inputs, (const uint8_t*)dco, nb, np,
                                                  (XYZZ<Fq>*)dterms, (Affine<Fq>*)dmisc, flags);
        B200ZK_TRY(check_launch(ctx, "agg_inputs"));
        xyzz_sum_kernel<<<sum_blocks, VT, 0, st>>>(rc, nb, rc_part);
        B200ZK_TRY(check_launch(ctx, "xyzz_sum"));
        xyzz_sum_kernel<<<1, VT, 0, st>>>(rc_part, sum_blocks, rc_sum);
        B200ZK_TRY(check_launch(ctx, "xyzz_sum"));
    }
    {
        ProfScope p(ctx, "verify_miller");
        agg_miller_kernel<<<div_up(batch + 3, VT), VT, 0, st>>>(vk->d_vk, (const Affine<Fq>*)dra, (const Affine<Fq2>*)db,
                                                                (const Affine<Fq>*)dmisc, (const XYZZ<Fq>*)dterms, np + 1, rc_sum, nb, f);
        B200ZK_TRY(check_launch(ctx, "agg_miller"));
    }
    {
        ProfScope p(ctx, "verify_final");
        const uint32_t n = nb + 3, blocks = std::min<uint32_t>(div_up(n, VT), 64);
        Fq12* part = f + n;
        fq12_product_kernel<<<blocks, VT, 0, st>>>(f, n, part);
        B200ZK_TRY(check_launch(ctx, "fq12_product"));
        fq12_product_kernel<<<1, VT, 0, st>>>(part, blocks, part + blocks);
        B200ZK_TRY(check_launch(ctx, "fq12_product"));
        agg_final_kernel<<<1, VT, 0, st>>>(part + blocks, flags, flags + 1);
        B200ZK_TRY(check_launch(ctx, "agg_final"));
    }
    int32_t verdict = 0;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(&verdict, flags + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(st));
    *all_valid = verdict;
    return B200ZK_OK;
}

int b200zk_points_compress(b200zk_ctx* ctx, int group, const uint8_t* affine, size_t n, uint8_t* out) {
    if (!ctx || (group != 1 && group != 2) || (n && (!affine || !out))) return B200ZK_ERR_BAD_ARG;
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t w = group == 1 ? 48 : 96;
    void *din, *dout;
    B200ZK_TRY(scratch(ctx, "wire_in", n * 2 * w, &din));
    B200ZK_TRY(scratch(ctx, "wire_out", n * 2 * w, &dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(din, affine, n * 2 * w, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(points_compress_device(ctx, group, din, n, (uint8_t*)dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, dout, n * w, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_points_decompress(b200zk_ctx* ctx, int group, const uint8_t* in, size_t n, int check_subgroup, uint8_t* out_affine,
                             int32_t* status) {
    if (!ctx || (group != 1 && group != 2) || (n && (!in || !out_affine || !status))) return B200ZK_ERR_BAD_ARG;
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t w = group == 1 ? 48 : 96;
    void *din, *dout, *dst;
    B200ZK_TRY(scratch(ctx, "wire_in", n * 2 * w, &din));
    B200ZK_TRY(scratch(ctx, "wire_out", n * 2 * w, &dout));
    B200ZK_TRY(scratch(ctx, "wire_status", n * sizeof(int32_t), &dst));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(din, in, n * w, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(points_decompress_device(ctx, group, (const uint8_t*)din, n, check_subgroup != 0, dout, (int32_t*)dst));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_affine, dout, n * 2 * w, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(status, dst, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_points_validate(b200zk_ctx* ctx, int group, const uint8_t* affine, size_t n, int check_subgroup,
                           int32_t* status) {
    if (!ctx || (group != 1 && group != 2) || (n && (!affine || !status))) return B200ZK_ERR_BAD_ARG;
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pt = group == 1 ? sizeof(Affine<Fq>) : sizeof(Affine<Fq2>);
    void *din, *dst;
    B200ZK_TRY(scratch(ctx, "wire_out", n * pt, &din));
    B200ZK_TRY(scratch(ctx, "wire_status", n * sizeof(int32_t), &dst));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(din, affine, n * pt, cudaMemcpyHostToDevice, ctx->stream));
    if (group == 1)
        points_validate_kernel<Fq><<<div_up(n, 64), 64, 0, ctx->stream>>>((const Affine<Fq>*)din, n, check_subgroup, (int32_t*)dst);
    else
        points_validate_kernel<Fq2><<<div_up(n, 64), 64, 0, ctx->stream>>>((const Affine<Fq2>*)din, n, check_subgroup, (int32_t*)dst);
    B200ZK_TRY(check_launch(ctx, "points_validate"));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(status, dst, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_vk_serialize(b200zk_ctx* ctx, const b200zk_vk* vk, uint8_t* out, size_t* len) {
    if (!ctx || !vk || !len) return B200ZK_ERR_BAD_ARG;
    const size_t need = 48 + 3 * 96 + 8 + (size_t)vk->num_inputs * 48;
    if (!out) {
        *len = need;
        return B200ZK_OK;
    }
    if (*len < need) return fail(ctx, B200ZK_ERR_BAD_LEN, "vk_serialize: buffer too small");
    *len = need;
    const uint8_t* raw = vk->raw.data();
    B200ZK_TRY(b200zk_points_compress(ctx, 1, raw, 1, out));
    B200ZK_TRY(b200zk_points_compress(ctx, 2, raw + 96, 3, out + 48));
    const uint64_t n = vk->num_inputs;
    for (int i = 0; i < 8; i++) out[336 + i] = (uint8_t)(n >> (8 * i));
    return b200zk_points_compress(ctx, 1, raw + 672, vk->num_inputs, out + 344);
}

int b200zk_vk_deserialize(b200zk_ctx* ctx, const uint8_t* in, size_t len, int check_subgroup, b200zk_vk** out) {
    if (!ctx || !in || !out) return B200ZK_ERR_BAD_ARG;
    if (len < 344) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "vk_deserialize: truncated");
    uint64_t n = 0;
    for (int i = 0; i < 8; i++) n |= (uint64_t)in[336 + i] << (8 * i);
    if (n == 0 || n > (1u << 24) || len != 344 + n * 48) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "vk_deserialize: bad length");
    std::vector<uint8_t> raw(672 + n * 96);
    std::vector<int32_t> st(n + 4);
    B200ZK_TRY(b200zk_points_decompress(ctx, 1, in, 1, check_subgroup, raw.data(), st.data()));
    B200ZK_TRY(b200zk_points_decompress(ctx, 2, in + 48, 3, check_subgroup, raw.data() + 96, st.data() + 1));
    B200ZK_TRY(b200zk_points_decompress(ctx, 1, in + 344, n, check_subgroup, raw.data() + 672, st.data() + 4));
    for (size_t i = 0; i < st.size(); i++)
        if (st[i] != POINT_OK)
            return fail(ctx, B200ZK_ERR_BAD_ENCODING, "vk_deserialize: invalid point " + std::to_string(i) + " (status " +
                                                          std::to_string(st[i]) + ")");
    return b200zk_vk_upload(ctx, raw.data(), (uint32_t)n, out);
}

}  // extern "C"
