// Groth16 prover pipeline on the GPU (BLS12-381), batched over proofs that share a proving key.
//
// Replaces ark_groth16::Groth16::<Bls12_381>::create_proof_with_reduction(circuit, &pk, r, s)
// with LibsnarkReduction ([recall]; no ark-groth16 pin and no
// prover of any kind exists in /root/reference -- section 0).  Per batch of B proofs:
//   1. r1cs_matvec:  a_i = <A_i,z>, b_i = <B_i,z>, c_i = <C_i,z>, a[nc + j] = z[j]  (CSR rows)
//   2. 3B iNTT -> 3B coset NTT (offset 7) -> h_pointwise (a*b - c) / Z(g)  -> B coset iNTT
//   3. five batched MSMs over the device-resident queries (a, b_g1, b_g2, l, h), scalars taken
//      straight from z / h in Montgomery form
//   4. assembly:  A = alpha + sum a_i z_i + r delta ;  B = beta + sum b_i z_i + s delta  (G1 and G2)
//                 C = sum l_i aux_i + sum h_i H_i + s A + r B1 - r s delta
//      and zcash-style compression to 48 + 96 + 48 bytes, all on the device.
#include <algorithm>
#include <cstring>
#include <memory>

#include "glv.cuh"
#include "types.cuh"

using namespace b200zk;

struct Csr {
    uint32_t* row_ptr = nullptr;
    uint32_t* cols = nullptr;
    Fr* vals = nullptr;
};

struct b200zk_pk {
    uint32_t num_constraints = 0, num_inputs = 0, num_aux = 0, log_n = 0;
    int kind = 0;
    int relation = 0;                  // host::RELATION_UPDATE_NOTE / _ACCOUNT: which witness generator feeds this key
    uint32_t tree_height = 0;
    uint32_t inputs_per_instance = 0;  // Fr elements per input row of that witness generator
    Csr m[3];
    b200zk_bases a_query, b_g1_query, b_g2_query, l_query, h_query;
    G1Affine alpha_g1, beta_g1, delta_g1;
    G2Affine beta_g2, delta_g2;
    void* d_singles = nullptr;  // alpha_g1, beta_g1, delta_g1 (G1Affine x3) then beta_g2, delta_g2 (G2Affine x2)
    void* d_delta_tab = nullptr;  // DeltaTable: 2^(8j) * delta for j < 32, G1 and G2 (fixed-base comb for r*delta, s*delta)
};

namespace {

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// abc layout: [3][batch][n]
__global__ void r1cs_matvec(Csr A, Csr B, Csr C, uint32_t nc, uint32_t num_inputs, uint32_t num_vars, uint32_t log_n,
                            const Fr* __restrict__ z_all, uint32_t batch, Fr* __restrict__ abc) {
    const uint32_t n = 1u << log_n;
    const uint32_t rows = nc + num_inputs;
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)rows * batch) return;
    const uint32_t b = (uint32_t)(t / rows), i = (uint32_t)(t % rows);
    const Fr* z = z_all + (size_t)b * num_vars;
    if (i >= nc) {  // input-consistency rows: a[nc + j] = z[j]
        st_fr(abc + ((size_t)0 * batch + b) * n + i, ld_fr(z + (i - nc)));
        return;
    }
    const Csr* M[3] = {&A, &B, &C};
#pragma unroll 1
    for (int m = 0; m < 3; m++) {
        Fr acc = Fr::zero();
        const uint32_t lo = M[m]->row_ptr[i], hi = M[m]->row_ptr[i + 1];
        for (uint32_t k = lo; k < hi; k++) acc = fp_add(acc, fp_mul(ld_fr(M[m]->vals + k), ld_fr(z + M[m]->cols[k])));
        st_fr(abc + ((size_t)m * batch + b) * n + i, acc);
    }
}

// a <- (a*b - c) * zinv over batch*n elements
__global__ void h_pointwise(Fr* __restrict__ abc, size_t count, Fr zinv) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    Fr a = ld_fr(abc + t), b = ld_fr(abc + count + t), c = ld_fr(abc + 2 * count + t);
    st_fr(abc + t, fp_mul(fp_sub(fp_mul(a, b), c), zinv));
}

struct Singles {
    G1Affine alpha_g1, beta_g1, delta_g1;
    G2Affine beta_g2, delta_g2;
};

struct DeltaTable {
    G1XYZZ g1[32];
    G2XYZZ g2[32];
};

// tab[j] = 2^(8j) * delta (one thread per group; runs once per key)
__global__ void delta_table_kernel(const Singles* __restrict__ sg, DeltaTable* __restrict__ tab) {
    if (threadIdx.x == 0) {
        G1XYZZ p = G1XYZZ::from_affine(sg->delta_g1);
        for (int j = 0; j < 32; j++) {
            tab->g1[j] = p;
            for (int d = 0; d < 8; d++) p = ec_dbl(p);
        }
    } else if (threadIdx.x == 32) {
        G2XYZZ p = G2XYZZ::from_affine(sg->delta_g2);
        for (int j = 0; j < 32; j++) {
            tab->g2[j] = p;
            for (int d = 0; d < 8; d++) p = ec_dbl(p);
        }
    }
}

// k * delta with the comb table: lane j multiplies byte j of k into tab[j] (8 double-and-adds), then the
// 32 partial sums are added in a shared-memory tree.  ~21 sequential group operations instead of ~380.
template <class F>
__device__ __forceinline__ XYZZ<F> comb_mul(const XYZZ<F>* __restrict__ tab, const uint32_t* k, XYZZ<F>* sh) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t digit = (k[lane >> 2] >> (8 * (lane & 3))) & 0xff;
    const XYZZ<F> base = tab[lane];
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int bit = 7; bit >= 0; bit--) {
        acc = ec_dbl(acc);
        if ((digit >> bit) & 1) ec_add(acc, base);
    }
    sh[lane] = acc;
    __syncwarp();
    for (uint32_t step = 16; step >= 1; step >>= 1) {
        if (lane < step) {
            XYZZ<F> b = sh[lane + step];
            ec_add(acc, b);
            sh[lane] = acc;
        }
        __syncwarp();
    }
    return acc;  // valid in lane 0
}

__device__ __forceinline__ void load_k(const uint8_t* p, uint32_t* k) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = q[i];
}

// The assembly is cut along its data dependencies so that every piece can run on the high-priority
// `fin` stream as soon as its inputs exist, hidden under the bucket accumulation of the slower MSMs
// (each piece is a handful of 255-bit scalar multiplications per proof: latency-bound, 4 warps).
//
// fin_scalars (needs r, s only): one warp per (j, proof); j = 0: r*delta1, 1: s*delta1, 2: s*delta2, 3: (r*s)*delta1
__global__ void __launch_bounds__(32) fin_scalars(const DeltaTable* __restrict__ tab, const uint8_t* __restrict__ r,
                                                  const uint8_t* __restrict__ s, uint32_t batch,
                                                  G1XYZZ* __restrict__ t_g1 /*[2][batch]*/, G2XYZZ* __restrict__ t_g2 /*[batch]*/,
                                                  G1XYZZ* __restrict__ u_g1 /*[3][batch], slot 2 written here*/) {
    __shared__ G2XYZZ sh2[32];
    G1XYZZ* sh1 = reinterpret_cast<G1XYZZ*>(sh2);
    const uint32_t j = blockIdx.x / batch, b = blockIdx.x % batch;
    uint32_t kr[8], ks[8];
    load_k(r + (size_t)b * 32, kr);
    load_k(s + (size_t)b * 32, ks);
    if (j < 2) {
        G1XYZZ v = comb_mul<Fq>(tab->g1, j == 0 ? kr : ks, sh1);
        if (threadIdx.x == 0) t_g1[(size_t)j * batch + b] = v;
    } else if (j == 2) {
        G2XYZZ v = comb_mul<Fq2>(tab->g2, ks, sh2);
        if (threadIdx.x == 0) t_g2[b] = v;
    } else {  // (r*s mod r) * delta1
        Fr fr_r, fr_s;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            fr_r.v[i] = kr[i];
            fr_s.v[i] = ks[i];
        }
        Fr rs = fp_from_mont(fp_mul(fp_to_mont(fr_r), fp_to_mont(fr_s)));
        G1XYZZ v = comb_mul<Fq>(tab->g1, rs.v, sh1);
        if (threadIdx.x == 0) u_g1[(size_t)2 * batch + b] = v;
    }
}

// k * P for P in G1 with the GLV endomorphism and Shamir's trick: k = k1 + k2 * lambda (glv.cuh), phi(P) = (beta X, Y, ZZ,
// ZZZ), one shared chain of 129 doublings with an addition of P, phi(P) or P + phi(P) where either bit is set:
// ~129 D + ~97 A instead of ~255 D + ~128 A.  This scalar multiplication is the longest latency chain of a single
// proof (one thread; nothing to overlap with once the MSMs are done).
__device__ G1XYZZ g1_mul_glv(const G1XYZZ& P, const uint32_t* k) {
    uint32_t k1[GLV_LIMBS], k2[GLV_LIMBS];
    glv_split(k, k1, k2);
    G1XYZZ Q = P;
    Q.x = glv_phi_x(P.x);
    G1XYZZ R = P;
    ec_add<Fq, CallOps>(R, Q);
    G1XYZZ acc = G1XYZZ::inf();
    bool started = false;
    for (int bit = 32 * GLV_LIMBS - 1; bit >= 0; bit--) {
        if (started) acc = ec_dbl(acc);
        const uint32_t b = ((k1[bit >> 5] >> (bit & 31)) & 1u) | (((k2[bit >> 5] >> (bit & 31)) & 1u) << 1);
        if (b) {
            ec_add<Fq, CallOps>(acc, b == 1 ? P : b == 2 ? Q : R);
            started = true;
        }
    }
    return acc;
}

// msm_g1 layout: [4][batch] = a, b_g1, l, h.
// fin_g1_mul: which = 0 (needs msm a):   u_g1[0] = s * A,   A  = alpha + msm_a  + r*delta1
//             which = 1 (needs msm b_g1): u_g1[1] = r * B1,  B1 = beta1 + msm_b1 + s*delta1
__global__ void __launch_bounds__(32) fin_g1_mul(const Singles* __restrict__ sg, const uint8_t* __restrict__ r,
                                                 const uint8_t* __restrict__ s, uint32_t batch, uint32_t which,
                                                 const G1Affine* __restrict__ msm_g1, const G1XYZZ* __restrict__ t_g1,
                                                 G1XYZZ* __restrict__ u_g1) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    uint32_t k[8];
    load_k((which == 0 ? s : r) + (size_t)b * 32, k);
    G1XYZZ P = t_g1[(size_t)which * batch + b];
    ec_madd(P, which == 0 ? sg->alpha_g1 : sg->beta_g1);
    ec_madd(P, msm_g1[(size_t)which * batch + b]);
    u_g1[(size_t)which * batch + b] = g1_mul_glv(P, k);
}

// canonical big-endian bytes of an Fq (48 B)
__device__ __forceinline__ void fq_to_be(const Fq& mont, uint8_t* out) {
    Fq c = fp_from_mont(mont);
    for (int k = 0; k < 48; k++) {
        int byte = 47 - k;
        out[k] = (uint8_t)(c.v[byte >> 2] >> (8 * (byte & 3)));
    }
}
// y > (p-1)/2  <=>  2y >= p  (canonical y)
__device__ __forceinline__ bool fq_lex_largest(const Fq& mont) {
    Fq c = fp_from_mont(mont);
    uint32_t d[13];
    uint32_t carry = 0;
    for (int i = 0; i < 12; i++) {
        d[i] = (c.v[i] << 1) | carry;
        carry = c.v[i] >> 31;
    }
    d[12] = carry;
    if (d[12]) return true;
    for (int i = 11; i >= 0; i--) {
        if (d[i] > FqCfg::mod(i)) return true;
        if (d[i] < FqCfg::mod(i)) return false;
    }
    return true;  // 2y == p cannot happen (p odd)
}
__device__ void compress_g1(const G1Affine& p, uint8_t* out) {
    if (p.is_inf()) {
        for (int i = 0; i < 48; i++) out[i] = 0;
        out[0] = 0xC0;
        return;
    }
    fq_to_be(p.x, out);
    out[0] |= 0x80 | (fq_lex_largest(p.y) ? 0x20 : 0);
}
__device__ void compress_g2(const G2Affine& p, uint8_t* out) {
    if (p.is_inf()) {
        for (int i = 0; i < 96; i++) out[i] = 0;
        out[0] = 0xC0;
        return;
    }
    fq_to_be(p.x.c1, out);
    fq_to_be(p.x.c0, out + 48);
    const bool largest = p.y.c1.is_zero() ? fq_lex_largest(p.y.c0) : fq_lex_largest(p.y.c1);
    out[0] |= 0x80 | (largest ? 0x20 : 0);
}

// fin_output: part 0: A (needs msm a), 1: B in G2 (needs msm b_g2), 2: C (needs everything) -> affine + compressed bytes
__global__ void __launch_bounds__(32) fin_output(const Singles* __restrict__ sg, uint32_t batch, uint32_t part,
                                                 const G1Affine* __restrict__ msm_g1,
                                                 const G2Affine* __restrict__ msm_g2,
                                                 const G1XYZZ* __restrict__ t_g1, const G2XYZZ* __restrict__ t_g2,
                                                 const G1XYZZ* __restrict__ u_g1, uint8_t* __restrict__ proofs,
                                                 uint8_t* __restrict__ points /* may be null: 384 B per proof */) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    uint8_t* out = proofs + (size_t)b * 192;
    if (part == 0) {
        G1XYZZ A = t_g1[b];
        ec_madd(A, sg->alpha_g1);
        ec_madd(A, msm_g1[(size_t)0 * batch + b]);
        G1Affine a = ec_to_affine(A);
        compress_g1(a, out);
        if (points) memcpy(points + (size_t)b * 384, &a, 96);
    } else if (part == 1) {
        G2XYZZ B = t_g2[b];
        ec_madd(B, sg->beta_g2);
        ec_madd(B, msm_g2[b]);
        G2Affine a = ec_to_affine(B);
        compress_g2(a, out + 48);
        if (points) memcpy(points + (size_t)b * 384 + 96, &a, 192);
    } else {
        G1XYZZ Cc = u_g1[(size_t)0 * batch + b];  // s*A
        G1XYZZ x = u_g1[(size_t)1 * batch + b];   // r*B1
        ec_add(Cc, x);
        x = ec_neg(u_g1[(size_t)2 * batch + b]);  // - r*s*delta1
        ec_add(Cc, x);
        ec_madd(Cc, msm_g1[(size_t)2 * batch + b]);  // l
        ec_madd(Cc, msm_g1[(size_t)3 * batch + b]);  // h
        G1Affine a = ec_to_affine(Cc);
        compress_g1(a, out + 144);
        if (points) memcpy(points + (size_t)b * 384 + 288, &a, 96);
    }
}

int upload_csr(b200zk_ctx* ctx, const std::vector<host::LC>& M, Csr* out) {
    std::vector<uint32_t> rp(M.size() + 1), cols;
    std::vector<Fr> vals;
    for (size_t i = 0; i < M.size(); i++) {
        rp[i] = (uint32_t)cols.size();
        for (auto& e : M[i].t) {
            cols.push_back(e.first);
            vals.push_back(e.second);
        }
    }
    rp[M.size()] = (uint32_t)cols.size();
    B200ZK_CUDA(ctx, cudaMalloc(&out->row_ptr, rp.size() * 4));
    B200ZK_CUDA(ctx, cudaMalloc(&out->cols, std::max<size_t>(1, cols.size()) * 4));
    B200ZK_CUDA(ctx, cudaMalloc(&out->vals, std::max<size_t>(1, vals.size()) * sizeof(Fr)));
    B200ZK_CUDA(ctx, cudaMemcpy(out->row_ptr, rp.data(), rp.size() * 4, cudaMemcpyHostToDevice));
    B200ZK_CUDA(ctx, cudaMemcpy(out->cols, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice));
    B200ZK_CUDA(ctx, cudaMemcpy(out->vals, vals.data(), vals.size() * sizeof(Fr), cudaMemcpyHostToDevice));
    return B200ZK_OK;
}

void free_pk(b200zk_pk* pk) {
    for (auto& m : pk->m) {
        if (m.row_ptr) cudaFree(m.row_ptr);
        if (m.cols) cudaFree(m.cols);
        if (m.vals) cudaFree(m.vals);
    }
    for (b200zk_bases* h : {&pk->a_query, &pk->b_g1_query, &pk->b_g2_query, &pk->l_query, &pk->h_query})
    {
        if (h->d_points) cudaFree(h->d_points);
        if (h->d_skip) cudaFree(h->d_skip);
        if (h->d_table) cudaFree(h->d_table);
    }
    if (pk->d_singles) cudaFree(pk->d_singles);
    if (pk->d_delta_tab) cudaFree(pk->d_delta_tab);
    delete pk;
}

int pk_common(b200zk_ctx* ctx, const b200zk_r1cs* r, b200zk_pk* pk) {
    const host::R1CS& cs = r->cs;
    pk->num_constraints = (uint32_t)cs.A.size();
    pk->num_inputs = cs.num_inputs;
    pk->num_aux = cs.num_aux;
    pk->kind = cs.kind;
    pk->relation = cs.relation;
    pk->tree_height = cs.tree_height;
    pk->inputs_per_instance = cs.inputs_per_instance;
    uint32_t lg = 0;
    while ((1ull << lg) < (uint64_t)pk->num_constraints + pk->num_inputs) lg++;
    pk->log_n = lg;
    B200ZK_TRY(upload_csr(ctx, cs.A, &pk->m[0]));
    B200ZK_TRY(upload_csr(ctx, cs.B, &pk->m[1]));
    B200ZK_TRY(upload_csr(ctx, cs.C, &pk->m[2]));
    return B200ZK_OK;
}

int pk_singles(b200zk_ctx* ctx, b200zk_pk* pk) {
    Singles sg{pk->alpha_g1, pk->beta_g1, pk->delta_g1, pk->beta_g2, pk->delta_g2};
    B200ZK_CUDA(ctx, cudaMalloc(&pk->d_singles, sizeof(Singles)));
    B200ZK_CUDA(ctx, cudaMemcpy(pk->d_singles, &sg, sizeof(sg), cudaMemcpyHostToDevice));
    B200ZK_CUDA(ctx, cudaMalloc(&pk->d_delta_tab, sizeof(DeltaTable)));
    delta_table_kernel<<<1, 64, 0, ctx->stream>>>((const Singles*)pk->d_singles, (DeltaTable*)pk->d_delta_tab);
    B200ZK_TRY(check_launch(ctx, "delta_table_kernel"));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

Fr fr_from_le(const uint8_t* p) {  // canonical LE -> Montgomery (host)
    Fr c;
    memcpy(&c, p, 32);
    return fp_to_mont(c);
}

}  // namespace

namespace b200zk {

// Everything that needs only (r, s): upload them and start the delta multiples on the `fin` stream.  Called
// before witness generation by the user-facing path, so this latency-bound piece hides under it.
int groth16_prove_begin(b200zk_ctx* ctx, const b200zk_pk* pk, size_t batch, const uint8_t* r, const uint8_t* s) {
    const size_t B = batch;
    void *d_rs, *d_t1, *d_t2, *d_u1;
    B200ZK_TRY(scratch(ctx, "g16_rs", 2 * B * 32, &d_rs));
    B200ZK_TRY(scratch(ctx, "g16_t1", 2 * B * sizeof(G1XYZZ), &d_t1));
    B200ZK_TRY(scratch(ctx, "g16_t2", B * sizeof(G2XYZZ), &d_t2));
    B200ZK_TRY(scratch(ctx, "g16_u1", 3 * B * sizeof(G1XYZZ), &d_u1));
    uint8_t* d_r = (uint8_t*)d_rs;
    uint8_t* d_s = d_r + B * 32;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(d_r, r, B * 32, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(d_s, s, B * 32, cudaMemcpyHostToDevice, ctx->stream));
    const cudaStream_t fin = ctx->concurrency ? ctx->fin : ctx->stream;
    if (ctx->concurrency) {
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        B200ZK_CUDA(ctx, cudaStreamWaitEvent(fin, ctx->ev_fork, 0));
    }
    {
        ProfScope ps(ctx, "finalize", fin);
        fin_scalars<<<(unsigned)(4 * B), 32, 0, fin>>>((const DeltaTable*)pk->d_delta_tab, d_r, d_s, (uint32_t)B,
                                                       (G1XYZZ*)d_t1, (G2XYZZ*)d_t2, (G1XYZZ*)d_u1);
        B200ZK_TRY(check_launch(ctx, "fin_scalars"));
    }
    if (ctx->concurrency) B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork2, fin));
    return B200ZK_OK;
}

// d_z: batch * num_vars Fr (Montgomery).  groth16_prove_begin must have been called for this batch.
// Enqueues the whole batch on the active lane's main stream and the shared MSM / assembly streams; the proof bytes
// (and points) are copied to proofs_out / points_out by the last operations on the main stream.  Does NOT
// synchronise the host: with pinned destinations the call returns while the GPU is still at the witness.
int groth16_prove_enqueue(b200zk_ctx* ctx, const b200zk_pk* pk, const Fr* d_z, size_t batch, uint8_t* proofs_out,
                          uint8_t* points_out) {
    const uint32_t n = 1u << pk->log_n, nv = pk->num_inputs + pk->num_aux;
    const size_t B = batch;
    void *d_abc, *d_rs, *d_msm1, *d_msm2, *d_t1, *d_t2, *d_u1, *d_proofs, *d_points = nullptr;
    B200ZK_TRY(scratch(ctx, "g16_abc", 3 * B * n * sizeof(Fr), &d_abc));
    B200ZK_TRY(scratch(ctx, "g16_rs", 2 * B * 32, &d_rs));
    B200ZK_TRY(scratch(ctx, "g16_msm1", 4 * B * sizeof(G1Affine), &d_msm1));
    B200ZK_TRY(scratch(ctx, "g16_msm2", B * sizeof(G2Affine), &d_msm2));
    B200ZK_TRY(scratch(ctx, "g16_t1", 2 * B * sizeof(G1XYZZ), &d_t1));
    B200ZK_TRY(scratch(ctx, "g16_t2", B * sizeof(G2XYZZ), &d_t2));
    B200ZK_TRY(scratch(ctx, "g16_u1", 3 * B * sizeof(G1XYZZ), &d_u1));
    B200ZK_TRY(scratch(ctx, "g16_proofs", B * 192, &d_proofs));
    if (points_out) B200ZK_TRY(scratch(ctx, "g16_points", B * 384, &d_points));
    uint8_t* d_r = (uint8_t*)d_rs;
    uint8_t* d_s = d_r + B * 32;
    const Singles* sg = (const Singles*)pk->d_singles;
    G1Affine* m1 = (G1Affine*)d_msm1;
    const uint32_t* zs = (const uint32_t*)d_z;
    // ---- the four MSMs that only need z run on auxiliary streams, concurrently with H(x) below:
    //      their latency-bound tails (bucket reduction) hide behind each other's bucket accumulation.
    //      The proof assembly runs piecewise on the high-priority `fin` stream as its inputs appear.
    const bool cc = ctx->concurrency;
    const cudaStream_t fin = cc ? ctx->fin : ctx->stream;
    const unsigned gb = div_up(B, 32);
    if (cc) {
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++) B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[i], ctx->ev_fork, 0));
    }
    const cudaStream_t fin2 = cc ? ctx->fin2 : ctx->stream;
    auto wait_msm = [&](cudaStream_t who, int slot, int ev) -> int {  // `who` waits for the MSM that just went to `slot`
        if (cc) {
            B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_msm[ev], slot_stream(ctx, slot)));
            B200ZK_CUDA(ctx, cudaStreamWaitEvent(who, ctx->ev_msm[ev], 0));
        }
        return B200ZK_OK;
    };
    B200ZK_TRY(msm_device<Fq>(ctx, &pk->a_query, zs, nv, nv, B, true, m1 + 0 * B, 1));
    B200ZK_TRY(wait_msm(fin, 1, 0));
    {   // fin: s*A, then the bytes of A
        ProfScope ps(ctx, "finalize", fin);
        fin_g1_mul<<<gb, 32, 0, fin>>>(sg, d_r, d_s, (uint32_t)B, 0, m1, (const G1XYZZ*)d_t1, (G1XYZZ*)d_u1);
        B200ZK_TRY(check_launch(ctx, "fin_g1_mul"));
        fin_output<<<gb, 32, 0, fin>>>(sg, (uint32_t)B, 0, m1, (const G2Affine*)d_msm2, (const G1XYZZ*)d_t1,
                                       (const G2XYZZ*)d_t2, (const G1XYZZ*)d_u1, (uint8_t*)d_proofs, (uint8_t*)d_points);
        B200ZK_TRY(check_launch(ctx, "fin_output"));
    }
    B200ZK_TRY(msm_device<Fq>(ctx, &pk->b_g1_query, zs, nv, nv, B, true, m1 + 1 * B, 2));
    if (cc) {  // fin2 starts after fin_scalars (first thing on fin) ...
        B200ZK_CUDA(ctx, cudaStreamWaitEvent(fin2, ctx->ev_fork2, 0));
    }
    B200ZK_TRY(wait_msm(fin2, 2, 1));  // ... and after the b_g1 MSM
    {   // fin2: r*B1, concurrently with s*A
        ProfScope ps(ctx, "finalize", fin2);
        fin_g1_mul<<<gb, 32, 0, fin2>>>(sg, d_r, d_s, (uint32_t)B, 1, m1, (const G1XYZZ*)d_t1, (G1XYZZ*)d_u1);
        B200ZK_TRY(check_launch(ctx, "fin_g1_mul"));
    }
    B200ZK_TRY(msm_device<Fq2>(ctx, &pk->b_g2_query, zs, nv, nv, B, true, (G2Affine*)d_msm2, 4));
    B200ZK_TRY(wait_msm(fin2, 4, 2));
    {   // fin2: the bytes of B
        ProfScope ps(ctx, "finalize", fin2);
        fin_output<<<gb, 32, 0, fin2>>>(sg, (uint32_t)B, 1, m1, (const G2Affine*)d_msm2, (const G1XYZZ*)d_t1,
                                        (const G2XYZZ*)d_t2, (const G1XYZZ*)d_u1, (uint8_t*)d_proofs, (uint8_t*)d_points);
        B200ZK_TRY(check_launch(ctx, "fin_output"));
    }
    B200ZK_TRY(msm_device<Fq>(ctx, &pk->l_query, zs + (size_t)pk->num_inputs * 8, pk->num_aux, nv, B, true, m1 + 2 * B, 3));

    // ---- H(x) = (A*B - C) / Z
    Fr* abc = (Fr*)d_abc;
    B200ZK_CUDA(ctx, cudaMemsetAsync(d_abc, 0, 3 * B * n * sizeof(Fr), ctx->stream));
    {
        ProfScope ps(ctx, "r1cs_matvec");
        const size_t threads = (size_t)(pk->num_constraints + pk->num_inputs) * B;
        r1cs_matvec<<<div_up(threads, 128), 128, 0, ctx->stream>>>(pk->m[0], pk->m[1], pk->m[2], pk->num_constraints,
                                                                   pk->num_inputs, nv, pk->log_n, d_z, (uint32_t)B, abc);
        B200ZK_TRY(check_launch(ctx, "r1cs_matvec"));
    }
    const Fr g = host::fr_from_u64(7);  // Fr::GENERATOR, the coset offset arkworks' Groth16 uses
    B200ZK_TRY(ntt_device(ctx, abc, pk->log_n, true, nullptr, 3 * B));
    B200ZK_TRY(ntt_device(ctx, abc, pk->log_n, false, &g, 3 * B));
    {
        // Z(g) = g^n - 1 on the coset
        Fr gn = g;
        for (uint32_t i = 0; i < pk->log_n; i++) gn = fp_sqr(gn);
        const Fr zinv = fp_inv(fp_sub(gn, Fr::one()));
        ProfScope ps(ctx, "h_pointwise");
        h_pointwise<<<div_up(B * n, 256), 256, 0, ctx->stream>>>(abc, B * n, zinv);
        B200ZK_TRY(check_launch(ctx, "h_pointwise"));
    }
    B200ZK_TRY(ntt_device(ctx, abc, pk->log_n, true, &g, B));
    // the h MSM has a stream of its own among the shared MSM streams (slot 5): with two batches in flight the
    // main stream of the other lane must not queue behind it, and its scratch is then shared through the stream
    if (cc) {
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_h, ctx->stream));
        B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[4], ctx->ev_h, 0));
    }
    B200ZK_TRY(msm_device<Fq>(ctx, &pk->h_query, (const uint32_t*)abc, n - 1, n, B, true, m1 + 3 * B, cc ? 5 : 0));
    if (cc) {
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_join[4], ctx->aux[4]));  // slot 5: h
        B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[4], 0));
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_join[2], ctx->aux[2]));  // slot 3: l
        B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[2], 0));
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_join[0], fin));          // fin has waited for a
        B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0));
        B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fin2, fin2));            // fin2 has waited for b_g1 and b_g2
        B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_fin2, 0));
    }

    // ---- C = s*A + r*B1 - rs*delta1 + l + h, and compression
    {
        ProfScope ps(ctx, "finalize");
        fin_output<<<gb, 32, 0, ctx->stream>>>(sg, (uint32_t)B, 2, m1, (const G2Affine*)d_msm2, (const G1XYZZ*)d_t1,
                                               (const G2XYZZ*)d_t2, (const G1XYZZ*)d_u1, (uint8_t*)d_proofs,
                                               (uint8_t*)d_points);
        B200ZK_TRY(check_launch(ctx, "fin_output"));
    }
    B200ZK_CUDA(ctx, cudaMemcpyAsync(proofs_out, d_proofs, B * 192, cudaMemcpyDeviceToHost, ctx->stream));
    if (points_out) B200ZK_CUDA(ctx, cudaMemcpyAsync(points_out, d_points, B * 384, cudaMemcpyDeviceToHost, ctx->stream));
    return B200ZK_OK;
}

int groth16_prove_device(b200zk_ctx* ctx, const b200zk_pk* pk, const Fr* d_z, size_t batch, uint8_t* proofs_out,
                         uint8_t* points_out) {
    B200ZK_TRY(groth16_prove_enqueue(ctx, pk, d_z, batch, proofs_out, points_out));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

}  // namespace b200zk

extern "C" {

int b200zk_groth16_setup(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t toxic[160], int precompute,
                         b200zk_pk** out, uint8_t* vk_out) {
    if (!ctx || !r || !toxic || !out) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const host::R1CS& cs = r->cs;
    std::unique_ptr<b200zk_pk, void (*)(b200zk_pk*)> pk(new b200zk_pk(), free_pk);
    B200ZK_TRY(pk_common(ctx, r, pk.get()));
    const uint32_t nc = pk->num_constraints, ni = pk->num_inputs, nv = ni + pk->num_aux, n = 1u << pk->log_n;
    const Fr alpha = fr_from_le(toxic), beta = fr_from_le(toxic + 32), gamma = fr_from_le(toxic + 64),
             delta = fr_from_le(toxic + 96), tau = fr_from_le(toxic + 128);
    if (gamma.is_zero() || delta.is_zero()) return fail(ctx, B200ZK_ERR_BAD_ARG, "gamma/delta must be non-zero");
    // ---- QAP at tau: Lagrange coefficients u_i = Z(tau) w^i / (n (tau - w^i))   (batch inversion)
    Fr w = Fr{{0x5f0e466au, 0xb9b58d8cu, 0x1819d7ecu, 0x5b1b4c80u, 0x52a31e64u, 0x0af53ae3u, 0x19e9b27bu, 0x5bf3addau}};
    for (uint32_t i = pk->log_n; i < 32; i++) w = fp_sqr(w);
    Fr tn = tau;
    for (uint32_t i = 0; i < pk->log_n; i++) tn = fp_sqr(tn);
    const Fr zt = fp_sub(tn, Fr::one());
    if (zt.is_zero()) return fail(ctx, B200ZK_ERR_BAD_ARG, "tau lies in the evaluation domain");
    std::vector<Fr> wi(n), den(n), pref(n), u(n);
    Fr cur = Fr::one();
    for (uint32_t i = 0; i < n; i++) {
        wi[i] = cur;
        den[i] = fp_sub(tau, cur);
        cur = fp_mul(cur, w);
    }
    Fr run = Fr::one();
    for (uint32_t i = 0; i < n; i++) {
        pref[i] = run;
        run = fp_mul(run, den[i]);
    }
    Fr inv_all = fp_inv(run);
    const Fr scale = fp_mul(zt, fp_inv(host::fr_from_u64(n)));
    for (int i = (int)n - 1; i >= 0; i--) {
        Fr inv_i = fp_mul(inv_all, pref[i]);
        inv_all = fp_mul(inv_all, den[i]);
        u[i] = fp_mul(fp_mul(scale, wi[i]), inv_i);
    }
    std::vector<Fr> a(nv, Fr::zero()), b(nv, Fr::zero()), c(nv, Fr::zero());
    for (uint32_t i = 0; i < ni; i++) a[i] = u[nc + i];
    for (uint32_t i = 0; i < nc; i++) {
        for (auto& e : cs.A[i].t) a[e.first] = fp_add(a[e.first], fp_mul(u[i], e.second));
        for (auto& e : cs.B[i].t) b[e.first] = fp_add(b[e.first], fp_mul(u[i], e.second));
        for (auto& e : cs.C[i].t) c[e.first] = fp_add(c[e.first], fp_mul(u[i], e.second));
    }
    const Fr ginv = fp_inv(gamma), dinv = fp_inv(delta);
    // scalar list (canonical): [alpha, beta, delta | gamma_abc (ni) | a (nv) | b (nv) | l (num_aux) | h (n-1)]
    std::vector<Fr> sc;
    sc.reserve(3 + ni + 2 * nv + pk->num_aux + n);
    sc.push_back(alpha);
    sc.push_back(beta);
    sc.push_back(delta);
    for (uint32_t i = 0; i < nv; i++) {
        Fr comb = fp_add(fp_add(fp_mul(beta, a[i]), fp_mul(alpha, b[i])), c[i]);
        if (i < ni) sc.push_back(fp_mul(comb, ginv));
    }
    for (auto& x : a) sc.push_back(x);
    for (auto& x : b) sc.push_back(x);
    for (uint32_t i = ni; i < nv; i++) {
        Fr comb = fp_add(fp_add(fp_mul(beta, a[i]), fp_mul(alpha, b[i])), c[i]);
        sc.push_back(fp_mul(comb, dinv));
    }
    Fr ti = fp_mul(zt, dinv);
    for (uint32_t i = 0; i + 1 < n; i++) {
        sc.push_back(ti);
        ti = fp_mul(ti, tau);
    }
    for (auto& x : sc) x = fp_from_mont(x);
    std::vector<Fr> sc2 = {fp_from_mont(beta), fp_from_mont(gamma), fp_from_mont(delta)};
    sc2.insert(sc2.end(), sc.begin() + 3 + ni + nv, sc.begin() + 3 + ni + 2 * nv);  // b (already canonical)

    void *ds1, *ds2, *dp1, *dp2;
    B200ZK_TRY(scratch(ctx, "setup_s1", sc.size() * 32, &ds1));
    B200ZK_TRY(scratch(ctx, "setup_s2", sc2.size() * 32, &ds2));
    B200ZK_TRY(scratch(ctx, "setup_p1", sc.size() * sizeof(G1Affine), &dp1));
    B200ZK_TRY(scratch(ctx, "setup_p2", sc2.size() * sizeof(G2Affine), &dp2));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(ds1, sc.data(), sc.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(ds2, sc2.data(), sc2.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(b200zk_fixed_base_mul_device(ctx, 1, ds1, sc.size(), dp1));
    B200ZK_TRY(b200zk_fixed_base_mul_device(ctx, 2, ds2, sc2.size(), dp2));
    const G1Affine* p1 = (const G1Affine*)dp1;
    const G2Affine* p2 = (const G2Affine*)dp2;
    G1Affine s1[3];
    G2Affine s2[3];
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s1, p1, sizeof(s1), cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(s2, p2, sizeof(s2), cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pk->alpha_g1 = s1[0];
    pk->beta_g1 = s1[1];
    pk->delta_g1 = s1[2];
    pk->beta_g2 = s2[0];
    pk->delta_g2 = s2[2];
    if (vk_out) {  // alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | gamma_abc_g1[ni]
        memcpy(vk_out, &s1[0], 96);
        memcpy(vk_out + 96, &s2[0], 192);
        memcpy(vk_out + 288, &s2[1], 192);
        memcpy(vk_out + 480, &s2[2], 192);
        B200ZK_CUDA(ctx, cudaMemcpy(vk_out + 672, p1 + 3, (size_t)ni * 96, cudaMemcpyDeviceToHost));
    }
    pk->a_query.group = pk->b_g1_query.group = pk->l_query.group = pk->h_query.group = 1;
    pk->b_g2_query.group = 2;
    size_t off = 3 + ni;
    B200ZK_TRY(bases_build<Fq>(ctx, &pk->a_query, p1 + off, true, nullptr, nv, precompute));
    off += nv;
    B200ZK_TRY(bases_build<Fq>(ctx, &pk->b_g1_query, p1 + off, true, nullptr, nv, precompute));
    off += nv;
    B200ZK_TRY(bases_build<Fq>(ctx, &pk->l_query, p1 + off, true, nullptr, pk->num_aux, precompute));
    off += pk->num_aux;
    B200ZK_TRY(bases_build<Fq>(ctx, &pk->h_query, p1 + off, true, nullptr, n - 1, precompute));
    B200ZK_TRY(bases_build<Fq2>(ctx, &pk->b_g2_query, p2 + 3, true, nullptr, nv, precompute));
    B200ZK_TRY(pk_singles(ctx, pk.get()));
    *out = pk.release();
    return B200ZK_OK;
}

int b200zk_pk_upload(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t* alpha_g1, const uint8_t* beta_g1,
                     const uint8_t* beta_g2, const uint8_t* delta_g1, const uint8_t* delta_g2, const uint8_t* a_query,
                     const uint8_t* b_g1_query, const uint8_t* b_g2_query, const uint8_t* l_query,
                     const uint8_t* h_query, int precompute, b200zk_pk** out) {
    if (!ctx || !r || !out || !alpha_g1 || !beta_g1 || !beta_g2 || !delta_g1 || !delta_g2 || !a_query || !b_g1_query ||
        !b_g2_query || !l_query || !h_query)
        return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<b200zk_pk, void (*)(b200zk_pk*)> pk(new b200zk_pk(), free_pk);
    B200ZK_TRY(pk_common(ctx, r, pk.get()));
    const uint32_t nv = pk->num_inputs + pk->num_aux, n = 1u << pk->log_n;
    memcpy(&pk->alpha_g1, alpha_g1, 96);
    memcpy(&pk->beta_g1, beta_g1, 96);
    memcpy(&pk->delta_g1, delta_g1, 96);
    memcpy(&pk->beta_g2, beta_g2, 192);
    memcpy(&pk->delta_g2, delta_g2, 192);
    pk->a_query.group = pk->b_g1_query.group = pk->l_query.group = pk->h_query.group = 1;
    pk->b_g2_query.group = 2;
    B200ZK_TRY(bases_build<Fq>(ctx, &pk->a_query, (const G1Affine*)a_query, false, nullptr, nv, precompute));
    B200ZK_TRY(bases_build<Fq>(ctx, &pk->b_g1_query, (const G1Affine*)b_g1_query, false, nullptr, nv, precompute));
    B200ZK_TRY(bases_build<Fq>(ctx, &pk->l_query, (const G1Affine*)l_query, false, nullptr, pk->num_aux, precompute));
    B200ZK_TRY(bases_build<Fq>(ctx, &pk->h_query, (const G1Affine*)h_query, false, nullptr, n - 1, precompute));
    B200ZK_TRY(bases_build<Fq2>(ctx, &pk->b_g2_query, (const G2Affine*)b_g2_query, false, nullptr, nv, precompute));
    B200ZK_TRY(pk_singles(ctx, pk.get()));
    *out = pk.release();
    return B200ZK_OK;
}

void b200zk_pk_free(b200zk_ctx* ctx, b200zk_pk* pk) {
    if (!pk) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    free_pk(pk);
}

int b200zk_pk_export_query(b200zk_ctx* ctx, const b200zk_pk* pk, int which, uint8_t* out, size_t* count) {
    if (!ctx || !pk || which < 0 || which > 4) return B200ZK_ERR_BAD_ARG;
    const b200zk_bases* h = which == 0 ? &pk->a_query : which == 1 ? &pk->b_g1_query : which == 2 ? &pk->b_g2_query
                            : which == 3 ? &pk->l_query : &pk->h_query;
    if (count) *count = h->n;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (out) B200ZK_CUDA(ctx, cudaMemcpy(out, h->d_points, h->n * (h->group == 1 ? 96 : 192), cudaMemcpyDeviceToHost));
    return B200ZK_OK;
}

// ---- ark CanonicalSerialize (compressed) of ProvingKey [recall; SURVEY.md Appendix B]:
//   vk | beta_g1 | delta_g1 | a_query | b_g1_query | b_g2_query | h_query | l_query, Vec = u64 LE length + points.
int b200zk_pk_serialize(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_vk* vk, uint8_t* out, size_t* len) {
    if (!ctx || !pk || !vk || !len) return B200ZK_ERR_BAD_ARG;
    size_t vk_len = 0;
    B200ZK_TRY(b200zk_vk_serialize(ctx, vk, nullptr, &vk_len));
    const b200zk_bases* q[5] = {&pk->a_query, &pk->b_g1_query, &pk->b_g2_query, &pk->h_query, &pk->l_query};
    size_t need = vk_len + 96;
    for (auto* h : q) need += 8 + h->n * (h->group == 1 ? 48 : 96);
    if (!out) {
        *len = need;
        return B200ZK_OK;
    }
    if (*len < need) return fail(ctx, B200ZK_ERR_BAD_LEN, "pk_serialize: buffer too small");
    *len = need;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t off = vk_len;
    B200ZK_TRY(b200zk_vk_serialize(ctx, vk, out, &vk_len));
    uint8_t singles[192];
    memcpy(singles, &pk->beta_g1, 96);
    memcpy(singles + 96, &pk->delta_g1, 96);
    B200ZK_TRY(b200zk_points_compress(ctx, 1, singles, 2, out + off));
    off += 96;
    for (auto* h : q) {
        const uint64_t n = h->n;
        const size_t w = h->group == 1 ? 48 : 96;
        for (int i = 0; i < 8; i++) out[off + i] = (uint8_t)(n >> (8 * i));
        off += 8;
        if (n) {
            void* d;
            B200ZK_TRY(scratch(ctx, "wire_out", n * 2 * w, &d));
            B200ZK_TRY(points_compress_device(ctx, h->group, h->d_points, n, (uint8_t*)d));
            B200ZK_CUDA(ctx, cudaMemcpyAsync(out + off, d, n * w, cudaMemcpyDeviceToHost, ctx->stream));
            B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        off += n * w;
    }
    return B200ZK_OK;
}

int b200zk_pk_deserialize(b200zk_ctx* ctx, const b200zk_r1cs* r, const uint8_t* in, size_t len, int check_subgroup,
                          int precompute, b200zk_pk** pk_out, b200zk_vk** vk_out) {
    if (!ctx || !r || !in || !pk_out) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    auto rd_u64 = [&](size_t off) {
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v |= (uint64_t)in[off + i] << (8 * i);
        return v;
    };
    if (len < 344) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "pk_deserialize: truncated");
    const uint64_t ni = rd_u64(336);
    if (ni > (1u << 24)) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "pk_deserialize: bad gamma_abc length");
    const size_t vk_len = 344 + (size_t)ni * 48;
    if (len < vk_len + 96) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "pk_deserialize: truncated");
    std::unique_ptr<b200zk_pk, void (*)(b200zk_pk*)> pk(new b200zk_pk(), free_pk);
    B200ZK_TRY(pk_common(ctx, r, pk.get()));
    const uint32_t nv = pk->num_inputs + pk->num_aux, n = 1u << pk->log_n;
    if (ni != pk->num_inputs) return fail(ctx, B200ZK_ERR_BAD_LEN, "pk_deserialize: key does not match the relation (inputs)");
    b200zk_vk* vk = nullptr;
    B200ZK_TRY(b200zk_vk_deserialize(ctx, in, vk_len, check_subgroup, &vk));
    std::unique_ptr<b200zk_vk, void (*)(b200zk_vk*)> vk_guard(vk, [](b200zk_vk* v) { b200zk_vk_free(nullptr, v); });
    std::vector<uint8_t> vk_raw(672 + (size_t)ni * 96);
    B200ZK_TRY(b200zk_vk_export(vk, vk_raw.data()));
    memcpy(&pk->alpha_g1, vk_raw.data(), 96);
    memcpy(&pk->beta_g2, vk_raw.data() + 96, 192);
    memcpy(&pk->delta_g2, vk_raw.data() + 480, 192);
    size_t off = vk_len;
    {
        uint8_t singles[192];
        int32_t st[2];
        B200ZK_TRY(b200zk_points_decompress(ctx, 1, in + off, 2, check_subgroup, singles, st));
        if (st[0] || st[1]) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "pk_deserialize: invalid beta_g1 / delta_g1");
        memcpy(&pk->beta_g1, singles, 96);
        memcpy(&pk->delta_g1, singles + 96, 96);
        off += 96;
    }
    b200zk_bases* q[5] = {&pk->a_query, &pk->b_g1_query, &pk->b_g2_query, &pk->h_query, &pk->l_query};
    const size_t want[5] = {nv, nv, nv, (size_t)n - 1, pk->num_aux};
    const char* names[5] = {"a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"};
    for (int k = 0; k < 5; k++) {
        const int group = k == 2 ? 2 : 1;
        const size_t w = group == 1 ? 48 : 96;
        if (len < off + 8) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "pk_deserialize: truncated");
        const uint64_t cnt = rd_u64(off);
        off += 8;
        if (cnt != want[k])
            return fail(ctx, B200ZK_ERR_BAD_LEN, std::string("pk_deserialize: key does not match the relation (") + names[k] + ")");
        if (len < off + cnt * w) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "pk_deserialize: truncated");
        void *din, *dpts, *dst;
        B200ZK_TRY(scratch(ctx, "wire_in", std::max<size_t>(cnt, 1) * 2 * w, &din));
        B200ZK_TRY(scratch(ctx, "wire_out", std::max<size_t>(cnt, 1) * 2 * w, &dpts));
        B200ZK_TRY(scratch(ctx, "wire_status", std::max<size_t>(cnt, 1) * sizeof(int32_t), &dst));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(din, in + off, cnt * w, cudaMemcpyHostToDevice, ctx->stream));
        B200ZK_TRY(points_decompress_device(ctx, group, (const uint8_t*)din, cnt, check_subgroup != 0, dpts, (int32_t*)dst));
        std::vector<int32_t> st(cnt);
        B200ZK_CUDA(ctx, cudaMemcpyAsync(st.data(), dst, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < cnt; i++)
            if (st[i] != 0)
                return fail(ctx, B200ZK_ERR_BAD_ENCODING, std::string("pk_deserialize: invalid point in ") + names[k] + " at " +
                                                              std::to_string(i) + " (status " + std::to_string(st[i]) + ")");
        q[k]->group = group;
        if (group == 1) B200ZK_TRY(bases_build<Fq>(ctx, q[k], (const G1Affine*)dpts, true, nullptr, cnt, precompute));
        else B200ZK_TRY(bases_build<Fq2>(ctx, q[k], (const G2Affine*)dpts, true, nullptr, cnt, precompute));
        off += cnt * w;
    }
    if (off != len) return fail(ctx, B200ZK_ERR_BAD_ENCODING, "pk_deserialize: trailing bytes");
    B200ZK_TRY(pk_singles(ctx, pk.get()));
    *pk_out = pk.release();
    if (vk_out) *vk_out = vk_guard.release();
    return B200ZK_OK;
}

int b200zk_groth16_prove_batch(b200zk_ctx* ctx, const b200zk_pk* pk, const void* assignments, int on_device,
                               size_t batch, const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                               uint8_t* points_out) {
    if (!ctx || !pk || !assignments || !r || !s || !proofs_out) return B200ZK_ERR_BAD_ARG;
    if (batch == 0) return B200ZK_OK;
    for (auto& pb : ctx->pending)
        if (pb.active) return fail(ctx, B200ZK_ERR_BAD_ARG, "a submitted batch is still in flight: b200zk_prove_wait first");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nv = pk->num_inputs + pk->num_aux;
    const Fr* dz = (const Fr*)assignments;
    if (!on_device) {
        void* d;
        B200ZK_TRY(scratch(ctx, "g16_z", batch * nv * sizeof(Fr), &d));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(d, assignments, batch * nv * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
        dz = (const Fr*)d;
    }
    B200ZK_TRY(groth16_prove_begin(ctx, pk, batch, r, s));
    return groth16_prove_device(ctx, pk, dz, batch, proofs_out, points_out);
}

static int grow_pinned(b200zk_ctx* ctx, void** p, size_t* cap, size_t need) {
    if (*cap >= need) return B200ZK_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    B200ZK_CUDA(ctx, cudaMallocHost(p, need + need / 4));
    *cap = need + need / 4;
    return B200ZK_OK;
}

// Enqueue one batch (witness generation + proving) on the next free lane.  Host buffers are staged through pinned
// memory owned by the library, so nothing here waits for the GPU and the caller's input buffers are free again on
// return; proofs_out / out_status are filled by b200zk_prove_wait.
static int relation_prove_submit(b200zk_ctx* ctx, const b200zk_pk* pk, int relation, const uint8_t* inputs,
                                 int inputs_on_device, size_t batch, const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                                 uint8_t* out_status, uint64_t* ticket) {
    if (!ctx || !pk || !inputs || !r || !s || !proofs_out || !ticket) return B200ZK_ERR_BAD_ARG;
    if (pk->relation != relation) return fail(ctx, B200ZK_ERR_BAD_ARG, "the proving key belongs to the other relation");
    if (batch == 0) return fail(ctx, B200ZK_ERR_BAD_ARG, "empty batch");
    const int lane = (int)(ctx->next_ticket & 1);
    b200zk_ctx::PendingBatch& pb = ctx->pending[lane];
    if (pb.active)
        return fail(ctx, B200ZK_ERR_BAD_ARG, "two batches are already in flight: b200zk_prove_wait(ticket " +
                                                 std::to_string(pb.ticket) + ") first");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t H = pk->tree_height, nv = pk->num_inputs + pk->num_aux;
    const size_t in_bytes = batch * (size_t)pk->inputs_per_instance * 32;
    B200ZK_TRY(grow_pinned(ctx, (void**)&pb.h_proofs, &pb.h_proofs_cap, batch * 192));
    B200ZK_TRY(grow_pinned(ctx, (void**)&pb.h_status, &pb.h_status_cap, batch * 4));
    B200ZK_TRY(grow_pinned(ctx, (void**)&pb.h_in, &pb.h_in_cap, (inputs_on_device ? 0 : in_bytes) + 2 * batch * 32));
    uint8_t* h_r = pb.h_in;
    uint8_t* h_s = h_r + batch * 32;
    uint8_t* h_inputs = h_s + batch * 32;
    memcpy(h_r, r, batch * 32);
    memcpy(h_s, s, batch * 32);
    if (!inputs_on_device) memcpy(h_inputs, inputs, in_bytes);
    set_lane(ctx, lane);
    struct LaneGuard {  // every other entry point works on lane 0
        b200zk_ctx* c;
        ~LaneGuard() { set_lane(c, 0); }
    } guard{ctx};
    void *din = (void*)inputs, *dz, *dst;
    if (!inputs_on_device) B200ZK_TRY(scratch(ctx, "wit_in", in_bytes, &din));
    B200ZK_TRY(scratch(ctx, "g16_z", batch * (size_t)nv * sizeof(Fr), &dz));
    B200ZK_TRY(scratch(ctx, "wit_status", batch * 4, &dst));
    if (!inputs_on_device)
        B200ZK_CUDA(ctx, cudaMemcpyAsync(din, h_inputs, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(groth16_prove_begin(ctx, pk, batch, h_r, h_s));
    B200ZK_TRY(relation_witness_device(ctx, pk->relation, pk->kind, H, nv, (const Fr*)din, batch, (Fr*)dz, (uint32_t*)dst));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(pb.h_status, dst, batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    // An unsatisfied instance is an error, not a proof (ark: SynthesisError::Unsatisfiable) -- it is reported by
    // b200zk_prove_wait; the batch is proven regardless so that the host never has to wait for the witness here.
    B200ZK_TRY(groth16_prove_enqueue(ctx, pk, (const Fr*)dz, batch, pb.h_proofs, nullptr));
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->lane_done[lane], ctx->stream));
    pb.active = true;
    pb.ticket = ctx->next_ticket++;
    pb.batch = batch;
    pb.proofs_out = proofs_out;
    pb.status_out = out_status;
    *ticket = pb.ticket;
    return B200ZK_OK;
}

static int prove_wait(b200zk_ctx* ctx, uint64_t ticket) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    b200zk_ctx::PendingBatch& pb = ctx->pending[ticket & 1];
    if (!pb.active || pb.ticket != ticket) return fail(ctx, B200ZK_ERR_BAD_ARG, "no batch with this ticket is in flight");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    pb.active = false;
    B200ZK_CUDA(ctx, cudaEventSynchronize(ctx->lane_done[ticket & 1]));
    bool bad = false;
    for (size_t i = 0; i < pb.batch; i++) {
        if (pb.status_out) pb.status_out[i] = (uint8_t)pb.h_status[i];
        if (pb.h_status[i] & 2) return fail(ctx, B200ZK_ERR_BAD_ARG, "witness layout does not match the R1CS (internal error)");
        bad = bad || (pb.h_status[i] & 1);
    }
    // like arkworks, an unsatisfied instance is an error, not a proof (SynthesisError::Unsatisfiable in debug builds)
    if (bad) return fail(ctx, B200ZK_ERR_UNSATISFIED, "a witness does not satisfy the relation");
    memcpy(pb.proofs_out, pb.h_proofs, pb.batch * 192);
    return B200ZK_OK;
}

static int prove_update_note_impl(b200zk_ctx* ctx, const b200zk_pk* pk, const uint8_t* inputs, int inputs_on_device,
                                  size_t batch, const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                                  uint8_t* out_status, int relation = host::RELATION_UPDATE_NOTE) {
    if (!ctx || !pk || !inputs || !r || !s || !proofs_out) return B200ZK_ERR_BAD_ARG;
    if (batch == 0) return B200ZK_OK;
    uint64_t ticket = 0;
    B200ZK_TRY(relation_prove_submit(ctx, pk, relation, inputs, inputs_on_device, batch, r, s, proofs_out, out_status, &ticket));
    return prove_wait(ctx, ticket);
}

int b200zk_update_note_prove_submit(b200zk_ctx* ctx, const b200zk_pk* pk, const void* inputs, int inputs_on_device,
                                    size_t batch, const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                                    uint8_t* out_status, uint64_t* ticket) {
    return relation_prove_submit(ctx, pk, host::RELATION_UPDATE_NOTE, (const uint8_t*)inputs, inputs_on_device, batch, r, s,
                                 proofs_out, out_status, ticket);
}

int b200zk_update_account_prove_submit(b200zk_ctx* ctx, const b200zk_pk* pk, const void* inputs, int inputs_on_device,
                                       size_t batch, const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                                       uint8_t* out_status, uint64_t* ticket) {
    return relation_prove_submit(ctx, pk, host::RELATION_UPDATE_ACCOUNT, (const uint8_t*)inputs, inputs_on_device, batch, r, s,
                                 proofs_out, out_status, ticket);
}

int b200zk_prove_wait(b200zk_ctx* ctx, uint64_t ticket) { return prove_wait(ctx, ticket); }

int b200zk_update_note_prove_batch(b200zk_ctx* ctx, const b200zk_pk* pk, const uint8_t* inputs, size_t batch,
                                   const uint8_t* r, const uint8_t* s, uint8_t* proofs_out, uint8_t* out_status) {
    return prove_update_note_impl(ctx, pk, inputs, 0, batch, r, s, proofs_out, out_status);
}

int b200zk_update_account_prove_batch(b200zk_ctx* ctx, const b200zk_pk* pk, const uint8_t* inputs, size_t batch,
                                      const uint8_t* r, const uint8_t* s, uint8_t* proofs_out, uint8_t* out_status) {
    return prove_update_note_impl(ctx, pk, inputs, 0, batch, r, s, proofs_out, out_status, host::RELATION_UPDATE_ACCOUNT);
}

int b200zk_update_note_prove_batch_device(b200zk_ctx* ctx, const b200zk_pk* pk, const void* d_inputs, size_t batch,
                                          const uint8_t* r, const uint8_t* s, uint8_t* proofs_out,
                                          uint8_t* out_status) {
    return prove_update_note_impl(ctx, pk, (const uint8_t*)d_inputs, 1, batch, r, s, proofs_out, out_status);
}

}  // extern "C"
