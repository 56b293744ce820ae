// K1 test kernels (element-wise field ops), integer-pipe micro-benchmarks that fix the
// roofline denominator, and the fixed-base k*G kernel used to make synthetic MSM bases on
// the device (2^24+ CPU scalar multiplications are infeasible).
#include "common.cuh"
#include "ec.cuh"
#include "field_dfma.cuh"
#include "ec_batch_affine.cuh"
#include "fq28.cuh"

using namespace b200zk;

namespace {

template <class F>
__global__ void field_op_kernel(int op, const F* a, const F* b, F* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i];
    F y = b ? b[i] : x;
    F r;
    switch (op) {
        case B200ZK_OP_ADD: r = fp_add(x, y); break;
        case B200ZK_OP_SUB: r = fp_sub(x, y); break;
        case B200ZK_OP_MUL: r = fp_mul(x, y); break;
        case B200ZK_OP_SQR: r = fp_sqr(x); break;
        case B200ZK_OP_INV: r = fp_inv(x); break;
        case B200ZK_OP_MUL_DFMA:  // experiment: the same product on the FP64 pipe (Fq only)
            if constexpr (sizeof(F) == sizeof(Fq)) r = dfma::fq_mul_dfma(x, y);
            else r = fp_mul(x, y);
            break;
        default: r = x; break;
    }
    out[i] = r;
}

template <class C>
__global__ void mont_conv_kernel(int op, const Fp<C>* a, Fp<C>* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = op == B200ZK_OP_TO_MONT ? fp_to_mont(a[i]) : fp_from_mont(a[i]);
}

template <class F>
int run_field_op(b200zk_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    size_t bytes = n * sizeof(F);
    void *da, *db, *dout;
    B200ZK_TRY(scratch(ctx, "dbg_a", bytes, &da));
    B200ZK_TRY(scratch(ctx, "dbg_b", bytes, &db));
    B200ZK_TRY(scratch(ctx, "dbg_o", bytes, &dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (b) B200ZK_CUDA(ctx, cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
    field_op_kernel<F><<<div_up(n, 128), 128, 0, ctx->stream>>>(op, (const F*)da, b ? (const F*)db : nullptr,
                                                                (F*)dout, n);
    B200ZK_TRY(check_launch(ctx, "field_op_kernel"));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

template <class C>
int run_mont_conv(b200zk_ctx* ctx, int op, const uint8_t* a, uint8_t* out, size_t n) {
    size_t bytes = n * sizeof(Fp<C>);
    void *da, *dout;
    B200ZK_TRY(scratch(ctx, "dbg_a", bytes, &da));
    B200ZK_TRY(scratch(ctx, "dbg_o", bytes, &dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    mont_conv_kernel<C><<<div_up(n, 128), 128, 0, ctx->stream>>>(op, (const Fp<C>*)da, (Fp<C>*)dout, n);
    B200ZK_TRY(check_launch(ctx, "mont_conv_kernel"));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

// ---------------------------------------------------------------- micro-benchmarks
constexpr int PEAK_ITERS = 2048;

__global__ void peak_imad32(uint32_t* out, uint32_t x, uint32_t y) {
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x + k;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(x), "r"(y));
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= a[k];
    if (s == 0x12345678u) out[0] = s;
}

// 16 INDEPENDENT accumulator chains per thread; the multiplicands are loop constants, so the only dependence is the
// addend (and the carry) of each chain -- the shape of a Montgomery row.  (The round-1 version fed the low word of the
// accumulator back as a multiplicand: a longer dependence than any field product has, and it measured 7.3 T/s, below
// the 9.1 T/s the Fq product itself sustains.)
__global__ void peak_imad_wide(uint64_t* out, uint32_t x) {
    uint64_t a[16];
    uint32_t m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        a[k] = threadIdx.x + k;
        m[k] = x + 2 * k + 1;
    }
    for (int it = 0; it < PEAK_ITERS / 2; it++) {
        // the second factor changes every iteration: with a loop-invariant product ptxas hoists the multiplication and
        // the loop degenerates into 64-bit additions (that is what the first version of this kernel measured)
        const uint32_t y = x + 0x9e3779b9u * (uint32_t)it;
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(m[k]), "r"(y));
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= a[k];
    if (s == 0x12345678ull) out[0] = s;
}

// carry-CHAINED wide multiply-adds (mad.lo.cc / madc.hi.cc rows as the Montgomery product issues them = IMAD.WIDE.U32
// with carry out and IMAD.WIDE.U32.X with carry in), four independent accumulators per thread: is the chained form
// issued at the rate of the plain IMAD.WIDE, or is the carry what halves the field product's multiplier use?
__global__ void peak_imad_wide_carry(uint32_t* out, const uint32_t* in) {
    uint32_t X[4][13], a[12];
#pragma unroll
    for (int k = 0; k < 12; k++) a[k] = in[k] + threadIdx.x;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 13; k++) X[c][k] = threadIdx.x + c + k;
#if defined(__CUDA_ARCH__)
    for (int it = 0; it < PEAK_ITERS / 4; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) ptx::row_even<12>(X[c], a, a[c] + it);   // 6 chained wide multiply-adds each
    }
#endif
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 13; k++) s ^= X[c][k];
    if (s == 0x12345678u) out[0] = s;
}

// two INDEPENDENT Fq product chains per thread in one instruction stream (what interleaving two additions, or the
// two halves of an Fq2 product, would give the scheduler)
__global__ void peak_fqmul_ilp2(Fq* out, const Fq* in) {
    Fq a = in[0], b = in[1], c = in[1], d = in[0];
    a.v[0] ^= threadIdx.x;
    c.v[1] ^= threadIdx.x;
    for (int it = 0; it < PEAK_ITERS / 2; it++) {
        a = fp_mul(a, b);
        c = fp_mul(c, d);
        b = fp_mul(b, a);
        d = fp_mul(d, c);
    }
    if (a.v[0] == 0x12345678u && b.v[1] == 0x9abcdef0u && c.v[2] == 1u && d.v[3] == 2u) out[0] = a;
}

// back-to-back Fq products in reduced radix (fq28.cuh: 14 x 28-bit limbs, carry-free IMAD.WIDE columns)
__global__ void peak_fq28mul(r28::Fq28* out, const uint32_t* in) {
    r28::Fq28 a, b;
#pragma unroll
    for (int k = 0; k < r28::NL; k++) {
        a.l[k] = (in[k] + threadIdx.x) & r28::LM;
        b.l[k] = (in[k + 1] + 3 * threadIdx.x) & r28::LM;
    }
    a.l[r28::NL - 1] &= 0xffff;
    b.l[r28::NL - 1] &= 0xffff;
    for (int it = 0; it < PEAK_ITERS / 2; it++) {
        a = r28::mul(a, b);
        b = r28::mul(b, a);
    }
    if (a.l[0] == 0x02345678u && b.l[1] == 0x0abcdef0u) out[0] = a;
}

__global__ void peak_dfma(double* out, double x, double y) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x + k;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fma(a[k], x, y);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    if (s == 0.123) out[0] = s;
}

template <class C>
__global__ void peak_fpmul(Fp<C>* out, const Fp<C>* in) {
    Fp<C> a = in[0], b = in[1];
    a.v[0] ^= threadIdx.x;  // still < p for the inputs we pass (top limb untouched)
    for (int it = 0; it < PEAK_ITERS / 2; it++) {
        a = fp_mul(a, b);
        b = fp_mul(b, a);
    }
    if (a.v[0] == 0x12345678u && b.v[1] == 0x9abcdef0u) out[0] = a;
}

// Fq products on the FP64 pipe (field_dfma.cuh), same chain as peak_fpmul
__global__ void peak_fqmul_dfma(Fq* out, const Fq* in) {
    Fq a = in[0], b = in[1];
    a.v[0] ^= threadIdx.x;
    for (int it = 0; it < PEAK_ITERS / 2; it++) {
        a = dfma::fq_mul_dfma(a, b);
        b = dfma::fq_mul_dfma(b, a);
    }
    if (a.v[0] == 0x12345678u && b.v[1] == 0x9abcdef0u) out[0] = a;
}

// two independent chains per thread, one on each pipe: what a kernel that splits its products between the
// integer multiplier and the FP64 unit could sustain
__global__ void peak_fqmul_both(Fq* out, const Fq* in) {
    Fq a = in[0], b = in[1], c = in[1], d = in[0];
    a.v[0] ^= threadIdx.x;
    c.v[0] ^= threadIdx.x;
    for (int it = 0; it < PEAK_ITERS / 2; it++) {
        a = fp_mul(a, b);
        c = dfma::fq_mul_dfma(c, d);
        b = fp_mul(b, a);
        d = dfma::fq_mul_dfma(d, c);
    }
    if (a.v[0] == 0x12345678u && b.v[1] == 0x9abcdef0u && c.v[2] == 1u && d.v[3] == 2u) out[0] = a;
}

// warp-specialised: even warps keep the integer multiplier busy, odd warps the FP64 unit (2 CTAs/SM = 8 + 8 warps)
__global__ void __launch_bounds__(256, 2) peak_fqmul_split(Fq* out, const Fq* in) {
    Fq a = in[0], b = in[1];
    a.v[0] ^= threadIdx.x;
    if ((threadIdx.x >> 5) & 1) {
        for (int it = 0; it < PEAK_ITERS / 2; it++) {
            a = dfma::fq_mul_dfma(a, b);
            b = dfma::fq_mul_dfma(b, a);
        }
    } else {
        for (int it = 0; it < PEAK_ITERS / 2; it++) {
            a = fp_mul(a, b);
            b = fp_mul(b, a);
        }
    }
    if (a.v[0] == 0x12345678u && b.v[1] == 0x9abcdef0u) out[0] = a;
}

// ---------------------------------------------------------------- batched affine addition (groundwork)
// thread t adds the pairs [t * chunk, (t + 1) * chunk) with one inversion; pre[] lives in local memory
constexpr int BA_MAX_CHUNK = 64;
template <class F>
__global__ void __launch_bounds__(128) batch_add_affine_kernel(const Affine<F>* __restrict__ p, const Affine<F>* __restrict__ q,
                                                               size_t n, int chunk, Affine<F>* __restrict__ out) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t lo = t * (size_t)chunk;
    if (lo >= n) return;
    const int m = (int)(n - lo < (size_t)chunk ? n - lo : (size_t)chunk);
    F pre[BA_MAX_CHUNK];
    ec_batch_add_affine(p + lo, q + lo, out + lo, m, pre);
}

// the same additions the way the bucket accumulation does them today: XYZZ accumulator += affine, then to affine
// is NOT included (the accumulation never leaves XYZZ); one mixed addition per pair, result discarded into a sum
template <class F>
__global__ void __launch_bounds__(128) xyzz_madd_kernel(const Affine<F>* __restrict__ p, const Affine<F>* __restrict__ q,
                                                        size_t n, int chunk, XYZZ<F>* __restrict__ out) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t lo = t * (size_t)chunk;
    if (lo >= n) return;
    const int m = (int)(n - lo < (size_t)chunk ? n - lo : (size_t)chunk);
    XYZZ<F> acc = XYZZ<F>::from_affine(p[lo]);
    for (int i = 0; i < m; i++) ec_madd(acc, q[lo + i]);
    out[t] = acc;
}

template <class F>
int run_batch_add(b200zk_ctx* ctx, const uint8_t* p, const uint8_t* q, size_t n, int chunk, uint8_t* out, double* ms_batch,
                  double* ms_xyzz) {
    const size_t bytes = n * sizeof(Affine<F>);
    void *dp, *dq, *dout, *dx;
    B200ZK_TRY(scratch(ctx, "dbg_a", bytes, &dp));
    B200ZK_TRY(scratch(ctx, "dbg_b", bytes, &dq));
    B200ZK_TRY(scratch(ctx, "dbg_o", bytes, &dout));
    const size_t threads = div_up(n, (size_t)chunk);
    B200ZK_TRY(scratch(ctx, "dbg_x", threads * sizeof(XYZZ<F>), &dx));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dp, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dq, q, bytes, cudaMemcpyHostToDevice, ctx->stream));
    cudaEvent_t e[3];
    for (auto& ev : e) B200ZK_CUDA(ctx, cudaEventCreate(&ev));
    float best_b = 1e30f, best_x = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        B200ZK_CUDA(ctx, cudaEventRecord(e[0], ctx->stream));
        batch_add_affine_kernel<F><<<div_up(threads, 128), 128, 0, ctx->stream>>>((const Affine<F>*)dp, (const Affine<F>*)dq, n,
                                                                                 chunk, (Affine<F>*)dout);
        B200ZK_TRY(check_launch(ctx, "batch_add_affine_kernel"));
        B200ZK_CUDA(ctx, cudaEventRecord(e[1], ctx->stream));
        xyzz_madd_kernel<F><<<div_up(threads, 128), 128, 0, ctx->stream>>>((const Affine<F>*)dp, (const Affine<F>*)dq, n, chunk,
                                                                          (XYZZ<F>*)dx);
        B200ZK_TRY(check_launch(ctx, "xyzz_madd_kernel"));
        B200ZK_CUDA(ctx, cudaEventRecord(e[2], ctx->stream));
        B200ZK_CUDA(ctx, cudaEventSynchronize(e[2]));
        float a = 0, b = 0;
        B200ZK_CUDA(ctx, cudaEventElapsedTime(&a, e[0], e[1]));
        B200ZK_CUDA(ctx, cudaEventElapsedTime(&b, e[1], e[2]));
        best_b = a < best_b ? a : best_b;
        best_x = b < best_x ? b : best_x;
    }
    for (auto& ev : e) cudaEventDestroy(ev);
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ms_batch) *ms_batch = best_b;
    if (ms_xyzz) *ms_xyzz = best_x;
    return B200ZK_OK;
}

// ---------------------------------------------------------------- fixed-base multiplication
template <class F>
__global__ void fixed_base_kernel(const Affine<F>* gen, const uint32_t* scalars, size_t n, Affine<F>* out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = scalars[i * 8 + j];
    Affine<F> g = *gen;
    XYZZ<F> acc = XYZZ<F>::inf();
    // double-and-add with mixed additions of the affine generator
    for (int limb = 7; limb >= 0; limb--)
        for (int bit = 31; bit >= 0; bit--) {
            acc = ec_dbl(acc);
            if ((k[limb] >> bit) & 1) ec_madd(acc, g);
        }
    out[i] = ec_to_affine(acc);
}

// generators in Montgomery form are produced on the host from canonical constants (see gen_*)
struct GenCache {
    bool ready = false;
    G1Affine g1;
    G2Affine g2;
};

Fq fq_from_hex_be(const char* hex) {  // canonical big-endian hex (96 digits) -> Montgomery
    Fq r = Fq::zero();
    for (int i = 0; i < 96; i++) {
        char c = hex[i];
        uint32_t d = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
        int bitpos = (95 - i) * 4;
        r.v[bitpos / 32] |= d << (bitpos % 32);
    }
    return fp_to_mont(r);
}

const GenCache& generators() {
    static GenCache g;
    if (!g.ready) {
        g.g1.x = fq_from_hex_be("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
        g.g1.y = fq_from_hex_be("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1");
        g.g2.x.c0 = fq_from_hex_be("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8");
        g.g2.x.c1 = fq_from_hex_be("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e");
        g.g2.y.c0 = fq_from_hex_be("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801");
        g.g2.y.c1 = fq_from_hex_be("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be");
        g.ready = true;
    }
    return g;
}

template <class F>
int fixed_base_device(b200zk_ctx* ctx, const Affine<F>& gen, const void* d_scalars, size_t n, void* d_out) {
    void* dgen;
    B200ZK_TRY(scratch(ctx, sizeof(F) == sizeof(Fq) ? "gen_g1" : "gen_g2", sizeof(Affine<F>), &dgen));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dgen, &gen, sizeof(Affine<F>), cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `gen` may be a temporary
    fixed_base_kernel<F><<<div_up(n, 64), 64, 0, ctx->stream>>>((const Affine<F>*)dgen, (const uint32_t*)d_scalars, n,
                                                               (Affine<F>*)d_out);
    return check_launch(ctx, "fixed_base_kernel");
}

}  // namespace

extern "C" {

int b200zk_dbg_field_op(b200zk_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out,
                        size_t n) {
    if (!ctx || !a || !out) return B200ZK_ERR_BAD_ARG;
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    bool conv = op == B200ZK_OP_TO_MONT || op == B200ZK_OP_FROM_MONT;
    bool binary = op == B200ZK_OP_ADD || op == B200ZK_OP_SUB || op == B200ZK_OP_MUL || op == B200ZK_OP_MUL_DFMA;
    if (binary && !b) return fail(ctx, B200ZK_ERR_BAD_ARG, "binary op needs b");
    if (!binary) b = nullptr;
    switch (field) {
        case B200ZK_FIELD_FR: return conv ? run_mont_conv<FrCfg>(ctx, op, a, out, n) : run_field_op<Fr>(ctx, op, a, b, out, n);
        case B200ZK_FIELD_FQ: return conv ? run_mont_conv<FqCfg>(ctx, op, a, out, n) : run_field_op<Fq>(ctx, op, a, b, out, n);
        case B200ZK_FIELD_FQ2:
            if (conv) return fail(ctx, B200ZK_ERR_BAD_ARG, "convert Fq2 as two Fq");
            return run_field_op<Fq2>(ctx, op, a, b, out, n);
        default: return fail(ctx, B200ZK_ERR_BAD_ARG, "unknown field");
    }
}

int b200zk_dbg_batch_add_affine(b200zk_ctx* ctx, int group, const uint8_t* p, const uint8_t* q, size_t n, int chunk,
                                uint8_t* out, double* ms_batch, double* ms_xyzz) {
    if (!ctx || !p || !q || !out) return B200ZK_ERR_BAD_ARG;
    if (chunk < 1 || chunk > BA_MAX_CHUNK) return fail(ctx, B200ZK_ERR_BAD_ARG, "chunk must be 1..64");
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (group == 1) return run_batch_add<Fq>(ctx, p, q, n, chunk, out, ms_batch, ms_xyzz);
    if (group == 2) return run_batch_add<Fq2>(ctx, p, q, n, chunk, out, ms_batch, ms_xyzz);
    return fail(ctx, B200ZK_ERR_BAD_ARG, "group must be 1 or 2");
}

int b200zk_dbg_int_peak(b200zk_ctx* ctx, int kind, double* ops_per_sec) {
    if (!ctx || !ops_per_sec) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    void* dbuf;
    B200ZK_TRY(scratch(ctx, "peak", 4096, &dbuf));
    Fq two[2];
    two[0] = Fq::one();
    two[1] = fp_add(Fq::one(), Fq::one());
    Fr twor[2];
    twor[0] = Fr::one();
    twor[1] = fp_add(Fr::one(), Fr::one());
    if (kind == 2) B200ZK_CUDA(ctx, cudaMemcpyAsync(dbuf, twor, sizeof(twor), cudaMemcpyHostToDevice, ctx->stream));
    if (kind == 3 || kind == 5 || kind == 6 || kind == 7 || kind == 8 || kind == 9 || kind == 10)
        B200ZK_CUDA(ctx, cudaMemcpyAsync(dbuf, two, sizeof(two), cudaMemcpyHostToDevice, ctx->stream));
    const int threads = 256;
    const int blocks = ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    B200ZK_CUDA(ctx, cudaEventCreate(&e0));
    B200ZK_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {  // rep 0 is the warm-up
        B200ZK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        double ops = 0;
        switch (kind) {
            case 0:
                peak_imad32<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)dbuf + 512, 3u + rep, 7u);
                ops = 8.0 * PEAK_ITERS;
                break;
            case 1:
                peak_imad_wide<<<blocks, threads, 0, ctx->stream>>>((uint64_t*)dbuf + 256, 3u + rep);
                ops = 8.0 * PEAK_ITERS;
                break;
            case 2:
                peak_fpmul<FrCfg><<<blocks, threads, 0, ctx->stream>>>((Fr*)dbuf + 16, (const Fr*)dbuf);
                ops = PEAK_ITERS;
                break;
            case 3:
                peak_fpmul<FqCfg><<<blocks, threads, 0, ctx->stream>>>((Fq*)dbuf + 16, (const Fq*)dbuf);
                ops = PEAK_ITERS;
                break;
            case 4:
                peak_dfma<<<blocks, threads, 0, ctx->stream>>>((double*)dbuf + 256, 1.0000001, 1e-9);
                ops = 8.0 * PEAK_ITERS;
                break;
            case 5:
                peak_fqmul_dfma<<<blocks, threads, 0, ctx->stream>>>((Fq*)dbuf + 16, (const Fq*)dbuf);
                ops = PEAK_ITERS;
                break;
            case 6:
                peak_fqmul_both<<<blocks, threads, 0, ctx->stream>>>((Fq*)dbuf + 16, (const Fq*)dbuf);
                ops = 2.0 * PEAK_ITERS;
                break;
            case 7:
                peak_fqmul_split<<<blocks, threads, 0, ctx->stream>>>((Fq*)dbuf + 16, (const Fq*)dbuf);
                ops = PEAK_ITERS;
                break;
            case 8:
                peak_imad_wide_carry<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)dbuf + 512, (const uint32_t*)dbuf);
                ops = 6.0 * PEAK_ITERS;          // 4 rows of 6 wide multiply-adds per iteration, PEAK_ITERS / 4 iterations
                break;
            case 9:
                peak_fqmul_ilp2<<<blocks, threads, 0, ctx->stream>>>((Fq*)dbuf + 16, (const Fq*)dbuf);
                ops = 2.0 * PEAK_ITERS;
                break;
            case 10:
                peak_fq28mul<<<blocks, threads, 0, ctx->stream>>>((r28::Fq28*)((uint32_t*)dbuf + 512), (const uint32_t*)dbuf);
                ops = PEAK_ITERS;
                break;
            default: return fail(ctx, B200ZK_ERR_BAD_ARG, "unknown kind");
        }
        B200ZK_TRY(check_launch(ctx, "peak kernel"));
        B200ZK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        B200ZK_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        B200ZK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double rate = ops * threads * (double)blocks / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_sec = best;
    return B200ZK_OK;
}

int b200zk_fixed_base_mul_device(b200zk_ctx* ctx, int group, const void* d_scalars, size_t n, void* d_out) {
    if (!ctx || !d_scalars || !d_out) return B200ZK_ERR_BAD_ARG;
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (group == 1) return fixed_base_device<Fq>(ctx, generators().g1, d_scalars, n, d_out);
    if (group == 2) return fixed_base_device<Fq2>(ctx, generators().g2, d_scalars, n, d_out);
    return fail(ctx, B200ZK_ERR_BAD_ARG, "group must be 1 or 2");
}

int b200zk_fixed_base_mul(b200zk_ctx* ctx, int group, const uint8_t* scalars, size_t n, uint8_t* out_points) {
    if (!ctx || !scalars || !out_points) return B200ZK_ERR_BAD_ARG;
    if (group != 1 && group != 2) return fail(ctx, B200ZK_ERR_BAD_ARG, "group must be 1 or 2");
    if (n == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t pt = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    void *ds, *dout;
    B200ZK_TRY(scratch(ctx, "fb_s", n * 32, &ds));
    B200ZK_TRY(scratch(ctx, "fb_o", n * pt, &dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(b200zk_fixed_base_mul_device(ctx, group, ds, n, dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_points, dout, n * pt, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

}  // extern "C"
