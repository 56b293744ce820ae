// K2/K3 -- radix-2 Fr NTT / iNTT / coset NTT as a multi-pass "four-step" transform.
//
// Replaces ark_poly::Radix2EvaluationDomain::{fft,ifft}_in_place and the coset forms
// (ark-poly 0.4.2 [recall]; absent from /root/reference -- SURVEY.md section 8 row a9).
// Natural order in, natural order out, X[k] = sum_j x[j] w^(jk), w = 7^((r-1)/n).
//
// Decomposition (P passes, n = n_0 * n_1 * ... * n_{P-1}, each n_p <= 2^11):
//   pass p views the data as [outer][n_p][inner] and runs, per CTA, G adjacent `inner` columns
//   of the length-n_p transform entirely in shared memory (DIF radix-2 stages, inner twiddles
//   staged in shared memory once per CTA; three stages at a time on 8 elements held in registers,
//   swizzled shared-memory slots), multiplies the result by the inter-pass twiddle
//   w_M^(inner_idx * k), M = n_p * inner (fused into the store), and writes it back in place.
//   The last pass has inner = 1 (rows are contiguous, fully coalesced loads) and scatters its
//   outputs to the digit-reversed natural position, G consecutive rows per CTA so that the
//   scattered stores are G*32 B contiguous.
// One master table T[i] = w_n^i (i < n/2) per domain size serves every twiddle of every pass
// (w_m^i = T[i * n/m], negation for the upper half, mirrored for the inverse).  Coset scaling is
// fused into the first load (forward) or the last store (inverse, together with 1/n).
//
// HBM traffic: one read + one write of the vector per pass; arithmetic (n/2) log2 n butterfly
// multiplications + n per pass boundary.  The kernel is bound by the integer pipe, not by HBM
// (SURVEY.md section 0 item 4), both fractions are reported by bench.py.
#include <cstring>

#include "common.cuh"
#include "field.cuh"

using namespace b200zk;

namespace {

constexpr int MAX_LOG_LEN = 11;       // longest in-CTA transform
constexpr int LOG_TILE = 11;          // elements per CTA tile (len * G)
constexpr int NTT_THREADS = 256;      // 8 elements per thread per fused stage group
constexpr int MAX_PASSES = 4;

struct PassParams {
    const Fr* in;
    Fr* out;
    const Fr* tw;          // master table, n/2 entries
    const Fr* pre_scale;   // n entries or null (first pass, indexed by input position)
    const Fr* post_scale;  // n entries or null (last pass, indexed by output position)
    Fr n_inv;              // used when inverse && !post_scale
    uint32_t log_n, log_len, log_inner, log_g;
    uint32_t lg[MAX_PASSES];  // log sizes of all passes (for the output digit reversal)
    uint32_t n_passes, pass;
    uint32_t inverse;
    uint32_t async_load;   // tile loads as cp.async straight into the swizzled slots (B200ZK_NTT_ASYNC=0: through registers)
    uint64_t rows_total;   // last pass: batch * n / len
};

// 16 bytes global -> shared without passing through registers (LDGSTS); completion: cp.async.wait_all
__device__ __forceinline__ void cp_async16(uint4* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}

__device__ __forceinline__ Fr load_fr(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void store_fr(Fr* p, const Fr& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// shared memory keeps the two 16-byte halves of an element in separate arrays: a warp touching
// consecutive elements then issues conflict-free LDS.128/STS.128.  The slot index is swizzled by
// XOR-folding its 3-bit groups into the low 3 bits, so that 8 lanes whose indices differ by ANY
// power-of-two stride (radix-8 groups, bit-reversed reads, strided twiddle reads) still land in 8
// different 16-byte bank groups; swz is a bijection of every aligned block of 8 slots.
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u); }
__device__ __forceinline__ Fr sload(const uint4* lo, const uint4* hi, uint32_t i) {
    Fr r;
    i = swz(i);
    uint4 a = lo[i], b = hi[i];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sstore(uint4* lo, uint4* hi, uint32_t i, const Fr& r) {
    i = swz(i);
    lo[i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[i] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// R fused DIF radix-2 stages (s .. s+R-1) of the in-CTA transform, 2^R elements per thread in registers:
// one shared-memory round trip and one barrier per R stages instead of per stage, and 2^R - 1 twiddle
// loads per R * 2^(R-1) butterflies.  Element k of a group sits at index (hi * 2^R + k) * q + lo with
// q = len >> (s + R); stage s+t pairs k with k + 2^(R-1-t); the twiddle of the pair is
// w_len^(j << (s+t)) with j = (k mod 2^(R-1-t)) * q + lo, the position inside the half block.
template <int R>
__device__ __forceinline__ void stage_group(uint4* d_lo, uint4* d_hi, const uint4* t_lo, const uint4* t_hi,
                                            uint32_t tid, uint32_t T, uint32_t tile, uint32_t log_g,
                                            uint32_t log_len, uint32_t s) {
    constexpr uint32_t K = 1u << R;
    const uint32_t G = 1u << log_g;
    const uint32_t log_q = log_len - s - R, q = 1u << log_q;
    for (uint32_t item = tid; item < (tile >> R); item += T) {
        const uint32_t g = item & (G - 1), grp = item >> log_g;
        const uint32_t lo = grp & (q - 1), hi = grp >> log_q;
        const uint32_t i0 = (hi << (log_q + R)) + lo;
        Fr x[K];
#pragma unroll
        for (uint32_t k = 0; k < K; k++) x[k] = sload(d_lo, d_hi, ((i0 + (k << log_q)) << log_g) + g);
#pragma unroll
        for (uint32_t t = 0; t < (uint32_t)R; t++) {
            const uint32_t span = K >> (t + 1);
#pragma unroll
            for (uint32_t jj = 0; jj < span; jj++) {
                const uint32_t j = (jj << log_q) + lo;
                Fr w;
                if (j != 0) w = sload(t_lo, t_hi, j << (s + t));
#pragma unroll
                for (uint32_t k = jj; k < K; k += 2 * span) {
                    const Fr u = x[k], v = x[k + span];
                    x[k] = fp_add(u, v);
                    Fr d = fp_sub(u, v);
                    if (j != 0) d = fp_mul(d, w);
                    x[k + span] = d;
                }
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < K; k++) sstore(d_lo, d_hi, ((i0 + (k << log_q)) << log_g) + g, x[k]);
    }
}

// w_n^E for E in [0, n), from the half table; `inverse` mirrors the exponent
__device__ __forceinline__ Fr twiddle(const Fr* tw, uint32_t log_n, uint64_t E, bool inverse) {
    const uint64_t n = 1ull << log_n;
    if (inverse) E = (n - E) & (n - 1);
    const uint64_t half = n >> 1;
    if (E < half) return load_fr(tw + E);
    return fp_neg(load_fr(tw + (E - half)));
}

__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_pass_kernel(PassParams p) {
    extern __shared__ uint4 smem[];
    const uint32_t len = 1u << p.log_len, G = 1u << p.log_g, tile = len * G;
    uint4* d_lo = smem;
    uint4* d_hi = smem + tile;
    uint4* t_lo = d_hi + tile;
    uint4* t_hi = t_lo + (len >> 1);
    const uint32_t tid = threadIdx.x, T = blockDim.x;
    const bool last = p.pass + 1 == p.n_passes;
    const bool inv = p.inverse != 0;
    const uint64_t n = 1ull << p.log_n;

    // ---- inner twiddles w_len^i, i < len/2
    for (uint32_t i = tid; i < (len >> 1); i += T) {
        Fr w = twiddle(p.tw, p.log_n, (uint64_t)i << (p.log_n - p.log_len), inv);
        sstore(t_lo, t_hi, i, w);
    }

    // ---- tile geometry
    uint64_t base = 0;       // non-last: element offset of (b, outer, r=0, inner0)
    uint64_t inner0 = 0;     // non-last: first inner index of the tile
    uint64_t row0 = 0;       // last: first row
    const uint64_t t_id = blockIdx.x;
    if (!last) {
        const uint64_t tiles_per_slab = (1ull << p.log_inner) >> p.log_g;  // per (b, outer)
        const uint64_t slab = t_id / tiles_per_slab;                       // = b * outer_count + outer
        inner0 = (t_id % tiles_per_slab) << p.log_g;
        base = (slab << (p.log_len + p.log_inner)) + inner0;
    } else {
        row0 = t_id << p.log_g;
    }

    // ---- load (smem index = r * G + g)
    // Asynchronous-copy variant of the tile load (the nearest thing to the "TMA-staged tiles" of BASELINE.json that fits
    // this layout: a tensor-map box cannot write the split, XOR-folded 16-byte planes the butterflies read conflict-free,
    // and a linear tile would cost a second pass through shared memory).  Measured (profiles/r02_ntt_async.log): 0.5-1.7 %
    // at 2^16...2^24 (2^22: 0.991 -> 0.986 ms) -- small, because with two CTAs per SM the loads of one tile already hide
    // under the butterflies of the other, but consistent: on by default (B200ZK_NTT_ASYNC=0 restores the register path).
    if (p.async_load && !p.pre_scale) {
        for (uint32_t e = tid; e < tile; e += T) {
            uint64_t addr;
            uint32_t slot;
            if (!last) {
                const uint32_t g = e & (G - 1), r = e >> p.log_g;
                addr = base + ((uint64_t)r << p.log_inner) + g;
                slot = e;
            } else {
                const uint32_t r = e & (len - 1), g = e >> p.log_len;
                const uint64_t row = row0 + g;
                slot = r * G + g;
                if (row >= p.rows_total) {
                    sstore(d_lo, d_hi, slot, Fr::zero());
                    continue;
                }
                addr = (row << p.log_len) + r;
            }
            const uint32_t sw = swz(slot);
            cp_async16(d_lo + sw, reinterpret_cast<const uint4*>(p.in + addr));
            cp_async16(d_hi + sw, reinterpret_cast<const uint4*>(p.in + addr) + 1);
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    } else if (!last) {
        for (uint32_t e = tid; e < tile; e += T) {
            const uint32_t g = e & (G - 1), r = e >> p.log_g;
            const uint64_t addr = base + ((uint64_t)r << p.log_inner) + g;
            Fr x = load_fr(p.in + addr);
            if (p.pre_scale) x = fp_mul(x, load_fr(p.pre_scale + (addr & (n - 1))));
            sstore(d_lo, d_hi, e, x);
        }
    } else {
        for (uint32_t e = tid; e < tile; e += T) {
            const uint32_t r = e & (len - 1), g = e >> p.log_len;
            const uint64_t row = row0 + g;
            Fr x = Fr::zero();
            if (row < p.rows_total) {
                const uint64_t addr = (row << p.log_len) + r;
                x = load_fr(p.in + addr);
                if (p.pre_scale) x = fp_mul(x, load_fr(p.pre_scale + (addr & (n - 1))));
            }
            sstore(d_lo, d_hi, r * G + g, x);
        }
    }
    __syncthreads();

    // ---- DIF radix-2 stages, three at a time in registers
    {
        uint32_t s = 0;
        for (; s + 3 <= p.log_len; s += 3) {
            stage_group<3>(d_lo, d_hi, t_lo, t_hi, tid, T, tile, p.log_g, p.log_len, s);
            __syncthreads();
        }
        if (p.log_len - s == 2) {
            stage_group<2>(d_lo, d_hi, t_lo, t_hi, tid, T, tile, p.log_g, p.log_len, s);
            __syncthreads();
        } else if (p.log_len - s == 1) {
            stage_group<1>(d_lo, d_hi, t_lo, t_hi, tid, T, tile, p.log_g, p.log_len, s);
            __syncthreads();
        }
    }

    // ---- store: position q holds output k = bitrev(q)
    const uint32_t rev_shift = 32 - p.log_len;
    if (!last) {
        const uint32_t log_m = p.log_len + p.log_inner;  // M = len * inner
        for (uint32_t e = tid; e < tile; e += T) {
            const uint32_t g = e & (G - 1), k = e >> p.log_g;
            const uint32_t q = p.log_len ? (__brev(k) >> rev_shift) : 0;
            Fr x = sload(d_lo, d_hi, q * G + g);
            const uint64_t ex = (inner0 + g) * (uint64_t)k;  // < M
            if (ex) x = fp_mul(x, twiddle(p.tw, p.log_n, ex << (p.log_n - log_m), inv));
            store_fr(p.out + base + ((uint64_t)k << p.log_inner) + g, x);
        }
    } else {
        const uint32_t log_rows = p.log_n - p.log_len;  // rows per transform
        for (uint32_t e = tid; e < tile; e += T) {
            const uint32_t g = e & (G - 1), k = e >> p.log_g;
            const uint64_t row = row0 + g;
            if (row >= p.rows_total) continue;
            const uint64_t b = row >> log_rows;
            uint64_t o = row & ((1ull << log_rows) - 1);
            // digit reversal of o = ((k_0 n_1 + k_1) n_2 + ...) -> k_0 + n_0 k_1 + n_0 n_1 k_2 ...
            uint64_t digits[MAX_PASSES];
            for (int qd = (int)p.n_passes - 2; qd >= 0; qd--) {
                digits[qd] = o & ((1ull << p.lg[qd]) - 1);
                o >>= p.lg[qd];
            }
            uint64_t pos = 0;
            uint32_t sh = 0;
            for (uint32_t qd = 0; qd + 1 < p.n_passes; qd++) {
                pos += digits[qd] << sh;
                sh += p.lg[qd];
            }
            pos += (uint64_t)k << log_rows;
            const uint32_t q = p.log_len ? (__brev(k) >> rev_shift) : 0;
            Fr x = sload(d_lo, d_hi, q * G + g);
            if (p.post_scale) x = fp_mul(x, load_fr(p.post_scale + pos));
            else if (inv) x = fp_mul(x, p.n_inv);
            store_fr(p.out + (b << p.log_n) + pos, x);
        }
    }
}


// Exchange step of the multi-GPU four-step transform: out[k][c] = in[c][k] *
// w_n^(+-(row0 + c) * k), a tiled transpose through shared memory with the inter-GPU twiddle fused into
// it.  `in` is rows x cols (row c = one length-`cols` column transform of the global matrix, global
// column index row0 + c), `out` is cols x rows, i.e. chunk-major by destination rank.  The twiddle
// comes from two small tables (w^e = hi[e >> S] * lo[e & (2^S - 1)]) instead of the n/2-entry master
// table, which at n = 2^26 would be 1 GiB per GPU.
constexpr int TT = 32;
__global__ void __launch_bounds__(TT * 8) ntt_twiddle_transpose_kernel(const Fr* __restrict__ in, Fr* __restrict__ out,
                                                                      uint64_t rows, uint64_t cols, uint64_t row0,
                                                                      const Fr* __restrict__ t_lo,
                                                                      const Fr* __restrict__ t_hi, uint32_t S,
                                                                      uint32_t log_n) {
    __shared__ uint4 s_lo[TT][TT + 1], s_hi[TT][TT + 1];
    const uint64_t c0 = (uint64_t)blockIdx.y * TT, k0 = (uint64_t)blockIdx.x * TT;
    const uint32_t tx = threadIdx.x & (TT - 1), ty = threadIdx.x / TT;
    for (uint32_t r = ty; r < TT; r += 8) {
        const uint64_t c = c0 + r, k = k0 + tx;
        if (c < rows && k < cols) {
            Fr x = load_fr(in + c * cols + k);
            const uint64_t e = ((row0 + c) * k) & ((1ull << log_n) - 1);
            if (e) {
                Fr w = fp_mul(load_fr(t_hi + (e >> S)), load_fr(t_lo + (e & ((1ull << S) - 1))));
                x = fp_mul(x, w);
            }
            s_lo[r][tx] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
            s_hi[r][tx] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
        }
    }
    __syncthreads();
    for (uint32_t r = ty; r < TT; r += 8) {
        const uint64_t k = k0 + r, c = c0 + tx;
        if (c < rows && k < cols) {
            uint4* q = reinterpret_cast<uint4*>(out + k * rows + c);
            q[0] = s_lo[tx][r];
            q[1] = s_hi[tx][r];
        }
    }
}

// out[i] = scale * base^i
__global__ void pow_table_kernel(Fr base, Fr scale, Fr* out, uint64_t count) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t e[2] = {(uint32_t)i, (uint32_t)(i >> 32)};
    Fr r = fp_pow(base, e, 2);
    store_fr(out + i, fp_mul(r, scale));
}

const Fr& root_2_32() {  // 7^((r-1)/2^32), Montgomery form
    static const Fr w = {{0x5f0e466au, 0xb9b58d8cu, 0x1819d7ecu, 0x5b1b4c80u, 0x52a31e64u, 0x0af53ae3u,
                          0x19e9b27bu, 0x5bf3addau}};
    return w;
}

Fr host_root_of_unity(uint32_t log_n) {
    Fr w = root_2_32();
    for (uint32_t i = log_n; i < 32; i++) w = fp_sqr(w);
    return w;
}

Fr host_fr_from_u64(uint64_t v) {
    Fr r = Fr::zero();
    r.v[0] = (uint32_t)v;
    r.v[1] = (uint32_t)(v >> 32);
    return fp_to_mont(r);
}

// table key: kind, log size, and the bytes that define the contents written out in full
std::string table_key(const char* kind, uint32_t log_n, const Fr* base = nullptr, int flag = 0) {
    std::string k = std::string(kind) + ":" + std::to_string(log_n) + ":" + std::to_string(flag);
    if (base) {
        static const char* hex = "0123456789abcdef";
        const uint8_t* b = reinterpret_cast<const uint8_t*>(base);
        k += ":";
        for (size_t i = 0; i < sizeof(Fr); i++) {
            k += hex[b[i] >> 4];
            k += hex[b[i] & 15];
        }
    }
    return k;
}

int get_table(b200zk_ctx* ctx, const std::string& key, const Fr& base, const Fr& scale, uint64_t count, const Fr** out) {
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        DeviceBuf b;
        size_t bytes = (size_t)(count ? count : 1) * sizeof(Fr);
        B200ZK_CUDA(ctx, cudaMalloc(&b.ptr, bytes));
        b.bytes = bytes;
        pow_table_kernel<<<div_up(count ? count : 1, 256), 256, 0, ctx->stream>>>(base, scale, (Fr*)b.ptr,
                                                                                  count ? count : 1);
        const int rc = check_launch(ctx, "pow_table_kernel");
        if (rc != B200ZK_OK) {  // never cache a table that was not written
            cudaFree(b.ptr);
            return rc;
        }
        it = ctx->tables.emplace(key, b).first;
    }
    *out = (const Fr*)it->second.ptr;
    return B200ZK_OK;
}

}  // namespace

namespace b200zk {

// device-resident transform; used by the C ABI below and by the Groth16 pipeline
int ntt_device(b200zk_ctx* ctx, Fr* d_data, uint32_t log_n, bool inverse, const Fr* coset_offset, size_t batch) {
    if (log_n > 32) return fail(ctx, B200ZK_ERR_DOMAIN_TOO_LARGE, "log_n > 32");
    if (batch == 0) return B200ZK_OK;
    const uint64_t n = 1ull << log_n;
    const bool coset = coset_offset && !(*coset_offset == Fr::one());
    if (log_n == 0) return B200ZK_OK;  // n = 1: identity in every mode

    const Fr w = host_root_of_unity(log_n);
    const Fr* tw;
    B200ZK_TRY(get_table(ctx, table_key("tw", log_n), w, Fr::one(), n >> 1, &tw));
    const Fr n_inv = fp_inv(host_fr_from_u64(n));
    const Fr *pre = nullptr, *post = nullptr;
    if (coset) {
        if (!inverse) B200ZK_TRY(get_table(ctx, table_key("coset_pre", log_n, coset_offset), *coset_offset, Fr::one(), n, &pre));
        else B200ZK_TRY(get_table(ctx, table_key("coset_post", log_n, coset_offset), fp_inv(*coset_offset), n_inv, n, &post));
    }

    // plan
    uint32_t P = (log_n + MAX_LOG_LEN - 1) / MAX_LOG_LEN;
    if (P > MAX_PASSES) return fail(ctx, B200ZK_ERR_DOMAIN_TOO_LARGE, "too many passes");
    uint32_t lg[MAX_PASSES] = {0, 0, 0, 0};
    for (uint32_t q = 0; q < P; q++) lg[q] = log_n / P + (q < log_n % P ? 1 : 0);

    Fr* tmp = nullptr;
    if (P > 1) {
        void* t;
        B200ZK_TRY(scratch(ctx, "ntt_tmp", batch * n * sizeof(Fr), &t));
        tmp = (Fr*)t;
    }
    const int max_smem = (int)(((1u << LOG_TILE) + (1u << (MAX_LOG_LEN - 1))) * 32);
    B200ZK_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));

    uint32_t log_inner = log_n;
    for (uint32_t q = 0; q < P; q++) {
        log_inner -= lg[q];
        PassParams p;
        p.in = (q == 0) ? d_data : tmp;
        p.out = (q + 1 == P) ? d_data : tmp;
        p.tw = tw;
        p.pre_scale = q == 0 ? pre : nullptr;
        p.post_scale = q + 1 == P ? post : nullptr;
        p.n_inv = n_inv;
        p.log_n = log_n;
        p.log_len = lg[q];
        p.log_inner = log_inner;
        for (int i = 0; i < MAX_PASSES; i++) p.lg[i] = lg[i];
        p.n_passes = P;
        p.pass = q;
        p.inverse = inverse ? 1 : 0;
        static const int async_env = getenv("B200ZK_NTT_ASYNC") ? atoi(getenv("B200ZK_NTT_ASYNC")) : 1;
        p.async_load = async_env ? 1u : 0u;
        const bool last = q + 1 == P;
        uint64_t tiles;
        if (!last) {
            uint32_t lgG = LOG_TILE - lg[q];
            if (lgG > log_inner) lgG = log_inner;
            p.log_g = lgG;
            p.rows_total = 0;
            tiles = (uint64_t)batch << (log_n - lg[q] - lgG);
        } else {
            uint64_t rows = (uint64_t)batch << (log_n - lg[q]);
            uint32_t lgG = LOG_TILE - lg[q];
            while (lgG > 0 && (1ull << lgG) > rows) lgG--;
            // keep G within one transform's rows unless the transform is a single row
            if (log_n - lg[q] > 0 && lgG > log_n - lg[q]) lgG = log_n - lg[q];
            p.log_g = lgG;
            p.rows_total = rows;
            tiles = (rows + (1ull << lgG) - 1) >> lgG;
        }
        if (tiles > 0x7fffffffull) return fail(ctx, B200ZK_ERR_BAD_ARG, "NTT grid too large");
        const size_t smem = (size_t)(((1u << (lg[q] + p.log_g)) + ((1u << lg[q]) >> 1)) * 32);
        {
            ProfScope ps(ctx, "ntt_pass");
            ntt_pass_kernel<<<(unsigned)tiles, NTT_THREADS, smem, ctx->stream>>>(p);
        }
        B200ZK_TRY(check_launch(ctx, "ntt_pass_kernel"));
    }
    return B200ZK_OK;
}

}  // namespace b200zk

extern "C" {

int b200zk_ntt_fr_device(b200zk_ctx* ctx, void* d_data, uint32_t log_n, int inverse, const uint8_t* coset_offset,
                         size_t batch) {
    if (!ctx || !d_data) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr off;
    if (coset_offset) memcpy(&off, coset_offset, sizeof(Fr));
    return ntt_device(ctx, (Fr*)d_data, log_n, inverse != 0, coset_offset ? &off : nullptr, batch);
}

int b200zk_ntt_fr(b200zk_ctx* ctx, uint8_t* data, uint32_t log_n, int inverse, const uint8_t* coset_offset,
                  size_t batch) {
    if (!ctx || !data) return B200ZK_ERR_BAD_ARG;
    if (log_n > 32) return fail(ctx, B200ZK_ERR_DOMAIN_TOO_LARGE, "log_n > 32");
    if (batch == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = (batch << log_n) * sizeof(Fr);
    void* d;
    B200ZK_TRY(scratch(ctx, "ntt_io", bytes, &d));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_TRY(b200zk_ntt_fr_device(ctx, d, log_n, inverse, coset_offset, batch));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_ntt_twiddle_transpose_device(b200zk_ctx* ctx, const void* d_in, void* d_out, uint32_t log_n, uint64_t rows,
                                        uint64_t cols, uint64_t row0, int inverse) {
    if (!ctx || !d_in || !d_out || d_in == d_out) return B200ZK_ERR_BAD_ARG;
    if (log_n > 32) return fail(ctx, B200ZK_ERR_DOMAIN_TOO_LARGE, "log_n > 32");
    if (rows == 0 || cols == 0) return B200ZK_OK;
    if (div_up(rows, TT) > 65535u) return fail(ctx, B200ZK_ERR_BAD_ARG, "too many rows for one launch");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr w = host_root_of_unity(log_n);
    if (inverse) w = fp_inv(w);
    const uint32_t S = (log_n + 1) / 2;
    Fr w_hi = w;
    for (uint32_t i = 0; i < S; i++) w_hi = fp_sqr(w_hi);
    const Fr *t_lo, *t_hi;
    B200ZK_TRY(get_table(ctx, table_key("tt_lo", log_n, nullptr, inverse ? 1 : 0), w, Fr::one(), 1ull << S, &t_lo));
    B200ZK_TRY(get_table(ctx, table_key("tt_hi", log_n, nullptr, inverse ? 1 : 0), w_hi, Fr::one(), 1ull << (log_n - S), &t_hi));
    dim3 grid(div_up(cols, TT), div_up(rows, TT));
    {
        ProfScope ps(ctx, "ntt_twiddle_transpose");
        ntt_twiddle_transpose_kernel<<<grid, TT * 8, 0, ctx->stream>>>((const Fr*)d_in, (Fr*)d_out, rows, cols, row0, t_lo,
                                                                        t_hi, S, log_n);
    }
    return check_launch(ctx, "ntt_twiddle_transpose_kernel");
}

int b200zk_copy2d_device(b200zk_ctx* ctx, void* d_dst, size_t dpitch, const void* d_src, size_t spitch, size_t width,
                         size_t height) {
    if (!ctx || !d_dst || !d_src) return B200ZK_ERR_BAD_ARG;
    if (width == 0 || height == 0) return B200ZK_OK;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    B200ZK_CUDA(ctx, cudaMemcpy2DAsync(d_dst, dpitch, d_src, spitch, width, height, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200ZK_OK;
}

}  // extern "C"
