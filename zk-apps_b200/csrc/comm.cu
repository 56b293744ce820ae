// Multi-GPU paths behind the C ABI (include/b200zk.h "multi-GPU"): one ctx per GPU, one
// NCCL communicator attached to it, and the two sharded hot paths enqueued on the ctx stream end to end --
// no host synchronisation between the local kernels, the collective and the combine step.
//
//   * large MSM: every rank runs the full bucket pipeline on its contiguous slice of the points; msm_finish writes
//     the rank's affine partial STRAIGHT INTO ITS SLOT of the all_gather buffer; ncclAllGather in place on the same
//     stream (world * 96 B for G1, 192 B for G2; curve addition is not an NCCL reduction op); points_sum_kernel
//     adds the world partials on every rank.
//   * large NTT: four-step transform with ONE exchange (column transforms, twiddle fused into a tiled transpose
//     that leaves the data chunk-major by destination rank, grouped ncclSend/ncclRecv all-to-all, interleave,
//     row transforms).
//
// NCCL is bound at run time (dlopen "libnccl.so.2", resolved by SONAME: inside a process that already loaded
// torch's bundled NCCL the same copy is reused; a Rust host gets the system one), so libb200zk.so has no link-time
// dependency on it and single-GPU users never load it.  The few NCCL declarations needed are restated here
// (nccl.h 2.27: opaque comm pointer, 128-byte unique id, ncclUint8 = 1).  No reference counterpart: the
// reference has no communication backend at all.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "types.cuh"

using namespace b200zk;

namespace {

typedef void* nccl_comm_t;
struct nccl_uid {
    char internal[128];
};
enum { NCCL_UINT8 = 1 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("B200ZK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
            return;
        }
        auto sym = [&](const char* s) {
            void* p = dlsym(api.lib, s);
            if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s;
            return p;
        };
        api.GetUniqueId = (int (*)(nccl_uid*))sym("ncclGetUniqueId");
        api.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_uid, int))sym("ncclCommInitRank");
        api.CommInitAll = (int (*)(nccl_comm_t*, int, const int*))sym("ncclCommInitAll");
        api.CommDestroy = (int (*)(nccl_comm_t))sym("ncclCommDestroy");
        api.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))sym("ncclAllGather");
        api.Send = (int (*)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclSend");
        api.Recv = (int (*)(void*, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclRecv");
        api.GroupStart = (int (*)())sym("ncclGroupStart");
        api.GroupEnd = (int (*)())sym("ncclGroupEnd");
        api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    });
    return api;
}

int nccl_ready(b200zk_ctx* ctx) {
    NcclApi& a = nccl();
    if (!a.error.empty()) return fail(ctx, B200ZK_ERR_NCCL, a.error);
    return B200ZK_OK;
}

#define B200ZK_NCCL(ctx, expr)                                                                                    \
    do {                                                                                                          \
        int _r = (expr);                                                                                          \
        if (_r != 0)                                                                                              \
            return ::b200zk::fail(ctx, B200ZK_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(_r)); \
    } while (0)

int need_comm(b200zk_ctx* ctx) {
    if (!ctx->nccl_comm && ctx->comm_world != 1)
        return fail(ctx, B200ZK_ERR_BAD_ARG, "no communicator: call b200zk_comm_init first");
    return B200ZK_OK;
}

// ---- phases of the sharded MSM (split so that the single-process variant can group the collective)
template <class F>
int msm_sharded_local(b200zk_ctx* ctx, const b200zk_bases* h, const void* scalars, int on_device, size_t n, void** gather) {
    const size_t pt = sizeof(Affine<F>);
    const int world = ctx->comm_world;
    B200ZK_TRY(scratch(ctx, sizeof(F) == sizeof(Fq) ? "comm_gather_g1" : "comm_gather_g2", (size_t)world * pt, gather));
    const uint32_t* ds = (const uint32_t*)scalars;
    if (!on_device && n) {
        void* d;
        B200ZK_TRY(scratch(ctx, "msm_scalars", std::max<size_t>(32, n * 32), &d));
        B200ZK_CUDA(ctx, cudaMemcpyAsync(d, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
        ds = (const uint32_t*)d;
    }
    // the partial lands in this rank's slot of the gather buffer (msm_finish writes it there)
    return msm_device<F>(ctx, h, ds, n, n, 1, false, (Affine<F>*)((uint8_t*)*gather + (size_t)ctx->comm_rank * pt));
}

int msm_sharded_collective(b200zk_ctx* ctx, void* gather, size_t pt) {
    if (ctx->comm_world == 1) return B200ZK_OK;
    ProfScope ps(ctx, "comm_all_gather");
    B200ZK_NCCL(ctx, nccl().AllGather((const uint8_t*)gather + (size_t)ctx->comm_rank * pt, gather, pt, NCCL_UINT8,
                                      (nccl_comm_t)ctx->nccl_comm, ctx->stream));
    return B200ZK_OK;
}

int msm_sharded_run(b200zk_ctx* ctx, const b200zk_bases* h, const void* scalars, int on_device, size_t n, void* d_out) {
    if (!ctx || !h || !d_out || (n && !scalars)) return B200ZK_ERR_BAD_ARG;
    if (n > h->n) return fail(ctx, B200ZK_ERR_BAD_LEN, "more scalars than bases");
    B200ZK_TRY(need_comm(ctx));
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    void* gather = nullptr;
    const size_t pt = h->group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    if (h->group == 1) B200ZK_TRY(msm_sharded_local<Fq>(ctx, h, scalars, on_device, n, &gather));
    else B200ZK_TRY(msm_sharded_local<Fq2>(ctx, h, scalars, on_device, n, &gather));
    B200ZK_TRY(msm_sharded_collective(ctx, gather, pt));
    ProfScope ps(ctx, "comm_points_sum");
    return b200zk_points_sum_device(ctx, h->group, gather, (size_t)ctx->comm_world, d_out);
}

// out[r][src * C + c] = recv[src][r][c]: the chunks of the all-to-all interleaved into rows (one launch instead
// of `world` strided copies)
__global__ void ntt_interleave_kernel(const uint4* __restrict__ recv, uint4* __restrict__ out, uint64_t R, uint64_t C,
                                      uint32_t world) {
    const uint64_t total = (uint64_t)world * R * C * 2;  // 16-byte halves
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t half = i & 1, e = i >> 1;
        const uint64_t c = e % C, r = (e / C) % R, src = e / (C * R);
        out[((r * world + src) * C + c) * 2 + half] = recv[i];
    }
}

}  // namespace

extern "C" {

int b200zk_comm_unique_id(uint8_t id_out[128]) {
    if (!id_out) return B200ZK_ERR_BAD_ARG;
    NcclApi& a = nccl();
    if (!a.error.empty()) return B200ZK_ERR_NCCL;
    nccl_uid id;
    if (a.GetUniqueId(&id) != 0) return B200ZK_ERR_NCCL;
    memcpy(id_out, id.internal, 128);
    return B200ZK_OK;
}

int b200zk_comm_init(b200zk_ctx* ctx, const uint8_t id[128], int rank, int world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) return B200ZK_ERR_BAD_ARG;
    if (ctx->nccl_comm) return fail(ctx, B200ZK_ERR_BAD_ARG, "the ctx already has a communicator");
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    if (world == 1) return B200ZK_OK;  // nothing to talk to: the sharded entry points degenerate to the local ones
    B200ZK_TRY(nccl_ready(ctx));
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid uid;
    memcpy(uid.internal, id, 128);
    nccl_comm_t comm = nullptr;
    int r = nccl().CommInitRank(&comm, world, uid, rank);
    if (r != 0) {
        ctx->comm_rank = 0;
        ctx->comm_world = 1;
        return fail(ctx, B200ZK_ERR_NCCL, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
    }
    ctx->nccl_comm = comm;
    return B200ZK_OK;
}

int b200zk_comm_destroy(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        nccl().CommDestroy((nccl_comm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    return B200ZK_OK;
}

int b200zk_comm_info(const b200zk_ctx* ctx, int* rank, int* world) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->comm_world;
    return B200ZK_OK;
}

int b200zk_init_multi(int n_gpus, b200zk_ctx** out) {
    if (!out || n_gpus < 1) return B200ZK_ERR_BAD_ARG;
    for (int i = 0; i < n_gpus; i++) out[i] = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return B200ZK_ERR_NO_DEVICE;
    if (n_gpus > count) return B200ZK_ERR_BAD_ARG;
    int rc = B200ZK_OK;
    for (int i = 0; i < n_gpus && rc == B200ZK_OK; i++) rc = b200zk_init(i, &out[i]);
    std::vector<nccl_comm_t> comms((size_t)n_gpus, nullptr);
    if (rc == B200ZK_OK && n_gpus > 1) {
        NcclApi& a = nccl();
        if (!a.error.empty()) rc = fail(out[0], B200ZK_ERR_NCCL, a.error);
        else {
            std::vector<int> devs((size_t)n_gpus);
            for (int i = 0; i < n_gpus; i++) devs[(size_t)i] = i;
            const int r = a.CommInitAll(comms.data(), n_gpus, devs.data());
            if (r != 0) rc = fail(out[0], B200ZK_ERR_NCCL, std::string("ncclCommInitAll: ") + a.GetErrorString(r));
        }
    }
    if (rc != B200ZK_OK) {
        for (int i = 0; i < n_gpus; i++) {
            if (out[i]) b200zk_destroy(out[i]);
            out[i] = nullptr;
        }
        return rc;
    }
    for (int i = 0; i < n_gpus; i++) {
        out[i]->comm_rank = i;
        out[i]->comm_world = n_gpus;
        out[i]->nccl_comm = comms[(size_t)i];
    }
    return B200ZK_OK;
}

int b200zk_msm_sharded_device(b200zk_ctx* ctx, const b200zk_bases* h_local, const void* scalars_local, int scalars_on_device,
                              size_t n_local, void* d_out_affine) {
    return msm_sharded_run(ctx, h_local, scalars_local, scalars_on_device, n_local, d_out_affine);
}

int b200zk_msm_sharded(b200zk_ctx* ctx, const b200zk_bases* h_local, const void* scalars_local, int scalars_on_device,
                       size_t n_local, uint8_t* out_affine, uint8_t* out_is_inf) {
    if (!ctx || !h_local || !out_affine) return B200ZK_ERR_BAD_ARG;
    const size_t pt = h_local->group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    void* dout;
    B200ZK_TRY(scratch(ctx, "comm_msm_out", pt, &dout));
    B200ZK_TRY(msm_sharded_run(ctx, h_local, scalars_local, scalars_on_device, n_local, dout));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(out_affine, dout, pt, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_is_inf) {
        bool z = true;
        for (size_t i = 0; i < pt; i++) z = z && out_affine[i] == 0;
        *out_is_inf = z ? 1 : 0;
    }
    return B200ZK_OK;
}

// Single-process form over the contexts of b200zk_init_multi: the collective of all ranks is issued inside one
// NCCL group by the calling thread (a thread driving several communicators must group their calls).
int b200zk_msm_sharded_multi(b200zk_ctx** ctxs, int n_gpus, const b200zk_bases** h_local, const void** scalars_local,
                             int scalars_on_device, const size_t* n_local, uint8_t* out_affine, uint8_t* out_is_inf) {
    if (!ctxs || n_gpus < 1 || !h_local || !scalars_local || !n_local || !out_affine) return B200ZK_ERR_BAD_ARG;
    b200zk_ctx* c0 = ctxs[0];
    for (int i = 0; i < n_gpus; i++)
        if (!ctxs[i] || !h_local[i] || ctxs[i]->comm_world != n_gpus || ctxs[i]->comm_rank != i || h_local[i]->group != h_local[0]->group)
            return fail(c0, B200ZK_ERR_BAD_ARG, "msm_sharded_multi: contexts must come from b200zk_init_multi(n_gpus)");
    const int group = h_local[0]->group;
    const size_t pt = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    std::vector<void*> gather((size_t)n_gpus, nullptr);
    for (int i = 0; i < n_gpus; i++) {
        b200zk_ctx* ctx = ctxs[i];
        if (n_local[i] > h_local[i]->n) return fail(c0, B200ZK_ERR_BAD_LEN, "more scalars than bases");
        B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
        const int rc = group == 1 ? msm_sharded_local<Fq>(ctx, h_local[i], scalars_local[i], scalars_on_device, n_local[i], &gather[(size_t)i])
                                  : msm_sharded_local<Fq2>(ctx, h_local[i], scalars_local[i], scalars_on_device, n_local[i], &gather[(size_t)i]);
        if (rc != B200ZK_OK) return fail(c0, rc, ctx->last_error);
    }
    if (n_gpus > 1) {
        B200ZK_NCCL(c0, nccl().GroupStart());
        for (int i = 0; i < n_gpus; i++) {
            const int rc = msm_sharded_collective(ctxs[i], gather[(size_t)i], pt);
            if (rc != B200ZK_OK) {
                nccl().GroupEnd();
                return fail(c0, rc, ctxs[i]->last_error);
            }
        }
        B200ZK_NCCL(c0, nccl().GroupEnd());
    }
    // every rank could sum; rank 0's result is the one returned
    B200ZK_CUDA(c0, cudaSetDevice(c0->device));
    void* dout;
    B200ZK_TRY(scratch(c0, "comm_msm_out", pt, &dout));
    B200ZK_TRY(b200zk_points_sum_device(c0, group, gather[0], (size_t)n_gpus, dout));
    B200ZK_CUDA(c0, cudaMemcpyAsync(out_affine, dout, pt, cudaMemcpyDeviceToHost, c0->stream));
    for (int i = 0; i < n_gpus; i++) {
        B200ZK_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        B200ZK_CUDA(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
    }
    if (out_is_inf) {
        bool z = true;
        for (size_t i = 0; i < pt; i++) z = z && out_affine[i] == 0;
        *out_is_inf = z ? 1 : 0;
    }
    return B200ZK_OK;
}

int b200zk_ntt_sharded_device(b200zk_ctx* ctx, void* d_local, uint32_t log_n, uint32_t log_n1, int inverse) {
    if (!ctx || !d_local) return B200ZK_ERR_BAD_ARG;
    if (log_n > 32) return fail(ctx, B200ZK_ERR_DOMAIN_TOO_LARGE, "log_n > 32");
    if (log_n1 > log_n) return fail(ctx, B200ZK_ERR_BAD_ARG, "log_n1 > log_n");
    B200ZK_TRY(need_comm(ctx));
    const uint32_t world = (uint32_t)ctx->comm_world, g = (uint32_t)ctx->comm_rank;
    const uint32_t log_n2 = log_n - log_n1;
    const uint64_t n1 = 1ull << log_n1, n2 = 1ull << log_n2;
    if ((world & (world - 1)) || n1 < world || n2 < world)
        return fail(ctx, B200ZK_ERR_BAD_ARG, "world size must be a power of two no larger than either factor of n");
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t C = n2 / world, R = n1 / world;
    const size_t bytes = (size_t)(C * n1) * sizeof(Fr);
    void *send, *recv = nullptr;
    B200ZK_TRY(scratch(ctx, "comm_ntt_send", bytes, &send));
    if (world > 1) B200ZK_TRY(scratch(ctx, "comm_ntt_recv", bytes, &recv));
    Fr* local = (Fr*)d_local;
    B200ZK_TRY(ntt_device(ctx, local, log_n1, inverse != 0, nullptr, (size_t)C));                       // columns
    B200ZK_TRY(b200zk_ntt_twiddle_transpose_device(ctx, local, send, log_n, C, n1, (uint64_t)g * C, inverse));  // -> [n1][C]
    if (world == 1) {
        B200ZK_CUDA(ctx, cudaMemcpyAsync(local, send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));  // [n1][n2] already
    } else {
        const size_t chunk = (size_t)(R * C) * sizeof(Fr);
        {
            ProfScope ps(ctx, "comm_all_to_all");
            B200ZK_NCCL(ctx, nccl().GroupStart());
            for (uint32_t p = 0; p < world; p++) {
                int r1 = nccl().Send((const uint8_t*)send + (size_t)p * chunk, chunk, NCCL_UINT8, (int)p, (nccl_comm_t)ctx->nccl_comm, ctx->stream);
                int r2 = nccl().Recv((uint8_t*)recv + (size_t)p * chunk, chunk, NCCL_UINT8, (int)p, (nccl_comm_t)ctx->nccl_comm, ctx->stream);
                if (r1 || r2) {
                    nccl().GroupEnd();
                    return fail(ctx, B200ZK_ERR_NCCL, std::string("ncclSend/ncclRecv: ") + nccl().GetErrorString(r1 ? r1 : r2));
                }
            }
            B200ZK_NCCL(ctx, nccl().GroupEnd());
        }
        {
            ProfScope ps(ctx, "ntt_interleave");
            ntt_interleave_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const uint4*)recv, (uint4*)local, R, C, world);
            B200ZK_TRY(check_launch(ctx, "ntt_interleave_kernel"));
        }
    }
    return ntt_device(ctx, local, log_n2, inverse != 0, nullptr, (size_t)R);                             // rows
}

}  // extern "C"
