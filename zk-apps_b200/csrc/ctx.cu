// Context, device memory helpers and profiling hooks of the C ABI (include/b200zk.h).
#include <algorithm>
#include <cstring>

#include "common.cuh"

using namespace b200zk;

extern "C" {

int b200zk_init(int device, b200zk_ctx** out) {
    if (!out) return B200ZK_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return B200ZK_ERR_NO_DEVICE;  // no CPU fallback by design
    if (device < 0 || device >= count) return B200ZK_ERR_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return B200ZK_ERR_CUDA;
    b200zk_ctx* ctx = new b200zk_ctx();
    ctx->device = device;
    // experiment / test overrides of the affine-level defaults (b200zk_set_option has the same knobs)
    if (const char* e = getenv("B200ZK_AFFINE_LEVELS")) ctx->msm_affine_levels = std::max(0, std::min(8, atoi(e)));
    if (const char* e = getenv("B200ZK_AFFINE_MIN_ENTRIES")) {
        ctx->msm_affine_min_entries = std::max(0ll, atoll(e));
        ctx->msm_affine_min_entries_buckets = ctx->msm_affine_min_entries ? std::max(ctx->msm_affine_min_entries, 1ll << 26) : 0;
    }
    if (const char* e = getenv("B200ZK_AFFINE_B")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 1024) ctx->msm_affine_b = v;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    // Stream priorities order the work of a proof batch (numerically lower = more urgent): the assembly pieces
    // (fin, fin2) first, then the main stream (witness -> H(x) pipeline -> h MSM), then the MSMs a > b_g1 > b_g2 > l,
    // one level each.  Blocks of a more urgent stream are dispatched first, so the MSMs complete one after the
    // other instead of all at the end: the latency-bound phases of one (fold, bucket reduction, the scalar
    // multiplications its result feeds) hide under the bucket accumulation of the next.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    auto prio = [&](int level) { return std::min(prio_lo, prio_hi + level); };
    for (int l = 0; l < b200zk_ctx::LANES; l++)
        if (cudaStreamCreateWithPriority(&ctx->main_lane[l], cudaStreamNonBlocking, prio(1)) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->lane_done[l], cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return B200ZK_ERR_CUDA;
        }
    ctx->stream = ctx->main_lane[0];
    for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++) {
        static const int level[b200zk_ctx::AUX_STREAMS] = {2, 3, 5, 4, 1};  // slots 1..5 = a, b_g1, l, b_g2, h
        if (cudaStreamCreateWithPriority(&ctx->aux[i], cudaStreamNonBlocking, prio(level[i])) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return B200ZK_ERR_CUDA;
        }
    }
    if (cudaStreamCreateWithPriority(&ctx->fin, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->fin2, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fin2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_h, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return B200ZK_ERR_CUDA;
    }
    for (int i = 0; i < 3; i++)
        if (cudaEventCreateWithFlags(&ctx->ev_msm[i], cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return B200ZK_ERR_CUDA;
        }
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return B200ZK_ERR_CUDA;
    }
    *out = ctx;
    return B200ZK_OK;
}

int b200zk_set_option(b200zk_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return B200ZK_ERR_BAD_ARG;
    if (strcmp(name, "concurrency") == 0) {
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        ctx->concurrency = value != 0;
        return B200ZK_OK;
    }
    if (strcmp(name, "msm_parts") == 0) {  // 0 = automatic, 1..4 = window groups of a single plain-bases MSM
        if (value < 0 || value > 4) return fail(ctx, B200ZK_ERR_BAD_ARG, "msm_parts must be 0..4");
        ctx->msm_parts = value;
        return B200ZK_OK;
    }
    if (strcmp(name, "msm_glv") == 0) {  // 1 (default): GLV half-length scalars for MSMs over plain bases
        ctx->msm_glv = value != 0;
        return B200ZK_OK;
    }
    if (strcmp(name, "table_c_g1") == 0 || strcmp(name, "table_c_g2") == 0) {  // window of full digit tables built from now on
        if (value < 2 || value > 16) return fail(ctx, B200ZK_ERR_BAD_ARG, "table window must be 2..16");
        (name[9] == '1' ? ctx->table_c_g1 : ctx->table_c_g2) = value;
        return B200ZK_OK;
    }
    if (strcmp(name, "msm_affine_levels") == 0) {  // rounds of batched-affine pairwise additions over digit tables; 0 = off
        if (value < 0 || value > 8) return fail(ctx, B200ZK_ERR_BAD_ARG, "msm_affine_levels must be 0..8");
        ctx->msm_affine_levels = value;
        return B200ZK_OK;
    }
    if (strcmp(name, "msm_affine_min_entries") == 0) {  // smaller batches keep the XYZZ running sums only
        if (value < 0) return fail(ctx, B200ZK_ERR_BAD_ARG, "msm_affine_min_entries must be >= 0");
        ctx->msm_affine_min_entries = value;
        ctx->msm_affine_min_entries_buckets = value ? std::max<long long>(value, 1ll << 26) : 0;   // 0 = always (tests)
        return B200ZK_OK;
    }
    if (strcmp(name, "msm_affine_b") == 0) {  // additions per lane that share one inversion per warp
        if (value < 1 || value > 1024) return fail(ctx, B200ZK_ERR_BAD_ARG, "msm_affine_b must be 1..1024");
        ctx->msm_affine_b = value;
        return B200ZK_OK;
    }
    return fail(ctx, B200ZK_ERR_BAD_ARG, std::string("unknown option ") + name);
}

void b200zk_destroy(b200zk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    b200zk_comm_destroy(ctx);
    for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++) {
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->fin) cudaStreamDestroy(ctx->fin);
    if (ctx->fin2) cudaStreamDestroy(ctx->fin2);
    if (ctx->ev_fin2) cudaEventDestroy(ctx->ev_fin2);
    if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
    if (ctx->ev_h) cudaEventDestroy(ctx->ev_h);
    for (int i = 0; i < 3; i++)
        if (ctx->ev_msm[i]) cudaEventDestroy(ctx->ev_msm[i]);
    for (auto& kv : ctx->scratch)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    for (auto& kv : ctx->tables)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    if (ctx->poseidon_consts) cudaFree(ctx->poseidon_consts);
    for (auto& t : ctx->prof_pending) {
        cudaEventDestroy(std::get<1>(t));
        cudaEventDestroy(std::get<2>(t));
    }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    for (int l = 0; l < b200zk_ctx::LANES; l++) {
        if (ctx->main_lane[l]) cudaStreamDestroy(ctx->main_lane[l]);
        if (ctx->lane_done[l]) cudaEventDestroy(ctx->lane_done[l]);
        if (ctx->pending[l].h_proofs) cudaFreeHost(ctx->pending[l].h_proofs);
        if (ctx->pending[l].h_status) cudaFreeHost(ctx->pending[l].h_status);
        if (ctx->pending[l].h_in) cudaFreeHost(ctx->pending[l].h_in);
    }
    delete ctx;
}

const char* b200zk_last_error(b200zk_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null ctx"; }

int b200zk_sync(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int l = 0; l < b200zk_ctx::LANES; l++) B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->main_lane[l]));
    return B200ZK_OK;
}

int b200zk_dev_alloc(b200zk_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    B200ZK_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return B200ZK_OK;
}

int b200zk_dev_free(b200zk_ctx* ctx, void* dptr) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    // work on any stream of the ctx may still read the buffer (auxiliary MSM streams, assembly streams)
    for (int l = 0; l < b200zk_ctx::LANES; l++) B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->main_lane[l]));
    for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++) B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->aux[i]));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->fin));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->fin2));
    B200ZK_CUDA(ctx, cudaFree(dptr));
    return B200ZK_OK;
}

int b200zk_dev_upload(b200zk_ctx* ctx, void* dptr, const void* host, size_t bytes) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_dev_download(b200zk_ctx* ctx, void* host, const void* dptr, size_t bytes) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    B200ZK_CUDA(ctx, cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

void* b200zk_stream(b200zk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

static void prof_resolve(b200zk_ctx* ctx) {
    for (int l = 0; l < b200zk_ctx::LANES; l++) cudaStreamSynchronize(ctx->main_lane[l]);
    for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++) cudaStreamSynchronize(ctx->aux[i]);
    if (ctx->fin) cudaStreamSynchronize(ctx->fin);
    if (ctx->fin2) cudaStreamSynchronize(ctx->fin2);
    for (auto& t : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t)) == cudaSuccess) {
            const std::string& tagged = std::get<0>(t);
            KernelTimer& k = ctx->prof[tagged.substr(0, tagged.find('@'))];
            k.ms += ms;
            k.launches++;
            // timeline: offsets from the first bracket since the last reset (events on any stream share one clock)
            float t0 = 0.f;
            cudaEvent_t origin = std::get<1>(ctx->prof_pending.front());
            if (ctx->timeline.size() < 4096 && cudaEventElapsedTime(&t0, origin, std::get<1>(t)) == cudaSuccess) {
                char line[160];
                snprintf(line, sizeof line, "%s %.4f %.4f\n", std::get<0>(t).c_str(), ctx->timeline_base + t0,
                         ctx->timeline_base + t0 + ms);
                ctx->timeline.push_back(line);
            }
        }
        ctx->event_pool.push_back(std::get<1>(t));
        ctx->event_pool.push_back(std::get<2>(t));
    }
    ctx->prof_pending.clear();
}

int b200zk_prof_enable(b200zk_ctx* ctx, int on) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    prof_resolve(ctx);
    ctx->prof_enabled = on != 0;
    return B200ZK_OK;
}

int b200zk_prof_timeline(b200zk_ctx* ctx, char* buf, size_t buflen) {
    if (!ctx || !buf || !buflen) return B200ZK_ERR_BAD_ARG;
    prof_resolve(ctx);
    std::string s;
    for (auto& l : ctx->timeline) {
        if (s.size() + l.size() + 1 >= buflen) break;
        s += l;
    }
    memcpy(buf, s.c_str(), s.size() + 1);
    return B200ZK_OK;
}

int b200zk_prof_reset(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    prof_resolve(ctx);
    ctx->timeline.clear();
    ctx->prof.clear();
    return B200ZK_OK;
}

int b200zk_prof_get(b200zk_ctx* ctx, const char* name, double* ms, long* launches) {
    if (!ctx || !name) return B200ZK_ERR_BAD_ARG;
    prof_resolve(ctx);
    auto it = ctx->prof.find(name);
    if (ms) *ms = it == ctx->prof.end() ? 0.0 : it->second.ms;
    if (launches) *launches = it == ctx->prof.end() ? 0 : it->second.launches;
    return B200ZK_OK;
}

int b200zk_prof_names(b200zk_ctx* ctx, char* buf, size_t buflen) {
    if (!ctx || !buf || !buflen) return B200ZK_ERR_BAD_ARG;
    prof_resolve(ctx);
    std::string s;
    for (auto& kv : ctx->prof) {
        if (!s.empty()) s += ",";
        s += kv.first;
    }
    strncpy(buf, s.c_str(), buflen - 1);
    buf[buflen - 1] = 0;
    return B200ZK_OK;
}

long b200zk_launch_count(b200zk_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200zk_stat_get(b200zk_ctx* ctx, const char* name, double* value) {
    if (!ctx || !name || !value) return B200ZK_ERR_BAD_ARG;
    auto it = ctx->stats.find(name);
    *value = it == ctx->stats.end() ? 0.0 : it->second;
    return B200ZK_OK;
}

int b200zk_stat_reset(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    ctx->stats.clear();
    return B200ZK_OK;
}

}  // extern "C"
