// Context, device memory helpers and profiling hooks of the C ABI (include/b200zk.h).
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
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return B200ZK_ERR_CUDA;
    }
    for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++) {
        if (cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return B200ZK_ERR_CUDA;
        }
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
        cudaDeviceSynchronize();
        ctx->concurrency = value != 0;
        return B200ZK_OK;
    }
    return fail(ctx, B200ZK_ERR_BAD_ARG, std::string("unknown option ") + name);
}

void b200zk_destroy(b200zk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++) {
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
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
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* b200zk_last_error(b200zk_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null ctx"; }

int b200zk_sync(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    B200ZK_CUDA(ctx, cudaFree(dptr));
    return B200ZK_OK;
}

int b200zk_dev_upload(b200zk_ctx* ctx, void* dptr, const void* host, size_t bytes) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_dev_download(b200zk_ctx* ctx, void* host, const void* dptr, size_t bytes) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    B200ZK_CUDA(ctx, cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

void* b200zk_stream(b200zk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

static void prof_resolve(b200zk_ctx* ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++) cudaStreamSynchronize(ctx->aux[i]);
    for (auto& t : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t)) == cudaSuccess) {
            KernelTimer& k = ctx->prof[std::get<0>(t)];
            k.ms += ms;
            k.launches++;
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

int b200zk_prof_reset(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_ERR_BAD_ARG;
    prof_resolve(ctx);
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
