// Host-side plumbing shared by every translation unit of libb200zk: the context behind the
// C ABI (include/b200zk.h), error propagation without exceptions across the boundary, and a
// small cache of device scratch buffers so steady-state calls never hit cudaMalloc.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/b200zk.h"

namespace b200zk {

struct KernelTimer {  // CUDA-event timing of one named kernel family, summed over launches
    double ms = 0.0;
    long launches = 0;
};

struct DeviceBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace b200zk

struct b200zk_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;                         // main stream (slot 0) of the ACTIVE lane, = main_lane[lane]
    // Two proof batches can be in flight (b200zk_*_prove_submit / b200zk_prove_wait): each has its own main stream
    // (witness -> H(x) -> assembly) and its own copy of every slot-0 scratch buffer ("lane"), while the MSM streams
    // below are shared, so the five MSMs of batch k+1 queue behind those of batch k and the head of one batch and
    // the tail of the other hide under bucket accumulation.
    static constexpr int LANES = 2;
    cudaStream_t main_lane[LANES] = {nullptr, nullptr};
    int lane = 0;
    cudaEvent_t lane_done[LANES] = {nullptr, nullptr};     // everything of the batch submitted on that lane has finished
    static constexpr int AUX_STREAMS = 5;                  // slots 1..5: the independent MSMs of a proof batch (a, b_g1, l, b_g2, h)
    cudaStream_t aux[AUX_STREAMS] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[AUX_STREAMS] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t fin = nullptr, fin2 = nullptr;            // high priority: latency-bound proof assembly pieces
    cudaEvent_t ev_fin2 = nullptr, ev_fork2 = nullptr;     // fin2 joined into main; fin_scalars done
    cudaEvent_t ev_h = nullptr;                            // H(x) coefficients ready: the h MSM may start
    cudaEvent_t ev_msm[3] = {nullptr, nullptr, nullptr};   // completion of the a / b_g1 / b_g2 MSMs
    bool concurrency = true;                               // b200zk_set_option("concurrency")
    int msm_parts = 0;                                     // b200zk_set_option("msm_parts"): 0 = automatic
    bool msm_glv = true;                                   // b200zk_set_option("msm_glv"): GLV for plain G1 bases
    int table_c_g1 = 13, table_c_g2 = 13;                  // window of a full digit table (precompute level 2): the table
                                                           // holds n * ceil(256 / c) * 2^(c-1) points
    // full digit tables: rounds of pairwise batched-affine additions (csrc/msm_affine.cuh) before the XYZZ running sums
    int msm_affine_levels = 4;                             // b200zk_set_option("msm_affine_levels"): 0 = off
    long long msm_affine_min_entries = 1ll << 22;          // ... only for batches with at least this many table entries
    long long msm_affine_min_entries_buckets = 1ll << 26;  // ... and, over plain bases (bucket method), for MSMs this large
                                                           // (measured: a gain from 2^22 points, a loss at 2^20)
    int msm_affine_b = 96;                                 // target additions per lane sharing one inversion per warp (the device plan rounds it to whole waves)
    std::string last_error;
    std::map<std::string, b200zk::DeviceBuf> scratch;      // named, grow-only
    std::map<std::string, b200zk::DeviceBuf> tables;       // twiddle / coset tables; the key spells out kind, size and
                                                           // the full 32-byte base (no hash: nothing to collide)
    // profiling (b200zk_prof_*): when enabled every launch of a tracked kernel is bracketed by
    // events on ctx->stream; resolved lazily at b200zk_prof_get.
    bool prof_enabled = false;
    std::map<std::string, b200zk::KernelTimer> prof;
    std::vector<std::tuple<std::string, cudaEvent_t, cudaEvent_t>> prof_pending;
    std::vector<cudaEvent_t> event_pool;
    std::vector<std::string> timeline;                     // "name start_ms end_ms" per bracket (b200zk_prof_timeline)
    float timeline_base = 0.f;
    long launches = 0;                                     // every kernel launch of this library
    std::map<std::string, double> stats;                   // work counters (b200zk_stat_get)
    void* poseidon_consts = nullptr;                       // device copy, see poseidon.cu
    struct PendingBatch {                                  // a submitted, not yet awaited proof batch (one per lane)
        bool active = false;
        uint64_t ticket = 0;
        size_t batch = 0;
        uint8_t* proofs_out = nullptr;                     // caller's buffers, filled at wait time
        uint8_t* status_out = nullptr;
        uint8_t* h_proofs = nullptr;                       // pinned staging (grow-only)
        uint32_t* h_status = nullptr;
        uint8_t* h_in = nullptr;                           // pinned staging of inputs | r | s
        size_t h_proofs_cap = 0, h_status_cap = 0, h_in_cap = 0;
    } pending[LANES];
    uint64_t next_ticket = 1;
    void* nccl_comm = nullptr;                             // ncclComm_t of this rank (csrc/comm.cu), null without one
    int comm_rank = 0, comm_world = 1;
};

namespace b200zk {

inline int fail(b200zk_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return code;
}

#define B200ZK_CUDA(ctx, expr)                                                                  \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::b200zk::fail(ctx, B200ZK_ERR_CUDA,                                         \
                                  std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                                      __FILE__ + ":" + std::to_string(__LINE__));               \
    } while (0)

#define B200ZK_TRY(expr)            \
    do {                            \
        int _rc = (expr);           \
        if (_rc != B200ZK_OK) return _rc; \
    } while (0)

inline cudaStream_t slot_stream(b200zk_ctx* ctx, int slot) {
    return (slot <= 0 || !ctx->concurrency) ? ctx->stream : ctx->aux[(slot - 1) % b200zk_ctx::AUX_STREAMS];
}

// grow-only named scratch allocation (slot > 0: a private copy for work running on an aux stream)
inline int scratch(b200zk_ctx* ctx, const char* name, size_t bytes, void** out, int slot = 0) {
    // slot > 0: the buffer belongs to a shared MSM stream (work on it is serialised by the stream);
    // slot 0: it belongs to the active lane's main stream, so lane 1 gets its own copy
    DeviceBuf& b = ctx->scratch[slot > 0 ? std::string(name) + "#" + std::to_string(slot)
                                : ctx->lane ? std::string(name) + "@" + std::to_string(ctx->lane) : std::string(name)];
    if (b.bytes < bytes) {
        if (b.ptr) B200ZK_CUDA(ctx, cudaFree(b.ptr));
        b.ptr = nullptr;
        b.bytes = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&b.ptr, want);
        if (e != cudaSuccess) {
            want = bytes;
            B200ZK_CUDA(ctx, cudaMalloc(&b.ptr, want));
        }
        b.bytes = want;
    }
    *out = b.ptr;
    return B200ZK_OK;
}

inline int check_launch(b200zk_ctx* ctx, const char* what) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(ctx, B200ZK_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    return B200ZK_OK;
}

// RAII bracket used as:  { ProfScope p(ctx, "msm_accumulate"); kernel<<<...>>>(...); }
struct ProfScope {
    b200zk_ctx* ctx;
    const char* name;
    cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get_event(b200zk_ctx* ctx) {
        if (!ctx->event_pool.empty()) {
            cudaEvent_t e = ctx->event_pool.back();
            ctx->event_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    cudaStream_t st;
    ProfScope(b200zk_ctx* c, const char* n, cudaStream_t s = nullptr) : ctx(c), name(n), st(s ? s : c->stream) {
        if (ctx->prof_enabled) {
            a = get_event(ctx);
            b = get_event(ctx);
            cudaEventRecord(a, st);
        }
    }
    ~ProfScope() {
        if (ctx->prof_enabled) {
            cudaEventRecord(b, st);
            int slot = 0;  // which stream: 0 main, 1..4 aux, 9 fin (shown in the timeline, stripped for the sums)
            for (int i = 0; i < b200zk_ctx::AUX_STREAMS; i++)
                if (st == ctx->aux[i]) slot = i + 1;
            if (st == ctx->fin) slot = 9;
            if (st == ctx->fin2) slot = 8;
            ctx->prof_pending.emplace_back(std::string(name) + "@" + std::to_string(slot), a, b);
        }
    }
};

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

inline void set_lane(b200zk_ctx* ctx, int lane) {
    ctx->lane = lane;
    ctx->stream = ctx->main_lane[lane];
}

}  // namespace b200zk
