// engine_ctx.h — the context object behind `b200ols_ctx*` and the small host helpers (error string, device bump
// arena, double-buffered pinned metadata staging) shared by the translation units of libb200ols.so.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200ols.h"

namespace b200 {
struct Stager;     // staging.h: threaded pageable <-> device copies through a pinned ring
struct GroupPlan;  // plan_host.cu: device-side `.over()` key planning state
}

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...);

#define CU(expr)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(B200OLS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                        \
    } while (0)

#define TRY(expr)              \
    do {                       \
        int rc__ = (expr);     \
        if (rc__ != 0) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct b200ols_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    int smem_optin = 0;
    int64_t launches = 0;
    int tile_rows = 0, warps_per_cta = 0, ctas_per_sm = 0;
    bool multi_enabled = true;      // test hook B200OLS_MULTI=0: never use gram_multi_kernel
    int pred_lag = 4;               // test hook B200OLS_PRED_LAG: groups the stream may run ahead of the predictions (per SM)
    bool pred_enabled = true;       // test hook B200OLS_PRED=0: never use the fused Gram -> solve -> predict kernel
    long long fuse_min_bytes = -1;  // < 0: default; test hook B200OLS_FUSE_MIN_BYTES (0 = always fuse the solve)
    int variant = 3, unroll = 0;  // Gram kernel variant (b200ols_set_variant); 3 = CTA-cooperative TMA pipeline
    // bump arena in device memory, reset at the start of every call
    char *arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    bool arena_overflow = false;
    std::vector<void *> retired;  // old arenas, freed after the next synchronise
    // pinned host staging for small metadata (offsets, segment tables)
    // two halves used by alternating calls; a half is recycled only after the event recorded at the end of the
    // call that used it has completed (device-memspace calls return before their copies ran)
    char *pinned = nullptr;
    size_t pinned_cap = 0, pinned_off = 0, pinned_base = 0;
    int pinned_half = 0;
    cudaEvent_t pinned_ev[2] = {nullptr, nullptr};
    bool pinned_ev_set[2] = {false, false};
    // diagnostics of the last static call
    int32_t *last_flags = nullptr;  // device pointer inside the arena
    int64_t last_flags_n = 0;
    // last uploaded group-offset table (steady-state loops re-use it instead of re-copying every call)
    std::vector<int64_t> plan_offsets;
    int64_t plan_max_rows = 0, plan_wide_group = -1;
    int plan_F = -1;
    int64_t *plan_dev = nullptr;
    size_t plan_cap = 0;
    // tile table of gram_multi_kernel (runs of whole groups) for the cached grouping
    int64_t *tile_dev = nullptr;
    size_t tile_cap = 0;
    int64_t tile_count = 0, tile_rows_built = 0;
    bool tile_valid = false;
    // fused multi-GPU gather (b200ols_set_peer_gather)
    int n_peers = 0;
    double *peer_coef[8] = {};
    int64_t peer_group_base = 0, peer_total_groups = 0;
    // per-step completion of the fused gather (b200ols_set_peer_flags / b200ols_peer_step_complete)
    int n_flag_peers = 0, flag_rank = 0;
    unsigned long long *peer_flags[8] = {};
    int *flag_timeout = nullptr;  // device: set when a spin gave up
    unsigned int *done_counter = nullptr;              // device: arrival count of the in-kernel completion
    unsigned long long armed_signal = 0, armed_wait = 0;  // b200ols_peer_arm_step: consumed by the next fused-gather launch
    // coordinate-descent kernel choice (run_static_impl)
    int cd_thread = 1;                 // B200OLS_CD_THREAD: 0 = sub-warp kernel only, 1 = thread-per-group kernel from 1024 groups, 2 = always (k <= 16)
    int cd_thread_blocks = 0;          // B200OLS_CD_THREAD_BLOCKS: resident warps per SM of cd_thread_kernel (0 = what fits)
    // optional device-side timing of the dominant kernel
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof;
    std::vector<cudaEvent_t> event_pool;
    // pageable <-> device staging ring (staging.cu) and the device-side `.over()` key planner (plan_host.cu), lazily created
    b200::Stager *stager = nullptr;
    b200::GroupPlan *gplan = nullptr;
};

// staging.cu
struct StageSeg {
    const void *host;
    void *dev;
    size_t bytes;
};
int stage_h2d(b200ols_ctx *c, const StageSeg *segs, int n);  // enqueue; the compute stream waits for the copies
int stage_d2h(b200ols_ctx *c, const StageSeg *segs, int n);  // returns with the data in host memory (stream synchronised)
bool host_ptr_is_pinned(const void *p);
void stager_destroy(b200::Stager *s);
// plan_host.cu
void group_plan_destroy(b200::GroupPlan *g);

struct ProfScope {  // brackets a launch with events when profiling is on
    b200ols_ctx *c;
    cudaEvent_t a = nullptr, b = nullptr;
    explicit ProfScope(b200ols_ctx *ctx) : c(ctx) {
        if (!c->profiling) return;
        auto get = [&](cudaEvent_t *ev) {
            if (!c->event_pool.empty()) { *ev = c->event_pool.back(); c->event_pool.pop_back(); return true; }
            return cudaEventCreate(ev) == cudaSuccess;
        };
        if (get(&a) && get(&b)) cudaEventRecord(a, c->stream);
    }
    ~ProfScope() {
        if (a && b) {
            cudaEventRecord(b, c->stream);
            c->prof.emplace_back(a, b);
        }
    }
};

inline int arena_reserve(b200ols_ctx *c, size_t bytes) {
    if (bytes <= c->arena_cap) return 0;
    size_t cap = std::max(bytes + (bytes >> 2), static_cast<size_t>(64) << 20);
    void *p = nullptr;
    CU(cudaMalloc(&p, cap));
    if (c->arena) c->retired.push_back(c->arena);
    c->arena = static_cast<char *>(p);
    c->arena_cap = cap;
    return 0;
}

template <typename U>
inline U *arena_alloc(b200ols_ctx *c, size_t count) {
    const size_t bytes = (count * sizeof(U) + 255) & ~static_cast<size_t>(255);
    U *p = reinterpret_cast<U *>(c->arena + c->arena_off);
    c->arena_off += bytes;
    if (c->arena_off > c->arena_cap) c->arena_overflow = true;  // checked by ARENA_GUARD before any launch
    return p;
}

#define ARENA_GUARD(c)                                                                                  \
    do {                                                                                                \
        if ((c)->arena_overflow) {                                                                      \
            (c)->arena_overflow = false;                                                                \
            return fail(B200OLS_ERR_CUDA, "internal: device arena under-sized (%zu > %zu)", (c)->arena_off, \
                        (c)->arena_cap);                                                                \
        }                                                                                               \
    } while (0)

// `bytes` = end offset needed (c->pinned_off based).  Offsets handed out are absolute inside the buffer.
inline int pinned_reserve(b200ols_ctx *c, size_t bytes) {
    const size_t need = bytes - c->pinned_base;  // bytes needed inside the current half
    if (need <= c->pinned_cap / 2) return 0;
    CU(cudaStreamSynchronize(c->stream));
    char *old = c->pinned;
    const size_t used = c->pinned_off - c->pinned_base;
    size_t half = std::max(need + (need >> 1), static_cast<size_t>(2) << 20);
    half = (half + 4095) & ~static_cast<size_t>(4095);
    char *fresh = nullptr;
    CU(cudaMallocHost(reinterpret_cast<void **>(&fresh), 2 * half));
    const size_t new_base = c->pinned_half ? half : 0;
    if (old && used) std::memcpy(fresh + new_base, old + c->pinned_base, used);
    if (old) cudaFreeHost(old);
    c->pinned = fresh;
    c->pinned_cap = 2 * half;
    c->pinned_base = new_base;
    c->pinned_off = new_base + used;
    c->pinned_ev_set[0] = c->pinned_ev_set[1] = false;  // everything was synchronised above
    return 0;
}

// start of a call: switch to the other half of the pinned staging buffer
inline int pinned_begin(b200ols_ctx *c) {
    c->pinned_half ^= 1;
    const int h = c->pinned_half;
    if (c->pinned_ev_set[h]) {
        CU(cudaEventSynchronize(c->pinned_ev[h]));
        c->pinned_ev_set[h] = false;
    }
    c->pinned_base = h ? c->pinned_cap / 2 : 0;
    c->pinned_off = c->pinned_base;
    return 0;
}

// end of a call: everything staged in this half has been enqueued
inline int pinned_end(b200ols_ctx *c) {
    const int h = c->pinned_half;
    if (!c->pinned_ev[h]) CU(cudaEventCreateWithFlags(&c->pinned_ev[h], cudaEventDisableTiming));
    CU(cudaEventRecord(c->pinned_ev[h], c->stream));
    c->pinned_ev_set[h] = true;
    return 0;
}

inline int free_retired(b200ols_ctx *c) {
    if (c->retired.empty()) return 0;
    CU(cudaStreamSynchronize(c->stream));
    for (void *p : c->retired) cudaFree(p);
    c->retired.clear();
    return 0;
}
