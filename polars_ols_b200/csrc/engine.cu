// engine.cu — host side of libb200ols.so: the C ABI of include/b200ols.h over the sm_100a kernels.
//
// What the reference does per group and per call (polars `.over()` -> one `_polars_plugin_*` FFI call
// per group -> Series -> row-major ndarray -> solver -> Series; src/expressions.rs:390-446) is done
// here once per frame: stage columns (H2D / gather / null policy), ONE streaming Gram+solve launch for
// all groups, optionally one prediction pass, copy results back.  No CPU fallback exists: without a
// CUDA device every entry point fails with B200OLS_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200ols.h"
#include "gram_cta.cuh"
#include "gram_multi.cuh"
#include "gram_pred.cuh"
#include "gram_wide.cuh"
#include "gram_ldg.cuh"
#include "gram_stream.cuh"
#include "moving.cuh"
#include "predict.cuh"
#include "prep.cuh"
#include "qr_fallback.cuh"
#include "small_solve.cuh"
#include "svd_solve.cuh"
#include "big.cuh"
#include "big_stats.cuh"
#include "batch_solve.cuh"
#include "cd_solve.cuh"
#include "cd_thread.h"
#include "stats.cuh"
#include "multi_target.cuh"

using namespace b200;

static __global__ void nan_mask_kernel(const double *v, uint8_t *m, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) m[i] = (v[i] == v[i]) ? 1 : 0;
}


#include "engine_ctx.h"

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}



extern "C" int b200ols_version(void) { return B200OLS_VERSION; }
extern "C" const char *b200ols_last_error(void) { return g_last_error.c_str(); }

extern "C" int b200ols_create_on_stream(int device, void *cuda_stream, b200ols_ctx **out) {
    if (!out) return fail(B200OLS_ERR_INVALID, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(B200OLS_ERR_NO_DEVICE,
                    "no CUDA device available (%s): libb200ols has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(B200OLS_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    b200ols_ctx *c = new b200ols_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) {
        delete c;
        return fail(B200OLS_ERR_NO_DEVICE, "device sm_%d%d: this library is built for sm_100a only", prop.major,
                    prop.minor);
    }
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    if (cuda_stream) {
        c->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    if (const char *v = std::getenv("B200OLS_VARIANT")) c->variant = std::atoi(v);  // test hook: force a Gram kernel variant
    if (const char *v = std::getenv("B200OLS_MULTI")) c->multi_enabled = std::atoi(v) != 0;
    if (const char *v = std::getenv("B200OLS_PRED")) c->pred_enabled = std::atoi(v) != 0;
    if (const char *v = std::getenv("B200OLS_PRED_LAG")) c->pred_lag = std::max(1, std::atoi(v));
    if (const char *v = std::getenv("B200OLS_FUSE_MIN_BYTES")) c->fuse_min_bytes = std::atoll(v);  // test hook: fused-solve threshold
    if (c->variant != 1 && c->variant != 3) c->variant = 3;
    *out = c;
    return 0;
}

extern "C" int b200ols_create(int device, b200ols_ctx **out) { return b200ols_create_on_stream(device, nullptr, out); }

extern "C" void b200ols_destroy(b200ols_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (void *p : c->retired) cudaFree(p);
    if (c->arena) cudaFree(c->arena);
    if (c->plan_dev) cudaFree(c->plan_dev);
    if (c->tile_dev) cudaFree(c->tile_dev);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->flag_timeout) cudaFree(c->flag_timeout);
    if (c->done_counter) cudaFree(c->done_counter);
    stager_destroy(c->stager);
    group_plan_destroy(c->gplan);
    for (int h = 0; h < 2; ++h)
        if (c->pinned_ev[h]) cudaEventDestroy(c->pinned_ev[h]);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int b200ols_synchronize(b200ols_ctx *c) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" void *b200ols_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        fail(B200OLS_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
extern "C" void b200ols_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" int64_t b200ols_launch_count(const b200ols_ctx *c) { return c ? c->launches : 0; }

extern "C" void *b200ols_device_alloc(b200ols_ctx *c, size_t bytes) {
    if (!c) return nullptr;
    void *p = nullptr;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        fail(B200OLS_ERR_CUDA, "cudaMalloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
extern "C" void b200ols_device_free(b200ols_ctx *c, void *p) {
    if (c && p) {
        cudaSetDevice(c->device);
        cudaFree(p);
    }
}
extern "C" int b200ols_device_memset(b200ols_ctx *c, void *dev_ptr, int value, size_t bytes) {
    if (!c || !dev_ptr) return fail(B200OLS_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(dev_ptr, value, bytes, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int b200ols_ipc_export(b200ols_ctx *c, const void *dev_ptr, uint8_t handle[64]) {
    if (!c || !dev_ptr || !handle) return fail(B200OLS_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    std::memcpy(handle, &h, 64);
    return 0;
}
extern "C" int b200ols_ipc_open(b200ols_ctx *c, const uint8_t handle[64], void **dev_ptr) {
    if (!c || !dev_ptr || !handle) return fail(B200OLS_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int b200ols_ipc_close(b200ols_ctx *c, void *dev_ptr) {
    if (!c || !dev_ptr) return fail(B200OLS_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}
extern "C" int b200ols_copy_to_host(b200ols_ctx *c, void *dst_host, const void *src_dev, size_t bytes) {
    if (!c || !dst_host || !src_dev) return fail(B200OLS_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int b200ols_set_peer_gather(b200ols_ctx *c, int n_peers, void *const *peer_coef, int64_t group_base,
                                       int64_t total_groups) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (n_peers < 0 || n_peers > 8) return fail(B200OLS_ERR_INVALID, "n_peers must be in [0, 8]");
    if (n_peers > 0 && (!peer_coef || group_base < 0 || total_groups <= 0)) return fail(B200OLS_ERR_INVALID, "bad peer table");
    c->n_peers = n_peers;
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_coef[r]) return fail(B200OLS_ERR_INVALID, "peer_coef[%d] is NULL", r);
        c->peer_coef[r] = static_cast<double *>(peer_coef[r]);
    }
    c->peer_group_base = group_base;
    c->peer_total_groups = total_groups;
    return 0;
}

// ---- per-step completion of the fused gather -------------------------------------------------------------------------
// Every rank owns a flag array flags[8] (in its IPC-exported buffer).  After a step's kernels, one tiny kernel on the same
// stream (so the step's peer stores are ordered before it) RELEASES `step` into slot [my rank] of every peer's array and
// then ACQUIRE-spins until every slot of its own array has reached `step`: when it retires, all ranks' coefficient rows of
// this step have landed in this rank's buffer — the guarantee an all-gather gives, without a collective launch.
struct PeerFlagParams {
    unsigned long long *peer[8];
    int n_peers, rank;
    unsigned long long step;       // value to signal
    unsigned long long wait_step;  // value to wait for (may trail `step`: double-buffered gathers)
    int *timeout_flag;
};
// what: bit 0 = signal (release `step` into every peer's slot [rank]), bit 1 = wait (acquire-spin on the own slots)
static __global__ void peer_step_complete_kernel(const PeerFlagParams p, int what) {
    const int r = threadIdx.x;
    if (r >= p.n_peers) return;
    if (what & 1) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer[r] + p.rank), "l"(p.step) : "memory");
    }
    if (!(what & 2)) return;
    const unsigned long long *mine = p.peer[p.rank] + r;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= p.wait_step) break;
        if (clock64() - t0 > 4000000000LL) {  // ~2 s: a peer died; do not hang the device
            *p.timeout_flag = 1;
            break;
        }
    }
}

extern "C" int b200ols_set_peer_flags(b200ols_ctx *c, int n_peers, void *const *peer_flags, int rank) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && (!peer_flags || rank < 0 || rank >= n_peers)))
        return fail(B200OLS_ERR_INVALID, "bad peer flag table");
    CU(cudaSetDevice(c->device));
    c->n_flag_peers = n_peers;
    c->flag_rank = rank;
    for (int r = 0; r < n_peers; ++r) c->peer_flags[r] = static_cast<unsigned long long *>(peer_flags[r]);
    if (n_peers > 0 && !c->flag_timeout) {
        CU(cudaMalloc(reinterpret_cast<void **>(&c->flag_timeout), sizeof(int)));
        CU(cudaMemsetAsync(c->flag_timeout, 0, sizeof(int), c->stream));
        CU(cudaMalloc(reinterpret_cast<void **>(&c->done_counter), sizeof(unsigned int)));
        CU(cudaMemsetAsync(c->done_counter, 0, sizeof(unsigned int), c->stream));
    }
    c->armed_signal = 0;
    return 0;
}

// Arms the in-kernel completion for the NEXT b200ols_least_squares_coefficients call on this context: when that call takes
// the fused peer-gather kernel, its last solver warp signals `signal_step` and waits for `wait_step` (no extra launch);
// on any other route the engine enqueues the equivalent flag kernel behind the call's kernels.
extern "C" int b200ols_peer_arm_step(b200ols_ctx *c, uint64_t signal_step, uint64_t wait_step) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (c->n_flag_peers <= 0) return fail(B200OLS_ERR_INVALID, "b200ols_set_peer_flags was not called");
    if (signal_step == 0) return fail(B200OLS_ERR_INVALID, "signal_step must be >= 1");
    c->armed_signal = signal_step;
    c->armed_wait = wait_step;
    return 0;
}

static int peer_step(b200ols_ctx *c, uint64_t step, uint64_t wait_step, int what);
extern "C" int b200ols_peer_step_complete(b200ols_ctx *c, uint64_t step) { return peer_step(c, step, step, 3); }
extern "C" int b200ols_peer_step_signal(b200ols_ctx *c, uint64_t step) { return peer_step(c, step, step, 1); }
extern "C" int b200ols_peer_step_wait(b200ols_ctx *c, uint64_t step) { return peer_step(c, step, step, 2); }
extern "C" int b200ols_peer_step_signal_wait(b200ols_ctx *c, uint64_t signal_step, uint64_t wait_step) {
    return peer_step(c, signal_step, wait_step, wait_step > 0 ? 3 : 1);
}
static int peer_step(b200ols_ctx *c, uint64_t step, uint64_t wait_step, int what) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (c->n_flag_peers <= 0) return fail(B200OLS_ERR_INVALID, "b200ols_set_peer_flags was not called");
    PeerFlagParams p;
    std::memset(&p, 0, sizeof(p));
    for (int r = 0; r < c->n_flag_peers; ++r) p.peer[r] = c->peer_flags[r];
    p.n_peers = c->n_flag_peers;
    p.rank = c->flag_rank;
    p.step = step;
    p.wait_step = wait_step;
    p.timeout_flag = c->flag_timeout;
    peer_step_complete_kernel<<<1, 32, 0, c->stream>>>(p, what);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int b200ols_peer_timed_out(b200ols_ctx *c) {
    if (!c || !c->flag_timeout) return 0;
    int v = 0;
    if (cudaMemcpyAsync(&v, c->flag_timeout, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    return v;
}

// generic fallback of the fused gather: after kernels that rewrite beta (QR / SVD / CD / split groups)
struct PeerScatterParams {
    const double *beta;
    double *peer[8];
    int n_peers;
    int64_t n, base_elems;
};
static __global__ void peer_scatter_kernel(const PeerScatterParams p) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const double v = p.beta[i];
    for (int r = 0; r < p.n_peers; ++r) p.peer[r][p.base_elems + i] = v;
}

extern "C" int b200ols_set_profiling(b200ols_ctx *c, int enabled) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    c->profiling = enabled != 0;
    return 0;
}

extern "C" int b200ols_profile_drain(b200ols_ctx *c, float *ms, int max) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    CU(cudaStreamSynchronize(c->stream));
    int n = 0;
    for (auto &pr : c->prof) {
        float t = 0.f;
        cudaEventElapsedTime(&t, pr.first, pr.second);
        if (ms && n < max) ms[n++] = t;
        c->event_pool.push_back(pr.first);
        c->event_pool.push_back(pr.second);
    }
    c->prof.clear();
    return n;
}

extern "C" int b200ols_set_tuning(b200ols_ctx *c, int tile_rows, int warps_per_cta, int ctas_per_sm) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (tile_rows < 0 || (tile_rows % 8) != 0) return fail(B200OLS_ERR_INVALID, "tile_rows must be a multiple of 8");
    if (warps_per_cta < 0 || warps_per_cta > GRAM_MAX_WARPS) return fail(B200OLS_ERR_INVALID, "warps_per_cta out of range");
    c->tile_rows = tile_rows;
    c->warps_per_cta = warps_per_cta;
    c->ctas_per_sm = ctas_per_sm;
    return 0;
}

extern "C" int b200ols_set_variant(b200ols_ctx *c, int variant, int unroll) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (variant != 1 && variant != 3)
        return fail(B200OLS_ERR_INVALID, "variant must be 3 (CTA-cooperative TMA pipeline, default) or 1 (direct-load DMMA, the non-TMA fallback for k <= 16)");
    c->variant = variant;
    c->unroll = unroll;
    return 0;
}

extern "C" int b200ols_last_group_flags(b200ols_ctx *c, int32_t *flags, int64_t n_groups) {
    if (!c || !flags) return fail(B200OLS_ERR_INVALID, "NULL argument");
    if (!c->last_flags || n_groups != c->last_flags_n) return fail(B200OLS_ERR_INVALID, "no flags for %lld groups", (long long)n_groups);
    CU(cudaMemcpyAsync(flags, c->last_flags, sizeof(int32_t) * n_groups, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// staging: frame -> device-resident, packed, null-cleaned SoA columns
// ------------------------------------------------------------------------------------------------
struct Staged {
    int kd = 0, has_w = 0, w_is_sqrt = 0, F = 0, intercept = 0;
    size_t esz = 8;
    int64_t n = 0, n_pad = 0;
    const void *feat[GRAM_MAX_COLS] = {};  // fit + prediction features (null-cleaned)
    const void *y = nullptr;               // fit target (cleaned)
    const void *w = nullptr;
    const void *mask = nullptr;            // T-typed row mask or nullptr
    const void *y_raw = nullptr;           // target for residuals
    int y_raw_packed = 0;                  // y_raw indexed by packed position
    const uint8_t *y_validity = nullptr;   // device bitmap (original rows) or nullptr
    const int64_t *row_index = nullptr;    // device
    int64_t n_groups = 1;
    std::vector<int64_t> offsets;          // host copy [G+1]
    int64_t max_group_rows = 0;
    int64_t wide_group = -1;               // first group with 0 < n <= k (needs the min-norm SVD path), or -1
    bool prepped = false;
    const void *raw[GRAM_MAX_COLS] = {};       // device copies of the caller's columns, ORIGINAL row order (features, target, w)
    const uint8_t *raw_bm[GRAM_MAX_COLS] = {};  // their validity bitmaps (device) or nullptr
};

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int validate_frame(const b200ols_frame *f) {
    if (!f) return fail(B200OLS_ERR_INVALID, "frame is NULL");
    if (f->n_rows < 0) return fail(B200OLS_ERR_INVALID, "n_rows < 0");
    if (f->n_features < 0) return fail(B200OLS_ERR_INVALID, "n_features < 0");
    // src/expressions.rs:72 `assert!(m > 1, "must pass at least 2 series")`
    if (f->n_features + (f->add_intercept ? 1 : 0) < 1) return fail(B200OLS_ERR_INVALID, "must pass at least 2 series");
    if (f->n_features + (f->add_intercept ? 1 : 0) > BIG_MAX_F)
        return fail(B200OLS_ERR_UNSUPPORTED, "more than %d coefficients (%d) is not implemented on the device", BIG_MAX_F,
                    f->n_features + (f->add_intercept ? 1 : 0));
    if (f->dtype != B200OLS_F64 && f->dtype != B200OLS_F32) return fail(B200OLS_ERR_INVALID, "bad dtype %d", f->dtype);
    if (f->memspace != B200OLS_HOST && f->memspace != B200OLS_DEVICE) return fail(B200OLS_ERR_INVALID, "bad memspace");
    if (f->n_rows > 0 && !f->target.values) return fail(B200OLS_ERR_INVALID, "target.values is NULL");
    if (f->n_features > 0 && !f->features) return fail(B200OLS_ERR_INVALID, "features is NULL");
    for (int j = 0; j < f->n_features; ++j)
        if (f->n_rows > 0 && !f->features[j].values) return fail(B200OLS_ERR_INVALID, "features[%d].values is NULL", j);
    if (f->n_groups < 1) return fail(B200OLS_ERR_INVALID, "n_groups must be >= 1");
    if (f->group_offsets) {
        if (f->group_offsets[0] != 0 || f->group_offsets[f->n_groups] != f->n_rows)
            return fail(B200OLS_ERR_INVALID, "group_offsets must start at 0 and end at n_rows");
    } else if (f->n_groups != 1) {
        return fail(B200OLS_ERR_INVALID, "group_offsets is NULL but n_groups != 1");
    }
    return 0;
}

// mask_kind / fill derived from the null policy; `moving` = rls / rolling (always zero fill,
// src/expressions.rs:603,629,656,683)
static void policy_to_prep(int policy, bool moving, int *fill, int *mask_kind) {
    *fill = PREP_ZERO;
    *mask_kind = MASK_NONE;
    switch (policy) {
        case B200OLS_NULL_IGNORE: *fill = moving ? PREP_ZERO : PREP_NAN; break;
        case B200OLS_NULL_ZERO: break;
        case B200OLS_NULL_DROP:
        case B200OLS_NULL_DROP_ZERO:
        case B200OLS_NULL_DROP_WINDOW: *mask_kind = MASK_ALL; break;
        case B200OLS_NULL_DROP_Y_ZERO_X: *mask_kind = MASK_TARGET; break;
    }
}

template <typename T>
static void launch_prep(b200ols_ctx *c, const PrepParams &pp) {
    const int64_t blocks = std::min<int64_t>((pp.n_rows_pad + 255) / 256, static_cast<int64_t>(c->sm_count) * 16);
    prep_kernel<T><<<static_cast<unsigned>(std::max<int64_t>(blocks, 1)), 256, 0, c->stream>>>(pp);
    c->launches++;
}

// Brings every column of the frame onto the device (arena), applies gather + null policy when needed.
// Arena must already be reserved.  `moving`: rls / rolling semantics for the null policy.
static int stage_frame(b200ols_ctx *c, const b200ols_frame *f, int policy, bool moving, Staged *st, int extra_targets = 0) {
    const int kd = f->n_features;
    const size_t esz = f->dtype == B200OLS_F64 ? 8 : 4;
    const int64_t n = f->n_rows;
    const int64_t A = 16 / static_cast<int64_t>(esz);
    const int64_t n_pad = static_cast<int64_t>(round_up(static_cast<size_t>(n), static_cast<size_t>(A)));
    st->kd = kd;
    st->esz = esz;
    st->n = n;
    st->n_pad = n_pad;
    st->intercept = f->add_intercept ? 1 : 0;
    st->F = kd + st->intercept;
    st->has_w = f->sample_weights ? 1 : 0;
    st->n_groups = f->n_groups;
    if (f->group_offsets && c->plan_offsets.size() == static_cast<size_t>(f->n_groups) + 1 &&
        std::memcmp(c->plan_offsets.data(), f->group_offsets, sizeof(int64_t) * (f->n_groups + 1)) == 0) {
        // same grouping as the previous call (steady-state loops): reuse the validated host copy
        st->offsets = c->plan_offsets;
        st->max_group_rows = c->plan_max_rows;
        st->wide_group = c->plan_wide_group;
        if (c->plan_F != st->F) {  // the "n <= k" test depends on the number of coefficients
            st->wide_group = -1;
            for (int64_t g = 0; g < f->n_groups && st->wide_group < 0; ++g) {
                const int64_t len = st->offsets[g + 1] - st->offsets[g];
                if (len > 0 && len <= st->F) st->wide_group = g;
            }
        }
    } else if (f->group_offsets) {
        st->offsets.resize(static_cast<size_t>(f->n_groups) + 1);
        std::memcpy(st->offsets.data(), f->group_offsets, sizeof(int64_t) * (f->n_groups + 1));
        for (int64_t g = 0; g < f->n_groups; ++g) {
            const int64_t len = st->offsets[g + 1] - st->offsets[g];
            if (len < 0) return fail(B200OLS_ERR_INVALID, "group_offsets must be non-decreasing");
            st->max_group_rows = std::max(st->max_group_rows, len);
            if (len > 0 && len <= st->F && st->wide_group < 0) st->wide_group = g;
        }
    } else {
        st->offsets.resize(2);
        st->offsets[0] = 0;
        st->offsets[1] = n;
        st->max_group_rows = n;
        if (n > 0 && n <= st->F) st->wide_group = 0;
    }

    const int ncol = kd + 1 + st->has_w;  // features, target, weights
    const b200ols_column *cols[GRAM_MAX_COLS];
    for (int j = 0; j < kd; ++j) cols[j] = &f->features[j];
    cols[kd] = &f->target;
    if (st->has_w) cols[kd + 1] = f->sample_weights;
    bool any_validity = false;
    for (int cidx = 0; cidx < ncol; ++cidx) any_validity = any_validity || (cols[cidx]->validity != nullptr);
    const size_t bm_bytes = static_cast<size_t>((n + 7) / 8);

    // 1) raw columns on the device
    const void *dev_vals[GRAM_MAX_COLS];
    const uint8_t *dev_bm[GRAM_MAX_COLS];
    const int64_t *dev_rowidx = nullptr;
    std::vector<StageSeg> segs;  // host -> device copies of this frame (pageable sources are pipelined through the pinned ring)
    for (int cidx = 0; cidx < ncol; ++cidx) {
        dev_bm[cidx] = nullptr;
        if (f->memspace == B200OLS_DEVICE) {
            dev_vals[cidx] = cols[cidx]->values;
            dev_bm[cidx] = cols[cidx]->validity;
        } else {
            char *d = arena_alloc<char>(c, static_cast<size_t>(n_pad) * esz);
            if (n > 0) segs.push_back({cols[cidx]->values, d, static_cast<size_t>(n) * esz});
            if (n_pad > n) CU(cudaMemsetAsync(d + static_cast<size_t>(n) * esz, 0, static_cast<size_t>(n_pad - n) * esz, c->stream));
            dev_vals[cidx] = d;
            if (cols[cidx]->validity) {
                uint8_t *b = arena_alloc<uint8_t>(c, bm_bytes);
                segs.push_back({cols[cidx]->validity, b, bm_bytes});
                dev_bm[cidx] = b;
            }
        }
    }
    if (f->row_index) {
        if (f->memspace == B200OLS_DEVICE || f->row_index_on_device) {
            dev_rowidx = f->row_index;
        } else {
            int64_t *d = arena_alloc<int64_t>(c, static_cast<size_t>(n));
            segs.push_back({f->row_index, d, sizeof(int64_t) * static_cast<size_t>(n)});
            dev_rowidx = d;
        }
    }
    if (!segs.empty()) {
        if (c->arena_overflow) return fail(B200OLS_ERR_CUDA, "internal: device arena under-sized before staging");
        TRY(stage_h2d(c, segs.data(), static_cast<int>(segs.size())));
    }
    for (int cidx = 0; cidx < ncol; ++cidx) {
        st->raw[cidx] = dev_vals[cidx];
        st->raw_bm[cidx] = dev_bm[cidx];
    }
    st->row_index = dev_rowidx;
    st->y_validity = dev_bm[kd];
    st->y_raw = dev_vals[kd];
    st->y_raw_packed = 0;

    // 2) gather / null policy pass (only when needed)
    int fill, mask_kind;
    policy_to_prep(policy, moving, &fill, &mask_kind);
    bool misaligned = false;  // e.g. a torch slice x[1:]: the streaming kernels need 16-byte aligned columns -> copy pass
    if (f->memspace == B200OLS_DEVICE)
        for (int cidx = 0; cidx < ncol; ++cidx) misaligned = misaligned || (reinterpret_cast<uintptr_t>(dev_vals[cidx]) & 15u) != 0;
    const bool need_prep = any_validity || dev_rowidx != nullptr || misaligned;
    if (!need_prep) {
        for (int j = 0; j < kd; ++j) st->feat[j] = dev_vals[j];
        st->y = dev_vals[kd];
        st->w = st->has_w ? dev_vals[kd + 1] : nullptr;
        st->w_is_sqrt = 0;
        st->mask = nullptr;
        return 0;
    }
    PrepParams pp;
    std::memset(&pp, 0, sizeof(pp));
    pp.kd = kd;
    pp.has_w = st->has_w;
    pp.fill = fill;
    pp.mask_kind = any_validity ? mask_kind : MASK_NONE;
    pp.n_rows = n;
    pp.n_rows_pad = n_pad;
    pp.row_index = dev_rowidx;
    pp.n_targets = extra_targets + 1;
    for (int cidx = 0; cidx < ncol; ++cidx) {
        pp.in[cidx] = dev_vals[cidx];
        pp.validity[cidx] = dev_bm[cidx];
        pp.out[cidx] = arena_alloc<char>(c, static_cast<size_t>(n_pad) * esz);
    }
    if (pp.mask_kind != MASK_NONE) pp.mask_out = arena_alloc<char>(c, static_cast<size_t>(n_pad) * esz);
    if (f->dtype == B200OLS_F64) launch_prep<double>(c, pp); else launch_prep<float>(c, pp);
    CU(cudaGetLastError());
    for (int j = 0; j < kd; ++j) st->feat[j] = pp.out[j];
    st->y = pp.out[kd];
    st->w = st->has_w ? pp.out[kd + 1] : nullptr;
    st->w_is_sqrt = 1;
    st->mask = pp.mask_out;
    st->prepped = true;
    return 0;
}

// upper bound of the arena bytes stage_frame() may take
static size_t stage_bytes_bound(const b200ols_frame *f) {
    const size_t esz = f->dtype == B200OLS_F64 ? 8 : 4;
    const size_t n_pad = round_up(static_cast<size_t>(f->n_rows), 16 / esz) + 64;
    const size_t ncol = static_cast<size_t>(f->n_features) + 2;
    size_t b = 0;
    const size_t per_col = round_up(n_pad * esz, 256) + 256;
    if (f->memspace == B200OLS_HOST) b += ncol * (per_col + round_up(static_cast<size_t>(f->n_rows) / 8 + 1, 256) + 256);
    if (f->row_index && f->memspace == B200OLS_HOST) b += round_up(static_cast<size_t>(f->n_rows) * 8, 256) + 256;
    b += (ncol + 1) * per_col;  // prep outputs + mask
    return b;
}

// ------------------------------------------------------------------------------------------------
// segment plan: groups longer than SEG_MAX rows are split so that one warp never owns more than that
// ------------------------------------------------------------------------------------------------
struct Plan {
    int64_t nseg = 0;
    const int64_t *seg_off = nullptr;        // device [nseg+1]
    const int32_t *seg_group = nullptr;      // device [nseg] or nullptr when segment == group
    const int64_t *group_seg_off = nullptr;  // device [G+1] or nullptr
    bool split = false;
};

static int build_plan(b200ols_ctx *c, const Staged &st, int64_t seg_max, Plan *pl) {
    const int64_t G = st.n_groups;
    std::vector<int64_t> seg_off;
    std::vector<int32_t> seg_group;
    std::vector<int64_t> gso;
    pl->split = st.max_group_rows > seg_max;
    if (!pl->split) {
        pl->nseg = G;
        const size_t bytes = sizeof(int64_t) * (G + 1);
        if (c->plan_dev && c->plan_offsets.size() == st.offsets.size() &&
            std::memcmp(c->plan_offsets.data(), st.offsets.data(), bytes) == 0) {
            pl->seg_off = c->plan_dev;  // identical grouping as the previous call: table is already on the device
            return 0;
        }
        if (bytes > c->plan_cap) {
            if (c->plan_dev) c->retired.push_back(c->plan_dev);
            c->plan_dev = nullptr;
            c->plan_cap = 0;
            void *q = nullptr;
            CU(cudaMalloc(&q, bytes + 4096));
            c->plan_dev = static_cast<int64_t *>(q);
            c->plan_cap = bytes + 4096;
        }
        c->plan_offsets.clear();  // invalid until the copy below is enqueued
        c->tile_valid = false;
        TRY(pinned_reserve(c, c->pinned_off + bytes + 256));
        char *h = c->pinned + c->pinned_off;
        std::memcpy(h, st.offsets.data(), bytes);
        c->pinned_off += round_up(bytes, 256);
        CU(cudaMemcpyAsync(c->plan_dev, h, bytes, cudaMemcpyHostToDevice, c->stream));
        c->plan_offsets = st.offsets;
        c->plan_max_rows = st.max_group_rows;
        c->plan_wide_group = st.wide_group;
        c->plan_F = st.F;
        pl->seg_off = c->plan_dev;
        return 0;
    }
    gso.resize(G + 1);
    seg_off.push_back(0);
    for (int64_t g = 0; g < G; ++g) {
        gso[g] = static_cast<int64_t>(seg_group.size());
        const int64_t a = st.offsets[g], b = st.offsets[g + 1];
        if (b == a) {  // empty group still owns one (empty) segment so that it gets its zero coefficients
            seg_group.push_back(static_cast<int32_t>(g));
            seg_off.push_back(b);
            continue;
        }
        const int64_t parts = (b - a + seg_max - 1) / seg_max;
        const int64_t step = ((b - a + parts - 1) / parts + 15) & ~static_cast<int64_t>(15);
        for (int64_t s = a; s < b; s += step) {
            seg_group.push_back(static_cast<int32_t>(g));
            seg_off.push_back(std::min(s + step, b));
        }
    }
    gso[G] = static_cast<int64_t>(seg_group.size());
    pl->nseg = static_cast<int64_t>(seg_group.size());
    const size_t b0 = sizeof(int64_t) * seg_off.size(), b1 = sizeof(int32_t) * seg_group.size(), b2 = sizeof(int64_t) * gso.size();
    TRY(pinned_reserve(c, c->pinned_off + b0 + b1 + b2 + 1024));
    char *h0 = c->pinned + c->pinned_off; c->pinned_off += round_up(b0, 256);
    char *h1 = c->pinned + c->pinned_off; c->pinned_off += round_up(b1, 256);
    char *h2 = c->pinned + c->pinned_off; c->pinned_off += round_up(b2, 256);
    std::memcpy(h0, seg_off.data(), b0);
    std::memcpy(h1, seg_group.data(), b1);
    std::memcpy(h2, gso.data(), b2);
    int64_t *d0 = arena_alloc<int64_t>(c, seg_off.size());
    int32_t *d1 = arena_alloc<int32_t>(c, seg_group.size());
    int64_t *d2 = arena_alloc<int64_t>(c, gso.size());
    CU(cudaMemcpyAsync(d0, h0, b0, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d1, h1, b1, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d2, h2, b2, cudaMemcpyHostToDevice, c->stream));
    pl->seg_off = d0;
    pl->seg_group = d1;
    pl->group_seg_off = d2;
    return 0;
}

static size_t plan_bytes_bound(const b200ols_frame *f, int64_t seg_max) {
    const size_t nseg = static_cast<size_t>(f->n_groups) + static_cast<size_t>(f->n_rows / std::max<int64_t>(seg_max / 2, 1)) + 2;
    return (nseg + static_cast<size_t>(f->n_groups) + 4) * 24 + 4096;
}

// ------------------------------------------------------------------------------------------------
// Gram launch
// ------------------------------------------------------------------------------------------------
// choose tile rows / stages / warps so that the CTA's shared memory fits, then launch
template <typename T>
static int launch_gram(b200ols_ctx *c, GramParams &gp) {
    const int F = gp.F;
    const int KB = (F + 7) / 8;
    const int NC = gp.kd + 1 + gp.has_w + gp.has_mask;
    if (KB > 2 && !gp.fused) {  // 17 <= k <= 64: block pairs split across the consumer warps
        const size_t budget = static_cast<size_t>(c->smem_optin) - 1024;
        int R = c->tile_rows > 0 ? c->tile_rows : 96;
        int S = 0;
        for (;;) {
            const size_t sb = static_cast<size_t>(NC) * gram_col_stride<T>(R);
            S = static_cast<int>(std::min<size_t>(GRAM_MAX_STAGES, budget / sb));
            if (S >= 2 || R <= 16) break;
            R -= 8;
        }
        if (S < 2) return fail(B200OLS_ERR_UNSUPPORTED, "Gram tile does not fit in shared memory (%d columns)", NC);
        if (c->warps_per_cta > 0) S = std::min(S, std::max(2, c->warps_per_cta));
        gp.tile_rows = R;
        gp.stages = S;
        const size_t smem = static_cast<size_t>(S) * NC * gram_col_stride<T>(R);
        const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(c->sm_count, gp.nseg));
        ProfScope prof(c);
        CU(sizeof(T) == 8 ? gram_wide_launch_f64(gp, static_cast<unsigned>(grid), smem, c->stream)
                          : gram_wide_launch_f32(gp, static_cast<unsigned>(grid), smem, c->stream));
        c->launches++;
        return 0;
    }
    if (c->variant == 3 && KB <= 2 && !gp.fused && !gp.seg_group && c->multi_enabled &&
        c->plan_offsets.size() == static_cast<size_t>(gp.nseg) + 1) {
        // short groups: tiles of whole groups, one consumer warp per group (gram_multi.cuh)
        const size_t budget = static_cast<size_t>(c->smem_optin) - 1024;
        const size_t row_bytes = static_cast<size_t>(NC) * sizeof(T);
        // Tile = the largest one that still leaves THREE stages (~75 KB): tools/sweep_tiles.py, gram_ms default-28 KB-tile
        // -> three-stage tile: f64 k=8 n=64 0.263 -> 0.141, f32 k=16 n=64 0.542 -> 0.304, f32 k=16 n=256 (C3) 0.73 ->
        // 0.43 (four groups per tile instead of gram_cta teams), f64 k=3 n=100 0.222 -> 0.208; half / three-quarter
        // size tiles are slower everywhere.  Used when such a tile holds at least four groups.
        (void)row_bytes;
        int64_t R = static_cast<int64_t>(budget / 3 / NC / sizeof(T)) / 8 * 8;
        while (R > 64 && 3 * static_cast<size_t>(NC) * gram_col_stride<T>(static_cast<int>(R)) > budget) R -= 8;
        R = std::min<int64_t>(std::max<int64_t>(R, 64), 8192);
        {   // small frames: rather one (smaller) tile per SM than a few big ones
            const int64_t n_total = c->plan_offsets.back(), r4 = (4 * gp.max_seg_rows + 7) / 8 * 8;
            if (n_total / R < c->sm_count) R = std::max<int64_t>(std::min<int64_t>(R, (n_total / c->sm_count + 7) / 8 * 8), std::max<int64_t>(r4, 64));
        }
        if (c->tile_rows > 0) R = std::max<int64_t>((gp.max_seg_rows + 7) / 8 * 8, c->tile_rows);  // sweep hook
        const size_t sb = static_cast<size_t>(NC) * gram_col_stride<T>(static_cast<int>(R));
        int S = static_cast<int>(std::min<size_t>(GRAM_MAX_STAGES, budget / sb));
        if (c->warps_per_cta > 0) S = std::min(S, std::max(2, c->warps_per_cta));
        // one consumer warp per group only pays while a tile holds several groups (profiles/r01_sweep_team.json,
        // r01_sweep_c3.json: 64-row groups 0.14 ms vs 0.73 ms with gram_cta teams; two or three 256-row groups per tile
        // run slower than teams, four run 1.7x faster)
        if (S >= 3 && R <= 8192 && (c->tile_rows > 0 || R >= 4 * gp.max_seg_rows)) {
            if (!c->tile_valid || c->tile_rows_built != R) {
                // greedy runs of consecutive groups whose rows fit one tile (A-aligned start included)
                const std::vector<int64_t> &off = c->plan_offsets;
                const int64_t G = gp.nseg, A = 16 / static_cast<int64_t>(sizeof(T));
                std::vector<int64_t> tg;
                tg.reserve(static_cast<size_t>(G / 2 + 2));
                tg.push_back(0);
                int64_t start = 0;
                for (int64_t g = 0; g < G; ++g) {
                    const int64_t a_al = off[start] & ~(A - 1);
                    if (off[g + 1] - a_al > R && g > start) {  // group g does not fit any more: close the tile before it
                        tg.push_back(g);
                        start = g;
                    }
                }
                tg.push_back(G);
                const size_t bytes = sizeof(int64_t) * tg.size();
                if (bytes > c->tile_cap) {
                    if (c->tile_dev) c->retired.push_back(c->tile_dev);
                    c->tile_dev = nullptr;
                    void *qd = nullptr;
                    CU(cudaMalloc(&qd, bytes + 4096));
                    c->tile_dev = static_cast<int64_t *>(qd);
                    c->tile_cap = bytes + 4096;
                }
                TRY(pinned_reserve(c, c->pinned_off + bytes + 256));
                char *h = c->pinned + c->pinned_off;
                c->pinned_off += round_up(bytes, 256);
                std::memcpy(h, tg.data(), bytes);
                CU(cudaMemcpyAsync(c->tile_dev, h, bytes, cudaMemcpyHostToDevice, c->stream));
                c->tile_count = static_cast<int64_t>(tg.size()) - 1;
                c->tile_rows_built = R;
                c->tile_valid = true;
            }
            gp.tile_rows = static_cast<int>(R);
            gp.stages = S;
            MultiPlan mp{c->tile_dev, c->tile_count};
            const size_t smem = static_cast<size_t>(S) * sb;
            const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(c->sm_count, c->tile_count));
            ProfScope prof(c);
            CU(sizeof(T) == 8 ? gram_multi_launch_f64(KB, gp, mp, static_cast<unsigned>(grid), smem, c->stream)
                              : gram_multi_launch_f32(KB, gp, mp, static_cast<unsigned>(grid), smem, c->stream));
            c->launches++;
            return 0;
        }
    }
    if (c->variant == 3 && KB <= 2) {  // CTA-cooperative warp-specialised TMA pipeline (k <= 16)
        const size_t budget = static_cast<size_t>(c->smem_optin) - 1024;
        // Ring of published accumulator buffers (18 KB per entry for a 16-coefficient Gram).  A parity wait may be one
        // phase behind its barrier or level with it, never ahead, so: without teams (all consumers publish every
        // segment in order) the ring must be at least CTA_SOLVERS deep; with teams every barrier must be waited on
        // by the SAME warps in every phase: stages % teams == 0 (a stage always belongs to one team) and
        // ring % CTA_SOLVERS == 0 (a published buffer always goes to the same solver warp).
        int R = 0, S = 0;
        size_t fixed = 0;
        auto fit = [&](int depth) {
            fixed = cta_fixed_smem<T>(KB, F, depth);
            // default tile = the longest segment (whole groups per stage: fewest, largest bulk copies), shrunk until
            // at least two stages fit
            R = c->tile_rows > 0 ? c->tile_rows
                                 : static_cast<int>(std::min<int64_t>(std::max<int64_t>((gp.max_seg_rows + 7) / 8 * 8, 64), 4096));
            for (;;) {
                const size_t sb = static_cast<size_t>(NC) * gram_col_stride<T>(R);
                S = static_cast<int>(std::min<size_t>(GRAM_MAX_STAGES, (budget - fixed) / sb));
                if (S >= 2 || R <= 16) break;
                R -= 8;
            }
            if (c->warps_per_cta > 0) S = std::min(S, std::max(2, c->warps_per_cta));  // sweep hook: warps_per_cta caps the stages
        };
        gp.red_depth = cta_plain_depth(KB);
        fit(gp.red_depth);
        if (S < 2) return fail(B200OLS_ERR_UNSUPPORTED, "Gram tile does not fit in shared memory (%d columns)", NC);
        // short groups: fewer consumer warps per segment (teams), >= ~16 row octets per warp; needs one tile per segment
        gp.team = 0;
        if (gp.max_seg_rows <= R) {
            int team = CTA_CONSUMERS;
            while (team > 1 && gp.max_seg_rows <= 128 * (team / 2)) team >>= 1;
            if (c->ctas_per_sm > 0) {  // sweep hook: largest power of two <= min(8, ctas_per_sm)
                team = 1;
                while (team * 2 <= std::min(CTA_CONSUMERS, c->ctas_per_sm)) team *= 2;
            }
            if (team < CTA_CONSUMERS) {
                const int R0 = R, S0 = S;
                const int nt = CTA_CONSUMERS / team;
                const int depth = KB == 1 ? CTA_RED_DEPTH : (nt == 2 ? 2 : 3);  // depth * nt is a multiple of CTA_SOLVERS
                fit(depth);
                if (R == R0 && S / nt >= 1) {
                    S = S / nt * nt;
                    gp.red_depth = depth;
                    gp.team = team;
                } else {  // not enough stages for that many teams: keep the plain ring
                    fit(gp.red_depth);
                    (void)S0;
                }
            }
        }
        gp.tile_rows = R;
        gp.stages = S;
        const size_t smem = static_cast<size_t>(S) * NC * gram_col_stride<T>(R) + fixed;
        const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(c->sm_count, gp.nseg));
        ProfScope prof(c);
        CU(sizeof(T) == 8 ? gram_cta_launch_f64(KB, gp, static_cast<unsigned>(grid), smem, c->stream)
                          : gram_cta_launch_f32(KB, gp, static_cast<unsigned>(grid), smem, c->stream));
        c->launches++;
        return 0;
    }
    if (c->variant == 1 && KB <= 2) {  // direct-load variant
        int warps = std::min(c->warps_per_cta > 0 ? c->warps_per_cta : 8, 8);
        const int ctas = c->ctas_per_sm > 0 ? c->ctas_per_sm : 3;
        int64_t grid = std::min<int64_t>(static_cast<int64_t>(c->sm_count) * ctas, (gp.nseg + warps - 1) / warps);
        grid = std::max<int64_t>(grid, 1);
        const int U = c->unroll > 0 ? c->unroll : 2;
        ProfScope prof(c);
        CU(sizeof(T) == 8 ? gram_ldg_launch_f64(KB, U, gp, static_cast<unsigned>(grid), warps, c->stream)
                          : gram_ldg_launch_f32(KB, U, gp, static_cast<unsigned>(grid), warps, c->stream));
        c->launches++;
        return 0;
    }
    return fail(B200OLS_ERR_UNSUPPORTED, "no Gram kernel for %d coefficients (fused = %d)", F, gp.fused);
}

// ------------------------------------------------------------------------------------------------
// static models: least_squares / least_squares_coefficients
// ------------------------------------------------------------------------------------------------
struct StaticRoute {
    int route = ROUTE_CHOL;   // ROUTE_*
    bool ols_qr_guard = false;  // OLS branch (QR in the reference): re-solve ill-conditioned groups by QR
    bool svd_all = false;       // solve_method = "svd": every group by the Jacobi SVD kernel
    bool svd_wide = false;      // solve_method = None on the OLS branch: groups with n <= k take the SVD path
    bool svd_ridge = false;     // ridge-SVD formula (solve_ridge_svd)
    double rcond = NAN;
    double alpha = 0.0, l1_ratio = 0.0, tol = 1e-5;
    int64_t max_iter = 1000;
    int positive = 0;
};

// _get_least_squares_coefficients dispatch (src/expressions.rs:361-387) + argument checks of
// solve_ols / solve_ridge / solve_elastic_net (src/least_squares.rs:231,349,366,402-413)
static int resolve_route(const b200ols_ols_kwargs *kw, StaticRoute *r) {
    const double alpha = std::isnan(kw->alpha) ? 0.0 : kw->alpha;
    const bool positive = kw->positive != 0;
    const int m = kw->solve_method;
    if (m < B200OLS_SOLVE_NONE || m > B200OLS_SOLVE_CD_ACTIVE_SET) return fail(B200OLS_ERR_INVALID, "invalid solve_method detected!");
    if (kw->null_policy < B200OLS_NULL_IGNORE || kw->null_policy > B200OLS_NULL_DROP_Y_ZERO_X)
        return fail(B200OLS_ERR_INVALID, "Invalid null_policy detected!");  // drop_window is rolling-only
    r->alpha = alpha;
    r->positive = positive ? 1 : 0;
    r->max_iter = kw->max_iter < 0 ? 1000 : kw->max_iter;
    r->tol = std::isnan(kw->tol) ? 1e-5 : kw->tol;
    if (alpha == 0.0 && !positive && (m == B200OLS_SOLVE_NONE || m == B200OLS_SOLVE_SVD || m == B200OLS_SOLVE_QR)) {
        r->route = ROUTE_CHOL;
        r->ols_qr_guard = m != B200OLS_SOLVE_SVD;
        r->svd_all = m == B200OLS_SOLVE_SVD;          // solve_ols_svd (src/least_squares.rs:183-191)
        r->svd_wide = m == B200OLS_SOLVE_NONE;        // n <= k -> SVD (src/least_squares.rs:225-229)
        return 0;
    }
    const double l1 = std::isnan(kw->l1_ratio) ? 0.0 : kw->l1_ratio;
    if (alpha >= 0.0 && l1 == 0.0 && !positive) {
        if (m == B200OLS_SOLVE_NONE || m == B200OLS_SOLVE_CHOL) r->route = ROUTE_CHOL;
        else if (m == B200OLS_SOLVE_LU) r->route = ROUTE_LU;
        else if (m == B200OLS_SOLVE_SVD) {  // solve_ridge_svd (src/least_squares.rs:106-168)
            r->route = ROUTE_CHOL;
            r->svd_all = true;
            r->svd_ridge = true;
            r->rcond = kw->rcond;
        }
        else return fail(B200OLS_ERR_INVALID, "Only 'Cholesky', 'LU', & 'SVD' are currently supported solver methods for Ridge.");
        return 0;
    }
    // elastic net
    if (!(m == B200OLS_SOLVE_NONE || m == B200OLS_SOLVE_CD || m == B200OLS_SOLVE_CD_ACTIVE_SET))
        return fail(B200OLS_ERR_INVALID, "Only solve_method 'CD' (coordinate descent) is currently supported for Elastic Net / Lasso problems.");
    if (!(alpha > 0.0)) return fail(B200OLS_ERR_INVALID, "'alpha' must be strictly positive");
    r->l1_ratio = std::isnan(kw->l1_ratio) ? 0.5 : kw->l1_ratio;
    if (!(r->l1_ratio >= 0.0 && r->l1_ratio <= 1.0)) return fail(B200OLS_ERR_INVALID, "'l1_ratio' must be strictly between 0. and 1.");
    r->route = (m == B200OLS_SOLVE_CD_ACTIVE_SET) ? ROUTE_CD_ACTIVE : ROUTE_CD;
    return 0;
}

static constexpr int64_t SEG_MAX_ROWS = 4096;

// Groups are cut into segments only to create parallelism: the CTA-level Gram kernels walk a group of any
// length tile by tile, so with at least ~2 groups per SM nothing is split (and the solve stays fused).
// With few groups the segment length is chosen to give ~4 segments per SM (never below 1024 rows).
static int64_t choose_seg_max(const b200ols_ctx *c, int64_t n_groups, int64_t n_rows) {
    if (c->variant != 3) return SEG_MAX_ROWS;  // warp-per-segment kernels: bound the work of one warp
    if (n_groups >= 2 * static_cast<int64_t>(c->sm_count)) return INT64_MAX / 4;
    int64_t s = n_rows / (4 * static_cast<int64_t>(c->sm_count)) + 1;
    s = (s + 15) & ~static_cast<int64_t>(15);
    return std::max<int64_t>(s, 1024);
}
static constexpr double ILLCOND_RATIO = 1.0e7;  // squared-pivot ratio above which OLS is re-solved by QR

// ------------------------------------------------------------------------------------------------
// static models with more than 64 coefficients (big.cuh): general, slower path
// ------------------------------------------------------------------------------------------------
static int upload_small(b200ols_ctx *c, const void *src, size_t bytes, void **dev) {
    TRY(pinned_reserve(c, c->pinned_off + bytes + 256));
    char *h = c->pinned + c->pinned_off;
    c->pinned_off += round_up(bytes, 256);
    std::memcpy(h, src, bytes);
    char *d = arena_alloc<char>(c, bytes);
    CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
    *dev = d;
    return 0;
}

static int run_statistics_big(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw, const BigParams &bp,
                              const b200ols_statistics_output *so);

static int run_static_big(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw, const StaticRoute &rt, int mode,
                          b200ols_output *out, bool peer_mode, const b200ols_statistics_output *stats = nullptr) {
    const int kd = f->n_features, F = kd + (f->add_intercept ? 1 : 0), has_w = f->sample_weights ? 1 : 0;
    const int ncol = kd + 1 + has_w;
    const int64_t G = f->n_groups, N = f->n_rows;
    const size_t esz = f->dtype == B200OLS_F64 ? 8 : 4;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    const bool cd = rt.route == ROUTE_CD || rt.route == ROUTE_CD_ACTIVE;
    const bool any_svd = rt.svd_all || rt.svd_wide;

    std::vector<int64_t> offsets(static_cast<size_t>(G) + 1);
    if (f->group_offsets) std::memcpy(offsets.data(), f->group_offsets, sizeof(int64_t) * (G + 1));
    else { offsets[0] = 0; offsets[1] = N; }
    bool any_wide = false, any_tall = false;
    for (int64_t g = 0; g < G; ++g) {
        const int64_t len = offsets[g + 1] - offsets[g];
        if (len < 0) return fail(B200OLS_ERR_INVALID, "group_offsets must be non-decreasing");
        if (len > 0 && len <= F) any_wide = true;
        if (len > F) any_tall = true;
    }

    const size_t bm_bytes = static_cast<size_t>((N + 7) / 8);
    size_t bytes = static_cast<size_t>(ncol) * 64 + (static_cast<size_t>(G) + 8) * 16 + (1 << 20);
    if (f->memspace == B200OLS_HOST)
        bytes += static_cast<size_t>(ncol) * (round_up(static_cast<size_t>(N) * esz, 256) + round_up(bm_bytes, 256) + 512) +
                 round_up(static_cast<size_t>(N) * 8, 256) + static_cast<size_t>(N) * 9 + 8192;
    bytes += static_cast<size_t>(F + 1) * N * 8 + static_cast<size_t>(N) + static_cast<size_t>(G) * P * 8 + 4096;   // W, mask, records
    if (rt.route == ROUTE_CHOL || (stats && rt.route == ROUTE_LU)) bytes += static_cast<size_t>(G) * F * F * 8 + 256;      // factorisation copy
    bytes += static_cast<size_t>(G) * F * (8 + 4 + 8 + 2 * 4 + 2 * 8) + static_cast<size_t>(G) * 4 + 8192;           // beta, perm, z, cd scratch, flags
    if (any_svd) bytes += 2 * static_cast<size_t>(N) * F * 8 + static_cast<size_t>(G) * F * F * 8 + 4096;             // X^T, J, V
    if (stats) bytes += static_cast<size_t>(G) * (2 * static_cast<size_t>(F) * F + 5 * static_cast<size_t>(F) + STATS_GS + 3) * 8 + 8192;  // factor, L^-1, metrics
    TRY(arena_reserve(c, bytes));
    c->arena_off = 0;
    c->arena_overflow = false;
    TRY(pinned_begin(c));

    // columns on the device + pointer tables
    std::vector<const void *> vals(ncol);
    std::vector<const uint8_t *> valid(ncol);
    for (int cidx = 0; cidx < ncol; ++cidx) {
        const b200ols_column *col = cidx < kd ? &f->features[cidx] : (cidx == kd ? &f->target : f->sample_weights);
        if (N > 0 && !col->values) return fail(B200OLS_ERR_INVALID, "column %d: values is NULL", cidx);
        if (f->memspace == B200OLS_DEVICE) {
            vals[cidx] = col->values;
            valid[cidx] = col->validity;
        } else {
            char *d = arena_alloc<char>(c, static_cast<size_t>(N) * esz + 16);
            if (N > 0) {
                const StageSeg sg = {col->values, d, static_cast<size_t>(N) * esz};
                TRY(stage_h2d(c, &sg, 1));
            }
            vals[cidx] = d;
            valid[cidx] = nullptr;
            if (col->validity) {
                uint8_t *b = arena_alloc<uint8_t>(c, bm_bytes);
                CU(cudaMemcpyAsync(b, col->validity, bm_bytes, cudaMemcpyHostToDevice, c->stream));
                valid[cidx] = b;
            }
        }
    }
    BigParams bp;
    std::memset(&bp, 0, sizeof(bp));
    {
        void *d = nullptr;
        TRY(upload_small(c, vals.data(), sizeof(void *) * ncol, &d));
        bp.vals = static_cast<const void *const *>(d);
        TRY(upload_small(c, valid.data(), sizeof(void *) * ncol, &d));
        bp.valid = static_cast<const uint8_t *const *>(d);
        TRY(upload_small(c, offsets.data(), sizeof(int64_t) * (G + 1), &d));
        bp.group_off = static_cast<const int64_t *>(d);
    }
    if (f->row_index) {
        if (f->memspace == B200OLS_DEVICE || f->row_index_on_device) bp.row_index = f->row_index;
        else {
            int64_t *d = arena_alloc<int64_t>(c, static_cast<size_t>(N));
            CU(cudaMemcpyAsync(d, f->row_index, sizeof(int64_t) * N, cudaMemcpyHostToDevice, c->stream));
            bp.row_index = d;
        }
    }
    bool any_validity = false;
    for (int cidx = 0; cidx < ncol; ++cidx) any_validity = any_validity || valid[cidx] != nullptr;
    int fill, mask_kind;
    policy_to_prep(kw->null_policy, false, &fill, &mask_kind);
    bp.kd = kd;
    bp.intercept = f->add_intercept ? 1 : 0;
    bp.F = F;
    bp.has_w = has_w;
    bp.f32 = f->dtype == B200OLS_F32;
    bp.fill = fill;
    bp.mask_kind = any_validity ? mask_kind : MASK_NONE;
    bp.n_rows = N;
    bp.n_groups = G;
    bp.W = arena_alloc<double>(c, static_cast<size_t>(F + 1) * N + 8);
    bp.mask = arena_alloc<uint8_t>(c, static_cast<size_t>(N) + 8);
    bp.rec = arena_alloc<double>(c, static_cast<size_t>(G) * P);
    bp.work = (rt.route == ROUTE_CHOL || (stats && rt.route == ROUTE_LU)) ? arena_alloc<double>(c, static_cast<size_t>(G) * F * F) : nullptr;
    bp.beta = arena_alloc<double>(c, static_cast<size_t>(G) * F);
    bp.flags = arena_alloc<int32_t>(c, static_cast<size_t>(G));
    bp.route = rt.svd_all ? ROUTE_FLAGS_ONLY : rt.route;
    bp.alpha = rt.alpha;
    bp.l1_ratio = rt.l1_ratio;
    bp.tol = rt.tol;
    bp.illcond_ratio = ILLCOND_RATIO;
    bp.max_iter = rt.max_iter;
    bp.positive = rt.positive;
    bp.svd_ridge = rt.svd_ridge ? 1 : 0;
    bp.skip_wide = rt.svd_wide ? 1 : 0;
    bp.rcond = rt.rcond;
    bp.max_sweeps = 60;
    c->last_flags = bp.flags;
    c->last_flags_n = G;
    ARENA_GUARD(c);

    const unsigned row_blocks = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((N + 255) / 256, static_cast<int64_t>(c->sm_count) * 16)));
    if (N > 0) {
        big_materialise_kernel<<<row_blocks, 256, 0, c->stream>>>(bp);
        c->launches++;
    }
    {
        const int ntile = (F + BIG_TILE - 1) / BIG_TILE;
        const int64_t blocks = G * (static_cast<int64_t>(ntile) * (ntile + 1) / 2);
        if (blocks > 0x7fffffffLL) return fail(B200OLS_ERR_UNSUPPORTED, "too many groups x coefficient tiles for one launch");
        ProfScope prof(c);
        big_gram_kernel<<<static_cast<unsigned>(blocks), 256, 0, c->stream>>>(bp, ntile);
        c->launches++;
    }
    big_xty_kernel<<<static_cast<unsigned>((G * F * 32 + 255) / 256), 256, 0, c->stream>>>(bp);
    c->launches++;
    {
        const size_t smem = big_solve_smem(F);
        static size_t attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || smem > attr_set[dev])) {
            CU(cudaFuncSetAttribute(big_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            if (dev >= 0 && dev < 64) attr_set[dev] = smem;
        }
        big_solve_kernel<<<static_cast<unsigned>(G), BIG_SOLVE_THREADS, smem, c->stream>>>(bp);
        c->launches++;
    }
    CU(cudaGetLastError());
    if (cd && any_wide) {
        int *iws = arena_alloc<int>(c, static_cast<size_t>(G) * 2 * F);
        double *dws = arena_alloc<double>(c, static_cast<size_t>(G) * 2 * F);
        ARENA_GUARD(c);
        big_cd_wide_kernel<<<static_cast<unsigned>((G * 32 + 127) / 128), 128, 0, c->stream>>>(bp, iws, dws);
        c->launches++;
    }
    if (rt.ols_qr_guard && any_tall) {
        QrParams qp;
        std::memset(&qp, 0, sizeof(qp));
        qp.kd = kd;
        qp.intercept = bp.intercept;
        qp.F = F;
        qp.n_groups = G;
        qp.n_rows = N;
        qp.group_off = bp.group_off;
        qp.ws = bp.W;
        qp.beta = bp.beta;
        qp.flags = bp.flags;
        qp.materialised = 1;
        qp.perm_ws = arena_alloc<int>(c, static_cast<size_t>(G) * F);
        qp.z_ws = arena_alloc<double>(c, static_cast<size_t>(G) * F);
        ARENA_GUARD(c);
        CU(launch_qr_fallback_kernel(c->stream, qp, true));
        c->launches++;
    }
    if (any_svd) {
        if (any_tall && rt.svd_all) {  // n > k groups with solve_method = "svd": one-sided Jacobi on X itself
            SvdParams sv;
            std::memset(&sv, 0, sizeof(sv));
            sv.q.kd = kd;
            sv.q.intercept = bp.intercept;
            sv.q.F = F;
            sv.q.n_groups = G;
            sv.q.n_rows = N;
            sv.q.group_off = bp.group_off;
            sv.q.ws = bp.W;
            sv.q.beta = bp.beta;
            sv.q.flags = bp.flags;
            sv.vws = arena_alloc<double>(c, static_cast<size_t>(G) * F * F + 8);
            sv.all_groups = 1;
            sv.ridge = rt.svd_ridge ? 1 : 0;
            sv.alpha = rt.alpha;
            sv.rcond = rt.rcond;
            sv.max_sweeps = 60;
            sv.materialised = 1;
            sv.skip_rows_le_F = 1;
            ARENA_GUARD(c);
            CU(launch_svd_solve(c->stream, sv, true));
            c->launches++;
        }
        if (any_wide) {
            bp.T = arena_alloc<double>(c, static_cast<size_t>(N) * F + 8);
            bp.J = arena_alloc<double>(c, static_cast<size_t>(N) * F + 8);
            ARENA_GUARD(c);
            big_svd_wide_kernel<<<static_cast<unsigned>((G * 32 + 127) / 128), 128, 0, c->stream>>>(bp, rt.svd_all ? 1 : 0);
            c->launches++;
        }
    }
    CU(cudaGetLastError());
    if (stats) return run_statistics_big(c, f, kw, bp, stats);

    if (peer_mode) {
        PeerScatterParams ps;
        std::memset(&ps, 0, sizeof(ps));
        ps.beta = bp.beta;
        ps.n_peers = c->n_peers;
        for (int r = 0; r < c->n_peers; ++r) ps.peer[r] = c->peer_coef[r];
        ps.n = static_cast<int64_t>(G) * F;
        ps.base_elems = c->peer_group_base * F;
        peer_scatter_kernel<<<static_cast<unsigned>((ps.n + 255) / 256), 256, 0, c->stream>>>(ps);
        c->launches++;
        CU(cudaGetLastError());
        if (!out->values) return 0;
    }
    if (mode == B200OLS_COEFFICIENTS) {
        const size_t ob = static_cast<size_t>(G) * F * sizeof(double);
        if (f->memspace == B200OLS_HOST) {
            CU(cudaMemcpyAsync(out->values, bp.beta, ob, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (out->validity)
                for (size_t i = 0; i < static_cast<size_t>(G) * F; ++i) out->validity[i] = std::isnan(out->values[i]) ? 0 : 1;
        } else {
            CU(cudaMemcpyAsync(out->values, bp.beta, ob, cudaMemcpyDeviceToDevice, c->stream));
            if (out->validity) {
                nan_mask_kernel<<<static_cast<unsigned>((static_cast<size_t>(G) * F + 255) / 256), 256, 0, c->stream>>>(out->values, out->validity, static_cast<int64_t>(G) * F);
                c->launches++;
            }
        }
        return 0;
    }
    bp.residuals = mode == B200OLS_RESIDUALS;
    bp.drop_mask = (kw->null_policy == B200OLS_NULL_DROP && any_validity) ? 1 : 0;
    double *dout = out->values;
    uint8_t *dval = out->validity;
    if (f->memspace == B200OLS_HOST) {
        dout = arena_alloc<double>(c, static_cast<size_t>(N));
        dval = out->validity ? arena_alloc<uint8_t>(c, static_cast<size_t>(N)) : nullptr;
        ARENA_GUARD(c);
    }
    bp.out = dout;
    bp.out_valid = dval;
    if (N > 0) {
        big_predict_kernel<<<row_blocks, 256, 0, c->stream>>>(bp);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (f->memspace == B200OLS_HOST) {
        const StageSeg d2h[2] = {{out->values, dout, sizeof(double) * N}, {out->validity, dval, dval ? static_cast<size_t>(N) : 0}};
        TRY(stage_d2h(c, d2h, 2));  // pageable outputs drain through the pinned ring; returns with the stream synchronised
    }
    return 0;
}

static int run_statistics(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw, const Staged &st, const Plan &pl,
                          const double *partial, const double *beta, const b200ols_statistics_output *so);

// `stats` != nullptr: mode = "statistics" (b200ols_least_squares_statistics) — the coefficients are dispatched as
// for mode = coefficients, the Gram records are always written (no fused solve) and run_statistics() finishes
static void launch_predict(b200ols_ctx *c, const PredictParams &pr, bool f64, unsigned blocks) {
    if (f64) predict_kernel<double><<<blocks, 256, 0, c->stream>>>(pr);
    else predict_kernel<float><<<blocks, 256, 0, c->stream>>>(pr);
}

static int run_static_impl(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw, int mode, b200ols_output *out,
                           const b200ols_statistics_output *stats = nullptr) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    const bool peer_mode = !stats && c->n_peers > 0 && mode == B200OLS_COEFFICIENTS && f && f->memspace == B200OLS_DEVICE;
    b200ols_output no_out = {nullptr, nullptr};
    if (stats) out = &no_out;
    if (!kw || !out || (!out->values && !peer_mode && !stats)) return fail(B200OLS_ERR_INVALID, "NULL argument");
    TRY(validate_frame(f));
    if (peer_mode && c->peer_group_base + f->n_groups > c->peer_total_groups)
        return fail(B200OLS_ERR_INVALID, "peer gather: shard [%lld, %lld) exceeds total_groups %lld", (long long)c->peer_group_base,
                    (long long)(c->peer_group_base + f->n_groups), (long long)c->peer_total_groups);
    if (mode != B200OLS_PREDICTIONS && mode != B200OLS_RESIDUALS && mode != B200OLS_COEFFICIENTS)
        return fail(B200OLS_ERR_INVALID, "bad mode %d", mode);
    StaticRoute rt;
    TRY(resolve_route(kw, &rt));
    CU(cudaSetDevice(c->device));
    TRY(free_retired(c));

    const int F = f->n_features + (f->add_intercept ? 1 : 0);
    if (F > 64) return run_static_big(c, f, kw, rt, mode, out, peer_mode, stats);
    const int64_t G = f->n_groups, N = f->n_rows;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    // arena budget
    const int64_t seg_max = choose_seg_max(c, G, N);
    size_t bytes = stage_bytes_bound(f) + plan_bytes_bound(f, seg_max);
    const size_t nseg_bound = static_cast<size_t>(G) + static_cast<size_t>(N / std::max<int64_t>(seg_max / 2, 1)) + 2;
    bytes += nseg_bound * P * 8 + static_cast<size_t>(G) * (static_cast<size_t>(F) * F + 4 * F) * 8;  // partials + work
    bytes += static_cast<size_t>(G) * F * 8 + static_cast<size_t>(G) * 4 + 4096;                     // beta + flags
    if (f->memspace == B200OLS_HOST) bytes += static_cast<size_t>(N) * 9 + 4096;                         // out + validity
    if (rt.ols_qr_guard || rt.svd_all || rt.svd_wide)
        bytes += static_cast<size_t>(F + 1) * static_cast<size_t>(N) * 8 + static_cast<size_t>(G) * F * F * 8 + 8192;  // QR / SVD workspace
    if (stats) bytes += static_cast<size_t>(G) * (2 * static_cast<size_t>(F) * F + 6 * F + STATS_GS + 3) * 8 + 8 * static_cast<size_t>(G + 1) + 8192;
    bytes += 1 << 20;
    TRY(arena_reserve(c, bytes));
    c->arena_off = 0;
    c->arena_overflow = false;
    TRY(pinned_begin(c));

    Staged st;
    TRY(stage_frame(c, f, kw->null_policy, false, &st));
    Plan pl;
    TRY(build_plan(c, st, seg_max, &pl));

    double *beta = arena_alloc<double>(c, static_cast<size_t>(G) * F);
    int32_t *flags = arena_alloc<int32_t>(c, static_cast<size_t>(G));
    c->last_flags = flags;
    c->last_flags_n = G;

    // wide / under-determined groups take the LAPACK SVD path in the reference (src/least_squares.rs:225-229)
    GramParams gp;
    std::memset(&gp, 0, sizeof(gp));
    for (int j = 0; j < st.kd; ++j) gp.cols[j] = st.feat[j];
    gp.cols[st.kd] = st.y;
    int nc = st.kd + 1;
    if (st.has_w) gp.cols[nc++] = st.w;
    if (st.mask) gp.cols[nc++] = st.mask;
    gp.kd = st.kd;
    gp.intercept = st.intercept;
    gp.F = F;
    gp.has_w = st.has_w;
    gp.w_is_sqrt = st.w_is_sqrt;
    gp.has_mask = st.mask ? 1 : 0;
    gp.n_rows_pad = st.n_pad;
    gp.nseg = pl.nseg;
    gp.max_seg_rows = pl.split ? std::min<int64_t>(seg_max, st.max_group_rows) : st.max_group_rows;
    gp.seg_off = pl.seg_off;
    gp.seg_group = pl.seg_group;
    gp.alpha = rt.alpha;
    gp.use_lu = rt.route == ROUTE_LU;
    gp.illcond_ratio = ILLCOND_RATIO;
    gp.beta = beta;
    gp.flags = flags;
    const bool cd = rt.route == ROUTE_CD || rt.route == ROUTE_CD_ACTIVE;
    // short groups: the four solver warps of a CTA cannot keep up with the stream (a k x k solve is a 3-10 us
    // dependent chain), so the streaming kernel only writes the Gram records and batch_solve_kernel solves all
    // groups at full occupancy; long groups keep the solve fused (no record traffic, no second launch)
    const size_t group_bytes = static_cast<size_t>(st.max_group_rows) * (st.kd + 1 + st.has_w) * st.esz;
    const size_t fuse_min = c->fuse_min_bytes >= 0 ? static_cast<size_t>(c->fuse_min_bytes) : (F <= 8 ? 64u << 10 : 256u << 10);
    gp.fused = (!pl.split && !cd && F <= 16 && group_bytes >= fuse_min && !stats) ? 1 : 0;
    if (!gp.fused) gp.partial = arena_alloc<double>(c, static_cast<size_t>(pl.nseg) * P);
    // fused gather: only when the streaming kernel's beta is final (no QR / SVD re-solve afterwards)
    const bool peer_direct = peer_mode && gp.fused && !rt.ols_qr_guard && !rt.svd_all && !rt.svd_wide;
    if (peer_direct) {
        gp.n_peers = c->n_peers;
        for (int r = 0; r < c->n_peers; ++r) gp.peer_beta[r] = c->peer_coef[r];
        gp.peer_group_base = c->peer_group_base;
    }
    // per-step completion armed by b200ols_peer_arm_step: inside the fused-gather kernel when that is what runs
    const unsigned long long arm_signal = peer_mode ? c->armed_signal : 0, arm_wait = c->armed_wait;
    if (peer_mode) c->armed_signal = 0;
    bool arm_in_kernel = false;
    if (arm_signal && peer_direct && c->variant == 3 && F <= 16 && mode == B200OLS_COEFFICIENTS) {
        arm_in_kernel = true;
        gp.flag_n = c->n_flag_peers;
        gp.flag_rank = c->flag_rank;
        for (int r = 0; r < c->n_flag_peers; ++r) gp.flag_peer[r] = c->peer_flags[r];
        gp.flag_step = arm_signal;
        gp.flag_wait = arm_wait;
        gp.done_counter = c->done_counter;
        gp.flag_timeout = c->flag_timeout;
    }

    // mode = predictions | residuals with whole groups per tile: ONE kernel (gram_pred.cuh) keeps the tile in shared
    // memory until beta is known and predicts from it, so the features are read from HBM once instead of twice
    bool pred_fused = false;
    double *dout = out->values;
    uint8_t *dval = out->validity;
    if (mode != B200OLS_COEFFICIENTS && f->memspace == B200OLS_HOST) {
        dout = arena_alloc<double>(c, static_cast<size_t>(N));
        dval = out->validity ? arena_alloc<uint8_t>(c, static_cast<size_t>(N)) : nullptr;
    }
    if (mode != B200OLS_COEFFICIENTS && !stats && c->variant == 3 && c->pred_enabled && gp.fused && !st.prepped && !st.mask &&
        !st.row_index && !rt.svd_all && st.max_group_rows <= 8192) {
        const int KBp = (F + 7) / 8;
        const int NCp = st.kd + 1 + st.has_w;
        const int R = static_cast<int>(std::max<int64_t>((st.max_group_rows + 7) / 8 * 8, 64));
        const size_t budget = static_cast<size_t>(c->smem_optin) - 1024;
        const size_t fixed = st.esz == 8 ? pred_fixed_smem<double>(KBp, F) : pred_fixed_smem<float>(KBp, F);
        const size_t sb = static_cast<size_t>(NCp) * (st.esz == 8 ? gram_col_stride<double>(R) : gram_col_stride<float>(R));
        int S = budget > fixed ? static_cast<int>(std::min<size_t>(GRAM_MAX_STAGES, (budget - fixed) / sb)) : 0;
        if (c->warps_per_cta > 0) S = std::min(S, std::max(2, c->warps_per_cta));  // sweep hook, as for gram_cta
        if (S >= 2) {
            pred_fused = true;
            gp.tile_rows = R;
            gp.stages = S;
            PredOut po{dout, mode == B200OLS_RESIDUALS ? 1 : 0, c->pred_lag};
            const size_t smem = static_cast<size_t>(S) * sb + fixed;
            const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(c->sm_count, gp.nseg));
            ARENA_GUARD(c);
            if (dval) CU(cudaMemsetAsync(dval, 1, static_cast<size_t>(N), c->stream));  // null-free frame: every row is valid
            {
                ProfScope prof(c);
                CU(st.esz == 8 ? gram_pred_launch_f64(KBp, gp, po, static_cast<unsigned>(grid), smem, c->stream)
                               : gram_pred_launch_f32(KBp, gp, po, static_cast<unsigned>(grid), smem, c->stream));
            }
            c->launches++;
        }
    }

    ARENA_GUARD(c);
    if (!pred_fused) {
        if (f->dtype == B200OLS_F64) TRY(launch_gram<double>(c, gp)); else TRY(launch_gram<float>(c, gp));
    }

    auto make_predict_params = [&]() {
        PredictParams pr;
        std::memset(&pr, 0, sizeof(pr));
        for (int j = 0; j < st.kd; ++j) pr.cols[j] = st.feat[j];
        if (st.has_w) pr.cols[st.kd] = st.w;
        pr.kd = st.kd;
        pr.intercept = st.intercept;
        pr.F = F;
        pr.has_w = st.has_w;
        pr.w_is_sqrt = st.w_is_sqrt;
        pr.target = st.y_raw;
        pr.target_is_packed = st.y_raw_packed;
        pr.target_validity = st.y_validity;
        pr.mask = (kw->null_policy == B200OLS_NULL_DROP) ? st.mask : nullptr;
        pr.nseg = pl.nseg;
        pr.n_rows = N;
        pr.seg_off = pl.seg_off;
        pr.seg_group = pl.seg_group;
        pr.beta = beta;
        pr.row_index = st.row_index;
        pr.residuals = mode == B200OLS_RESIDUALS;
        pr.out = dout;
        pr.out_valid = dval;
        return pr;
    };
    if (!gp.fused) {
        SolveParams sp;
        std::memset(&sp, 0, sizeof(sp));
        sp.F = F;
        sp.n_groups = G;
        sp.partial = gp.partial;
        sp.group_seg_off = pl.group_seg_off;
        sp.work = arena_alloc<double>(c, static_cast<size_t>(G) * (static_cast<size_t>(F) * F + 4 * F));
        sp.beta = beta;
        sp.flags = flags;
        sp.route = rt.route;
        sp.alpha = rt.alpha;
        sp.l1_ratio = rt.l1_ratio;
        sp.tol = rt.tol;
        sp.illcond_ratio = ILLCOND_RATIO;
        sp.max_iter = rt.max_iter;
        sp.positive = rt.positive;
        if (cd) {
            // (Round 2 tried to write the predictions from the sub-warp that ran the coordinate descent — no predict pass,
            // coefficients never leave the registers.  Bit-identical, but slower on C3: 1.82 ms with scalar loads, 1.52 ms
            // with 16-byte loads against 1.31 ms for CD + predict_kernel; profiles/r02_c3_experiments.json.  Removed.)
            // k <= 16 and many groups: one thread per group (cd_thread.cu) instead of one sub-warp per group — same
            // arithmetic in the same order, ~2x fewer issue slots per coordinate step for 16x the groups per warp.
            // (Round 2 also cut the groups into chunks and ran predict(chunk i) on a second stream under cd(chunk i + 1):
            // 1.30 -> 1.19 ms on C3 with the sub-warp kernel, nothing with this one (0.98 ms either way), so it was
            // removed; profiles/r02_c3_experiments.json.)
            if (const char *v = std::getenv("B200OLS_CD_THREAD")) c->cd_thread = std::atoi(v);  // test hooks, read per call
            if (const char *v = std::getenv("B200OLS_CD_THREAD_BLOCKS")) c->cd_thread_blocks = std::atoi(v);
            if (c->cd_thread != 0 && F <= 16 && G >= (c->cd_thread > 1 ? 1 : 1024)) CU(launch_cd_thread(c->stream, sp, c->sm_count, c->cd_thread_blocks));
            else CU(launch_cd_solve(c->stream, sp, c->sm_count));
        } else if (F <= 16) {
            CU(launch_batch_solve(c->stream, sp));
        } else {
            const unsigned blocks = static_cast<unsigned>((G + 63) / 64);
            small_solve_kernel<<<blocks, 64, 0, c->stream>>>(sp);
            CU(cudaGetLastError());
        }
        c->launches++;
    }

    double *svd_ws = nullptr;  // the QR and SVD kernels share one absolute-row workspace
    if (rt.ols_qr_guard) {
        // groups whose Gram is too ill-conditioned for 1e-6 parity with the reference's QR are re-solved
        // from the data by Householder QR with column pivoting (faer col_piv_qr, src/least_squares.rs:195-205)
        QrParams qp;
        std::memset(&qp, 0, sizeof(qp));
        for (int j = 0; j < st.kd; ++j) qp.cols[j] = st.feat[j];
        qp.cols[st.kd] = st.y;
        qp.w = st.w;
        qp.mask = st.mask;
        qp.kd = st.kd;
        qp.intercept = st.intercept;
        qp.F = F;
        qp.w_is_sqrt = st.w_is_sqrt;
        qp.n_groups = G;
        qp.group_off = pl.split ? nullptr : pl.seg_off;
        if (pl.split) {
            // group offsets on the device (the split plan only holds segment offsets)
            const size_t ob = sizeof(int64_t) * (G + 1);
            TRY(pinned_reserve(c, c->pinned_off + ob + 256));
            char *h = c->pinned + c->pinned_off;
            c->pinned_off += round_up(ob, 256);
            std::memcpy(h, st.offsets.data(), ob);
            int64_t *d = arena_alloc<int64_t>(c, static_cast<size_t>(G) + 1);
            CU(cudaMemcpyAsync(d, h, ob, cudaMemcpyHostToDevice, c->stream));
            qp.group_off = d;
        }
        qp.beta = beta;
        qp.flags = flags;
        qp.n_rows = st.n;
        qp.ws = arena_alloc<double>(c, static_cast<size_t>(F + 1) * static_cast<size_t>(st.n) + 8);
        svd_ws = qp.ws;
        ARENA_GUARD(c);
        CU(launch_qr_fallback_kernel(c->stream, qp, f->dtype == B200OLS_F64));
        c->launches++;
    }

    if (rt.svd_all || rt.svd_wide) {
        // minimum-norm / ridge SVD path (solve_ols_svd, solve_ridge_svd): one-sided Jacobi on the data
        SvdParams sv;
        std::memset(&sv, 0, sizeof(sv));
        QrParams &qp = sv.q;
        for (int j = 0; j < st.kd; ++j) qp.cols[j] = st.feat[j];
        qp.cols[st.kd] = st.y;
        qp.w = st.w;
        qp.mask = st.mask;
        qp.kd = st.kd;
        qp.intercept = st.intercept;
        qp.F = F;
        qp.w_is_sqrt = st.w_is_sqrt;
        qp.n_groups = G;
        qp.n_rows = st.n;
        qp.beta = beta;
        qp.flags = flags;
        if (!pl.split) {
            qp.group_off = pl.seg_off;
        } else {
            const size_t ob = sizeof(int64_t) * (G + 1);
            TRY(pinned_reserve(c, c->pinned_off + ob + 256));
            char *h = c->pinned + c->pinned_off;
            c->pinned_off += round_up(ob, 256);
            std::memcpy(h, st.offsets.data(), ob);
            int64_t *d = arena_alloc<int64_t>(c, static_cast<size_t>(G) + 1);
            CU(cudaMemcpyAsync(d, h, ob, cudaMemcpyHostToDevice, c->stream));
            qp.group_off = d;
        }
        qp.ws = svd_ws ? svd_ws : arena_alloc<double>(c, static_cast<size_t>(F + 1) * static_cast<size_t>(st.n) + 8);
        sv.vws = arena_alloc<double>(c, static_cast<size_t>(G) * F * F + 8);
        sv.all_groups = rt.svd_all ? 1 : 0;
        sv.ridge = rt.svd_ridge ? 1 : 0;
        sv.alpha = rt.alpha;
        sv.rcond = rt.rcond;
        sv.max_sweeps = 40;
        ARENA_GUARD(c);
        CU(launch_svd_solve(c->stream, sv, f->dtype == B200OLS_F64));
        c->launches++;
    }

    if (stats) return run_statistics(c, f, kw, st, pl, gp.partial, beta, stats);

    // outputs
    if (peer_mode && !peer_direct) {
        PeerScatterParams ps;
        std::memset(&ps, 0, sizeof(ps));
        ps.beta = beta;
        ps.n_peers = c->n_peers;
        for (int r = 0; r < c->n_peers; ++r) ps.peer[r] = c->peer_coef[r];
        ps.n = static_cast<int64_t>(G) * F;
        ps.base_elems = c->peer_group_base * F;
        peer_scatter_kernel<<<static_cast<unsigned>((ps.n + 255) / 256), 256, 0, c->stream>>>(ps);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (arm_signal && !arm_in_kernel) TRY(peer_step(c, arm_signal, arm_wait, arm_wait > 0 ? 3 : 1));
    if (peer_mode && !out->values) return 0;
    if (mode == B200OLS_COEFFICIENTS) {
        const size_t ob = static_cast<size_t>(G) * F * sizeof(double);
        if (f->memspace == B200OLS_HOST) {
            CU(cudaMemcpyAsync(out->values, beta, ob, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (out->validity)
                for (size_t i = 0; i < static_cast<size_t>(G) * F; ++i) out->validity[i] = std::isnan(out->values[i]) ? 0 : 1;
        } else {
            CU(cudaMemcpyAsync(out->values, beta, ob, cudaMemcpyDeviceToDevice, c->stream));
            if (out->validity) {
                nan_mask_kernel<<<static_cast<unsigned>((static_cast<size_t>(G) * F + 255) / 256), 256, 0, c->stream>>>(out->values, out->validity, static_cast<int64_t>(G) * F);
                c->launches++;
            }
        }
        return 0;
    }

    PredictParams pr = make_predict_params();
    if (pred_fused) {  // only the groups whose beta was re-solved after the fused kernel (pivoted QR / min-norm SVD) are redone
        pr.flags = flags;
        pr.only_flags = FLAG_QR | FLAG_SVD;
    }
    if (!pred_fused || rt.ols_qr_guard || rt.svd_wide) {
        const int64_t warps_needed = (N + PREDICT_CHUNK - 1) / PREDICT_CHUNK;
        const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((warps_needed + 7) / 8, static_cast<int64_t>(c->sm_count) * 8));
        launch_predict(c, pr, f->dtype == B200OLS_F64, static_cast<unsigned>(blocks));
        c->launches++;
        CU(cudaGetLastError());
    }
    if (f->memspace == B200OLS_HOST) {
        const StageSeg d2h[2] = {{out->values, dout, sizeof(double) * N}, {out->validity, dval, dval ? static_cast<size_t>(N) : 0}};
        TRY(stage_d2h(c, d2h, 2));  // pageable outputs drain through the pinned ring; returns with the stream synchronised
    }
    return 0;
}

static int run_static(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw, int mode, b200ols_output *out) {
    const int rc = run_static_impl(c, f, kw, mode, out);
    if (c && c->pinned) pinned_end(c);
    return rc;
}

extern "C" int b200ols_least_squares(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw, int mode,
                                     b200ols_output *out) {
    if (mode == B200OLS_COEFFICIENTS) return fail(B200OLS_ERR_INVALID, "use b200ols_least_squares_coefficients for mode=coefficients");
    return run_static(c, f, kw, mode, out);
}

extern "C" int b200ols_least_squares_coefficients(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw,
                                                  b200ols_output *out) {
    return run_static(c, f, kw, B200OLS_COEFFICIENTS, out);
}

// ------------------------------------------------------------------------------------------------
// mode = "statistics" (src/expressions.rs:469-509, src/statistics.rs): stats.cuh after the Gram records
// ------------------------------------------------------------------------------------------------
static int device_group_offsets(b200ols_ctx *c, const Staged &st, const Plan &pl, const int64_t **out) {
    if (!pl.split) {
        *out = pl.seg_off;
        return 0;
    }
    const size_t ob = sizeof(int64_t) * (st.n_groups + 1);  // the split plan only holds segment offsets
    TRY(pinned_reserve(c, c->pinned_off + ob + 256));
    char *h = c->pinned + c->pinned_off;
    c->pinned_off += round_up(ob, 256);
    std::memcpy(h, st.offsets.data(), ob);
    int64_t *d = arena_alloc<int64_t>(c, static_cast<size_t>(st.n_groups) + 1);
    CU(cudaMemcpyAsync(d, h, ob, cudaMemcpyHostToDevice, c->stream));
    *out = d;
    return 0;
}

// mode = "statistics" behind the general path (more than 64 coefficients): big_stats.cuh on what run_static_big left
static int run_statistics_big(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw, const BigParams &bp,
                              const b200ols_statistics_output *so) {
    if (!so->r2 || !so->mae || !so->mse || !so->coefficients || !so->standard_errors || !so->t_values || !so->p_values)
        return fail(B200OLS_ERR_INVALID, "statistics output: NULL array");
    if (kw->alpha != kw->alpha) return fail(B200OLS_ERR_INVALID, "mode=statistics requires alpha (called `Option::unwrap()` on a `None` value)");
    const int F = bp.F;
    const int64_t G = bp.n_groups;
    const size_t GF = static_cast<size_t>(G) * F;
    BigStatsParams bs;
    std::memset(&bs, 0, sizeof(bs));
    bs.F = F;
    bs.n_groups = G;
    bs.n_rows = bp.n_rows;
    bs.group_off = bp.group_off;
    bs.W = bp.W;
    bs.mask = bp.mask;
    bs.rec = bp.rec;
    bs.beta = bp.beta;
    bs.alpha = kw->alpha;
    bs.A = arena_alloc<double>(c, GF * F);
    bs.M = arena_alloc<double>(c, GF * F);
    bs.beta2 = arena_alloc<double>(c, GF);
    bs.inv_diag = arena_alloc<double>(c, GF);
    bs.gstat = arena_alloc<double>(c, static_cast<size_t>(G) * STATS_GS);
    StatsParams sp;  // stats_final_kernel's view of the same arrays
    std::memset(&sp, 0, sizeof(sp));
    sp.F = F;
    sp.n_groups = G;
    sp.alpha = kw->alpha;
    sp.beta2 = bs.beta2;
    sp.inv_diag = bs.inv_diag;
    sp.gstat = bs.gstat;
    const bool host = f->memspace == B200OLS_HOST;
    double *blk = host ? arena_alloc<double>(c, 3 * static_cast<size_t>(G) + 3 * GF) : nullptr;
    sp.r2 = host ? blk : so->r2;
    sp.mae = host ? blk + G : so->mae;
    sp.mse = host ? blk + 2 * G : so->mse;
    sp.se = host ? blk + 3 * G : so->standard_errors;
    sp.tv = host ? blk + 3 * G + GF : so->t_values;
    sp.pv = host ? blk + 3 * G + 2 * GF : so->p_values;
    sp.bad_df = arena_alloc<int32_t>(c, 4);
    ARENA_GUARD(c);
    CU(cudaMemsetAsync(sp.bad_df, 0, sizeof(int32_t), c->stream));
    CU(cudaMemsetAsync(bs.gstat, 0, sizeof(double) * G * STATS_GS, c->stream));
    if (G > 0) {
        const size_t smem = big_stats_factor_smem(F);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(big_stats_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        big_stats_factor_kernel<<<static_cast<unsigned>(G), BIG_SOLVE_THREADS, smem, c->stream>>>(bs);
        big_stats_invdiag_kernel<<<dim3(static_cast<unsigned>((F + 255) / 256), static_cast<unsigned>(std::min<int64_t>(G, 65535))), 256, 0, c->stream>>>(bs);
        const unsigned rb = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(G, static_cast<int64_t>(c->sm_count) * 8)));
        big_stats_resid_kernel<<<rb, 256, 0, c->stream>>>(bs);
        stats_final_kernel<<<static_cast<unsigned>((GF + 127) / 128), 128, 0, c->stream>>>(sp);
        c->launches += 4;
        CU(cudaGetLastError());
    }
    int32_t bad = 0;
    if (host) {
        CU(cudaMemcpyAsync(so->r2, sp.r2, sizeof(double) * G, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->mae, sp.mae, sizeof(double) * G, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->mse, sp.mse, sizeof(double) * G, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->standard_errors, sp.se, sizeof(double) * GF, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->t_values, sp.tv, sizeof(double) * GF, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->p_values, sp.pv, sizeof(double) * GF, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->coefficients, bp.beta, sizeof(double) * GF, cudaMemcpyDeviceToHost, c->stream));
    } else {
        CU(cudaMemcpyAsync(so->coefficients, bp.beta, sizeof(double) * GF, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaMemcpyAsync(&bad, sp.bad_df, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (bad > 0) return fail(B200OLS_ERR_INVALID, "Degrees of freedom <= 0. Cannot compute standard errors. (%d group(s))", bad);
    return 0;
}

static int run_statistics(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw, const Staged &st, const Plan &pl,
                          const double *partial, const double *beta, const b200ols_statistics_output *so) {
    if (!so->r2 || !so->mae || !so->mse || !so->coefficients || !so->standard_errors || !so->t_values || !so->p_values)
        return fail(B200OLS_ERR_INVALID, "statistics output: NULL array");
    // `let lambda = kwargs.alpha.unwrap()` (src/expressions.rs:475): alpha = None panics in the reference
    if (kw->alpha != kw->alpha) return fail(B200OLS_ERR_INVALID, "mode=statistics requires alpha (called `Option::unwrap()` on a `None` value)");
    const int F = st.F;
    const int64_t G = st.n_groups;
    const size_t GF = static_cast<size_t>(G) * F;
    StatsParams sp;
    std::memset(&sp, 0, sizeof(sp));
    for (int j = 0; j < st.kd; ++j) sp.cols[j] = st.feat[j];
    sp.cols[st.kd] = st.y;
    sp.w = st.w;
    sp.mask = st.mask;
    sp.kd = st.kd;
    sp.intercept = st.intercept;
    sp.F = F;
    sp.w_is_sqrt = st.w_is_sqrt;
    sp.n_groups = G;
    TRY(device_group_offsets(c, st, pl, &sp.group_off));
    sp.partial = partial;
    sp.group_seg_off = pl.group_seg_off;
    sp.alpha = kw->alpha;
    sp.beta = beta;
    sp.work = arena_alloc<double>(c, static_cast<size_t>(G) * (2 * static_cast<size_t>(F) * F + F));
    sp.beta2 = arena_alloc<double>(c, GF);
    sp.inv_diag = arena_alloc<double>(c, GF);
    sp.gstat = arena_alloc<double>(c, static_cast<size_t>(G) * STATS_GS);
    const bool host = f->memspace == B200OLS_HOST;
    // host frames: results are produced in one arena block [r2 | mae | mse | se | t | p] and copied back
    double *blk = host ? arena_alloc<double>(c, 3 * static_cast<size_t>(G) + 3 * GF) : nullptr;
    sp.r2 = host ? blk : so->r2;
    sp.mae = host ? blk + G : so->mae;
    sp.mse = host ? blk + 2 * G : so->mse;
    sp.se = host ? blk + 3 * G : so->standard_errors;
    sp.tv = host ? blk + 3 * G + GF : so->t_values;
    sp.pv = host ? blk + 3 * G + 2 * GF : so->p_values;
    sp.bad_df = arena_alloc<int32_t>(c, 4);
    ARENA_GUARD(c);
    CU(cudaMemsetAsync(sp.bad_df, 0, sizeof(int32_t), c->stream));
    stats_inverse_kernel<<<static_cast<unsigned>((G + 63) / 64), 64, 0, c->stream>>>(sp);
    const unsigned rb = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(G, static_cast<int64_t>(c->sm_count) * 8)));
    if (f->dtype == B200OLS_F64) stats_resid_kernel<double><<<rb, 256, 0, c->stream>>>(sp);
    else stats_resid_kernel<float><<<rb, 256, 0, c->stream>>>(sp);
    stats_final_kernel<<<static_cast<unsigned>((GF + 127) / 128), 128, 0, c->stream>>>(sp);
    c->launches += 3;
    CU(cudaGetLastError());
    int32_t bad = 0;
    if (host) {
        CU(cudaMemcpyAsync(so->r2, sp.r2, sizeof(double) * G, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->mae, sp.mae, sizeof(double) * G, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->mse, sp.mse, sizeof(double) * G, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->standard_errors, sp.se, sizeof(double) * GF, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->t_values, sp.tv, sizeof(double) * GF, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->p_values, sp.pv, sizeof(double) * GF, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(so->coefficients, beta, sizeof(double) * GF, cudaMemcpyDeviceToHost, c->stream));
    } else {
        CU(cudaMemcpyAsync(so->coefficients, beta, sizeof(double) * GF, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaMemcpyAsync(&bad, sp.bad_df, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    // src/statistics.rs:131-134 asserts per call; one batched call carries every group, so any failing group fails it
    if (bad > 0) return fail(B200OLS_ERR_INVALID, "Degrees of freedom <= 0. Cannot compute standard errors. (%d group(s))", bad);
    return 0;
}

extern "C" int b200ols_least_squares_statistics(b200ols_ctx *c, const b200ols_frame *f, const b200ols_ols_kwargs *kw,
                                                const b200ols_statistics_output *out) {
    if (!out) return fail(B200OLS_ERR_INVALID, "NULL argument");
    const int rc = run_static_impl(c, f, kw, B200OLS_COEFFICIENTS, nullptr, out);
    if (c && c->pinned) pinned_end(c);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// multi_target_least_squares (src/expressions.rs:521-591): multi_target.cuh
// ------------------------------------------------------------------------------------------------
static int run_multi_target_impl(b200ols_ctx *c, const b200ols_frame *f0, int32_t m, const b200ols_column *targets,
                                 const b200ols_ols_kwargs *kw, int mode, b200ols_output *out) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (!f0 || !kw || !out || !out->values || !targets) return fail(B200OLS_ERR_INVALID, "NULL argument");
    if (m < 1) return fail(B200OLS_ERR_INVALID, "the first series in a multi-target regression must be a struct (n_targets >= 1)");
    // compute_multi_target_least_squares (polars_ols/least_squares.py:303-318)
    if (kw->positive || !(kw->l1_ratio != kw->l1_ratio || kw->l1_ratio == 0.0))
        return fail(B200OLS_ERR_INVALID, "Multi-target regression is only supported for unconstrained OLS & Ridge problems.");
    if (kw->solve_method != B200OLS_SOLVE_NONE && kw->solve_method != B200OLS_SOLVE_SVD)
        return fail(B200OLS_ERR_INVALID, "only solve_method='svd' is supported for multi-target regressions");
    if (mode != B200OLS_PREDICTIONS && mode != B200OLS_RESIDUALS)
        return fail(B200OLS_ERR_INVALID, "Only mode={'predictions', 'residuals'} is currently supported.");
    const double alpha = (kw->alpha == kw->alpha) ? kw->alpha : 0.0;
    const bool ridge = alpha > 0.0;  // src/least_squares.rs:254-259
    const int kd = f0->n_features, icpt = f0->add_intercept ? 1 : 0;
    const int F = kd + icpt, kdx = kd + m - 1, FX = kdx + icpt;
    if (F < 1) return fail(B200OLS_ERR_INVALID, "must pass at least 2 series");
    if (FX > 64 || kdx + 3 > GRAM_MAX_COLS)
        return fail(B200OLS_ERR_UNSUPPORTED, "multi-target: features + targets (%d) above 64 is not implemented on the device", FX);
    // the extended frame: targets 0..m-2 ride along as feature columns, target m-1 is the frame's target
    std::vector<b200ols_column> xcols(static_cast<size_t>(kdx));
    for (int j = 0; j < kd; ++j) {
        if (!f0->features) return fail(B200OLS_ERR_INVALID, "features is NULL");
        xcols[j] = f0->features[j];
    }
    for (int t = 0; t < m - 1; ++t) xcols[kd + t] = targets[t];
    b200ols_frame fx = *f0;
    fx.n_features = kdx;
    fx.features = xcols.data();
    fx.target = targets[m - 1];
    const b200ols_frame *f = &fx;
    TRY(validate_frame(f));
    CU(cudaSetDevice(c->device));
    TRY(free_retired(c));
    // every array of the reference is built with fill_zero = true (src/expressions.rs:549-550,569): 'ignore' fills 0 too
    const int policy = kw->null_policy == B200OLS_NULL_IGNORE ? B200OLS_NULL_ZERO : kw->null_policy;

    const int64_t G = f->n_groups, N = f->n_rows;
    const size_t PX = static_cast<size_t>(FX) * FX + FX + 1;
    const int64_t seg_max = choose_seg_max(c, G, N);
    size_t bytes = stage_bytes_bound(f) + plan_bytes_bound(f, seg_max);
    const size_t nseg_bound = static_cast<size_t>(G) + static_cast<size_t>(N / std::max<int64_t>(seg_max / 2, 1)) + 2;
    bytes += nseg_bound * PX * 8 + static_cast<size_t>(G) * (2 * static_cast<size_t>(F) * F + F + static_cast<size_t>(m) * F) * 8;
    bytes += static_cast<size_t>(G) * 4 + 8 * static_cast<size_t>(G + 1) + 8192;
    bytes += static_cast<size_t>(F + m) * static_cast<size_t>(N) * 8 + 8192;  // SVD workspace
    if (f->memspace == B200OLS_HOST) bytes += static_cast<size_t>(m) * static_cast<size_t>(N) * 9 + 8192;
    bytes += 1 << 20;
    TRY(arena_reserve(c, bytes));
    c->arena_off = 0;
    c->arena_overflow = false;
    TRY(pinned_begin(c));

    Staged st;
    TRY(stage_frame(c, f, policy, false, &st, m - 1));
    Plan pl;
    TRY(build_plan(c, st, seg_max, &pl));

    double *beta = arena_alloc<double>(c, static_cast<size_t>(G) * m * F);
    int32_t *flags = arena_alloc<int32_t>(c, static_cast<size_t>(G));
    c->last_flags = flags;
    c->last_flags_n = G;

    GramParams gp;
    std::memset(&gp, 0, sizeof(gp));
    for (int j = 0; j < st.kd; ++j) gp.cols[j] = st.feat[j];
    gp.cols[st.kd] = st.y;
    int nc = st.kd + 1;
    if (st.has_w) gp.cols[nc++] = st.w;
    if (st.mask) gp.cols[nc++] = st.mask;
    gp.kd = st.kd;
    gp.intercept = st.intercept;
    gp.F = FX;
    gp.has_w = st.has_w;
    gp.w_is_sqrt = st.w_is_sqrt;
    gp.has_mask = st.mask ? 1 : 0;
    gp.n_rows_pad = st.n_pad;
    gp.nseg = pl.nseg;
    gp.max_seg_rows = pl.split ? std::min<int64_t>(seg_max, st.max_group_rows) : st.max_group_rows;
    gp.seg_off = pl.seg_off;
    gp.seg_group = pl.seg_group;
    gp.illcond_ratio = ILLCOND_RATIO;
    gp.beta = beta;  // unused: the solve is never fused here
    gp.flags = flags;
    gp.fused = 0;
    gp.partial = arena_alloc<double>(c, static_cast<size_t>(pl.nseg) * PX);
    ARENA_GUARD(c);
    if (f->dtype == B200OLS_F64) TRY(launch_gram<double>(c, gp)); else TRY(launch_gram<float>(c, gp));

    MultiSolveParams ms;
    std::memset(&ms, 0, sizeof(ms));
    ms.kd = kd;
    ms.intercept = icpt;
    ms.F = F;
    ms.m = m;
    ms.FX = FX;
    ms.n_groups = G;
    ms.partial = gp.partial;
    ms.group_seg_off = pl.group_seg_off;
    ms.alpha = ridge ? alpha : 0.0;
    ms.illcond_ratio = ILLCOND_RATIO;
    ms.work = arena_alloc<double>(c, static_cast<size_t>(G) * (static_cast<size_t>(F) * F + F));
    ms.beta = beta;
    ms.flags = flags;
    ARENA_GUARD(c);
    multi_solve_kernel<<<static_cast<unsigned>((G + 63) / 64), 64, 0, c->stream>>>(ms);
    c->launches++;
    CU(cudaGetLastError());

    {   // flagged groups (ill-conditioned, failed factorisation, n <= k): the reference's SVD itself, m right-hand sides
        SvdParams sv;
        std::memset(&sv, 0, sizeof(sv));
        QrParams &qp = sv.q;
        for (int j = 0; j < kd; ++j) qp.cols[j] = st.feat[j];
        for (int t = 0; t < m - 1; ++t) qp.cols[kd + t] = st.feat[kd + t];
        qp.cols[kd + m - 1] = st.y;
        qp.w = st.w;
        qp.mask = st.mask;
        qp.kd = kd;
        qp.intercept = icpt;
        qp.F = F;
        qp.w_is_sqrt = st.w_is_sqrt;
        qp.n_groups = G;
        qp.n_rows = st.n;
        qp.beta = beta;
        qp.flags = flags;
        TRY(device_group_offsets(c, st, pl, &qp.group_off));
        qp.ws = arena_alloc<double>(c, static_cast<size_t>(F + m) * static_cast<size_t>(st.n) + 8);
        sv.vws = arena_alloc<double>(c, static_cast<size_t>(G) * F * F + 8);
        sv.all_groups = 0;
        sv.flag_mask = FLAG_WIDE | FLAG_ILLCOND | FLAG_LU_FALLBACK;
        sv.n_rhs = m;
        sv.ridge = ridge ? 1 : 0;
        sv.alpha = alpha;
        sv.rcond = kw->rcond;
        sv.max_sweeps = 40;
        ARENA_GUARD(c);
        CU(launch_svd_solve(c->stream, sv, f->dtype == B200OLS_F64));
        c->launches++;
    }

    // predictions / residuals: one pass of the row-parallel predict kernel per target
    double *dout = out->values;
    uint8_t *dval = out->validity;
    if (f->memspace == B200OLS_HOST) {
        dout = arena_alloc<double>(c, static_cast<size_t>(N) * m);
        dval = out->validity ? arena_alloc<uint8_t>(c, static_cast<size_t>(N) * m) : nullptr;
        ARENA_GUARD(c);
    }
    for (int t = 0; t < m; ++t) {
        PredictParams pr;
        std::memset(&pr, 0, sizeof(pr));
        for (int j = 0; j < kd; ++j) pr.cols[j] = st.feat[j];
        if (st.has_w) pr.cols[kd] = st.w;
        pr.kd = kd;
        pr.intercept = icpt;
        pr.F = F;
        pr.has_w = st.has_w;
        pr.w_is_sqrt = st.w_is_sqrt;
        const int raw_idx = (t < m - 1) ? kd + t : kdx;  // position of target t among the staged columns
        pr.target = st.raw[raw_idx];
        pr.target_is_packed = 0;
        pr.target_validity = st.raw_bm[raw_idx];
        pr.mask = (kw->null_policy == B200OLS_NULL_DROP) ? st.mask : nullptr;
        pr.nseg = pl.nseg;
        pr.n_rows = N;
        pr.seg_off = pl.seg_off;
        pr.seg_group = pl.seg_group;
        pr.beta = beta + static_cast<size_t>(t) * F;
        pr.beta_stride = static_cast<int64_t>(m) * F;
        pr.nan_is_null = 1;
        pr.row_index = st.row_index;
        pr.residuals = mode == B200OLS_RESIDUALS;
        pr.out = dout + static_cast<size_t>(t) * N;
        pr.out_valid = dval ? dval + static_cast<size_t>(t) * N : nullptr;
        const int64_t warps_needed = (N + PREDICT_CHUNK - 1) / PREDICT_CHUNK;
        const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((warps_needed + 7) / 8, static_cast<int64_t>(c->sm_count) * 8));
        launch_predict(c, pr, f->dtype == B200OLS_F64, static_cast<unsigned>(blocks));
        c->launches++;
        CU(cudaGetLastError());
    }
    if (f->memspace == B200OLS_HOST) {
        const StageSeg d2h[2] = {{out->values, dout, sizeof(double) * N * m}, {out->validity, dval, dval ? static_cast<size_t>(N) * m : 0}};
        TRY(stage_d2h(c, d2h, 2));  // pageable outputs drain through the pinned ring; returns with the stream synchronised
    }
    return 0;
}

extern "C" int b200ols_multi_target_least_squares(b200ols_ctx *c, const b200ols_frame *f, int32_t n_targets,
                                                  const b200ols_column *targets, const b200ols_ols_kwargs *kw, int mode,
                                                  b200ols_output *out) {
    const int rc = run_multi_target_impl(c, f, n_targets, targets, kw, mode, out);
    if (c && c->pinned) pinned_end(c);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// predict (src/expressions.rs:706-741): row-wise dot of features with a per-row coefficient struct
// ------------------------------------------------------------------------------------------------
struct RowDotParams {
    const void *feat[GRAM_MAX_COLS];
    const uint8_t *feat_valid[GRAM_MAX_COLS];
    const double *coef[GRAM_MAX_COLS];
    const uint8_t *coef_valid[GRAM_MAX_COLS];
    int n_coef, n_feat, fill_nan, drop;
    int accumulate;  // continue the row sums of an earlier batch of coefficient columns (more than 64 coefficients)
    int64_t n_rows;
    double *out;
    uint8_t *out_valid;
};

template <typename T>
static __global__ void __launch_bounds__(256) rowdot_kernel(const RowDotParams p) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < p.n_rows; r += stride) {
        double acc = p.accumulate ? p.out[r] : 0.0;
        bool all_valid = (p.accumulate && p.drop && p.out_valid) ? (p.out_valid[r] != 0) : true;
        for (int j = 0; j < p.n_coef; ++j) {
            bool cv = p.coef_valid[j] ? ((p.coef_valid[j][r >> 3] >> (r & 7)) & 1) : true;
            const double c = cv ? p.coef[j][r] : static_cast<double>(NAN);
            double x = 1.0;  // the `const` feature appended by add_intercept (polars_ols/least_squares.py:479-483)
            if (j < p.n_feat) {
                const bool xv = p.feat_valid[j] ? ((p.feat_valid[j][r >> 3] >> (r & 7)) & 1) : true;
                all_valid = all_valid && xv;
                x = xv ? static_cast<double>(static_cast<const T *>(p.feat[j])[r]) : (p.fill_nan ? static_cast<double>(NAN) : 0.0);
            }
            all_valid = all_valid && cv;
            acc += x * c;  // (&features * &coefficients).sum_axis(Axis(1))
        }
        p.out[r] = acc;
        if (p.out_valid) p.out_valid[r] = (!p.drop || all_valid) ? 1 : 0;
    }
}

extern "C" int b200ols_predict(b200ols_ctx *c, int64_t n_rows, int32_t n_coef, int32_t dtype, int32_t memspace,
                               const b200ols_column *coefficients, const b200ols_column *features, int32_t add_intercept,
                               int32_t null_policy, b200ols_output *out) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (!coefficients || !out || !out->values) return fail(B200OLS_ERR_INVALID, "NULL argument");
    if (n_rows < 0 || n_coef < 1 || n_coef > BIG_MAX_F) return fail(B200OLS_ERR_INVALID, "bad shape");
    const int n_feat = n_coef - (add_intercept ? 1 : 0);
    // src/expressions.rs:717-721 "number of coefficients must match number of features!"
    if (n_feat < 0 || (n_feat > 0 && !features)) return fail(B200OLS_ERR_INVALID, "number of coefficients must match number of features!");
    if (dtype != B200OLS_F64 && dtype != B200OLS_F32) return fail(B200OLS_ERR_INVALID, "bad dtype %d", dtype);
    if (null_policy < B200OLS_NULL_IGNORE || null_policy > B200OLS_NULL_DROP_WINDOW) return fail(B200OLS_ERR_INVALID, "Invalid null_policy detected!");
    CU(cudaSetDevice(c->device));
    TRY(free_retired(c));
    const size_t esz = dtype == B200OLS_F64 ? 8 : 4;
    const size_t bm = static_cast<size_t>((n_rows + 7) / 8);
    size_t bytes = 1 << 20;
    if (memspace == B200OLS_HOST) bytes += static_cast<size_t>(n_coef) * 2 * (static_cast<size_t>(n_rows) * 8 + bm + 1024) + static_cast<size_t>(n_rows) * 9 + 4096;
    TRY(arena_reserve(c, bytes));
    c->arena_off = 0;
    c->arena_overflow = false;
    TRY(pinned_begin(c));
    c->last_flags = nullptr;
    auto stage = [&](const b200ols_column &col, size_t es, const void **v, const uint8_t **m) -> int {
        if (memspace == B200OLS_DEVICE) {
            *v = col.values;
            *m = col.validity;
            return 0;
        }
        char *d = arena_alloc<char>(c, static_cast<size_t>(n_rows) * es + 16);
        if (n_rows) {
            const StageSeg sg = {col.values, d, static_cast<size_t>(n_rows) * es};
            TRY(stage_h2d(c, &sg, 1));
        }
        *v = d;
        *m = nullptr;
        if (col.validity) {
            uint8_t *b = arena_alloc<uint8_t>(c, bm + 16);
            CU(cudaMemcpyAsync(b, col.validity, bm, cudaMemcpyHostToDevice, c->stream));
            *m = b;
        }
        return 0;
    };
    double *dout = out->values;
    uint8_t *dval = out->validity;
    if (memspace == B200OLS_HOST) {
        dout = arena_alloc<double>(c, static_cast<size_t>(n_rows));
        dval = out->validity ? arena_alloc<uint8_t>(c, static_cast<size_t>(n_rows)) : nullptr;
    }
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((n_rows + 255) / 256, static_cast<int64_t>(c->sm_count) * 16));
    // the kernel takes up to 64 coefficient columns by value; more are summed batch after batch (same order of
    // additions per row as one pass)
    for (int b0 = 0; b0 < n_coef; b0 += 64) {
        RowDotParams rp;
        std::memset(&rp, 0, sizeof(rp));
        rp.n_coef = std::min(64, n_coef - b0);
        rp.n_feat = std::max(0, std::min(rp.n_coef, n_feat - b0));
        rp.fill_nan = null_policy == B200OLS_NULL_IGNORE;
        rp.drop = null_policy == B200OLS_NULL_DROP;
        rp.accumulate = b0 > 0;
        rp.n_rows = n_rows;
        for (int j = 0; j < rp.n_coef; ++j) {
            const void *v;
            TRY(stage(coefficients[b0 + j], 8, &v, &rp.coef_valid[j]));
            rp.coef[j] = static_cast<const double *>(v);
            if (j < rp.n_feat) TRY(stage(features[b0 + j], esz, &rp.feat[j], &rp.feat_valid[j]));
        }
        rp.out = dout;
        rp.out_valid = dval;
        ARENA_GUARD(c);
        ProfScope prof(c);
        if (dtype == B200OLS_F64) rowdot_kernel<double><<<static_cast<unsigned>(blocks), 256, 0, c->stream>>>(rp);
        else rowdot_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, c->stream>>>(rp);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (memspace == B200OLS_HOST) {
        const StageSeg d2h[2] = {{out->values, dout, sizeof(double) * n_rows}, {out->validity, dval, dval ? static_cast<size_t>(n_rows) : 0}};
        TRY(stage_d2h(c, d2h, 2));  // pageable outputs drain through the pinned ring; returns with the stream synchronised
    }
    if (c->pinned) pinned_end(c);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// moving-window models: rls / rolling (moving.cuh)
// ------------------------------------------------------------------------------------------------
int b200::launch_moving(cudaStream_t stream, MovingParams &p, const int64_t *offsets, bool f64, int sm_count, char *ws, size_t ws_bytes,
                         int64_t *launches) {
    const int64_t G = p.n_groups;
    // three execution paths (moving.cuh): k <= 8 null-free frames stream through per-thread staging rings
    // (moving_fast.cuh); k <= 8 with a row mask / min_periods > window keep the chunk-interleaved copies; 9 <= k <= 64 run
    // one block per chunk (moving_wide.cuh)
    // test hooks, read per call: B200OLS_MOVING_FAST=0 / B200OLS_MOVING_NBR=0 switch the staged / window-chunk kernels off,
    // B200OLS_MOVING_NBR_MIN_CHUNKS lowers the frame size from which the window-chunk kernel is used
    auto env_int = [](const char *name, long long dflt) { const char *v = std::getenv(name); return v ? std::atoll(v) : dflt; };
    const bool fast_enabled = env_int("B200OLS_MOVING_FAST", 1) != 0;
    const bool wide = p.F > 8;
    p.fast = (!wide && fast_enabled && !p.mask && (p.kind == MOVING_RLS || p.min_periods <= p.window)) ? 1 : 0;
    int64_t L;
    p.nbr = 0;
    const bool nbr_enabled = env_int("B200OLS_MOVING_NBR", 1) != 0;
    const int64_t nbr_min_chunks = env_int("B200OLS_MOVING_NBR_MIN_CHUNKS", static_cast<long long>(sm_count) * 128);
    if (p.fast && p.kind == MOVING_ROLLING && nbr_enabled && p.window >= 64 && p.n_rows / p.window >= nbr_min_chunks) {
        // window-length chunks: the row leaving a window is the neighbouring thread's current row (rolling_nbr_kernel)
        p.nbr = 1;
        L = p.window;
    } else if (p.fast) {
        // one chunk per resident thread (a single wave: every thread carries the same work), never below 64 rows
        const int nb = std::max(1, moving_fast_blocks_per_sm(f64, p.F, p.kind, p.kd + 1 + (p.w ? 1 : 0)));
        const int64_t resident = static_cast<int64_t>(sm_count) * nb * 128;
        L = std::max<int64_t>(64, (p.n_rows + resident - 1) / resident);
        L = (L + 7) & ~static_cast<int64_t>(7);
    } else if (wide) {
        L = 64;
        const int64_t want = p.n_rows / (static_cast<int64_t>(sm_count) * 16);
        while (L < want) L <<= 1;
    } else {
        L = moving_chunk_len(p.n_rows, sm_count, p.kind, p.window);
        // the chunk-interleaved copies take L * n_chunks elements per column: with many series much shorter than
        // L that pads badly, so shrink L until the padded size is at most 2N + 64G (always true at L = 64)
        for (;;) {
            int64_t nch = 0;
            for (int64_t g = 0; g < G; ++g) nch += (offsets[g + 1] - offsets[g] + L - 1) / L;
            if (L <= 64 || L * nch <= 2 * p.n_rows + 64 * G) break;
            L >>= 1;
        }
    }
    std::vector<int64_t> r0, r1, gco(static_cast<size_t>(G) + 1);
    std::vector<int32_t> cg;
    for (int64_t g = 0; g < G; ++g) {
        gco[g] = static_cast<int64_t>(r0.size());
        for (int64_t a = offsets[g]; a < offsets[g + 1]; a += L) {
            r0.push_back(a);
            r1.push_back(std::min(a + L, offsets[g + 1]));
            cg.push_back(static_cast<int32_t>(g));
        }
    }
    gco[G] = static_cast<int64_t>(r0.size());
    const size_t nc = r0.size();
    // super-chunks for the rls scan: runs of <= SCAN_SUPER chunks inside one series
    std::vector<int64_t> s0, s1, gso(static_cast<size_t>(G) + 1);
    if (p.kind == MOVING_RLS) {
        for (int64_t g = 0; g < G; ++g) {
            gso[g] = static_cast<int64_t>(s0.size());
            for (int64_t c = gco[g]; c < gco[g + 1]; c += SCAN_SUPER) {
                s0.push_back(c);
                s1.push_back(std::min<int64_t>(c + SCAN_SUPER, gco[g + 1]));
            }
        }
        gso[G] = static_cast<int64_t>(s0.size());
    }
    const size_t ns = s0.size();
    p.n_chunks = static_cast<int64_t>(nc);
    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = ws + off; off += (bytes + 255) & ~static_cast<size_t>(255); return q; };
    int64_t *d_r0 = reinterpret_cast<int64_t *>(take(nc * 8 + 8));
    int64_t *d_r1 = reinterpret_cast<int64_t *>(take(nc * 8 + 8));
    int32_t *d_cg = reinterpret_cast<int32_t *>(take(nc * 4 + 8));
    int64_t *d_gco = reinterpret_cast<int64_t *>(take((G + 1) * 8));
    p.series_info = reinterpret_cast<int64_t *>(take(static_cast<size_t>(G) * 32 + 8));
    p.summaries = reinterpret_cast<double *>(take((p.kind == MOVING_RLS || p.nbr) ? nc * moving_rec(p.F) * 8 + 8 : 8));
    int64_t *d_s0 = reinterpret_cast<int64_t *>(take(ns * 8 + 8));
    int64_t *d_s1 = reinterpret_cast<int64_t *>(take(ns * 8 + 8));
    int64_t *d_gso = reinterpret_cast<int64_t *>(take((G + 1) * 8));
    p.sup = reinterpret_cast<double *>(take(ns * moving_rec(p.F) * 8 + 8));
    p.n_super = static_cast<int64_t>(ns);
    if (!p.fast && !wide) {   // chunk-interleaved column copies (filled by chunk_transpose_kernel)
        p.chunk_len = L;
        p.chunk_shift = 0;
        while ((int64_t{1} << p.chunk_shift) < L) ++p.chunk_shift;
        const size_t esz = f64 ? 8 : 4;
        const size_t col_bytes = static_cast<size_t>(L) * nc * esz + 256;
        for (int j = 0; j <= p.kd; ++j) p.tcols[j] = take(col_bytes);
        p.tw = p.w ? take(col_bytes) : nullptr;
        p.tmask = p.mask ? take(col_bytes) : nullptr;
    }
    p.sup_c0 = d_s0;
    p.sup_c1 = d_s1;
    p.group_sup_off = d_gso;
    if (off > ws_bytes) return fail(B200OLS_ERR_CUDA, "internal: moving workspace under-sized (%zu > %zu)", off, ws_bytes);
    // pageable -> device async copies are staged by the driver before the call returns
    if (nc) {
        CU(cudaMemcpyAsync(d_r0, r0.data(), nc * 8, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(d_r1, r1.data(), nc * 8, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(d_cg, cg.data(), nc * 4, cudaMemcpyHostToDevice, stream));
    }
    CU(cudaMemcpyAsync(d_gco, gco.data(), (G + 1) * 8, cudaMemcpyHostToDevice, stream));
    if (ns) {
        CU(cudaMemcpyAsync(d_s0, s0.data(), ns * 8, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(d_s1, s1.data(), ns * 8, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(d_gso, gso.data(), (G + 1) * 8, cudaMemcpyHostToDevice, stream));
    }
    CU(cudaStreamSynchronize(stream));  // host vectors go out of scope; tables are tiny
    p.chunk_r0 = d_r0;
    p.chunk_r1 = d_r1;
    p.chunk_group = d_cg;
    CU(f64 ? moving_launch_f64(stream, p, d_gco, launches) : moving_launch_f32(stream, p, d_gco, launches));
    return 0;
}

static int run_moving_impl(b200ols_ctx *c, const b200ols_frame *f, int kind, const b200ols_rls_kwargs *rk,
                           const b200ols_rolling_kwargs *wk, int mode, b200ols_output *out, double *state_host = nullptr) {
    if (!c) return fail(B200OLS_ERR_INVALID, "ctx is NULL");
    if (!state_host && (!out || !out->values)) return fail(B200OLS_ERR_INVALID, "NULL argument");
    TRY(validate_frame(f));
    const int policy = kind == MOVING_RLS ? rk->null_policy : wk->null_policy;
    if (policy < B200OLS_NULL_IGNORE || policy > B200OLS_NULL_DROP_WINDOW) return fail(B200OLS_ERR_INVALID, "Invalid null_policy detected!");
    const int F = f->n_features + (f->add_intercept ? 1 : 0);
    if (F > MOVING_MAX_K)
        return fail(B200OLS_ERR_UNSUPPORTED, "rls / rolling_ols with more than %d coefficients (%d) is not implemented on the device", MOVING_MAX_K, F);
    CU(cudaSetDevice(c->device));
    TRY(free_retired(c));
    const int64_t N = f->n_rows, G = f->n_groups;

    MovingParams mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.kind = kind;
    mp.F = F;
    if (kind == MOVING_RLS) {
        const double hl = rk->half_life;
        mp.lambda = (!std::isnan(hl) && hl > 0.0) ? std::exp(std::log(0.5) / hl) : 1.0;  // src/least_squares.rs:513-517
        // Some(half_life <= 0) is not rejected by the reference either: exp(ln(.5)/hl)
        if (!std::isnan(hl) && hl <= 0.0) mp.lambda = std::exp(std::log(0.5) / hl);
        mp.p0 = std::isnan(rk->initial_state_covariance) ? 10.0 : rk->initial_state_covariance;
        mp.has_mean = (rk->initial_state_mean != nullptr && mode == B200OLS_COEFFICIENTS) ? 1 : 0;  // quirk A.5.1
        if (mp.has_mean)
            for (int j = 0; j < F; ++j) mp.mean[j] = rk->initial_state_mean[j];
    } else {
        if (wk->window_size < 1) return fail(B200OLS_ERR_INVALID, "window_size must be >= 1");
        mp.window = wk->window_size;
        mp.min_periods = wk->min_periods < 0 ? std::min<int64_t>(F, wk->window_size) : wk->min_periods;  // :860
        if (mp.min_periods < 1) return fail(B200OLS_ERR_INVALID, "min_periods must be >= 1");
        mp.alpha = std::isnan(wk->alpha) ? 0.0 : wk->alpha;
        mp.fixed_window = !(policy == B200OLS_NULL_DROP || policy == B200OLS_NULL_DROP_ZERO || policy == B200OLS_NULL_DROP_Y_ZERO_X);  // :947-950
    }

    // the chunk-interleaved copies (2 x the frame) are only reserved when this call can take the kernels that read them
    bool any_bitmap = f->target.validity != nullptr || (f->sample_weights && f->sample_weights->validity);
    for (int j = 0; j < f->n_features; ++j) any_bitmap = any_bitmap || f->features[j].validity != nullptr;
    const bool may_mask = any_bitmap && policy != B200OLS_NULL_ZERO && policy != B200OLS_NULL_IGNORE;
    const bool fast_off = [] { const char *v = std::getenv("B200OLS_MOVING_FAST"); return v && std::atoi(v) == 0; }();
    const bool transposed = F <= 8 && (may_mask || fast_off || (kind == MOVING_ROLLING && mp.min_periods > mp.window));
    const size_t ws_bytes = moving_workspace_bytes(N, G, F, transposed);
    size_t bytes = stage_bytes_bound(f) + (static_cast<size_t>(G) + 8) * 64 + ws_bytes + (1 << 20);
    if (f->memspace == B200OLS_HOST) bytes += static_cast<size_t>(N) * (static_cast<size_t>(F) * 9 + 16) + 4096;
    TRY(arena_reserve(c, bytes));
    c->arena_off = 0;
    c->arena_overflow = false;
    TRY(pinned_begin(c));
    c->last_flags = nullptr;

    Staged st;
    TRY(stage_frame(c, f, policy, true, &st));
    // group offsets on the device
    {
        const size_t ob = sizeof(int64_t) * (G + 1);
        TRY(pinned_reserve(c, c->pinned_off + ob + 256));
        char *h = c->pinned + c->pinned_off;
        c->pinned_off += round_up(ob, 256);
        std::memcpy(h, st.offsets.data(), ob);
        int64_t *d = arena_alloc<int64_t>(c, static_cast<size_t>(G) + 1);
        CU(cudaMemcpyAsync(d, h, ob, cudaMemcpyHostToDevice, c->stream));
        mp.group_off = d;
    }
    for (int j = 0; j < st.kd; ++j) mp.cols[j] = st.feat[j];
    mp.cols[st.kd] = st.y;
    mp.w = st.w;
    mp.w_is_sqrt = st.w_is_sqrt;
    mp.mask = st.mask;
    mp.kd = st.kd;
    mp.intercept = st.intercept;
    mp.n_groups = G;
    mp.n_rows = N;
    mp.row_index = st.row_index;
    mp.mode = mode;
    mp.target = st.y_raw;
    mp.target_is_packed = st.y_raw_packed;
    mp.target_validity = st.y_validity;
    mp.mask_predictions = (st.mask != nullptr) ? 1 : 0;  // src/expressions.rs:640-645,695-700
    const size_t NE = static_cast<size_t>(F) * F + F;
    if (kind == MOVING_RLS && rk->initial_information) {  // series continued from an earlier time shard
        const size_t ib = sizeof(double) * NE * static_cast<size_t>(G);
        TRY(pinned_reserve(c, c->pinned_off + ib + 256));
        char *h = c->pinned + c->pinned_off;
        c->pinned_off += round_up(ib, 256);
        std::memcpy(h, rk->initial_information, ib);
        double *d = arena_alloc<double>(c, NE * static_cast<size_t>(G));
        CU(cudaMemcpyAsync(d, h, ib, cudaMemcpyHostToDevice, c->stream));
        mp.init_info = d;
    }
    if (state_host) {
        mp.state_out = arena_alloc<double>(c, (NE + 1) * static_cast<size_t>(G));
        mp.state_only = 1;
    }
    const size_t out_elems = state_host ? 0 : (mode == B200OLS_COEFFICIENTS ? static_cast<size_t>(N) * F : static_cast<size_t>(N));
    double *dout = state_host ? nullptr : out->values;
    uint8_t *dval = state_host ? nullptr : out->validity;
    if (!state_host && f->memspace == B200OLS_HOST) {
        dout = arena_alloc<double>(c, out_elems);
        dval = out->validity ? arena_alloc<uint8_t>(c, out_elems) : nullptr;
    }
    mp.out = dout;
    mp.out_valid = dval;
    char *ws = arena_alloc<char>(c, ws_bytes);
    ARENA_GUARD(c);
    {
        ProfScope prof(c);
        TRY(launch_moving(c->stream, mp, st.offsets.data(), f->dtype == B200OLS_F64, c->sm_count, ws, ws_bytes, &c->launches));
    }
    if (state_host) {
        // state leaving each series: dense [A (F x F, symmetric), b (F), D]
        if (N == 0) {  // nothing launched: the state is the entering one
            for (int64_t g = 0; g < G; ++g) {
                double *o = state_host + g * (NE + 1);
                for (size_t e = 0; e < NE; ++e) {
                    if (rk->initial_information) o[e] = rk->initial_information[g * NE + e];
                    else if (e < static_cast<size_t>(F) * F) o[e] = (e / F == e % F) ? 1.0 / mp.p0 : 0.0;
                    else o[e] = (mp.has_mean ? mp.mean[e - static_cast<size_t>(F) * F] : 0.0) / mp.p0;
                }
                o[NE] = 1.0;
            }
            return 0;
        }
        CU(cudaMemcpyAsync(state_host, mp.state_out, sizeof(double) * (NE + 1) * G, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (int64_t g = 0; g < G; ++g) {  // the device carries the lower triangle
            double *A = state_host + g * (NE + 1);
            for (int i = 0; i < F; ++i)
                for (int j = i + 1; j < F; ++j) A[i * F + j] = A[j * F + i];
        }
        return 0;
    }
    if (f->memspace == B200OLS_HOST) {
        const StageSeg d2h[2] = {{out->values, dout, sizeof(double) * out_elems}, {out->validity, dval, dval ? out_elems : 0}};
        TRY(stage_d2h(c, d2h, 2));  // pageable outputs drain through the pinned ring; returns with the stream synchronised
    }
    return 0;
}

static int run_moving(b200ols_ctx *c, const b200ols_frame *f, int kind, const b200ols_rls_kwargs *rk,
                      const b200ols_rolling_kwargs *wk, int mode, b200ols_output *out, double *state_host = nullptr) {
    const int rc = run_moving_impl(c, f, kind, rk, wk, mode, out, state_host);
    if (c && c->pinned) pinned_end(c);
    return rc;
}

extern "C" int b200ols_recursive_least_squares(b200ols_ctx *c, const b200ols_frame *f, const b200ols_rls_kwargs *kw, int mode,
                                               b200ols_output *out) {
    if (!kw) return fail(B200OLS_ERR_INVALID, "kwargs is NULL");
    if (mode != B200OLS_PREDICTIONS && mode != B200OLS_RESIDUALS) return fail(B200OLS_ERR_INVALID, "bad mode %d", mode);
    return run_moving(c, f, MOVING_RLS, kw, nullptr, mode, out);
}
extern "C" int b200ols_recursive_least_squares_coefficients(b200ols_ctx *c, const b200ols_frame *f,
                                                            const b200ols_rls_kwargs *kw, b200ols_output *out) {
    if (!kw) return fail(B200OLS_ERR_INVALID, "kwargs is NULL");
    return run_moving(c, f, MOVING_RLS, kw, nullptr, B200OLS_COEFFICIENTS, out);
}
extern "C" int b200ols_recursive_least_squares_state(b200ols_ctx *c, const b200ols_frame *f, const b200ols_rls_kwargs *kw,
                                                     double *state) {
    if (!kw) return fail(B200OLS_ERR_INVALID, "kwargs is NULL");
    if (!state) return fail(B200OLS_ERR_INVALID, "NULL argument");
    // the prior mean belongs to the state whatever the mode (the caller composes states, not predictions)
    return run_moving(c, f, MOVING_RLS, kw, nullptr, B200OLS_COEFFICIENTS, nullptr, state);
}
extern "C" int b200ols_rolling_least_squares(b200ols_ctx *c, const b200ols_frame *f, const b200ols_rolling_kwargs *kw, int mode,
                                             b200ols_output *out) {
    if (!kw) return fail(B200OLS_ERR_INVALID, "kwargs is NULL");
    if (mode != B200OLS_PREDICTIONS && mode != B200OLS_RESIDUALS) return fail(B200OLS_ERR_INVALID, "bad mode %d", mode);
    return run_moving(c, f, MOVING_ROLLING, nullptr, kw, mode, out);
}
extern "C" int b200ols_rolling_least_squares_coefficients(b200ols_ctx *c, const b200ols_frame *f,
                                                          const b200ols_rolling_kwargs *kw, b200ols_output *out) {
    if (!kw) return fail(B200OLS_ERR_INVALID, "kwargs is NULL");
    return run_moving(c, f, MOVING_ROLLING, nullptr, kw, B200OLS_COEFFICIENTS, out);
}
