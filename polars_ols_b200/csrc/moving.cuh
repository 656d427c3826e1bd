// moving.cuh — kernels + launcher for recursive_least_squares{,_coefficients} and
// rolling_least_squares{,_coefficients} (src/expressions.rs:594-701 of /root/reference) over the
// chunk algorithms of moving_core.cuh.  One thread per (series, time chunk); all series (groups of the
// `.over()` context) and all chunks of one frame run in a single launch per pass:
//   rolling : [prepass per series (only with a row mask)] -> rolling_main
//   rls     : rls_summary -> rls_scan (per series, lanes = matrix elements) -> rls_main
// Outputs are scattered to the original row order; predictions are (X o Theta).sum(1) with the WLS
// un-scaling and the null masks of src/expressions.rs:640-645,695-700 and
// polars_ols/least_squares.py:234-239,407-408 fused in.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "gram_stream.cuh"
#include "moving_core.cuh"

namespace b200 {

// The chunk walks are latency-bound (one thread per chunk, a dependent FP64 chain and a memory round trip per row):
// throughput follows the number of resident warps, but only while the state stays in registers.  Measured on C4
// (rolling, k = 6, 50M rows): 162 registers / 12 warps per SM 4.87 ms; forced to 128 registers / 16 warps
// (MOVING_MIN_BLOCKS = 4, part of the Cholesky factor in local memory) 5.75 ms; 204 registers / 8 warps with the next
// row's loads software-pipelined 6.7 ms.  So the default stays at one block minimum (no register cap).
#ifndef MOVING_MIN_BLOCKS
#define MOVING_MIN_BLOCKS 1
#endif
constexpr int MOVING_OCC_K = 6;

struct MovingParams {
    const void *cols[GRAM_MAX_COLS];  // [0,kd) features, [kd] target (cleaned: zero filled)
    const void *w;                    // weights / sqrt-weights or nullptr
    const void *mask;                 // T-typed row validity or nullptr (all valid)
    int kd, intercept, F, w_is_sqrt;
    int kind;                         // MOVING_RLS | MOVING_ROLLING
    int mode;                         // B200OLS_PREDICTIONS(0) | RESIDUALS(1) | COEFFICIENTS(2)
    int mask_predictions;             // policy has a validity mask: invalid rows -> null predictions
    int64_t n_groups, n_rows;
    const int64_t *group_off;         // device [G+1]
    const int64_t *row_index;         // packed -> original or nullptr
    const void *target;               // raw target (residuals)
    int target_is_packed;
    const uint8_t *target_validity;   // bitmap over original rows or nullptr
    double *out;
    uint8_t *out_valid;
    // rls
    double lambda, p0;
    int has_mean;
    double mean[MOVING_MAX_K];
    // rolling
    int64_t window, min_periods;
    double alpha;
    int fixed_window;
    // chunk table (device) + workspace
    int64_t n_chunks;
    const int64_t *chunk_r0, *chunk_r1;
    const int32_t *chunk_group;
    double *summaries;                // [n_chunks][REC]  rls
    // chunk-interleaved ("transposed") copies of the input columns: element (column, row i of chunk ch) at
    // tcols[column][i * n_chunks + ch], so that the lanes of a warp (consecutive chunks) read consecutive
    // addresses while every thread walks its own chunk sequentially
    const void *tcols[GRAM_MAX_COLS];  // [0,kd) features, [kd] target
    const void *tw, *tmask;
    int64_t chunk_len;                // L (power of two); every chunk of a series except its last has L rows
    int chunk_shift;                  // log2(L)
    int64_t n_super;                  // rls scan: runs of <= SCAN_SUPER chunks of one series
    const int64_t *sup_c0, *sup_c1, *group_sup_off;
    double *sup;                      // [n_super][REC]
    int64_t *series_info;             // [G][4] rolling: mpv, n_valid, all_nan, m_warm
    // rls continued from / summarised for another time shard (single long series):
    const double *init_info;          // [G][F*F+F] information state ENTERING each series (A row-major, b) or nullptr = prior
    double *state_out;                // [G][F*F+F+1] state LEAVING each series (lower triangle of A, b) and the decay D, or nullptr
    int state_only;                   // stop after the scan (b200ols_recursive_least_squares_state)
    int fast;                         // staged thread-private rings (moving_fast.cuh): no mask, k <= 8, no transposed copies
    int nbr;                          // rolling, fast: chunks of exactly `window` rows, lag rows from the neighbour's slot
};

__host__ __device__ constexpr int moving_rec(int K) { return K * K + K + 1; }  // doubles per chunk record: A, b, D

template <typename T, int K>
struct DevSrc {
    const T *x[K];
    const T *y, *w, *mask;
    int kd, w_is_sqrt;
    __device__ __forceinline__ bool valid(int64_t r) const { return mask ? (mask[r] != T(0)) : true; }
    __device__ __forceinline__ T scale(int64_t r) const {
        if (!w) return T(1);
        const T v = w[r];
        return w_is_sqrt ? v : static_cast<T>(sqrt(v));
    }
    __device__ __forceinline__ void load(int64_t r, double (&xo)[K], double &yo) const {
        const T s = scale(r);
#pragma unroll
        for (int j = 0; j < K; ++j) xo[j] = (j < kd) ? static_cast<double>(static_cast<T>(x[j][r] * s)) : static_cast<double>(s);
        yo = static_cast<double>(static_cast<T>(y[r] * s));
    }
    __device__ __forceinline__ void prefetch(int64_t) const {}
};

template <typename T, int K>
__device__ __forceinline__ DevSrc<T, K> make_src(const MovingParams &p) {
    DevSrc<T, K> s;
#pragma unroll
    for (int j = 0; j < K; ++j) s.x[j] = (j < p.kd) ? static_cast<const T *>(p.cols[j]) : nullptr;
    s.y = static_cast<const T *>(p.cols[p.kd]);
    s.w = static_cast<const T *>(p.w);
    s.mask = static_cast<const T *>(p.mask);
    s.kd = p.kd;
    s.w_is_sqrt = p.w_is_sqrt;
    return s;
}

// Same interface over the chunk-interleaved copies; r is still the packed row index.  Rows of earlier
// chunks of the same series (window halo, warm-up) resolve to (chunk - back, row inside that chunk).
template <typename T, int K>
struct DevSrcT {
    const T *x[K];
    const T *y, *w, *mask;
    int64_t r0, c, nch, L;
    int shift, kd, w_is_sqrt;
    __device__ __forceinline__ int64_t at(int64_t r) const {
        if (r >= r0) return (r - r0) * nch + c;
        const int64_t d = r0 - r;
        const int64_t back = (d + L - 1) >> shift;
        return ((back << shift) - d) * nch + (c - back);
    }
    __device__ __forceinline__ bool valid(int64_t r) const { return mask ? (mask[at(r)] != T(0)) : true; }
    __device__ __forceinline__ T scale_at(int64_t t) const {
        if (!w) return T(1);
        const T v = w[t];
        return w_is_sqrt ? v : static_cast<T>(sqrt(v));
    }
    __device__ __forceinline__ T scale(int64_t r) const { return scale_at(at(r)); }
    __device__ __forceinline__ void load(int64_t r, double (&xo)[K], double &yo) const {
        const int64_t t = at(r);
        const T s = scale_at(t);
#pragma unroll
        for (int j = 0; j < K; ++j) xo[j] = (j < kd) ? static_cast<double>(static_cast<T>(x[j][t] * s)) : static_cast<double>(s);
        yo = static_cast<double>(static_cast<T>(y[t] * s));
    }
    // L1 prefetch hint for every column of row r (the chunk loops call it MOVING_PF rows ahead: a thread walks its
    // chunk row by row and would otherwise pay one HBM / L2 round trip per row — ncu: long_scoreboard 5.2 of 6.2
    // stalled warps per issue slot at 12 warps per SM)
    __device__ __forceinline__ void prefetch(int64_t r) const {
        const int64_t t = at(r);
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j < kd) asm volatile("prefetch.global.L1 [%0];" ::"l"(x[j] + t));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(y + t));
        if (w) asm volatile("prefetch.global.L1 [%0];" ::"l"(w + t));
        if (mask) asm volatile("prefetch.global.L1 [%0];" ::"l"(mask + t));
    }
};

template <typename T, int K>
__device__ __forceinline__ DevSrcT<T, K> make_src_t(const MovingParams &p, int64_t chunk) {
    DevSrcT<T, K> s;
#pragma unroll
    for (int j = 0; j < K; ++j) s.x[j] = (j < p.kd) ? static_cast<const T *>(p.tcols[j]) : nullptr;
    s.y = static_cast<const T *>(p.tcols[p.kd]);
    s.w = static_cast<const T *>(p.tw);
    s.mask = static_cast<const T *>(p.tmask);
    s.r0 = p.chunk_r0[chunk];
    s.c = chunk;
    s.nch = p.n_chunks;
    s.L = p.chunk_len;
    s.shift = p.chunk_shift;
    s.kd = p.kd;
    s.w_is_sqrt = p.w_is_sqrt;
    return s;
}

// dst[i * n_chunks + ch] = src[chunk_r0[ch] + i]: 32 x 32 tiles through shared memory, coalesced on both sides.
// blockIdx.z = column slot (features, target, weights, mask).
struct TransposeParams {
    const void *src[GRAM_MAX_COLS];
    void *dst[GRAM_MAX_COLS];
    const int64_t *chunk_r0, *chunk_r1;
    int64_t n_chunks, chunk_len;
};

template <typename T>
__global__ void __launch_bounds__(256) chunk_transpose_kernel(const TransposeParams p) {
    __shared__ T tile[32][33];
    const T *src = static_cast<const T *>(p.src[blockIdx.z]);
    T *dst = static_cast<T *>(p.dst[blockIdx.z]);
    const int64_t ch0 = static_cast<int64_t>(blockIdx.x) * 32, i0 = static_cast<int64_t>(blockIdx.y) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
    for (int k = ty; k < 32; k += 8) {  // chunk ch0 + k, rows i0 + tx
        const int64_t ch = ch0 + k;
        T v = T(0);
        if (ch < p.n_chunks) {
            const int64_t r = p.chunk_r0[ch] + i0 + tx;
            if (r < p.chunk_r1[ch]) v = src[r];
        }
        tile[k][tx] = v;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {  // row i0 + k, chunks ch0 + tx
        const int64_t ch = ch0 + tx, i = i0 + k;
        if (ch < p.n_chunks && i < p.chunk_len) dst[i * p.n_chunks + ch] = tile[tx][k];
    }
}

template <typename T, int K, typename SrcT>
struct DevEmit {
    const MovingParams &p;
    const SrcT &src;
    __device__ __forceinline__ void operator()(int64_t r, const double (&beta)[K], bool) const {
        if (p.mode == 2) {
            coefficients(r, beta);
            return;
        }
        double x[K], y;
        src.load(r, x, y);
        row(r, beta, x);
    }
    __device__ __forceinline__ void coefficients(int64_t r, const double (&beta)[K]) const {
        const int64_t orow = p.row_index ? p.row_index[r] : r;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            p.out[orow * K + j] = beta[j];
            if (p.out_valid) p.out_valid[orow * K + j] = (beta[j] == beta[j]) ? 1 : 0;
        }
    }
    // x = the row's (scaled) features as src.load returns them — the pipelined chunk loops already hold them
    __device__ __forceinline__ void row(int64_t r, const double (&beta)[K], const double (&x)[K]) const {
        if (p.mode == 2) {
            coefficients(r, beta);
            return;
        }
        const int64_t orow = p.row_index ? p.row_index[r] : r;
        double pred = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) pred += x[j] * beta[j];  // (features * coefficients).sum_axis(1)
        if (p.w) pred *= static_cast<double>(T(1) / src.scale(r));
        bool valid = true;
        if (p.mask_predictions) valid = src.valid(r);
        if (p.mode == 1) {
            const int64_t trow = p.target_is_packed ? r : orow;
            pred = static_cast<double>(static_cast<const T *>(p.target)[trow]) - pred;
            if (p.target_validity) valid = valid && ((p.target_validity[orow >> 3] >> (orow & 7)) & 1);
        }
        if (p.kind == MOVING_ROLLING) valid = valid && (pred == pred);  // fill_nan(None)
        p.out[orow] = pred;
        if (p.out_valid) p.out_valid[orow] = valid ? 1 : 0;
    }
};

template <typename T, int K>
__global__ void __launch_bounds__(128) rolling_prepass_kernel(const MovingParams p) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= p.n_groups) return;
    const DevSrc<T, K> src = make_src<T, K>(p);
    const RollingSeries rs = rolling_prepass(src, p.group_off[g], p.group_off[g + 1], p.min_periods);
    p.series_info[g * 4 + 0] = rs.mpv;
    p.series_info[g * 4 + 1] = rs.n_valid;
    p.series_info[g * 4 + 2] = rs.all_nan;
    p.series_info[g * 4 + 3] = rs.m_warm;
}

template <typename T, int K>
__global__ void __launch_bounds__(128, (K <= MOVING_OCC_K ? MOVING_MIN_BLOCKS : 1)) rolling_main_kernel(const MovingParams p) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= p.n_chunks) return;
    const DevSrcT<T, K> src = make_src_t<T, K>(p, c);
    const int64_t g = p.chunk_group[c];
    const int64_t g0 = p.group_off[g], g1 = p.group_off[g + 1];
    RollingSeries rs;
    if (p.series_info) {
        rs.mpv = p.series_info[g * 4 + 0];
        rs.n_valid = p.series_info[g * 4 + 1];
        rs.all_nan = static_cast<int>(p.series_info[g * 4 + 2]);
        rs.m_warm = p.series_info[g * 4 + 3];
    } else {  // no mask: every row is valid
        rs.mpv = p.min_periods;
        rs.n_valid = (g1 - g0 < p.min_periods) ? (g1 - g0) : p.min_periods;
        rs.all_nan = (g1 - g0) < p.min_periods;
        rs.m_warm = p.min_periods;
    }
    RollingCfg cfg{p.window, p.min_periods, p.alpha, p.fixed_window};
    DevEmit<T, K, DevSrcT<T, K>> emit{p, src};
    rolling_chunk<K>(src, cfg, rs, g0, g1, p.chunk_r0[c], p.chunk_r1[c], emit);
}

template <typename T, int K>
__global__ void __launch_bounds__(128) rls_summary_kernel(const MovingParams p) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= p.n_chunks) return;
    const DevSrcT<T, K> src = make_src_t<T, K>(p, c);
    RlsCfg cfg{p.lambda, p.p0};
    RlsSummary<K> s;
    rls_summarise<K>(src, cfg, p.chunk_r0[c], p.chunk_r1[c], s);
    double *rec = p.summaries + c * moving_rec(K);
#pragma unroll
    for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int j = 0; j < K; ++j) rec[i * K + j] = (j <= i) ? s.ab.S[i][j] : 0.0;
        rec[K * K + i] = s.ab.v[i];
    }
    rec[K * K + K] = s.D;
}

// Exclusive scan over the chunks of each series: rec[c] <- information state ENTERING chunk c.
// The per-chunk summaries are affine maps  s -> D s + A  (associative), so the scan is hierarchical:
//   phase 0: one warp per super-chunk (run of <= 256 chunks of one series) composes its chunks
//            (lanes = the K*K+K matrix elements) -> sup[sc] = (D, A) of the run;
//   phase 1: one warp per series walks its super-chunks: sup[sc] <- state entering the run, starting from
//            the prior A0 = I/p0, b0 = A0 theta0;
//   phase 2: one warp per super-chunk replays its chunks from that state and overwrites rec[c] in place.
// Records are prefetched SCAN_PF at a time (the addresses do not depend on the carry), so no phase is a
// chain of dependent HBM round trips.
constexpr int SCAN_SUPER = 256;
constexpr int SCAN_PF = 8;

template <int K>
__global__ void __launch_bounds__(128) rls_scan_kernel(const MovingParams p, int phase, int64_t n_super,
                                                       const int64_t *__restrict__ sup_c0, const int64_t *__restrict__ sup_c1,
                                                       const int64_t *__restrict__ group_sup_off, double *__restrict__ sup) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    constexpr int NE = K * K + K;
    constexpr int NT = (NE + 31) / 32;
    double carry[NT];
    if (phase == 1) {
        if (wid >= p.n_groups) return;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int e = lane + 32 * t;
            double v = 0.0;
            if (p.init_info) {                                                            // continued series
                if (e < K * K) v = ((e % K) <= (e / K)) ? p.init_info[wid * NE + e] : 0.0;  // lower triangle only
                else if (e < NE) v = p.init_info[wid * NE + e];
            } else if (e < K * K) v = ((e / K) == (e % K)) ? 1.0 / p.p0 : 0.0;           // A0 = I / p0
            else if (e < NE) v = (p.has_mean ? p.mean[e - K * K] : 0.0) / p.p0;         // b0 = A0 theta0
            carry[t] = v;
        }
        double Dall = 1.0;
        for (int64_t sc = group_sup_off[wid]; sc < group_sup_off[wid + 1]; ++sc) {
            double *rec = sup + sc * moving_rec(K);
            const double D = rec[NE];
            Dall *= D;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int e = lane + 32 * t;
                if (e < NE) {
                    const double add = rec[e];
                    rec[e] = carry[t];
                    carry[t] = fma(D, carry[t], add);
                }
            }
            __syncwarp();
        }
        if (p.state_out) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int e = lane + 32 * t;
                if (e < NE) p.state_out[wid * (NE + 1) + e] = carry[t];
            }
            if (lane == 0) p.state_out[wid * (NE + 1) + NE] = Dall;
        }
        return;
    }
    if (wid >= n_super) return;
    const int64_t c0 = sup_c0[wid], c1 = sup_c1[wid];
    double Dtot = 1.0;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int e = lane + 32 * t;
        carry[t] = (phase == 2 && e < NE) ? sup[wid * moving_rec(K) + e] : 0.0;  // phase 0 starts from the zero map offset
    }
    for (int64_t cb = c0; cb < c1; cb += SCAN_PF) {
        double add[SCAN_PF][NT], Dv[SCAN_PF];
#pragma unroll
        for (int u = 0; u < SCAN_PF; ++u) {
            const int64_t c = (cb + u < c1) ? cb + u : c1 - 1;
            const double *rec = p.summaries + c * moving_rec(K);
            Dv[u] = rec[NE];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int e = lane + 32 * t;
                add[u][t] = (e < NE) ? rec[e] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < SCAN_PF; ++u) {
            if (cb + u < c1) {
                double *rec = p.summaries + (cb + u) * moving_rec(K);
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const int e = lane + 32 * t;
                    if (phase == 2 && e < NE) rec[e] = carry[t];
                    carry[t] = fma(Dv[u], carry[t], add[u][t]);
                }
                Dtot *= Dv[u];
            }
        }
    }
    if (phase == 0) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int e = lane + 32 * t;
            if (e < NE) sup[wid * moving_rec(K) + e] = carry[t];
        }
        if (lane == 0) sup[wid * moving_rec(K) + NE] = Dtot;
    }
}

template <typename T, int K>
__global__ void __launch_bounds__(128, (K <= MOVING_OCC_K ? MOVING_MIN_BLOCKS : 1)) rls_main_kernel(const MovingParams p) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= p.n_chunks) return;
    const DevSrcT<T, K> src = make_src_t<T, K>(p, c);
    const int64_t g = p.chunk_group[c];
    const int64_t r0 = p.chunk_r0[c];
    const bool first = (r0 == p.group_off[g]) && !p.init_info;  // a continued series enters through its information state
    RlsCfg cfg{p.lambda, p.p0};
    NormalState<K> in;
    if (!first) {
        const double *rec = p.summaries + c * moving_rec(K);
#pragma unroll
        for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int j = 0; j < K; ++j) in.S[i][j] = rec[i * K + j];
            in.v[i] = rec[K * K + i];
        }
    }
    DevEmit<T, K, DevSrcT<T, K>> emit{p, src};
    rls_chunk<K>(src, cfg, first, p.has_mean ? p.mean : nullptr, &in, r0, p.chunk_r1[c], emit);
}

// ---- host side -------------------------------------------------------------------------------------
inline int64_t moving_chunk_len(int64_t n_rows, int sm_count, int kind, int64_t window) {
    int64_t want = (n_rows + static_cast<int64_t>(sm_count) * 1024 - 1) / (static_cast<int64_t>(sm_count) * 1024);
    if (kind == MOVING_ROLLING) want = std::max<int64_t>(want, std::min<int64_t>(window, 1 << 16) / 4);
    int64_t L = 64;  // power of two (the chunk-interleaved addressing shifts instead of dividing)
    while (L < want) L <<= 1;
    return L;
}

// `transposed`: the call may take the round-1 kernels that read chunk-interleaved column copies (k <= 8 with a row mask or
// min_periods > window); the staged and block-per-chunk kernels read the caller's columns in place
inline size_t moving_workspace_bytes(int64_t n_rows, int64_t n_groups, int F, bool transposed = true) {
    const size_t max_chunks = static_cast<size_t>(n_rows / 64 + n_groups + 2);
    const size_t max_super = max_chunks / 256 + static_cast<size_t>(n_groups) + 2;
    size_t b = max_chunks * (moving_rec(F) * 8 + 8 + 8 + 4) + static_cast<size_t>(n_groups + 2) * (4 * 8 + 8 + 8) +
               max_super * (moving_rec(F) * 8 + 16) + 16384;
    // chunk-interleaved column copies: (F + 3) columns x padded rows (every series pads < one chunk)
    if (transposed) b += static_cast<size_t>(F + 3) * (static_cast<size_t>(n_rows) * 2 + static_cast<size_t>(n_groups + 1) * 64 + 4096) * 8;
    return b;
}

// thread-per-chunk kernels for 1..8 coefficients (moving_f{64,32}_r0.cu); 9..MOVING_WIDE_MAX_K coefficients run the
// block-per-chunk kernels of moving_wide.cuh (moving_wide.cu)
cudaError_t moving_launch_f64_r0(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches);
cudaError_t moving_launch_f32_r0(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches);
cudaError_t moving_launch_wide(cudaStream_t s, MovingParams &p, bool f64, const int64_t *gco, int64_t *launches);
// resident thread blocks per SM of the staged (moving_fast.cuh) kernels for this shape: sizes the chunks of the fast path
int moving_fast_blocks_per_sm(bool f64, int F, int kind, int n_cols);

// defined in moving_f64.cu / moving_f32.cu: dispatch on the number of coefficients
cudaError_t moving_launch_f64(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches);
cudaError_t moving_launch_f32(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches);

// builds the chunk table on the host (group offsets are host metadata), uploads it into `ws` and launches
int launch_moving(cudaStream_t stream, MovingParams &p, const int64_t *offsets_host, bool f64, int sm_count, char *ws, size_t ws_bytes,
                  int64_t *launches);

}  // namespace b200
