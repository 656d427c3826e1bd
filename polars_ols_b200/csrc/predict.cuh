// predict.cuh — second streaming pass for mode = predictions | residuals of the static models:
//   make_predictions            src/expressions.rs:175-195   (features . coefficients, validity mask)
//   predictions *= 1/sqrt_w     polars_ols/least_squares.py:234-235
//   residuals = target - pred   polars_ols/least_squares.py:238-239
// Row-parallel: a warp takes a chunk of PREDICT_CHUNK consecutive packed rows (so the grid balances by
// rows, not by groups), lanes stride the rows; plain coalesced loads (every
// column element is read once), the row's group is found by one binary search per lane and chunk and then
// advanced incrementally, beta comes from L1.  Writes go to the ORIGINAL row order through row_index
// (the scatter of `.over()`).
#pragma once
#include <cstdint>

#include "gram_stream.cuh"

namespace b200 {

constexpr int PREDICT_CHUNK = 512;

struct PredictParams {
    const void *cols[GRAM_MAX_COLS];  // [0,kd) features (null-cleaned: zero / NaN filled), [kd] weights
    int kd, intercept, F, has_w, w_is_sqrt;
    const void *target;               // raw target values (residuals); may be nullptr for predictions
    const uint8_t *target_validity;   // Arrow bitmap of the raw target (indexed by ORIGINAL row) or nullptr
    const void *mask;                 // T-typed row mask (policy "drop"): 0 -> prediction is null
    int64_t nseg, n_rows;
    const int64_t *seg_off;
    const int32_t *seg_group;
    const double *beta;               // [n_groups][F]  (group stride `beta_stride` doubles when != 0)
    int64_t beta_stride;              // multi-target: coefficients of target t live at beta + t*F, stride n_targets*F
    const int32_t *flags;             // [n_groups] or nullptr
    int only_flags;                   // != 0: only rows of groups with (flags & only_flags) are (re)written
    int nan_is_null;                  // multi-target: a NaN prediction is a null (convert_array_to_struct_series fill_nan)
    const int64_t *row_index;         // packed -> original row, or nullptr
    int target_is_packed;             // target pointer is indexed by packed position (cleaned copy)
    int residuals;
    double *out;                      // [n_rows] original order
    uint8_t *out_valid;               // bytes or nullptr
};

template <typename T> struct PV;
template <> struct PV<double> { using type = double2; static constexpr int N = 2; };
template <> struct PV<float> { using type = float4; static constexpr int N = 4; };

template <typename T>
__device__ __forceinline__ void predict_store(const PredictParams &p, int64_t r, double acc, T s, int64_t group) {
    if (p.only_flags && !(p.flags[group] & p.only_flags)) return;
    if (p.has_w) acc *= static_cast<double>(T(1) / s);  // predictions *= 1.0 / sqrt_w
    const int64_t orow = p.row_index ? p.row_index[r] : r;
    bool valid = true;
    if (p.mask) valid = static_cast<const T *>(p.mask)[r] != T(0);
    if (p.nan_is_null) valid = valid && (acc == acc);
    if (p.residuals) {
        const int64_t trow = p.target_is_packed ? r : orow;
        acc = static_cast<double>(static_cast<const T *>(p.target)[trow]) - acc;
        if (p.target_validity) valid = valid && ((p.target_validity[orow >> 3] >> (orow & 7)) & 1);
    }
    p.out[orow] = acc;
    if (p.out_valid) p.out_valid[orow] = valid ? 1 : 0;
}

template <typename T>
__device__ __forceinline__ T predict_scale(const PredictParams &p, T w) {
    return p.w_is_sqrt ? w : static_cast<T>(sqrt(w));
}

// one row (vector tails and rows of a vector that straddle a group boundary)
template <typename T>
__device__ __forceinline__ void predict_row(const PredictParams &p, int64_t r, const double *beta, int64_t group) {
    const int kd = p.kd;
    T s = T(1);
    if (p.has_w) s = predict_scale<T>(p, static_cast<const T *>(p.cols[kd])[r]);
    double acc = 0.0;
#pragma unroll 8
    for (int j = 0; j < kd; ++j) {
        const T x = static_cast<const T *>(p.cols[j])[r];
        acc = fma(static_cast<double>(static_cast<T>(x * s)), __ldg(beta + j), acc);
    }
    if (p.intercept) acc = fma(static_cast<double>(s), __ldg(beta + kd), acc);
    predict_store<T>(p, r, acc, s, group);
}

// PREDICT_COLS feature columns have their 16-byte loads issued together before any arithmetic; 4 columns at 64 registers
// (4 blocks per SM) measured best on C3: 2 / 3 / 6 / 8 / 16 columns, 48 / 80 registers and a double-buffered variant were
// all slower (profiles/r02_c3_experiments.json)
constexpr int PREDICT_COLS = 4;
template <typename T>
__global__ void __launch_bounds__(256, 4) predict_kernel(const PredictParams p) {
    constexpr int U = PREDICT_COLS;
    using Vec = typename PV<T>::type;
    constexpr int VN = PV<T>::N;  // rows per lane and iteration: one 16-byte load per column
    const int lane = threadIdx.x & 31;
    const int64_t wg = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t nchunks = (p.n_rows + PREDICT_CHUNK - 1) / PREDICT_CHUNK;
    const int kd = p.kd;
    const int64_t bs = p.beta_stride ? p.beta_stride : p.F;
    for (int64_t ch = wg; ch < nchunks; ch += nwarps) {
        const int64_t c0 = ch * PREDICT_CHUNK;
        const int64_t c1 = (c0 + PREDICT_CHUNK < p.n_rows) ? c0 + PREDICT_CHUNK : p.n_rows;
        int64_t r = c0 + static_cast<int64_t>(lane) * VN;
        if (r >= c1) continue;
        // segment of row r: largest s with seg_off[s] <= r (empty segments are skipped by the search)
        int64_t lo = 0, hi = p.nseg;  // invariant: seg_off[lo] <= r < seg_off[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (p.seg_off[mid] <= r) lo = mid; else hi = mid;
        }
        int64_t seg = lo;
        int64_t seg_end = p.seg_off[seg + 1];
        int64_t grp = p.seg_group ? p.seg_group[seg] : seg;
        const double *beta = p.beta + grp * bs;
        for (; r < c1; r += 32 * VN) {
            while (r >= seg_end) {
                ++seg;
                seg_end = p.seg_off[seg + 1];
                grp = p.seg_group ? p.seg_group[seg] : seg;
                beta = p.beta + grp * bs;
            }
            if (r + VN <= c1 && r + VN <= seg_end) {
                // whole vector inside one group: 16-byte loads, beta_j loaded once for the VN rows
                T sv[VN];
                double acc[VN];
#pragma unroll
                for (int v = 0; v < VN; ++v) { sv[v] = T(1); acc[v] = 0.0; }
                if (p.has_w) {
                    const Vec w4 = *reinterpret_cast<const Vec *>(static_cast<const T *>(p.cols[kd]) + r);
                    const T *wp = reinterpret_cast<const T *>(&w4);
#pragma unroll
                    for (int v = 0; v < VN; ++v) sv[v] = predict_scale<T>(p, wp[v]);
                }
                auto load_batch = [&](Vec (&xv)[U], int j0) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (j0 + u < kd) xv[u] = *reinterpret_cast<const Vec *>(static_cast<const T *>(p.cols[j0 + u]) + r);
                };
                auto use_batch = [&](const Vec (&xv)[U], int j0) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (j0 + u < kd) {
                            const T *xp = reinterpret_cast<const T *>(&xv[u]);
                            const double b = __ldg(beta + j0 + u);
#pragma unroll
                            for (int v = 0; v < VN; ++v) acc[v] = fma(static_cast<double>(static_cast<T>(xp[v] * sv[v])), b, acc[v]);
                        }
                    }
                };
                for (int j0 = 0; j0 < kd; j0 += U) {
                    Vec xv[U];
                    load_batch(xv, j0);
                    use_batch(xv, j0);
                }
                if (p.intercept) {
                    const double b = __ldg(beta + kd);
#pragma unroll
                    for (int v = 0; v < VN; ++v) acc[v] = fma(static_cast<double>(sv[v]), b, acc[v]);
                }
#pragma unroll
                for (int v = 0; v < VN; ++v) predict_store<T>(p, r + v, acc[v], sv[v], grp);
            } else {
                int64_t sg = seg, se = seg_end, gg = grp;
                const double *bb = beta;
                for (int v = 0; v < VN && r + v < c1; ++v) {
                    while (r + v >= se) {
                        ++sg;
                        se = p.seg_off[sg + 1];
                        gg = p.seg_group ? p.seg_group[sg] : sg;
                        bb = p.beta + gg * bs;
                    }
                    predict_row<T>(p, r + v, bb, gg);
                }
            }
        }
    }
}

}  // namespace b200
