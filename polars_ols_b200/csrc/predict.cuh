// predict.cuh — second streaming pass for mode = predictions | residuals of the static models:
//   make_predictions            src/expressions.rs:175-195   (features . coefficients, validity mask)
//   predictions *= 1/sqrt_w     polars_ols/least_squares.py:234-235
//   residuals = target - pred   polars_ols/least_squares.py:238-239
// One warp per segment, lanes stride the rows; plain coalesced 8/4-byte loads (each column element is
// read once), beta comes from L1 (warp-uniform address).  Writes go to the ORIGINAL row order through
// row_index (the scatter of `.over()`).
#pragma once
#include <cstdint>

#include "gram_stream.cuh"

namespace b200 {

struct PredictParams {
    const void *cols[GRAM_MAX_COLS];  // [0,kd) features (null-cleaned: zero / NaN filled), [kd] weights
    int kd, intercept, F, has_w, w_is_sqrt;
    const void *target;               // raw target values (residuals); may be nullptr for predictions
    const uint8_t *target_validity;   // Arrow bitmap of the raw target (indexed by ORIGINAL row) or nullptr
    const void *mask;                 // T-typed row mask (policy "drop"): 0 -> prediction is null
    int64_t nseg;
    const int64_t *seg_off;
    const int32_t *seg_group;
    const double *beta;               // [n_groups][F]
    const int64_t *row_index;         // packed -> original row, or nullptr
    int target_is_packed;             // target pointer is indexed by packed position (cleaned copy)
    int residuals;
    double *out;                      // [n_rows] original order
    uint8_t *out_valid;               // bytes or nullptr
};

template <typename T>
__global__ void __launch_bounds__(256) predict_kernel(const PredictParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t wg = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int kd = p.kd, F = p.F;
    for (int64_t seg = wg; seg < p.nseg; seg += nwarps) {
        const int64_t r0 = p.seg_off[seg], r1 = p.seg_off[seg + 1];
        const int64_t g = p.seg_group ? p.seg_group[seg] : seg;
        const double *beta = p.beta + g * F;
        for (int64_t r = r0 + lane; r < r1; r += 32) {
            T s = T(1);
            if (p.has_w) {
                const T w = static_cast<const T *>(p.cols[kd])[r];
                s = p.w_is_sqrt ? w : static_cast<T>(sqrt(w));
            }
            double acc = 0.0;
#pragma unroll 4
            for (int j = 0; j < kd; ++j) {
                const T x = static_cast<const T *>(p.cols[j])[r];
                acc = fma(static_cast<double>(static_cast<T>(x * s)), __ldg(beta + j), acc);
            }
            if (p.intercept) acc = fma(static_cast<double>(s), __ldg(beta + kd), acc);
            if (p.has_w) acc *= static_cast<double>(T(1) / s);  // predictions *= 1.0 / sqrt_w
            const int64_t orow = p.row_index ? p.row_index[r] : r;
            bool valid = true;
            if (p.mask) valid = static_cast<const T *>(p.mask)[r] != T(0);
            if (p.residuals) {
                const int64_t trow = p.target_is_packed ? r : orow;
                acc = static_cast<double>(static_cast<const T *>(p.target)[trow]) - acc;
                if (p.target_validity) valid = valid && ((p.target_validity[orow >> 3] >> (orow & 7)) & 1);
            }
            p.out[orow] = acc;
            if (p.out_valid) p.out_valid[orow] = valid ? 1 : 0;
        }
    }
}

}  // namespace b200
