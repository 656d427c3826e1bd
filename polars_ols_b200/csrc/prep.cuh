// prep.cuh — column staging on the device: null policies + `.over()` gather in ONE pass.
// Restates compute_is_valid_mask / handle_nulls / construct_features_array
// (src/expressions.rs:201-296, :22-63) and the sqrt-weight fill of polars_ols/least_squares.py:193:
//   out[c][p] = clean(in[c][row_index[p]])      p = packed position
//   ignore        : null -> NaN                 zero : null -> 0
//   drop*         : null -> 0 (prediction features) + T-typed row mask (1 = row enters the fit)
//   weights       : sqrt(w), null -> 1e-12
// Only launched when a validity bitmap or a row_index is present; null-free contiguous frames are
// streamed in place by the Gram kernel.
#pragma once
#include <cmath>
#include <cstdint>

#include "gram_stream.cuh"

namespace b200 {

enum : int { PREP_NAN = 0, PREP_ZERO = 1 };
enum : int { MASK_NONE = 0, MASK_ALL = 1, MASK_TARGET = 2 };

struct PrepParams {
    const void *in[GRAM_MAX_COLS];          // [0,kd) features, [kd] target, [kd+1] weights (optional)
    const uint8_t *validity[GRAM_MAX_COLS];  // Arrow bitmaps indexed by ORIGINAL row, or nullptr
    void *out[GRAM_MAX_COLS];               // packed, cleaned copies (same order)
    void *mask_out;                         // T-typed row mask or nullptr
    int kd, has_w;
    int fill;       // PREP_NAN | PREP_ZERO
    int mask_kind;  // MASK_*
    int n_targets;  // multi-target: the last n_targets - 1 "features" and the target are all targets (MASK_TARGET ANDs
                    // them, compute_is_valid_mask(.., Some(m)) src/expressions.rs:218-226); 0 / 1 = single target
    int64_t n_rows;
    int64_t n_rows_pad;  // outputs are zero-padded up to here
    const int64_t *row_index;
};

__device__ __forceinline__ bool bit_at(const uint8_t *bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

template <typename T>
__global__ void __launch_bounds__(256) prep_kernel(const PrepParams p) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int nc = p.kd + 1 + (p.has_w ? 1 : 0);
    const int ft = p.n_targets > 1 ? p.kd - (p.n_targets - 1) : p.kd;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < p.n_rows_pad; r += stride) {
        if (r >= p.n_rows) {  // padding rows: finite zeros so that masked smem reads stay harmless
            for (int c = 0; c < nc; ++c) static_cast<T *>(p.out[c])[r] = T(0);
            if (p.mask_out) static_cast<T *>(p.mask_out)[r] = T(0);
            continue;
        }
        const int64_t src = p.row_index ? p.row_index[r] : r;
        bool all_valid = true, y_valid = true;
        for (int c = 0; c <= p.kd; ++c) {
            const bool v = p.validity[c] ? bit_at(p.validity[c], src) : true;
            all_valid = all_valid && v;
            if (c >= ft) y_valid = y_valid && v;
            const T x = static_cast<const T *>(p.in[c])[src];
            static_cast<T *>(p.out[c])[r] = v ? x : (p.fill == PREP_NAN ? static_cast<T>(NAN) : T(0));
        }
        if (p.has_w) {
            const int c = p.kd + 1;
            const bool v = p.validity[c] ? bit_at(p.validity[c], src) : true;
            const T w = static_cast<const T *>(p.in[c])[src];
            static_cast<T *>(p.out[c])[r] = v ? static_cast<T>(sqrt(w)) : static_cast<T>(1.0e-12);
        }
        if (p.mask_out) {
            const bool m = (p.mask_kind == MASK_ALL) ? all_valid : ((p.mask_kind == MASK_TARGET) ? y_valid : true);
            static_cast<T *>(p.mask_out)[r] = m ? T(1) : T(0);
        }
    }
}

}  // namespace b200
