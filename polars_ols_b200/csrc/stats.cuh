// stats.cuh — mode = "statistics"  for every group of a frame.
//
// Reference restated (relative to /root/reference):
//   least_squares_statistics   src/expressions.rs:469-509   coefficients by the ordinary dispatch, then
//   compute_residual_metrics   src/statistics.rs:15-36      mse, mae, r2 of the FIT rows (scaled by sqrt(w) under WLS)
//   compute_feature_metrics    src/statistics.rs:77-156     (X^T X + lambda I)^-1 by Cholesky (failure -> NaN metrics),
//                                                           its own ridge coefficients inv . X^T y, rss, df = n - p
//                                                           (lambda = 0) or n - trace(inv), standard errors, t, p
// Three small passes after the streaming Gram kernel has written the per-group records:
//   stats_inverse_kernel  one thread per group: record -> explicit inverse (stats_math.cuh), diag, trace, ridge beta
//   stats_resid_kernel    one CTA per group (grid-stride): mean of y, then sum e^2, sum |e|, sum (y - mean)^2 for the
//                         dispatched coefficients and rss for the ridge coefficients — one extra read of the group
//                         (the L1 norm cannot come from the Gram), fixed-order block reductions (deterministic)
//   stats_final_kernel    one thread per (group, coefficient): df, sigma^2, se, t and the Student-t p-value
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "gram_stream.cuh"
#include "stats_math.cuh"

namespace b200 {

constexpr int STATS_GS = 8;  // doubles per group in `gstat`: trace, chol_ok, nfit, sse, sae, sst, rss(ridge beta), spare

struct StatsParams {
    const void *cols[GRAM_MAX_COLS];  // [0,kd) features, [kd] target — cleaned, packed (as the Gram kernel reads them)
    const void *w, *mask;
    int kd, intercept, F, w_is_sqrt;
    int64_t n_groups;
    const int64_t *group_off;       // [n_groups + 1] packed row ranges
    const double *partial;          // [nseg][F*F + F + 1] Gram records
    const int64_t *group_seg_off;   // [n_groups + 1] or nullptr (one record per group)
    double alpha;                   // lambda of compute_feature_metrics (kwargs.alpha)
    const double *beta;             // [n_groups][F] dispatched coefficients
    double *work;                   // [n_groups][2 F F + F]
    double *beta2, *inv_diag;       // [n_groups][F]
    double *gstat;                  // [n_groups][STATS_GS]
    double *r2, *mae, *mse;         // [n_groups]
    double *se, *tv, *pv;           // [n_groups][F]
    int32_t *bad_df;                // groups with df <= 0 (the reference asserts, src/statistics.rs:131-134)
};

__global__ void __launch_bounds__(64) stats_inverse_kernel(const StatsParams p) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= p.n_groups) return;
    const int F = p.F;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    const int64_t s0 = p.group_seg_off ? p.group_seg_off[g] : g;
    const int64_t s1 = p.group_seg_off ? p.group_seg_off[g + 1] : g + 1;
    double *A = p.work + static_cast<size_t>(g) * (2 * static_cast<size_t>(F) * F + F);
    double *M = A + static_cast<size_t>(F) * F;
    double *c = M + static_cast<size_t>(F) * F;
    for (int e = 0; e < F * F; ++e) {  // fixed-order sum of the segment records
        double s = 0.0;
        for (int64_t sg = s0; sg < s1; ++sg) s += p.partial[static_cast<size_t>(sg) * P + e];
        A[e] = s;
    }
    for (int e = 0; e < F; ++e) {
        double s = 0.0;
        for (int64_t sg = s0; sg < s1; ++sg) s += p.partial[static_cast<size_t>(sg) * P + static_cast<size_t>(F) * F + e];
        c[e] = s;
    }
    double nfit = 0.0;
    for (int64_t sg = s0; sg < s1; ++sg) nfit += p.partial[static_cast<size_t>(sg) * P + static_cast<size_t>(F) * F + F];
    for (int i = 0; i < F; ++i) A[i * F + i] += p.alpha;
    double *b2 = p.beta2 + g * F, *dg = p.inv_diag + g * F;
    double *gs = p.gstat + g * STATS_GS;
    gs[2] = nfit;
    if (chol_inverse(A, M, F, dg /* scratch, overwritten below */) != 0) {
        for (int i = 0; i < F; ++i) { b2[i] = NAN; dg[i] = NAN; }
        gs[0] = NAN;
        gs[1] = 0.0;
        return;
    }
    double tr = 0.0;
    for (int i = 0; i < F; ++i) {
        double s = 0.0;
        for (int j = 0; j < F; ++j) s += A[i * F + j] * c[j];  // coefficients = inv . X^T y (src/statistics.rs:116)
        b2[i] = s;
        dg[i] = A[i * F + i];
        tr += A[i * F + i];
    }
    gs[0] = tr;
    gs[1] = 1.0;
}

__device__ __forceinline__ double stats_block_sum(double v, double *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // sh may still be read from the previous reduction
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += sh[i];  // every thread sums in the same fixed order
    return s;
}

template <typename T>
__global__ void __launch_bounds__(256) stats_resid_kernel(const StatsParams p) {
    __shared__ double sh[8];
    __shared__ double sb[2 * 64];
    const int F = p.F, kd = p.kd;
    for (int64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
        const int64_t r0 = p.group_off[g], r1 = p.group_off[g + 1];
        __syncthreads();
        for (int i = threadIdx.x; i < F; i += blockDim.x) {
            sb[i] = p.beta[g * F + i];
            sb[64 + i] = p.beta2[g * F + i];
        }
        double sy = 0.0, cnt = 0.0;
        for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            if (p.mask && static_cast<const T *>(p.mask)[r] == T(0)) continue;
            T s = T(1);
            if (p.w) {
                const T wv = static_cast<const T *>(p.w)[r];
                s = p.w_is_sqrt ? wv : static_cast<T>(sqrt(wv));
            }
            sy += static_cast<double>(static_cast<T>(static_cast<const T *>(p.cols[kd])[r] * s));
            cnt += 1.0;
        }
        sy = stats_block_sum(sy, sh);
        cnt = stats_block_sum(cnt, sh);
        const double mean = cnt > 0.0 ? sy / cnt : 0.0;  // targets.mean().unwrap_or(0.0)
        double sse = 0.0, sae = 0.0, sst = 0.0, rss2 = 0.0;
        for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            if (p.mask && static_cast<const T *>(p.mask)[r] == T(0)) continue;
            T s = T(1);
            if (p.w) {
                const T wv = static_cast<const T *>(p.w)[r];
                s = p.w_is_sqrt ? wv : static_cast<T>(sqrt(wv));
            }
            const double y = static_cast<double>(static_cast<T>(static_cast<const T *>(p.cols[kd])[r] * s));
            double p1 = 0.0, p2 = 0.0;
            for (int c = 0; c < F; ++c) {
                const T xv = (c < kd) ? static_cast<const T *>(p.cols[c])[r] : T(1);
                const double x = static_cast<double>(static_cast<T>(xv * s));
                p1 = fma(x, sb[c], p1);
                p2 = fma(x, sb[64 + c], p2);
            }
            const double e1 = y - p1, e2 = y - p2, d = y - mean;
            sse = fma(e1, e1, sse);
            sae += fabs(e1);
            sst = fma(d, d, sst);
            rss2 = fma(e2, e2, rss2);
        }
        sse = stats_block_sum(sse, sh);
        sae = stats_block_sum(sae, sh);
        sst = stats_block_sum(sst, sh);
        rss2 = stats_block_sum(rss2, sh);
        if (threadIdx.x == 0) {
            double *gs = p.gstat + g * STATS_GS;
            gs[2] = cnt;
            gs[3] = sse;
            gs[4] = sae;
            gs[5] = sst;
            gs[6] = rss2;
        }
    }
}

__global__ void __launch_bounds__(128) stats_final_kernel(const StatsParams p) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int F = p.F;
    if (i >= p.n_groups * F) return;
    const int64_t g = i / F;
    const int j = static_cast<int>(i - g * F);
    const double *gs = p.gstat + g * STATS_GS;
    const double n = gs[2];
    if (j == 0) {
        p.mse[g] = gs[3] / n;
        p.mae[g] = gs[4] / n;
        p.r2[g] = 1.0 - gs[3] / gs[5];
    }
    if (gs[1] == 0.0) {  // Cholesky failed: NaN metrics (src/statistics.rs:101-111)
        p.se[i] = NAN;
        p.tv[i] = NAN;
        p.pv[i] = NAN;
        return;
    }
    const double df = p.alpha > 0.0 ? n - gs[0] : n - static_cast<double>(F);
    if (!(df > 0.0)) {
        if (j == 0) atomicAdd(p.bad_df, 1);
        p.se[i] = NAN;
        p.tv[i] = NAN;
        p.pv[i] = NAN;
        return;
    }
    const double sigma2 = gs[6] / df;
    const double se = sqrt(sigma2 * fabs(p.inv_diag[i]));
    const double t = p.beta2[i] / se;
    p.se[i] = se;
    p.tv[i] = t;
    p.pv[i] = students_t_two_sided_p(t, df);
}

}  // namespace b200
