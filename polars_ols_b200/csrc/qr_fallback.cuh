// qr_fallback.cuh — Householder QR with column pivoting for the groups whose Gram matrix is too
// ill-conditioned for the normal equations to reproduce the reference's OLS answer to 1e-6.
// The reference's default OLS path IS a pivoted QR (faer col_piv_qr().solve_lstsq,
// src/least_squares.rs:195-205, chosen at :225-229 when n > k); for well-conditioned groups
// Cholesky on the Gram matches it to ~1e-12, so only flagged groups (squared-pivot ratio above
// ILLCOND_RATIO, or a failed Cholesky) come here.  One warp per flagged group; lanes stride the rows;
// the (scaled, masked) group matrix lives column-major in a global workspace indexed by absolute row,
// so no per-group allocation is needed.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "gram_stream.cuh"
#include "solvers.cuh"

namespace b200 {

struct QrParams {
    const void *cols[GRAM_MAX_COLS];  // [0,kd) features, [kd] target
    const void *w, *mask;
    int kd, intercept, F, w_is_sqrt;
    int64_t n_groups, n_rows;
    const int64_t *group_off;
    double *ws;                       // [(F+1)][n_rows] column-major workspace
    double *beta;
    int32_t *flags;
    // more than 64 coefficients (big.cuh): the fit matrix is already in `ws` and the per-group scratch that
    // otherwise lives in registers / local memory comes from global memory
    int materialised;
    int *perm_ws;                     // [n_groups][F] or nullptr (F <= 64)
    double *z_ws;                     // [n_groups][F] or nullptr
    int all_flagged;                  // 1: every non-empty, non-wide group (no Cholesky pre-pass decided)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(128) qr_fallback_kernel(const QrParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (g >= p.n_groups) return;
    const int fl = p.flags[g];
    if ((!p.all_flagged && !(fl & (FLAG_ILLCOND | FLAG_LU_FALLBACK))) || (fl & (FLAG_EMPTY | FLAG_WIDE))) return;  // wide groups -> SVD kernel
    const int F = p.F, kd = p.kd;
    const int64_t r0 = p.group_off[g], r1 = p.group_off[g + 1], n = r1 - r0;
    const int64_t N = p.n_rows;
    double *b = p.ws + static_cast<size_t>(F) * N;
    // 1) materialise the fit matrix (sqrt-weight scaling, intercept, dropped rows -> zero rows)
    for (int64_t r = r0 + lane; r < r1 && !p.materialised; r += 32) {
        T s = T(1);
        if (p.w) {
            const T wv = static_cast<const T *>(p.w)[r];
            s = p.w_is_sqrt ? wv : static_cast<T>(sqrt(wv));
        }
        const bool keep = p.mask ? (static_cast<const T *>(p.mask)[r] != T(0)) : true;
        for (int c = 0; c < F; ++c) {
            const T x = (c < kd) ? static_cast<const T *>(p.cols[c])[r] : T(1);
            p.ws[static_cast<size_t>(c) * N + r] = keep ? static_cast<double>(static_cast<T>(x * s)) : 0.0;
        }
        b[r] = keep ? static_cast<double>(static_cast<T>(static_cast<const T *>(p.cols[kd])[r] * s)) : 0.0;
    }
    __syncwarp();
    int perm_local[64];
    int *perm = p.perm_ws ? p.perm_ws + g * F : perm_local;  // global scratch: written by lane 0 only
    const bool perm_shared = p.perm_ws != nullptr;
    for (int c = 0; c < F; ++c)
        if (!perm_shared || lane == 0) perm[c] = c;
    const int steps = (n < F) ? static_cast<int>(n) : F;
    for (int j = 0; j < steps; ++j) {
        // pivot: remaining column with the largest norm over rows j..n-1
        int piv = j;
        double best = -1.0;
        for (int c = j; c < F; ++c) {
            const double *a = p.ws + static_cast<size_t>(c) * N + r0;
            double s = 0.0;
            for (int64_t i = j + lane; i < n; i += 32) s = fma(a[i], a[i], s);
            s = warp_sum(s);
            if (s > best) { best = s; piv = c; }
        }
        if (piv != j) {
            double *a = p.ws + static_cast<size_t>(j) * N + r0, *c2 = p.ws + static_cast<size_t>(piv) * N + r0;
            for (int64_t i = lane; i < n; i += 32) { const double t = a[i]; a[i] = c2[i]; c2[i] = t; }
            if (!perm_shared || lane == 0) { const int t = perm[j]; perm[j] = perm[piv]; perm[piv] = t; }
            __syncwarp();
        }
        double *aj = p.ws + static_cast<size_t>(j) * N + r0;
        const double norm = sqrt(best);
        if (norm == 0.0) continue;
        const double ajj = aj[j];
        const double alpha = (ajj > 0.0) ? -norm : norm;
        // v = a_j[j:] - alpha e_1, stored in place (a_j[j] := v_0); R_jj = alpha kept in a register array below
        const double v0 = ajj - alpha;
        const double vtv = best - ajj * ajj + v0 * v0;
        __syncwarp();
        if (lane == 0) aj[j] = v0;
        __syncwarp();
        if (vtv != 0.0) {
            for (int c = j + 1; c <= F; ++c) {  // c == F -> right-hand side
                double *ac = (c < F) ? p.ws + static_cast<size_t>(c) * N + r0 : b + r0;
                double s = 0.0;
                for (int64_t i = j + lane; i < n; i += 32) s = fma(aj[i], ac[i], s);
                s = 2.0 * warp_sum(s) / vtv;
                for (int64_t i = j + lane; i < n; i += 32) ac[i] = fma(-s, aj[i], ac[i]);
                __syncwarp();
            }
        }
        __syncwarp();
        if (lane == 0) aj[j] = alpha;  // diagonal of R
        __syncwarp();
    }
    // back substitution on the F x F upper triangle (rows 0..F-1 of the workspace), R z = (Q^T b)[:F]
    if (lane == 0) {
        double z_local[64];
        double *z = p.z_ws ? p.z_ws + g * F : z_local;
        for (int i = F - 1; i >= 0; --i) {
            double s = (i < n) ? b[r0 + i] : 0.0;
            for (int c = i + 1; c < F; ++c) s -= ((i < n) ? p.ws[static_cast<size_t>(c) * N + r0 + i] : 0.0) * z[c];
            z[i] = (i < n) ? s / p.ws[static_cast<size_t>(i) * N + r0 + i] : 0.0;
        }
        for (int jx = 0; jx < F; ++jx) p.beta[g * F + perm[jx]] = z[jx];
        p.flags[g] = (fl & ~FLAG_ILLCOND) | FLAG_QR;
    }
}

inline cudaError_t launch_qr_fallback_kernel(cudaStream_t stream, const QrParams &qp, bool f64) {
    const unsigned blocks = static_cast<unsigned>((qp.n_groups * 32 + 127) / 128);
    if (f64) qr_fallback_kernel<double><<<blocks, 128, 0, stream>>>(qp);
    else qr_fallback_kernel<float><<<blocks, 128, 0, stream>>>(qp);
    return cudaGetLastError();
}

}  // namespace b200
