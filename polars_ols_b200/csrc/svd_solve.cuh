// svd_solve.cuh — minimum-norm least squares / ridge by singular value decomposition, on the device.
//
// Reference call sites restated (relative to /root/reference):
//   solve_ols_svd   src/least_squares.rs:183-191  LAPACK dgelsd through ndarray-linalg, rcond < 0 (machine eps):
//                   the minimum-norm solution, singular values <= eps * s_max treated as zero.  Chosen by
//                   solve_ols when solve_method == "svd" or n <= k (:225-229).
//   solve_ridge_svd src/least_squares.rs:106-168  beta = V diag(s / (s^2 + alpha)) U^T y, s < cutoff zeroed,
//                   cutoff = (rcond or eps * max(n, k)) * s_max (:140-148).
// dgelsd (bidiagonal divide & conquer) is third-party code that is not under /root/reference; what is
// restated here is its contract — the truncated pseudo-inverse solution — computed by one-sided Jacobi
// (Hestenes): column pairs of the n x k group matrix are rotated until mutually orthogonal, A V = U S; the
// column norms are the singular values (computed to high RELATIVE accuracy, so the eps * s_max truncation
// decides like LAPACK's).  One warp per selected group, lanes stride the rows; the group matrix lives
// column-major in the same absolute-row workspace as the QR fallback, V in a per-group k x k global record.
// This is the slow, robust path: it runs only for solve_method = "svd" and for groups with n <= k.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "qr_fallback.cuh"

namespace b200 {

struct SvdParams {
    QrParams q;          // columns, weights, mask, group offsets, workspace, beta, flags
    double *vws;         // [n_groups][F * F] right singular vectors, column c contiguous: V[i][c] at vws[c * F + i]
    int all_groups;      // 1: solve_method = "svd" (every non-empty group); 0: only groups flagged FLAG_WIDE
    int ridge;           // 1: ridge-SVD formula
    double alpha;
    double rcond;        // NaN = None
    int max_sweeps;
    int materialised;    // big.cuh: q.ws already holds the fit matrix
    int skip_rows_le_F;  // big.cuh: groups with n <= F rows are solved by big_svd_wide_kernel
    int n_rhs;           // multi-target: q.cols[kd .. kd + n_rhs) are the targets, workspace holds F + n_rhs columns,
                         // beta is [n_groups][n_rhs][F]; 0 = 1 (single target)
    int flag_mask;       // groups selected when !all_groups (0 = FLAG_WIDE)
};

template <typename T>
__global__ void __launch_bounds__(128) svd_solve_kernel(const SvdParams sp) {
    const QrParams &p = sp.q;
    const int lane = threadIdx.x & 31;
    const int64_t g = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (g >= p.n_groups) return;
    const int fl = p.flags[g];
    if (fl & FLAG_EMPTY) return;
    if (!sp.all_groups && !(fl & (sp.flag_mask ? sp.flag_mask : FLAG_WIDE))) return;
    const int n_rhs = sp.n_rhs > 0 ? sp.n_rhs : 1;
    const int F = p.F, kd = p.kd;
    const int64_t r0 = p.group_off[g], r1 = p.group_off[g + 1], n = r1 - r0;
    const int64_t N = p.n_rows;
    double *b = p.ws + static_cast<size_t>(F) * N;
    if (sp.skip_rows_le_F && n <= F) return;
    double *V = sp.vws + static_cast<size_t>(g) * F * F;
    // 1) materialise the fit matrix (sqrt-weight scaling, intercept, dropped rows -> zero rows) and V = I
    for (int64_t r = r0 + lane; r < r1 && !sp.materialised; r += 32) {
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
        for (int t = 0; t < n_rhs; ++t)
            b[static_cast<size_t>(t) * N + r] = keep ? static_cast<double>(static_cast<T>(static_cast<const T *>(p.cols[kd + t])[r] * s)) : 0.0;
    }
    for (int e = lane; e < F * F; e += 32) V[e] = ((e / F) == (e % F)) ? 1.0 : 0.0;
    __syncwarp();
    // 2) cyclic one-sided Jacobi sweeps
    for (int sweep = 0; sweep < sp.max_sweeps; ++sweep) {
        int rotated = 0;
        for (int pi = 0; pi < F - 1; ++pi) {
            for (int qi = pi + 1; qi < F; ++qi) {
                double *ap = p.ws + static_cast<size_t>(pi) * N + r0, *aq = p.ws + static_cast<size_t>(qi) * N + r0;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int64_t i = lane; i < n; i += 32) {
                    const double x = ap[i], y = aq[i];
                    al = fma(x, x, al);
                    be = fma(y, y, be);
                    ga = fma(x, y, ga);
                }
                al = warp_sum(al);
                be = warp_sum(be);
                ga = warp_sum(ga);
                if (ga == 0.0 || fabs(ga) <= DBL_EPSILON * sqrt(al * be)) continue;  // already orthogonal
                ++rotated;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int64_t i = lane; i < n; i += 32) {
                    const double x = ap[i], y = aq[i];
                    ap[i] = cs * x - sn * y;
                    aq[i] = sn * x + cs * y;
                }
                for (int i = lane; i < F; i += 32) {
                    const double x = V[pi * F + i], y = V[qi * F + i];
                    V[pi * F + i] = cs * x - sn * y;
                    V[qi * F + i] = sn * x + cs * y;
                }
                __syncwarp();
            }
        }
        if (rotated == 0) break;
    }
    // 3) singular values, projections a_c^T y, truncated pseudo-inverse
    double smax = 0.0;
    for (int c = 0; c < F; ++c) {
        const double *a = p.ws + static_cast<size_t>(c) * N + r0;
        double s2 = 0.0;
        for (int64_t i = lane; i < n; i += 32) s2 = fma(a[i], a[i], s2);
        s2 = warp_sum(s2);
        smax = fmax(smax, sqrt(s2));
    }
    const int64_t mx = (n > F) ? n : F;
    const double cutoff = sp.ridge ? ((sp.rcond == sp.rcond) ? sp.rcond : DBL_EPSILON * static_cast<double>(mx)) * smax
                                   : DBL_EPSILON * smax;
    for (int t = 0; t < n_rhs; ++t) {
        double *beta = p.beta + (g * n_rhs + t) * F;  // lane accumulates coefficients lane, lane + 32, ...
        const double *bt = b + static_cast<size_t>(t) * N;
        for (int i = lane; i < F; i += 32) beta[i] = 0.0;
        for (int c = 0; c < F; ++c) {
            const double *a = p.ws + static_cast<size_t>(c) * N + r0;
            double s2 = 0.0, ay = 0.0;
            for (int64_t i = lane; i < n; i += 32) {
                s2 = fma(a[i], a[i], s2);
                ay = fma(a[i], bt[r0 + i], ay);
            }
            s2 = warp_sum(s2);
            ay = warp_sum(ay);
            const double sv = sqrt(s2);
            double coef;
            if (sp.ridge) coef = (sv < cutoff) ? 0.0 : ay / (s2 + sp.alpha);   // V d U^T y with d = s / (s^2 + alpha)
            else coef = (sv <= cutoff) ? 0.0 : ay / s2;                         // V S^+ U^T y
            for (int i = lane; i < F; i += 32) beta[i] = fma(V[c * F + i], coef, beta[i]);
        }
    }
    if (lane == 0) p.flags[g] = (fl & ~(FLAG_ILLCOND | FLAG_LU_FALLBACK | FLAG_QR)) | FLAG_SVD;
}

inline cudaError_t launch_svd_solve(cudaStream_t stream, const SvdParams &sp, bool f64) {
    const unsigned blocks = static_cast<unsigned>((sp.q.n_groups * 32 + 127) / 128);
    if (f64) svd_solve_kernel<double><<<blocks, 128, 0, stream>>>(sp);
    else svd_solve_kernel<float><<<blocks, 128, 0, stream>>>(sp);
    return cudaGetLastError();
}

}  // namespace b200
