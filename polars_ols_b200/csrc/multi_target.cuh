// multi_target.cuh — shared-X, many-y least squares.
//
// Reference restated (relative to /root/reference):
//   multi_target_least_squares src/expressions.rs:521-591   one SVD of the group's X, K x M coefficients, predictions
//   solve_multi_target         src/least_squares.rs:243-260 alpha > 0 -> solve_ridge_svd (:106-168) else solve_ols_svd
// Device design: the m targets ride through the streaming Gram kernel as extra columns — the record of the
// EXTENDED column set [x_0 .. x_kd-1, y_0 .. y_m-2, (const) | y_m-1] contains X^T X and every X^T y_t — so X is read
// ONCE for all targets.  This kernel solves the m right-hand sides of each group from that record by Cholesky
// (one thread per group: factor once, m substitutions).  For a full-rank, well-conditioned X this IS the SVD answer
// of the reference to ~cond * eps; groups where it is not (squared-pivot ratio above illcond_ratio, failed
// factorisation, n <= k) are flagged and re-solved from the data by the one-sided Jacobi SVD kernel with the same m
// right-hand sides (svd_solve.cuh: truncated pseudo-inverse / ridge-SVD formula with rcond).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "solvers.cuh"

namespace b200 {

struct MultiSolveParams {
    int kd, intercept, F;       // true features (without the targets): F = kd + intercept
    int m;                      // targets
    int FX;                     // coefficients of the EXTENDED Gram record: kd + (m - 1) + intercept
    int64_t n_groups;
    const double *partial;      // [nseg][FX*FX + FX + 1]
    const int64_t *group_seg_off;
    double alpha;               // ridge (0 = OLS)
    double illcond_ratio;
    double *work;               // [n_groups][F*F + F]
    double *beta;               // [n_groups][m][F]
    int32_t *flags;
};

__global__ void __launch_bounds__(64) multi_solve_kernel(const MultiSolveParams p) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= p.n_groups) return;
    const int F = p.F, FX = p.FX, kd = p.kd, m = p.m;
    const size_t P = static_cast<size_t>(FX) * FX + FX + 1;
    const int64_t s0 = p.group_seg_off ? p.group_seg_off[g] : g;
    const int64_t s1 = p.group_seg_off ? p.group_seg_off[g + 1] : g + 1;
    double *A = p.work + static_cast<size_t>(g) * (static_cast<size_t>(F) * F + F);
    double *diag = A + static_cast<size_t>(F) * F;
    double *B = p.beta + static_cast<size_t>(g) * m * F;
    auto ext = [&](int i) { return i < kd ? i : FX - 1; };  // the intercept is the LAST extended coefficient
    auto rec = [&](size_t e) {
        double s = 0.0;
        for (int64_t sg = s0; sg < s1; ++sg) s += p.partial[static_cast<size_t>(sg) * P + e];  // fixed order
        return s;
    };
    const double nfit = rec(static_cast<size_t>(FX) * FX + FX);
    if (nfit == 0.0) {  // solve_multi_target: x.is_empty() -> zeros (src/least_squares.rs:250-252)
        for (int e = 0; e < m * F; ++e) B[e] = 0.0;
        p.flags[g] = FLAG_EMPTY;
        return;
    }
    for (int i = 0; i < F; ++i)
        for (int j = 0; j < F; ++j) A[i * F + j] = rec(static_cast<size_t>(ext(i)) * FX + ext(j)) + ((i == j) ? p.alpha : 0.0);
    for (int t = 0; t < m; ++t)
        for (int i = 0; i < F; ++i)
            B[t * F + i] = (t < m - 1) ? rec(static_cast<size_t>(ext(i)) * FX + (kd + t)) : rec(static_cast<size_t>(FX) * FX + ext(i));
    int fl = 0;
    double mn, mx;
    if (chol_factor_lower(A, F, F, diag, &mn, &mx) != 0) {
        fl |= FLAG_ILLCOND | FLAG_LU_FALLBACK;  // the SVD kernel overwrites B
    } else {
        if (mx > p.illcond_ratio * mn) fl |= FLAG_ILLCOND;
        for (int t = 0; t < m; ++t) chol_solve_lower(A, F, F, B + t * F);
    }
    if (nfit <= static_cast<double>(F)) fl |= FLAG_WIDE;
    p.flags[g] = fl;
}

}  // namespace b200
