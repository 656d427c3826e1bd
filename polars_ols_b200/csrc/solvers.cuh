// solvers.cuh — small dense f64 solvers for the k x k normal systems (k small), shared by the fused
// Gram epilogue, the standalone per-group solve kernel and the moving-window kernel.
// `B200_HD` functions also compile for the host so tests/hostcheck can exercise exactly this code on a
// CPU-only box (host logic tests; the product never runs them on the CPU).
//
// Reference semantics restated (files relative to /root/reference):
//   solve_normal_equations  src/least_squares.rs:277-337  (Cholesky -> LU fallback, :284-316, :362)
//   solve_ols_lu            src/least_squares.rs:264-273
//   solve_elastic_net       src/least_squares.rs:386-492  (cyclic CD, alpha*n, ||dw||_2 < tol stop)
//   soft_threshold          src/least_squares.rs:373-379
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

enum : int {
    FLAG_LU_FALLBACK = 1,  // Cholesky hit a non-positive pivot, LU with partial pivoting ran
    FLAG_EMPTY = 2,        // no rows after null filtering: coefficients = 0
    FLAG_QR = 4,           // ill-conditioned: re-solved by the pivoted-QR kernel
    FLAG_ILLCOND = 8,      // internal: Cholesky pivot ratio says cond(G) is too large for 1e-6 parity
    FLAG_WIDE = 16,        // 0 < n <= k: the reference takes the LAPACK min-norm SVD path (src/least_squares.rs:225-229)
    FLAG_SVD = 32          // solved by the one-sided Jacobi SVD kernel
};

// In-place LL^T on the lower triangle of the symmetric matrix A (row-major, leading dim ld).
// The strict upper triangle is left untouched and `diag` receives a copy of the original diagonal so
// the matrix can be restored for the LU fallback.  Returns 0 on success, 1 on a non-positive / NaN pivot.
// min_piv / max_piv receive the extreme squared pivots (a cheap cond(G) lower bound).
B200_HD int chol_factor_lower(double *A, int ld, int n, double *diag, double *min_piv, double *max_piv) {
    double mn = INFINITY, mx = 0.0;
    for (int j = 0; j < n; ++j) diag[j] = A[j * ld + j];
    for (int j = 0; j < n; ++j) {
        double d = A[j * ld + j];
        for (int p = 0; p < j; ++p) d -= A[j * ld + p] * A[j * ld + p];
        if (!(d > 0.0)) return 1;
        mn = fmin(mn, d);
        mx = fmax(mx, d);
        d = sqrt(d);
        A[j * ld + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * ld + j];
            for (int p = 0; p < j; ++p) s -= A[i * ld + p] * A[j * ld + p];
            A[i * ld + j] = s / d;
        }
    }
    *min_piv = mn;
    *max_piv = mx;
    return 0;
}

B200_HD void chol_solve_lower(const double *L, int ld, int n, double *b) {
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int p = 0; p < i; ++p) s -= L[i * ld + p] * b[p];
        b[i] = s / L[i * ld + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int p = i + 1; p < n; ++p) s -= L[p * ld + i] * b[p];
        b[i] = s / L[i * ld + i];
    }
}

// restore the full symmetric matrix from its strict upper triangle + saved diagonal
B200_HD void restore_from_upper(double *A, int ld, int n, const double *diag) {
    for (int i = 0; i < n; ++i) {
        A[i * ld + i] = diag[i];
        for (int j = 0; j < i; ++j) A[i * ld + j] = A[j * ld + i];
    }
}

// LU with row partial pivoting, in place, then solve A x = b (b overwritten).  Row swaps are applied
// to b on the fly so no pivot vector is needed.
B200_HD void lu_solve_inplace(double *A, int ld, int n, double *b) {
    for (int j = 0; j < n; ++j) {
        int p = j;
        double best = fabs(A[j * ld + j]);
        for (int i = j + 1; i < n; ++i) {
            const double v = fabs(A[i * ld + j]);
            if (v > best) { best = v; p = i; }
        }
        if (p != j) {
            for (int c = 0; c < n; ++c) {
                const double t = A[j * ld + c];
                A[j * ld + c] = A[p * ld + c];
                A[p * ld + c] = t;
            }
            const double t = b[j]; b[j] = b[p]; b[p] = t;
        }
        const double d = A[j * ld + j];
        for (int i = j + 1; i < n; ++i) {
            const double f = A[i * ld + j] / d;
            A[i * ld + j] = f;
            for (int c = j + 1; c < n; ++c) A[i * ld + c] -= f * A[j * ld + c];
            b[i] -= f * b[j];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int p = i + 1; p < n; ++p) s -= A[i * ld + p] * b[p];
        b[i] = s / A[i * ld + i];
    }
}

// (G) beta = c with the reference's method ladder.  G must already contain +alpha*I.
// use_lu != 0 -> LU directly ("lu"); else Cholesky with LU fallback (None / "chol").
// scratch: n doubles.  Returns FLAG_* bits.
B200_HD int normal_equations_solve(double *G, int ld, int n, double *c, int use_lu, double *scratch,
                                   double illcond_ratio) {
    int flags = 0;
    if (!use_lu) {
        double mn, mx;
        if (chol_factor_lower(G, ld, n, scratch, &mn, &mx) == 0) {
            chol_solve_lower(G, ld, n, c);
            if (mx > illcond_ratio * mn) flags |= FLAG_ILLCOND;
            return flags;
        }
        restore_from_upper(G, ld, n, scratch);
        flags |= FLAG_LU_FALLBACK;
    }
    lu_solve_inplace(G, ld, n, c);
    return flags;
}

B200_HD double soft_threshold(double x, double t, bool positive) {
    // x.signum() * (|x| - t).max(0); signum(+-0) = +-1 but then (|x|-t).max(0) = 0 for t >= 0
    double r = copysign(fmax(fabs(x) - t, 0.0), x);
    if (positive) r = fmax(r, 0.0);
    return r;
}

// Cyclic coordinate descent on the Gram form of the reference's residual-form loop
// (src/least_squares.rs:422-489).  With q = c - G w maintained incrementally,
//   x_j^T (r + x_j w_j) = q_j + G_jj w_j      (the argument of soft_threshold at :430)
// and the iterate sequence, sweep count and stopping test are those of the reference (differences are
// rounding-level; verified in tests/test_hostcheck.py against oracle/ols_oracle.c).
// a = alpha * n_samples (:419).  scratch: 2n doubles + n ints (as doubles: 3n doubles).
// Returns the number of sweeps.
B200_HD int cd_gram_solve(const double *G, int ld, int n, const double *c, double a, double l1_ratio,
                          int64_t max_iter, double tol, bool positive, bool active_set, double *w,
                          double *scratch) {
    double *q = scratch;
    double *w_old = scratch + n;
    int *active = reinterpret_cast<int *>(scratch + 2 * n);
    int n_active = n;
    for (int j = 0; j < n; ++j) {
        w[j] = 0.0;
        q[j] = c[j];
        active[j] = j;
    }
    const double l1 = a * l1_ratio, l2 = a * (1.0 - l1_ratio);
    int sweeps = 0;
    for (int64_t it = 0; it < max_iter; ++it) {
        ++sweeps;
        for (int j = 0; j < n; ++j) w_old[j] = w[j];
        // the reference iterates over a clone of the active list taken at sweep start (:459) and
        // removes converged-to-zero coordinates from the live list (:472-476)
        const int n_loop = n_active;
        int wr = 0;
        for (int t = 0; t < n_loop; ++t) {
            const int j = active[t];
            const double gjj = G[j * ld + j];
            const double wj = w[j];
            const double rho = q[j] + gjj * wj;
            const double wn = soft_threshold(rho, l1, positive) / (gjj + l2);
            const double delta = wn - wj;
            w[j] = wn;
            if (delta != 0.0)
                for (int l = 0; l < n; ++l) q[l] -= G[j * ld + l] * delta;  // G symmetric: row j == column j
            if (!(active_set && fabs(wn) < tol)) active[wr++] = j;  // stable in-place compaction
        }
        n_active = wr;
        double d2 = 0.0;
        for (int j = 0; j < n; ++j) {
            const double d = w[j] - w_old[j];
            d2 += d * d;
        }
        if (sqrt(d2) < tol) break;
    }
    return sweeps;
}

}  // namespace b200
