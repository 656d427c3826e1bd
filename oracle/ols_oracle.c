/*
 * ols_oracle.c — CPU restatement (plain C, f64) of the polars_ols solver library.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (polars_ols_b200/) never imports or links this file.
 *
 * Every function cites the reference lines (relative to /root/reference) whose arithmetic it
 * restates.  Third-party arithmetic that is NOT under /root/reference (faer 0.18.2, ndarray
 * 0.15.6, LAPACK dgelsd through ndarray-linalg 0.16) is restated from its published algorithm:
 *   - faer col_piv_qr().solve_lstsq  -> Householder QR with column pivoting (Golub & Van Loan 5.4.2)
 *   - faer cholesky(Side::Lower)     -> LL^T, fails on a non-positive pivot
 *   - faer partial_piv_lu()          -> LU with row partial pivoting
 *   - LAPACK dgelsd                  -> not restated here; oracle/semantics.py calls numpy.linalg.lstsq
 *                                       (the same LAPACK routine) for the SVD paths.
 *
 * Parity pin: see oracle/README.md — checked against outputs the reference itself produced: its README
 * known-answer frame (tests/golden/readme_frame.json) and the executed cells of its demo notebook on the
 * notebook's own seeded data (tests/golden/notebook_outputs.json, 6 decimals: OLS / WLS / ridge / NNLS /
 * rolling / RLS / expanding / predict); then the Rust unit-test cases of src/lib.rs:47-171 and the
 * numpy / scikit-learn oracles the reference's own tests/test_ols.py uses.
 *
 * Layout conventions follow the reference after marshalling (src/expressions.rs:22-103):
 * X is ROW-MAJOR [n, k] f64, y is [n] f64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                         */
/* ------------------------------------------------------------------------------------------ */

/* XtX = X^T X, Xty = X^T y  (src/least_squares.rs:352-354: x.t().dot(x), x.t().dot(y)) */
static void gram(const double *x, const double *y, int64_t n, int k, double *xtx, double *xty) {
    memset(xtx, 0, sizeof(double) * (size_t)k * k);
    if (xty) memset(xty, 0, sizeof(double) * (size_t)k);
    for (int64_t i = 0; i < n; ++i) {
        const double *r = x + i * k;
        for (int a = 0; a < k; ++a) {
            const double ra = r[a];
            double *row = xtx + (size_t)a * k;
            for (int b = 0; b < k; ++b) row[b] += ra * r[b];
            if (xty) xty[a] += ra * y[i];
        }
    }
}

/* faer cholesky(Side::Lower): A = L L^T in place (lower), returns 0 on success, 1 on a
 * non-positive (or NaN) pivot.  Call sites: src/least_squares.rs:23,289. */
static int chol_factor(double *a, int k) {
    for (int j = 0; j < k; ++j) {
        double d = a[(size_t)j * k + j];
        for (int p = 0; p < j; ++p) d -= a[(size_t)j * k + p] * a[(size_t)j * k + p];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        a[(size_t)j * k + j] = d;
        for (int i = j + 1; i < k; ++i) {
            double s = a[(size_t)i * k + j];
            for (int p = 0; p < j; ++p) s -= a[(size_t)i * k + p] * a[(size_t)j * k + p];
            a[(size_t)i * k + j] = s / d;
        }
    }
    return 0;
}

static void chol_solve(const double *l, int k, double *b) {
    for (int i = 0; i < k; ++i) {
        double s = b[i];
        for (int p = 0; p < i; ++p) s -= l[(size_t)i * k + p] * b[p];
        b[i] = s / l[(size_t)i * k + i];
    }
    for (int i = k - 1; i >= 0; --i) {
        double s = b[i];
        for (int p = i + 1; p < k; ++p) s -= l[(size_t)p * k + i] * b[p];
        b[i] = s / l[(size_t)i * k + i];
    }
}

/* faer partial_piv_lu(): in-place LU with row pivoting; piv[i] = row swapped into i. */
static void lu_factor(double *a, int k, int *piv) {
    for (int j = 0; j < k; ++j) {
        int p = j;
        double best = fabs(a[(size_t)j * k + j]);
        for (int i = j + 1; i < k; ++i) {
            double v = fabs(a[(size_t)i * k + j]);
            if (v > best) { best = v; p = i; }
        }
        piv[j] = p;
        if (p != j)
            for (int c = 0; c < k; ++c) {
                double t = a[(size_t)j * k + c];
                a[(size_t)j * k + c] = a[(size_t)p * k + c];
                a[(size_t)p * k + c] = t;
            }
        double d = a[(size_t)j * k + j];
        for (int i = j + 1; i < k; ++i) {
            double f = a[(size_t)i * k + j] / d;
            a[(size_t)i * k + j] = f;
            for (int c = j + 1; c < k; ++c) a[(size_t)i * k + c] -= f * a[(size_t)j * k + c];
        }
    }
}

static void lu_solve(const double *lu, const int *piv, int k, double *b) {
    for (int j = 0; j < k; ++j) {
        int p = piv[j];
        if (p != j) { double t = b[j]; b[j] = b[p]; b[p] = t; }
    }
    for (int i = 0; i < k; ++i) {
        double s = b[i];
        for (int p = 0; p < i; ++p) s -= lu[(size_t)i * k + p] * b[p];
        b[i] = s;
    }
    for (int i = k - 1; i >= 0; --i) {
        double s = b[i];
        for (int p = i + 1; p < k; ++p) s -= lu[(size_t)i * k + p] * b[p];
        b[i] = s / lu[(size_t)i * k + i];
    }
}

/* inv(array, use_cholesky=false): LU inverse (src/least_squares.rs:20-39). out may not alias a. */
static void lu_inverse(const double *a, int k, double *out) {
    double *w = (double *)malloc(sizeof(double) * (size_t)k * k);
    int *piv = (int *)malloc(sizeof(int) * (size_t)k);
    double *col = (double *)malloc(sizeof(double) * (size_t)k);
    memcpy(w, a, sizeof(double) * (size_t)k * k);
    lu_factor(w, k, piv);
    for (int c = 0; c < k; ++c) {
        for (int i = 0; i < k; ++i) col[i] = (i == c) ? 1.0 : 0.0;
        lu_solve(w, piv, k, col);
        for (int i = 0; i < k; ++i) out[(size_t)i * k + c] = col[i];
    }
    free(w); free(piv); free(col);
}

/* ------------------------------------------------------------------------------------------ */
/* OLS / ridge / normal equations                                                              */
/* ------------------------------------------------------------------------------------------ */

/* solve_ols_qr (src/least_squares.rs:195-205): faer col_piv_qr().solve_lstsq.
 * Householder QR with column pivoting on a copy of X (n >= k expected), then R z = (Q^T y)[:k],
 * beta = P z.  No rank truncation (faer solves with the full R). */
ORC_API void orc_solve_ols_qr(const double *y, const double *x, int64_t n, int k, double *beta) {
    double *a = (double *)malloc(sizeof(double) * (size_t)n * k);
    double *b = (double *)malloc(sizeof(double) * (size_t)n);
    int *perm = (int *)malloc(sizeof(int) * (size_t)k);
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(a, x, sizeof(double) * (size_t)n * k);
    memcpy(b, y, sizeof(double) * (size_t)n);
    for (int j = 0; j < k; ++j) perm[j] = j;
    int steps = (int)((n < k) ? n : k);
    for (int j = 0; j < steps; ++j) {
        /* pivot: remaining column with the largest 2-norm */
        int p = j; double best = -1.0;
        for (int c = j; c < k; ++c) {
            double s = 0.0;
            for (int64_t i = j; i < n; ++i) s += a[i * k + c] * a[i * k + c];
            if (s > best) { best = s; p = c; }
        }
        if (p != j) {
            for (int64_t i = 0; i < n; ++i) { double t = a[i * k + j]; a[i * k + j] = a[i * k + p]; a[i * k + p] = t; }
            int t = perm[j]; perm[j] = perm[p]; perm[p] = t;
        }
        /* Householder vector for column j, rows j..n-1 */
        double norm = 0.0;
        for (int64_t i = j; i < n; ++i) norm += a[i * k + j] * a[i * k + j];
        norm = sqrt(norm);
        if (norm == 0.0) continue;
        double alpha = (a[(int64_t)j * k + j] > 0.0) ? -norm : norm;
        for (int64_t i = j; i < n; ++i) v[i] = a[i * k + j];
        v[j] -= alpha;
        double vtv = 0.0;
        for (int64_t i = j; i < n; ++i) vtv += v[i] * v[i];
        if (vtv == 0.0) continue;
        for (int c = j; c < k; ++c) {
            double s = 0.0;
            for (int64_t i = j; i < n; ++i) s += v[i] * a[i * k + c];
            s = 2.0 * s / vtv;
            for (int64_t i = j; i < n; ++i) a[i * k + c] -= s * v[i];
        }
        double s = 0.0;
        for (int64_t i = j; i < n; ++i) s += v[i] * b[i];
        s = 2.0 * s / vtv;
        for (int64_t i = j; i < n; ++i) b[i] -= s * v[i];
    }
    /* back substitution on the k x k upper triangle */
    double *z = (double *)malloc(sizeof(double) * (size_t)k);
    for (int i = k - 1; i >= 0; --i) {
        double s = (i < n) ? b[i] : 0.0;
        for (int c = i + 1; c < k; ++c) s -= ((i < n) ? a[(int64_t)i * k + c] : 0.0) * z[c];
        z[i] = (i < n) ? s / a[(int64_t)i * k + i] : 0.0;
    }
    for (int j = 0; j < k; ++j) beta[perm[j]] = z[j];
    free(a); free(b); free(perm); free(v); free(z);
}

/* solve_normal_equations (src/least_squares.rs:277-337) restricted to the methods reachable from
 * solve_ridge / rolling: method 0 = Cholesky with LU fallback (:284-316,:362), 1 = LU (:330-333).
 * a (k x k) and b (k) are overwritten; solution returned in b.  Returns 1 if the LU fallback ran. */
ORC_API int orc_solve_normal_equations(double *a, double *b, int k, int method) {
    int fell_back = 0;
    if (method == 0) {
        double *l = (double *)malloc(sizeof(double) * (size_t)k * k);
        memcpy(l, a, sizeof(double) * (size_t)k * k);
        if (chol_factor(l, k) == 0) {
            chol_solve(l, k, b);
            free(l);
            return 0;
        }
        free(l);
        fell_back = 1;
    }
    int *piv = (int *)malloc(sizeof(int) * (size_t)k);
    lu_factor(a, k, piv);
    lu_solve(a, piv, k, b);
    free(piv);
    return fell_back;
}

/* solve_ridge (src/least_squares.rs:342-371), Cholesky / LU methods: (X^T X + alpha I) beta = X^T y,
 * alpha NOT scaled by n (:352-356).  method: 0 = None/"chol" (Cholesky -> LU), 1 = "lu". */
ORC_API int orc_solve_ridge(const double *y, const double *x, int64_t n, int k, double alpha,
                            int method, double *beta) {
    double *xtx = (double *)malloc(sizeof(double) * (size_t)k * k);
    gram(x, y, n, k, xtx, beta);
    for (int j = 0; j < k; ++j) xtx[(size_t)j * k + j] += alpha;
    int r = orc_solve_normal_equations(xtx, beta, k, method);
    free(xtx);
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* elastic net by cyclic coordinate descent                                                    */
/* ------------------------------------------------------------------------------------------ */

/* soft_threshold (src/least_squares.rs:373-379) */
static double soft_threshold(double x, double alpha, int positive) {
    double sgn = (x > 0.0) - (x < 0.0);          /* f64::signum is +-1 for +-0, irrelevant: |x|-a<=0 */
    if (x == 0.0) sgn = signbit(x) ? -1.0 : 1.0;
    double r = sgn * fmax(fabs(x) - alpha, 0.0);
    if (positive) r = fmax(r, 0.0);
    return r;
}

/* solve_elastic_net (src/least_squares.rs:386-492).  Residual-form ("naive update") cyclic CD,
 * alpha *= n (:419), stop when ||w - w_old||_2 < tol (:436-444).  active_set != 0 selects
 * cd_active_set (:446-489).  Returns the number of sweeps executed. */
ORC_API int orc_solve_elastic_net(const double *y, const double *x, int64_t n, int k, double alpha,
                                  double l1_ratio, int64_t max_iter, double tol, int positive,
                                  int active_set, double *w) {
    double *diag = (double *)malloc(sizeof(double) * (size_t)k);
    double *res = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *w_old = (double *)malloc(sizeof(double) * (size_t)k);
    int *active = (int *)malloc(sizeof(int) * (size_t)k);
    int n_active = k;
    for (int j = 0; j < k; ++j) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += x[i * k + j] * x[i * k + j];
        diag[j] = s;               /* only the diagonal of xtx is used (:417,:431) */
        w[j] = 0.0;
        active[j] = j;
    }
    memcpy(res, y, sizeof(double) * (size_t)n);
    const double a = alpha * (double)n;
    int sweeps = 0;
    for (int64_t it = 0; it < max_iter; ++it) {
        ++sweeps;
        memcpy(w_old, w, sizeof(double) * (size_t)k);
        int n_loop = n_active;                       /* `for j in active_indices.clone()` (:459) */
        int *loop = (int *)malloc(sizeof(int) * (size_t)(n_loop > 0 ? n_loop : 1));
        memcpy(loop, active, sizeof(int) * (size_t)n_loop);
        for (int q = 0; q < n_loop; ++q) {
            const int j = loop[q];
            const double wj = w[j];
            for (int64_t i = 0; i < n; ++i) res[i] = res[i] + x[i * k + j] * wj;       /* :428 */
            double rho = 0.0;
            for (int64_t i = 0; i < n; ++i) rho += x[i * k + j] * res[i];              /* :430 */
            const double wn = soft_threshold(rho, a * l1_ratio, positive) / (diag[j] + a * (1.0 - l1_ratio));
            w[j] = wn;
            for (int64_t i = 0; i < n; ++i) res[i] = res[i] - x[i * k + j] * wn;       /* :433 */
            if (active_set && fabs(wn) < tol) {                                         /* :472-476 */
                int pos = -1;
                for (int t = 0; t < n_active; ++t) if (active[t] == j) { pos = t; break; }
                if (pos >= 0) {
                    for (int t = pos; t + 1 < n_active; ++t) active[t] = active[t + 1];
                    --n_active;
                }
            }
        }
        free(loop);
        double d2 = 0.0;
        for (int j = 0; j < k; ++j) d2 += (w[j] - w_old[j]) * (w[j] - w_old[j]);
        if (sqrt(d2) < tol) break;
    }
    free(diag); free(res); free(w_old); free(active);
    return sweeps;
}

/* ------------------------------------------------------------------------------------------ */
/* recursive least squares                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* RecursiveLeastSquares::{new,update} + solve_recursive_least_squares
 * (src/least_squares.rs:494-598).  half_life < 0 or NaN means None (lambda = 1).
 * coef_out is [n, k] row-major: theta AFTER the update with row t; invalid rows forward-fill. */
ORC_API void orc_solve_recursive_least_squares(const double *y, const double *x, int64_t n, int k,
                                               double half_life, double initial_state_covariance,
                                               const double *initial_state_mean,
                                               const uint8_t *is_valid, double *coef_out) {
    const double lam = (half_life == half_life && half_life > 0.0) ? exp(log(0.5) / half_life) : 1.0;
    double *p = (double *)calloc((size_t)k * k, sizeof(double));
    double *coef = (double *)calloc((size_t)k, sizeof(double));
    double *kg = (double *)calloc((size_t)k, sizeof(double));
    double *xp = (double *)malloc(sizeof(double) * (size_t)k);
    double *px = (double *)malloc(sizeof(double) * (size_t)k);
    for (int j = 0; j < k; ++j) p[(size_t)j * k + j] = initial_state_covariance;
    if (initial_state_mean) memcpy(coef, initial_state_mean, sizeof(double) * (size_t)k);
    for (int64_t t = 0; t < n; ++t) {
        if (!is_valid || is_valid[t]) {
            const double *xt = x + t * k;
            /* r = 1 + x^T P x / lambda (:532) evaluated as (x^T P) x */
            for (int j = 0; j < k; ++j) {
                double s = 0.0;
                for (int i = 0; i < k; ++i) s += xt[i] * p[(size_t)i * k + j];
                xp[j] = s;
            }
            double q = 0.0;
            for (int j = 0; j < k; ++j) q += xp[j] * xt[j];
            const double r = 1.0 + q / lam;
            /* K = P x / (r lambda) (:533-534) */
            for (int i = 0; i < k; ++i) {
                double s = 0.0;
                for (int j = 0; j < k; ++j) s += p[(size_t)i * k + j] * xt[j];
                px[i] = s;
            }
            for (int i = 0; i < k; ++i) kg[i] = px[i] / (r * lam);
            /* theta += K (y - x^T theta) (:535-536) */
            double pred = 0.0;
            for (int j = 0; j < k; ++j) pred += xt[j] * coef[j];
            const double resid = y[t] - pred;
            for (int j = 0; j < k; ++j) coef[j] = coef[j] + kg[j] * resid;
            /* P = P / lambda - K K^T r (:537-539) */
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                    p[(size_t)i * k + j] = p[(size_t)i * k + j] / lam - (kg[i] * kg[j]) * r;
        }
        memcpy(coef_out + t * k, coef, sizeof(double) * (size_t)k);
    }
    free(p); free(coef); free(kg); free(xp); free(px);
}

/* ------------------------------------------------------------------------------------------ */
/* rolling OLS                                                                                 */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int k;
    int woodbury;
    double *m;   /* xtx (non-Woodbury) or xtx_inv (Woodbury), k x k */
    double *xty; /* k */
} rolling_state;

/* woodbury_update (src/least_squares.rs:629-648) with U = x_update^T (K x r), V = U^T and a
 * diagonal C (update_xtx_inv :651-666): A^-1 <- A^-1 - A^-1 U (C^-1 + V A^-1 U)^-1 V A^-1. */
static void woodbury_rank_r(double *a_inv, int k, const double *xu /* r x k */, const double *cdiag, int r) {
    double *v_inv_a = (double *)malloc(sizeof(double) * (size_t)r * k); /* r x K */
    double *inv_a_u = (double *)malloc(sizeof(double) * (size_t)k * r); /* K x r */
    double mid[4], mid_inv[4];
    for (int a = 0; a < r; ++a)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int i = 0; i < k; ++i) s += xu[(size_t)a * k + i] * a_inv[(size_t)i * k + j];
            v_inv_a[(size_t)a * k + j] = s;
        }
    for (int i = 0; i < k; ++i)
        for (int a = 0; a < r; ++a) {
            double s = 0.0;
            for (int j = 0; j < k; ++j) s += a_inv[(size_t)i * k + j] * xu[(size_t)a * k + j];
            inv_a_u[(size_t)i * r + a] = s;
        }
    for (int a = 0; a < r; ++a)
        for (int b = 0; b < r; ++b) {
            double s = 0.0;
            for (int i = 0; i < k; ++i) s += xu[(size_t)a * k + i] * inv_a_u[(size_t)i * r + b];
            mid[a * r + b] = s + ((a == b) ? 1.0 / cdiag[a] : 0.0);
        }
    lu_inverse(mid, r, mid_inv);                                   /* inv(.., false) (:646) */
    double *tmp = (double *)malloc(sizeof(double) * (size_t)k * r);
    for (int i = 0; i < k; ++i)
        for (int b = 0; b < r; ++b) {
            double s = 0.0;
            for (int a = 0; a < r; ++a) s += inv_a_u[(size_t)i * r + a] * mid_inv[a * r + b];
            tmp[(size_t)i * r + b] = s;
        }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int b = 0; b < r; ++b) s += tmp[(size_t)i * r + b] * v_inv_a[(size_t)b * k + j];
            a_inv[(size_t)i * k + j] -= s;
        }
    free(v_inv_a); free(inv_a_u); free(tmp);
}

/* exported for the Rust unit-test analogues (src/lib.rs:124-171) */
ORC_API void orc_update_xtx_inv(double *xtx_inv, int k, const double *x_update, const double *cdiag, int r) {
    woodbury_rank_r(xtx_inv, k, x_update, cdiag, r);
}

ORC_API void orc_inv_lu(const double *a, int k, double *out) { lu_inverse(a, k, out); }

/* RollingOLSUpdate::update for both states (src/least_squares.rs:707-725, :749-776) */
static void state_update(rolling_state *s, const double *x_new, double y_new, const double *x_prev,
                         double y_prev) {
    const int k = s->k;
    if (!s->woodbury) {
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) s->m[(size_t)i * k + j] += x_new[i] * x_new[j];
        for (int i = 0; i < k; ++i) s->xty[i] = s->xty[i] + x_new[i] * y_new;
        if (x_prev) {
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) s->m[(size_t)i * k + j] -= x_prev[i] * x_prev[j];
            for (int i = 0; i < k; ++i) s->xty[i] = s->xty[i] - x_prev[i] * y_prev;
        }
    } else if (x_prev) {
        double *xu = (double *)malloc(sizeof(double) * 2 * (size_t)k);
        for (int j = 0; j < k; ++j) { xu[j] = -x_prev[j]; xu[k + j] = x_new[j]; }   /* :762-765 */
        const double c[2] = {-1.0, 1.0};                                             /* :744 */
        woodbury_rank_r(s->m, k, xu, c, 2);
        for (int i = 0; i < k; ++i) s->xty[i] = s->xty[i] + x_new[i] * y_new - x_prev[i] * y_prev;
        free(xu);
    } else {
        const double c[1] = {1.0};
        woodbury_rank_r(s->m, k, x_new, c, 1);
        for (int i = 0; i < k; ++i) s->xty[i] = s->xty[i] + x_new[i] * y_new;
    }
}

/* RollingOLSUpdate::subtract (:727-730, :778-782) */
static void state_subtract(rolling_state *s, const double *x_prev, double y_prev) {
    const int k = s->k;
    if (!s->woodbury) {
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) s->m[(size_t)i * k + j] -= x_prev[i] * x_prev[j];
        for (int i = 0; i < k; ++i) s->xty[i] = s->xty[i] - x_prev[i] * y_prev;
    } else {
        for (int i = 0; i < k; ++i) s->xty[i] = s->xty[i] - x_prev[i] * y_prev;
        const double c[1] = {-1.0};
        woodbury_rank_r(s->m, k, x_prev, c, 1);
    }
}

/* RollingOLSUpdate::solve (:732-734 Cholesky -> LU on a copy; :784-786 xtx_inv . xty) */
static void state_solve(rolling_state *s, double *out) {
    const int k = s->k;
    if (!s->woodbury) {
        double *a = (double *)malloc(sizeof(double) * (size_t)k * k);
        memcpy(a, s->m, sizeof(double) * (size_t)k * k);
        memcpy(out, s->xty, sizeof(double) * (size_t)k);
        orc_solve_normal_equations(a, out, k, 0);
        free(a);
    } else {
        for (int i = 0; i < k; ++i) {
            double v = 0.0;
            for (int j = 0; j < k; ++j) v += s->m[(size_t)i * k + j] * s->xty[j];
            out[i] = v;
        }
    }
}

/* solve_rolling_ols (src/least_squares.rs:848-1032).
 * min_periods < 0 -> None (min(k, window)); use_woodbury < 0 -> None (k > 60); alpha NaN -> 0.
 * null_branch: 0 = "last W valid rows" loop (Drop | DropZero | DropYZeroX, :947-986),
 *              1 = fixed row window loop (DropWindow | Zero | Ignore, :987-1029).
 * coef_out [n, k] row-major, NaN before warm-up. */
ORC_API void orc_solve_rolling_ols(const double *y, const double *x, int64_t n, int k,
                                   int64_t window_size, int64_t min_periods, int use_woodbury,
                                   double alpha, const uint8_t *is_valid_in, int null_branch,
                                   double *coef_out) {
    const double nan = NAN;
    for (int64_t i = 0; i < n * k; ++i) coef_out[i] = nan;
    if (n == 0) return;
    if (min_periods < 0) min_periods = (k < window_size) ? k : window_size;        /* :860 */
    const int woodbury = (use_woodbury < 0) ? (k > 60) : (use_woodbury != 0);     /* :863 */
    if (!(alpha == alpha)) alpha = 0.0;                                           /* :865 */
    uint8_t *is_valid = (uint8_t *)malloc((size_t)n);
    for (int64_t i = 0; i < n; ++i) is_valid[i] = is_valid_in ? is_valid_in[i] : 1;

    int64_t min_periods_valid = min_periods, n_valid = 0;                         /* :881-891 */
    for (int64_t i = 0; i < n; ++i) {
        if (is_valid[i]) n_valid += 1;
        if (n_valid == min_periods) { min_periods_valid = i + 1; break; }
    }
    if (n < ((n_valid > min_periods) ? n_valid : min_periods)) { free(is_valid); return; } /* :893-900 */

    int64_t *hist = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));         /* VecDeque :903 */
    int64_t h_head = 0, h_tail = 0;
    rolling_state st;
    st.k = k; st.woodbury = woodbury;
    st.m = (double *)calloc((size_t)k * k, sizeof(double));
    st.xty = (double *)calloc((size_t)k, sizeof(double));
    for (int64_t i = 0; i < min_periods_valid; ++i) {                             /* :909-921 */
        if (is_valid[i]) {
            const double *xi = x + i * k;
            for (int a = 0; a < k; ++a)
                for (int b = 0; b < k; ++b) st.m[(size_t)a * k + b] += xi[a] * xi[b];
            for (int a = 0; a < k; ++a) st.xty[a] = st.xty[a] + xi[a] * y[i];
            if (h_tail - h_head != window_size) hist[h_tail++] = i;
        }
    }
    if (alpha > 0.0) for (int a = 0; a < k; ++a) st.m[(size_t)a * k + a] += alpha; /* :924-926 */
    if (woodbury) {                                                                /* :929-932 */
        double *inv = (double *)malloc(sizeof(double) * (size_t)k * k);
        lu_inverse(st.m, k, inv);
        memcpy(st.m, inv, sizeof(double) * (size_t)k * k);
        free(inv);
    }
    double *coef_i = (double *)malloc(sizeof(double) * (size_t)k);
    state_solve(&st, coef_i);                                                      /* :939-943 */
    memcpy(coef_out + (min_periods_valid - 1) * k, coef_i, sizeof(double) * (size_t)k);

    if (null_branch == 0) {                                                        /* :947-986 */
        int saturated = (h_tail - h_head) == window_size;
        for (int64_t i = min_periods_valid; i < n; ++i) {
            if (is_valid[i]) {
                if (saturated) {
                    const int64_t i_start = hist[h_head];
                    state_update(&st, x + i * k, y[i], x + i_start * k, y[i_start]);
                    ++h_head;
                } else {
                    state_update(&st, x + i * k, y[i], NULL, 0.0);
                }
                state_solve(&st, coef_i);
                memcpy(coef_out + i * k, coef_i, sizeof(double) * (size_t)k);
                hist[h_tail++] = i;
                if (!saturated) saturated = (h_tail - h_head) == window_size;
            } else {
                memcpy(coef_out + i * k, coef_i, sizeof(double) * (size_t)k);
            }
        }
    } else {                                                                       /* :987-1029 */
        for (int64_t i = min_periods_valid; i < n; ++i) {
            const int64_t i_start = (i >= window_size) ? i - window_size : 0;      /* saturating_sub */
            const int v_i = is_valid[i], v_s = is_valid[i_start];
            int64_t n_valid_window = 0;
            for (int64_t t = i_start + 1; t <= i; ++t) n_valid_window += is_valid[t] ? 1 : 0;
            if (v_i) {
                if ((i >= window_size) && v_s)
                    state_update(&st, x + i * k, y[i], x + i_start * k, y[i_start]);
                else
                    state_update(&st, x + i * k, y[i], NULL, 0.0);
                if (n_valid_window >= n_valid) state_solve(&st, coef_i);
            } else if (v_s && !v_i && (i >= window_size)) {
                state_subtract(&st, x + i_start * k, y[i_start]);
                if (n_valid_window >= n_valid) state_solve(&st, coef_i);
            }
            memcpy(coef_out + i * k, coef_i, sizeof(double) * (size_t)k);
        }
    }
    free(is_valid); free(hist); free(st.m); free(st.xty); free(coef_i);
}

/* ------------------------------------------------------------------------------------------ */
/* grouped driver used as the timed CPU baseline ("port")                                      */
/* ------------------------------------------------------------------------------------------ */

/* What polars + the plugin do for `expr.over(group)` with contiguous groups, per group g:
 *   gather the k+1 column slices, copy them into a row-major [n_g, k] matrix
 *   (construct_features_array, src/expressions.rs:22-63) and call the solver
 *   (_get_least_squares_coefficients, src/expressions.rs:351-388).
 * Groups run in parallel (polars' rayon pool, README.md:19) -> OpenMP here.
 * model: 0 = ridge/OLS via normal equations (Cholesky -> LU), 1 = OLS via pivoted QR,
 *        2 = elastic net CD, 3 = elastic net CD with active set.
 * cols[0] = y, cols[1..k] = features (SoA, f64).  coef_out [G, k]. */
ORC_API void orc_grouped_least_squares_coefficients(const double *const *cols, int k,
                                                    const int64_t *offsets, int64_t n_groups,
                                                    int model, double alpha, double l1_ratio,
                                                    int64_t max_iter, double tol, int positive,
                                                    int n_threads, double *coef_out) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        double *xb = NULL, *yb = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t g = 0; g < n_groups; ++g) {
            const int64_t r0 = offsets[g], n = offsets[g + 1] - r0;
            double *beta = coef_out + g * k;
            if (n == 0) { for (int j = 0; j < k; ++j) beta[j] = 0.0; continue; }  /* :357-359 */
            if (n > cap) {
                free(xb); free(yb);
                cap = n;
                xb = (double *)malloc(sizeof(double) * (size_t)cap * k);
                yb = (double *)malloc(sizeof(double) * (size_t)cap);
            }
            memcpy(yb, cols[0] + r0, sizeof(double) * (size_t)n);
            for (int j = 0; j < k; ++j) {
                const double *c = cols[1 + j] + r0;
                for (int64_t i = 0; i < n; ++i) xb[i * k + j] = c[i];
            }
            if (model == 0) orc_solve_ridge(yb, xb, n, k, alpha, 0, beta);
            else if (model == 1) orc_solve_ols_qr(yb, xb, n, k, beta);
            else orc_solve_elastic_net(yb, xb, n, k, alpha, l1_ratio, max_iter, tol, positive, model == 3, beta);
        }
        free(xb); free(yb);
    }
}

/* CPU baseline of config C3 (bench only): `pl.col("y").least_squares.elastic_net(*x, sample_weights=w, mode="predictions")
 * .over(group)`: per group the reference multiplies target and features by sqrt(w) (polars_ols/least_squares.py:190-196),
 * copies into row-major f64 (src/expressions.rs:22-63), solves (model as above), predicts X beta (make_predictions,
 * src/expressions.rs:175-195) and un-scales by 1/sqrt(w) (least_squares.py:234-235).  weights may be NULL.  out [N]. */
ORC_API void orc_grouped_least_squares_predictions(const double *const *cols, int k, const double *weights,
                                                   const int64_t *offsets, int64_t n_groups, int model, double alpha,
                                                   double l1_ratio, int64_t max_iter, double tol, int positive,
                                                   int n_threads, double *out) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        double *xb = NULL, *yb = NULL, *beta = (double *)malloc(sizeof(double) * (size_t)k);
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t g = 0; g < n_groups; ++g) {
            const int64_t r0 = offsets[g], n = offsets[g + 1] - r0;
            if (n == 0) continue;
            if (n > cap) {
                free(xb); free(yb);
                cap = n;
                xb = (double *)malloc(sizeof(double) * (size_t)cap * k);
                yb = (double *)malloc(sizeof(double) * (size_t)cap);
            }
            for (int64_t i = 0; i < n; ++i) {
                const double s = weights ? sqrt(weights[r0 + i]) : 1.0;
                yb[i] = cols[0][r0 + i] * s;
                for (int j = 0; j < k; ++j) xb[i * k + j] = cols[1 + j][r0 + i] * s;
            }
            if (model == 0) orc_solve_ridge(yb, xb, n, k, alpha, 0, beta);
            else if (model == 1) orc_solve_ols_qr(yb, xb, n, k, beta);
            else orc_solve_elastic_net(yb, xb, n, k, alpha, l1_ratio, max_iter, tol, positive, model == 3, beta);
            for (int64_t i = 0; i < n; ++i) {
                double p = 0.0;
                for (int j = 0; j < k; ++j) p += xb[i * k + j] * beta[j];
                const double s = weights ? sqrt(weights[r0 + i]) : 1.0;
                out[r0 + i] = p * (1.0 / s);
            }
        }
        free(xb); free(yb); free(beta);
    }
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
