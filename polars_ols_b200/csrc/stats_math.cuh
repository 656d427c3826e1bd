// stats_math.cuh — Student-t two-sided p-value for mode = "statistics", host + device.
//
// Reference call site: t_value_to_p_value src/statistics.rs:44-48
//     2.0 * (1.0 - StudentsT::new(0, 1, df).cdf(|t|))
// StudentsT::cdf is statrs 0.17.1 (Cargo.lock; third-party, not under /root/reference).  Its published formula:
//     k = (x - loc) / scale,  h = df / (df + k^2),  ib = 0.5 * I_h(df / 2, 1 / 2),
//     cdf = ib if x <= loc else 1 - ib
// with I the regularised incomplete beta function.  Restated here with the classic modified-Lentz continued
// fraction (any implementation accurate to ~1e-14 agrees with statrs far inside the 1e-6 parity bar).  The
// reference's 1 - (1 - ib) cancellation is reproduced on purpose: p-values below ~1e-16 come out as exactly 0.
// `B200_HD` so tests/hostcheck can run exactly this code against scipy on a CPU-only box.
#pragma once
#include <cmath>

#include "solvers.cuh"

namespace b200 {

// continued fraction of the incomplete beta function (modified Lentz)
B200_HD double beta_cf(double a, double b, double x) {
    const double tiny = 1.0e-300, eps = 1.0e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 20000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}

// I_x(a, b) with y = 1 - x supplied separately (no cancellation when x is close to 1)
B200_HD double beta_reg_xy(double a, double b, double x, double y) {
    if (!(x > 0.0)) return 0.0;
    if (!(y > 0.0)) return 1.0;
    // ln B(a, b)^-1 = lgamma(a + b) - lgamma(a) - lgamma(b); for b = 1/2 and large a (a = df / 2: millions of rows) the
    // difference of two ~1e8-sized lgamma values would lose 8 digits, so the asymptotic ratio
    // Gamma(a + 1/2) / Gamma(a) = sqrt(a) (1 - 1/(8a) + 1/(128a^2) + 5/(1024a^3) - 21/(32768a^4) - 399/(262144a^5)) is used
    double lnorm;
    if (b == 0.5 && a > 200.0) {
        const double r = 1.0 / a;
        const double ser = 1.0 + r * (-0.125 + r * (1.0 / 128.0 + r * (5.0 / 1024.0 + r * (-21.0 / 32768.0 + r * (-399.0 / 262144.0)))));
        lnorm = 0.5 * log(a) + log(ser) - 0.5723649429247001 /* ln Gamma(1/2) */;
    } else {
        lnorm = lgamma(a + b) - lgamma(a) - lgamma(b);
    }
    const double lbt = lnorm + a * log(x) + b * log(y);
    const double bt = exp(lbt);
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * beta_cf(a, b, x) / a;
    return 1.0 - bt * beta_cf(b, a, y) / b;
}

B200_HD double students_t_two_sided_p(double t, double df) {
    if (t != t || df != df) return NAN;
    const double k = fabs(t);
    if (k == 0.0) return 1.0;                      // cdf(0) = ib = 0.5 -> 2 * (1 - 0.5)
    if (k > 1.0e300) return 0.0;                   // +-inf
    // h is rounded exactly as in statrs (df / (df + k^2)); 1 - h is taken from the ROUNDED h, so that for huge df and
    // tiny |t| the same point of the beta function is evaluated as in the reference (not the more accurate k^2 / den)
    const double h = df / (df + k * k);
    const double ib = 0.5 * beta_reg_xy(0.5 * df, 0.5, h, 1.0 - h);
    const double cdf = 1.0 - ib;                   // x > loc branch
    return 2.0 * (1.0 - cdf);
}

// Explicit inverse of A = X^T X + lambda I through its Cholesky factor, as compute_feature_metrics does
// (src/statistics.rs:95-116: faer cholesky(Lower).inverse(), Err -> NaN metrics).  A (row-major n x n, lower
// triangle read) is overwritten by the full symmetric inverse; M is n x n scratch; diag n doubles of scratch.
// Returns 0, or 1 when the factorisation hits a non-positive / NaN pivot (A is then garbage).
B200_HD int chol_inverse(double *A, double *M, int n, double *diag) {
    double mn, mx;
    if (chol_factor_lower(A, n, n, diag, &mn, &mx) != 0) return 1;
    for (int j = 0; j < n; ++j) {                       // M = L^-1 (lower triangular), column by column
        M[j * n + j] = 1.0 / A[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            double s = 0.0;
            for (int p = j; p < i; ++p) s += A[i * n + p] * M[p * n + j];
            M[i * n + j] = -s / A[i * n + i];
        }
    }
    for (int i = 0; i < n; ++i)                         // A^-1 = M^T M
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            for (int p = i; p < n; ++p) s += M[p * n + i] * M[p * n + j];
            A[i * n + j] = s;
            A[j * n + i] = s;
        }
    return 0;
}

}  // namespace b200
