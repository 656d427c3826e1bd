// hostcheck.cpp — compiles the host/device templates of polars_ols_b200/csrc/{solvers,moving_core}.cuh
// with g++ so that the EXACT device algorithms (Cholesky->LU ladder, Gram-form coordinate descent,
// chunked rolling / RLS with halo + scan restarts) can be compared with the oracle on a CPU-only box.
// TEST HARNESS ONLY — never linked into, imported by or shipped with the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../polars_ols_b200/csrc/moving_core.cuh"
#include "../../polars_ols_b200/csrc/solvers.cuh"
#include "../../polars_ols_b200/csrc/stats_math.cuh"

using namespace b200;

#define HC_API extern "C" __attribute__((visibility("default")))

HC_API int hc_normal_equations(double *G, int F, double *c, int use_lu, double illcond_ratio) {
    std::vector<double> scratch(F);
    return normal_equations_solve(G, F, F, c, use_lu, scratch.data(), illcond_ratio);
}

HC_API int hc_cd_gram(const double *G, int F, const double *c, double a, double l1_ratio, int64_t max_iter, double tol,
                      int positive, int active_set, double *w) {
    std::vector<double> scratch(3 * F + 2);
    return cd_gram_solve(G, F, F, c, a, l1_ratio, max_iter, tol, positive != 0, active_set != 0, w, scratch.data());
}

HC_API double hc_students_t_p(double t, double df) { return students_t_two_sided_p(t, df); }

HC_API int hc_chol_inverse(double *A, int n) {
    std::vector<double> M(static_cast<size_t>(n) * n), diag(n);
    return chol_inverse(A, M.data(), n, diag.data());
}

template <int K>
struct HostSrc {
    const double *x;  // row-major [n, K]
    const double *y;
    const uint8_t *valid_;
    bool valid(int64_t r) const { return valid_ ? valid_[r] != 0 : true; }
    void load(int64_t r, double (&xo)[K], double &yo) const {
        for (int j = 0; j < K; ++j) xo[j] = x[r * K + j];
        yo = y[r];
    }
};

template <int K>
struct HostEmit {
    double *out;
    void operator()(int64_t r, const double (&beta)[K], bool) const {
        for (int j = 0; j < K; ++j) out[r * K + j] = beta[j];
    }
};

template <int K>
static void rolling_t(const double *y, const double *x, const uint8_t *valid, int64_t n, int64_t window,
                      int64_t min_periods, double alpha, int fixed_window, int64_t chunk, double *out) {
    HostSrc<K> src{x, y, valid};
    RollingCfg cfg{window, min_periods, alpha, fixed_window};
    RollingSeries rs = rolling_prepass(src, 0, n, min_periods);
    HostEmit<K> emit{out};
    for (int64_t c0 = 0; c0 < n; c0 += chunk) rolling_chunk<K>(src, cfg, rs, 0, n, c0, (c0 + chunk < n) ? c0 + chunk : n, emit);
}

HC_API void hc_rolling(const double *y, const double *x, const uint8_t *valid, int64_t n, int K, int64_t window,
                       int64_t min_periods, double alpha, int fixed_window, int64_t chunk, double *out) {
    switch (K) {
        case 1: rolling_t<1>(y, x, valid, n, window, min_periods, alpha, fixed_window, chunk, out); break;
        case 2: rolling_t<2>(y, x, valid, n, window, min_periods, alpha, fixed_window, chunk, out); break;
        case 3: rolling_t<3>(y, x, valid, n, window, min_periods, alpha, fixed_window, chunk, out); break;
        case 4: rolling_t<4>(y, x, valid, n, window, min_periods, alpha, fixed_window, chunk, out); break;
        case 5: rolling_t<5>(y, x, valid, n, window, min_periods, alpha, fixed_window, chunk, out); break;
        case 6: rolling_t<6>(y, x, valid, n, window, min_periods, alpha, fixed_window, chunk, out); break;
        case 7: rolling_t<7>(y, x, valid, n, window, min_periods, alpha, fixed_window, chunk, out); break;
        default: rolling_t<8>(y, x, valid, n, window, min_periods, alpha, fixed_window, chunk, out); break;
    }
}

template <int K>
static void rls_t(const double *y, const double *x, const uint8_t *valid, int64_t n, double lambda, double p0,
                  const double *mean, int64_t chunk, double *out) {
    HostSrc<K> src{x, y, valid};
    RlsCfg cfg{lambda, p0};
    HostEmit<K> emit{out};
    const int64_t nc = (n + chunk - 1) / chunk;
    std::vector<RlsSummary<K>> sums(nc);
    for (int64_t c = 0; c < nc; ++c) rls_summarise<K>(src, cfg, c * chunk, std::min<int64_t>((c + 1) * chunk, n), sums[c]);
    // exclusive scan (the device does this per element in rls_scan_kernel)
    NormalState<K> carry;
    carry.clear();
    for (int i = 0; i < K; ++i) {
        carry.S[i][i] = 1.0 / p0;
        carry.v[i] = (mean ? mean[i] : 0.0) / p0;
    }
    for (int64_t c = 0; c < nc; ++c) {
        NormalState<K> in = carry;
        for (int i = 0; i < K; ++i) {
            carry.v[i] = std::fma(sums[c].D, carry.v[i], sums[c].ab.v[i]);
            for (int j = 0; j <= i; ++j) carry.S[i][j] = std::fma(sums[c].D, carry.S[i][j], sums[c].ab.S[i][j]);
        }
        rls_chunk<K>(src, cfg, c == 0, mean, &in, c * chunk, std::min<int64_t>((c + 1) * chunk, n), emit);
    }
}

HC_API void hc_rls(const double *y, const double *x, const uint8_t *valid, int64_t n, int K, double lambda, double p0,
                   const double *mean, int64_t chunk, double *out) {
    switch (K) {
        case 1: rls_t<1>(y, x, valid, n, lambda, p0, mean, chunk, out); break;
        case 2: rls_t<2>(y, x, valid, n, lambda, p0, mean, chunk, out); break;
        case 3: rls_t<3>(y, x, valid, n, lambda, p0, mean, chunk, out); break;
        case 4: rls_t<4>(y, x, valid, n, lambda, p0, mean, chunk, out); break;
        case 5: rls_t<5>(y, x, valid, n, lambda, p0, mean, chunk, out); break;
        case 6: rls_t<6>(y, x, valid, n, lambda, p0, mean, chunk, out); break;
        case 7: rls_t<7>(y, x, valid, n, lambda, p0, mean, chunk, out); break;
        default: rls_t<8>(y, x, valid, n, lambda, p0, mean, chunk, out); break;
    }
}
