// big_stats.cuh — mode = "statistics" with MORE THAN 64 coefficients (the reference has no such bound:
// compute_feature_metrics, src/statistics.rs:76-156, inverts whatever X^T X it is handed).
//
// Runs behind the general path of big.cuh, which leaves per group the dense f64 fit matrix W[(F+1)][N] (scaled, null
// policy applied, column F = target), the row mask, the Gram record [F*F + F + 1] and the dispatched coefficients.
//   big_stats_factor_kernel   one CTA per group: A = lower(X^T X) + lambda I, Cholesky in global memory (big_chol;
//                             failure -> NaN metrics, src/statistics.rs:101-111), the feature metrics' own ridge
//                             coefficients (X^T X + lambda I)^-1 X^T y by two triangular solves
//   big_stats_invdiag_kernel  one THREAD per column j of L^-1 (forward substitution, no barriers; the columns of a warp
//                             advance together so L[i][p] is a broadcast and M[p][j] a coalesced access):
//                             diag((X^T X + lambda I)^-1)_j = |L^-1 e_j|^2 — the statistics only need the diagonal and
//                             the trace of the inverse (df = n - trace(inv), standard errors), never the inverse itself
//   big_stats_resid_kernel    one CTA per group: trace, mean of y, then sum e^2 / sum |e| / sum (y - mean)^2 for the
//                             dispatched coefficients and rss for the ridge coefficients over the FIT rows of W
//                             (fixed-order block reductions, as stats_resid_kernel)
// stats_final_kernel (stats.cuh) finishes: df, sigma^2, standard errors, t, Student-t p-values.
#pragma once
#include "big.cuh"
#include "stats.cuh"

namespace b200 {

struct BigStatsParams {
    int F;
    int64_t n_groups, n_rows;
    const int64_t *group_off;   // [G + 1] packed row ranges
    const double *W;            // [(F + 1)][n_rows]
    const uint8_t *mask;        // [n_rows] 1 = fit row
    const double *rec;          // [G][F*F + F + 1]
    const double *beta;         // [G][F] dispatched coefficients
    double alpha;
    double *A;                  // [G][F*F] Cholesky factor (lower)
    double *M;                  // [G][F*F] columns of L^-1, M[p * F + j]
    double *beta2, *inv_diag;   // [G][F]
    double *gstat;              // [G][STATS_GS]
};

inline size_t big_stats_factor_smem(int F) { return static_cast<size_t>(F) * 2 * 8 + 256; }

__global__ void __launch_bounds__(BIG_SOLVE_THREADS) big_stats_factor_kernel(const BigStatsParams p) {
    extern __shared__ __align__(16) unsigned char bs_smem[];
    const int F = p.F, tid = threadIdx.x;
    double *col = reinterpret_cast<double *>(bs_smem);
    double *z = col + F;
    const int64_t g = blockIdx.x;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    const double *G = p.rec + static_cast<size_t>(g) * P;
    const double *c = G + static_cast<size_t>(F) * F;
    double *A = p.A + static_cast<size_t>(g) * F * F;
    double *gs = p.gstat + g * STATS_GS;
    for (size_t e = tid; e < static_cast<size_t>(F) * F; e += BIG_SOLVE_THREADS) {
        const int i = static_cast<int>(e / F), j = static_cast<int>(e % F);
        if (j <= i) A[e] = G[e] + (i == j ? p.alpha : 0.0);
    }
    for (int i = tid; i < F; i += BIG_SOLVE_THREADS) z[i] = c[i];
    __syncthreads();
    double mn, mx;
    const bool ok = big_chol(A, F, col, &mn, &mx);  // uniform
    if (!ok) {
        for (int i = tid; i < F; i += BIG_SOLVE_THREADS) {
            p.beta2[g * F + i] = NAN;
            p.inv_diag[g * F + i] = NAN;
        }
        if (tid == 0) {
            gs[0] = NAN;
            gs[1] = 0.0;
            gs[2] = c[F];
        }
        return;
    }
    big_chol_solve(A, F, z);  // coefficients = inv . X^T y (src/statistics.rs:116)
    for (int i = tid; i < F; i += BIG_SOLVE_THREADS) p.beta2[g * F + i] = z[i];
    if (tid == 0) {
        gs[1] = 1.0;
        gs[2] = c[F];
    }
}

// grid (ceil(F / 256), min(G, 65535)), 256 threads: thread <-> column j of M = L^-1
__global__ void __launch_bounds__(256) big_stats_invdiag_kernel(const BigStatsParams p) {
    const int F = p.F;
    const int j = static_cast<int>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int jw = j & ~31;  // first column of this warp: the warp's columns share the loop bounds
    const bool live = j < F;
    if (jw >= F) return;
    for (int64_t g = blockIdx.y; g < p.n_groups; g += gridDim.y) {
    if (p.gstat[g * STATS_GS + 1] == 0.0) continue;  // factorisation failed: NaNs already written
    const double *L = p.A + static_cast<size_t>(g) * F * F;
    double *M = p.M + static_cast<size_t>(g) * F * F;
    double diag = 0.0;
    for (int i = jw; i < F; ++i) {
        const double *Li = L + static_cast<size_t>(i) * F;
        double s = (i == j) ? 1.0 : 0.0;
        for (int q = jw; q < i; ++q) {
            const double l = Li[q];                                             // broadcast
            const double m = (live && q >= j) ? M[static_cast<size_t>(q) * F + j] : 0.0;  // coalesced
            s = fma(-l, m, s);
        }
        const double v = (live && i >= j) ? s / Li[i] : 0.0;
        if (live && i >= j) {
            M[static_cast<size_t>(i) * F + j] = v;
            diag = fma(v, v, diag);
        }
    }
    if (live) p.inv_diag[g * F + j] = diag;
    }
}

__global__ void __launch_bounds__(256) big_stats_resid_kernel(const BigStatsParams p) {
    __shared__ double sh[8];
    const int F = p.F;
    const int64_t N = p.n_rows;
    for (int64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
        const int64_t r0 = p.group_off[g], r1 = p.group_off[g + 1];
        const double *b1 = p.beta + g * F, *b2 = p.beta2 + g * F;
        const double *y = p.W + static_cast<size_t>(F) * N;
        double *gs = p.gstat + g * STATS_GS;
        double tr = 0.0;
        for (int i = threadIdx.x; i < F; i += blockDim.x) tr += p.inv_diag[g * F + i];
        tr = stats_block_sum(tr, sh);
        double sy = 0.0, cnt = 0.0;
        for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x)
            if (p.mask[r]) {
                sy += y[r];
                cnt += 1.0;
            }
        sy = stats_block_sum(sy, sh);
        cnt = stats_block_sum(cnt, sh);
        const double mean = cnt > 0.0 ? sy / cnt : 0.0;  // targets.mean().unwrap_or(0.0)
        double sse = 0.0, sae = 0.0, sst = 0.0, rss2 = 0.0;
        for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            if (!p.mask[r]) continue;
            double p1 = 0.0, p2 = 0.0;
            for (int c = 0; c < F; ++c) {
                const double x = p.W[static_cast<size_t>(c) * N + r];
                p1 = fma(x, b1[c], p1);
                p2 = fma(x, b2[c], p2);
            }
            const double e1 = y[r] - p1, e2 = y[r] - p2, d = y[r] - mean;
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
            if (gs[1] != 0.0) gs[0] = tr;
            gs[2] = cnt;
            gs[3] = sse;
            gs[4] = sae;
            gs[5] = sst;
            gs[6] = rss2;
        }
    }
}

}  // namespace b200
