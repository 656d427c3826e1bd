// big.cuh — static models with MORE THAN 64 coefficients (the reference's tests/test_ols.py:272-313
// `test_fit_wide` runs ols / ridge / lasso with 100 and 1000 features on 10 rows).
//
// The fast path (gram_cta / gram_wide) keeps every column pointer and a whole k x k accumulator on chip; this
// path trades speed for generality (k up to BIG_MAX_F) and is never taken for k <= 64:
//   big_materialise : raw columns (device pointer TABLES: any number of columns) -> dense f64 fit matrix
//                     W[(F+1)][N] in packed (group-sorted) row order: sqrt-weight scaling in the column dtype,
//                     intercept column, null policy (fill / row mask -> zero rows); column F is the target.
//                     (_pre_process_data polars_ols/least_squares.py:163-196, handle_nulls /
//                     construct_features_array src/expressions.rs:22-63,257-296)
//   big_gram / big_xty : per group G = W^T W (64 x 64 tiles, f64 FMA), c = W^T y, n_fit -> the same
//                     [F*F + F + 1] record the small-k kernels hand to their solvers
//   big_solve       : one CTA per group, in global memory (L2 resident): Cholesky -> LU fallback / LU
//                     (solve_normal_equations src/least_squares.rs:277-337) or Gram-form cyclic coordinate
//                     descent (solve_elastic_net :386-492) for groups with n > k
//   big_cd_wide     : residual-form coordinate descent, exactly the reference's loop, one warp per group with
//                     n <= k (O(n) per coordinate instead of O(k))
//   big_svd_wide    : minimum-norm OLS / ridge-SVD for groups with n <= k by one-sided Jacobi on X^T
//                     (n columns of length k; solve_ols_svd :183-191, solve_ridge_svd :106-168)
//   big_predict     : predictions / residuals through the pointer tables (make_predictions
//                     src/expressions.rs:175-195 + the post-processing of polars_ols/least_squares.py:234-239)
// n > k groups that need the QR guard or solve_method="svd" reuse qr_fallback_kernel / svd_solve_kernel on W
// (`materialised` inputs, per-group scratch in global memory).
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "prep.cuh"
#include "qr_fallback.cuh"
#include "small_solve.cuh"
#include "solvers.cuh"

namespace b200 {

constexpr int BIG_MAX_F = 4096;
constexpr int BIG_TILE = 64;
constexpr int BIG_KC = 16;

struct BigParams {
    const void *const *vals;        // device table [kd + 1 + has_w]: features, target, weights (ORIGINAL row order)
    const uint8_t *const *valid;    // device table of Arrow bitmaps (entries may be nullptr)
    int kd, intercept, F, has_w, f32;
    int fill;                       // PREP_NAN | PREP_ZERO
    int mask_kind;                  // MASK_*
    int64_t n_rows, n_groups;
    const int64_t *row_index;       // packed -> original row, or nullptr
    const int64_t *group_off;       // device [G+1]
    double *W;                      // [(F+1)][n_rows]
    uint8_t *mask;                  // [n_rows] 1 = row enters the fit
    double *rec;                    // [G][F*F + F + 1]
    double *work;                   // [G][F*F] factorisation scratch (Cholesky route: always; LU route: when the record must survive) or nullptr
    double *beta;                   // [G][F]
    int32_t *flags;                 // [G]
    int route;                      // ROUTE_*
    double alpha, l1_ratio, tol, illcond_ratio;
    int64_t max_iter;
    int positive;
    int skip_wide;                  // n <= k groups are re-solved by the SVD kernel anyway: set the flag, skip the factorisation
    // svd (wide groups)
    int svd_ridge;
    double rcond;
    double *T;                      // [n_rows * F] per wide group: X^T, column r (a row of X) contiguous
    double *J;                      // [n_rows * F] per wide group: rotations, column-contiguous [n][n]
    int max_sweeps;
    // predict
    int residuals, drop_mask;
    double *out;
    uint8_t *out_valid;
};

__device__ __forceinline__ double big_round(double v, int f32) { return f32 ? static_cast<double>(static_cast<float>(v)) : v; }

__device__ __forceinline__ double big_load(const void *col, int64_t i, int f32) {
    return f32 ? static_cast<double>(static_cast<const float *>(col)[i]) : static_cast<const double *>(col)[i];
}

// sqrt-weight of a row in the column dtype (polars_ols/least_squares.py:193: null -> 1e-12)
__device__ __forceinline__ double big_sqrt_w(const BigParams &p, int64_t src) {
    if (!p.has_w) return 1.0;
    const int c = p.kd + 1;
    const bool v = p.valid[c] ? bit_at(p.valid[c], src) : true;
    if (!v) return big_round(1.0e-12, p.f32);
    if (p.f32) return static_cast<double>(sqrtf(static_cast<const float *>(p.vals[c])[src]));
    return sqrt(static_cast<const double *>(p.vals[c])[src]);
}

// x * s in the column dtype, as the reference's polars expression does before the cast to f64
__device__ __forceinline__ double big_scale(double x, double s, const BigParams &p) {
    if (!p.has_w) return x;
    return p.f32 ? static_cast<double>(static_cast<float>(x) * static_cast<float>(s)) : x * s;
}

__global__ void __launch_bounds__(256) big_materialise_kernel(const BigParams p) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t N = p.n_rows;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < N; r += stride) {
        const int64_t src = p.row_index ? p.row_index[r] : r;
        bool keep = true;
        if (p.mask_kind == MASK_ALL) {
            for (int c = 0; c <= p.kd; ++c)
                if (p.valid[c] && !bit_at(p.valid[c], src)) { keep = false; break; }
        } else if (p.mask_kind == MASK_TARGET) {
            keep = p.valid[p.kd] ? bit_at(p.valid[p.kd], src) : true;
        }
        p.mask[r] = keep ? 1 : 0;
        const double s = big_sqrt_w(p, src);
        const double fillv = (p.fill == PREP_NAN) ? NAN : 0.0;
        for (int c = 0; c <= p.kd; ++c) {
            double v = 0.0;
            if (keep) {
                const bool ok = p.valid[c] ? bit_at(p.valid[c], src) : true;
                v = big_scale(ok ? big_load(p.vals[c], src, p.f32) : fillv, s, p);
            }
            const int dst = (c < p.kd) ? c : p.F;  // the target goes last; the intercept sits at column kd
            p.W[static_cast<size_t>(dst) * N + r] = v;
        }
        if (p.intercept) p.W[static_cast<size_t>(p.kd) * N + r] = keep ? s : 0.0;
    }
}

// G tile (bi, bj), bi >= bj, of one group: 16 x 16 threads, 4 x 4 outputs each
__global__ void __launch_bounds__(256) big_gram_kernel(const BigParams p, int ntile) {
    __shared__ double As[BIG_KC][BIG_TILE + 1], Bs[BIG_KC][BIG_TILE + 1];
    const int64_t npair = static_cast<int64_t>(ntile) * (ntile + 1) / 2;
    const int64_t g = blockIdx.x / npair;
    const int pair = static_cast<int>(blockIdx.x % npair);
    int bi = static_cast<int>((sqrt(8.0 * pair + 1.0) - 1.0) * 0.5);
    while (static_cast<int64_t>(bi + 1) * (bi + 2) / 2 <= pair) ++bi;
    while (static_cast<int64_t>(bi) * (bi + 1) / 2 > pair) --bi;
    const int bj = pair - bi * (bi + 1) / 2;
    const int F = p.F;
    const int64_t N = p.n_rows, r0 = p.group_off[g], r1 = p.group_off[g + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const int lk = threadIdx.x & 15, lc = threadIdx.x >> 4;  // load: row lk of the chunk, columns lc + 16 m
    for (int64_t k0 = r0; k0 < r1; k0 += BIG_KC) {
        const int64_t r = k0 + lk;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int ci = bi * BIG_TILE + lc + 16 * m, cj = bj * BIG_TILE + lc + 16 * m;
            As[lk][lc + 16 * m] = (r < r1 && ci < F) ? p.W[static_cast<size_t>(ci) * N + r] : 0.0;
            Bs[lk][lc + 16 * m] = (r < r1 && cj < F) ? p.W[static_cast<size_t>(cj) * N + r] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BIG_KC; ++kk) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[kk][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
    double *G = p.rec + static_cast<size_t>(g) * (static_cast<size_t>(F) * F + F + 1);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = bi * BIG_TILE + ty * 4 + a, j = bj * BIG_TILE + tx * 4 + b;
            if (i < F && j < F) {
                G[static_cast<size_t>(i) * F + j] = acc[a][b];
                G[static_cast<size_t>(j) * F + i] = acc[a][b];
            }
        }
}

// one warp per (group, coefficient): c_j = W_j . y; the j == 0 warp also counts the fitted rows
__global__ void __launch_bounds__(256) big_xty_kernel(const BigParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int F = p.F;
    if (wid >= p.n_groups * F) return;
    const int64_t g = wid / F;
    const int j = static_cast<int>(wid % F);
    const int64_t N = p.n_rows, r0 = p.group_off[g], r1 = p.group_off[g + 1];
    const double *wj = p.W + static_cast<size_t>(j) * N, *y = p.W + static_cast<size_t>(F) * N;
    double s = 0.0, cnt = 0.0;
    for (int64_t r = r0 + lane; r < r1; r += 32) {
        s = fma(wj[r], y[r], s);
        if (j == 0) cnt += p.mask[r] ? 1.0 : 0.0;
    }
    s = warp_sum(s);
    double *rec = p.rec + static_cast<size_t>(g) * (static_cast<size_t>(F) * F + F + 1);
    if (lane == 0) rec[static_cast<size_t>(F) * F + j] = s;
    if (j == 0) {
        cnt = warp_sum(cnt);
        if (lane == 0) rec[static_cast<size_t>(F) * F + F] = cnt;
    }
}

// ---- one CTA per group: factorisations / Gram-form CD in global memory -------------------------------
constexpr int BIG_SOLVE_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < BIG_SOLVE_THREADS / 32; ++w) t += red[w];  // fixed order
    return t;
}

// in-place right-looking LL^T of the lower triangle of A (row-major F x F); same elimination order per
// element as chol_factor_lower.  Returns false on a non-positive / NaN pivot.  col: F doubles of shared memory.
__device__ bool big_chol(double *A, int F, double *col, double *mn_out, double *mx_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double mn = INFINITY, mx = 0.0;
    for (int j = 0; j < F; ++j) {
        __syncthreads();
        const double d = A[static_cast<size_t>(j) * F + j];
        if (!(d > 0.0)) return false;  // uniform
        mn = fmin(mn, d);
        mx = fmax(mx, d);
        const double sd = sqrt(d);
        for (int i = j + 1 + tid; i < F; i += BIG_SOLVE_THREADS) {
            const double v = A[static_cast<size_t>(i) * F + j] / sd;
            A[static_cast<size_t>(i) * F + j] = v;
            col[i] = v;
        }
        __syncthreads();
        if (tid == 0) A[static_cast<size_t>(j) * F + j] = sd;
        for (int i = j + 1 + warp; i < F; i += BIG_SOLVE_THREADS / 32) {
            const double li = col[i];
            double *row = A + static_cast<size_t>(i) * F;
            for (int l = j + 1 + lane; l <= i; l += 32) row[l] = fma(-li, col[l], row[l]);
        }
    }
    __syncthreads();
    *mn_out = mn;
    *mx_out = mx;
    return true;
}

// L z = c, L^T x = z  (z: F doubles of shared memory holding c on entry, x on exit)
__device__ void big_chol_solve(const double *L, int F, double *z) {
    const int tid = threadIdx.x;
    for (int j = 0; j < F; ++j) {
        __syncthreads();
        if (tid == 0) z[j] /= L[static_cast<size_t>(j) * F + j];
        __syncthreads();
        const double zj = z[j];
        for (int i = j + 1 + tid; i < F; i += BIG_SOLVE_THREADS) z[i] = fma(-L[static_cast<size_t>(i) * F + j], zj, z[i]);
    }
    for (int j = F - 1; j >= 0; --j) {
        __syncthreads();
        if (tid == 0) z[j] /= L[static_cast<size_t>(j) * F + j];
        __syncthreads();
        const double zj = z[j];
        const double *row = L + static_cast<size_t>(j) * F;
        for (int i = tid; i < j; i += BIG_SOLVE_THREADS) z[i] = fma(-row[i], zj, z[i]);
    }
    __syncthreads();
}

// LU with row partial pivoting in place + solve (lu_solve_inplace); z: right-hand side / solution, col: scratch
__device__ void big_lu_solve(double *A, int F, double *z, double *col, double *red, int *redi) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = 0; j < F; ++j) {
        __syncthreads();
        // pivot: first row with the largest |A[i][j]|, i >= j
        double best = -1.0;
        int bi = j;
        for (int i = j + tid; i < F; i += BIG_SOLVE_THREADS) {
            const double v = fabs(A[static_cast<size_t>(i) * F + j]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { red[warp] = best; redi[warp] = bi; }
        __syncthreads();
        best = red[0];
        bi = redi[0];
        for (int w = 1; w < BIG_SOLVE_THREADS / 32; ++w)
            if (red[w] > best || (red[w] == best && redi[w] < bi)) { best = red[w]; bi = redi[w]; }
        const int pv = (best >= 0.0) ? bi : j;  // all-NaN column: keep the row (NaNs propagate like the serial code)
        __syncthreads();
        if (pv != j) {
            double *rj = A + static_cast<size_t>(j) * F, *rp = A + static_cast<size_t>(pv) * F;
            for (int c = tid; c < F; c += BIG_SOLVE_THREADS) { const double t = rj[c]; rj[c] = rp[c]; rp[c] = t; }
            if (tid == 0) { const double t = z[j]; z[j] = z[pv]; z[pv] = t; }
            __syncthreads();
        }
        const double d = A[static_cast<size_t>(j) * F + j];
        for (int i = j + 1 + tid; i < F; i += BIG_SOLVE_THREADS) {
            const double f = A[static_cast<size_t>(i) * F + j] / d;
            A[static_cast<size_t>(i) * F + j] = f;
            col[i] = f;
        }
        __syncthreads();
        const double zj = z[j];
        const double *rj = A + static_cast<size_t>(j) * F;
        for (int i = j + 1 + warp; i < F; i += BIG_SOLVE_THREADS / 32) {
            const double f = col[i];
            double *row = A + static_cast<size_t>(i) * F;
            for (int c = j + 1 + lane; c < F; c += 32) row[c] = fma(-f, rj[c], row[c]);
            if (lane == 0) z[i] = fma(-f, zj, z[i]);
        }
    }
    for (int j = F - 1; j >= 0; --j) {
        __syncthreads();
        if (tid == 0) z[j] /= A[static_cast<size_t>(j) * F + j];
        __syncthreads();
        const double zj = z[j];
        for (int i = tid; i < j; i += BIG_SOLVE_THREADS) z[i] = fma(-A[static_cast<size_t>(i) * F + j], zj, z[i]);
    }
    __syncthreads();
}

// shared memory: col[F] z[F] wold[F] doubles + active[2][F] ints + reductions
inline size_t big_solve_smem(int F) { return static_cast<size_t>(F) * (3 * 8 + 2 * 4) + 256; }

__global__ void __launch_bounds__(BIG_SOLVE_THREADS) big_solve_kernel(const BigParams p) {
    extern __shared__ __align__(16) unsigned char big_smem[];
    const int F = p.F, tid = threadIdx.x;
    double *col = reinterpret_cast<double *>(big_smem);
    double *z = col + F;
    double *wold = z + F;
    int *act0 = reinterpret_cast<int *>(wold + F);
    int *act1 = act0 + F;
    double *red = reinterpret_cast<double *>(act1 + F);  // 2F ints: still 8-byte aligned
    int *redi = reinterpret_cast<int *>(red + 8);
    const int64_t g = blockIdx.x;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    double *G = p.rec + static_cast<size_t>(g) * P;
    const double *c = G + static_cast<size_t>(F) * F;
    const double nfit = c[F];
    double *beta = p.beta + g * F;
    if (nfit == 0.0) {  // src/expressions.rs:357-359
        for (int i = tid; i < F; i += BIG_SOLVE_THREADS) beta[i] = 0.0;
        if (tid == 0) p.flags[g] = FLAG_EMPTY;
        return;
    }
    const bool wide = nfit <= static_cast<double>(F);
    if (p.route == ROUTE_FLAGS_ONLY || (wide && p.skip_wide)) {  // an SVD kernel solves: only the flags it selects on
        if (tid == 0) p.flags[g] = wide ? FLAG_WIDE : 0;
        return;
    }
    if (p.route == ROUTE_CD || p.route == ROUTE_CD_ACTIVE) {
        if (wide) return;  // big_cd_wide_kernel
        // Gram-form cyclic coordinate descent: cd_gram_solve (solvers.cuh) with the q update spread over the CTA
        double *w = z, *q = col;
        const double a = p.alpha * nfit;  // src/least_squares.rs:419
        const double l1 = a * p.l1_ratio, l2 = a * (1.0 - p.l1_ratio);
        const bool active_set = p.route == ROUTE_CD_ACTIVE;
        for (int j = tid; j < F; j += BIG_SOLVE_THREADS) { w[j] = 0.0; q[j] = c[j]; act0[j] = j; }
        int n_active = F;
        int *cur = act0, *nxt = act1;
        for (int64_t it = 0; it < p.max_iter; ++it) {
            __syncthreads();
            for (int j = tid; j < F; j += BIG_SOLVE_THREADS) wold[j] = w[j];
            int wr = 0;
            for (int t = 0; t < n_active; ++t) {
                __syncthreads();
                const int j = cur[t];
                const double gjj = G[static_cast<size_t>(j) * F + j];
                const double wj = w[j];
                const double rho = q[j] + gjj * wj;
                const double wn = soft_threshold(rho, l1, p.positive != 0) / (gjj + l2);
                const double delta = wn - wj;
                __syncthreads();
                if (tid == 0) w[j] = wn;
                if (delta != 0.0) {
                    const double *row = G + static_cast<size_t>(j) * F;
                    for (int l = tid; l < F; l += BIG_SOLVE_THREADS) q[l] = fma(-row[l], delta, q[l]);
                }
                if (!(active_set && fabs(wn) < p.tol)) {
                    if (tid == 0) nxt[wr] = j;
                    ++wr;
                }
            }
            n_active = wr;
            int *tsw = cur; cur = nxt; nxt = tsw;
            __syncthreads();
            double d2 = 0.0;
            for (int j = tid; j < F; j += BIG_SOLVE_THREADS) { const double d = w[j] - wold[j]; d2 = fma(d, d, d2); }
            d2 = block_sum(d2, red);
            if (sqrt(d2) < p.tol) break;
        }
        __syncthreads();
        for (int i = tid; i < F; i += BIG_SOLVE_THREADS) beta[i] = w[i];
        if (tid == 0) p.flags[g] = 0;
        return;
    }
    // the factorisations run on a copy when scratch is provided (always for the Cholesky route, whose LU fallback needs
    // the matrix again; for the LU route when mode = "statistics" still needs the record afterwards), else in place
    double *A = p.work ? p.work + static_cast<size_t>(g) * F * F : G;
    auto load_regularised = [&]() {
        if (A != G)
            for (size_t e = tid; e < static_cast<size_t>(F) * F; e += BIG_SOLVE_THREADS) A[e] = G[e];
        __syncthreads();
        for (int i = tid; i < F; i += BIG_SOLVE_THREADS) A[static_cast<size_t>(i) * F + i] += p.alpha;
        __syncthreads();
    };
    for (int i = tid; i < F; i += BIG_SOLVE_THREADS) z[i] = c[i];
    load_regularised();
    int fl = 0;
    bool done = false;
    if (p.route == ROUTE_CHOL) {
        double mn, mx;
        if (big_chol(A, F, col, &mn, &mx)) {
            big_chol_solve(A, F, z);
            if (mx > p.illcond_ratio * mn) fl |= FLAG_ILLCOND;
            done = true;
        } else {
            fl |= FLAG_LU_FALLBACK;
            __syncthreads();
            load_regularised();  // the failed factorisation overwrote part of the copy
        }
    }
    if (!done) big_lu_solve(A, F, z, col, red, redi);
    if (wide) fl |= FLAG_WIDE;
    for (int i = tid; i < F; i += BIG_SOLVE_THREADS) beta[i] = z[i];
    if (tid == 0) p.flags[g] = fl;
}

// Residual-form coordinate descent for groups with n <= k: the reference's own loop
// (src/least_squares.rs:422-489; oracle/ols_oracle.c orc_solve_elastic_net), one warp per group, lanes stride
// the rows; the residual lives in column F of W (the target, updated in place), the coefficients in beta.
// iws: [G][2F] ints (active lists), dws: [G][2F] doubles (column norms, w_old).
__global__ void __launch_bounds__(128) big_cd_wide_kernel(const BigParams p, int *iws, double *dws) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (g >= p.n_groups) return;
    const int F = p.F;
    const double nfit = p.rec[static_cast<size_t>(g) * (static_cast<size_t>(F) * F + F + 1) + static_cast<size_t>(F) * F + F];
    if (nfit == 0.0 || nfit > static_cast<double>(F)) return;  // empty: big_solve wrote zeros; n > k: big_solve
    const int64_t N = p.n_rows, r0 = p.group_off[g], n = p.group_off[g + 1] - r0;
    double *res = p.W + static_cast<size_t>(F) * N + r0;
    double *w = p.beta + g * F;
    double *diag = dws + static_cast<size_t>(g) * 2 * F, *wold = diag + F;
    int *cur = iws + static_cast<size_t>(g) * 2 * F, *nxt = cur + F;
    for (int j = 0; j < F; ++j) {
        const double *xj = p.W + static_cast<size_t>(j) * N + r0;
        double s = 0.0;
        for (int64_t i = lane; i < n; i += 32) s = fma(xj[i], xj[i], s);
        s = warp_sum(s);
        if (lane == 0) { diag[j] = s; w[j] = 0.0; cur[j] = j; }
    }
    __syncwarp();
    const double a = p.alpha * nfit;
    const double l1 = a * p.l1_ratio, l2 = a * (1.0 - p.l1_ratio);
    const bool active_set = p.route == ROUTE_CD_ACTIVE;
    int n_active = F;
    for (int64_t it = 0; it < p.max_iter; ++it) {
        for (int j = lane; j < F; j += 32) wold[j] = w[j];
        __syncwarp();
        int wr = 0;
        for (int t = 0; t < n_active; ++t) {
            const int j = cur[t];
            const double *xj = p.W + static_cast<size_t>(j) * N + r0;
            const double wj = w[j];
            double rho = 0.0;
            for (int64_t i = lane; i < n; i += 32) {
                const double ri = fma(xj[i], wj, res[i]);  // residual + x_j w_j (:428)
                res[i] = ri;
                rho = fma(xj[i], ri, rho);                 // :430
            }
            rho = warp_sum(rho);
            const double wn = soft_threshold(rho, l1, p.positive != 0) / (diag[j] + l2);
            for (int64_t i = lane; i < n; i += 32) res[i] = fma(-xj[i], wn, res[i]);  // :433
            if (lane == 0) w[j] = wn;
            if (!(active_set && fabs(wn) < p.tol)) {
                if (lane == 0) nxt[wr] = j;
                ++wr;
            }
            __syncwarp();
        }
        n_active = wr;
        int *tsw = cur; cur = nxt; nxt = tsw;
        double d2 = 0.0;
        for (int j = lane; j < F; j += 32) { const double d = w[j] - wold[j]; d2 = fma(d, d, d2); }
        d2 = warp_sum(d2);
        if (sqrt(d2) < p.tol) break;
    }
    if (lane == 0) p.flags[g] = 0;
}

// Minimum-norm least squares / ridge-SVD for groups with n <= k: one-sided Jacobi on A = X^T (k x n), i.e. the
// n rows of X are rotated until mutually orthogonal, A Jr = B with orthogonal columns b_c = sigma_c q_c, so
// X = Jr Sigma Q^T and   beta = sum_c  b_c * (Jr_c . y) * d_c,
//   d_c = 1 / sigma_c^2 (sigma_c > eps * sigma_max, else 0)            solve_ols_svd  (LAPACK dgelsd contract)
//   d_c = 1 / (sigma_c^2 + alpha) (sigma_c >= cutoff, else 0)          solve_ridge_svd (:140-148)
// One warp per wide group; p.T / p.J hold the group's transposed block and rotation matrix at offset r0 * F.
__global__ void __launch_bounds__(128) big_svd_wide_kernel(const BigParams p, int all_groups) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (g >= p.n_groups) return;
    const int fl = p.flags[g];
    if (fl & FLAG_EMPTY) return;
    const int F = p.F;
    const int64_t N = p.n_rows, r0 = p.group_off[g], n = p.group_off[g + 1] - r0;
    if (n > F || n == 0) return;                       // n > k: svd_solve_kernel
    if (!all_groups && !(fl & FLAG_WIDE)) return;
    double *A = p.T + static_cast<size_t>(r0) * F;    // column r (length F) at A + r * F
    double *Jr = p.J + static_cast<size_t>(r0) * F;   // n x n, column c at Jr + c * n
    const double *y = p.W + static_cast<size_t>(F) * N + r0;
    for (int64_t r = 0; r < n; ++r)
        for (int c = lane; c < F; c += 32) A[r * F + c] = p.W[static_cast<size_t>(c) * N + r0 + r];
    for (int64_t e = lane; e < n * n; e += 32) Jr[e] = ((e / n) == (e % n)) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < p.max_sweeps; ++sweep) {
        int rotated = 0;
        for (int64_t pi = 0; pi < n - 1; ++pi)
            for (int64_t qi = pi + 1; qi < n; ++qi) {
                double *ap = A + pi * F, *aq = A + qi * F;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int i = lane; i < F; i += 32) {
                    const double x = ap[i], v = aq[i];
                    al = fma(x, x, al);
                    be = fma(v, v, be);
                    ga = fma(x, v, ga);
                }
                al = warp_sum(al);
                be = warp_sum(be);
                ga = warp_sum(ga);
                if (ga == 0.0 || fabs(ga) <= DBL_EPSILON * sqrt(al * be)) continue;
                ++rotated;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int i = lane; i < F; i += 32) {
                    const double x = ap[i], v = aq[i];
                    ap[i] = cs * x - sn * v;
                    aq[i] = sn * x + cs * v;
                }
                double *jp = Jr + pi * n, *jq = Jr + qi * n;
                for (int64_t i = lane; i < n; i += 32) {
                    const double x = jp[i], v = jq[i];
                    jp[i] = cs * x - sn * v;
                    jq[i] = sn * x + cs * v;
                }
                __syncwarp();
            }
        if (rotated == 0) break;
    }
    double smax = 0.0;
    for (int64_t c = 0; c < n; ++c) {
        double s2 = 0.0;
        for (int i = lane; i < F; i += 32) s2 = fma(A[c * F + i], A[c * F + i], s2);
        smax = fmax(smax, sqrt(warp_sum(s2)));
    }
    const int64_t mx = (n > F) ? n : F;
    const double cutoff = p.svd_ridge ? ((p.rcond == p.rcond) ? p.rcond : DBL_EPSILON * static_cast<double>(mx)) * smax
                                      : DBL_EPSILON * smax;
    double *beta = p.beta + g * F;
    for (int i = lane; i < F; i += 32) beta[i] = 0.0;
    __syncwarp();
    for (int64_t c = 0; c < n; ++c) {
        double s2 = 0.0, jy = 0.0;
        for (int i = lane; i < F; i += 32) s2 = fma(A[c * F + i], A[c * F + i], s2);
        for (int64_t i = lane; i < n; i += 32) jy = fma(Jr[c * n + i], y[i], jy);
        s2 = warp_sum(s2);
        jy = warp_sum(jy);
        const double sv = sqrt(s2);
        double coef;
        if (p.svd_ridge) coef = (sv < cutoff) ? 0.0 : jy / (s2 + p.alpha);
        else coef = (sv <= cutoff) ? 0.0 : jy / s2;
        for (int i = lane; i < F; i += 32) beta[i] = fma(A[c * F + i], coef, beta[i]);
    }
    if (lane == 0) p.flags[g] = (fl & ~(FLAG_ILLCOND | FLAG_LU_FALLBACK | FLAG_QR)) | FLAG_SVD;
}

// predictions / residuals: one thread per packed row, features read through the tables
__global__ void __launch_bounds__(256) big_predict_kernel(const BigParams p) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t N = p.n_rows;
    const int F = p.F, kd = p.kd;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < N; r += stride) {
        int64_t lo = 0, hi = p.n_groups;  // group of row r: largest g with group_off[g] <= r
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (p.group_off[mid] <= r) lo = mid; else hi = mid;
        }
        const double *beta = p.beta + lo * F;
        const int64_t src = p.row_index ? p.row_index[r] : r;
        const double s = big_sqrt_w(p, src);
        const double fillv = (p.fill == PREP_NAN) ? NAN : 0.0;
        double acc = 0.0;
        bool all_valid = true;
        for (int c = 0; c < kd; ++c) {
            const bool ok = p.valid[c] ? bit_at(p.valid[c], src) : true;
            all_valid = all_valid && ok;
            acc = fma(big_scale(ok ? big_load(p.vals[c], src, p.f32) : fillv, s, p), beta[c], acc);
        }
        if (p.intercept) acc = fma(s, beta[kd], acc);
        if (p.has_w) acc *= p.f32 ? static_cast<double>(1.0f / static_cast<float>(s)) : 1.0 / s;  // predictions *= 1 / sqrt_w
        const bool y_ok = p.valid[kd] ? bit_at(p.valid[kd], src) : true;
        bool valid = true;
        if (p.drop_mask) valid = all_valid && y_ok;  // policy "drop": null where any input is null
        if (p.residuals) {
            acc = big_load(p.vals[kd], src, p.f32) - acc;
            valid = valid && y_ok;
        }
        p.out[src] = acc;
        if (p.out_valid) p.out_valid[src] = valid ? 1 : 0;
    }
}

}  // namespace b200
