// moving_wide.cuh — rolling_least_squares / recursive_least_squares for 9 <= k <= 64 coefficients: ONE THREAD BLOCK
// per (series, time chunk), the k x k state in shared memory, every step of the recurrence spread over the block.
//
// The reference runs any k sequentially and switches the rolling solver to a Woodbury rank-2 update of (X^T X)^-1
// above 60 coefficients (src/least_squares.rs:737-787, :863); that is the same beta = (X^T X)^-1 X^T y as the direct
// solve (NonWoodburyState, :700-735), so the device always solves S beta = v by Cholesky (LU fallback, :732-734) and
// `use_woodbury` only selects the reference's arithmetic, not the result.  Same chunk algorithms as the register
// kernels (rolling_chunk_impl / rls restart from the information form, moving_core.cuh), same null semantics.
// This is the general path: correct for every policy and k <= 64, not tuned (k^3 / 3 flops per row and two block
// barriers per Cholesky column).
#pragma once
#include "moving.cuh"

namespace b200 {

constexpr int MWIDE_THREADS = 128;

struct WideSmem {
    double *S;     // [K*K] window sums (lower triangle authoritative) | rls: P (full)
    double *Lw;    // [K*K] Cholesky factor / LU scratch
    double *v;     // [K]
    double *beta;  // [K]
    double *xs;    // [K+1] current row: scaled features, then scaled target
    double *t;     // [K]   substitution scratch | rls: P x
    double *inv;   // [K]   1 / L[j][j]
    double *theta; // [K]   rls coefficients
    double *kg;    // [K]   rls gain
};

__host__ __device__ inline size_t moving_wide_smem(int K) {
    return (static_cast<size_t>(2) * K * K + 8 * static_cast<size_t>(K) + 8) * sizeof(double);
}

__device__ __forceinline__ WideSmem wide_carve(unsigned char *raw, int K) {
    WideSmem s;
    double *p = reinterpret_cast<double *>(raw);
    s.S = p; p += K * K;
    s.Lw = p; p += K * K;
    s.v = p; p += K;
    s.beta = p; p += K;
    s.xs = p; p += K + 1;
    s.t = p; p += K;
    s.inv = p; p += K;
    s.theta = p; p += K;
    s.kg = p;
    return s;
}

// scaled row r into sm.xs[0..K] (features incl. intercept, then the target); returns the row's scale and raw target
template <typename T>
__device__ __forceinline__ void wide_load_row(const MovingParams &p, const WideSmem &sm, int K, int64_t r, T *scale_out = nullptr,
                                              T *y_raw_out = nullptr) {
    T s = T(1);
    if (p.w) {
        const T wv = static_cast<const T *>(p.w)[r];
        s = p.w_is_sqrt ? wv : static_cast<T>(sqrt(wv));
    }
    const int j = threadIdx.x;
    if (j < K) sm.xs[j] = (j < p.kd) ? static_cast<double>(static_cast<T>(static_cast<const T *>(p.cols[j])[r] * s)) : static_cast<double>(s);
    const T yr = static_cast<const T *>(p.cols[p.kd])[r];
    if (j == K) sm.xs[K] = static_cast<double>(static_cast<T>(yr * s));
    if (scale_out) *scale_out = s;
    if (y_raw_out) *y_raw_out = yr;
    __syncthreads();
}

// Lw <- Cholesky factor of S (strict lower part + inv[j] = 1 / L[j][j]); false on a non-positive pivot (uniform)
__device__ inline bool wide_factor(const WideSmem &sm, int K) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int i = ty; i < K; i += MWIDE_THREADS / 16)
        for (int j = tx; j < K; j += 16) sm.Lw[i * K + j] = (j <= i) ? sm.S[i * K + j] : sm.S[j * K + i];
    __syncthreads();
    for (int j = 0; j < K; ++j) {
        const double d = sm.Lw[j * K + j];
        if (!(d > 0.0)) return false;  // every thread reads the same pivot
        const double rinv = rsqrt(d);
        for (int i = j + 1 + tid; i < K; i += MWIDE_THREADS) sm.Lw[i * K + j] *= rinv;
        if (tid == 0) sm.inv[j] = rinv;
        __syncthreads();
        for (int ii = j + 1 + ty; ii < K; ii += MWIDE_THREADS / 16) {
            const double lij = sm.Lw[ii * K + j];
            for (int kk = j + 1 + tx; kk <= ii; kk += 16) sm.Lw[ii * K + kk] = fma(-lij, sm.Lw[kk * K + j], sm.Lw[ii * K + kk]);
        }
        __syncthreads();
    }
    return true;
}

// out = (L L^T)^-1 v by warp 0 (column-oriented substitutions, lanes own rows); ends with a block barrier
__device__ inline void wide_substitute(const WideSmem &sm, int K, double *out) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int i = lane; i < K; i += 32) sm.t[i] = sm.v[i];
        __syncwarp();
        for (int pcol = 0; pcol < K; ++pcol) {  // forward: z = L^-1 v
            const double z = sm.t[pcol] * sm.inv[pcol];
            __syncwarp();
            if (lane == 0) sm.t[pcol] = z;
            for (int i = pcol + 1 + lane; i < K; i += 32) sm.t[i] = fma(-sm.Lw[i * K + pcol], z, sm.t[i]);
            __syncwarp();
        }
        for (int pcol = K - 1; pcol >= 0; --pcol) {  // backward: out = L^-T z
            const double bv = sm.t[pcol] * sm.inv[pcol];
            __syncwarp();
            if (lane == 0) out[pcol] = bv;
            for (int i = lane; i < pcol; i += 32) sm.t[i] = fma(-sm.Lw[pcol * K + i], bv, sm.t[i]);
            __syncwarp();
        }
    }
    __syncthreads();
}

// beta = S^-1 v: Cholesky, or on a non-positive pivot LU with partial pivoting on one thread (the reference's
// fallback, src/least_squares.rs:732-734)
__device__ inline void wide_solve(const WideSmem &sm, int K) {
    const int tid = threadIdx.x;
    if (wide_factor(sm, K)) {
        wide_substitute(sm, K, sm.beta);
        return;
    }
    __syncthreads();
    for (int e = tid; e < K * K; e += MWIDE_THREADS) {
        const int i = e / K, j = e - i * K;
        sm.Lw[e] = (j <= i) ? sm.S[e] : sm.S[j * K + i];
    }
    if (tid < K) sm.beta[tid] = sm.v[tid];
    __syncthreads();
    if (tid == 0) lu_solve_inplace(sm.Lw, K, K, sm.beta);
    __syncthreads();
}

// rows of beta / predictions for row r (rules of DevEmit, moving.cuh)
template <typename T>
__device__ inline void wide_emit(const MovingParams &p, const WideSmem &sm, const double *beta, int K, int64_t r, bool is_nan, bool row_valid) {
    const int tid = threadIdx.x;
    const int64_t orow = p.row_index ? p.row_index[r] : r;
    if (p.mode == 2) {
        if (tid < K) {
            const double bv = is_nan ? NAN : beta[tid];
            p.out[orow * K + tid] = bv;
            if (p.out_valid) p.out_valid[orow * K + tid] = (bv == bv) ? 1 : 0;
        }
        return;
    }
    T s, y_raw;
    wide_load_row<T>(p, sm, K, r, &s, &y_raw);
    if (tid < 32) {
        double acc = 0.0;
        for (int j = tid; j < K; j += 32) acc = fma(sm.xs[j], is_nan ? NAN : beta[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (tid == 0) {
            double pred = acc;
            if (p.w) pred *= static_cast<double>(T(1) / s);
            bool valid = true;
            if (p.mask_predictions) valid = row_valid;
            if (p.mode == 1) {
                const int64_t trow = p.target_is_packed ? r : orow;
                pred = static_cast<double>(static_cast<const T *>(p.target)[trow]) - pred;
                if (p.target_validity) valid = valid && ((p.target_validity[orow >> 3] >> (orow & 7)) & 1);
            }
            if (p.kind == MOVING_ROLLING) valid = valid && (pred == pred);
            p.out[orow] = pred;
            if (p.out_valid) p.out_valid[orow] = valid ? 1 : 0;
        }
    }
    __syncthreads();
}

// block-cooperative backend of rolling_chunk_impl
template <typename T>
struct RollingWide {
    const MovingParams &p;
    WideSmem sm;
    int K;
    bool nan_state;
    __device__ bool valid(int64_t r) const { return p.mask ? (static_cast<const T *>(p.mask)[r] != T(0)) : true; }
    __device__ void prefetch(int64_t) const {}
    __device__ void clear() {
        for (int e = threadIdx.x; e < K * K; e += MWIDE_THREADS) sm.S[e] = 0.0;
        if (threadIdx.x < K) sm.v[threadIdx.x] = 0.0;
        __syncthreads();
    }
    __device__ void add_row(int64_t r, double sign) {
        wide_load_row<T>(p, sm, K, r);
        for (int i = threadIdx.x >> 4; i < K; i += MWIDE_THREADS / 16) {
            const double xi = sign * sm.xs[i];
            for (int j = threadIdx.x & 15; j <= i; j += 16) sm.S[i * K + j] = fma(xi, sm.xs[j], sm.S[i * K + j]);
        }
        if (threadIdx.x < K) sm.v[threadIdx.x] = fma(sign * sm.xs[threadIdx.x], sm.xs[K], sm.v[threadIdx.x]);
        __syncthreads();
    }
    __device__ void add_diag(double a) {
        if (threadIdx.x < K) sm.S[threadIdx.x * K + threadIdx.x] += a;
        __syncthreads();
    }
    __device__ void solve() {
        wide_solve(sm, K);
        nan_state = false;
    }
    __device__ void set_nan() { nan_state = true; }
    __device__ void emit(int64_t r, bool is_nan) { wide_emit<T>(p, sm, sm.beta, K, r, is_nan || nan_state, valid(r)); }
};

template <typename T>
__global__ void __launch_bounds__(MWIDE_THREADS) rolling_wide_kernel(const MovingParams p) {
    extern __shared__ __align__(16) unsigned char wide_raw[];
    const int K = p.F;
    const int64_t c = blockIdx.x;
    const int64_t g = p.chunk_group[c];
    const int64_t g0 = p.group_off[g], g1 = p.group_off[g + 1];
    RollingSeries rs;
    if (p.series_info) {
        rs.mpv = p.series_info[g * 4 + 0];
        rs.n_valid = p.series_info[g * 4 + 1];
        rs.all_nan = static_cast<int>(p.series_info[g * 4 + 2]);
        rs.m_warm = p.series_info[g * 4 + 3];
    } else {
        rs.mpv = p.min_periods;
        rs.n_valid = (g1 - g0 < p.min_periods) ? (g1 - g0) : p.min_periods;
        rs.all_nan = (g1 - g0) < p.min_periods;
        rs.m_warm = p.min_periods;
    }
    RollingCfg cfg{p.window, p.min_periods, p.alpha, p.fixed_window};
    RollingWide<T> b{p, wide_carve(wide_raw, K), K, true};
    rolling_chunk_impl(b, cfg, rs, g0, g1, p.chunk_r0[c], p.chunk_r1[c]);
}

// one thread per series (any k): min_periods_valid etc., as rolling_prepass_kernel
template <typename T>
__global__ void __launch_bounds__(128) rolling_wide_prepass_kernel(const MovingParams p) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= p.n_groups) return;
    struct MaskSrc {
        const T *mask;
        __device__ bool valid(int64_t r) const { return mask ? (mask[r] != T(0)) : true; }
    } src{static_cast<const T *>(p.mask)};
    const RollingSeries rs = rolling_prepass(src, p.group_off[g], p.group_off[g + 1], p.min_periods);
    p.series_info[g * 4 + 0] = rs.mpv;
    p.series_info[g * 4 + 1] = rs.n_valid;
    p.series_info[g * 4 + 2] = rs.all_nan;
    p.series_info[g * 4 + 3] = rs.m_warm;
}

// ---- recursive least squares ---------------------------------------------------------------------------------
// pass 1: information-form summary of a chunk -> summaries[c] = { A (K*K, lower triangle), b (K), D }
template <typename T>
__global__ void __launch_bounds__(MWIDE_THREADS) rls_wide_summary_kernel(const MovingParams p) {
    extern __shared__ __align__(16) unsigned char wide_raw[];
    const int K = p.F;
    const WideSmem sm = wide_carve(wide_raw, K);
    const int64_t c = blockIdx.x;
    const int64_t c0 = p.chunk_r0[c], c1 = p.chunk_r1[c];
    const T *mask = static_cast<const T *>(p.mask);
    for (int e = threadIdx.x; e < K * K; e += MWIDE_THREADS) sm.S[e] = 0.0;
    if (threadIdx.x < K) sm.v[threadIdx.x] = 0.0;
    __syncthreads();
    double D = 1.0;
    const double lam = p.lambda;
    for (int64_t r = c0; r < c1; ++r) {
        if (mask && mask[r] == T(0)) continue;
        wide_load_row<T>(p, sm, K, r);
        for (int i = threadIdx.x >> 4; i < K; i += MWIDE_THREADS / 16) {
            const double xi = sm.xs[i];
            for (int j = threadIdx.x & 15; j <= i; j += 16) sm.S[i * K + j] = fma(xi, sm.xs[j], sm.S[i * K + j] * lam);
        }
        if (threadIdx.x < K) sm.v[threadIdx.x] = fma(sm.xs[threadIdx.x], sm.xs[K], sm.v[threadIdx.x] * lam);
        D *= lam;
        __syncthreads();
    }
    double *rec = p.summaries + c * moving_rec(K);
    for (int e = threadIdx.x; e < K * K; e += MWIDE_THREADS) {
        const int i = e / K, j = e - i * K;
        rec[e] = (j <= i) ? sm.S[e] : 0.0;
    }
    if (threadIdx.x < K) rec[K * K + threadIdx.x] = sm.v[threadIdx.x];
    if (threadIdx.x == 0) rec[K * K + K] = D;
}

// pass 2: exclusive scan over the chunks of a series (one block per series, threads = matrix elements):
// summaries[c] <- information state ENTERING chunk c, starting from the prior (or init_info); state_out as rls_scan_kernel
__global__ void __launch_bounds__(256) rls_wide_scan_kernel(const MovingParams p, const int64_t *__restrict__ group_chunk_off) {
    const int K = p.F;
    const int NE = K * K + K;
    const int64_t g = blockIdx.x;
    const int64_t ca = group_chunk_off[g], cb = group_chunk_off[g + 1];
    for (int e = threadIdx.x; e < NE; e += blockDim.x) {
        double carry;
        if (p.init_info) {
            carry = (e < K * K) ? (((e % K) <= (e / K)) ? p.init_info[g * NE + e] : 0.0) : p.init_info[g * NE + e];
        } else if (e < K * K) {
            carry = ((e / K) == (e % K)) ? 1.0 / p.p0 : 0.0;
        } else {
            carry = (p.has_mean ? p.mean[e - K * K] : 0.0) / p.p0;
        }
        for (int64_t c = ca; c < cb; ++c) {
            double *rec = p.summaries + c * moving_rec(K);
            const double D = rec[NE], add = rec[e];
            rec[e] = carry;
            carry = fma(D, carry, add);
        }
        if (p.state_out) p.state_out[g * (NE + 1) + e] = carry;
    }
    if (p.state_out && threadIdx.x == 0) {
        double Dall = 1.0;
        for (int64_t c = ca; c < cb; ++c) Dall *= p.summaries[c * moving_rec(K) + NE];
        p.state_out[g * (NE + 1) + NE] = Dall;
    }
}

// pass 3: covariance-form recurrence of a chunk (src/least_squares.rs:531-540), P in shared memory.
// P stays exactly symmetric (P/lambda - K_i K_j r is symmetric term by term), so P x == (x^T P)^T bit for bit and
// one block-wide product serves both.
template <typename T>
__global__ void __launch_bounds__(MWIDE_THREADS) rls_wide_main_kernel(const MovingParams p) {
    extern __shared__ __align__(16) unsigned char wide_raw[];
    const int K = p.F;
    const WideSmem sm = wide_carve(wide_raw, K);
    const int tid = threadIdx.x;
    const int64_t c = blockIdx.x;
    const int64_t g = p.chunk_group[c];
    const int64_t c0 = p.chunk_r0[c], c1 = p.chunk_r1[c];
    const bool first = (c0 == p.group_off[g]) && !p.init_info;
    double *P = sm.S, *theta = sm.theta, *px = sm.t, *kg = sm.kg;
    if (first) {
        for (int e = tid; e < K * K; e += MWIDE_THREADS) P[e] = ((e / K) == (e % K)) ? p.p0 : 0.0;
        if (tid < K) theta[tid] = p.has_mean ? p.mean[tid] : 0.0;
        __syncthreads();
    } else {
        // information state (A, b) entering the chunk (from the scan) -> theta = A^-1 b, P = A^-1: one block-cooperative
        // Cholesky, then k + 1 substitutions; the columns of P overwrite A, which the factor no longer needs
        const double *rec = p.summaries + c * moving_rec(K);
        for (int e = tid; e < K * K; e += MWIDE_THREADS) sm.S[e] = rec[e];
        if (tid < K) sm.v[tid] = rec[K * K + tid];
        __syncthreads();
        const bool ok = wide_factor(sm, K);  // A = prior + PSD terms: positive definite unless the data hold NaN / inf
        if (ok) {
            wide_substitute(sm, K, theta);
            for (int col = 0; col < K; ++col) {
                if (tid < K) sm.v[tid] = (tid == col) ? 1.0 : 0.0;
                __syncthreads();
                wide_substitute(sm, K, sm.beta);
                if (tid < K) P[tid * K + col] = sm.beta[tid];
                __syncthreads();
            }
        } else {
            __syncthreads();
            for (int e = tid; e < K * K; e += MWIDE_THREADS) P[e] = NAN;
            if (tid < K) theta[tid] = NAN;
            __syncthreads();
        }
    }
    const T *mask = static_cast<const T *>(p.mask);
    const double lam = p.lambda, inv_lam = 1.0 / p.lambda;
    int64_t exact_left = first ? RLS_EXACT_ROWS : 0;
    for (int64_t r = c0; r < c1; ++r) {
        const bool rv = mask ? (mask[r] != T(0)) : true;
        if (rv) {
            const bool exact = exact_left > 0;
            if (exact) --exact_left;
            wide_load_row<T>(p, sm, K, r);
            if (tid < K) {  // (x^T P)_j = sum_i x_i P[i][j], i ascending as the reference's dot
                double s = 0.0;
                if (exact) for (int i = 0; i < K; ++i) s = B200_ADD(s, B200_MUL(sm.xs[i], P[i * K + tid]));
                else for (int i = 0; i < K; ++i) s = fma(sm.xs[i], P[i * K + tid], s);
                px[tid] = s;
            }
            __syncthreads();
            double q = 0.0, pred = 0.0;  // every thread forms the two scalars itself (same order everywhere)
            if (exact) {
                for (int j = 0; j < K; ++j) q = B200_ADD(q, B200_MUL(px[j], sm.xs[j]));
                for (int j = 0; j < K; ++j) pred = B200_ADD(pred, B200_MUL(sm.xs[j], theta[j]));
            } else {
                for (int j = 0; j < K; ++j) q = fma(px[j], sm.xs[j], q);
                for (int j = 0; j < K; ++j) pred = fma(sm.xs[j], theta[j], pred);
            }
            const double rr = exact ? B200_ADD(1.0, B200_DIV(q, lam)) : 1.0 + q / lam;
            const double resid = sm.xs[K] - pred;
            __syncthreads();
            if (tid < K) {
                const double kv = exact ? B200_DIV(px[tid], B200_MUL(rr, lam)) : px[tid] * (1.0 / (rr * lam));
                kg[tid] = kv;
                theta[tid] = exact ? B200_ADD(theta[tid], B200_MUL(kv, resid)) : fma(kv, resid, theta[tid]);
            }
            __syncthreads();
            for (int e = tid; e < K * K; e += MWIDE_THREADS) {
                const int i = e / K, j = e - i * K;
                P[e] = exact ? B200_ADD(B200_DIV(P[e], lam), -B200_MUL(B200_MUL(kg[i], kg[j]), rr))
                             : fma(P[e], inv_lam, -(kg[i] * kg[j]) * rr);
            }
            __syncthreads();
        }
        wide_emit<T>(p, sm, theta, K, r, false, rv);
    }
}

// host launcher (moving_wide.cu)
template <typename T>
static cudaError_t launch_moving_wide_t(cudaStream_t stream, MovingParams &p, const int64_t *group_chunk_off_dev, int64_t *launches) {
    if (p.n_chunks == 0) return cudaSuccess;
    const size_t smem = moving_wide_smem(p.F);
    const unsigned grid = static_cast<unsigned>(p.n_chunks);
    if (p.kind == MOVING_ROLLING) {
        cudaFuncSetAttribute(rolling_wide_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (p.mask) {
            rolling_wide_prepass_kernel<T><<<static_cast<unsigned>((p.n_groups + 127) / 128), 128, 0, stream>>>(p);
            ++*launches;
        } else {
            p.series_info = nullptr;
        }
        rolling_wide_kernel<T><<<grid, MWIDE_THREADS, smem, stream>>>(p);
        ++*launches;
        return cudaGetLastError();
    }
    cudaFuncSetAttribute(rls_wide_summary_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(rls_wide_main_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    rls_wide_summary_kernel<T><<<grid, MWIDE_THREADS, smem, stream>>>(p);
    rls_wide_scan_kernel<<<static_cast<unsigned>(p.n_groups), 256, 0, stream>>>(p, group_chunk_off_dev);
    *launches += 2;
    if (p.state_only) return cudaGetLastError();
    rls_wide_main_kernel<T><<<grid, MWIDE_THREADS, smem, stream>>>(p);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace b200
