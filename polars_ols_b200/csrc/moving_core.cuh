// moving_core.cuh — time-parallel recursive / rolling least squares, one THREAD per time chunk with the
// whole k x k state in registers (k <= 8).  Host/device templates: the kernels in moving.cuh and the CPU
// host-check harness (tests/hostcheck) instantiate exactly this code.
//
// Reference recurrences restated (files relative to /root/reference):
//   RecursiveLeastSquares::update           src/least_squares.rs:531-540
//   solve_recursive_least_squares           src/least_squares.rs:568-598   (forward fill on invalid rows)
//   solve_rolling_ols                       src/least_squares.rs:848-1032  (warm-up, both null branches)
//   NonWoodburyState::{update,subtract,solve} src/least_squares.rs:700-735 (Cholesky -> LU per row)
// The reference walks each series strictly sequentially.  Here a series is cut into chunks that run
// concurrently; what a chunk needs from its past is rebuilt exactly:
//   rolling : the window sums at the chunk start = a direct Gram over the previous W (valid) rows;
//   rls     : information form A_t = lam A_{t-1} + x x^T, b_t = lam b_{t-1} + x y is associative, so a
//             per-chunk summary pass + a scan over chunks gives (A, b) entering every chunk and
//             P = A^-1, theta = P b restart the covariance-form recurrence (equal to the sequential
//             result to rounding: SURVEY.md §7 hard part 2, verified in tests/test_hostcheck.py).
#pragma once
#include <cmath>
#include <cstdint>

#include "solvers.cuh"

namespace b200 {

// prefetch hint `MOVING_PF` rows ahead of a chunk walk (device sources only; see DevSrcT::prefetch in moving.cuh)
constexpr int MOVING_PF = 4;
#if defined(__CUDA_ARCH__)
#define B200_PREFETCH(src, r) (src).prefetch(r)
#else
#define B200_PREFETCH(src, r) ((void)0)
#endif

constexpr int MOVING_MAX_K = 64;  // K <= 8: one thread per chunk, state in registers; 9..64: one block per chunk, state in shared memory (moving_wide.cuh)
enum : int { MOVING_RLS = 0, MOVING_ROLLING = 1 };

// ---- register-resident small matrices --------------------------------------------------------------
template <int K>
struct NormalState {   // S = sum x x^T (+ alpha I), v = sum x y   (lower triangle of S is authoritative)
    double S[K][K];
    double v[K];
    B200_HD void clear() {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            v[i] = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) S[i][j] = 0.0;
        }
    }
    B200_HD void add(const double (&x)[K], double y, double sign) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const double xi = sign * x[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) S[i][j] = fma(xi, x[j], S[i][j]);
            v[i] = fma(xi, y, v[i]);
        }
    }
    B200_HD void decay(double lam) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            v[i] *= lam;
#pragma unroll
            for (int j = 0; j <= i; ++j) S[i][j] *= lam;
        }
    }
    B200_HD void add_diag(double a) {
#pragma unroll
        for (int i = 0; i < K; ++i) S[i][i] += a;
    }
};

// beta = S^-1 v by Cholesky (register resident); on a non-positive pivot fall back to LU with partial
// pivoting on a local copy (solve_normal_equations(.., None, Some(LU)), src/least_squares.rs:732-734).
template <int K>
B200_HD void solve_normal(const NormalState<K> &st, double (&beta)[K]) {
    double L[K][K], inv[K];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double d = st.S[j][j];
#pragma unroll
        for (int p = 0; p < j; ++p) d = fma(-L[j][p], L[j][p], d);
        if (!(d > 0.0)) ok = false;
        // one reciprocal square root per column instead of K divisions (f64 division and sqrt are ~25-instruction
        // sequences on the GPU and this runs once per row); results agree with the reference's LL^T to rounding
        // one reciprocal square root per column instead of K divisions; a hand-rolled MUFU.RSQ + two Newton steps was
        // measured against the library routine on C4 (profiles/r02_c4_variants.json): no difference, the library one stays
#if defined(__CUDA_ARCH__)
        const double r = rsqrt(d);
#else
        const double r = 1.0 / sqrt(d);
#endif
        inv[j] = r;
        L[j][j] = d * r;
#pragma unroll
        for (int i = j + 1; i < K; ++i) {
            double s = st.S[i][j];
#pragma unroll
            for (int p = 0; p < j; ++p) s = fma(-L[i][p], L[j][p], s);
            L[i][j] = s * r;
        }
    }
    if (ok) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            double s = st.v[i];
#pragma unroll
            for (int p = 0; p < i; ++p) s = fma(-L[i][p], beta[p], s);
            beta[i] = s * inv[i];
        }
#pragma unroll
        for (int i = K - 1; i >= 0; --i) {
            double s = beta[i];
#pragma unroll
            for (int p = i + 1; p < K; ++p) s = fma(-L[p][i], beta[p], s);
            beta[i] = s * inv[i];
        }
        return;
    }
    double A[K * K], b[K];
    for (int i = 0; i < K; ++i) {
        b[i] = st.v[i];
        for (int j = 0; j < K; ++j) A[i * K + j] = (j <= i) ? st.S[i][j] : st.S[j][i];
    }
    lu_solve_inplace(A, K, K, b);
    for (int i = 0; i < K; ++i) beta[i] = b[i];
}

// ---- row access ------------------------------------------------------------------------------------
// A RowSource provides, for a packed row index r of the series:
//   bool valid(r)                      row enters the fit
//   void load(r, double (&x)[K], double &y)   scaled features (incl. intercept) and target, f64
template <int K, typename Src>
B200_HD void gram_range(const Src &src, int64_t a, int64_t b, NormalState<K> &st) {
    double x[K], y;
    for (int64_t r = a; r < b; ++r)
        if (src.valid(r)) {
            src.load(r, x, y);
            st.add(x, y, 1.0);
        }
}

// ---- rolling ---------------------------------------------------------------------------------------
struct RollingCfg {
    int64_t window;       // W
    int64_t min_periods;  // resolved (>= 1)
    double alpha;         // >= 0, added once to the diagonal (src/least_squares.rs:924-926)
    int fixed_window;     // 0: "last W valid rows" (drop*), 1: fixed row window (drop_window / zero / ignore)
};

// Per-series facts computed once (rolling_prepass): position of the min_periods-th valid row etc.
struct RollingSeries {
    int64_t mpv;      // min_periods_valid (src/least_squares.rs:881-891); rows < mpv-1 are NaN
    int64_t n_valid;  // the reference's `n_valid` after that loop (= min_periods when reached)
    int all_nan;      // n < max(n_valid, min_periods)  (:893-900)
    int64_t m_warm;   // valid rows inside the warm-up [g0, g0 + mpv)
};

template <typename Src>
B200_HD RollingSeries rolling_prepass(const Src &src, int64_t g0, int64_t g1, int64_t min_periods) {
    RollingSeries rs;
    rs.mpv = min_periods;
    rs.n_valid = 0;
    bool reached = false;
    for (int64_t i = g0; i < g1; ++i) {
        if (src.valid(i)) rs.n_valid += 1;
        if (rs.n_valid == min_periods) {
            rs.mpv = (i - g0) + 1;
            reached = true;
            break;
        }
    }
    const int64_t n = g1 - g0;
    rs.all_nan = n < ((rs.n_valid > min_periods) ? rs.n_valid : min_periods);
    rs.m_warm = rs.n_valid;
    if (!reached) {  // fewer than min_periods valid rows in the whole series: the warm-up is the first min_periods ROWS
        rs.m_warm = 0;
        for (int64_t i = g0; i < g0 + rs.mpv && i < g1; ++i) rs.m_warm += src.valid(i) ? 1 : 0;
    }
    return rs;
}

// Register backend of rolling_chunk_impl: one thread owns the chunk, window sums and Cholesky factor in registers.
template <int K, typename Src, typename Emit>
struct RollingRegs {
    const Src &src;
    Emit &sink;
    NormalState<K> st;
    double beta[K];
    B200_HD bool valid(int64_t r) const { return src.valid(r); }
    B200_HD void prefetch(int64_t r) const { B200_PREFETCH(src, r); (void)r; }
    B200_HD void clear() { st.clear(); }
    B200_HD void add_row(int64_t r, double sign) {
        double x[K], y;
        src.load(r, x, y);
        st.add(x, y, sign);
    }
    B200_HD void add_diag(double a) { st.add_diag(a); }
    B200_HD void solve() { solve_normal<K>(st, beta); }
    B200_HD void set_nan() {
#pragma unroll
        for (int j = 0; j < K; ++j) beta[j] = NAN;
    }
    B200_HD void emit(int64_t r, bool is_nan) { sink(r, beta, is_nan); }
};

// Processes rows [c0, c1) of the series [g0, g1) (packed indices); b.emit(r, is_nan) is called once per row in
// increasing r.  `B` is the execution backend (state + row access + solve + output): RollingRegs for one thread
// per chunk, the block-cooperative backend of moving_wide.cuh for 9 <= k <= 64; every call on it is made in the
// same order by all threads of a cooperative backend (the control flow only depends on validity bits).
template <typename B>
B200_HD void rolling_chunk_impl(B &b, const RollingCfg &cfg, const RollingSeries &rs, int64_t g0, int64_t g1,
                                int64_t c0, int64_t c1) {
    b.set_nan();
    if (rs.all_nan) {
        for (int64_t r = c0; r < c1; ++r) b.emit(r, true);
        return;
    }
    const int64_t W = cfg.window;
    const int64_t first = g0 + rs.mpv - 1;  // first row that carries coefficients
    int64_t r = c0;
    for (; r < c1 && r < first; ++r) b.emit(r, true);
    if (r >= c1) return;
    b.clear();

    if (!cfg.fixed_window) {
        // ---- branch 0: window = the reference's deque of valid row indices (src/least_squares.rs:903-986) ----
        // The deque starts as the first min(m_warm, W) valid warm-up rows (:917-919) and every later valid row is
        // appended after the front was popped (and subtracted) once W entries are held.  Valid warm-up rows beyond
        // the W-th never enter the deque, so they are never subtracted ("stuck", only when min_periods > window).
        const int64_t wend = g0 + rs.mpv;  // first row after the warm-up
        const int64_t ninit = rs.m_warm < W ? rs.m_warm : W;
        int64_t cnt = 0, tail = wend, init_left = 0;
        if (r == first) {
            // warm-up (:909-921): every valid row of [g0, wend) enters the sums, the first W of them the deque
            for (int64_t i = g0; i < wend; ++i)
                if (b.valid(i)) {
                    b.add_row(i, 1.0);
                    if (cnt == 0) tail = i;
                    if (cnt < W) ++cnt;
                }
            init_left = cnt;
            if (cfg.alpha > 0.0) b.add_diag(cfg.alpha);
            b.solve();
            b.emit(r, false);
            ++r;
        } else {
            // state entering row r: the newest post-warm-up valid rows before r (at most W) ...
            int64_t post = 0;
            for (int64_t i = r - 1; i >= wend && post < W; --i)
                if (b.valid(i)) {
                    if (i - MOVING_PF >= wend) b.prefetch(i - MOVING_PF);
                    b.add_row(i, 1.0);
                    tail = i;
                    ++post;
                }
            cnt = post;
            if (post < W) {
                // ... and the warm-up entries still in the deque: ninit + post entries were pushed so far, so
                // max(0, ninit + post - W) were popped from the front
                const int64_t pushed = ninit + post;
                const int64_t popped = pushed > W ? pushed - W : 0;
                int64_t rank = 0;
                for (int64_t i = g0; i < wend && rank < ninit; ++i)
                    if (b.valid(i)) {
                        if (rank >= popped) {
                            b.add_row(i, 1.0);
                            if (init_left == 0) tail = i;
                            ++init_left;
                        }
                        ++rank;
                    }
                cnt += init_left;
            }
            if (rs.m_warm > W) {  // stuck warm-up rows
                int64_t rank = 0;
                for (int64_t i = g0; i < wend; ++i)
                    if (b.valid(i)) {
                        if (rank >= W) b.add_row(i, 1.0);
                        ++rank;
                    }
            }
            if (cfg.alpha > 0.0) b.add_diag(cfg.alpha);
            b.solve();  // coefficients carried into the chunk (forward fill source)
        }
        for (; r < c1; ++r) {
            if (r + MOVING_PF < c1) b.prefetch(r + MOVING_PF);
            if (b.valid(r)) {
                if (cnt == W) {  // saturated: pop + subtract the deque's front
                    b.prefetch(tail + MOVING_PF);  // W >= 1 rows behind r + MOVING_PF: always inside the series
                    b.add_row(r, 1.0);
                    b.add_row(tail, -1.0);
                    if (init_left > 0 && --init_left == 0) tail = wend - 1;  // next entry: first valid row after the warm-up
                    do { ++tail; } while (!b.valid(tail));                     // terminates at r at the latest
                } else {
                    b.add_row(r, 1.0);
                    if (cnt == 0) tail = r;
                    ++cnt;
                }
                b.solve();
            }
            b.emit(r, false);
        }
        return;
    }

    // ---- branch 1: fixed row window (src/least_squares.rs:987-1029) ---------------------------------
    // series-relative index i = r - g0.  State after processing row i:
    //   S = alpha I + Gram(valid rows in (i-W, i])  [+ valid rows < mpv-W that the reference never
    //   subtracts when the warm-up is longer than the window]
    //   cnt_i = #valid in [max(i-W,0)+1, i]   (index 0 is excluded while i < W: quirk at :990-997)
    const int64_t mpv = rs.mpv;
    auto valid_rel = [&](int64_t i) { return b.valid(g0 + i); };
    auto warmup = [&]() {
        b.clear();
        for (int64_t t = 0; t < mpv; ++t)
            if (valid_rel(t)) b.add_row(g0 + t, 1.0);
        if (cfg.alpha > 0.0) b.add_diag(cfg.alpha);
    };
    int64_t i = r - g0;
    int64_t cnt = 0;
    if (r == first) {
        warmup();
        b.solve();
        b.emit(r, false);
        // cnt after row mpv-1 under the reference's sliding definition
        {
            const int64_t ii = mpv - 1, lo = ((ii >= W) ? ii - W : 0) + 1;
            for (int64_t t = lo; t <= ii; ++t) cnt += valid_rel(t) ? 1 : 0;
        }
        ++r;
        ++i;
    } else {
        // rebuild the state after row i-1 and the coefficients carried into the chunk
        const int64_t ip = i - 1;
        auto build = [&](int64_t at) {
            b.clear();
            // rows of the warm-up [0, mpv) are all present until they are subtracted at step t+W (t+W >= mpv);
            // rows >= mpv are present from their own step.  Row t is subtracted at step t+W iff t+W >= mpv.
            const int64_t lo = (at - W + 1 > 0) ? at - W + 1 : 0;
            for (int64_t t = lo; t <= at; ++t)
                if (valid_rel(t)) {
                    if (t + MOVING_PF <= at) b.prefetch(g0 + t + MOVING_PF);
                    b.add_row(g0 + t, 1.0);
                }
            // never-subtracted warm-up rows: t + W < mpv  and t < lo
            const int64_t stuck_hi = (mpv - W < lo) ? mpv - W : lo;
            for (int64_t t = 0; t < stuck_hi; ++t)
                if (valid_rel(t)) b.add_row(g0 + t, 1.0);
            if (cfg.alpha > 0.0) b.add_diag(cfg.alpha);
        };
        {
            const int64_t lo = ((ip >= W) ? ip - W : 0) + 1;
            for (int64_t t = lo; t <= ip; ++t) cnt += valid_rel(t) ? 1 : 0;
        }
        // walk back to the last row whose coefficients were refreshed
        int64_t j = ip, cj = cnt;
        bool found = false;
        while (j >= mpv) {
            const bool vi = valid_rel(j);
            const bool vs = (j >= W) && valid_rel(j - W);
            const bool changed = vi || vs;
            if (changed && cj >= rs.n_valid) {
                found = true;
                break;
            }
            // cnt_{j-1} = cnt_j - valid[j] + (j > W ? valid[j-W] : 0)
            cj = cj - (vi ? 1 : 0) + ((j > W && valid_rel(j - W)) ? 1 : 0);
            --j;
        }
        if (found) build(j); else warmup();  // nothing refreshed since the warm-up: carry the warm-up coefficients
        b.solve();
        build(ip);
    }
    for (; r < c1; ++r, ++i) {
        if (r + MOVING_PF < c1) {
            b.prefetch(r + MOVING_PF);
            if (i + MOVING_PF >= W) b.prefetch(r + MOVING_PF - W);
        }
        const bool vi = valid_rel(i);
        const bool sat = i >= W;
        const bool vs = sat && valid_rel(i - W);
        cnt += (vi ? 1 : 0) - ((i > W && valid_rel(i - W)) ? 1 : 0);
        if (vi) {
            b.add_row(r, 1.0);
            if (vs) b.add_row(r - W, -1.0);
            if (cnt >= rs.n_valid) b.solve();
        } else if (vs) {
            b.add_row(r - W, -1.0);
            if (cnt >= rs.n_valid) b.solve();
        }
        b.emit(r, false);
    }
}

template <int K, typename Src, typename Emit>
B200_HD void rolling_chunk(const Src &src, const RollingCfg &cfg, const RollingSeries &rs, int64_t g0, int64_t g1,
                           int64_t c0, int64_t c1, Emit &emit) {
    RollingRegs<K, Src, Emit> b{src, emit};
    rolling_chunk_impl(b, cfg, rs, g0, g1, c0, c1);
}

// ---- recursive least squares -----------------------------------------------------------------------
struct RlsCfg {
    double lambda;  // forgetting factor (src/least_squares.rs:513-517)
    double p0;      // initial_state_covariance
};

// chunk summary in information form: after the chunk, (A, b)_out = D (A, b)_in + (A_c, b_c)
template <int K>
struct RlsSummary {
    NormalState<K> ab;
    double D;
};

template <int K, typename Src>
B200_HD void rls_summarise(const Src &src, const RlsCfg &cfg, int64_t c0, int64_t c1, RlsSummary<K> &out) {
    out.ab.clear();
    out.D = 1.0;
    double x[K], y;
    for (int64_t r = c0; r < c1; ++r)
        if (src.valid(r)) {
            if (r + MOVING_PF < c1) B200_PREFETCH(src, r + MOVING_PF);
            src.load(r, x, y);
            out.ab.decay(cfg.lambda);
            out.ab.add(x, y, 1.0);
            out.D *= cfg.lambda;
        }
}

// One covariance-form update (src/least_squares.rs:531-540), operation order as in the reference.
template <int K>
B200_HD void rls_update(double (&P)[K][K], double (&theta)[K], const double (&x)[K], double y, double lam, double inv_lam) {
    double xp[K], px[K], kg[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) s = fma(x[i], P[i][j], s);
        xp[j] = s;
    }
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) q = fma(xp[j], x[j], q);
    const double r = 1.0 + q / lam;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) s = fma(P[i][j], x[j], s);
        px[i] = s;
    }
    const double irl = 1.0 / (r * lam);  // one division per row; K = P x / (r lambda)
#pragma unroll
    for (int i = 0; i < K; ++i) kg[i] = px[i] * irl;
    double pred = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) pred = fma(x[j], theta[j], pred);
    const double resid = y - pred;
#pragma unroll
    for (int j = 0; j < K; ++j) theta[j] = fma(kg[j], resid, theta[j]);
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) P[i][j] = fma(P[i][j], inv_lam, -(kg[i] * kg[j]) * r);  // P / lambda - K K^T r
}

// The same update with the reference's literal operation sequence — (x^T P) x, true divisions, no fused multiply-add
// (every product and sum rounded separately, as the scalar CPU code does): used while the recurrence is still
// dominated by the prior.  There P / lambda - (K K^T) r cancels ~p0 |x|^2 digits, so a 1-ulp difference in K is
// amplified by the same factor and the short-cuts of rls_update (one reciprocal, fused products) would show at 1e-6.
#if defined(__CUDA_ARCH__)
#define B200_MUL(a, b) __dmul_rn((a), (b))
#define B200_ADD(a, b) __dadd_rn((a), (b))
#define B200_DIV(a, b) __ddiv_rn((a), (b))
#else
#define B200_MUL(a, b) ((a) * (b))
#define B200_ADD(a, b) ((a) + (b))
#define B200_DIV(a, b) ((a) / (b))
#endif
template <int K>
B200_HD void rls_update_exact(double (&P)[K][K], double (&theta)[K], const double (&x)[K], double y, double lam) {
    double xp[K], px[K], kg[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) s = B200_ADD(s, B200_MUL(x[i], P[i][j]));
        xp[j] = s;
    }
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) q = B200_ADD(q, B200_MUL(xp[j], x[j]));
    const double r = B200_ADD(1.0, B200_DIV(q, lam));
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) s = B200_ADD(s, B200_MUL(P[i][j], x[j]));
        px[i] = s;
    }
    const double rl = B200_MUL(r, lam);
#pragma unroll
    for (int i = 0; i < K; ++i) kg[i] = B200_DIV(px[i], rl);
    double pred = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) pred = B200_ADD(pred, B200_MUL(x[j], theta[j]));
    const double resid = B200_ADD(y, -pred);
#pragma unroll
    for (int j = 0; j < K; ++j) theta[j] = B200_ADD(theta[j], B200_MUL(kg[j], resid));
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) P[i][j] = B200_ADD(B200_DIV(P[i][j], lam), -B200_MUL(B200_MUL(kg[i], kg[j]), r));
}

// rows of a series' first chunk that run rls_update_exact: long enough for the covariance to have contracted from the
// prior to the data's scale (tests: diffuse priors up to 1e6, k up to 8; measured margin in tools/rls_diag.py)
constexpr int64_t RLS_EXACT_ROWS = 64;

// covariance-form state from the information state entering a chunk: P = A^-1 column by column, theta = A^-1 b
// (A is SPD: prior I/p0 plus PSD terms)
template <int K>
B200_HD void rls_restart(NormalState<K> &t, double (&P)[K][K], double (&theta)[K]) {
    solve_normal<K>(t, theta);
#pragma unroll
    for (int c = 0; c < K; ++c) {
        double col[K];
#pragma unroll
        for (int i = 0; i < K; ++i) t.v[i] = (i == c) ? 1.0 : 0.0;
        solve_normal<K>(t, col);
#pragma unroll
        for (int i = 0; i < K; ++i) P[i][c] = col[i];
    }
}

// Runs rows [c0, c1).  `first_chunk`: the series starts here -> P = p0 I, theta = theta0 exactly as the
// reference.  Otherwise (A_in, b_in) is the information state entering the chunk (prior included).
template <int K, typename Src, typename Emit>
B200_HD void rls_chunk(const Src &src, const RlsCfg &cfg, bool first_chunk, const double *theta0,
                       const NormalState<K> *in, int64_t c0, int64_t c1, Emit &emit) {
    double P[K][K], theta[K];
    if (first_chunk) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            theta[i] = theta0 ? theta0[i] : 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) P[i][j] = (i == j) ? cfg.p0 : 0.0;
        }
    } else {
        NormalState<K> t = *in;
        rls_restart<K>(t, P, theta);
    }
    const double inv_lam = 1.0 / cfg.lambda;  // exact (1.0) for the expanding case
    double x[K], y;
    int64_t exact_left = first_chunk ? RLS_EXACT_ROWS : 0;  // valid rows still to run in the reference's literal arithmetic
    for (int64_t r = c0; r < c1; ++r) {
        if (r + MOVING_PF < c1) B200_PREFETCH(src, r + MOVING_PF);
        if (src.valid(r)) {
            src.load(r, x, y);
            if (exact_left > 0) {
                rls_update_exact<K>(P, theta, x, y, cfg.lambda);
                --exact_left;
            } else {
                rls_update<K>(P, theta, x, y, cfg.lambda, inv_lam);
            }
        }
        emit(r, theta, false);
    }
}

}  // namespace b200
