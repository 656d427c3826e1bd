// moving_fast.cuh — the streaming kernels of rolling_least_squares / recursive_least_squares for null-free frames
// with k <= 8 coefficients (the C4 shape: ONE series of 50M rows x 6 features, and every `.over()` frame without
// nulls).  Same chunk algorithms as moving_core.cuh (one thread per time chunk, window / information state rebuilt
// exactly at the chunk start), but the rows reach the thread through a PRIVATE staging ring in shared memory:
//
//   * every thread walks its own chunk, so consecutive lanes read addresses a whole chunk apart.  Instead of first
//     transposing the frame into a chunk-interleaved copy (an extra read + write pass over all inputs,
//     chunk_transpose_kernel) and then paying one exposed HBM round trip per row, each thread issues 16-byte
//     cp.async copies (SASS LDGSTS) for its next rows straight from the caller's SoA columns into its own slots
//     of a shared-memory ring, `MF_DEPTH` stages ahead of the row it is working on.  A 16-byte unit is two f64
//     (four f32) consecutive rows of one column: sectors are consumed whole within one stage or the next.
//   * the ring is laid out [slot][column][thread] in 16-byte units: the lanes of a warp write and read consecutive
//     units (no bank conflicts), and since a thread only ever reads what it copied itself no block barrier is
//     needed — cp.async.wait_group orders a thread's own copies.
//   * rolling: the row leaving the window (r - W) is a second stream through a second ring, delayed by W rows;
//     the window sums entering a chunk are accumulated from the W rows before it through the same lead stream.
//
// Reference semantics restated (null-free): src/least_squares.rs:848-986 (warm-up over the first min_periods rows,
// first coefficients at row min_periods - 1, rank-1 add / subtract + Cholesky (LU fallback) per row; both null
// branches coincide without nulls while min_periods <= window) and :505-598 (RLS).  Frames with a row mask, with
// min_periods > window, or with k > 8 take the general kernels of moving.cuh.
#pragma once
#include "moving.cuh"

namespace b200 {

#ifndef MF_DEPTH
#define MF_DEPTH 1          // stages (16-byte units per column) in flight ahead of the one being consumed
#endif
// resident blocks per SM the register allocation aims at (ptxas -v, k = 6 f64): rolling 168 registers at 3 blocks
// (12 warps / SM, 56 bytes of spills in the cold LU fallback); rls 240 registers at 2 blocks, heavy spills at 3
#ifndef MF_MIN_BLOCKS_ROLLING
#define MF_MIN_BLOCKS_ROLLING 3
#endif
#ifndef MF_MIN_BLOCKS_RLS
#define MF_MIN_BLOCKS_RLS 2
#endif
constexpr int MF_THREADS = 128;
constexpr int MF_LEAD_SLOTS = MF_DEPTH + 1;
constexpr int MF_LAG_SLOTS = MF_DEPTH + 2;

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One thread's view of its staging rings.  Stage q of a stream = rows [a0 + q * RPU, a0 + (q + 1) * RPU) of every
// column (RPU rows per 16-byte unit), kept in slot q % SLOTS.
template <typename T, int K, bool WT>
struct FastSrc {
    static constexpr int RPU = 16 / static_cast<int>(sizeof(T));
    const T *col[K + 2];   // [0, kd) features, [kd] target, [kd + 1] weights
    int nc;                // columns staged
    int kd, has_w, w_is_sqrt;
    uint32_t lead, lag;    // shared-memory byte address of this thread's unit 0 in each ring
    int64_t a0;            // first staged row (multiple of RPU: the copies are 16-byte aligned)
    int64_t n_lim;         // rows at or beyond this are never fetched (end of the frame, padded to 16 bytes)

    __device__ __forceinline__ void issue(uint32_t ring, int slots, int64_t q) const {
        if (q < 0) return;
        issue_slot(ring, static_cast<int>(q % slots), q);
    }
    // stage q into ring slot `slot` (callers that walk the stages in order carry the slot index along)
    __device__ __forceinline__ void issue_slot(uint32_t ring, int slot, int64_t q) const {
        const int64_t row = a0 + q * RPU;
        if (row >= n_lim) return;
        const uint32_t dst = ring + static_cast<uint32_t>(slot * nc) * (MF_THREADS * 16u);
#pragma unroll
        for (int c = 0; c < K + 2; ++c)  // compile-time bound: `col` stays in registers
            if (c < nc) cp_async16(dst + static_cast<uint32_t>(c) * (MF_THREADS * 16u), col[c] + row);
    }
    // scaled features (incl. intercept), scaled target, raw target and 1/sqrt(w)-able scale of staged row r
    __device__ __forceinline__ void read(uint32_t ring, int slots, int64_t r, double (&x)[K], double &y, T &y_raw, T &s) const {
        read_from(ring, slots, r, a0, x, y, y_raw, s);
    }
    // same, for a ring whose stage 0 starts at row `origin` (a neighbouring thread's ring in rolling_nbr_kernel)
    __device__ __forceinline__ void read_from(uint32_t ring, int slots, int64_t r, int64_t origin, double (&x)[K], double &y, T &y_raw,
                                              T &s) const {
        const int64_t off = r - origin;
        const int64_t q = off / RPU;
        read_slot(ring, static_cast<int>(q % slots), static_cast<int>(off - q * RPU), x, y, y_raw, s);
    }
    // row `elem` of the unit in ring slot `slot`
    __device__ __forceinline__ void read_slot(uint32_t ring, int slot, int elem, double (&x)[K], double &y, T &y_raw, T &s) const {
        const uint32_t src = ring + static_cast<uint32_t>(slot * nc) * (MF_THREADS * 16u) + static_cast<uint32_t>(elem) * static_cast<uint32_t>(sizeof(T));
        auto lds = [&](int c) -> T {
            T v;
            if constexpr (sizeof(T) == 8) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(src + static_cast<uint32_t>(c) * (MF_THREADS * 16u)));
            else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(src + static_cast<uint32_t>(c) * (MF_THREADS * 16u)));
            return v;
        };
        s = T(1);
        if constexpr (WT) {
            const T wv = lds(kd + 1);
            s = w_is_sqrt ? wv : static_cast<T>(sqrt(wv));
#pragma unroll
            for (int j = 0; j < K; ++j) x[j] = (j < kd) ? static_cast<double>(static_cast<T>(lds(j) * s)) : static_cast<double>(s);
            y_raw = lds(kd);
            y = static_cast<double>(static_cast<T>(y_raw * s));
        } else {  // unweighted instantiation: no multiply by 1 on the FP64 pipe
#pragma unroll
            for (int j = 0; j < K; ++j) x[j] = (j < kd) ? static_cast<double>(lds(j)) : 1.0;
            y_raw = lds(kd);
            y = static_cast<double>(y_raw);
        }
    }
};

template <typename T, int K, bool WT>
__device__ __forceinline__ FastSrc<T, K, WT> make_fast_src(const MovingParams &p, unsigned char *smem) {
    FastSrc<T, K, WT> s;
#pragma unroll
    for (int j = 0; j < K + 2; ++j) s.col[j] = nullptr;
    s.kd = p.kd;
    s.has_w = p.w ? 1 : 0;
#pragma unroll
    for (int j = 0; j < K + 2; ++j) {
        if (j <= p.kd) s.col[j] = static_cast<const T *>(p.cols[j]);
        else if (j == p.kd + 1 && s.has_w) s.col[j] = static_cast<const T *>(p.w);
    }
    s.nc = p.kd + 1 + s.has_w;
    s.w_is_sqrt = p.w_is_sqrt;
    const uint32_t base = smem_u32(smem) + threadIdx.x * 16u;
    s.lead = base;
    s.lag = base + static_cast<uint32_t>(MF_LEAD_SLOTS * s.nc) * (MF_THREADS * 16u);
    s.n_lim = p.n_rows;
    s.a0 = 0;
    return s;
}

__host__ __device__ inline size_t moving_fast_smem(int nc, bool rolling) {
    return static_cast<size_t>(MF_LEAD_SLOTS + (rolling ? MF_LAG_SLOTS : 0)) * nc * MF_THREADS * 16;
}

// output of one row (same rules as DevEmit in moving.cuh): coefficients, or prediction / residual with the WLS
// un-scaling, `fill_nan(None)` for rolling
template <typename T, int K, bool WT>
__device__ __forceinline__ void fast_emit(const MovingParams &p, int64_t r, const double (&beta)[K], const double (&x)[K], T y_raw, T s) {
    const int64_t orow = p.row_index ? p.row_index[r] : r;
    if (p.mode == 2) {
        // one thread owns the row: 16-byte stores where the row starts on a 16-byte boundary (every row when K is even);
        // the validity bytes (NaN <=> null) are written by a coalesced pass after the kernel (launch_moving_t)
        double *o = p.out + orow * K;
        if ((K % 2 == 0) || ((orow & 1) == 0)) {
#pragma unroll
            for (int j = 0; j + 1 < K; j += 2) *reinterpret_cast<double2 *>(o + j) = make_double2(beta[j], beta[j + 1]);
            if (K % 2) o[K - 1] = beta[K - 1];
        } else {
            o[0] = beta[0];
#pragma unroll
            for (int j = 1; j + 1 < K; j += 2) *reinterpret_cast<double2 *>(o + j) = make_double2(beta[j], beta[j + 1]);
            if (K % 2 == 0) o[K - 1] = beta[K - 1];
        }
        return;
    }
    double pred = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) pred += x[j] * beta[j];
    if constexpr (WT) pred *= static_cast<double>(T(1) / s);
    bool valid = true;
    if (p.mode == 1) {
        double t = static_cast<double>(y_raw);
        if (p.target != p.cols[p.kd]) t = static_cast<double>(static_cast<const T *>(p.target)[p.target_is_packed ? r : orow]);
        pred = t - pred;
        if (p.target_validity) valid = (p.target_validity[orow >> 3] >> (orow & 7)) & 1;
    }
    if (p.kind == MOVING_ROLLING) valid = valid && (pred == pred);
    p.out[orow] = pred;
    if (p.out_valid && p.mode == 1 && p.target_validity) p.out_valid[orow] = valid ? 1 : 0;  // else: coalesced pass afterwards
}

// validity bytes of the staged kernels' outputs in one coalesced pass: coefficient / rolling-prediction nulls are exactly
// the NaNs (fill_nan(None)); rls predictions of a null-free frame are all valid
static __global__ void __launch_bounds__(256) moving_validity_kernel(const double *__restrict__ v, uint8_t *__restrict__ m, int64_t n, int nan_is_null) {
    const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    if (i + 8 <= n) {
        unsigned long long bits = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double x = v[i + u];
            bits |= static_cast<unsigned long long>((!nan_is_null || x == x) ? 1u : 0u) << (8 * u);
        }
        *reinterpret_cast<unsigned long long *>(m + i) = bits;
    } else {
        for (int64_t u = i; u < n; ++u) m[u] = (!nan_is_null || v[u] == v[u]) ? 1 : 0;
    }
}

template <typename T, int K>
__device__ __forceinline__ void fast_emit_nan(const MovingParams &p, int64_t r) {
    const int64_t orow = p.row_index ? p.row_index[r] : r;
    if (p.mode == 2) {
        for (int j = 0; j < K; ++j) p.out[orow * K + j] = NAN;
        return;
    }
    p.out[orow] = NAN;
    if (p.out_valid && p.mode == 1 && p.target_validity) p.out_valid[orow] = 0;
}

// ---- rolling ---------------------------------------------------------------------------------------------
// Requires: no row mask, min_periods <= window.  Chunk [c0, c1) of series [g0, g1):
//   rows before g0 + min_periods - 1 are NaN; the state entering the first coefficient row rs of the chunk is
//   alpha I + Gram(rows [max(g0, rs - W), rs)); then per row r: + row r, - row r - W (when r - W >= g0), solve.
template <typename T, int K, bool WT>
__global__ void __launch_bounds__(MF_THREADS, MF_MIN_BLOCKS_ROLLING) rolling_fast_kernel(const MovingParams p) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    constexpr int RPU = FastSrc<T, K, WT>::RPU;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= p.n_chunks) return;
    const int64_t g = p.chunk_group[c];
    const int64_t g0 = p.group_off[g], g1 = p.group_off[g + 1];
    const int64_t c0 = p.chunk_r0[c], c1 = p.chunk_r1[c];
    const int64_t W = p.window;
    const int64_t first = g0 + p.min_periods - 1;
    int64_t r = c0;
    if ((g1 - g0) < p.min_periods) {  // fewer rows than min_periods: all NaN (src/least_squares.rs:893-900)
        for (; r < c1; ++r) fast_emit_nan<T, K>(p, r);
        return;
    }
    for (; r < c1 && r < first; ++r) fast_emit_nan<T, K>(p, r);
    if (r >= c1) return;
    const int64_t rs = r;                                  // first row of this chunk that carries coefficients
    const int64_t s0 = (rs - W > g0) ? rs - W : g0;         // oldest row of the window entering rs
    FastSrc<T, K, WT> src = make_fast_src<T, K, WT>(p, mf_smem);
    src.a0 = s0 & ~static_cast<int64_t>(RPU - 1);
    // lag stage consumed while the lead works on stage q:  q - lagq (and q - lagq + 1 when W is not a multiple of RPU)
    const int64_t lagq = (W + RPU - 1) / RPU;

    NormalState<K> st;
    st.clear();
    if (p.alpha > 0.0) st.add_diag(p.alpha);
    double beta[K], x[K], xo[K], y, yo;
    T y_raw, s, yr2, s2;

    // prologue: the stages a boundary `b` in [-MF_DEPTH, 0) would have issued
#pragma unroll
    for (int b = -MF_DEPTH; b < 0; ++b) {
        src.issue(src.lead, MF_LEAD_SLOTS, b + MF_DEPTH);
        src.issue(src.lag, MF_LAG_SLOTS, b - lagq + MF_DEPTH + 1);
        cp_async_commit();
    }
    const int64_t q_end = (c1 - src.a0 + RPU - 1) / RPU;
    for (int64_t q = 0; q < q_end; ++q) {
        src.issue(src.lead, MF_LEAD_SLOTS, q + MF_DEPTH);
        src.issue(src.lag, MF_LAG_SLOTS, q - lagq + MF_DEPTH + 1);
        cp_async_commit();
        cp_async_wait<MF_DEPTH>();  // lead stage q and lag stages <= q - lagq + 1 have landed
#pragma unroll
        for (int wi = 0; wi < RPU; ++wi) {
            const int64_t row = src.a0 + q * RPU + wi;
            if (row < s0 || row >= c1) continue;
            src.read(src.lead, MF_LEAD_SLOTS, row, x, y, y_raw, s);
            st.add(x, y, 1.0);
            if (row < rs) continue;  // still accumulating the window that enters the chunk
            if (row - W >= g0) {
                src.read(src.lag, MF_LAG_SLOTS, row - W, xo, yo, yr2, s2);
                st.add(xo, yo, -1.0);
            }
            solve_normal<K>(st, beta);
            fast_emit<T, K, WT>(p, row, beta, x, y_raw, s);
        }
    }
    cp_async_wait<0>();
}

// ---- rolling, window-length chunks with neighbour sharing ("nbr") ------------------------------------------------
// The private lag stream of rolling_fast_kernel doubles the number of concurrent 16-byte streams; at 12 warps / SM the
// lines they touch (14 streams x 57k threads x 128 B) no longer fit the L2 and every 16-byte unit costs a 64-byte DRAM
// fetch (ncu, 10M rows: 557 B/row read for 56 B/row of input).  Here every chunk is exactly W rows long, so the row
// leaving the window of (chunk c, row i) is (chunk c - 1, row i): the row the NEIGHBOURING thread consumes at the same
// step, already sitting in ITS staging slot.  One lead stream per thread, the lag row is read from the neighbour's
// slot; the window sums entering a chunk are the Gram totals of the previous chunk (chunk_totals_kernel, one coalesced
// pass).  The block advances in lock step, one barrier per stage; thread 0 of a block only stages the chunk before the
// block's first one.  Rows are consumed one stage behind the newest landed one because a neighbour's unit boundaries
// may be shifted by up to a unit (W or the series start need not be multiples of the unit).
constexpr int NB_SLOTS = MF_DEPTH + 4;

// One warp per chunk, lanes stride over the rows (coalesced 256-byte requests per column), four rows in flight per
// lane.  DECAY = false: plain Gram totals of the chunk (rolling_nbr_kernel's window sums entering the next chunk).
// DECAY = true: the information-form summary of rls (A_c = sum lam^(n-1-i) x_i x_i^T, b_c likewise, D = lam^n, as
// rls_summarise in moving_core.cuh): row i carries the weight lam^(n-1-i); a lane starts from one exp() and steps its
// weight by lam^-32 per row of its own.
template <typename T, int K, bool DECAY>
__global__ void __launch_bounds__(256) chunk_totals_kernel(const MovingParams p, double *__restrict__ totals) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (c >= p.n_chunks) return;
    const int64_t r0 = p.chunk_r0[c], r1 = p.chunk_r1[c];
    if (!DECAY && r1 == p.group_off[p.chunk_group[c] + 1]) return;  // last chunk of its series: nobody continues from it
    const DevSrc<T, K> src = make_src<T, K>(p);
    NormalState<K> st;
    st.clear();
    const int64_t n = r1 - r0;
    double wgt = 1.0, step = 1.0;
    const double ll = DECAY ? log(p.lambda) : 0.0;
    bool direct = false;  // absurdly short half-lives: lam^-32 overflows, take every weight from exp() instead
    if (DECAY) {
        wgt = exp(ll * static_cast<double>(n - 1 - lane));   // lam^(n-1-i) of this lane's first row
        step = exp(-32.0 * ll);                              // lam^-32: the next row of this lane is 32 rows later
        direct = !(step < 1.0e300);
    }
    auto next_weight = [&](int64_t row_done) {
        if (!DECAY) return;
        if (direct) wgt = exp(ll * static_cast<double>(r1 - 1 - (row_done + 32)));
        else wgt *= step;
    };
    constexpr int U = 4;
    int64_t r = r0 + lane;
    for (; r + 32 * (U - 1) < r1; r += 32 * U) {
        double x[U][K], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) src.load(r + 32 * u, x[u], y[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            st.add(x[u], y[u], wgt);
            next_weight(r + 32 * u);
        }
    }
    for (; r < r1; r += 32) {
        double x[K], y;
        src.load(r, x, y);
        st.add(x, y, wgt);
        next_weight(r);
    }
    double *rec = totals + c * moving_rec(K);
#pragma unroll
    for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double v = st.S[i][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) rec[i * K + j] = v;
            if (DECAY && lane == 0 && j < i) rec[j * K + i] = 0.0;
        }
        double v = st.v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) rec[K * K + i] = v;
    }
    if (DECAY && lane == 0) rec[K * K + K] = exp(log(p.lambda) * static_cast<double>(n));
}

__host__ __device__ inline size_t moving_nbr_smem(int nc) { return static_cast<size_t>(NB_SLOTS) * nc * MF_THREADS * 16; }

template <typename T, int K, bool WT>
__global__ void __launch_bounds__(MF_THREADS, MF_MIN_BLOCKS_ROLLING) rolling_nbr_kernel(const MovingParams p, const double *__restrict__ totals) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    constexpr int RPU = FastSrc<T, K, WT>::RPU;
    const int tid = threadIdx.x;
    const int64_t ci = static_cast<int64_t>(blockIdx.x) * (MF_THREADS - 1) + tid - 1;
    const bool has = ci >= 0 && ci < p.n_chunks;
    const bool compute = has && tid > 0;
    const int64_t W = p.window;
    int64_t r0 = 0, r1 = 0, g0 = 0, g1 = 0;
    if (has) {
        r0 = p.chunk_r0[ci];
        r1 = p.chunk_r1[ci];
        const int64_t g = p.chunk_group[ci];
        g0 = p.group_off[g];
        g1 = p.group_off[g + 1];
    }
    FastSrc<T, K, WT> src = make_fast_src<T, K, WT>(p, mf_smem);
    src.a0 = r0 & ~static_cast<int64_t>(RPU - 1);
    src.n_lim = r1;                                     // nobody needs rows of another chunk from this thread
    const bool is_first = r0 == g0;                     // first chunk of its series: the window only grows
    const int64_t origin_prev = (r0 - W) & ~static_cast<int64_t>(RPU - 1);
    const uint32_t nbr = src.lead - 16u;                // thread tid - 1 staged chunk ci - 1
    const bool all_nan = (g1 - g0) < p.min_periods;     // src/least_squares.rs:893-900
    const int64_t first = g0 + p.min_periods - 1;

    NormalState<K> st;
    st.clear();
    if (compute && !is_first) {
        const double *rec = totals + (ci - 1) * moving_rec(K);
#pragma unroll
        for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) st.S[i][j] = rec[i * K + j];
            st.v[i] = rec[K * K + i];
        }
    }
    if (p.alpha > 0.0) st.add_diag(p.alpha);
    double beta[K], x[K], xo[K], y, yo;
    T y_raw, s, yr2, s2;

    // lag row of (stage q - 1, element wi) in the neighbour's numbering: stage q - 1 + dlt[wi], element el[wi]
    // (the neighbour's units may be shifted by up to one unit: W or the series start need not be unit multiples)
    int dlt[RPU], el[RPU];
    {
        const int c = static_cast<int>(src.a0 - W - origin_prev);  // in (-RPU, RPU)
#pragma unroll
        for (int wi = 0; wi < RPU; ++wi) {
            const int t = wi + c + RPU;  // > 0
            dlt[wi] = t / RPU - 1;
            el[wi] = t % RPU;
        }
    }
#pragma unroll
    for (int b = -MF_DEPTH; b < 0; ++b) {
        if (has) src.issue_slot(src.lead, b + MF_DEPTH, b + MF_DEPTH);
        cp_async_commit();
    }
    const int64_t q_end = W / RPU + 3;  // block-uniform: covers a chunk starting anywhere inside a unit + the one-behind drain
    int slot_issue = MF_DEPTH % NB_SLOTS;  // slot of stage q + MF_DEPTH
    int slot_own = NB_SLOTS - 1;           // slot of stage q - 1
    int64_t row0 = src.a0 - RPU;           // first row of stage q - 1
    for (int64_t q = 0; q < q_end; ++q) {
        if (has) src.issue_slot(src.lead, slot_issue, q + MF_DEPTH);
        cp_async_commit();
        cp_async_wait<MF_DEPTH>();
        __syncthreads();  // every thread's stage q has landed; stage q - 3 may be overwritten from now on
        if (compute && q > 0) {
#pragma unroll
            for (int wi = 0; wi < RPU; ++wi) {
                const int64_t row = row0 + wi;
                if (row < r0 || row >= r1) continue;
                src.read_slot(src.lead, slot_own, wi, x, y, y_raw, s);
                st.add(x, y, 1.0);
                if (!is_first) {
                    int sl = slot_own + dlt[wi];
                    sl = sl < 0 ? sl + NB_SLOTS : (sl >= NB_SLOTS ? sl - NB_SLOTS : sl);
                    src.read_slot(nbr, sl, el[wi], xo, yo, yr2, s2);
                    st.add(xo, yo, -1.0);
                }
                if (all_nan || row < first) {
                    fast_emit_nan<T, K>(p, row);
                } else {
                    solve_normal<K>(st, beta);
                    fast_emit<T, K, WT>(p, row, beta, x, y_raw, s);
                }
            }
        }
        slot_issue = slot_issue + 1 == NB_SLOTS ? 0 : slot_issue + 1;
        slot_own = slot_own + 1 == NB_SLOTS ? 0 : slot_own + 1;
        row0 += RPU;
    }
    cp_async_wait<0>();
}

// ---- recursive least squares -------------------------------------------------------------------------------
// pass 1 (chunk summaries in information form): chunk_totals_kernel<T, K, true> above
// pass 3 (after the scan): the covariance-form recurrence of every chunk, restarted from the information state
// entering it (rls_chunk of moving_core.cuh, rows through the staging ring)
template <typename T, int K, bool WT>
__global__ void __launch_bounds__(MF_THREADS, MF_MIN_BLOCKS_RLS) rls_fast_main_kernel(const MovingParams p) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    constexpr int RPU = FastSrc<T, K, WT>::RPU;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= p.n_chunks) return;
    const int64_t g = p.chunk_group[c];
    const int64_t c0 = p.chunk_r0[c], c1 = p.chunk_r1[c];
    const bool first = (c0 == p.group_off[g]) && !p.init_info;
    FastSrc<T, K, WT> src = make_fast_src<T, K, WT>(p, mf_smem);
    src.a0 = c0 & ~static_cast<int64_t>(RPU - 1);
#pragma unroll
    for (int b = -MF_DEPTH; b < 0; ++b) {
        src.issue(src.lead, MF_LEAD_SLOTS, b + MF_DEPTH);
        cp_async_commit();
    }
    double P[K][K], theta[K];
    if (first) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            theta[i] = p.has_mean ? p.mean[i] : 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) P[i][j] = (i == j) ? p.p0 : 0.0;
        }
    } else {
        NormalState<K> t;
        const double *rec = p.summaries + c * moving_rec(K);
#pragma unroll
        for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int j = 0; j < K; ++j) t.S[i][j] = rec[i * K + j];
            t.v[i] = rec[K * K + i];
        }
        rls_restart<K>(t, P, theta);
    }
    const double inv_lam = 1.0 / p.lambda;
    double x[K], y;
    T y_raw, s;
    const int64_t exact_end = first ? c0 + RLS_EXACT_ROWS : c0;  // prior-dominated head of a series: literal arithmetic
    const int64_t q_end = (c1 - src.a0 + RPU - 1) / RPU;
    for (int64_t q = 0; q < q_end; ++q) {
        src.issue(src.lead, MF_LEAD_SLOTS, q + MF_DEPTH);
        cp_async_commit();
        cp_async_wait<MF_DEPTH>();
#pragma unroll
        for (int wi = 0; wi < RPU; ++wi) {
            const int64_t row = src.a0 + q * RPU + wi;
            if (row < c0 || row >= c1) continue;
            src.read(src.lead, MF_LEAD_SLOTS, row, x, y, y_raw, s);
            if (row < exact_end) rls_update_exact<K>(P, theta, x, y, p.lambda);
            else rls_update<K>(P, theta, x, y, p.lambda, inv_lam);
            fast_emit<T, K, WT>(p, row, theta, x, y_raw, s);
        }
    }
    cp_async_wait<0>();
}

}  // namespace b200
