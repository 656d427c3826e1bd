// gram_stream.cuh — the hot kernel: row-streaming X^T X / X^T y per (group) segment with a fused
// k x k solve.  Replaces, for every group of a `.over()` batch at once,
//   construct_features_array + convert_polars_to_ndarray   src/expressions.rs:22-103  (SoA -> AoS copy)
//   x.t().dot(x), x.t().dot(y), + alpha I                  src/least_squares.rs:352-356
//   solve_normal_equations (Cholesky -> LU)                src/least_squares.rs:277-337
// of /root/reference.  Nothing is copied to row-major: the SoA columns are streamed once from HBM.
//
// Structure (one warp = one independent pipeline, persistent over its segments):
//   * the warp's elected lanes issue 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP, the TMA
//     engine) of the k+1(+w)(+mask) column slices of a row tile into its private shared-memory stage,
//     completion counted on an mbarrier (complete_tx); STAGES tiles are in flight per warp;
//   * the warp waits on the stage's mbarrier and feeds FP64 tensor-core MMAs (mma.sync.m8n8k4.f64 ->
//     SASS DMMA): with A = X^T tile and B = X tile the A and B fragments are THE SAME register
//     (lane holds X[row = lane&3][feature = lane>>2]), so a k<=8 group costs one 16-byte LDS per lane per
//     8 rows; X^T y rides along as one DFMA per block;
//   * after the last tile of a segment the 8x8 accumulator fragments go to a per-warp scratch and lane 0
//     runs the reference's Cholesky -> LU ladder; beta goes straight to HBM (fused path), or the raw
//     Gram partial is written for the multi-segment / elastic-net solve kernel.
// HBM traffic = algorithmic bytes (each column element is read exactly once) + 8k bytes out per group.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx.cuh"
#include "solvers.cuh"

namespace b200 {

constexpr int GRAM_MAX_COLS = 72;   // 64 features + y + w + mask (+ slack)
constexpr int GRAM_MAX_WARPS = 16;
constexpr int GRAM_MAX_STAGES = 16;

struct GramParams {
    const void *cols[GRAM_MAX_COLS];  // [0,kd) features, [kd] y, then sqrt-weights/weights, then row mask
    int kd;                           // data feature columns
    int intercept;                    // 1: feature index kd is the constant 1.0 (appended last)
    int F;                            // kd + intercept = number of coefficients
    int has_w;                        // weights column present
    int w_is_sqrt;                    // weights column already holds sqrt(w) (sanitised path)
    int has_mask;                     // row-mask column present (T typed, 0 = row dropped from the fit)
    int64_t n_rows_pad;               // readable rows of every column (multiple of 16/sizeof(T))
    int64_t nseg;
    const int64_t *seg_off;           // [nseg+1] packed row offsets
    const int32_t *seg_group;         // [nseg] group of each segment, or nullptr (segment == group)
    int64_t max_seg_rows;             // host-side hint: longest segment (tile sizing)
    int tile_rows;                    // R, multiple of 8
    int stages;                       // tiles in flight per warp
    int team;                         // gram_cta: consumer warps per segment (0 = all)
    int red_depth;                    // gram_cta: ring depth of published accumulator buffers (0 = max)
    // outputs
    int fused;                        // 1: solve in the epilogue and write beta; 0: write partials
    double *partial;                  // [nseg][F*F + F + 1]  (G row-major, c, n_fit)
    double *beta;                     // [n_groups][F]
    int32_t *flags;                   // [n_groups]
    double alpha;                     // ridge penalty added to the diagonal (NOT scaled by n)
    int use_lu;                       // "lu": skip Cholesky
    double illcond_ratio;             // flag groups whose squared-pivot ratio exceeds this
    // fused multi-GPU gather: beta is also stored into every rank's full buffer (P2P over NVLink)
    int n_peers;
    double *peer_beta[8];             // [total_groups][F] on rank r
    int64_t peer_group_base;          // first global group index of this rank's shard
};

template <typename T> struct V2;
template <> struct V2<double> { using type = double2; };
template <> struct V2<float> { using type = float2; };

// column stride inside a stage: conflict-free for the (feature = lane>>2, row pair = lane&3) access
// pattern: 16-byte lanes (f64) want stride % 128 == 64, 8-byte lanes (f32) want stride % 128 == 32.
template <typename T>
__host__ __device__ inline uint32_t gram_col_stride(int tile_rows) {
    const uint32_t raw = static_cast<uint32_t>(tile_rows + 8) * sizeof(T);
    return ((raw + 127u) & ~127u) + (sizeof(T) == 8 ? 64u : 32u);
}

template <typename T>
__host__ __device__ inline uint32_t gram_scratch_bytes(int F, int fused) {
    if (!fused) return 0;
    const int FP = (F + 7) & ~7;
    return static_cast<uint32_t>(((FP * (FP + 1) + 2 * FP + 1) / 2 * 2) * sizeof(double));  // G (row stride FP+1), c, scratch
}

// Warp-cooperative Cholesky solve in registers (k <= FP <= 16): lane i holds row i of the SPD matrix in
// g[0..FP) and the right-hand side c_i.  Right-looking LL^T with full symmetric trailing updates, so that
// afterwards lane i holds row i of L in g[0..i] and COLUMN i of L in g[i+1..) — exactly what the forward
// (L z = c) and backward (L^T x = z) substitutions need without any transposition.  Values are exchanged
// with warp shuffles; one rsqrt per column replaces the divisions (results agree with the reference's
// faer LL^T to rounding).  Returns false on a non-positive / NaN pivot (-> LU fallback); mn / mx are the
// extreme squared pivots.  On success lane i returns beta_i in c.
// WIDTH < 32: a sub-warp of WIDTH lanes per system (batch_solve.cuh), `lane` is the lane inside the sub-warp and
// every sub-warp runs the whole sequence (no early exit: the shuffles need the full warp).
template <int FP, int WIDTH = 32>
__device__ __forceinline__ bool warp_chol_solve(double (&g)[FP], double &c, int F, int lane, double &mn, double &mx) {
    constexpr unsigned FULL = 0xffffffffu;
    bool ok = true;
    mn = INFINITY;
    mx = 0.0;
    double inv[FP];
#pragma unroll
    for (int j = 0; j < FP; ++j) {
        inv[j] = 1.0;
        if (j < F) {
            const double d = __shfl_sync(FULL, g[j], j, WIDTH);
            ok = ok && (d > 0.0);
            mn = fmin(mn, d);
            mx = fmax(mx, d);
            const double r = rsqrt(d);
            inv[j] = r;  // 1 / L_jj
            const double lij = (lane == j) ? d * r : g[j] * r;
            if (lane >= j) g[j] = lij;
#pragma unroll
            for (int cc = j + 1; cc < FP; ++cc) {
                if (cc < F) {
                    const double lcj = __shfl_sync(FULL, lij, cc, WIDTH);
                    if (lane > j) g[cc] = fma(-lij, lcj, g[cc]);
                    else if (lane == j) g[cc] = lcj;
                }
            }
        }
    }
    if (WIDTH == 32 && !ok) return false;
#pragma unroll
    for (int j = 0; j < FP; ++j) {
        if (j < F) {
            const double zj = __shfl_sync(FULL, c * inv[j], j, WIDTH);
            if (lane == j) c = zj;
            else if (lane > j) c = fma(-g[j], zj, c);
        }
    }
#pragma unroll
    for (int j = FP - 1; j >= 0; --j) {
        if (j < F) {
            const double xj = __shfl_sync(FULL, c * inv[j], j, WIDTH);
            if (lane == j) c = xj;
            else if (lane < j) c = fma(-g[j], xj, c);
        }
    }
    return ok;
}

// Second half of the epilogue: the per-warp scratch holds the symmetric Gram matrix (row stride FP + 1)
// followed by X^T y.  Fused mode: solve and write beta; otherwise (direct-FMA kernel only) dump the raw
// partial record for the standalone solve kernel.  nfit must be warp-uniform.
template <int KB>
__device__ __forceinline__ void gram_finish(const GramParams &p, int nfit, int64_t seg, double *Gs, int lane, double *beta_out = nullptr) {
    constexpr int FP = 8 * KB;
    constexpr int LD = FP + 1;
    const int F = p.F;
    double *cs = Gs + FP * LD;
    const int64_t g = p.seg_group ? p.seg_group[seg] : seg;
    if (!p.fused) {
        double *out = p.partial + static_cast<size_t>(seg) * (static_cast<size_t>(F) * F + F + 1);
        for (int e = lane; e < F * F; e += 32) out[e] = Gs[(e / F) * LD + (e % F)];
        if (lane < F) out[F * F + lane] = cs[lane];
        if (lane == 0) out[F * F + F] = static_cast<double>(nfit);
        __syncwarp();
        return;
    }
    // lane i takes row i of the (ridge) matrix and c_i into registers
    double grow[FP];
    const int li = (lane < FP) ? lane : 0;
#pragma unroll
    for (int c = 0; c < FP; ++c) grow[c] = Gs[li * LD + c];
    double ci = (lane < FP) ? cs[lane] : 0.0;
#pragma unroll
    for (int c = 0; c < FP; ++c)
        if (c == lane && lane < F) grow[c] += p.alpha;  // + alpha I, NOT scaled by n (src/least_squares.rs:352-356)
    int fl = 0;
    if (nfit == 0) {  // src/expressions.rs:357-359: no rows -> zeros
        ci = 0.0;
        fl = FLAG_EMPTY;
    } else {
        bool solved = false;
        if (!p.use_lu) {
            double mn, mx;
            const bool ok = warp_chol_solve<FP>(grow, ci, F, lane, mn, mx);
            if (ok) {
                solved = true;
                if (mx > p.illcond_ratio * mn) fl |= FLAG_ILLCOND;
            } else {
                fl |= FLAG_LU_FALLBACK;
            }
        }
        if (!solved) {  // "lu", or Cholesky hit a non-positive pivot: LU with partial pivoting (rare; lane 0)
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < F; ++i) Gs[i * LD + i] += p.alpha;
                lu_solve_inplace(Gs, LD, F, cs);
            }
            __syncwarp();
            ci = (lane < FP) ? cs[lane] : 0.0;
        }
    }
    if (nfit > 0 && nfit <= F) fl |= FLAG_WIDE;
    if (lane == 0) p.flags[g] = fl;
    if (lane < F) {
        p.beta[g * F + lane] = ci;
        for (int r = 0; r < p.n_peers; ++r) p.peer_beta[r][(p.peer_group_base + g) * F + lane] = ci;  // fused gather
    }
    if (beta_out) *beta_out = ci;  // lane i < F: beta_i (gram_pred.cuh predicts from it)
    __syncwarp();
}

// Epilogue shared by the TMA-staged and the direct-load kernels: reduce X^T y over the 4 lanes that share
// a feature, then either solve in shared memory (fused) or write the raw Gram partial.
template <int KB>
__device__ __forceinline__ void gram_epilogue(const GramParams &p, double (&acc)[KB * (KB + 1) / 2][2], double (&cy)[KB],
                                              int nfit, int64_t seg, double *Gs, int lane, double *beta_out = nullptr) {
    constexpr int FP = 8 * KB;
    const int fb = lane >> 2, q = lane & 3;
    const int F = p.F;
    // reduce X^T y and the row count over the 4 lanes that share a feature
#pragma unroll
    for (int bk = 0; bk < KB; ++bk) {
        cy[bk] += __shfl_xor_sync(0xffffffffu, cy[bk], 1);
        cy[bk] += __shfl_xor_sync(0xffffffffu, cy[bk], 2);
    }
    nfit += __shfl_xor_sync(0xffffffffu, nfit, 1);
    nfit += __shfl_xor_sync(0xffffffffu, nfit, 2);
    nfit = __shfl_sync(0xffffffffu, nfit, 0);
    const int64_t g = p.seg_group ? p.seg_group[seg] : seg;

    if (p.fused) {
        constexpr int LD = FP + 1;  // odd row stride: conflict-free row reads
        double *cs = Gs + FP * LD;
        int idx = 0;
#pragma unroll
        for (int bi = 0; bi < KB; ++bi) {
#pragma unroll
            for (int bj = bi; bj < KB; ++bj) {
                const int rr = 8 * bi + fb, cc = 8 * bj + 2 * q;
                Gs[rr * LD + cc] = acc[idx][0];
                Gs[rr * LD + cc + 1] = acc[idx][1];
                if (bi != bj) {
                    Gs[cc * LD + rr] = acc[idx][0];
                    Gs[(cc + 1) * LD + rr] = acc[idx][1];
                }
                ++idx;
            }
            if (q == 0) cs[8 * bi + fb] = cy[bi];
        }
        __syncwarp();
        gram_finish<KB>(p, nfit, seg, Gs, lane, beta_out);
    } else {
        double *out = p.partial + static_cast<size_t>(seg) * (static_cast<size_t>(F) * F + F + 1);
        int idx = 0;
#pragma unroll
        for (int bi = 0; bi < KB; ++bi) {
#pragma unroll
            for (int bj = bi; bj < KB; ++bj) {
                const int rr = 8 * bi + fb, cc = 8 * bj + 2 * q;
                if (rr < F) {
                    if (cc < F) out[rr * F + cc] = acc[idx][0];
                    if (cc + 1 < F) out[rr * F + cc + 1] = acc[idx][1];
                    if (bi != bj) {
                        if (cc < F) out[cc * F + rr] = acc[idx][0];
                        if (cc + 1 < F) out[(cc + 1) * F + rr] = acc[idx][1];
                    }
                }
                ++idx;
            }
            if (q == 0 && 8 * bi + fb < F) out[F * F + 8 * bi + fb] = cy[bi];
        }
        if (lane == 0) out[F * F + F] = static_cast<double>(nfit);
    }
}

template <typename T, int KB, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) gram_stream_kernel(const GramParams p) {
    using Vec = typename V2<T>::type;
    constexpr int FP = 8 * KB;
    constexpr int NPAIR = KB * (KB + 1) / 2;
    constexpr int A = 16 / sizeof(T);  // tile starts are aligned down to 16 bytes

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[GRAM_MAX_WARPS * GRAM_MAX_STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int fb = lane >> 2, q = lane & 3;
    const int kd = p.kd, F = p.F;
    const int ycol = kd, wcol = kd + 1, mcol = kd + 1 + (p.has_w ? 1 : 0);
    const int NC = kd + 1 + (p.has_w ? 1 : 0) + (p.has_mask ? 1 : 0);
    const int R = p.tile_rows, S = p.stages;
    const uint32_t stride = gram_col_stride<T>(R);
    const uint32_t stage_bytes = static_cast<uint32_t>(NC) * stride;
    const uint32_t scratch_bytes = gram_scratch_bytes<T>(F, p.fused);
    unsigned char *wbase = smem + static_cast<size_t>(warp) * (static_cast<size_t>(S) * stage_bytes + scratch_bytes);
    double *Gs = reinterpret_cast<double *>(wbase + static_cast<size_t>(S) * stage_bytes);
    uint64_t *bar = bars + warp * GRAM_MAX_STAGES;

    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bar[s], 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int64_t nseg = p.nseg;
    const int64_t wg = static_cast<int64_t>(blockIdx.x) * W + warp;
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * W;

    // ---- issue cursor: runs `S` tiles ahead of the consumer over the same (segment, tile) sequence ----
    int64_t iseg = wg, irow = 0, iend = 0;
    if (iseg < nseg) {
        irow = p.seg_off[iseg];
        iend = p.seg_off[iseg + 1];
    }
    int istage = 0;
    auto issue = [&]() {
        while (iseg < nseg && irow >= iend) {
            iseg += nwarps;
            if (iseg < nseg) {
                irow = p.seg_off[iseg];
                iend = p.seg_off[iseg + 1];
            }
        }
        if (iseg >= nseg) return;
        const int64_t a = irow;
        const int64_t b = (a + R < iend) ? a + R : iend;
        const int64_t a_al = a & ~static_cast<int64_t>(A - 1);
        int64_t b_al = (b + (A - 1)) & ~static_cast<int64_t>(A - 1);
        if (b_al > p.n_rows_pad) b_al = p.n_rows_pad;
        const uint32_t bytes = static_cast<uint32_t>(b_al - a_al) * sizeof(T);
        unsigned char *sb = wbase + static_cast<size_t>(istage) * stage_bytes;
        if (lane == 0) {
            fence_proxy_async_smem();  // our generic-proxy reads of this stage are done (WAR vs async proxy)
            mbar_arrive_expect_tx(&bar[istage], bytes * static_cast<uint32_t>(NC));
        }
        __syncwarp();
        for (int c = lane; c < NC; c += 32)
            bulk_g2s(sb + static_cast<size_t>(c) * stride, static_cast<const T *>(p.cols[c]) + a_al, bytes,
                     &bar[istage]);
        irow = b;
        istage = (istage + 1 == S) ? 0 : istage + 1;
    };
    for (int s = 0; s < S; ++s) issue();

    int cstage = 0;
    uint32_t phase = 0;

    for (int64_t seg = wg; seg < nseg; seg += nwarps) {
        const int64_t r0 = p.seg_off[seg], r1 = p.seg_off[seg + 1];
        constexpr bool DUAL = KB <= 2;  // two independent DMMA chains while the accumulators are few
        double acc[NPAIR][2], acc2[DUAL ? NPAIR : 1][2];
        double cy[KB];
#pragma unroll
        for (int i = 0; i < NPAIR; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
        for (int i = 0; i < (DUAL ? NPAIR : 1); ++i) acc2[i][0] = acc2[i][1] = 0.0;
#pragma unroll
        for (int i = 0; i < KB; ++i) cy[i] = 0.0;
        int nfit = 0;
        const bool plain = !p.has_mask && !p.has_w;  // warp-uniform: mask-free interior loop allowed

        for (int64_t row = r0; row < r1; row += R) {
            const int64_t b = (row + R < r1) ? row + R : r1;
            const int o = static_cast<int>(row & (A - 1));
            const int hi = o + static_cast<int>(b - row);  // valid local rows are [o, hi)
            mbar_wait(&bar[cstage], phase);
            const unsigned char *sb = wbase + static_cast<size_t>(cstage) * stage_bytes;
            const unsigned char *xs[KB];
#pragma unroll
            for (int bk = 0; bk < KB; ++bk) xs[bk] = sb + static_cast<size_t>(8 * bk + fb) * stride + 2 * q * sizeof(T);
            const unsigned char *ys = sb + static_cast<size_t>(ycol) * stride + 2 * q * sizeof(T);

            // predicated octet (segment edges, weights, row mask)
            auto masked_octet = [&](int j) {
                const int lr = 8 * j + 2 * q;
                bool v0 = (lr >= o) && (lr < hi);
                bool v1 = (lr + 1 >= o) && (lr + 1 < hi);
                const Vec y2 = *reinterpret_cast<const Vec *>(ys + 8 * j * sizeof(T));
                T s0 = T(1), s1 = T(1);
                if (p.has_mask) {
                    const Vec m2 = *reinterpret_cast<const Vec *>(sb + static_cast<size_t>(mcol) * stride + lr * sizeof(T));
                    v0 = v0 && (m2.x != T(0));
                    v1 = v1 && (m2.y != T(0));
                }
                if (p.has_w) {
                    const Vec w2 = *reinterpret_cast<const Vec *>(sb + static_cast<size_t>(wcol) * stride + lr * sizeof(T));
                    // polars: sqrt_w = w.sqrt() in the column dtype (polars_ols/least_squares.py:193)
                    s0 = p.w_is_sqrt ? w2.x : static_cast<T>(sqrt(w2.x));
                    s1 = p.w_is_sqrt ? w2.y : static_cast<T>(sqrt(w2.y));
                }
                // target * sqrt_w and feature * sqrt_w are evaluated in the column dtype, then cast to f64
                const double y0 = v0 ? static_cast<double>(static_cast<T>(y2.x * s0)) : 0.0;
                const double y1 = v1 ? static_cast<double>(static_cast<T>(y2.y * s1)) : 0.0;
                if (fb == 0) nfit += (v0 ? 1 : 0) + (v1 ? 1 : 0);
                double f0[KB], f1[KB];
#pragma unroll
                for (int bk = 0; bk < KB; ++bk) {
                    const int f = 8 * bk + fb;
                    T x0 = T(0), x1 = T(0);
                    if (f < kd) {
                        const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                        x0 = x2.x;
                        x1 = x2.y;
                    } else if (f == kd && p.intercept) {
                        x0 = T(1);
                        x1 = T(1);
                    }
                    f0[bk] = v0 ? static_cast<double>(static_cast<T>(x0 * s0)) : 0.0;
                    f1[bk] = v1 ? static_cast<double>(static_cast<T>(x1 * s1)) : 0.0;
                }
                int idx = 0;
#pragma unroll
                for (int bi = 0; bi < KB; ++bi) {
#pragma unroll
                    for (int bj = bi; bj < KB; ++bj) {
                        dmma_m8n8k4(acc[idx][0], acc[idx][1], f0[bi], f0[bj]);
                        if (DUAL) dmma_m8n8k4(acc2[idx][0], acc2[idx][1], f1[bi], f1[bj]);
                        else dmma_m8n8k4(acc[idx][0], acc[idx][1], f1[bi], f1[bj]);
                        ++idx;
                    }
                    cy[bi] = fma(f0[bi], y0, cy[bi]);
                    cy[bi] = fma(f1[bi], y1, cy[bi]);
                }
            };

            const int noct = (hi + 7) >> 3;
            if (!plain) {
                for (int j = 0; j < noct; ++j) masked_octet(j);
            } else {
                int j = 0;
                if (o != 0) {
                    masked_octet(0);
                    j = 1;
                }
                const int jfull = hi >> 3;  // octets [j, jfull) lie entirely inside [o, hi)
#pragma unroll 4
                for (; j < jfull; ++j) {
                    const Vec y2 = *reinterpret_cast<const Vec *>(ys + 8 * j * sizeof(T));
                    double f0[KB], f1[KB];
#pragma unroll
                    for (int bk = 0; bk < KB; ++bk) {
                        const int f = 8 * bk + fb;
                        if (f < kd) {
                            const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                            f0[bk] = static_cast<double>(x2.x);
                            f1[bk] = static_cast<double>(x2.y);
                        } else {
                            f0[bk] = f1[bk] = (f == kd && p.intercept) ? 1.0 : 0.0;
                        }
                    }
                    const double y0 = static_cast<double>(y2.x), y1 = static_cast<double>(y2.y);
                    int idx = 0;
#pragma unroll
                    for (int bi = 0; bi < KB; ++bi) {
#pragma unroll
                        for (int bj = bi; bj < KB; ++bj) {
                            dmma_m8n8k4(acc[idx][0], acc[idx][1], f0[bi], f0[bj]);
                            if (DUAL) dmma_m8n8k4(acc2[idx][0], acc2[idx][1], f1[bi], f1[bj]);
                        else dmma_m8n8k4(acc[idx][0], acc[idx][1], f1[bi], f1[bj]);
                            ++idx;
                        }
                        cy[bi] = fma(f0[bi], y0, cy[bi]);
                        cy[bi] = fma(f1[bi], y1, cy[bi]);
                    }
                }
                if (j < noct) masked_octet(j);
            }
            __syncwarp();
            issue();  // refill the stage we just drained
            if (++cstage == S) {
                cstage = 0;
                phase ^= 1u;
            }
        }
#pragma unroll
        for (int i = 0; i < (DUAL ? NPAIR : 0); ++i) {
            acc[i][0] += acc2[i][0];
            acc[i][1] += acc2[i][1];
        }
        if (plain) nfit = (lane == 0) ? static_cast<int>(r1 - r0) : 0;

        gram_epilogue<KB>(p, acc, cy, nfit, seg, Gs, lane);
    }
}

// max consumer warps per CTA for a given number of 8-feature blocks (register budget: the 8x8 f64
// accumulator fragments of all block pairs live in registers)
__host__ __device__ constexpr int gram_max_warps(int KB) { return KB <= 2 ? 16 : (KB <= 4 ? 8 : 4); }

// defined in gram_f64.cu / gram_f32.cu (one translation unit per dtype keeps the build parallel)
cudaError_t gram_launch_f64(int KB, const GramParams &p, unsigned grid, int warps, size_t smem, cudaStream_t s);
cudaError_t gram_launch_f32(int KB, const GramParams &p, unsigned grid, int warps, size_t smem, cudaStream_t s);

template <typename T, int KB>
cudaError_t gram_launch_t(const GramParams &p, unsigned grid, int warps, size_t smem, cudaStream_t s) {
    auto kern = gram_stream_kernel<T, KB, gram_max_warps(KB)>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, warps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

template <typename T>
cudaError_t gram_launch_any(int KB, const GramParams &p, unsigned grid, int warps, size_t smem, cudaStream_t s) {
    switch (KB) {
        case 1: return gram_launch_t<T, 1>(p, grid, warps, smem, s);
        case 2: return gram_launch_t<T, 2>(p, grid, warps, smem, s);
        case 4: return gram_launch_t<T, 4>(p, grid, warps, smem, s);
        default: return gram_launch_t<T, 8>(p, grid, warps, smem, s);
    }
}

}  // namespace b200
