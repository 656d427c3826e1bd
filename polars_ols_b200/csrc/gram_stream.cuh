// gram_stream.cuh — shared definitions of the row-streaming Gram kernels (gram_cta.cuh, gram_multi.cuh, gram_wide.cuh,
// gram_pred.cuh, gram_ldg.cuh): the launch parameter block, the shared-memory stage geometry and the Gram epilogues
// (fused k x k solve or partial-record write).  They replace, for every group of a `.over()` batch at once,
//   construct_features_array + convert_polars_to_ndarray   src/expressions.rs:22-103  (SoA -> AoS copy)
//   x.t().dot(x), x.t().dot(y), + alpha I                  src/least_squares.rs:352-356
//   solve_normal_equations (Cholesky -> LU)                src/least_squares.rs:277-337
// of /root/reference.  Nothing is copied to row-major: the SoA columns are streamed once from HBM.
// (Round 1 also kept a per-warp TMA pipeline and an FP64-FMA row-per-lane kernel selectable; both lost to the
// CTA-cooperative pipeline everywhere they were measured — 0.65 / 0.65 against 0.94 of the HBM peak on C2,
// profiles/r01_sweep_variant3.json — and were removed in round 2.  The direct-load DMMA kernel of gram_ldg.cuh is the
// one non-TMA fallback.)
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
    // in-kernel completion of the fused gather (gram_cta_kernel): the LAST solver warp of the grid release-stores
    // `flag_step` into slot [flag_rank] of every peer's flag array and acquire-spins until its own slots reached `flag_wait`
    int flag_n;                       // 0 = off
    int flag_rank;
    unsigned long long *flag_peer[8];
    unsigned long long flag_step, flag_wait;
    unsigned int *done_counter;       // device: solver warps that finished (re-armed by the last one)
    int *flag_timeout;
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

}  // namespace b200
