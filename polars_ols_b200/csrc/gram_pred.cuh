// gram_pred.cuh — ONE kernel for mode = predictions | residuals of the static models when a whole group fits one
// shared-memory tile (k <= 16, null-free contiguous groups): Gram -> solve -> predictions, every input byte read from
// HBM once.
//
// The two-pass route (gram_cta_kernel, then predict_kernel) reads the features twice: 72 KB + 64 KB per C2 group for
// 8 KB of output.  Here the group's tile STAYS in shared memory until its coefficients are known and the predictions
//   make_predictions          src/expressions.rs:175-195   (features . coefficients)
//   predictions *= 1/sqrt_w   polars_ols/least_squares.py:234-235
//   residuals = target - pred polars_ols/least_squares.py:238-239
// are computed straight from it — algorithmic bytes n (k + 1) s in + 8 n out.
//
// Same role split as gram_cta.cuh around mbarrier rings, with one more hand-off:
//   warp 0         PRODUCER  one tile == one group (1-D bulk async copies per column, `full` mbarrier)
//   warps 1..8     CONSUMERS Gram of group i (DMMA, as gram_cta); they only publish, they never release a stage
//   warps 9..12    SOLVERS   sum the published fragments, register Cholesky (LU fallback), beta -> HBM and -> the
//                            stage's shared beta slot, `beta` mbarrier
//   warps 13..16   PREDICTORS wait for `beta`, compute the group's predictions from the tile (two rows per lane and
//                            step, 16-byte shared loads and global stores) and release the stage (`empty`).  A first
//                            version let the consumers predict group i-1 after the Gram of group i: the stage was
//                            then held until the NEXT group had landed and the copy engine idled (0.24 ms for C2).
// A stage is held for load + Gram + solve + predict (~5 us for a C2 group), so three stages are needed to keep the
// copy engine busy; to make them fit next to 3 x 73 KB of tiles the eight consumer warps do not publish eight
// fragment sets (8 KB per ring entry) but accumulate them IN PLACE, in fixed order inside each half of four warps
// (named barriers), into two 1 KB half-sums that the solver adds — deterministic, 2 KB per ring entry.
#pragma once
#include "gram_cta.cuh"

namespace b200 {

constexpr int PRED_DEPTH = CTA_SOLVERS;  // ring of published half-sums; >= CTA_SOLVERS (parity waits, see gram_cta.cuh)
constexpr int PRED_WARPS = 4;            // predictor warps
constexpr int PRED_THREADS = CTA_THREADS + PRED_WARPS * 32;

struct PredOut {
    double *out;    // [n_rows] (groups are contiguous, packed order == original order)
    int residuals;  // 1: target - prediction
};

template <typename T>
__host__ __device__ inline size_t pred_fixed_smem(int KB, int F) {
    const size_t red = static_cast<size_t>(PRED_DEPTH) * 2 * 32 * (KB * (KB + 1) + KB) * sizeof(double);
    return red + CTA_SOLVERS * gram_scratch_bytes<T>(F, 1) + GRAM_MAX_STAGES * 16 * sizeof(double) + 128;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <typename T, int KB>
__global__ void __launch_bounds__(PRED_THREADS, 1) gram_pred_kernel(const GramParams p, const PredOut po) {
    using Vec = typename V2<T>::type;
    constexpr int NPAIR = KB * (KB + 1) / 2;
    constexpr int A = 16 / sizeof(T);
    constexpr int W = CTA_CONSUMERS;
    constexpr int RED = KB * (KB + 1) + KB;  // acc fragments + cy (the row count is the group length: no mask here)
    constexpr bool DUAL = KB <= 2;

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[GRAM_MAX_STAGES], empty_bar[GRAM_MAX_STAGES], beta_bar[GRAM_MAX_STAGES], red_full[PRED_DEPTH],
        red_empty[PRED_DEPTH];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fb = lane >> 2, q = lane & 3;
    const int kd = p.kd, F = p.F;
    const int ycol = kd, wcol = kd + 1;
    const int NC = kd + 1 + (p.has_w ? 1 : 0);
    const int R = p.tile_rows, S = p.stages;
    const uint32_t stride = gram_col_stride<T>(R);
    const uint32_t stage_bytes = static_cast<uint32_t>(NC) * stride;
    double *red = reinterpret_cast<double *>(smem + static_cast<size_t>(S) * stage_bytes);
    double *Gs_base = red + PRED_DEPTH * 2 * 32 * RED;
    double *beta_s = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(Gs_base) + CTA_SOLVERS * gram_scratch_bytes<T>(F, 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PRED_WARPS);
            mbar_init(&beta_bar[s], 1);
        }
        for (int b = 0; b < PRED_DEPTH; ++b) {
            mbar_init(&red_full[b], 2);
            mbar_init(&red_empty[b], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t nseg = p.nseg;

    if (warp == 0) {
        // ================================ PRODUCER: one (possibly empty) tile per group ================================
        int stage = 0;
        uint32_t phase = 0;
        int64_t nr0 = 0, nr1 = 0;
        if (static_cast<int64_t>(blockIdx.x) < nseg) { nr0 = p.seg_off[blockIdx.x]; nr1 = p.seg_off[blockIdx.x + 1]; }
        for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
            const int64_t r0 = nr0, r1 = nr1;
            if (seg + gridDim.x < nseg) { nr0 = p.seg_off[seg + gridDim.x]; nr1 = p.seg_off[seg + gridDim.x + 1]; }
            const int64_t a_al = r0 & ~static_cast<int64_t>(A - 1);
            int64_t b_al = (r1 + (A - 1)) & ~static_cast<int64_t>(A - 1);
            if (b_al > p.n_rows_pad) b_al = p.n_rows_pad;
            const uint32_t bytes = (r1 > r0) ? static_cast<uint32_t>(b_al - a_al) * sizeof(T) : 0u;
            mbar_wait(&empty_bar[stage], phase ^ 1u);  // predictions of the previous tenant are stored
            unsigned char *sb = smem + static_cast<size_t>(stage) * stage_bytes;
            if (lane == 0) {
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(&full_bar[stage], bytes * static_cast<uint32_t>(NC));
            }
            __syncwarp();
            if (bytes)
                for (int c = lane; c < NC; c += 32)
                    bulk_g2s(sb + static_cast<size_t>(c) * stride, static_cast<const T *>(p.cols[c]) + a_al, bytes, &full_bar[stage]);
            if (++stage == S) {
                stage = 0;
                phase ^= 1u;
            }
        }
    } else if (warp <= W) {
        // ================================ CONSUMERS ================================
        const int cw = warp - 1, half = cw >> 2, pos = cw & 3;
        int64_t nr0 = 0, nr1 = 0;
        if (static_cast<int64_t>(blockIdx.x) < nseg) { nr0 = p.seg_off[blockIdx.x]; nr1 = p.seg_off[blockIdx.x + 1]; }
        uint32_t li = 0;            // CTA-local group index

        for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x, ++li) {
            const int64_t r0 = nr0, r1 = nr1;
            if (seg + gridDim.x < nseg) { nr0 = p.seg_off[seg + gridDim.x]; nr1 = p.seg_off[seg + gridDim.x + 1]; }
            const int stage = static_cast<int>(li % S);
            double acc[NPAIR][2], acc2[DUAL ? NPAIR : 1][2];
            double cy[KB];
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
            for (int i = 0; i < (DUAL ? NPAIR : 1); ++i) acc2[i][0] = acc2[i][1] = 0.0;
#pragma unroll
            for (int i = 0; i < KB; ++i) cy[i] = 0.0;

            const int o = static_cast<int>(r0 & (A - 1));
            const int hi = o + static_cast<int>(r1 - r0);  // valid local rows are [o, hi)
            const int noct = (hi + 7) >> 3;
            const int per = (noct + W - 1) / W;
            const int j0 = cw * per;
            const int j1 = (j0 + per < noct) ? j0 + per : noct;
            mbar_wait(&full_bar[stage], (li / S) & 1u);
            const unsigned char *sb = smem + static_cast<size_t>(stage) * stage_bytes;
            const unsigned char *xs[KB];
            bool has_x[KB];
            double xconst[KB];
#pragma unroll
            for (int bk = 0; bk < KB; ++bk) {
                xs[bk] = sb + static_cast<size_t>((8 * bk + fb < kd) ? 8 * bk + fb : 0) * stride + 2 * q * sizeof(T);
                has_x[bk] = 8 * bk + fb < kd;
                xconst[bk] = ((8 * bk + fb == kd) && p.intercept) ? 1.0 : 0.0;
            }
            const unsigned char *ys = sb + static_cast<size_t>(ycol) * stride + 2 * q * sizeof(T);
            const unsigned char *ws = sb + static_cast<size_t>(wcol) * stride + 2 * q * sizeof(T);

            auto mma_octet = [&](const double (&f0)[KB], const double (&f1)[KB], double y0, double y1) {
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
            // one octet; EDGE: rows outside [o, hi) contribute zeros (same arithmetic as gram_cta's masked_octet)
            auto octet = [&](int j, bool edge) {
                const int lr = 8 * j + 2 * q;
                const bool v0 = !edge || ((lr >= o) && (lr < hi));
                const bool v1 = !edge || ((lr + 1 >= o) && (lr + 1 < hi));
                const Vec y2 = *reinterpret_cast<const Vec *>(ys + 8 * j * sizeof(T));
                T s0 = T(1), s1 = T(1);
                if (p.has_w) {
                    const Vec w2 = *reinterpret_cast<const Vec *>(ws + 8 * j * sizeof(T));
                    s0 = p.w_is_sqrt ? w2.x : static_cast<T>(sqrt(w2.x));
                    s1 = p.w_is_sqrt ? w2.y : static_cast<T>(sqrt(w2.y));
                }
                double f0[KB], f1[KB];
#pragma unroll
                for (int bk = 0; bk < KB; ++bk) {
                    const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                    const T x0 = has_x[bk] ? x2.x : static_cast<T>(xconst[bk]);
                    const T x1 = has_x[bk] ? x2.y : static_cast<T>(xconst[bk]);
                    f0[bk] = v0 ? static_cast<double>(static_cast<T>(x0 * s0)) : 0.0;
                    f1[bk] = v1 ? static_cast<double>(static_cast<T>(x1 * s1)) : 0.0;
                }
                mma_octet(f0, f1, v0 ? static_cast<double>(static_cast<T>(y2.x * s0)) : 0.0,
                          v1 ? static_cast<double>(static_cast<T>(y2.y * s1)) : 0.0);
            };
            {
                int j = j0;
                if (j < j1 && j == 0 && o != 0) {
                    octet(0, true);
                    j = 1;
                }
                const int jfull = hi >> 3;  // octets below jfull lie entirely inside [o, hi)
                const int jend = (j1 < jfull) ? j1 : jfull;
#pragma unroll 4
                for (; j < jend; ++j) octet(j, false);
                for (; j < j1; ++j) octet(j, true);
            }

            // ---- publish: in-place accumulation inside each half of four warps, fixed order (deterministic) ----
            const int buf = static_cast<int>(li % PRED_DEPTH);
            double *entry = red + static_cast<size_t>(buf * 2 + half) * 32 * RED;
            if (pos == 0) mbar_wait(&red_empty[buf], ((li / PRED_DEPTH) & 1u) ^ 1u);
#pragma unroll
            for (int step = 0; step < 4; ++step) {
                if (pos == step) {
#pragma unroll
                    for (int i = 0; i < NPAIR; ++i) {
                        const double a0 = DUAL ? acc[i][0] + acc2[i][0] : acc[i][0];
                        const double a1 = DUAL ? acc[i][1] + acc2[i][1] : acc[i][1];
                        entry[(2 * i) * 32 + lane] = (step == 0) ? a0 : entry[(2 * i) * 32 + lane] + a0;
                        entry[(2 * i + 1) * 32 + lane] = (step == 0) ? a1 : entry[(2 * i + 1) * 32 + lane] + a1;
                    }
#pragma unroll
                    for (int i = 0; i < KB; ++i)
                        entry[(2 * NPAIR + i) * 32 + lane] = (step == 0) ? cy[i] : entry[(2 * NPAIR + i) * 32 + lane] + cy[i];
                }
                named_bar_sync(1 + half, 128);
            }
            if (pos == 3 && lane == 0) mbar_arrive(&red_full[buf]);
        }
    } else if (warp <= W + CTA_SOLVERS) {
        // ================================ SOLVERS ================================
        const int sw = warp - (W + 1);
        double *Gs = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(Gs_base) + static_cast<size_t>(sw) * gram_scratch_bytes<T>(F, 1));
        uint32_t li = sw;
        for (int64_t seg = blockIdx.x + static_cast<int64_t>(sw) * gridDim.x; seg < nseg;
             seg += static_cast<int64_t>(CTA_SOLVERS) * gridDim.x, li += CTA_SOLVERS) {
            const int buf = static_cast<int>(li % PRED_DEPTH);
            mbar_wait(&red_full[buf], (li / PRED_DEPTH) & 1u);
            double acc[NPAIR][2], cy[KB];
            const double *e0 = red + static_cast<size_t>(buf * 2) * 32 * RED, *e1 = e0 + 32 * RED;
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
                acc[i][0] = e0[(2 * i) * 32 + lane] + e1[(2 * i) * 32 + lane];
                acc[i][1] = e0[(2 * i + 1) * 32 + lane] + e1[(2 * i + 1) * 32 + lane];
            }
#pragma unroll
            for (int i = 0; i < KB; ++i) cy[i] = e0[(2 * NPAIR + i) * 32 + lane] + e1[(2 * NPAIR + i) * 32 + lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_empty[buf]);
            const int nfit = (lane == 0) ? static_cast<int>(p.seg_off[seg + 1] - p.seg_off[seg]) : 0;
            gram_epilogue<KB>(p, acc, cy, nfit, seg, Gs, lane);  // beta -> HBM (+ flags), as gram_cta
            const int stage = static_cast<int>(li % S);
            if (lane < F) beta_s[stage * 16 + lane] = p.beta[seg * F + lane];  // this lane's own store
            __syncwarp();
            if (lane == 0) mbar_arrive(&beta_bar[stage]);
        }
    } else {
        // ================================ PREDICTORS ================================
        const int pw = warp - (W + CTA_SOLVERS + 1);
        const bool out_al = (reinterpret_cast<uintptr_t>(po.out) & 15u) == 0;
        int64_t nr0 = 0, nr1 = 0;
        if (static_cast<int64_t>(blockIdx.x) < nseg) { nr0 = p.seg_off[blockIdx.x]; nr1 = p.seg_off[blockIdx.x + 1]; }
        uint32_t li = 0;
        for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x, ++li) {
            const int64_t r0 = nr0, r1 = nr1;
            if (seg + gridDim.x < nseg) { nr0 = p.seg_off[seg + gridDim.x]; nr1 = p.seg_off[seg + gridDim.x + 1]; }
            const int stage = static_cast<int>(li % S);
            const uint32_t par = (li / S) & 1u;
            mbar_wait(&beta_bar[stage], par);
            mbar_wait(&full_bar[stage], par);  // completed long ago; orders this warp's reads after the bulk copies
            const unsigned char *sb = smem + static_cast<size_t>(stage) * stage_bytes;
            const double *bs = beta_s + stage * 16;
            double beta[8 * KB];
#pragma unroll
            for (int c = 0; c < 8 * KB; ++c) beta[c] = (c < F) ? bs[c] : 0.0;
            const double b_int = p.intercept ? bs[kd] : 0.0;
            const int o = static_cast<int>(r0 & (A - 1));
            const int hi = o + static_cast<int>(r1 - r0);  // valid local rows are [o, hi)
            const int64_t a_al = r0 - o;
            const int npair = (hi + 1) >> 1;
            for (int pp = pw * 32 + lane; pp < npair; pp += PRED_WARPS * 32) {
                const int lr = 2 * pp;
                const bool v0 = lr >= o && lr < hi, v1 = lr + 1 >= o && lr + 1 < hi;
                const size_t off = static_cast<size_t>(lr) * sizeof(T);
                T s0 = T(1), s1 = T(1);
                if (p.has_w) {
                    const Vec w2 = *reinterpret_cast<const Vec *>(sb + static_cast<size_t>(wcol) * stride + off);
                    s0 = p.w_is_sqrt ? w2.x : static_cast<T>(sqrt(w2.x));
                    s1 = p.w_is_sqrt ? w2.y : static_cast<T>(sqrt(w2.y));
                }
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int c = 0; c < 8 * KB; ++c)
                    if (c < kd) {
                        const Vec x2 = *reinterpret_cast<const Vec *>(sb + static_cast<size_t>(c) * stride + off);
                        a0 = fma(static_cast<double>(static_cast<T>(x2.x * s0)), beta[c], a0);
                        a1 = fma(static_cast<double>(static_cast<T>(x2.y * s1)), beta[c], a1);
                    }
                if (p.intercept) {
                    a0 = fma(static_cast<double>(s0), b_int, a0);
                    a1 = fma(static_cast<double>(s1), b_int, a1);
                }
                if (p.has_w) {  // predictions *= 1.0 / sqrt_w
                    a0 *= static_cast<double>(T(1) / s0);
                    a1 *= static_cast<double>(T(1) / s1);
                }
                if (po.residuals) {
                    const Vec y2 = *reinterpret_cast<const Vec *>(sb + static_cast<size_t>(ycol) * stride + off);
                    a0 = static_cast<double>(y2.x) - a0;
                    a1 = static_cast<double>(y2.y) - a1;
                }
                double *dst = po.out + a_al + lr;
                if (v0 && v1 && out_al) {
                    *reinterpret_cast<double2 *>(dst) = make_double2(a0, a1);
                } else {
                    if (v0) dst[0] = a0;
                    if (v1) dst[1] = a1;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
        }
    }
}

template <typename T, int KB>
cudaError_t gram_pred_launch_t(const GramParams &p, const PredOut &po, unsigned grid, size_t smem, cudaStream_t s) {
    auto kern = gram_pred_kernel<T, KB>;
    static size_t attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem > attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = smem;
    }
    kern<<<grid, PRED_THREADS, smem, s>>>(p, po);
    return cudaGetLastError();
}

cudaError_t gram_pred_launch_f64(int KB, const GramParams &p, const PredOut &po, unsigned grid, size_t smem, cudaStream_t s);
cudaError_t gram_pred_launch_f32(int KB, const GramParams &p, const PredOut &po, unsigned grid, size_t smem, cudaStream_t s);

}  // namespace b200
