// gram_pred.cuh — ONE kernel for mode = predictions | residuals of the static models when a whole group fits one
// shared-memory tile (k <= 16, null-free contiguous groups): Gram -> solve -> predictions, every input byte read from
// HBM once.
//
// The two-pass route (gram_cta_kernel, then predict_kernel) reads the features from HBM twice: 72 KB + 64 KB per C2
// group for 8 KB of output.  Here the group's predictions
//   make_predictions          src/expressions.rs:175-195   (features . coefficients)
//   predictions *= 1/sqrt_w   polars_ols/least_squares.py:234-235
//   residuals = target - pred polars_ols/least_squares.py:238-239
// are computed inside the streaming kernel as soon as its coefficients are known, re-reading the group's rows with
// plain 16-byte loads while they are still in L2 (they went through it a few microseconds earlier on their way to
// shared memory) — DRAM traffic n (k + 1) s in + 8 n out (ncu: dram read = 1.00 x the input bytes).
//
// Role split (gram_cta.cuh's pipeline plus one more role), 24 warps (k <= 8) or 16:
//   warp 0         PRODUCER   one tile == one group (1-D bulk async copies per column, `full` mbarrier); THROTTLED:
//                             group i is loaded only when group i - lag has been predicted, so that the rows a
//                             predictor re-reads are still in L2 (chip-wide footprint ~ 148 x lag x 73 KB)
//   warps 1..8     CONSUMERS  Gram of the group (DMMA, as gram_cta), release the stage, publish their fragments
//   warps 9..12    SOLVERS    (round-robin over groups) sum the fragments in fixed order, register Cholesky (LU
//                             fallback), beta -> HBM and -> the group's shared beta slot, `beta` mbarrier
//   warps 13..     PREDICTORS all work on one group at a time, in order: two rows per lane and load, straight from
//                             L2; streaming (evict-first) stores.  Each publishes its progress; the producer
//                             throttles on the slowest one (and a beta slot is reused only 16 groups later)
// What was tried first (profiles/r01_pred_ncu.txt): keeping the tile in shared memory until beta is known (consumers
// or dedicated warps predicting from it) holds a stage for load + Gram + solve + predict ~ 9 us, and with the three
// 73 KB stages that fit the copy engine idles (0.24 / 0.22 ms for C2, no better than two passes); predicting in the
// solver warp from L2 without the throttle lets the stream run a whole ring (~100 MB) ahead: 4 % L2 hits, 1.8 x DRAM.
#pragma once
#include "gram_cta.cuh"

namespace b200 {

// predictor warps: their loads are L2 round trips (~0.4 us under load), so what counts is how many are in flight —
// k <= 8: 11 warps x 8 loads per lane (24 warps, 80 registers each); k <= 16: 3 warps (16 warps, 128 registers)
__host__ __device__ constexpr int pred_predictors(int KB) { return KB == 1 ? 11 : 3; }
__host__ __device__ constexpr int pred_threads(int KB) { return (1 + CTA_CONSUMERS + CTA_SOLVERS + pred_predictors(KB)) * 32; }
constexpr int PRED_MAX_PREDICTORS = 11;
constexpr int PRED_DEPTH = CTA_SOLVERS;  // ring of published fragment sets; == CTA_SOLVERS (parity waits, see gram_cta.cuh)
constexpr int PRED_SLOTS = 16;           // ring of beta slots; > max lag, so a slot is never rewritten before it was read
constexpr int PRED_MAX_LAG = 12;

struct PredOut {
    double *out;    // [n_rows] (groups are contiguous, packed order == original order)
    int residuals;  // 1: target - prediction
    int lag;        // the producer loads group i only when groups <= i - lag are finished (bounds the L2 footprint)
};

template <int KB>
__host__ __device__ constexpr int pred_red_doubles() { return KB * (KB + 1) + KB; }  // acc fragments + cy

template <typename T>
__host__ __device__ inline size_t pred_fixed_smem(int KB, int F) {
    const size_t red = static_cast<size_t>(PRED_DEPTH) * CTA_CONSUMERS * 32 * (KB * (KB + 1) + KB) * sizeof(double);
    return red + CTA_SOLVERS * gram_scratch_bytes<T>(F, 1) + PRED_SLOTS * 16 * sizeof(double) + 128;
}

template <typename T, int KB>
__global__ void __launch_bounds__(pred_threads(KB), 1) gram_pred_kernel(const GramParams p, const PredOut po) {
    using Vec = typename V2<T>::type;
    constexpr int NPAIR = KB * (KB + 1) / 2;
    constexpr int A = 16 / sizeof(T);
    constexpr int W = CTA_CONSUMERS;
    constexpr int RED = pred_red_doubles<KB>();
    constexpr bool DUAL = KB <= 2;
    constexpr int NP = pred_predictors(KB);

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[GRAM_MAX_STAGES], empty_bar[GRAM_MAX_STAGES], red_full[PRED_DEPTH], red_empty[PRED_DEPTH],
        beta_bar[PRED_SLOTS];
    __shared__ unsigned int progress[PRED_MAX_PREDICTORS];  // groups finished by each predictor warp (they may drift apart)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fb = lane >> 2, q = lane & 3;
    const int kd = p.kd, F = p.F;
    const int ycol = kd, wcol = kd + 1;
    const int NC = kd + 1 + (p.has_w ? 1 : 0);
    const int R = p.tile_rows, S = p.stages;
    const uint32_t stride = gram_col_stride<T>(R);
    const uint32_t stage_bytes = static_cast<uint32_t>(NC) * stride;
    double *red = reinterpret_cast<double *>(smem + static_cast<size_t>(S) * stage_bytes);
    double *Gs_base = red + PRED_DEPTH * W * 32 * RED;
    double *beta_s = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(Gs_base) + CTA_SOLVERS * gram_scratch_bytes<T>(F, 1));

    if (threadIdx.x == 0) {
        for (int i = 0; i < PRED_MAX_PREDICTORS; ++i) progress[i] = 0;
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], W);
        }
        for (int b = 0; b < PRED_DEPTH; ++b) {
            mbar_init(&red_full[b], W);
            mbar_init(&red_empty[b], 1);
        }
        for (int b = 0; b < PRED_SLOTS; ++b) mbar_init(&beta_bar[b], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t nseg = p.nseg;

    if (warp == 0) {
        // ================================ PRODUCER: one (possibly empty) tile per group ================================
        int stage = 0;
        uint32_t phase = 0, li = 0;
        const uint32_t lag = po.lag < 1 ? 1u : (po.lag > PRED_MAX_LAG ? static_cast<uint32_t>(PRED_MAX_LAG) : static_cast<uint32_t>(po.lag));
        int64_t nr0 = 0, nr1 = 0;
        if (static_cast<int64_t>(blockIdx.x) < nseg) { nr0 = p.seg_off[blockIdx.x]; nr1 = p.seg_off[blockIdx.x + 1]; }
        for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x, ++li) {
            if (li >= lag) {  // throttle (see the header): every predictor warp has finished group li - lag
                for (;;) {
                    const unsigned int pr = (lane < NP) ? *reinterpret_cast<volatile unsigned int *>(&progress[lane]) : 0xffffffffu;
                    if (__all_sync(0xffffffffu, pr >= li - lag + 1)) break;
                    __nanosleep(20);
                }
            }
            const int64_t r0 = nr0, r1 = nr1;
            if (seg + gridDim.x < nseg) { nr0 = p.seg_off[seg + gridDim.x]; nr1 = p.seg_off[seg + gridDim.x + 1]; }
            const int64_t a_al = r0 & ~static_cast<int64_t>(A - 1);
            int64_t b_al = (r1 + (A - 1)) & ~static_cast<int64_t>(A - 1);
            if (b_al > p.n_rows_pad) b_al = p.n_rows_pad;
            const uint32_t bytes = (r1 > r0) ? static_cast<uint32_t>(b_al - a_al) * sizeof(T) : 0u;
            mbar_wait(&empty_bar[stage], phase ^ 1u);  // all consumers released this stage
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
        const int cw = warp - 1;
        int64_t nr0 = 0, nr1 = 0;
        if (static_cast<int64_t>(blockIdx.x) < nseg) { nr0 = p.seg_off[blockIdx.x]; nr1 = p.seg_off[blockIdx.x + 1]; }
        uint32_t li = 0;  // CTA-local group index
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x, ++li) {
            const int64_t r0 = nr0, r1 = nr1;
            if (seg + gridDim.x < nseg) { nr0 = p.seg_off[seg + gridDim.x]; nr1 = p.seg_off[seg + gridDim.x + 1]; }
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
            mbar_wait(&full_bar[stage], phase);
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
            if (p.has_w && !p.w_is_sqrt) {  // sqrt(w) once per row, in place (see gram_cta.cuh)
                T *wc = reinterpret_cast<T *>(const_cast<unsigned char *>(sb) + static_cast<size_t>(wcol) * stride);
                for (int i = 8 * j0 + lane; i < 8 * j1; i += 32) wc[i] = static_cast<T>(sqrt(wc[i]));
                __syncwarp();
            }

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
                    s0 = w2.x;
                    s1 = w2.y;
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
                if (!p.has_w) {  // no scaling: x * 1 is exact, skip the multiplies (as gram_cta's interior loop)
#pragma unroll 4
                    for (; j < jend; ++j) {
                        const Vec y2 = *reinterpret_cast<const Vec *>(ys + 8 * j * sizeof(T));
                        double f0[KB], f1[KB];
#pragma unroll
                        for (int bk = 0; bk < KB; ++bk) {
                            const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                            f0[bk] = has_x[bk] ? static_cast<double>(x2.x) : xconst[bk];
                            f1[bk] = has_x[bk] ? static_cast<double>(x2.y) : xconst[bk];
                        }
                        mma_octet(f0, f1, static_cast<double>(y2.x), static_cast<double>(y2.y));
                    }
                } else {
#pragma unroll 2
                    for (; j < jend; ++j) octet(j, false);
                }
                for (; j < j1; ++j) octet(j, true);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == S) {
                stage = 0;
                phase ^= 1u;
            }

            // ---- publish this warp's partial fragments (ring of PRED_DEPTH sets of W slots) ----
            const int buf = static_cast<int>(li % PRED_DEPTH);
            mbar_wait(&red_empty[buf], ((li / PRED_DEPTH) & 1u) ^ 1u);
            double *slot = red + (static_cast<size_t>(buf) * W + cw) * 32 * RED;
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
                slot[(2 * i) * 32 + lane] = DUAL ? acc[i][0] + acc2[i][0] : acc[i][0];
                slot[(2 * i + 1) * 32 + lane] = DUAL ? acc[i][1] + acc2[i][1] : acc[i][1];
            }
#pragma unroll
            for (int i = 0; i < KB; ++i) slot[(2 * NPAIR + i) * 32 + lane] = cy[i];
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_full[buf]);
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
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
            for (int i = 0; i < KB; ++i) cy[i] = 0.0;
            for (int cw = 0; cw < W; ++cw) {  // fixed order: deterministic sums
                const double *slot = red + (static_cast<size_t>(buf) * W + cw) * 32 * RED;
#pragma unroll
                for (int i = 0; i < NPAIR; ++i) {
                    acc[i][0] += slot[(2 * i) * 32 + lane];
                    acc[i][1] += slot[(2 * i + 1) * 32 + lane];
                }
#pragma unroll
                for (int i = 0; i < KB; ++i) cy[i] += slot[(2 * NPAIR + i) * 32 + lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_empty[buf]);
            const int nfit = (lane == 0) ? static_cast<int>(p.seg_off[seg + 1] - p.seg_off[seg]) : 0;
            double bl = 0.0;
            gram_epilogue<KB>(p, acc, cy, nfit, seg, Gs, lane, &bl);  // beta -> HBM (+ flags), as gram_cta; lane i keeps beta_i
            const int bslot = static_cast<int>(li % PRED_SLOTS);
            if (lane < 16) beta_s[bslot * 16 + lane] = (lane < F) ? bl : 0.0;
            __syncwarp();
            if (lane == 0) mbar_arrive(&beta_bar[bslot]);
        }
    } else {
        // ================================ PREDICTORS ================================
        const int pw = warp - (W + CTA_SOLVERS + 1);
        const int pl = pw * 32 + lane;
        constexpr int PL = NP * 32;
        const bool out_al = (reinterpret_cast<uintptr_t>(po.out) & 15u) == 0;
        const T *ycol_g = static_cast<const T *>(p.cols[ycol]);
        const T *wcol_g = static_cast<const T *>(p.cols[p.has_w ? wcol : ycol]);
        int64_t nr0 = 0, nr1 = 0;
        if (static_cast<int64_t>(blockIdx.x) < nseg) { nr0 = p.seg_off[blockIdx.x]; nr1 = p.seg_off[blockIdx.x + 1]; }
        uint32_t li = 0;
        for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x, ++li) {
            const int64_t r0 = nr0, r1 = nr1;
            if (seg + gridDim.x < nseg) { nr0 = p.seg_off[seg + gridDim.x]; nr1 = p.seg_off[seg + gridDim.x + 1]; }
            const int bslot = static_cast<int>(li % PRED_SLOTS);
            mbar_wait(&beta_bar[bslot], (li / PRED_SLOTS) & 1u);
            double beta[8 * KB];
#pragma unroll
            for (int c = 0; c < 8 * KB; ++c) beta[c] = beta_s[bslot * 16 + c];
            const double b_int = p.intercept ? beta_s[bslot * 16 + kd] : 0.0;

            const int o = static_cast<int>(r0 & (A - 1));
            const int hi = o + static_cast<int>(r1 - r0);  // valid local rows are [o, hi)
            const int64_t a_al = r0 - o;                   // multiple of A >= 2: vector loads / stores are aligned
            const int npair = (hi + 1) >> 1;
            // intercept, un-scaling, residual and the store of one row pair
            auto emit = [&](int pp, double a0, double a1, T s0, T s1, Vec y2) {
                const int lr = 2 * pp;
                const bool v0 = lr >= o && lr < hi, v1 = lr + 1 >= o && lr + 1 < hi;
                if (p.intercept) {
                    a0 = fma(static_cast<double>(s0), b_int, a0);
                    a1 = fma(static_cast<double>(s1), b_int, a1);
                }
                if (p.has_w) {  // predictions *= 1.0 / sqrt_w
                    a0 *= static_cast<double>(T(1) / s0);
                    a1 *= static_cast<double>(T(1) / s1);
                }
                if (po.residuals) {
                    a0 = static_cast<double>(y2.x) - a0;
                    a1 = static_cast<double>(y2.y) - a1;
                }
                double *dst = po.out + a_al + lr;
                if (v0 && v1 && out_al) {
                    __stcs(reinterpret_cast<double2 *>(dst), make_double2(a0, a1));
                } else {
                    if (v0) dst[0] = a0;
                    if (v1) dst[1] = a1;
                }
            };
            auto scales = [&](int pp, T &s0, T &s1) {
                s0 = s1 = T(1);
                if (p.has_w) {
                    const Vec w2 = __ldg(reinterpret_cast<const Vec *>(wcol_g + a_al + 2 * pp));
                    s0 = p.w_is_sqrt ? w2.x : static_cast<T>(sqrt(w2.x));
                    s1 = p.w_is_sqrt ? w2.y : static_cast<T>(sqrt(w2.y));
                }
            };
            int pp = pl;
            if (NP == 3 && KB == 1) {  // (unused with 11 predictor warps) two loads deep
                for (; pp + PL < npair; pp += 2 * PL) {
                    Vec xa[8], xb[8], ya, yb;
                    ya.x = ya.y = yb.x = yb.y = T(0);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < kd) {
                            xa[c] = __ldg(reinterpret_cast<const Vec *>(static_cast<const T *>(p.cols[c]) + a_al + 2 * pp));
                            xb[c] = __ldg(reinterpret_cast<const Vec *>(static_cast<const T *>(p.cols[c]) + a_al + 2 * (pp + PL)));
                        }
                    if (po.residuals) {
                        ya = __ldg(reinterpret_cast<const Vec *>(ycol_g + a_al + 2 * pp));
                        yb = __ldg(reinterpret_cast<const Vec *>(ycol_g + a_al + 2 * (pp + PL)));
                    }
                    T sa0, sa1, sb0, sb1;
                    scales(pp, sa0, sa1);
                    scales(pp + PL, sb0, sb1);
                    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < kd) {
                            a0 = fma(static_cast<double>(static_cast<T>(xa[c].x * sa0)), beta[c], a0);
                            a1 = fma(static_cast<double>(static_cast<T>(xa[c].y * sa1)), beta[c], a1);
                            b0 = fma(static_cast<double>(static_cast<T>(xb[c].x * sb0)), beta[c], b0);
                            b1 = fma(static_cast<double>(static_cast<T>(xb[c].y * sb1)), beta[c], b1);
                        }
                    emit(pp, a0, a1, sa0, sa1, ya);
                    emit(pp + PL, b0, b1, sb0, sb1, yb);
                }
            }
            for (; pp < npair; pp += PL) {
                T s0, s1;
                scales(pp, s0, s1);
                Vec y2;
                y2.x = y2.y = T(0);
                if (po.residuals) y2 = __ldg(reinterpret_cast<const Vec *>(ycol_g + a_al + 2 * pp));
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int c = 0; c < 8 * KB; ++c)
                    if (c < kd) {
                        const Vec x2 = __ldg(reinterpret_cast<const Vec *>(static_cast<const T *>(p.cols[c]) + a_al + 2 * pp));
                        a0 = fma(static_cast<double>(static_cast<T>(x2.x * s0)), beta[c], a0);
                        a1 = fma(static_cast<double>(static_cast<T>(x2.y * s1)), beta[c], a1);
                    }
                emit(pp, a0, a1, s0, s1, y2);
            }
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile unsigned int *>(&progress[pw]) = li + 1;
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
    kern<<<grid, pred_threads(KB), smem, s>>>(p, po);
    return cudaGetLastError();
}

cudaError_t gram_pred_launch_f64(int KB, const GramParams &p, const PredOut &po, unsigned grid, size_t smem, cudaStream_t s);
cudaError_t gram_pred_launch_f32(int KB, const GramParams &p, const PredOut &po, unsigned grid, size_t smem, cudaStream_t s);

}  // namespace b200
