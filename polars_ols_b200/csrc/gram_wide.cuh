// gram_wide.cuh — row-streaming Gram kernel for 17 <= k <= 64 (C5: lasso on 64 features).
//
// Same producer / consumer mbarrier pipeline as gram_cta.cuh (one persistent CTA per SM, every column
// slice of a row tile fetched by one TMA bulk copy), but here the work of a tile is split across the
// consumer warps by OUTPUT TILES instead of by rows: the k x k Gram is an 8 x 8 grid of 8 x 8 DMMA tiles;
// the 36 tiles of its upper triangle are cut into 12 triples (see WideItem) and consumer warp w owns triple w
// for ALL rows.  A warp therefore keeps 6 accumulator registers, reads only the 2-4 feature blocks it needs
// from shared memory, and no cross-warp reduction is needed: at the end of a segment every warp writes its
// own tiles (and their transposes) of the partial record.  X^T y rides with the four diagonal triples.  At k = 64 the kernel is DMMA-bound (8 flop/B vs ~5.5 flop/B machine balance):
// 72 DMMA per 8 rows.
// The k x k solve (coordinate descent / Cholesky on the partials) runs in cd_solve.cuh / small_solve.cuh.
#pragma once
#include "gram_cta.cuh"

namespace b200 {

constexpr int WIDE_KB = 8;
constexpr int WIDE_ITEMS = 12;  // consumer warps: 36 upper-triangle tiles, 3 per warp
constexpr int WIDE_THREADS = (WIDE_ITEMS + 1) * 32;

// Work split: the upper triangle of the 8 x 8 grid of 8 x 8 DMMA tiles (36 tiles) is cut into 12 triples, one per
// consumer warp, each touching at most 4 feature blocks (fragments): 6 DMMA + 2-4 LDS.128 per 8 rows and warp for
// EVERY warp.  With consumer warps 1..12 three of them land on each of the SM's four schedulers, i.e. 18 DMMA per 8
// rows on each FP64 tensor datapath.  (Round 1 split the triangle into 4 diagonal + 6 off-diagonal 2 x 2 super-blocks:
// 6 resp. 8 DMMA per warp, 22 / 22 / 14 / 14 per scheduler — the busiest datapath set the pace, 0.65 of the DMMA limit.)
//   shape 0  D  : frags [a, b]           tiles (a,a) (a,b) (b,b)      + X^T y of blocks a, b
//   shape 1  L1 : frags [r0, r1, c0, c1] tiles (r0,c0) (r0,c1) (r1,c0)
//   shape 2  L2 : frags [r0, r1, c0, c1] tiles (r0,c1) (r1,c0) (r1,c1)
//   shape 3  V  : frags [r0, r1, r2, c]  tiles (r0,c) (r1,c) (r2,c)
//   shape 4  L3 : frags [r0, r1, c0, c1] tiles (r0,c0) (r1,c0) (r1,c1)
struct WideItem {
    int shape;
    int blk[4];
};
__device__ __forceinline__ WideItem wide_item(int item) {
    switch (item) {
        case 0: return {0, {0, 1, 0, 0}};
        case 1: return {0, {2, 3, 0, 0}};
        case 2: return {0, {4, 5, 0, 0}};
        case 3: return {0, {6, 7, 0, 0}};
        case 4: return {1, {0, 1, 2, 3}};   // (0,2) (0,3) (1,2)
        case 5: return {2, {0, 1, 3, 4}};   // (0,4) (1,3) (1,4)
        case 6: return {1, {0, 1, 5, 6}};   // (0,5) (0,6) (1,5)
        case 7: return {2, {0, 1, 6, 7}};   // (0,7) (1,6) (1,7)
        case 8: return {1, {2, 3, 4, 5}};   // (2,4) (2,5) (3,4)
        case 9: return {2, {2, 3, 5, 6}};   // (2,6) (3,5) (3,6)
        case 10: return {3, {2, 3, 4, 7}};  // (2,7) (3,7) (4,7)
        default: return {4, {4, 5, 6, 7}};  // (4,6) (5,6) (5,7)
    }
}
// fragment indices (into WideItem::blk) of tile t of a shape; constexpr functions: usable in device code, folded
// away in the unrolled loops
__host__ __device__ constexpr int wide_pa(int shape, int t) {
    return shape == 0 ? (t == 2 ? 1 : 0)
         : shape == 1 ? (t == 2 ? 1 : 0)
         : shape == 2 ? (t == 0 ? 0 : 1)
         : shape == 3 ? t
                      : (t == 0 ? 0 : 1);
}
__host__ __device__ constexpr int wide_pb(int shape, int t) {
    return shape == 0 ? (t == 0 ? 0 : 1)
         : shape == 1 ? (t == 1 ? 3 : 2)
         : shape == 2 ? (t == 1 ? 2 : 3)
         : shape == 3 ? 3
                      : (t == 2 ? 3 : 2);
}
template <int SHAPE>
struct WidePairs {
    static constexpr int NF = SHAPE == 0 ? 2 : 4;
};

// consumer loop of one warp, specialised for its tile-triple shape
template <typename T, bool EXTRA, int SHAPE>
__device__ __forceinline__ void wide_consume(const GramParams &p, const WideItem &it, int item, unsigned char *smem, uint64_t *full_bar,
                                             uint64_t *empty_bar) {
    using Vec = typename V2<T>::type;
    using PR = WidePairs<SHAPE>;
    constexpr int A = 16 / sizeof(T);
    constexpr int NF = PR::NF;
    constexpr bool diag = SHAPE == 0;
    const int lane = threadIdx.x & 31;
    const int fb = lane >> 2, q = lane & 3;
    const int kd = p.kd, F = p.F;
    const int ycol = kd, wcol = kd + 1, mcol = kd + 1 + (p.has_w ? 1 : 0);
    const int NC = kd + 1 + (p.has_w ? 1 : 0) + (p.has_mask ? 1 : 0);
    const int R = p.tile_rows, S = p.stages;
    const uint32_t stride = gram_col_stride<T>(R);
    const uint32_t stage_bytes = static_cast<uint32_t>(NC) * stride;
    const int nblk = (F + 7) >> 3;  // feature blocks actually present
    bool tile_on[3];                // tiles beyond F have nothing to do
    bool active = false;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        tile_on[t] = it.blk[wide_pa(SHAPE, t)] < nblk && it.blk[wide_pb(SHAPE, t)] < nblk;
        active = active || tile_on[t];
    }
    bool has_x[NF];
    double xconst[NF];
    uint32_t xoff[NF];  // byte offset of this lane's element pair inside a stage, per fragment
#pragma unroll
    for (int t = 0; t < NF; ++t) {
        const int f = 8 * it.blk[t] + fb;
        has_x[t] = f < kd;
        xoff[t] = static_cast<uint32_t>(has_x[t] ? f : 0) * stride + 2 * q * sizeof(T);  // padding lanes alias column 0
        xconst[t] = (f == kd && p.intercept) ? 1.0 : 0.0;
    }
    const uint32_t yoff = static_cast<uint32_t>(ycol) * stride + 2 * q * sizeof(T);
    const uint32_t woff = static_cast<uint32_t>(wcol) * stride + 2 * q * sizeof(T);
    const uint32_t moff = static_cast<uint32_t>(mcol) * stride + 2 * q * sizeof(T);
    const bool has_w = EXTRA && p.has_w, has_mask = EXTRA && p.has_mask;
    const bool plain = !has_mask;
    const int64_t nseg = p.nseg;

    int stage = 0;
    uint32_t phase = 0;
    for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const int64_t r0 = p.seg_off[seg], r1 = p.seg_off[seg + 1];
        double acc[3][2], cy[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[i][0] = acc[i][1] = 0.0;
        cy[0] = cy[1] = 0.0;
        int nfit = 0;
        for (int64_t row = r0; row < r1; row += R) {
            const int64_t b = (row + R < r1) ? row + R : r1;
            const int o = static_cast<int>(row & (A - 1));
            const int hi = o + static_cast<int>(b - row);
            const int noct = (hi + 7) >> 3;
            mbar_wait(&full_bar[stage], phase);
            if (active) {
                const unsigned char *sbp = smem + static_cast<size_t>(stage) * stage_bytes;
                auto mma_all = [&](const double (&f0)[NF], const double (&f1)[NF]) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) dmma_m8n8k4(acc[t][0], acc[t][1], f0[wide_pa(SHAPE, t)], f0[wide_pb(SHAPE, t)]);
#pragma unroll
                    for (int t = 0; t < 3; ++t) dmma_m8n8k4(acc[t][0], acc[t][1], f1[wide_pa(SHAPE, t)], f1[wide_pb(SHAPE, t)]);
                };
                // predicated / scaled octet (segment edges, weights, row mask)
                auto masked_octet = [&](int j) {
                    const int lr = 8 * j + 2 * q;
                    bool v0 = (lr >= o) && (lr < hi);
                    bool v1 = (lr + 1 >= o) && (lr + 1 < hi);
                    const uint32_t jo = 8 * j * sizeof(T);
                    if (has_mask) {
                        const Vec m2 = *reinterpret_cast<const Vec *>(sbp + moff + jo);
                        v0 = v0 && (m2.x != T(0));
                        v1 = v1 && (m2.y != T(0));
                    }
                    T s0 = T(1), s1 = T(1);
                    if (has_w) {
                        const Vec w2 = *reinterpret_cast<const Vec *>(sbp + woff + jo);
                        s0 = p.w_is_sqrt ? w2.x : static_cast<T>(sqrt(w2.x));
                        s1 = p.w_is_sqrt ? w2.y : static_cast<T>(sqrt(w2.y));
                    }
                    double f0[NF], f1[NF];
#pragma unroll
                    for (int t = 0; t < NF; ++t) {
                        const Vec x2 = *reinterpret_cast<const Vec *>(sbp + xoff[t] + jo);
                        const T x0 = has_x[t] ? x2.x : static_cast<T>(xconst[t]);
                        const T x1 = has_x[t] ? x2.y : static_cast<T>(xconst[t]);
                        f0[t] = v0 ? static_cast<double>(static_cast<T>(x0 * s0)) : 0.0;
                        f1[t] = v1 ? static_cast<double>(static_cast<T>(x1 * s1)) : 0.0;
                    }
                    mma_all(f0, f1);
                    if (diag) {
                        const Vec y2 = *reinterpret_cast<const Vec *>(sbp + yoff + jo);
                        const double y0 = v0 ? static_cast<double>(static_cast<T>(y2.x * s0)) : 0.0;
                        const double y1 = v1 ? static_cast<double>(static_cast<T>(y2.y * s1)) : 0.0;
                        cy[0] = fma(f0[0], y0, fma(f1[0], y1, cy[0]));
                        cy[1] = fma(f0[1], y0, fma(f1[1], y1, cy[1]));
                        if (item == 0 && fb == 0) nfit += (v0 ? 1 : 0) + (v1 ? 1 : 0);
                    }
                };
                if (EXTRA) {
                    for (int j = 0; j < noct; ++j) masked_octet(j);
                } else {
                    int j = 0;
                    if (o != 0) {
                        masked_octet(0);
                        j = 1;
                    }
                    const int jfull = hi >> 3;
#pragma unroll 2
                    for (; j < jfull; ++j) {
                        const uint32_t jo = 8 * j * sizeof(T);
                        double f0[NF], f1[NF];
#pragma unroll
                        for (int t = 0; t < NF; ++t) {
                            const Vec x2 = *reinterpret_cast<const Vec *>(sbp + xoff[t] + jo);
                            f0[t] = has_x[t] ? static_cast<double>(x2.x) : xconst[t];
                            f1[t] = has_x[t] ? static_cast<double>(x2.y) : xconst[t];
                        }
                        mma_all(f0, f1);
                        if (diag) {
                            const Vec y2 = *reinterpret_cast<const Vec *>(sbp + yoff + jo);
                            const double y0 = static_cast<double>(y2.x), y1 = static_cast<double>(y2.y);
                            cy[0] = fma(f0[0], y0, fma(f1[0], y1, cy[0]));
                            cy[1] = fma(f0[1], y0, fma(f1[1], y1, cy[1]));
                        }
                    }
                    for (; j < noct; ++j) masked_octet(j);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == S) {
                stage = 0;
                phase ^= 1u;
            }
        }
        // ---- write this warp's tiles of the partial record (G row-major F x F, then c, then n_fit) ----
        double *out = p.partial + static_cast<size_t>(seg) * (static_cast<size_t>(F) * F + F + 1);
        if (active) {
            auto put = [&](int bi, int bj, double a0, double a1) {
                const int rr = 8 * bi + fb, cc = 8 * bj + 2 * q;
                if (rr < F) {
                    if (cc < F) out[rr * F + cc] = a0;
                    if (cc + 1 < F) out[rr * F + cc + 1] = a1;
                    if (bi != bj) {
                        if (cc < F) out[cc * F + rr] = a0;
                        if (cc + 1 < F) out[(cc + 1) * F + rr] = a1;
                    }
                }
            };
#pragma unroll
            for (int t = 0; t < 3; ++t)
                if (tile_on[t]) put(it.blk[wide_pa(SHAPE, t)], it.blk[wide_pb(SHAPE, t)], acc[t][0], acc[t][1]);
            if (diag) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    double c = cy[t];
                    c += __shfl_xor_sync(0xffffffffu, c, 1);
                    c += __shfl_xor_sync(0xffffffffu, c, 2);
                    const int f = 8 * it.blk[t] + fb;
                    if (q == 0 && f < F) out[F * F + f] = c;
                }
                if (item == 0) {
                    nfit += __shfl_xor_sync(0xffffffffu, nfit, 1);
                    nfit += __shfl_xor_sync(0xffffffffu, nfit, 2);
                    if (lane == 0) out[F * F + F] = static_cast<double>(plain ? static_cast<int>(r1 - r0) : nfit);
                }
            }
        }
    }
}

// EXTRA = weights and/or a row mask are present (predicated, scaled path); otherwise the loop body is
// 2-4 LDS.128 + 6 DMMA per 8 rows and warp.
template <typename T, bool EXTRA>
__global__ void __launch_bounds__(WIDE_THREADS, 1) gram_wide_kernel(const GramParams p) {
    constexpr int A = 16 / sizeof(T);
    constexpr int W = WIDE_ITEMS;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[GRAM_MAX_STAGES], empty_bar[GRAM_MAX_STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kd = p.kd;
    const int NC = kd + 1 + (p.has_w ? 1 : 0) + (p.has_mask ? 1 : 0);
    const int R = p.tile_rows, S = p.stages;
    const uint32_t stride = gram_col_stride<T>(R);
    const uint32_t stage_bytes = static_cast<uint32_t>(NC) * stride;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], W);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int64_t nseg = p.nseg;

    if (warp == 0) {
        // ================================ PRODUCER ================================
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
            const int64_t r0 = p.seg_off[seg], r1 = p.seg_off[seg + 1];
            for (int64_t row = r0; row < r1; row += R) {
                const int64_t b = (row + R < r1) ? row + R : r1;
                const int64_t a_al = row & ~static_cast<int64_t>(A - 1);
                int64_t b_al = (b + (A - 1)) & ~static_cast<int64_t>(A - 1);
                if (b_al > p.n_rows_pad) b_al = p.n_rows_pad;
                const uint32_t bytes = static_cast<uint32_t>(b_al - a_al) * sizeof(T);
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                unsigned char *sb = smem + static_cast<size_t>(stage) * stage_bytes;
                if (lane == 0) {
                    fence_proxy_async_smem();
                    mbar_arrive_expect_tx(&full_bar[stage], bytes * static_cast<uint32_t>(NC));
                }
                __syncwarp();
                for (int c = lane; c < NC; c += 32)
                    bulk_g2s(sb + static_cast<size_t>(c) * stride, static_cast<const T *>(p.cols[c]) + a_al, bytes,
                             &full_bar[stage]);
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        return;
    }

    // ================================ CONSUMERS ================================
    const int item = warp - 1;
    const WideItem it = wide_item(item);
    switch (it.shape) {  // warp-uniform
        case 0: wide_consume<T, EXTRA, 0>(p, it, item, smem, full_bar, empty_bar); break;
        case 1: wide_consume<T, EXTRA, 1>(p, it, item, smem, full_bar, empty_bar); break;
        case 2: wide_consume<T, EXTRA, 2>(p, it, item, smem, full_bar, empty_bar); break;
        case 3: wide_consume<T, EXTRA, 3>(p, it, item, smem, full_bar, empty_bar); break;
        default: wide_consume<T, EXTRA, 4>(p, it, item, smem, full_bar, empty_bar); break;
    }
}

template <typename T, bool EXTRA>
cudaError_t gram_wide_launch_e(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s) {
    auto kern = gram_wide_kernel<T, EXTRA>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, WIDE_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

template <typename T>
cudaError_t gram_wide_launch_t(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s) {
    return (p.has_w || p.has_mask) ? gram_wide_launch_e<T, true>(p, grid, smem, s) : gram_wide_launch_e<T, false>(p, grid, smem, s);
}

cudaError_t gram_wide_launch_f64(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s);
cudaError_t gram_wide_launch_f32(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s);

}  // namespace b200
