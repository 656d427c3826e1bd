// gram_wide.cuh — row-streaming Gram kernel for 17 <= k <= 64 (C5: lasso on 64 features).
//
// Same producer / consumer mbarrier pipeline as gram_cta.cuh (one persistent CTA per SM, every column
// slice of a row tile fetched by one TMA bulk copy), but here the work of a tile is split across the
// consumer warps by OUTPUT BLOCK PAIRS instead of by rows: the k x k Gram is an 8 x 8 grid of 8 x 8 DMMA
// tiles; its upper triangle is cut into 2 x 2 super-blocks — 4 diagonal ones (3 distinct tile pairs, 2
// feature blocks each) and 6 off-diagonal ones (4 tile pairs, 4 feature blocks) — and consumer warp w owns
// super-block w for ALL rows.  A warp therefore keeps at most 8 accumulator registers, reads only the
// 2-4 feature blocks it needs from shared memory, and no cross-warp reduction is needed: at the end of a
// segment every warp writes its own tiles (and their transposes) of the partial record.  X^T y rides with
// the diagonal super-blocks.  At k = 64 the kernel is DMMA-bound (8 flop/B vs ~5.5 flop/B machine balance):
// 72 DMMA per 8 rows.
// The k x k solve (coordinate descent / Cholesky on the partials) runs in cd_solve.cuh / small_solve.cuh.
#pragma once
#include "gram_cta.cuh"

namespace b200 {

constexpr int WIDE_KB = 8;
constexpr int WIDE_SB = WIDE_KB / 2;                          // super-blocks per side
constexpr int WIDE_ITEMS = WIDE_SB + WIDE_SB * (WIDE_SB - 1) / 2;  // 4 diagonal + 6 off-diagonal = 10 consumer warps
constexpr int WIDE_THREADS = (WIDE_ITEMS + 1) * 32;

// EXTRA = weights and/or a row mask are present (predicated, scaled path); otherwise the loop body is
// 2-4 LDS.128 + 6-8 DMMA per 8 rows and warp.
template <typename T, bool EXTRA>
__global__ void __launch_bounds__(WIDE_THREADS, 1) gram_wide_kernel(const GramParams p) {
    using Vec = typename V2<T>::type;
    constexpr int A = 16 / sizeof(T);
    constexpr int W = WIDE_ITEMS;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[GRAM_MAX_STAGES], empty_bar[GRAM_MAX_STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fb = lane >> 2, q = lane & 3;
    const int kd = p.kd, F = p.F;
    const int ycol = kd, wcol = kd + 1, mcol = kd + 1 + (p.has_w ? 1 : 0);
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
    // item < SB: diagonal super-block a (feature blocks 2a, 2a+1); else off-diagonal (sa < sb2)
    int sa, sb2;
    if (item < WIDE_SB) {
        sa = sb2 = item;
    } else {
        int t = item - WIDE_SB;
        sa = 0;
        while (t >= WIDE_SB - 1 - sa) {
            t -= WIDE_SB - 1 - sa;
            ++sa;
        }
        sb2 = sa + 1 + t;
    }
    const bool diag = sa == sb2;
    const int nblk = (F + 7) >> 3;                // feature blocks actually present
    const int blk[4] = {2 * sa, 2 * sa + 1, 2 * sb2, 2 * sb2 + 1};
    // tile pairs of this item: diagonal (b0,b0) (b0,b1) (b1,b1); off-diagonal (b0,b2) (b0,b3) (b1,b2) (b1,b3)
    const bool active = blk[0] < nblk && (diag || blk[2] < nblk);  // super-blocks beyond F have nothing to do
    bool has_x[4];
    double xconst[4];
    uint32_t xoff[4];  // byte offset of this lane's element pair inside a stage, per feature block
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int f = 8 * blk[t] + fb;
        has_x[t] = f < kd;
        xoff[t] = static_cast<uint32_t>(has_x[t] ? f : 0) * stride + 2 * q * sizeof(T);  // padding lanes alias column 0
        xconst[t] = (f == kd && p.intercept) ? 1.0 : 0.0;
    }
    const uint32_t yoff = static_cast<uint32_t>(ycol) * stride + 2 * q * sizeof(T);
    const uint32_t woff = static_cast<uint32_t>(wcol) * stride + 2 * q * sizeof(T);
    const uint32_t moff = static_cast<uint32_t>(mcol) * stride + 2 * q * sizeof(T);
    const bool has_w = EXTRA && p.has_w, has_mask = EXTRA && p.has_mask;
    const bool plain = !has_mask;

    int stage = 0;
    uint32_t phase = 0;
    for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const int64_t r0 = p.seg_off[seg], r1 = p.seg_off[seg + 1];
        double acc[4][2], cy[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0;
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
                auto mma_all = [&](const double (&f0)[4], const double (&f1)[4]) {
                    if (diag) {
                        dmma_m8n8k4(acc[0][0], acc[0][1], f0[0], f0[0]);
                        dmma_m8n8k4(acc[1][0], acc[1][1], f0[0], f0[1]);
                        dmma_m8n8k4(acc[2][0], acc[2][1], f0[1], f0[1]);
                        dmma_m8n8k4(acc[0][0], acc[0][1], f1[0], f1[0]);
                        dmma_m8n8k4(acc[1][0], acc[1][1], f1[0], f1[1]);
                        dmma_m8n8k4(acc[2][0], acc[2][1], f1[1], f1[1]);
                    } else {
                        dmma_m8n8k4(acc[0][0], acc[0][1], f0[0], f0[2]);
                        dmma_m8n8k4(acc[1][0], acc[1][1], f0[0], f0[3]);
                        dmma_m8n8k4(acc[2][0], acc[2][1], f0[1], f0[2]);
                        dmma_m8n8k4(acc[3][0], acc[3][1], f0[1], f0[3]);
                        dmma_m8n8k4(acc[0][0], acc[0][1], f1[0], f1[2]);
                        dmma_m8n8k4(acc[1][0], acc[1][1], f1[0], f1[3]);
                        dmma_m8n8k4(acc[2][0], acc[2][1], f1[1], f1[2]);
                        dmma_m8n8k4(acc[3][0], acc[3][1], f1[1], f1[3]);
                    }
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
                    double f0[4], f1[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        f0[t] = f1[t] = 0.0;
                        if (t < 2 || !diag) {
                            const Vec x2 = *reinterpret_cast<const Vec *>(sbp + xoff[t] + jo);
                            const T x0 = has_x[t] ? x2.x : static_cast<T>(xconst[t]);
                            const T x1 = has_x[t] ? x2.y : static_cast<T>(xconst[t]);
                            f0[t] = v0 ? static_cast<double>(static_cast<T>(x0 * s0)) : 0.0;
                            f1[t] = v1 ? static_cast<double>(static_cast<T>(x1 * s1)) : 0.0;
                        }
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
                    if (diag) {
#pragma unroll 2
                        for (; j < jfull; ++j) {
                            const uint32_t jo = 8 * j * sizeof(T);
                            const Vec a2 = *reinterpret_cast<const Vec *>(sbp + xoff[0] + jo);
                            const Vec b2 = *reinterpret_cast<const Vec *>(sbp + xoff[1] + jo);
                            const Vec y2 = *reinterpret_cast<const Vec *>(sbp + yoff + jo);
                            const double a0 = has_x[0] ? static_cast<double>(a2.x) : xconst[0], a1 = has_x[0] ? static_cast<double>(a2.y) : xconst[0];
                            const double b0 = has_x[1] ? static_cast<double>(b2.x) : xconst[1], b1 = has_x[1] ? static_cast<double>(b2.y) : xconst[1];
                            dmma_m8n8k4(acc[0][0], acc[0][1], a0, a0);
                            dmma_m8n8k4(acc[1][0], acc[1][1], a0, b0);
                            dmma_m8n8k4(acc[2][0], acc[2][1], b0, b0);
                            dmma_m8n8k4(acc[0][0], acc[0][1], a1, a1);
                            dmma_m8n8k4(acc[1][0], acc[1][1], a1, b1);
                            dmma_m8n8k4(acc[2][0], acc[2][1], b1, b1);
                            const double y0 = static_cast<double>(y2.x), y1 = static_cast<double>(y2.y);
                            cy[0] = fma(a0, y0, fma(a1, y1, cy[0]));
                            cy[1] = fma(b0, y0, fma(b1, y1, cy[1]));
                        }
                    } else {
#pragma unroll 2
                        for (; j < jfull; ++j) {
                            const uint32_t jo = 8 * j * sizeof(T);
                            const Vec a2 = *reinterpret_cast<const Vec *>(sbp + xoff[0] + jo);
                            const Vec b2 = *reinterpret_cast<const Vec *>(sbp + xoff[1] + jo);
                            const Vec c2 = *reinterpret_cast<const Vec *>(sbp + xoff[2] + jo);
                            const Vec d2 = *reinterpret_cast<const Vec *>(sbp + xoff[3] + jo);
                            const double a0 = has_x[0] ? static_cast<double>(a2.x) : xconst[0], a1 = has_x[0] ? static_cast<double>(a2.y) : xconst[0];
                            const double b0 = has_x[1] ? static_cast<double>(b2.x) : xconst[1], b1 = has_x[1] ? static_cast<double>(b2.y) : xconst[1];
                            const double c0 = has_x[2] ? static_cast<double>(c2.x) : xconst[2], c1 = has_x[2] ? static_cast<double>(c2.y) : xconst[2];
                            const double d0 = has_x[3] ? static_cast<double>(d2.x) : xconst[3], d1 = has_x[3] ? static_cast<double>(d2.y) : xconst[3];
                            dmma_m8n8k4(acc[0][0], acc[0][1], a0, c0);
                            dmma_m8n8k4(acc[1][0], acc[1][1], a0, d0);
                            dmma_m8n8k4(acc[2][0], acc[2][1], b0, c0);
                            dmma_m8n8k4(acc[3][0], acc[3][1], b0, d0);
                            dmma_m8n8k4(acc[0][0], acc[0][1], a1, c1);
                            dmma_m8n8k4(acc[1][0], acc[1][1], a1, d1);
                            dmma_m8n8k4(acc[2][0], acc[2][1], b1, c1);
                            dmma_m8n8k4(acc[3][0], acc[3][1], b1, d1);
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
            if (diag) {
                put(2 * sa, 2 * sa, acc[0][0], acc[0][1]);
                put(2 * sa, 2 * sa + 1, acc[1][0], acc[1][1]);
                put(2 * sa + 1, 2 * sa + 1, acc[2][0], acc[2][1]);
            } else {
                put(2 * sa, 2 * sb2, acc[0][0], acc[0][1]);
                put(2 * sa, 2 * sb2 + 1, acc[1][0], acc[1][1]);
                put(2 * sa + 1, 2 * sb2, acc[2][0], acc[2][1]);
                put(2 * sa + 1, 2 * sb2 + 1, acc[3][0], acc[3][1]);
            }
            if (diag) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    double c = cy[t];
                    c += __shfl_xor_sync(0xffffffffu, c, 1);
                    c += __shfl_xor_sync(0xffffffffu, c, 2);
                    const int f = 8 * blk[t] + fb;
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
