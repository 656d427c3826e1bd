// gram_ldg.cuh — direct-load variant of the row-streaming Gram kernel (same maths, same epilogue as
// gram_stream.cuh): every lane fetches its DMMA fragment element pairs straight from HBM with 16-byte
// (f64) / 8-byte (f32) loads, `U` row octets in flight per lane, no shared-memory staging.
// One warp per segment.  The interior of a segment runs a mask-free loop (2 loads + 2 DMMA + 2 DFMA per
// 8 rows and lane); only the first / last chunk of a segment and frames with weights or a row mask take
// the predicated path.
#pragma once
#include "gram_stream.cuh"

namespace b200 {

template <typename T, int KB, int U, int MAXW, bool EXTRA>
__global__ void __launch_bounds__(MAXW * 32) gram_ldg_kernel(const GramParams p) {
    using Vec = typename V2<T>::type;
    constexpr int NPAIR = KB * (KB + 1) / 2;
    constexpr int A = 16 / sizeof(T);
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int fb = lane >> 2, q = lane & 3;
    const int kd = p.kd, F = p.F;
    double *Gs = reinterpret_cast<double *>(smem + static_cast<size_t>(warp) * gram_scratch_bytes<T>(F, p.fused));
    const T *ycol = static_cast<const T *>(p.cols[kd]);
    // EXTRA = weights and/or row mask present (keeps their registers out of the plain fast path)
    const T *wcol = (EXTRA && p.has_w) ? static_cast<const T *>(p.cols[kd + 1]) : nullptr;
    const T *mcol = (EXTRA && p.has_mask) ? static_cast<const T *>(p.cols[kd + 1 + (p.has_w ? 1 : 0)]) : nullptr;
    // lanes whose feature slot is padding (or the synthetic intercept) still load — from column 0, the same
    // sectors the fb = 0 lanes fetch anyway — so the streaming loop has no divergent branch
    const T *xcol[KB];
    bool has_x[KB];
    double xconst[KB];
#pragma unroll
    for (int bk = 0; bk < KB; ++bk) {
        has_x[bk] = 8 * bk + fb < kd;
        xcol[bk] = static_cast<const T *>(p.cols[has_x[bk] ? 8 * bk + fb : 0]);
        xconst[bk] = ((8 * bk + fb == kd) && p.intercept) ? 1.0 : 0.0;
    }
    const int64_t wg = static_cast<int64_t>(blockIdx.x) * W + warp;
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * W;

    for (int64_t seg = wg; seg < p.nseg; seg += nwarps) {
        const int64_t r0 = p.seg_off[seg], r1 = p.seg_off[seg + 1];
        const int64_t a_al = r0 & ~static_cast<int64_t>(A - 1);
        const int head = static_cast<int>(r0 - a_al);          // local rows [head, total) are valid
        const int total = head + static_cast<int>(r1 - r0);
        const int last = static_cast<int>(p.n_rows_pad - a_al) - 2;  // last loadable local row pair
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
        const T *yb = ycol + a_al + 2 * q;
        const T *wb = wcol ? wcol + a_al + 2 * q : nullptr;
        const T *mb = mcol ? mcol + a_al + 2 * q : nullptr;
        const T *xb[KB];
#pragma unroll
        for (int bk = 0; bk < KB; ++bk) xb[bk] = xcol[bk] + a_al + 2 * q;

        // predicated chunk: rows [off, off + 8U) against [head, total), loads clamped to the array end
        auto masked_chunk = [&](int off) {
            Vec xv[U][KB], yv[U], wv[EXTRA ? U : 1], mv[EXTRA ? U : 1];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int lr = off + 8 * u;
                if (lr + 2 * q > last) lr = last - 2 * q;
                yv[u] = *reinterpret_cast<const Vec *>(yb + lr);
                if (EXTRA) {
                    if (wb) wv[u] = *reinterpret_cast<const Vec *>(wb + lr);
                    if (mb) mv[u] = *reinterpret_cast<const Vec *>(mb + lr);
                }
#pragma unroll
                for (int bk = 0; bk < KB; ++bk) xv[u][bk] = *reinterpret_cast<const Vec *>(xb[bk] + lr);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int lr = off + 8 * u + 2 * q;
                bool v0 = (lr >= head) && (lr < total), v1 = (lr + 1 >= head) && (lr + 1 < total);
                T s0 = T(1), s1 = T(1);
                if (EXTRA) {
                    if (mb) {
                        v0 = v0 && (mv[u].x != T(0));
                        v1 = v1 && (mv[u].y != T(0));
                    }
                    if (wb) {
                        s0 = p.w_is_sqrt ? wv[u].x : static_cast<T>(sqrt(wv[u].x));
                        s1 = p.w_is_sqrt ? wv[u].y : static_cast<T>(sqrt(wv[u].y));
                    }
                }
                const double y0 = v0 ? static_cast<double>(static_cast<T>(yv[u].x * s0)) : 0.0;
                const double y1 = v1 ? static_cast<double>(static_cast<T>(yv[u].y * s1)) : 0.0;
                if (fb == 0) nfit += (v0 ? 1 : 0) + (v1 ? 1 : 0);
                double f0[KB], f1[KB];
#pragma unroll
                for (int bk = 0; bk < KB; ++bk) {
                    const T x0 = has_x[bk] ? xv[u][bk].x : static_cast<T>(xconst[bk]);
                    const T x1 = has_x[bk] ? xv[u][bk].y : static_cast<T>(xconst[bk]);
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
            }
        };

        int off = 0;
        if (EXTRA) {
            for (; off < total; off += 8 * U) masked_chunk(off);
        } else {
            if (head != 0) {
                masked_chunk(0);
                off = 8 * U;
            }
            // mask-free interior
            for (; off + 8 * U <= total; off += 8 * U) {
                Vec xv[U][KB], yv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    yv[u] = *reinterpret_cast<const Vec *>(yb + off + 8 * u);
#pragma unroll
                    for (int bk = 0; bk < KB; ++bk) xv[u][bk] = *reinterpret_cast<const Vec *>(xb[bk] + off + 8 * u);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double f0[KB], f1[KB];
#pragma unroll
                    for (int bk = 0; bk < KB; ++bk) {
                        f0[bk] = has_x[bk] ? static_cast<double>(xv[u][bk].x) : xconst[bk];
                        f1[bk] = has_x[bk] ? static_cast<double>(xv[u][bk].y) : xconst[bk];
                    }
                    const double y0 = static_cast<double>(yv[u].x), y1 = static_cast<double>(yv[u].y);
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
            }
            if (off < total) masked_chunk(off);
            nfit = 0;  // recomputed below: without a mask every row of the segment is fitted
        }
#pragma unroll
        for (int i = 0; i < (DUAL ? NPAIR : 0); ++i) {
            acc[i][0] += acc2[i][0];
            acc[i][1] += acc2[i][1];
        }
        if (!EXTRA) nfit = (fb == 0 && q == 0) ? (total - head) : 0;
        gram_epilogue<KB>(p, acc, cy, nfit, seg, Gs, lane);
    }
}

template <typename T, int KB, int U, bool EXTRA>
cudaError_t gram_ldg_launch_e(const GramParams &p, unsigned grid, int warps, cudaStream_t s) {
    auto kern = gram_ldg_kernel<T, KB, U, 8, EXTRA>;
    const size_t smem = static_cast<size_t>(warps) * gram_scratch_bytes<T>(p.F, p.fused);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, warps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

template <typename T, int KB, int U>
cudaError_t gram_ldg_launch_t(const GramParams &p, unsigned grid, int warps, cudaStream_t s) {
    return (p.has_w || p.has_mask) ? gram_ldg_launch_e<T, KB, (U > 4 ? 4 : U), true>(p, grid, warps, s)
                                   : gram_ldg_launch_e<T, KB, U, false>(p, grid, warps, s);
}

// only KB <= 2 (k <= 16) is instantiated: wider problems use the staged kernel
template <typename T>
cudaError_t gram_ldg_launch_any(int KB, int U, const GramParams &p, unsigned grid, int warps, cudaStream_t s) {
    if (KB == 1) {
        switch (U) {
            case 1: return gram_ldg_launch_t<T, 1, 1>(p, grid, warps, s);
            case 2: return gram_ldg_launch_t<T, 1, 2>(p, grid, warps, s);
            case 4: return gram_ldg_launch_t<T, 1, 4>(p, grid, warps, s);
            default: return gram_ldg_launch_t<T, 1, 8>(p, grid, warps, s);
        }
    }
    return (U <= 2) ? gram_ldg_launch_t<T, 2, 2>(p, grid, warps, s) : gram_ldg_launch_t<T, 2, 4>(p, grid, warps, s);
}

cudaError_t gram_ldg_launch_f64(int KB, int U, const GramParams &p, unsigned grid, int warps, cudaStream_t s);
cudaError_t gram_ldg_launch_f32(int KB, int U, const GramParams &p, unsigned grid, int warps, cudaStream_t s);

}  // namespace b200
