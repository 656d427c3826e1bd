// gram_multi.cuh — row-streaming Gram kernel for SHORT groups (k <= 16, a group is at most one tile).
//
// gram_cta_kernel spends one barrier round trip, one set of bulk copies and one accumulator hand-off per group;
// with groups of a few hundred rows that fixed cost, not HBM, sets the pace (n = 64: 0.9 TB/s).  Here a TILE is
// a run of consecutive WHOLE groups (their rows are contiguous in the packed frame), fetched by one bulk async
// copy per column, and every consumer warp owns whole groups of the tile:
//   warp 0      PRODUCER : tile table -> k+1(+w)(+mask) bulk copies (cp.async.bulk / UBLKCP) per tile into the
//                          stage ring (full / empty mbarriers, every consumer releases every stage: the plain
//                          ring, each barrier is waited on by the same warps in every phase);
//   warps 1..W  CONSUMERS: group j of the tile belongs to warp (first_group + j) % W; the warp streams the
//                          group's row octets through mma.sync.m8n8k4.f64 (DMMA; A = X^T and B = X fragments are
//                          the same registers), reduces X^T y over the four lanes of a feature and writes the
//                          [k*k + k + 1] Gram record straight from registers.  No cross-warp reduction, no
//                          solver warps: the records are solved by batch_solve_kernel / cd_solve_kernel at full
//                          occupancy.  Warps without a group in a tile release it at once and run ahead, so up to
//                          STAGES tiles are being consumed concurrently.
// Reference arithmetic restated: X^T X, X^T y of solve_ridge / solve_elastic_net (src/least_squares.rs:352-354,
// :417) with the WLS scaling of polars_ols/least_squares.py:189-196 applied on load in the column dtype.
#pragma once
#include "gram_stream.cuh"

namespace b200 {

constexpr int MULTI_CONSUMERS = 12;
constexpr int MULTI_THREADS = (MULTI_CONSUMERS + 1) * 32;

struct MultiPlan {
    const int64_t *tile_group;  // device [n_tiles + 1]: tile t holds groups [tile_group[t], tile_group[t + 1])
    int64_t n_tiles;
};

template <typename T, int KB>
__global__ void __launch_bounds__(MULTI_THREADS, 1) gram_multi_kernel(const GramParams p, const MultiPlan mp) {
    using Vec = typename V2<T>::type;
    constexpr int NPAIR = KB * (KB + 1) / 2;
    constexpr int A = 16 / sizeof(T);
    constexpr int W = MULTI_CONSUMERS;

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[GRAM_MAX_STAGES], empty_bar[GRAM_MAX_STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fb = lane >> 2, q = lane & 3;
    const int kd = p.kd;
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

    if (warp == 0) {
        // ================================ PRODUCER ================================
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < mp.n_tiles; t += gridDim.x) {
            const int64_t g0 = mp.tile_group[t], g1 = mp.tile_group[t + 1];
            const int64_t r0 = p.seg_off[g0], r1 = p.seg_off[g1];
            const int64_t a_al = r0 & ~static_cast<int64_t>(A - 1);
            int64_t b_al = (r1 + (A - 1)) & ~static_cast<int64_t>(A - 1);
            if (b_al > p.n_rows_pad) b_al = p.n_rows_pad;
            const uint32_t bytes = static_cast<uint32_t>(b_al - a_al) * sizeof(T);
            mbar_wait(&empty_bar[stage], phase ^ 1u);
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
        return;
    }
    // ================================ CONSUMERS ================================
    const int cw = warp - 1;
    int stage = 0;
    uint32_t phase = 0;
    const bool plain = !p.has_mask;
    bool has_x[KB];
    double xconst[KB];
#pragma unroll
    for (int bk = 0; bk < KB; ++bk) {
        has_x[bk] = 8 * bk + fb < kd;
        xconst[bk] = ((8 * bk + fb == kd) && p.intercept) ? 1.0 : 0.0;
    }
    for (int64_t t = blockIdx.x; t < mp.n_tiles; t += gridDim.x) {
        const int64_t g0 = mp.tile_group[t], g1 = mp.tile_group[t + 1];
        const int64_t base = p.seg_off[g0] & ~static_cast<int64_t>(A - 1);  // packed row of the tile's local row 0
        // first group of this tile that belongs to this warp: (g0 + j) % W == cw
        int64_t g = g0 + ((cw - static_cast<int>(g0 % W)) + W) % W;
        mbar_wait(&full_bar[stage], phase);
        const unsigned char *sb = smem + static_cast<size_t>(stage) * stage_bytes;
        const unsigned char *xs[KB];
#pragma unroll
        for (int bk = 0; bk < KB; ++bk)
            xs[bk] = sb + static_cast<size_t>((8 * bk + fb < kd) ? 8 * bk + fb : 0) * stride + 2 * q * sizeof(T);
        const unsigned char *ys = sb + static_cast<size_t>(ycol) * stride + 2 * q * sizeof(T);
        const unsigned char *ws = sb + static_cast<size_t>(wcol) * stride + 2 * q * sizeof(T);
        for (; g < g1; g += W) {
            const int o = static_cast<int>(p.seg_off[g] - base), hi = static_cast<int>(p.seg_off[g + 1] - base);  // local rows [o, hi)
            if (p.has_w && !p.w_is_sqrt) {  // sqrt(w) once per row, in place, this group's rows only (see gram_cta.cuh)
                T *wc = reinterpret_cast<T *>(const_cast<unsigned char *>(sb) + static_cast<size_t>(wcol) * stride);
                for (int i = o + lane; i < hi; i += 32) wc[i] = static_cast<T>(sqrt(wc[i]));
                __syncwarp();
            }
            double acc[NPAIR][2], acc2[NPAIR][2], cy[KB];
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) acc[i][0] = acc[i][1] = acc2[i][0] = acc2[i][1] = 0.0;
#pragma unroll
            for (int i = 0; i < KB; ++i) cy[i] = 0.0;
            int nfit = 0;
            auto mma_octet = [&](const double (&f0)[KB], const double (&f1)[KB], double y0, double y1) {
                int idx = 0;
#pragma unroll
                for (int bi = 0; bi < KB; ++bi) {
#pragma unroll
                    for (int bj = bi; bj < KB; ++bj) {
                        dmma_m8n8k4(acc[idx][0], acc[idx][1], f0[bi], f0[bj]);
                        dmma_m8n8k4(acc2[idx][0], acc2[idx][1], f1[bi], f1[bj]);
                        ++idx;
                    }
                    cy[bi] = fma(f0[bi], y0, cy[bi]);
                    cy[bi] = fma(f1[bi], y1, cy[bi]);
                }
            };
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
                    const Vec w2 = *reinterpret_cast<const Vec *>(ws + 8 * j * sizeof(T));
                    s0 = w2.x;
                    s1 = w2.y;
                }
                const double y0 = v0 ? static_cast<double>(static_cast<T>(y2.x * s0)) : 0.0;
                const double y1 = v1 ? static_cast<double>(static_cast<T>(y2.y * s1)) : 0.0;
                if (fb == 0) nfit += (v0 ? 1 : 0) + (v1 ? 1 : 0);
                double f0[KB], f1[KB];
#pragma unroll
                for (int bk = 0; bk < KB; ++bk) {
                    const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                    const T x0 = has_x[bk] ? x2.x : static_cast<T>(xconst[bk]);
                    const T x1 = has_x[bk] ? x2.y : static_cast<T>(xconst[bk]);
                    f0[bk] = v0 ? static_cast<double>(static_cast<T>(x0 * s0)) : 0.0;
                    f1[bk] = v1 ? static_cast<double>(static_cast<T>(x1 * s1)) : 0.0;
                }
                mma_octet(f0, f1, y0, y1);
            };
            const int jlo = o >> 3, jhi = (hi + 7) >> 3;  // octets [jlo, jhi) touch the group
            if (!plain) {
                for (int j = jlo; j < jhi; ++j) masked_octet(j);
            } else {
                int j = jlo;
                if (j < jhi && (o & 7)) masked_octet(j++);
                const int jfull = hi >> 3;  // octets below jfull end inside the group
                if (!p.has_w) {
#pragma unroll 4
                    for (; j < jfull; ++j) {
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
                    for (; j < jfull; ++j) {
                        const Vec y2 = *reinterpret_cast<const Vec *>(ys + 8 * j * sizeof(T));
                        const Vec w2 = *reinterpret_cast<const Vec *>(ws + 8 * j * sizeof(T));
                        const T s0 = w2.x;
                        const T s1 = w2.y;
                        double f0[KB], f1[KB];
#pragma unroll
                        for (int bk = 0; bk < KB; ++bk) {
                            const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                            f0[bk] = static_cast<double>(static_cast<T>((has_x[bk] ? x2.x : static_cast<T>(xconst[bk])) * s0));
                            f1[bk] = static_cast<double>(static_cast<T>((has_x[bk] ? x2.y : static_cast<T>(xconst[bk])) * s1));
                        }
                        mma_octet(f0, f1, static_cast<double>(static_cast<T>(y2.x * s0)), static_cast<double>(static_cast<T>(y2.y * s1)));
                    }
                }
                for (; j < jhi; ++j) masked_octet(j);
                nfit = (lane == 0) ? (hi - o) : 0;
            }
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
                acc[i][0] += acc2[i][0];
                acc[i][1] += acc2[i][1];
            }
            gram_epilogue<KB>(p, acc, cy, nfit, g, nullptr, lane);  // p.fused == 0: the record goes straight to global memory
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == S) {
            stage = 0;
            phase ^= 1u;
        }
    }
}

template <typename T, int KB>
cudaError_t gram_multi_launch_t(const GramParams &p, const MultiPlan &mp, unsigned grid, size_t smem, cudaStream_t s) {
    auto kern = gram_multi_kernel<T, KB>;
    static size_t attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem > attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = smem;
    }
    kern<<<grid, MULTI_THREADS, smem, s>>>(p, mp);
    return cudaGetLastError();
}

cudaError_t gram_multi_launch_f64(int KB, const GramParams &p, const MultiPlan &mp, unsigned grid, size_t smem, cudaStream_t s);
cudaError_t gram_multi_launch_f32(int KB, const GramParams &p, const MultiPlan &mp, unsigned grid, size_t smem, cudaStream_t s);

}  // namespace b200
