// gram_simt.cuh — row-streaming Gram kernel for k <= 8 on the FP64 FMA pipe (no tensor-core MMA).
//
// Why it exists: on this B200 the f64 tensor instruction (mma.sync.m8n8k4.f64 -> DMMA.8x8x4) has a long
// dependent-issue latency and a low issue rate (tools/microbench.cu, profiles/), so for small k the DMMA
// kernels are bound by the DMMA pipe at ~0.75x of the HBM roofline.  Here every lane owns whole rows:
// a warp iteration reads 32 consecutive rows of each column (one fully coalesced 256-byte request per
// column) and each lane updates its private 36 + 8 accumulators (lower triangle of X^T X and X^T y) with
// independent DFMAs — no dependent chains, 9 * 256 B in flight per warp and iteration.
// At the end of a segment the 44 per-lane partial sums are combined across the warp by recursive halving
// (44 -> 22 -> 11 -> 6 -> 3 -> 2 values per lane, 44 shuffle steps instead of 220 for a plain butterfly),
// written to the per-warp scratch as the symmetric matrix, and the common epilogue (gram_stream.cuh:
// warp-cooperative Cholesky -> LU fallback, or raw partial) finishes the group.
#pragma once
#include "gram_stream.cuh"

namespace b200 {

constexpr int SIMT_NV = 44;  // 36 lower-triangle entries + 8 entries of X^T y

// packed lower-triangle index p = j (j + 1) / 2 + l  (l <= j < 8)  ->  (j, l)
__device__ __forceinline__ void simt_unpack(int p, int &j, int &l) {
    j = 0;
#pragma unroll
    for (int t = 1; t < 8; ++t) j += (p >= t * (t + 1) / 2) ? 1 : 0;
    l = p - j * (j + 1) / 2;
}

// One recursive-halving round over lane bit BIT: v[0..N) -> v[0..H), H = ceil(N / 2).
// Lanes with the bit clear keep entries [0, H), lanes with it set keep [H, N) (zero padded).
template <int N, int BIT>
__device__ __forceinline__ void simt_halve(double (&v)[SIMT_NV], int lane) {
    constexpr int H = (N + 1) / 2;
    const bool up = (lane & BIT) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const double hi = (H + i < N) ? v[H + i] : 0.0;
        const double send = up ? v[i] : hi;
        const double keep = up ? hi : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
    }
}

template <typename T, int U, int MAXW, bool EXTRA>
__global__ void __launch_bounds__(MAXW * 32) gram_simt_kernel(const GramParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int FP = 8, LD = FP + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int kd = p.kd, F = p.F;
    double *Gs = reinterpret_cast<double *>(smem + static_cast<size_t>(warp) * gram_scratch_bytes<T>(F, 1));
    const T *ycol = static_cast<const T *>(p.cols[kd]);
    const T *wcol = (EXTRA && p.has_w) ? static_cast<const T *>(p.cols[kd + 1]) : nullptr;
    const T *mcol = (EXTRA && p.has_mask) ? static_cast<const T *>(p.cols[kd + 1 + (p.has_w ? 1 : 0)]) : nullptr;
    const T *xcol[FP];
#pragma unroll
    for (int j = 0; j < FP; ++j) xcol[j] = static_cast<const T *>(p.cols[j < kd ? j : 0]);  // padding slots alias column 0
    const double xone = p.intercept ? 1.0 : 0.0;  // value of feature slot kd (the synthetic `const` column)
    const int64_t wg = static_cast<int64_t>(blockIdx.x) * W + warp;
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * W;

    for (int64_t seg = wg; seg < p.nseg; seg += nwarps) {
        const int64_t r0 = p.seg_off[seg], r1 = p.seg_off[seg + 1];
        double v[SIMT_NV];
#pragma unroll
        for (int i = 0; i < SIMT_NV; ++i) v[i] = 0.0;
        int nfit = 0;

        auto accumulate = [&](const double (&x)[FP], double y) {
#pragma unroll
            for (int j = 0; j < FP; ++j) {
#pragma unroll
                for (int l = 0; l <= j; ++l) v[j * (j + 1) / 2 + l] = fma(x[j], x[l], v[j * (j + 1) / 2 + l]);
                v[36 + j] = fma(x[j], y, v[36 + j]);
            }
        };

        int64_t r = r0 + lane;
        if (!EXTRA) {
            // mask-free interior: U x 32 rows per iteration, all loads issued before the first use
            for (; (r - lane) + 32 * U <= r1; r += 32 * U) {
                T xv[U][FP], yv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    yv[u] = ycol[r + 32 * u];
#pragma unroll
                    for (int j = 0; j < FP; ++j) xv[u][j] = xcol[j][r + 32 * u];
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double x[FP];
#pragma unroll
                    for (int j = 0; j < FP; ++j) x[j] = (j < kd) ? static_cast<double>(xv[u][j]) : ((j == kd) ? xone : 0.0);
                    accumulate(x, static_cast<double>(yv[u]));
                }
            }
        }
        // predicated tail (and the whole segment when weights / a row mask are present)
        for (; r - lane < r1; r += 32) {
            const bool in = r < r1;
            const int64_t rr = in ? r : r0;  // r1 > r0 here, so r0 is a valid row to read
            bool valid = in;
            T s = T(1);
            if (EXTRA) {
                if (mcol) valid = valid && (mcol[rr] != T(0));
                if (wcol) {
                    const T w = wcol[rr];
                    s = p.w_is_sqrt ? w : static_cast<T>(sqrt(w));
                }
            }
            double x[FP];
#pragma unroll
            for (int j = 0; j < FP; ++j) {
                const T xv = (j < kd) ? xcol[j][rr] : static_cast<T>((j == kd) ? xone : 0.0);
                x[j] = valid ? static_cast<double>(static_cast<T>(xv * s)) : 0.0;
            }
            const double y = valid ? static_cast<double>(static_cast<T>(ycol[rr] * s)) : 0.0;
            nfit += valid ? 1 : 0;
            accumulate(x, y);
        }
        if (!EXTRA) {
            nfit = static_cast<int>(r1 - r0);
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nfit += __shfl_xor_sync(0xffffffffu, nfit, o);
        }

        // ---- cross-lane reduction by recursive halving: 44 -> 22 -> 11 -> 6 -> 3 -> 2 values per lane ----
        simt_halve<44, 16>(v, lane);
        simt_halve<22, 8>(v, lane);
        simt_halve<11, 4>(v, lane);
        simt_halve<6, 2>(v, lane);
        simt_halve<3, 1>(v, lane);
        // original index of final slot s held by this lane (inverse of the keep rule of every round)
        double *cs = Gs + FP * LD;
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            int idx = s;
            bool live = true;
            idx = (lane & 1) ? 2 + idx : idx;   live = live && idx < 3;
            idx = (lane & 2) ? 3 + idx : idx;   live = live && idx < 6;
            idx = (lane & 4) ? 6 + idx : idx;   live = live && idx < 11;
            idx = (lane & 8) ? 11 + idx : idx;  live = live && idx < 22;
            idx = (lane & 16) ? 22 + idx : idx; live = live && idx < 44;
            if (live) {
                if (idx < 36) {
                    int j, l;
                    simt_unpack(idx, j, l);
                    Gs[j * LD + l] = v[s];
                    Gs[l * LD + j] = v[s];
                } else {
                    cs[idx - 36] = v[s];
                }
            }
        }
        __syncwarp();
        gram_finish<1>(p, nfit, seg, Gs, lane);
    }
}

template <typename T, int U, bool EXTRA>
cudaError_t gram_simt_launch_e(const GramParams &p, unsigned grid, int warps, cudaStream_t s) {
    auto kern = gram_simt_kernel<T, U, (U == 1 ? 16 : 8), EXTRA>;  // U > 1 needs > 128 registers per thread
    const size_t smem = static_cast<size_t>(warps) * gram_scratch_bytes<T>(p.F, 1);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, warps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

template <typename T>
cudaError_t gram_simt_launch_any(int U, const GramParams &p, unsigned grid, int warps, cudaStream_t s) {
    if (p.has_w || p.has_mask) return gram_simt_launch_e<T, 1, true>(p, grid, warps, s);
    switch (U) {
        case 1: return gram_simt_launch_e<T, 1, false>(p, grid, warps, s);
        case 2: return gram_simt_launch_e<T, 2, false>(p, grid, warps, s);
        default: return gram_simt_launch_e<T, 4, false>(p, grid, warps, s);
    }
}

cudaError_t gram_simt_launch_f64(int U, const GramParams &p, unsigned grid, int warps, cudaStream_t s);
cudaError_t gram_simt_launch_f32(int U, const GramParams &p, unsigned grid, int warps, cudaStream_t s);

}  // namespace b200
