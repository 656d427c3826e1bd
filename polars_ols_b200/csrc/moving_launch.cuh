// moving_launch.cuh — host-side launch templates of the rls / rolling kernels (instantiated per dtype in
// moving_f64_r0.cu / moving_f32_r0.cu for 1..8 coefficients).
#pragma once
#include "moving_fast.cuh"

namespace b200 {

// null-free frame, k <= 8: rows reach the threads through private cp.async staging rings (moving_fast.cuh);
// WT = sample weights present (the unweighted instantiation carries no scaling arithmetic)
template <typename T, int K, bool WT>
static cudaError_t launch_moving_fast_t(cudaStream_t stream, MovingParams &p, int64_t *launches) {
    const unsigned cb = static_cast<unsigned>((p.n_chunks + 127) / 128);
    (void)cb;
        const int nc = p.kd + 1 + (p.w ? 1 : 0);
        // validity bytes: one coalesced pass after the kernels instead of a scattered byte store per row and thread
        // (residuals against a target with nulls are the exception: the kernels write those bytes themselves)
        struct ValidityPass {
            cudaStream_t s; const MovingParams &p; int64_t *launches;
            ~ValidityPass() {
                if (!p.out_valid || p.state_only || (p.mode == 1 && p.target_validity)) return;
                const int64_t n = p.mode == 2 ? p.n_rows * p.F : p.n_rows;
                if (n == 0) return;
                const int nan_is_null = (p.mode == 2 || p.kind == MOVING_ROLLING) ? 1 : 0;
                moving_validity_kernel<<<static_cast<unsigned>((n + 2047) / 2048), 256, 0, s>>>(p.out, p.out_valid, n, nan_is_null);
                ++*launches;
            }
        } validity_pass{stream, p, launches};
        const unsigned fb = static_cast<unsigned>((p.n_chunks + MF_THREADS - 1) / MF_THREADS);
        if (p.kind == MOVING_ROLLING && p.nbr) {
            const size_t smem = moving_nbr_smem(nc);
            cudaFuncSetAttribute(rolling_nbr_kernel<T, K, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            chunk_totals_kernel<T, K, false><<<static_cast<unsigned>((p.n_chunks * 32 + 255) / 256), 256, 0, stream>>>(p, p.summaries);
            const unsigned nb = static_cast<unsigned>((p.n_chunks + MF_THREADS - 2) / (MF_THREADS - 1));
            rolling_nbr_kernel<T, K, WT><<<nb, MF_THREADS, smem, stream>>>(p, p.summaries);
            *launches += 2;
            return cudaGetLastError();
        }
        if (p.kind == MOVING_ROLLING) {
            const size_t smem = moving_fast_smem(nc, true);
            cudaFuncSetAttribute(rolling_fast_kernel<T, K, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            rolling_fast_kernel<T, K, WT><<<fb, MF_THREADS, smem, stream>>>(p);
            ++*launches;
            return cudaGetLastError();
        }
        const size_t smem = moving_fast_smem(nc, false);
        cudaFuncSetAttribute(rls_fast_main_kernel<T, K, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        chunk_totals_kernel<T, K, true><<<static_cast<unsigned>((p.n_chunks * 32 + 255) / 256), 256, 0, stream>>>(p, p.summaries);  // pass 1: chunk summaries
        const unsigned sb = static_cast<unsigned>((p.n_super * 32 + 127) / 128);
        rls_scan_kernel<K><<<sb, 128, 0, stream>>>(p, 0, p.n_super, p.sup_c0, p.sup_c1, p.group_sup_off, p.sup);
        rls_scan_kernel<K><<<static_cast<unsigned>((p.n_groups * 32 + 127) / 128), 128, 0, stream>>>(p, 1, p.n_super, p.sup_c0, p.sup_c1, p.group_sup_off, p.sup);
        *launches += 3;
        if (p.state_only) return cudaGetLastError();
        rls_scan_kernel<K><<<sb, 128, 0, stream>>>(p, 2, p.n_super, p.sup_c0, p.sup_c1, p.group_sup_off, p.sup);
        rls_fast_main_kernel<T, K, WT><<<fb, MF_THREADS, smem, stream>>>(p);
        *launches += 2;
        return cudaGetLastError();
}

template <typename T, int K>
static cudaError_t launch_moving_t(cudaStream_t stream, MovingParams &p, const int64_t *group_chunk_off_dev,
                                   int64_t *launches) {
    const unsigned cb = static_cast<unsigned>((p.n_chunks + 127) / 128);
    if (p.n_chunks == 0) return cudaSuccess;
    if (p.fast) return p.w ? launch_moving_fast_t<T, K, true>(stream, p, launches) : launch_moving_fast_t<T, K, false>(stream, p, launches);
    {   // chunk-interleaved copies of every input column (one coalesced read + write of the data)
        TransposeParams tp;
        int nc = 0;
        for (int j = 0; j <= p.kd; ++j) { tp.src[nc] = p.cols[j]; tp.dst[nc] = const_cast<void *>(p.tcols[j]); ++nc; }
        if (p.w) { tp.src[nc] = p.w; tp.dst[nc] = const_cast<void *>(p.tw); ++nc; }
        if (p.mask) { tp.src[nc] = p.mask; tp.dst[nc] = const_cast<void *>(p.tmask); ++nc; }
        tp.chunk_r0 = p.chunk_r0;
        tp.chunk_r1 = p.chunk_r1;
        tp.n_chunks = p.n_chunks;
        tp.chunk_len = p.chunk_len;
        const dim3 grid(static_cast<unsigned>((p.n_chunks + 31) / 32), static_cast<unsigned>((p.chunk_len + 31) / 32), static_cast<unsigned>(nc));
        chunk_transpose_kernel<T><<<grid, 256, 0, stream>>>(tp);
        ++*launches;
    }
    if (p.kind == MOVING_ROLLING) {
        if (p.mask) {
            rolling_prepass_kernel<T, K><<<static_cast<unsigned>((p.n_groups + 127) / 128), 128, 0, stream>>>(p);
            ++*launches;
        } else {
            p.series_info = nullptr;
        }
        rolling_main_kernel<T, K><<<cb, 128, 0, stream>>>(p);
        ++*launches;
    } else {
        rls_summary_kernel<T, K><<<cb, 128, 0, stream>>>(p);
        const unsigned sb = static_cast<unsigned>((p.n_super * 32 + 127) / 128);
        rls_scan_kernel<K><<<sb, 128, 0, stream>>>(p, 0, p.n_super, p.sup_c0, p.sup_c1, p.group_sup_off, p.sup);
        rls_scan_kernel<K><<<static_cast<unsigned>((p.n_groups * 32 + 127) / 128), 128, 0, stream>>>(p, 1, p.n_super, p.sup_c0, p.sup_c1, p.group_sup_off, p.sup);
        *launches += 3;
        if (p.state_only) return cudaGetLastError();
        rls_scan_kernel<K><<<sb, 128, 0, stream>>>(p, 2, p.n_super, p.sup_c0, p.sup_c1, p.group_sup_off, p.sup);
        rls_main_kernel<T, K><<<cb, 128, 0, stream>>>(p);
        *launches += 2;
    }
    return cudaGetLastError();
}

// Coefficient counts KLO..KHI of one dtype (one translation unit each: the K >= 9 instantiations spill and take minutes
// to compile, so they are spread over several .cu files that nvcc builds in parallel — moving_f64_r*.cu / moving_f32_r*.cu)
template <typename T, int KLO, int KHI>
static cudaError_t launch_moving_range(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    if constexpr (KLO >= KHI) {
        return launch_moving_t<T, KHI>(s, p, gco, launches);
    } else {
        if (p.F <= KLO) return launch_moving_t<T, KLO>(s, p, gco, launches);
        return launch_moving_range<T, KLO + 1, KHI>(s, p, gco, launches);
    }
}


template <typename T, int K>
static int moving_fast_blocks_t(int kind, int nc) {
    int nb = 0;
    if (kind == MOVING_ROLLING) {
        const size_t smem = moving_fast_smem(nc, true);
        cudaFuncSetAttribute(rolling_fast_kernel<T, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rolling_fast_kernel<T, K, false>, MF_THREADS, smem);
    } else {
        const size_t smem = moving_fast_smem(nc, false);
        cudaFuncSetAttribute(rls_fast_main_kernel<T, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rls_fast_main_kernel<T, K, false>, MF_THREADS, smem);
    }
    return nb;
}

template <typename T, int KLO, int KHI>
static int moving_fast_blocks_range(int F, int kind, int nc) {
    if constexpr (KLO >= KHI) {
        return moving_fast_blocks_t<T, KHI>(kind, nc);
    } else {
        if (F <= KLO) return moving_fast_blocks_t<T, KLO>(kind, nc);
        return moving_fast_blocks_range<T, KLO + 1, KHI>(F, kind, nc);
    }
}

}  // namespace b200
