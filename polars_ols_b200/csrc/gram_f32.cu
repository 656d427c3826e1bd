// gram_f32.cu — f32-input instantiations of the row-streaming Gram kernel (see gram_stream.cuh)
#include "gram_cta.cuh"
#include "gram_multi.cuh"
#include "gram_wide.cuh"
#include "gram_ldg.cuh"
#include "gram_stream.cuh"
namespace b200 {
cudaError_t gram_ldg_launch_f32(int KB, int U, const GramParams &p, unsigned grid, int warps, cudaStream_t s) {
    return gram_ldg_launch_any<float>(KB, U, p, grid, warps, s);
}
cudaError_t gram_cta_launch_f32(int KB, const GramParams &p, unsigned grid, size_t smem, cudaStream_t s) {
    const bool teams = p.team > 0 && p.team < CTA_CONSUMERS;
    if (KB == 1) return teams ? gram_cta_launch_t<float, 1, true>(p, grid, smem, s) : gram_cta_launch_t<float, 1, false>(p, grid, smem, s);
    return teams ? gram_cta_launch_t<float, 2, true>(p, grid, smem, s) : gram_cta_launch_t<float, 2, false>(p, grid, smem, s);
}
cudaError_t gram_wide_launch_f32(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s) {
    return gram_wide_launch_t<float>(p, grid, smem, s);
}
cudaError_t gram_multi_launch_f32(int KB, const GramParams &p, const MultiPlan &mp, unsigned grid, size_t smem, cudaStream_t s) {
    return KB == 1 ? gram_multi_launch_t<float, 1>(p, mp, grid, smem, s) : gram_multi_launch_t<float, 2>(p, mp, grid, smem, s);
}
}  // namespace b200
