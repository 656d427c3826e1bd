// gram_f64.cu — f64 instantiations of the row-streaming Gram kernel (see gram_stream.cuh)
#include "gram_cta.cuh"
#include "gram_multi.cuh"
#include "gram_wide.cuh"
#include "gram_ldg.cuh"
#include "gram_stream.cuh"
namespace b200 {
cudaError_t gram_ldg_launch_f64(int KB, int U, const GramParams &p, unsigned grid, int warps, cudaStream_t s) {
    return gram_ldg_launch_any<double>(KB, U, p, grid, warps, s);
}
cudaError_t gram_cta_launch_f64(int KB, const GramParams &p, unsigned grid, size_t smem, cudaStream_t s) {
    const bool teams = p.team > 0 && p.team < CTA_CONSUMERS;
    if (KB == 1) return teams ? gram_cta_launch_t<double, 1, true>(p, grid, smem, s) : gram_cta_launch_t<double, 1, false>(p, grid, smem, s);
    return teams ? gram_cta_launch_t<double, 2, true>(p, grid, smem, s) : gram_cta_launch_t<double, 2, false>(p, grid, smem, s);
}
cudaError_t gram_wide_launch_f64(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s) {
    return gram_wide_launch_t<double>(p, grid, smem, s);
}
cudaError_t gram_multi_launch_f64(int KB, const GramParams &p, const MultiPlan &mp, unsigned grid, size_t smem, cudaStream_t s) {
    return KB == 1 ? gram_multi_launch_t<double, 1>(p, mp, grid, smem, s) : gram_multi_launch_t<double, 2>(p, mp, grid, smem, s);
}
}  // namespace b200
