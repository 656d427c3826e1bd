// pred_f32.cu — f32 instantiations of the fused Gram -> solve -> predict kernel (gram_pred.cuh)
#include "gram_pred.cuh"
namespace b200 {
cudaError_t gram_pred_launch_f32(int KB, const GramParams &p, const PredOut &po, unsigned grid, size_t smem, cudaStream_t s) {
    return KB == 1 ? gram_pred_launch_t<float, 1>(p, po, grid, smem, s) : gram_pred_launch_t<float, 2>(p, po, grid, smem, s);
}
}  // namespace b200
