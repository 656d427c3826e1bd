// moving_f32_r3.cu — f32 instantiations of the rls / rolling kernels for 15..16 coefficients (see moving.cuh)
#include "moving.cuh"
namespace b200 {
cudaError_t moving_launch_f32_r3(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    return launch_moving_range<float, 15, 16>(s, p, gco, launches);
}
}  // namespace b200
