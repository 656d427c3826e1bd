// moving_f32.cu — f32 instantiations of the rls / rolling kernels (see moving.cuh)
#include "moving.cuh"
namespace b200 {
cudaError_t moving_launch_f32(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    return launch_moving_k<float>(s, p, gco, launches);
}
}  // namespace b200
