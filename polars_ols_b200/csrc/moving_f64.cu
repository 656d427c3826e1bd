// moving_f64.cu — f64 instantiations of the rls / rolling kernels (see moving.cuh)
#include "moving.cuh"
namespace b200 {
cudaError_t moving_launch_f64(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    return launch_moving_k<double>(s, p, gco, launches);
}
}  // namespace b200
