// moving_f64_r2.cu — f64 instantiations of the rls / rolling kernels for 12..14 coefficients (see moving.cuh)
#include "moving.cuh"
namespace b200 {
cudaError_t moving_launch_f64_r2(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    return launch_moving_range<double, 12, 14>(s, p, gco, launches);
}
}  // namespace b200
