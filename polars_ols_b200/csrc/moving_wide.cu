// moving_wide.cu — rls / rolling kernels for 9..64 coefficients (block per chunk, state in shared memory; moving_wide.cuh)
#include "moving_wide.cuh"
namespace b200 {
cudaError_t moving_launch_wide(cudaStream_t s, MovingParams &p, bool f64, const int64_t *gco, int64_t *launches) {
    return f64 ? launch_moving_wide_t<double>(s, p, gco, launches) : launch_moving_wide_t<float>(s, p, gco, launches);
}
}  // namespace b200
