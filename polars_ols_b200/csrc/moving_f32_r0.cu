// moving_f32_r0.cu — f32 instantiations of the rls / rolling kernels for 1..8 coefficients (see moving.cuh, moving_fast.cuh)
#include "moving_launch.cuh"
namespace b200 {
cudaError_t moving_launch_f32_r0(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    return launch_moving_range<float, 1, 8>(s, p, gco, launches);
}
int moving_fast_blocks_f32(int F, int kind, int nc) { return moving_fast_blocks_range<float, 1, 8>(F, kind, nc); }
}  // namespace b200
