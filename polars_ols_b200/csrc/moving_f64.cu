// moving_f64.cu — f64 rls / rolling kernels: dispatch on the number of coefficients to the translation units that
// instantiate them (moving_f64_r0..r3.cu, see moving.cuh)
#include "moving.cuh"
namespace b200 {
cudaError_t moving_launch_f64(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    if (p.F <= 8) return moving_launch_f64_r0(s, p, gco, launches);
    if (p.F <= 11) return moving_launch_f64_r1(s, p, gco, launches);
    if (p.F <= 14) return moving_launch_f64_r2(s, p, gco, launches);
    return moving_launch_f64_r3(s, p, gco, launches);
}
}  // namespace b200
