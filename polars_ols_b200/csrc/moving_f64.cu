// moving_f64.cu / moving_f32.cu merged: dispatch of the rls / rolling launches on dtype and coefficient count
#include "moving.cuh"
namespace b200 {
int moving_fast_blocks_f64(int F, int kind, int nc);
int moving_fast_blocks_f32(int F, int kind, int nc);
cudaError_t moving_launch_f64(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    if (p.F <= 8) return moving_launch_f64_r0(s, p, gco, launches);
    return moving_launch_wide(s, p, true, gco, launches);
}
cudaError_t moving_launch_f32(cudaStream_t s, MovingParams &p, const int64_t *gco, int64_t *launches) {
    if (p.F <= 8) return moving_launch_f32_r0(s, p, gco, launches);
    return moving_launch_wide(s, p, false, gco, launches);
}
int moving_fast_blocks_per_sm(bool f64, int F, int kind, int n_cols) {
    return f64 ? moving_fast_blocks_f64(F, kind, n_cols) : moving_fast_blocks_f32(F, kind, n_cols);
}
}  // namespace b200
