// gram_f64.cu — f64 instantiations of the row-streaming Gram kernel (see gram_stream.cuh)
#include "gram_stream.cuh"
namespace b200 {
cudaError_t gram_launch_f64(int KB, const GramParams &p, unsigned grid, int warps, size_t smem, cudaStream_t s) {
    return gram_launch_any<double>(KB, p, grid, warps, smem, s);
}
}  // namespace b200
