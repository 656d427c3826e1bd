// cd_thread.h — launcher of the thread-per-group coordinate descent (cd_thread.cu)
#pragma once
#include <cuda_runtime.h>

#include "solve_params.h"

namespace b200 {
// blocks_per_sm > 0 caps the resident warps per SM (tuning hook; 0 = what shared memory and registers allow)
cudaError_t launch_cd_thread(cudaStream_t stream, const SolveParams &sp, int sm_count, int blocks_per_sm);
}  // namespace b200
