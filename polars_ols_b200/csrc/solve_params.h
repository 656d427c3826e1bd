// solve_params.h — parameter block shared by the per-group k x k solve kernels (small_solve.cuh, batch_solve.cuh,
// cd_solve.cuh, cd_thread.cu)
#pragma once
#include <cstdint>

namespace b200 {

enum : int { ROUTE_CHOL = 0, ROUTE_LU = 1, ROUTE_CD = 2, ROUTE_CD_ACTIVE = 3, ROUTE_FLAGS_ONLY = 4 /* big.cuh: an SVD kernel solves */ };

struct SolveParams {
    int F;
    int64_t n_groups;
    const double *partial;         // [nseg][F*F + F + 1]
    const int64_t *group_seg_off;  // [n_groups+1] or nullptr (one segment per group)
    double *work;                  // [n_groups][F*F + 4F]
    double *beta;                  // [n_groups][F]
    int32_t *flags;                // [n_groups]
    int route;
    double alpha, l1_ratio, tol, illcond_ratio;
    int64_t max_iter;
    int positive;
};

}  // namespace b200
