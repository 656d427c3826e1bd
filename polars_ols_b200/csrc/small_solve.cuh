// small_solve.cuh — per-group k x k solve from the Gram partials written by gram_stream_kernel.
// Used when the solve cannot be fused into the streaming epilogue: groups split into several
// segments, elastic-net / lasso (coordinate descent on the Gram), or k > 16.
// One thread per group; the group's (G, c) lives in a private global scratch record (L1/L2 resident).
//
// Reference call sites restated: _get_least_squares_coefficients src/expressions.rs:351-388 (dispatch,
// empty -> zeros), solve_ridge src/least_squares.rs:342-371, solve_elastic_net :386-492.
#pragma once
#include <cstdint>

#include "solve_params.h"
#include "solvers.cuh"

namespace b200 {

__global__ void __launch_bounds__(128) small_solve_kernel(const SolveParams p) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= p.n_groups) return;
    const int F = p.F;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    const int64_t s0 = p.group_seg_off ? p.group_seg_off[g] : g;
    const int64_t s1 = p.group_seg_off ? p.group_seg_off[g + 1] : g + 1;
    double *G = p.work + static_cast<size_t>(g) * (static_cast<size_t>(F) * F + 4 * F);
    double *c = G + static_cast<size_t>(F) * F;
    double *scratch = c + F;  // 3F doubles
    // fixed-order (deterministic) reduction of the segment partials
    double nfit = 0.0;
    for (int e = 0; e < F * F + F; ++e) {
        double s = 0.0;
        for (int64_t sg = s0; sg < s1; ++sg) s += p.partial[static_cast<size_t>(sg) * P + e];
        G[e] = s;
    }
    for (int64_t sg = s0; sg < s1; ++sg) nfit += p.partial[static_cast<size_t>(sg) * P + F * F + F];
    double *beta = p.beta + g * F;
    if (nfit == 0.0) {  // src/expressions.rs:357-359
        for (int i = 0; i < F; ++i) beta[i] = 0.0;
        p.flags[g] = FLAG_EMPTY;
        return;
    }
    int fl = 0;
    if (p.route == ROUTE_CHOL || p.route == ROUTE_LU) {
        for (int i = 0; i < F; ++i) G[i * F + i] += p.alpha;
        fl = normal_equations_solve(G, F, F, c, p.route == ROUTE_LU, scratch, p.illcond_ratio);
        if (nfit <= static_cast<double>(F)) fl |= FLAG_WIDE;
        for (int i = 0; i < F; ++i) beta[i] = c[i];
    } else {
        // alpha is scaled by the number of fitted samples (src/least_squares.rs:419)
        cd_gram_solve(G, F, F, c, p.alpha * nfit, p.l1_ratio, p.max_iter, p.tol, p.positive != 0,
                      p.route == ROUTE_CD_ACTIVE, beta, scratch);
    }
    p.flags[g] = fl;
}

}  // namespace b200
