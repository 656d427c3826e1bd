// batch_solve.cuh — normal-equation solve of many small systems from the Gram records (k <= 16).
//
// The fused epilogue of gram_cta_kernel solves one group per solver warp, a ~3-10 us dependent chain; with
// short groups (a few hundred rows) the four solver warps of a CTA cannot keep up with the stream.  For those
// frames the streaming kernel only writes the [k*k + k + 1] record and this kernel solves ALL groups at full
// occupancy: a sub-warp of FP = 8 or 16 lanes per group (lane = row of the matrix, the same register Cholesky
// as the fused path, shuffles restricted to the sub-warp), 2-4 groups per warp, thousands of warps in flight, so
// the latency of one chain is hidden by the others.  Groups split into several segments are summed in a fixed
// order.  Same ladder and flags as normal_equations_solve (solvers.cuh): Cholesky -> LU with partial pivoting
// (lane 0 of the sub-warp, in a global scratch record), alpha on the diagonal, empty -> zeros.
// Reference: solve_ridge / solve_normal_equations src/least_squares.rs:277-371.
#pragma once
#include <cuda_runtime.h>

#include "gram_stream.cuh"
#include "small_solve.cuh"

namespace b200 {

template <int FP>
__global__ void __launch_bounds__(128) batch_solve_kernel(const SolveParams p) {
    constexpr int SUBS = 32 / FP;
    const int lane = threadIdx.x & 31, sub = lane / FP, sl = lane % FP;
    const int F = p.F;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    const int64_t g_raw = (static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SUBS + sub;
    const bool live = g_raw < p.n_groups;
    const int64_t g = live ? g_raw : p.n_groups - 1;  // idle sub-warps shadow the last group (shuffles need all lanes)
    const int64_t s0 = p.group_seg_off ? p.group_seg_off[g] : g;
    const int64_t s1 = p.group_seg_off ? p.group_seg_off[g + 1] : g + 1;
    double grow[FP];
    double ci = 0.0, nfit = 0.0;
#pragma unroll
    for (int c = 0; c < FP; ++c) grow[c] = (c == sl) ? 1.0 : 0.0;  // rows / columns >= F: identity padding
    for (int64_t sg = s0; sg < s1; ++sg) nfit += p.partial[static_cast<size_t>(sg) * P + static_cast<size_t>(F) * F + F];
    if (sl < F) {
#pragma unroll
        for (int c = 0; c < FP; ++c)
            if (c < F) {
                double s = 0.0;
                for (int64_t sg = s0; sg < s1; ++sg) s += p.partial[static_cast<size_t>(sg) * P + static_cast<size_t>(sl) * F + c];
                grow[c] = s;
            }
        for (int64_t sg = s0; sg < s1; ++sg) ci += p.partial[static_cast<size_t>(sg) * P + static_cast<size_t>(F) * F + sl];
#pragma unroll
        for (int c = 0; c < FP; ++c)
            if (c == sl) grow[c] += p.alpha;  // + alpha I, NOT scaled by n (src/least_squares.rs:352-356)
    }
    int fl = 0;
    bool need_lu = p.route == ROUTE_LU;
    if (!need_lu) {
        double mn, mx;
        const bool ok = warp_chol_solve<FP, FP>(grow, ci, F, sl, mn, mx);
        if (ok) {
            if (mx > p.illcond_ratio * mn) fl |= FLAG_ILLCOND;
        } else {
            fl |= FLAG_LU_FALLBACK;
            need_lu = true;
        }
    }
    if (nfit == 0.0) {  // src/expressions.rs:357-359
        if (live && sl < F) p.beta[g * F + sl] = 0.0;
        if (live && sl == 0) p.flags[g] = FLAG_EMPTY;
        return;
    }
    if (nfit <= static_cast<double>(F)) fl |= FLAG_WIDE;
    if (need_lu) {
        if (live && sl == 0) {  // rare: serial LU with partial pivoting on a private copy
            double *G = p.work + static_cast<size_t>(g) * (static_cast<size_t>(F) * F + 4 * F);
            double *c = G + static_cast<size_t>(F) * F;
            for (int e = 0; e < F * F + F; ++e) {
                double s = 0.0;
                for (int64_t sg = s0; sg < s1; ++sg) s += p.partial[static_cast<size_t>(sg) * P + e];
                G[e] = s;
            }
            for (int i = 0; i < F; ++i) G[i * F + i] += p.alpha;
            lu_solve_inplace(G, F, F, c);
            for (int i = 0; i < F; ++i) p.beta[g * F + i] = c[i];
            p.flags[g] = fl;
        }
        return;
    }
    if (live && sl < F) p.beta[g * F + sl] = ci;
    if (live && sl == 0) p.flags[g] = fl;
}

inline cudaError_t launch_batch_solve(cudaStream_t stream, const SolveParams &sp) {
    const int FP = sp.F <= 8 ? 8 : 16;
    const int64_t per_block = 4 * (32 / FP);
    const unsigned blocks = static_cast<unsigned>((sp.n_groups + per_block - 1) / per_block);
    if (FP == 8) batch_solve_kernel<8><<<blocks, 128, 0, stream>>>(sp);
    else batch_solve_kernel<16><<<blocks, 128, 0, stream>>>(sp);
    return cudaGetLastError();
}

}  // namespace b200
