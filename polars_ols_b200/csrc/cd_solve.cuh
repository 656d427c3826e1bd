// cd_solve.cuh — elastic-net / lasso coordinate descent on the per-group Gram matrices, warp-cooperative.
//
// Restates solve_elastic_net (src/least_squares.rs:386-492 of /root/reference) in its Gram form — the
// reference's residual-form update  w_j <- S(x_j^T (r + x_j w_j), a l1) / (|x_j|^2 + a (1 - l1))  with
// q = X^T y - G w maintained incrementally: x_j^T (r + x_j w_j) = q_j + G_jj w_j.  Same cyclic order,
// alpha * n scaling (:419), active-set pruning (:472-476) and stop rule ||w - w_old||_2 < tol (:436-444)
// as cd_gram_solve() in solvers.cuh (the host-checked serial version); here a sub-warp of WIDTH lanes
// owns one group: lane l holds coordinates l, l + WIDTH, .. (q, w, G_ll in registers), G sits in shared
// memory and each coordinate step is: owner computes the new w_j, one shuffle broadcasts the step
// delta, every lane updates its q entries from row j of G (conflict-free).  32 / WIDTH groups per warp
// run concurrently, so the latency-bound sweeps of 100k small groups (C3) overlap.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "small_solve.cuh"
#include "solvers.cuh"

namespace b200 {

template <int WIDTH, int NPL>
__global__ void __launch_bounds__(128) cd_solve_kernel(const SolveParams p) {
    extern __shared__ __align__(16) unsigned char cd_smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int F = p.F;
    const int lane = threadIdx.x & 31;
    const int sub = lane / WIDTH, sl = lane % WIDTH;          // sub-warp inside the warp, lane inside the sub-warp
    constexpr int SUBS_PER_WARP = 32 / WIDTH;
    const int warp_in_block = threadIdx.x >> 5;
    const int subs_per_block = (blockDim.x >> 5) * SUBS_PER_WARP;
    double *Gs = reinterpret_cast<double *>(cd_smem) + static_cast<size_t>(warp_in_block * SUBS_PER_WARP + sub) * F * F;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    const bool active_set = p.route == ROUTE_CD_ACTIVE;
    const bool positive = p.positive != 0;
    const int64_t n_iters = (p.n_groups + static_cast<int64_t>(gridDim.x) * subs_per_block - 1) / (static_cast<int64_t>(gridDim.x) * subs_per_block);

    for (int64_t it = 0; it < n_iters; ++it) {
        // every sub-warp of a warp runs the same number of iterations (shuffles need the full warp)
        const int64_t g = (it * gridDim.x + blockIdx.x) * subs_per_block + warp_in_block * SUBS_PER_WARP + sub;
        const bool live = g < p.n_groups;
        const int64_t s0 = live ? (p.group_seg_off ? p.group_seg_off[g] : g) : 0;
        const int64_t s1 = live ? (p.group_seg_off ? p.group_seg_off[g + 1] : g + 1) : 0;
        // ---- gather the group's Gram matrix (fixed-order sum over its segments) ----
        for (int e = sl; e < F * F; e += WIDTH) {
            double s = 0.0;
            for (int64_t sg = s0; sg < s1; ++sg) s += p.partial[static_cast<size_t>(sg) * P + e];
            Gs[e] = s;
        }
        double q[NPL], w[NPL], w_old[NPL], gd[NPL], rinv[NPL];
        double nfit = 0.0;
        for (int64_t sg = s0; sg < s1; ++sg) nfit += p.partial[static_cast<size_t>(sg) * P + F * F + F];
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int l = sl + WIDTH * t;
            double s = 0.0;
            if (l < F)
                for (int64_t sg = s0; sg < s1; ++sg) s += p.partial[static_cast<size_t>(sg) * P + F * F + l];
            q[t] = s;
            w[t] = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int l = sl + WIDTH * t;
            gd[t] = (l < F) ? Gs[l * F + l] : 1.0;
        }
        // alpha *= n_samples (src/least_squares.rs:419); products rounded on their own (no contraction into the
        // subtraction / addition that follows), as the reference computes alpha * l1_ratio and alpha * (1 - l1_ratio)
        const double a = __dmul_rn(p.alpha, nfit);
        const double l1 = __dmul_rn(a, p.l1_ratio), l2 = __dmul_rn(a, 1.0 - p.l1_ratio);
        // 1 / (|x_j|^2 + a (1 - l1)) once per coordinate: the per-update division of the reference (:431) becomes
        // a multiplication (f64 division is a ~20-instruction sequence and this loop is instruction bound)
#pragma unroll
        for (int t = 0; t < NPL; ++t) rinv[t] = 1.0 / (gd[t] + l2);
        // sub-warps diverge freely (different sweep counts / active sets): shuffles use the sub-warp's own mask
        const unsigned submask = (WIDTH == 32) ? FULL : (((1u << WIDTH) - 1u) << (sub * WIDTH));
        if (NPL == 1) {
            // One coordinate per lane (F <= WIDTH).  The sub-warps of a warp sweep in LOCK-STEP until all of them have
            // stopped (a finished one freezes its state): the warp is busy until its slowest group is done anyway, and
            // full-mask shuffles with a compile-time source lane cost 2 instructions where sub-warp masks cost ~12
            // (MATCH / REDUX / BRA.DIV convergence checks — ncu: the kernel is issue-bound).  The coordinate loop is
            // fully unrolled and only the step delta is broadcast.  Same arithmetic as the general loop below: w_new
            // and delta = w_new - w_old come from the owner lane, q is updated from row j of G.  The stop test
            // sqrt(d2) < tol is evaluated as d2 <= cut, cut = the largest double whose correctly rounded square root is
            // below tol (identical decisions, no sqrt per sweep).
            const bool any_below = p.tol > 0.0;  // tol <= 0 (or NaN): nothing is ever below it
            double cut = p.tol * p.tol;
            if (any_below) {  // tol * tol is within a few ulps of the boundary: both loops run 0-2 times
                while (cut > 0.0 && sqrt(cut) >= p.tol) cut = __longlong_as_double(__double_as_longlong(cut) - 1);
                while (cut < 1.0e300 && sqrt(__longlong_as_double(__double_as_longlong(cut) + 1)) < p.tol)
                    cut = __longlong_as_double(__double_as_longlong(cut) + 1);
            }
            double q0 = q[0], w0 = 0.0;
            const double gd0 = gd[0], ri0 = rinv[0];
            const bool mine = sl < F;
            const double *gcol = Gs + (mine ? sl : 0);
            uint32_t active = (F >= 32) ? 0xffffffffu : ((1u << F) - 1u);  // NPL == 1: F <= WIDTH <= 32
            bool done = !(live && nfit != 0.0);
            const int sub_shift = sub * WIDTH;
            for (int64_t sweep = 0; sweep < p.max_iter; ++sweep) {
                if (__all_sync(FULL, done)) break;
                const double wold = w0;
                const uint32_t loop_set = done ? 0u : active;  // the reference iterates over a clone of the active list (:459)
#pragma unroll
                for (int j = 0; j < WIDTH; ++j) {
                    const bool on = (loop_set >> j) & 1u;  // uniform inside the sub-warp (bits >= F are never set)
                    // soft_threshold (src/least_squares.rs:373-379), branch- and fmax-free: keep = |rho| - l1 > 0 (a NaN
                    // is not kept, like f64::max), clamp at zero when `positive`, sign of rho through the high word
                    const double rho = fma(gd0, w0, q0);
                    const double av = fabs(rho) - l1;
                    bool keep = av > 0.0;
                    if (positive) keep = keep && (rho > 0.0);
                    const int hi = keep ? (__double2hiint(av) | (__double2hiint(rho) & static_cast<int>(0x80000000u))) : 0;
                    const int lo = keep ? __double2loint(av) : 0;
                    const double wn = __hiloint2double(hi, lo) * ri0;
                    const double delta = __shfl_sync(FULL, wn - w0, j, WIDTH);
                    if (active_set) {
                        const unsigned small = __ballot_sync(FULL, on && sl == j && fabs(wn) < p.tol);
                        if ((small >> sub_shift) & ((WIDTH == 32) ? 0xffffffffu : ((1u << WIDTH) - 1u))) active &= ~(1u << j);  // never re-admitted (:472-476)
                    }
                    w0 = (on && sl == j) ? wn : w0;
                    if (on && mine && delta != 0.0) q0 = fma(-gcol[j * F], delta, q0);  // G symmetric: row j == column j
                }
                const double d = w0 - wold;
                double d2 = mine ? d * d : 0.0;
#pragma unroll
                for (int o = WIDTH / 2; o > 0; o >>= 1) d2 += __shfl_xor_sync(FULL, d2, o, WIDTH);
                if (any_below && d2 <= cut) done = true;
            }
            if (live) {
                if (mine) p.beta[g * F + sl] = (nfit == 0.0) ? 0.0 : w0;  // src/expressions.rs:357-359: no rows -> zeros
                if (sl == 0) p.flags[g] = (nfit == 0.0) ? FLAG_EMPTY : 0;
            }
        } else
        if (live) {
            if (nfit == 0.0) {  // src/expressions.rs:357-359: no rows -> zeros
#pragma unroll
                for (int t = 0; t < NPL; ++t) {
                    const int l = sl + WIDTH * t;
                    if (l < F) p.beta[g * F + l] = 0.0;
                }
                if (sl == 0) p.flags[g] = FLAG_EMPTY;
            } else {
                uint64_t active = (F >= 64) ? ~0ull : ((1ull << F) - 1ull);
                for (int64_t sweep = 0; sweep < p.max_iter; ++sweep) {
#pragma unroll
                    for (int t = 0; t < NPL; ++t) w_old[t] = w[t];
                    const uint64_t loop_set = active;  // the reference iterates over a clone of the active list (:459)
                    for (int j = 0; j < F; ++j) {
                        if (!((loop_set >> j) & 1ull)) continue;
                        const int owner = j % WIDTH, tj = j / WIDTH;
                        double qj = q[0], wj = w[0], gj = gd[0], rj = rinv[0];
#pragma unroll
                        for (int t = 1; t < NPL; ++t)
                            if (t == tj) { qj = q[t]; wj = w[t]; gj = gd[t]; rj = rinv[t]; }
                        const double rho = fma(gj, wj, qj);
                        const double wn_own = soft_threshold(rho, l1, positive) * rj;
                        const double wn = __shfl_sync(submask, wn_own, owner, WIDTH);
                        const double wo = __shfl_sync(submask, wj, owner, WIDTH);
                        const double delta = wn - wo;
                        if (sl == owner) {
#pragma unroll
                            for (int t = 0; t < NPL; ++t)
                                if (t == tj) w[t] = wn;
                        }
                        if (delta != 0.0) {
#pragma unroll
                            for (int t = 0; t < NPL; ++t) {
                                const int l = sl + WIDTH * t;
                                if (l < F) q[t] = fma(-Gs[j * F + l], delta, q[t]);  // G symmetric: row j == column j
                            }
                        }
                        if (active_set && fabs(wn) < p.tol) active &= ~(1ull << j);  // never re-admitted (:472-476)
                    }
                    double d2 = 0.0;
#pragma unroll
                    for (int t = 0; t < NPL; ++t) {
                        const double d = w[t] - w_old[t];
                        d2 = fma(d, d, d2);
                    }
#pragma unroll
                    for (int o = WIDTH / 2; o > 0; o >>= 1) d2 += __shfl_xor_sync(submask, d2, o, WIDTH);
                    if (sqrt(d2) < p.tol) break;
                }
#pragma unroll
                for (int t = 0; t < NPL; ++t) {
                    const int l = sl + WIDTH * t;
                    if (l < F) p.beta[g * F + l] = w[t];
                }
                if (sl == 0) p.flags[g] = 0;
            }
        }
        __syncwarp();
    }
}

inline cudaError_t launch_cd_solve(cudaStream_t stream, const SolveParams &sp, int sm_count) {
    const int F = sp.F;
    int width, npl;
    if (F <= 8) { width = 8; npl = 1; }
    else if (F <= 16) { width = 16; npl = 1; }
    else if (F <= 32) { width = 32; npl = 1; }
    else { width = 32; npl = 2; }
    const int subs_per_warp = 32 / width;
    const size_t per_sub = static_cast<size_t>(F) * F * sizeof(double);
    int warps = 4;
    while (warps > 1 && per_sub * subs_per_warp * warps > 200 * 1024) --warps;
    const size_t smem = per_sub * subs_per_warp * warps;
    const int64_t subs_per_block = static_cast<int64_t>(warps) * subs_per_warp;
    int64_t blocks = (sp.n_groups + subs_per_block - 1) / subs_per_block;
    const int64_t max_blocks = static_cast<int64_t>(sm_count) * std::max<int64_t>(1, std::min<int64_t>(16, (200 * 1024) / std::max<size_t>(smem, 1)));
    blocks = std::max<int64_t>(1, std::min(blocks, max_blocks));
    cudaError_t e = cudaSuccess;
#define B200_CD_LAUNCH(WD, NP)                                                                                      \
    e = cudaFuncSetAttribute(cd_solve_kernel<WD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
    if (e == cudaSuccess) cd_solve_kernel<WD, NP><<<static_cast<unsigned>(blocks), warps * 32, smem, stream>>>(sp);
    if (width == 8) { B200_CD_LAUNCH(8, 1) }
    else if (width == 16) { B200_CD_LAUNCH(16, 1) }
    else if (npl == 1) { B200_CD_LAUNCH(32, 1) }
    else { B200_CD_LAUNCH(32, 2) }
#undef B200_CD_LAUNCH
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace b200
