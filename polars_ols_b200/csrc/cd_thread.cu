// cd_thread.cu — elastic-net / lasso coordinate descent, ONE THREAD PER GROUP (k <= 16, many groups: config C3).
//
// Same algorithm, order of operations and stop rule as cd_solve_kernel (cd_solve.cuh: solve_elastic_net,
// src/least_squares.rs:386-492 of /root/reference, in Gram form) — results are bit-identical — but a different mapping.
// cd_solve_kernel gives a group a sub-warp (one coordinate per lane) and pays ~25 issue slots per coordinate step for two
// groups per warp (shuffle broadcast of the step, ballots, selects): ncu shows it issue bound, 0.47 ms for C3's 100k
// groups.  Here a lane owns a whole group: w, q = X'y - G w, diag(G) and 1 / (G_jj + l2) live in registers (the
// coordinate loop is fully unrolled, every index is static), the strict upper triangle of G sits in shared memory
// transposed ([entry][lane], padded to 33 so both the cooperative fill and the per-lane reads are conflict-free: 31 KB
// per warp for k = 16, seven warps per SM), and a coordinate step is ~8 dependent FP64 operations plus FP fused
// multiply-adds on q — for 32 groups per warp, without a single shuffle.  Lanes leave the sweep loop one by one as their
// groups converge.
//
// The stop rule sums (w - w_old)^2 in the order of the sub-warp butterfly of cd_solve_kernel (xor 8, 4, 2, 1 resp.
// 4, 2, 1; floating-point addition is commutative, so every lane of that butterfly holds the same value), which keeps
// the sweep counts — and therefore the coefficients — identical to the other kernel's.
#include <cuda_runtime.h>

#include <cstdint>

#include "cd_thread.h"
#include "solvers.cuh"

namespace b200 {

constexpr int CDT_PAD = 33;

// strictly-upper-triangle slot of (a, b), a < b, rows packed one after another
template <int FP> __host__ __device__ constexpr int cdt_slot(int a, int b) { return a * (FP - 1) - a * (a - 1) / 2 + (b - a - 1); }
template <int FP> __host__ __device__ constexpr int cdt_entries() { return FP * (FP - 1) / 2; }

template <int FP, bool ACTIVE, bool POSITIVE>
__global__ void __launch_bounds__(32) cd_thread_kernel(const SolveParams p) {
    extern __shared__ __align__(16) unsigned char cdt_smem[];
    double *S = reinterpret_cast<double *>(cdt_smem);  // [cdt_entries][CDT_PAD]
    constexpr int NE = cdt_entries<FP>();
    constexpr int PER_LANE = (NE + 31) / 32;
    const int F = p.F, lane = threadIdx.x;
    const size_t P = static_cast<size_t>(F) * F + F + 1;
    const int64_t n_tiles = (p.n_groups + 31) / 32;
    const double tol = p.tol;
    const int64_t max_iter = p.max_iter;

    // record offset of every off-diagonal entry this lane fetches during the fill (-1: padding -> 0)
    int rec[PER_LANE];
#pragma unroll
    for (int c = 0; c < PER_LANE; ++c) {
        const int u = lane + 32 * c;
        int e = -1;
        if (u < NE) {
            int a = 0, rem = u;
            while (rem >= FP - 1 - a) { rem -= FP - 1 - a; ++a; }
            const int b = a + 1 + rem;
            if (b < F) e = a * F + b;
        }
        rec[c] = e;
    }
    const bool any_below = tol > 0.0;  // tol <= 0 (or NaN): nothing is ever below it
    double cut = tol * tol;            // d2 <= cut  <=>  sqrt(d2) < tol (see cd_solve.cuh)
    if (any_below) {
        while (cut > 0.0 && sqrt(cut) >= tol) cut = __longlong_as_double(__double_as_longlong(cut) - 1);
        while (cut < 1.0e300 && sqrt(__longlong_as_double(__double_as_longlong(cut) + 1)) < tol)
            cut = __longlong_as_double(__double_as_longlong(cut) + 1);
    }

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t g_base = tile * 32;
        const int64_t g = g_base + lane;
        const bool live = g < p.n_groups;
        __syncwarp();
        // ---- fill.  One segment per group (the usual case): the warp copies one record at a time with 8-byte cp.async
        // straight into the transposed tile, every load of the tile in flight at once (the fill is pure latency).
        // Groups split into segments: fixed-order sum over the segments, as cd_solve_kernel does. ----
        if (!p.group_seg_off) {
            const int n_live = static_cast<int>(p.n_groups - g_base < 32 ? p.n_groups - g_base : 32);
            const double *rec0 = p.partial + static_cast<size_t>(g_base) * P;
#pragma unroll
            for (int c = 0; c < PER_LANE; ++c) {
                const int u = lane + 32 * c;
                if (u >= NE) break;
                double *dst = S + u * CDT_PAD;
                if (rec[c] >= 0) {
                    const double *src = rec0 + rec[c];
                    for (int gi = 0; gi < n_live; ++gi)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst + gi))),
                                     "l"(src + static_cast<size_t>(gi) * P)
                                     : "memory");
                } else {
                    for (int gi = 0; gi < n_live; ++gi) dst[gi] = 0.0;
                }
            }
        } else {
            for (int gi = 0; gi < 32; ++gi) {
                const int64_t gg = g_base + gi;
                if (gg >= p.n_groups) break;
                const int64_t s0 = p.group_seg_off[gg], s1 = p.group_seg_off[gg + 1];
#pragma unroll
                for (int c = 0; c < PER_LANE; ++c) {
                    const int u = lane + 32 * c;
                    if (u >= NE) break;
                    double v = 0.0;
                    if (rec[c] >= 0)
                        for (int64_t sg = s0; sg < s1; ++sg) v += p.partial[static_cast<size_t>(sg) * P + rec[c]];
                    S[u * CDT_PAD + gi] = v;
                }
            }
        }
        // the lane's own group: diag(G), X'y and the row count straight into registers (under the copies above)
        double q[FP], w[FP], dsq[FP], gd[FP], rinv[FP];
        double nfit = 0.0;
        {
            const int64_t s0 = live ? (p.group_seg_off ? p.group_seg_off[g] : g) : 0;
            const int64_t s1 = live ? (p.group_seg_off ? p.group_seg_off[g + 1] : g + 1) : 0;
#pragma unroll
            for (int l = 0; l < FP; ++l) {
                q[l] = 0.0;
                gd[l] = (live && l < F) ? 0.0 : 1.0;
            }
            for (int64_t sg = s0; sg < s1; ++sg) {
                const double *r = p.partial + static_cast<size_t>(sg) * P;
#pragma unroll
                for (int l = 0; l < FP; ++l)
                    if (l < F) {
                        gd[l] += __ldg(r + l * F + l);
                        q[l] += __ldg(r + F * F + l);
                    }
                nfit += __ldg(r + F * F + F);
            }
        }
        // alpha *= n_samples (src/least_squares.rs:419); products rounded on their own (no contraction into the
        // subtraction / addition that follows), as the reference computes alpha * l1_ratio and alpha * (1 - l1_ratio)
        const double a = __dmul_rn(p.alpha, nfit);
        const double l1 = __dmul_rn(a, p.l1_ratio), l2 = __dmul_rn(a, 1.0 - p.l1_ratio);
#pragma unroll
        for (int l = 0; l < FP; ++l) {
            w[l] = 0.0;
            dsq[l] = 0.0;
            rinv[l] = 1.0 / (gd[l] + l2);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        const double *Sg = S + lane;
        uint32_t active = (F >= 32) ? 0xffffffffu : ((1u << F) - 1u);
        if (live && nfit != 0.0) {
            for (int64_t sweep = 0; sweep < max_iter; ++sweep) {
                const uint32_t loop_set = active;  // the reference iterates over a clone of the active list (:459)
#pragma unroll
                for (int j = 0; j < FP; ++j) {
                    if (ACTIVE) {
                        dsq[j] = 0.0;
                        if (!((loop_set >> j) & 1u)) continue;
                    } else if (j >= F) {
                        continue;  // padding coordinate (uniform)
                    }
                    // soft_threshold (src/least_squares.rs:373-379): |rho| - l1 kept when positive (a NaN is not kept, like
                    // f64::max), with the sign of rho; clamped at zero when `positive`
                    const double rho = fma(gd[j], w[j], q[j]);
                    const double av = fabs(rho) - l1;
                    const bool keep = POSITIVE ? (av > 0.0 && rho > 0.0) : (av > 0.0);
                    const double wn = (keep ? (POSITIVE ? av : copysign(av, rho)) : 0.0) * rinv[j];
                    const double delta = wn - w[j];  // == w_new - w_old of this sweep: a coordinate is visited once
                    dsq[j] = delta * delta;
                    if (ACTIVE && fabs(wn) < tol) active &= ~(1u << j);  // never re-admitted (:472-476)
                    w[j] = wn;
                    // q -= G[j][:] delta, row j of the symmetric G from its upper triangle.  No `delta != 0` shortcut: with
                    // finite data the update is then an exact no-op, and a NaN row poisons q — and with it every later
                    // coordinate — exactly as x_j * w_j poisons the reference's residual vector (:428-433).
#pragma unroll
                    for (int l = 0; l < FP; ++l) {
                        const double gjl = l == j ? gd[j] : Sg[(l < j ? cdt_slot<FP>(l, j) : cdt_slot<FP>(j, l)) * CDT_PAD];
                        q[l] = fma(-gjl, delta, q[l]);
                    }
                }
                // ||w - w_old||^2 summed in the order of cd_solve_kernel's sub-warp butterfly (xor FP/2, .., 1)
                double t[FP];
#pragma unroll
                for (int l = 0; l < FP; ++l) t[l] = dsq[l];
#pragma unroll
                for (int o = FP / 2; o > 0; o >>= 1) {
#pragma unroll
                    for (int l = 0; l < o; ++l) t[l] += t[l + o];
                }
                if (any_below && t[0] <= cut) break;
            }
        }
        if (live) {
            double *bo = p.beta + g * F;
#pragma unroll
            for (int l = 0; l < FP; ++l)
                if (l < F) bo[l] = (nfit == 0.0) ? 0.0 : w[l];  // src/expressions.rs:357-359: no rows -> zeros
            p.flags[g] = (nfit == 0.0) ? FLAG_EMPTY : 0;
        }
    }
}

template <int FP, bool ACTIVE, bool POSITIVE>
static cudaError_t cd_thread_launch_k(cudaStream_t stream, const SolveParams &sp, int sm_count, int blocks_per_sm) {
    const size_t smem = static_cast<size_t>(cdt_entries<FP>()) * CDT_PAD * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(cd_thread_kernel<FP, ACTIVE, POSITIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const int fit = static_cast<int>((227u * 1024u) / (smem + 1024));
    const int per_sm = blocks_per_sm > 0 ? std::min(blocks_per_sm, fit) : std::min(fit, 8);  // 8: registers (235 per thread)
    const int64_t tiles = (sp.n_groups + 31) / 32;
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(tiles, static_cast<int64_t>(sm_count) * per_sm));
    cd_thread_kernel<FP, ACTIVE, POSITIVE><<<static_cast<unsigned>(blocks), 32, smem, stream>>>(sp);
    return cudaGetLastError();
}

template <int FP>
static cudaError_t cd_thread_launch_t(cudaStream_t stream, const SolveParams &sp, int sm_count, int blocks_per_sm) {
    const bool active = sp.route == ROUTE_CD_ACTIVE, positive = sp.positive != 0;
    if (active) return positive ? cd_thread_launch_k<FP, true, true>(stream, sp, sm_count, blocks_per_sm) : cd_thread_launch_k<FP, true, false>(stream, sp, sm_count, blocks_per_sm);
    return positive ? cd_thread_launch_k<FP, false, true>(stream, sp, sm_count, blocks_per_sm) : cd_thread_launch_k<FP, false, false>(stream, sp, sm_count, blocks_per_sm);
}

cudaError_t launch_cd_thread(cudaStream_t stream, const SolveParams &sp, int sm_count, int blocks_per_sm) {
    if (sp.F <= 8) return cd_thread_launch_t<8>(stream, sp, sm_count, blocks_per_sm);
    if (sp.F <= 16) return cd_thread_launch_t<16>(stream, sp, sm_count, blocks_per_sm);
    return cudaErrorInvalidValue;
}

}  // namespace b200
