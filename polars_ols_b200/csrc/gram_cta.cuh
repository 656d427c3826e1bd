// gram_cta.cuh — CTA-cooperative, warp-specialised row-streaming Gram + solve kernel (k <= 16).
//
// The canonical Blackwell pipeline shape (producer / consumer / epilogue roles around mbarrier rings),
// applied to an HBM-bound f64 problem:
//   warp 0            PRODUCER : waits for a free stage, then every column slice of the next row tile of
//                                the current segment is fetched by ONE 1-D bulk async copy
//                                (cp.async.bulk -> UBLKCP, the TMA engine) — k+1(+w)(+mask) copies of up to
//                                ~8 KB each per tile — completion counted on the stage's `full` mbarrier;
//   warps 1..W        CONSUMERS: wait on `full`, split the tile's row octets between them, feed
//                                mma.sync.m8n8k4.f64 (DMMA) with 16-byte LDS fragments (A = X^T tile and
//                                B = X tile are the same register), release the stage on `empty`; at the
//                                end of a segment each publishes its accumulator fragments in a
//                                double-buffered shared-memory slot (`red_full` / `red_empty` mbarriers);
//   warps W+1..W+4    SOLVERS  : (round-robin over segments) each sums the W partial fragments in a fixed order (deterministic), runs the
//                                warp-cooperative register Cholesky (LU fallback) of gram_stream.cuh and
//                                writes beta — overlapped with the consumers' next segment.
// Short groups (team mode): when every segment fits one tile, the consumers split into W / team TEAMS of `team`
// warps and consecutive segments go to consecutive teams, so that a warp sees every (W / team)-th group only
// and its fixed per-group cost (barrier round trips, publishing the accumulator fragments) is amortised over
// team-times more rows; with team = 1 a warp owns whole groups and the solver has a single slot to read.
// One persistent CTA per SM; the unit of scheduling is the SM, so 10k groups spread over 148 SMs with
// < 1 % imbalance (a warp-per-group mapping quantises at 2.8 groups per warp), and up to STAGES whole
// tiles (~200 KB) are in flight per SM irrespective of occupancy or register pressure.
#pragma once
#include "gram_stream.cuh"

namespace b200 {

constexpr int CTA_CONSUMERS = 8;
constexpr int CTA_SOLVERS = 4;   // the k x k solve is a long dependent chain: several groups are solved concurrently
constexpr int CTA_RED_DEPTH = 6; // max ring depth of published partial-fragment buffers (GramParams::red_depth; consumers may run this far ahead)
constexpr int CTA_THREADS = (CTA_CONSUMERS + CTA_SOLVERS + 1) * 32;

template <int KB>
__host__ __device__ constexpr int cta_red_doubles() { return KB * (KB + 1) + KB + 1; }  // acc frags, cy, nfit

template <typename T>
__host__ __device__ inline size_t cta_fixed_smem(int KB, int F, int red_depth) {
    const size_t red = static_cast<size_t>(red_depth) * CTA_CONSUMERS * 32 * (KB * (KB + 1) + KB + 1) * sizeof(double);
    return red + CTA_SOLVERS * gram_scratch_bytes<T>(F, 1) + 128;
}

// ring depth without teams: every consumer publishes every segment in order, so a solver that got segment i knows
// segments < i are published; the ring must then be at least CTA_SOLVERS deep (see engine.cu: launch_gram)
__host__ __device__ constexpr int cta_plain_depth(int KB) { return KB == 1 ? CTA_RED_DEPTH : CTA_SOLVERS; }

// TEAMS = false compiles the plain pipeline (all consumers on every segment) with constant team / ring sizes
// WEIGHTED = false compiles every sample-weight path out (the C2 headline shape: 0.1175 ms; with the weight code resident
// the register allocation of the interior loop changed and the same launch took 0.1275 ms)
template <typename T, int KB, bool TEAMS, bool WEIGHTED>
__global__ void __launch_bounds__(CTA_THREADS, 1) gram_cta_kernel(const GramParams p) {
    using Vec = typename V2<T>::type;
    constexpr int NPAIR = KB * (KB + 1) / 2;
    constexpr int A = 16 / sizeof(T);
    constexpr int W = CTA_CONSUMERS;
    constexpr int RED = cta_red_doubles<KB>();

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[GRAM_MAX_STAGES], empty_bar[GRAM_MAX_STAGES], red_full[CTA_RED_DEPTH * CTA_CONSUMERS],
        red_empty[CTA_RED_DEPTH * CTA_CONSUMERS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fb = lane >> 2, q = lane & 3;
    const int kd = p.kd, F = p.F;
    const int ycol = kd, wcol = kd + 1, mcol = kd + 1 + (WEIGHTED ? 1 : 0);
    const int NC = kd + 1 + (WEIGHTED ? 1 : 0) + (p.has_mask ? 1 : 0);
    const int R = p.tile_rows, S = p.stages;
    const int TEAM = TEAMS ? ((p.team > 0 && p.team < W) ? p.team : W) : W;  // consumer warps per segment
    const int NT = W / TEAM;                                    // teams; segment i (CTA-local) -> team i % NT
    const int DEPTH = TEAMS ? ((p.red_depth > 0 && p.red_depth < CTA_RED_DEPTH) ? p.red_depth : CTA_RED_DEPTH) : cta_plain_depth(KB);
    const int RD = DEPTH * NT;                                  // ring of published buffers (TEAM slots each)
    const uint32_t stride = gram_col_stride<T>(R);
    const uint32_t stage_bytes = static_cast<uint32_t>(NC) * stride;
    double *red = reinterpret_cast<double *>(smem + static_cast<size_t>(S) * stage_bytes);
    double *Gs_base = red + DEPTH * W * 32 * RED;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], TEAM);
        }
        for (int b = 0; b < RD; ++b) {
            mbar_init(&red_full[b], TEAM);
            mbar_init(&red_empty[b], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t nseg = p.nseg;
    const bool plain = !p.has_mask;  // no row mask: interior octets need no predication (weights only scale)

    if (warp == 0) {
        // ================================ PRODUCER ================================
        int stage = 0;
        uint32_t phase = 0;
        // segment bounds are fetched one segment ahead (their HBM/L2 latency would otherwise sit on the critical
        // path of every small group)
        int64_t nr0 = 0, nr1 = 0;
        if (static_cast<int64_t>(blockIdx.x) < nseg) { nr0 = p.seg_off[blockIdx.x]; nr1 = p.seg_off[blockIdx.x + 1]; }
        for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
            const int64_t r0 = nr0, r1 = nr1;
            if (seg + gridDim.x < nseg) { nr0 = p.seg_off[seg + gridDim.x]; nr1 = p.seg_off[seg + gridDim.x + 1]; }
            // team mode maps tile index == segment index, so an empty segment still takes one (empty) tile
            for (int64_t row = r0; row < r1 || (NT > 1 && row == r0); row += R) {
                const int64_t b = (row + R < r1) ? row + R : r1;
                const int64_t a_al = row & ~static_cast<int64_t>(A - 1);
                int64_t b_al = (b + (A - 1)) & ~static_cast<int64_t>(A - 1);
                if (b_al > p.n_rows_pad) b_al = p.n_rows_pad;
                const uint32_t bytes = static_cast<uint32_t>(b_al - a_al) * sizeof(T);
                mbar_wait(&empty_bar[stage], phase ^ 1u);  // all consumers released this stage
                unsigned char *sb = smem + static_cast<size_t>(stage) * stage_bytes;
                if (lane == 0) {
                    fence_proxy_async_smem();
                    mbar_arrive_expect_tx(&full_bar[stage], bytes * static_cast<uint32_t>(NC));
                }
                __syncwarp();
                if (bytes)
                    for (int c = lane; c < NC; c += 32)
                        bulk_g2s(sb + static_cast<size_t>(c) * stride, static_cast<const T *>(p.cols[c]) + a_al, bytes,
                                 &full_bar[stage]);
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp <= W) {
        // ================================ CONSUMERS ================================
        const int cw = (warp - 1) % TEAM, team_id = (warp - 1) / TEAM;
        // team mode (NT > 1) requires one tile per segment, so tile index == CTA-local segment index
        uint32_t red_i = team_id;
        int stage = team_id % S;
        uint32_t phase = (team_id / S) & 1u;
        const int64_t seg_step = static_cast<int64_t>(NT) * gridDim.x;
        const int64_t seg_first = blockIdx.x + static_cast<int64_t>(team_id) * gridDim.x;
        int64_t nr0 = 0, nr1 = 0;
        if (seg_first < nseg) { nr0 = p.seg_off[seg_first]; nr1 = p.seg_off[seg_first + 1]; }
        for (int64_t seg = seg_first; seg < nseg; seg += seg_step) {
            const int64_t r0 = nr0, r1 = nr1;
            if (seg + seg_step < nseg) { nr0 = p.seg_off[seg + seg_step]; nr1 = p.seg_off[seg + seg_step + 1]; }
            constexpr bool DUAL = KB <= 2;
            double acc[NPAIR][2], acc2[DUAL ? NPAIR : 1][2];
            double cy[KB];
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
            for (int i = 0; i < (DUAL ? NPAIR : 1); ++i) acc2[i][0] = acc2[i][1] = 0.0;
#pragma unroll
            for (int i = 0; i < KB; ++i) cy[i] = 0.0;
            int nfit = 0;

            for (int64_t row = r0; row < r1 || (NT > 1 && row == r0); row += R) {
                const int64_t b = (row + R < r1) ? row + R : r1;
                const int o = static_cast<int>(row & (A - 1));
                const int hi = o + static_cast<int>(b - row);  // valid local rows are [o, hi)
                const int noct = (hi + 7) >> 3;
                const int per = (noct + TEAM - 1) / TEAM;
                const int j0 = cw * per;
                const int j1 = (j0 + per < noct) ? j0 + per : noct;
                mbar_wait(&full_bar[stage], phase);
                const unsigned char *sb = smem + static_cast<size_t>(stage) * stage_bytes;
                const unsigned char *xs[KB];
#pragma unroll
                for (int bk = 0; bk < KB; ++bk)
                    xs[bk] = sb + static_cast<size_t>((8 * bk + fb < kd) ? 8 * bk + fb : 0) * stride + 2 * q * sizeof(T);
                const unsigned char *ys = sb + static_cast<size_t>(ycol) * stride + 2 * q * sizeof(T);
                if (WEIGHTED && !p.w_is_sqrt) {
                    // sqrt(w) ONCE per row, in place in the stage (this warp's rows only), instead of in every lane of
                    // every fragment load: the IEEE sqrt is a ~25-instruction sequence and the 8 lanes that share a row
                    // pair all computed it (ncu on C3: 37 % of the kernel's instructions)
                    T *wc = reinterpret_cast<T *>(const_cast<unsigned char *>(sb) + static_cast<size_t>(wcol) * stride);
                    for (int i = 8 * j0 + lane; i < 8 * j1; i += 32) wc[i] = static_cast<T>(sqrt(wc[i]));
                    __syncwarp();
                }
                bool has_x[KB];
                double xconst[KB];
#pragma unroll
                for (int bk = 0; bk < KB; ++bk) {
                    has_x[bk] = 8 * bk + fb < kd;
                    xconst[bk] = ((8 * bk + fb == kd) && p.intercept) ? 1.0 : 0.0;
                }

                auto mma_octet = [&](const double (&f0)[KB], const double (&f1)[KB], double y0, double y1) {
                    int idx = 0;
#pragma unroll
                    for (int bi = 0; bi < KB; ++bi) {
#pragma unroll
                        for (int bj = bi; bj < KB; ++bj) {
                            dmma_m8n8k4(acc[idx][0], acc[idx][1], f0[bi], f0[bj]);
                            if (DUAL) dmma_m8n8k4(acc2[idx][0], acc2[idx][1], f1[bi], f1[bj]);
                            else dmma_m8n8k4(acc[idx][0], acc[idx][1], f1[bi], f1[bj]);
                            ++idx;
                        }
                        cy[bi] = fma(f0[bi], y0, cy[bi]);
                        cy[bi] = fma(f1[bi], y1, cy[bi]);
                    }
                };
                auto masked_octet = [&](int j) {
                    const int lr = 8 * j + 2 * q;
                    bool v0 = (lr >= o) && (lr < hi);
                    bool v1 = (lr + 1 >= o) && (lr + 1 < hi);
                    const Vec y2 = *reinterpret_cast<const Vec *>(ys + 8 * j * sizeof(T));
                    T s0 = T(1), s1 = T(1);
                    if (p.has_mask) {
                        const Vec m2 = *reinterpret_cast<const Vec *>(sb + static_cast<size_t>(mcol) * stride + lr * sizeof(T));
                        v0 = v0 && (m2.x != T(0));
                        v1 = v1 && (m2.y != T(0));
                    }
                    if (WEIGHTED) {
                        const Vec w2 = *reinterpret_cast<const Vec *>(sb + static_cast<size_t>(wcol) * stride + lr * sizeof(T));
                        s0 = w2.x;
                        s1 = w2.y;
                    }
                    const double y0 = v0 ? static_cast<double>(static_cast<T>(y2.x * s0)) : 0.0;
                    const double y1 = v1 ? static_cast<double>(static_cast<T>(y2.y * s1)) : 0.0;
                    if (fb == 0) nfit += (v0 ? 1 : 0) + (v1 ? 1 : 0);
                    double f0[KB], f1[KB];
#pragma unroll
                    for (int bk = 0; bk < KB; ++bk) {
                        const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                        const T x0 = has_x[bk] ? x2.x : static_cast<T>(xconst[bk]);
                        const T x1 = has_x[bk] ? x2.y : static_cast<T>(xconst[bk]);
                        f0[bk] = v0 ? static_cast<double>(static_cast<T>(x0 * s0)) : 0.0;
                        f1[bk] = v1 ? static_cast<double>(static_cast<T>(x1 * s1)) : 0.0;
                    }
                    mma_octet(f0, f1, y0, y1);
                };

                if (!plain) {
                    for (int j = j0; j < j1; ++j) masked_octet(j);
                } else {
                    int j = j0;
                    if (j < j1 && j == 0 && o != 0) {
                        masked_octet(0);
                        j = 1;
                    }
                    const int jfull = hi >> 3;  // octets below jfull lie entirely inside [o, hi)
                    const int jend = (j1 < jfull) ? j1 : jfull;
                    if (!WEIGHTED) {
#pragma unroll 4
                        for (; j < jend; ++j) {
                            const Vec y2 = *reinterpret_cast<const Vec *>(ys + 8 * j * sizeof(T));
                            double f0[KB], f1[KB];
#pragma unroll
                            for (int bk = 0; bk < KB; ++bk) {
                                const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                                f0[bk] = has_x[bk] ? static_cast<double>(x2.x) : xconst[bk];
                                f1[bk] = has_x[bk] ? static_cast<double>(x2.y) : xconst[bk];
                            }
                            mma_octet(f0, f1, static_cast<double>(y2.x), static_cast<double>(y2.y));
                        }
                    } else {
                        const unsigned char *ws = sb + static_cast<size_t>(wcol) * stride + 2 * q * sizeof(T);
#pragma unroll 2
                        for (; j < jend; ++j) {
                            const Vec y2 = *reinterpret_cast<const Vec *>(ys + 8 * j * sizeof(T));
                            const Vec w2 = *reinterpret_cast<const Vec *>(ws + 8 * j * sizeof(T));
                            const T s0 = w2.x;
                            const T s1 = w2.y;
                            double f0[KB], f1[KB];
#pragma unroll
                            for (int bk = 0; bk < KB; ++bk) {
                                const Vec x2 = *reinterpret_cast<const Vec *>(xs[bk] + 8 * j * sizeof(T));
                                f0[bk] = static_cast<double>(static_cast<T>((has_x[bk] ? x2.x : static_cast<T>(xconst[bk])) * s0));
                                f1[bk] = static_cast<double>(static_cast<T>((has_x[bk] ? x2.y : static_cast<T>(xconst[bk])) * s1));
                            }
                            mma_octet(f0, f1, static_cast<double>(static_cast<T>(y2.x * s0)), static_cast<double>(static_cast<T>(y2.y * s1)));
                        }
                    }
                    for (; j < j1; ++j) masked_octet(j);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
                if (NT == 1) {
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1u;
                    }
                } else {  // next tile of this team: NT tiles further on
                    const uint32_t t = red_i + NT;
                    stage = static_cast<int>(t % S);
                    phase = (t / S) & 1u;
                }
            }
            // ---- publish this warp's partial fragments ----
            const int buf = red_i % RD;
            mbar_wait(&red_empty[buf], ((red_i / RD) & 1u) ^ 1u);
            double *slot = red + (static_cast<size_t>(buf) * TEAM + cw) * 32 * RED;
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
                slot[(2 * i) * 32 + lane] = DUAL ? acc[i][0] + acc2[i][0] : acc[i][0];
                slot[(2 * i + 1) * 32 + lane] = DUAL ? acc[i][1] + acc2[i][1] : acc[i][1];
            }
#pragma unroll
            for (int i = 0; i < KB; ++i) slot[(2 * NPAIR + i) * 32 + lane] = cy[i];
            slot[(2 * NPAIR + KB) * 32 + lane] = static_cast<double>(nfit);
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_full[buf]);
            red_i += NT;
        }
    } else {
        // ================================ SOLVERS ================================
        // solver s takes the segments whose CTA-local index is congruent to s (mod CTA_SOLVERS)
        const int sw = warp - (W + 1);
        double *Gs = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(Gs_base) + static_cast<size_t>(sw) * gram_scratch_bytes<T>(F, 1));
        uint32_t red_i = sw;
        for (int64_t seg = blockIdx.x + static_cast<int64_t>(sw) * gridDim.x; seg < nseg;
             seg += static_cast<int64_t>(CTA_SOLVERS) * gridDim.x, red_i += CTA_SOLVERS) {
            const int buf = red_i % RD;
            mbar_wait(&red_full[buf], (red_i / RD) & 1u);
            double acc[NPAIR][2], cy[KB];
            double nf = 0.0;
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
            for (int i = 0; i < KB; ++i) cy[i] = 0.0;
            for (int cw = 0; cw < TEAM; ++cw) {  // fixed order: deterministic sums
                const double *slot = red + (static_cast<size_t>(buf) * TEAM + cw) * 32 * RED;
#pragma unroll
                for (int i = 0; i < NPAIR; ++i) {
                    acc[i][0] += slot[(2 * i) * 32 + lane];
                    acc[i][1] += slot[(2 * i + 1) * 32 + lane];
                }
#pragma unroll
                for (int i = 0; i < KB; ++i) cy[i] += slot[(2 * NPAIR + i) * 32 + lane];
                nf += slot[(2 * NPAIR + KB) * 32 + lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_empty[buf]);
            int nfit = static_cast<int>(nf);
            if (plain) nfit = (lane == 0) ? static_cast<int>(p.seg_off[seg + 1] - p.seg_off[seg]) : 0;
            gram_epilogue<KB>(p, acc, cy, nfit, seg, Gs, lane);
        }
        if (p.flag_n > 0) {
            // per-step completion of the fused gather without a second launch: every solver warp orders its peer stores
            // before a device-wide arrival count; the warp that arrives last knows every beta of this launch has been
            // stored into every peer, so it publishes the step (release) and waits for the peers' previous step (acquire)
            __threadfence_system();
            unsigned int old = 0;
            if (lane == 0) old = atomicAdd(p.done_counter, 1u);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == gridDim.x * CTA_SOLVERS - 1) {
                if (lane == 0) *p.done_counter = 0;  // re-arm for the next launch on this stream
                __threadfence_system();
                if (lane < p.flag_n) {
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.flag_peer[lane] + p.flag_rank), "l"(p.flag_step) : "memory");
                    const unsigned long long *mine = p.flag_peer[p.flag_rank] + lane;
                    const long long t0 = clock64();
                    for (;;) {
                        unsigned long long v;
                        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
                        if (v >= p.flag_wait) break;
                        if (clock64() - t0 > 4000000000LL) {  // ~2 s: a peer died; do not hang the device
                            *p.flag_timeout = 1;
                            break;
                        }
                    }
                }
            }
        }
    }
}

template <typename T, int KB, bool TEAMS, bool WEIGHTED>
cudaError_t gram_cta_launch_w(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s);

template <typename T, int KB, bool TEAMS>
cudaError_t gram_cta_launch_t(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s) {
    if (p.has_w) return gram_cta_launch_w<T, KB, TEAMS, true>(p, grid, smem, s);
    return gram_cta_launch_w<T, KB, TEAMS, false>(p, grid, smem, s);
}

template <typename T, int KB, bool TEAMS, bool WEIGHTED>
cudaError_t gram_cta_launch_w(const GramParams &p, unsigned grid, size_t smem, cudaStream_t s) {
    auto kern = gram_cta_kernel<T, KB, TEAMS, WEIGHTED>;
    static size_t attr_set[64] = {};  // per device: the opt-in shared-memory size already granted to this kernel
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem > attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = smem;
    }
    kern<<<grid, CTA_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t gram_cta_launch_f64(int KB, const GramParams &p, unsigned grid, size_t smem, cudaStream_t s);
cudaError_t gram_cta_launch_f32(int KB, const GramParams &p, unsigned grid, size_t smem, cudaStream_t s);

}  // namespace b200
