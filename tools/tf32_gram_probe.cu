// tf32_gram_probe.cu — the measurement behind DESIGN.md §4.11: what tcgen05 (kind::tf32, 3-way split) would give the
// f32 config C3 (100k groups x 256 rows x 16 columns).  Standalone (not part of libb200ols.so):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tf32_gram_probe tools/tf32_gram_probe.cu
//   tools/tf32_gram_probe            # numerics of 8 groups + tensor-side throughput on all SMs
//
// Mapping: UMMA tiles are M in {64, 128} rows, a group's Gram is 16 x 16, so 8 groups share one operand:
//   A[m][k] = X_g[k][f]   m = 16 g + f (8 groups x 16 columns = 128), k = data row;   D = A A^T  (128 x 128, f32 in TMEM)
// and only the eight 16 x 16 diagonal blocks of D are Grams (12.5 % of the tensor work is useful).  f32 -> tf32 keeps 10
// mantissa bits, so x = hi + lo with hi = tf32(x), lo = tf32(x - hi) and D = hi hi^T + hi lo^T + lo hi^T (3 MMAs per K-step).
// Operand tiles: canonical K-major, no swizzle — 8-row x 16-byte core matrices, LBO (next 16-byte K chunk) = 128 B,
// SBO (next 8 rows) = 256 B, 4 KB per K-step of 8 rows (cute/arch/mma_sm100_desc.hpp: SmemDescriptor / InstrDescriptor).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int M = 128, N = 128, KSTEP = 8;      // one tcgen05.mma: 128 x 128 x 8 (tf32)
constexpr int ROWS = 256, COLS = 16, GROUPS = 8;  // per CTA tile: 8 groups x 256 rows x 16 columns
constexpr int CHUNK_ROWS = 64;                   // data rows staged per shared-memory chunk (8 K-steps)
constexpr int STEP_BYTES = (M / 8) * 256;        // 4 KB per K-step
constexpr int CHUNK_BYTES = (CHUNK_ROWS / KSTEP) * STEP_BYTES;  // 32 KB per split tile

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (static_cast<uint64_t>((saddr >> 4) & 0x3FFF)) | (static_cast<uint64_t>(128 >> 4) << 16) | (static_cast<uint64_t>(256 >> 4) << 32) |
           (1ull << 46);  // version 1 (sm_100), layout_type 0 = SWIZZLE_NONE (interleaved core matrices)
}
// c_format F32 (1) @4, a/b_format TF32 (2) @7 / @10, K-major both, n_dim = N >> 3 @17, m_dim = M >> 4 @24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ int g_timeout = 0;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 2000000000LL) {  // ~1 s: the commit never arrived; give up instead of hanging the device
            g_timeout = 1;
            break;
        }
    }
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// x: [GROUPS][COLS][ROWS] f32 per CTA tile (SoA per group).  gram: [GROUPS][COLS][COLS] f32 (numerics mode).
// reps > 0: throughput mode — the staged chunk is multiplied `reps` times (tensor side only, results discarded).
__global__ void __launch_bounds__(128, 1) tf32_gram_kernel(const float *__restrict__ x, float *__restrict__ gram, int reps, int split) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char *hi = smem, *lo = smem + CHUNK_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5;
    const float *xt = x + static_cast<size_t>(blockIdx.x) * GROUPS * COLS * ROWS;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    uint32_t parity = 0;
    bool first = true;
    const int n_chunks = ROWS / CHUNK_ROWS;
    for (int ch = 0; ch < n_chunks; ++ch) {
        // stage: split every element into hi / lo tf32 tiles in the canonical K-major core-matrix layout
        for (int e = tid; e < GROUPS * COLS * CHUNK_ROWS; e += blockDim.x) {
            const int k = e % CHUNK_ROWS, m = e / CHUNK_ROWS;  // m = 16 g + f
            const float v = xt[static_cast<size_t>(m) * ROWS + ch * CHUNK_ROWS + k];
            const float h = to_tf32(v), l = to_tf32(v - h);
            const int off = (k / KSTEP) * STEP_BYTES + (m / 8) * 256 + ((k % KSTEP) / 4) * 128 + (m % 8) * 16 + (k % 4) * 4;
            *reinterpret_cast<float *>(hi + off) = h;
            *reinterpret_cast<float *>(lo + off) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const int loops = reps > 0 ? reps : 1;
            for (int rp = 0; rp < loops; ++rp) {
                for (int s = 0; s < CHUNK_ROWS / KSTEP; ++s) {
                    const uint64_t dh = make_desc(smem_u32(hi + s * STEP_BYTES)), dl = make_desc(smem_u32(lo + s * STEP_BYTES));
                    umma_tf32(tmem_d, dh, dh, first ? 0u : 1u);
                    first = false;
                    if (split) {
                        umma_tf32(tmem_d, dh, dl, 1u);
                        umma_tf32(tmem_d, dl, dh, 1u);
                    }
                }
            }
            umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (gram) {
        // epilogue: row m of D sits in TMEM lane m; warp w owns lanes 32 w .. 32 w + 31; the Gram of group g is the
        // 16 x 16 block at rows / columns 16 g
        const int m = tid, g = m / COLS, f = m % COLS;
        // tcgen05.ld is warp-collective with ONE address: the warp reads the 32 columns of its two groups and every
        // lane keeps the half that belongs to its own group
        uint32_t ra[16], rb[16], r[16];
        const uint32_t taddr = tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(warp * 32);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(ra[0]), "=r"(ra[1]), "=r"(ra[2]), "=r"(ra[3]), "=r"(ra[4]), "=r"(ra[5]), "=r"(ra[6]), "=r"(ra[7]), "=r"(ra[8]), "=r"(ra[9]),
                       "=r"(ra[10]), "=r"(ra[11]), "=r"(ra[12]), "=r"(ra[13]), "=r"(ra[14]), "=r"(ra[15])
                     : "r"(taddr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(rb[0]), "=r"(rb[1]), "=r"(rb[2]), "=r"(rb[3]), "=r"(rb[4]), "=r"(rb[5]), "=r"(rb[6]), "=r"(rb[7]), "=r"(rb[8]), "=r"(rb[9]),
                       "=r"(rb[10]), "=r"(rb[11]), "=r"(rb[12]), "=r"(rb[13]), "=r"(rb[14]), "=r"(rb[15])
                     : "r"(taddr + 16u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = (tid & 16) ? rb[j] : ra[j];
        float *out = gram + (static_cast<size_t>(blockIdx.x) * GROUPS + g) * COLS * COLS + f * COLS;
        for (int j = 0; j < 16; ++j) out[j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(128) : "memory");
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t tile = static_cast<size_t>(GROUPS) * COLS * ROWS;
    std::vector<float> hx(tile * sms);
    srand(1);
    for (auto &v : hx) v = static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f + 0.3f;
    float *dx = nullptr, *dg = nullptr;
    CK(cudaMalloc(&dx, hx.size() * sizeof(float)));
    CK(cudaMalloc(&dg, static_cast<size_t>(sms) * GROUPS * COLS * COLS * sizeof(float)));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * sizeof(float), cudaMemcpyHostToDevice));
    const size_t smem = 2 * CHUNK_BYTES + 1024;
    CK(cudaFuncSetAttribute(tf32_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));

    // ---- numerics: one-pass TF32 and the 3-way split against the f64 Gram of the f32 inputs (first CTA tile) ----
    for (int split = 0; split <= 1; ++split) {
        tf32_gram_kernel<<<1, 128, smem>>>(dx, dg, 0, split);
        CK(cudaDeviceSynchronize());
        std::vector<float> hg(static_cast<size_t>(GROUPS) * COLS * COLS);
        CK(cudaMemcpy(hg.data(), dg, hg.size() * sizeof(float), cudaMemcpyDeviceToHost));
        double worst = 0.0;
        for (int g = 0; g < GROUPS; ++g)
            for (int a = 0; a < COLS; ++a)
                for (int b = 0; b < COLS; ++b) {
                    double ref = 0.0;
                    for (int k = 0; k < ROWS; ++k)
                        ref += static_cast<double>(hx[(static_cast<size_t>(g) * COLS + a) * ROWS + k]) * static_cast<double>(hx[(static_cast<size_t>(g) * COLS + b) * ROWS + k]);
                    const double diag = 256.0 / 3.0 + 256 * 0.09;  // scale of a diagonal entry
                    worst = std::fmax(worst, std::fabs(hg[(static_cast<size_t>(g) * COLS + a) * COLS + b] - ref) / diag);
                }
        printf("{\"probe\": \"numerics\", \"split\": %d, \"max_abs_err_over_diag_scale\": %.3e, \"g0_row0_first4\": [%.6f, %.6f, %.6f, %.6f], "
               "\"g1_row1_first4\": [%.6f, %.6f, %.6f, %.6f]}\n",
               split ? 3 : 1, worst, hg[0], hg[1], hg[2], hg[3], hg[256 + 16], hg[256 + 17], hg[256 + 18], hg[256 + 19]);
        if (!split) {
            double r0[4] = {0, 0, 0, 0}, r1[4] = {0, 0, 0, 0};
            for (int b = 0; b < 4; ++b)
                for (int k = 0; k < ROWS; ++k) {
                    r0[b] += static_cast<double>(hx[k]) * hx[static_cast<size_t>(b) * ROWS + k];
                    r1[b] += static_cast<double>(hx[(16 + 1) * static_cast<size_t>(ROWS) + k]) * hx[(16 + b) * static_cast<size_t>(ROWS) + k];
                }
            printf("{\"probe\": \"reference\", \"g0_row0_first4\": [%.6f, %.6f, %.6f, %.6f], \"g1_row1_first4\": [%.6f, %.6f, %.6f, %.6f]}\n", r0[0], r0[1],
                   r0[2], r0[3], r1[0], r1[1], r1[2], r1[3]);
        }
    }

    // ---- tensor-side throughput: all SMs, 1 CTA each, the staged chunk multiplied `reps` times per chunk ----
    const int reps = 64;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int split = 0; split <= 1; ++split) {
        tf32_gram_kernel<<<sms, 128, smem>>>(dx, nullptr, reps, split);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        tf32_gram_kernel<<<sms, 128, smem>>>(dx, nullptr, reps, split);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double mmas = static_cast<double>(sms) * (ROWS / CHUNK_ROWS) * reps * (CHUNK_ROWS / KSTEP) * (split ? 3 : 1);
        const double rate = mmas / (ms * 1e-3);
        const double tflops = rate * 2.0 * M * N * KSTEP / 1e12;
        // C3: 100000 groups / 8 per tile x 32 K-steps x (3 or 1) MMAs
        const double c3_mmas = 100000.0 / GROUPS * (ROWS / KSTEP) * (split ? 3 : 1);
        printf("{\"probe\": \"throughput\", \"split\": %d, \"mma_128x128x8_per_s\": %.4e, \"tflops_issued\": %.1f, \"includes_staging_of_chunks\": true, "
               "\"c3_tensor_ms_at_this_rate\": %.3f, \"useful_fraction_of_flops\": 0.125}\n",
               split ? 3 : 1, rate, tflops, c3_mmas / rate * 1e3);
    }
    int to = 0;
    CK(cudaMemcpyFromSymbol(&to, g_timeout, sizeof(int)));
    if (to) printf("{\"probe\": \"error\", \"what\": \"an mbarrier wait timed out: the numbers above are not valid\"}\n");
    return 0;
}
