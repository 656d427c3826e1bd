// microbench.cu — FP64 pipe facts on the actual B200 the build runs on:
//   DMMA.8x8x4 dependent-issue latency, DMMA throughput per SM (independent chains, many warps),
//   DFMA throughput per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CHAINS>
__global__ void k_dmma(double *out, int iters, long long *cycles) {
    double acc[CHAINS][2];
    for (int c = 0; c < CHAINS; ++c) acc[c][0] = acc[c][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) dmma(acc[c][0], acc[c][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int CHAINS>
__global__ void k_dfma(double *out, int iters, long long *cycles) {
    double acc[CHAINS];
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x;
    double a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <typename K>
void run(const char *name, K kern, int blocks, int threads, int iters, double ops_per_thread_iter, double flop_per_op) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
    kern<<<blocks, threads>>>(out, 10, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); kern<<<blocks, threads>>>(out, iters, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double warp_ops = (double)blocks * threads / 32 * iters * ops_per_thread_iter;
    printf("%-34s blocks=%4d thr=%4d  %8.3f ms  cycles/iter(blk0)=%8.1f  warp-ops/s=%.3e  TFLOP/s=%.2f  err=%s\n", name, blocks, threads, ms,
           (double)h / iters, warp_ops / (ms * 1e-3), warp_ops * flop_per_op / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int S = p.multiProcessorCount;
    run("dmma latency (1 warp, 1 chain)", k_dmma<1>, 1, 32, 20000, 1, 512);
    run("dmma 1 warp, 4 chains", k_dmma<4>, 1, 32, 20000, 4, 512);
    run("dmma 1 warp/SMSP (4 warps), 4 ch", k_dmma<4>, S, 128, 20000, 4, 512);
    run("dmma 8 warps/SM, 4 chains", k_dmma<4>, S, 256, 20000, 4, 512);
    run("dmma 32 warps/SM, 4 chains", k_dmma<4>, S * 2, 512, 5000, 4, 512);
    run("dfma latency (1 warp, 1 chain)", k_dfma<1>, 1, 32, 100000, 1, 64);
    run("dfma 1 warp, 8 chains", k_dfma<8>, 1, 32, 100000, 8, 64);
    run("dfma 32 warps/SM, 8 chains", k_dfma<8>, S * 2, 512, 20000, 8, 64);
    return 0;
}
