// membench.cu — read-only streaming bandwidth of the C2 layout (9 SoA f64 columns x 10M rows) under the
// access patterns the Gram kernels use, without any math.  Answers: what read bandwidth is reachable,
// and which pattern / how many loads in flight reach it.
//   P0: plain grid-stride 16-byte reads over one big buffer (ideal streaming read)
//   P1: "DMMA fragment": lane (fb=lane>>2, q=lane&3) reads 16 B at col[fb][row0 + 8u + 2q]  (+ y)
//   P2: "row per lane":  lane reads col[j][row0 + lane + 32u], j = 0..8
// groups of 1000 rows, one warp per group (grid-stride), like the real kernels.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
constexpr int NCOL = 9;
struct Cols { const double *c[NCOL]; };

__global__ void p0(const double2 *buf, size_t n2, double *out) {
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = buf[i]; s += v.x + v.y;
    }
    if (s == 123.456) out[0] = s;
}
template <int U>
__global__ void p1(Cols cols, int ngroups, int rows, double *out) {
    const int lane = threadIdx.x & 31, fb = lane >> 2, q = lane & 3;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double s = 0;
    const double *xc = cols.c[fb], *yc = cols.c[8];
    for (int g = wg; g < ngroups; g += nw) {
        const size_t base = (size_t)g * rows + 2 * q;
        for (int off = 0; off + 8 * U <= rows; off += 8 * U) {
            double2 xv[U], yv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { xv[u] = *(const double2 *)(xc + base + off + 8 * u); yv[u] = *(const double2 *)(yc + base + off + 8 * u); }
#pragma unroll
            for (int u = 0; u < U; ++u) s += xv[u].x * yv[u].x + xv[u].y * yv[u].y;
        }
    }
    if (s == 123.456) out[0] = s;
}
template <int U>
__global__ void p2(Cols cols, int ngroups, int rows, double *out) {
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double s = 0;
    for (int g = wg; g < ngroups; g += nw) {
        const size_t base = (size_t)g * rows + lane;
        for (int off = 0; off + 32 * U <= rows; off += 32 * U) {
            double v[U][NCOL];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < NCOL; ++j) v[u][j] = cols.c[j][base + off + 32 * u];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < NCOL; ++j) s += v[u][j];
        }
    }
    if (s == 123.456) out[0] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
    const int G = 10000, R = 1000; const size_t N = (size_t)G * R;
    double *buf, *out; cudaMalloc(&buf, N * NCOL * 8); cudaMalloc(&out, 8);
    cudaMemset(buf, 0, N * NCOL * 8);
    Cols cols; for (int j = 0; j < NCOL; ++j) cols.c[j] = buf + j * N;
    const double gb = N * NCOL * 8 / 1e9; const int S = 148;
    printf("pattern            blocks/SM thr  U   ms     GB/s\n");
    for (int bps : {2, 4, 8}) for (int thr : {256, 512}) {
        float ms = timeit([&] { p0<<<S * bps, thr>>>((const double2 *)buf, N * NCOL / 2, out); });
        printf("P0 plain           %3d %5d  -  %6.3f %7.0f\n", bps, thr, ms, gb / ms * 1e3);
    }
#define RUN1(U) for (int bps : {1, 2, 3, 4, 6, 8}) { float ms = timeit([&] { p1<U><<<S * bps, 256>>>(cols, G, R, out); }); \
        printf("P1 dmma-fragment   %3d %5d %2d  %6.3f %7.0f\n", bps, 256, U, ms, gb / ms * 1e3); }
    RUN1(1) RUN1(2) RUN1(4) RUN1(5)
#define RUN2(U) for (int bps : {1, 2, 3, 4, 6}) { float ms = timeit([&] { p2<U><<<S * bps, 256>>>(cols, G, R, out); }); \
        printf("P2 row-per-lane    %3d %5d %2d  %6.3f %7.0f\n", bps, 256, U, ms, gb / ms * 1e3); }
    RUN2(1) RUN2(2) RUN2(4)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
