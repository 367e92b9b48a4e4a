// Development micro-benchmark: dependent-issue latencies of the fp64 operations the per-pixel solver and OASIS kernels
// chain (one warp, clock64 around N dependent ops), and DFMA throughput with 8 / 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void lat(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0, t1;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) x = a / x + b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // reciprocal intrinsic chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) x = __drcp_rn(x) + b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // shuffle chain (double = 2 x SHFL)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // shared-memory load chain (pointer chase through doubles)
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = 0.0;
    __syncthreads();
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int)sm[idx] + (threadIdx.x & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // sqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) x = sqrt(x * x + 1.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = x + idx;
}
__global__ void thr(double* out, long long* cyc, double a, double b) {
    double x0 = a + threadIdx.x, x1 = a * 2, x2 = a * 3, x3 = a * 4, x4 = a * 5, x5 = a * 6, x6 = a * 7, x7 = a * 8;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
        x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    double* out; long long* cyc; long long h[8];
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    lat<<<1, 32>>>(out, cyc, 1.000001, 0.999999);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("dependent latency (cycles/op): DADD %.1f  DFMA %.1f  DMUL %.1f  a/x+b %.1f  drcp+add %.1f  SHFL.f64 %.1f  LDS chase(+cvt) %.1f  sqrt(fma) %.1f\n",
           h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, h[3] / (N / 8.0), h[4] / (N / 8.0), h[5] / (double)N, h[6] / (double)N, h[7] / (N / 8.0));
    for (int threads = 128; threads <= 1024; threads *= 2) {
        thr<<<148, threads>>>(out, cyc, 1.000001, 0.999999);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA throughput, %4d threads/SM: %.1f DFMA/clk/SM\n", threads, 8.0 * N * threads / (double)h[0]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
