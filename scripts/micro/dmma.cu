// Development micro-benchmark: fp64 tensor-core MMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16 .f64) throughput and dependent latency
// on B200, next to plain DFMA -- decides whether ring_solve_kernel's trailing updates should go through DMMA fragments.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int MODE, int ILP>
__global__ void thr(double* out, long long* cyc, double a0, double b0) {
    double c[ILP][4];
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = a0 + i + threadIdx.x;
    for (int i = 0; i < 4; ++i) b[i] = b0 + i;
    for (int j = 0; j < ILP; ++j) for (int i = 0; i < 4; ++i) c[j][i] = j + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (MODE == 0) dmma884(c[j][0], c[j][1], a[j & 7], b[j & 3]);
            if (MODE == 1) dmma1688(c[j], a, b);
            if (MODE == 2) dmma16816(c[j], a, b);
            if (MODE == 3) { c[j][0] = fma(a[0], b[0], c[j][0]); c[j][1] = fma(a[1], b[1], c[j][1]); c[j][2] = fma(a[2], b[2], c[j][2]); c[j][3] = fma(a[3], b[3], c[j][3]); }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
    for (int j = 0; j < ILP; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int ILP>
void run(const char* name, double fma_per_instr, int threads, int blocks_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    int nb = sms * blocks_per_sm;
    cudaMalloc(&out, (size_t)nb * threads * 8); cudaMalloc(&cyc, nb * 8);
    thr<MODE, ILP><<<nb, threads>>>(out, cyc, 1.0, 2.0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    thr<MODE, ILP><<<nb, threads>>>(out, cyc, 1.0, 2.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double warps = threads / 32.0 * blocks_per_sm;
    double instr_per_sm = (double)N * ILP * warps;
    double fma_total = instr_per_sm * fma_per_instr * sms;
    printf("%-28s threads=%d x%d/SM ILP=%d: %.1f cycles/warp-instr/SM, %.1f FMA/clk/SM, %.2f TFLOP/s (%.3f ms)  err=%s\n", name, threads, blocks_per_sm, ILP,
           (double)h / instr_per_sm, fma_per_instr * instr_per_sm / (double)h, 2 * fma_total / ms / 1e9, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<3, 8>("DFMA x4 (128 FMA/instr-group)", 128, 256, 2);
    run<0, 1>("DMMA m8n8k4 dependent", 256, 32, 1);
    run<0, 8>("DMMA m8n8k4", 256, 256, 1);
    run<0, 8>("DMMA m8n8k4", 256, 256, 2);
    run<0, 4>("DMMA m8n8k4", 256, 512, 2);
    run<1, 1>("DMMA m16n8k8 dependent", 1024, 32, 1);
    run<1, 8>("DMMA m16n8k8", 1024, 256, 2);
    run<2, 1>("DMMA m16n8k16 dependent", 2048, 32, 1);
    run<2, 8>("DMMA m16n8k16", 2048, 256, 1);
    run<2, 8>("DMMA m16n8k16", 2048, 256, 2);
    return 0;
}
