// common.cuh -- shared helpers for the cnmfe_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

#ifndef CNMFE_BLOCK
#define CNMFE_BLOCK 256
#endif
// CTA size of the HALS_temporal sweeps: that kernel is a dependency chain on a mostly idle GPU (~100 runnable items on 148 SMs),
// so a work item gets a bigger CTA than the throughput-bound batch kernels (measured at configs[1]: 256 threads 29.4 ms,
// 512 threads 21.7 ms, 1024 threads 26.9 ms)
#ifndef CNMFE_HALS_BLOCK
#define CNMFE_HALS_BLOCK 512
#endif

namespace cnmfe {

// Launch counter: every kernel of this library launched through LAUNCH() bumps it (bench.py reports it).
extern unsigned long long g_launch_count;

#define CNMFE_CUDA_OK(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            cnmfe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

void set_error(const char* fmt, ...);

#define LAUNCH(kernel, grid, block, smem, stream, ...)            \
    do {                                                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
        ++cnmfe::g_launch_count;                                  \
    } while (0)

// NOTE: every warp-collective helper starts with __syncwarp().  After divergent code (e.g. an `if (threadIdx.x == 0)`
// section) the lanes of a warp are not guaranteed to have re-converged; shuffles issued by a divergent warp take the
// compiler's slow path (BRA.DIV: one lane group at a time), which cost 10x in the OASIS scans.
// Warp index as a value the compiler can see is warp-uniform (broadcast from lane 0).  A branch on it keeps the shuffles
// inside the branch on the fast path; a branch on threadIdx.x >> 5 (or threadIdx.x < 32) makes the compiler wrap each of
// them in divergence guards (WARPSYNC / ENDCOLLECTIVE): 1.6x the instructions of ring_solve_kernel's panel factorisation.
__device__ __forceinline__ int warp_id_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ double warp_sum(double v) {
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in all threads. `red` = shared double[32].
__device__ __forceinline__ double block_sum(double v, double* red) {
    int lane = threadIdx.x & 31, wid = warp_id_uniform();
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    double r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (wid == 0) {
        r = warp_sum(r);
        if (lane == 0) red[0] = r;
    }
    __syncthreads();
    r = red[0];
    __syncthreads();
    return r;
}

__device__ __forceinline__ double block_max(double v, double* red) {
    int lane = threadIdx.x & 31, wid = warp_id_uniform();
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    double r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
        if (lane == 0) red[0] = r;
    }
    __syncthreads();
    r = red[0];
    __syncthreads();
    return r;
}

}  // namespace cnmfe
