// kernels_video.cuh -- kernels that stream the resident video (pixel-major u16 rows Yt[q][Tpad]).
// The video is only ever read as integers; every continuous quantity is fp64 (DESIGN.md "exact-from-integers").
#pragma once
#include "common.cuh"

namespace cnmfe {

// Frame-major chunk (nf frames of d pixels, q contiguous) -> pixel-major rows (t contiguous) + byte planes.
template <typename TIN>
__global__ void transpose_chunk_kernel(const TIN* __restrict__ src, int d, int nf, int t0, int Tpad,
                                       uint16_t* __restrict__ Yt, uint8_t* __restrict__ hi,
                                       uint8_t* __restrict__ lo) {
    __shared__ uint16_t tile[32][33];
    int q0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int f = f0 + j, q = q0 + threadIdx.x;
        tile[j][threadIdx.x] = (f < nf && q < d) ? (uint16_t)src[(size_t)f * d + q] : (uint16_t)0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int q = q0 + j, f = f0 + threadIdx.x;
        if (q < d && f < nf) {
            uint16_t v = tile[threadIdx.x][j];
            size_t o = (size_t)q * Tpad + t0 + f;
            Yt[o] = v;
            if (hi) { hi[o] = (uint8_t)(v >> 8); lo[o] = (uint8_t)(v & 0xff); }
        }
    }
}

// Ysum[q] = sum_t Yt[q][t]  (exact: integers).  One warp per pixel.
__global__ void row_sum_kernel(const uint16_t* __restrict__ Yt, int d, int T, int Tpad, double* __restrict__ Ysum) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= d) return;
    const uint16_t* row = Yt + (size_t)warp * Tpad;
    unsigned long long s = 0;
    for (int t = lane; t < T; t += 32) s += row[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) Ysum[warp] = (double)s;
}

// S1[q] = sum over selected frames (t = 0, kf, 2kf, ...)
__global__ void row_sum_strided_kernel(const uint16_t* __restrict__ Yt, int d, int T, int Tpad, int kf,
                                       double* __restrict__ S1) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= d) return;
    const uint16_t* row = Yt + (size_t)warp * Tpad;
    unsigned long long s = 0;
    for (int t = lane * kf; t < T; t += 32 * kf) s += row[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) S1[warp] = (double)s;
}

// Gather rows of a [K][T] matrix: dst[i] = src[ids[i]]; also row means and centred copy.
__global__ void gather_center_rows_kernel(const double* __restrict__ src, const int* __restrict__ ids, int n, int T,
                                          double* __restrict__ dst_centered, double* __restrict__ mean_out) {
    __shared__ double red[32];
    int i = blockIdx.x;
    if (i >= n) return;
    const double* s = src + (size_t)ids[i] * T;
    double a = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) a += s[t];
    double m = block_sum(a, red) / (double)T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) dst_centered[(size_t)i * T + t] = s[t] - m;
    if (threadIdx.x == 0 && mean_out) mean_out[i] = m;
}

// G[i][j] = sum_{t in sel} X[i][t] * Z[j][t];  rowsumX[i] = sum_sel X[i][t] (optional).  Small dense (n,m <= ~1e3).
__global__ void small_gram_kernel(const double* __restrict__ X, int n, const double* __restrict__ Z, int m, int T,
                                  int kf, double* __restrict__ G, double* __restrict__ rowsumX) {
    __shared__ double red[32];
    int i = blockIdx.x, j = blockIdx.y;
    const double* x = X + (size_t)i * T;
    const double* z = Z + (size_t)j * T;
    double a = 0.0, r = 0.0;
    for (int t = threadIdx.x * kf; t < T; t += blockDim.x * kf) { a = fma(x[t], z[t], a); r += x[t]; }
    a = block_sum(a, red);
    if (threadIdx.x == 0) G[(size_t)i * m + j] = a;
    if (rowsumX && j == 0) {
        r = block_sum(r, red);
        if (threadIdx.x == 0) rowsumX[i] = r;
    }
}

// Centred projection on a spatially local pattern:
//   Mc[q][k] = sum_{t in sel} (Y[q,t] - Ymean[q]) * Cc[k][t]   for every neuron k whose bbox contains pixel q.
// bbox: [K][4] = (r0, r1, c0, c1) inclusive, block coordinates.  One warp handles PQ consecutive pixels of one
// column; neurons are processed in batches of NB.  Mc is dense [db][K] and must be zeroed by the caller.
#define PROJ_PQ 4
#define PROJ_NB 4
__global__ void __launch_bounds__(256)
proj_mc_kernel(const uint16_t* __restrict__ Yt, const double* __restrict__ Ymean, int nrb, int ncb, int T, int Tpad,
               int kf, const double* __restrict__ Cc, int K, const int* __restrict__ bbox,
               double* __restrict__ Mc) {
    __shared__ int s_list[8][64];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups_per_col = (nrb + PROJ_PQ - 1) / PROJ_PQ;
    const long long wid = (long long)blockIdx.x * 8 + wib;
    if (wid >= (long long)groups_per_col * ncb) return;
    const int c = (int)(wid / groups_per_col), r0 = (int)(wid % groups_per_col) * PROJ_PQ;
    const int r1 = min(nrb - 1, r0 + PROJ_PQ - 1);
    int* list = s_list[wib];
    // neurons whose bbox meets this pixel group (at most 64; the host checks the overlap bound)
    int cnt = 0;
    for (int kk = 0; kk < K; kk += 32) {
        int k = kk + lane;
        bool hit = false;
        if (k < K) {
            const int* b = bbox + 4 * k;
            hit = (c >= b[2] && c <= b[3] && r1 >= b[0] && r0 <= b[1]);
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            int pos = cnt + __popc(m & ((1u << lane) - 1));
            if (pos < 64) list[pos] = k;
        }
        cnt += __popc(m);
    }
    if (cnt > 64) cnt = 64;
    __syncwarp();
    if (cnt == 0) return;
    double ym[PROJ_PQ];
    const uint16_t* rows[PROJ_PQ];
#pragma unroll
    for (int p = 0; p < PROJ_PQ; ++p) {
        int r = min(r0 + p, nrb - 1);
        size_t q = (size_t)c * nrb + r;
        rows[p] = Yt + q * Tpad;
        ym[p] = Ymean[q];
    }
    for (int b0 = 0; b0 < cnt; b0 += PROJ_NB) {
        const double* crow[PROJ_NB];
#pragma unroll
        for (int n = 0; n < PROJ_NB; ++n) crow[n] = Cc + (size_t)list[min(b0 + n, cnt - 1)] * T;
        double acc[PROJ_PQ][PROJ_NB];
#pragma unroll
        for (int p = 0; p < PROJ_PQ; ++p)
#pragma unroll
            for (int n = 0; n < PROJ_NB; ++n) acc[p][n] = 0.0;
        for (int t = lane * kf; t < T; t += 32 * kf) {
            double y[PROJ_PQ], cv[PROJ_NB];
#pragma unroll
            for (int p = 0; p < PROJ_PQ; ++p) y[p] = (double)rows[p][t] - ym[p];
#pragma unroll
            for (int n = 0; n < PROJ_NB; ++n) cv[n] = crow[n][t];
#pragma unroll
            for (int p = 0; p < PROJ_PQ; ++p)
#pragma unroll
                for (int n = 0; n < PROJ_NB; ++n) acc[p][n] = fma(y[p], cv[n], acc[p][n]);
        }
#pragma unroll
        for (int p = 0; p < PROJ_PQ; ++p)
#pragma unroll
            for (int n = 0; n < PROJ_NB; ++n) {
                double v = warp_sum(acc[p][n]);
                if (lane == 0 && r0 + p <= r1 && b0 + n < cnt) {
                    size_t q = (size_t)c * nrb + r0 + p;
                    Mc[q * K + list[b0 + n]] = v;
                }
            }
    }
}

// Temporal projection: U[k][t] = sum_q B[q][k] * (Y[q,t] - Ymean[q]) over the pixels of neuron k's bbox.
// grid = (K, ceil(T/(blockDim*2))); each thread owns two consecutive frames (one 32-bit load per row).
__global__ void __launch_bounds__(256)
proj_bt_kernel(const uint16_t* __restrict__ Yt, const double* __restrict__ Ymean, int nrb, int T, int Tpad,
               const double* __restrict__ B, int K, const int* __restrict__ bbox, double* __restrict__ U) {
    const int k = blockIdx.x;
    const int t = (blockIdx.y * blockDim.x + threadIdx.x) * 2;
    const int* b = bbox + 4 * k;
    if (t >= T) return;
    double a0 = 0.0, a1 = 0.0;
    for (int c = b[2]; c <= b[3]; ++c) {
        for (int r = b[0]; r <= b[1]; ++r) {
            size_t q = (size_t)c * nrb + r;
            double w = B[q * K + k];
            if (w == 0.0) continue;
            unsigned v = *reinterpret_cast<const unsigned*>(Yt + q * Tpad + t);
            double ym = Ymean[q];
            a0 = fma(w, (double)(v & 0xffffu) - ym, a0);
            a1 = fma(w, (double)(v >> 16) - ym, a1);
        }
    }
    U[(size_t)k * T + t] = a0;
    if (t + 1 < T) U[(size_t)k * T + t + 1] = a1;
}

// U[k][t] += cst[k] + sum_j M[k][j] * X[j][t]   (small dense correction)
__global__ void add_small_matmul_kernel(double* __restrict__ U, int K, int T, const double* __restrict__ cst,
                                        const double* __restrict__ M, int J, const double* __restrict__ X) {
    int k = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double a = cst ? cst[k] : 0.0;
    for (int j = 0; j < J; ++j) {
        double m = M[(size_t)k * J + j];
        if (m != 0.0) a = fma(m, X[(size_t)j * T + t], a);
    }
    U[(size_t)k * T + t] += a;
}

}  // namespace cnmfe
