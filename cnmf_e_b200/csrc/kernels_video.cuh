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

// float / double frames holding integer counts (distribute_data.m:144-147 keeps the source class, which may be 'single'):
// exact conversion to uint16, with a flag raised for any value that is not an integer in [0, 65535] (or NaN)
template <typename TIN>
__global__ void float_to_u16_kernel(const TIN* __restrict__ src, size_t n, uint16_t* __restrict__ dst, int* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = (double)src[i];
    if (!(v >= 0.0 && v <= 65535.0) || v != floor(v)) { atomicExch(bad, 1); dst[i] = 0; return; }
    dst[i] = (uint16_t)v;
}

// Frame-subsampled byte planes for the tensor-core second moments when fit_ring_model.m:84-90 keeps every kf-th frame:
// hi_k/lo_k[q][j] = bytes of Yt[q][j * kf], j < Tk; columns Tk .. Tpadk-1 are zero.  grid = (d, ceil(Tpadk/256)).
__global__ void subsample_planes_kernel(const uint16_t* __restrict__ Yt, int Tpad, int kf, int Tk, int Tpadk,
                                        uint8_t* __restrict__ hi, uint8_t* __restrict__ lo) {
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    const size_t q = blockIdx.x;
    if (j >= Tpadk) return;
    const unsigned v = j < Tk ? Yt[q * Tpad + (size_t)j * kf] : 0u;
    hi[q * Tpadk + j] = (uint8_t)(v >> 8);
    lo[q * Tpadk + j] = (uint8_t)(v & 0xffu);
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

// rows of the resident video as doubles: dst[i][0..n) = Yt[q0(i)][f0 .. f0+n), q(i) = (c0 + i / nr) * nrb + r0 + i % nr
// (patch pixels of a block in MATLAB order).  grid = (ceil(n/256), rows).
__global__ void rows_u16_to_f64_kernel(const uint16_t* __restrict__ Yt, int Tpad, int nrb, int r0, int c0, int nr, int first,
                                       int f0, int n, double* __restrict__ dst) {
    const int i = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = first + i;
    const size_t q = (size_t)(c0 + p / nr) * nrb + r0 + p % nr;
    dst[(size_t)i * n + t] = (double)Yt[q * Tpad + f0 + t];
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

// ---------------------------------------------------------------------------------------------------------------
// Tiled versions of the two projections (all frames, kf = 1).  u16 -> double without I2F: the bit pattern
// 0x43300000'0000yyyy is the double 2^52 + y, so (that - 2^52) is y exactly; then - Ymean as in the reference.
__device__ __forceinline__ double u16_to_f64(unsigned y) {
    return __hiloint2double(0x43300000, (int)y) - 4503599627370496.0;
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Centred projection, tiled: one CTA = an 8 x 8 pixel tile; the video rows of the tile are staged chunk by chunk
// (PMC_TC frames, double-buffered cp.async) into shared memory, padded so that the 64 pixels read the same frames
// without bank conflicts.  A lane owns the pixels (l, l + 32) of the tile and PMC_NB neurons: per 4 frames it does two
// 8-byte shared loads, 2 x PMC_NB 16-byte warp-uniform loads of the traces and 64 DFMA.  The 8 warps split the
// neuron groups of the tile and, when there are fewer than 8 groups, the frames of each chunk; the partial sums meet
// in shared memory at the end (fixed order).  Same contract as proj_mc_kernel (Mc dense [db][K], zeroed by the caller,
// written only inside the neuron boxes); needs T even (16-byte trace loads).
#define PMC_TC 256
#define PMC_ROWB (PMC_TC * 2 + 8)
#define PMC_NB 8
#define PMC_STAGE (64 * PMC_ROWB)
__global__ void __launch_bounds__(256, 2)
proj_mc_tile_kernel(const uint16_t* __restrict__ Yt, const double* __restrict__ Ymean, int nrb, int ncb, int T, int Tpad,
                    const double* __restrict__ Cc, int K, const int* __restrict__ bbox, double* __restrict__ Mc) {
    extern __shared__ __align__(16) unsigned char pmc_smem[];
    __shared__ int s_list[64];
    __shared__ int s_cnt;
    const int tid = threadIdx.x, lane = tid & 31, warp = warp_id_uniform();
    const int tiles_r = (nrb + 7) >> 3;
    const int r0 = (blockIdx.x % tiles_r) * 8, c0 = (blockIdx.x / tiles_r) * 8;
    const int r1 = min(nrb - 1, r0 + 7), c1 = min(ncb - 1, c0 + 7);
    if (warp == 0) {
        int cnt = 0;
        for (int kk = 0; kk < K; kk += 32) {
            const int k = kk + lane;
            bool hit = false;
            if (k < K) {
                const int* b = bbox + 4 * k;
                hit = (c1 >= b[2] && c0 <= b[3] && r1 >= b[0] && r0 <= b[1]);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int pos = cnt + __popc(m & ((1u << lane) - 1));
                if (pos < 64) s_list[pos] = k;
            }
            cnt += __popc(m);
        }
        if (lane == 0) s_cnt = min(cnt, 64);
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt == 0) return;
    const int G = (cnt + PMC_NB - 1) / PMC_NB;
    const int TS = G == 1 ? 8 : (G == 2 ? 4 : (G <= 4 ? 2 : 1));
    const int Gp = 8 / TS;
    const int grp = warp % Gp, slice = warp / Gp;
    const bool active = grp < G;
    // the lane's two pixels
    int pq[2]; double ym[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        const int r = min(r0 + (i & 7), nrb - 1), c = min(c0 + (i >> 3), ncb - 1);
        pq[h] = c * nrb + r;
        ym[h] = Ymean[pq[h]];
    }
    const double* crow[PMC_NB];
#pragma unroll
    for (int j = 0; j < PMC_NB; ++j) crow[j] = Cc + (size_t)s_list[min(grp * PMC_NB + j, cnt - 1)] * T;
    double acc[2][PMC_NB];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < PMC_NB; ++j) acc[h][j] = 0.0;
    // staging: thread -> (row = pass * 4 + tid / 64, 8-byte column = tid % 64)
    const int srow = tid >> 6, scol = tid & 63;
    auto stage = [&](int ch, int buf) {
        const int t0 = ch * PMC_TC;
        if (t0 + scol * 4 < Tpad) {
            unsigned char* dst = pmc_smem + buf * PMC_STAGE + scol * 8;
#pragma unroll 4
            for (int ps = 0; ps < 16; ++ps) {
                const int i = ps * 4 + srow;
                const int r = min(r0 + (i & 7), nrb - 1), c = min(c0 + (i >> 3), ncb - 1);
                cp_async8(dst + i * PMC_ROWB, Yt + (size_t)(c * nrb + r) * Tpad + t0 + scol * 4);
            }
        }
        cp_async_commit();
    };
    const int nch = (Tpad + PMC_TC - 1) / PMC_TC;
    stage(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) { stage(ch + 1, (ch + 1) & 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        if (active) {
            const int t0 = ch * PMC_TC;
            const int nfr = min(PMC_TC, Tpad - t0);
            const int ta = slice * nfr / TS, tb = (slice + 1) * nfr / TS;
            const unsigned char* base = pmc_smem + (ch & 1) * PMC_STAGE;
            const unsigned char* rowA = base + lane * PMC_ROWB;
            const unsigned char* rowB = base + (lane + 32) * PMC_ROWB;
            for (int t = ta; t < tb; t += 4) {
                const uint2 ya = *reinterpret_cast<const uint2*>(rowA + t * 2);
                const uint2 yb = *reinterpret_cast<const uint2*>(rowB + t * 2);
                double y[2][4];
                y[0][0] = u16_to_f64(ya.x & 0xffffu) - ym[0]; y[0][1] = u16_to_f64(ya.x >> 16) - ym[0];
                y[0][2] = u16_to_f64(ya.y & 0xffffu) - ym[0]; y[0][3] = u16_to_f64(ya.y >> 16) - ym[0];
                y[1][0] = u16_to_f64(yb.x & 0xffffu) - ym[1]; y[1][1] = u16_to_f64(yb.x >> 16) - ym[1];
                y[1][2] = u16_to_f64(yb.y & 0xffffu) - ym[1]; y[1][3] = u16_to_f64(yb.y >> 16) - ym[1];
                const int tg = t0 + t;
                if (tg + 4 <= T) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        double2 cv[PMC_NB];
#pragma unroll
                        for (int j = 0; j < PMC_NB; ++j) cv[j] = __ldg(reinterpret_cast<const double2*>(crow[j] + tg + 2 * u));
#pragma unroll
                        for (int j = 0; j < PMC_NB; ++j) {
                            acc[0][j] = fma(y[0][2 * u], cv[j].x, acc[0][j]);
                            acc[1][j] = fma(y[1][2 * u], cv[j].x, acc[1][j]);
                            acc[0][j] = fma(y[0][2 * u + 1], cv[j].y, acc[0][j]);
                            acc[1][j] = fma(y[1][2 * u + 1], cv[j].y, acc[1][j]);
                        }
                    }
                } else {
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        if (tg + f < T) {
#pragma unroll
                            for (int j = 0; j < PMC_NB; ++j) {
                                const double cvv = __ldg(crow[j] + tg + f);
                                acc[0][j] = fma(y[0][f], cvv, acc[0][j]);
                                acc[1][j] = fma(y[1][f], cvv, acc[1][j]);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    // partial sums -> shared memory [warp][pixel][PMC_NB]; summed over the frame slices in ascending order
    double* red = reinterpret_cast<double*>(pmc_smem);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < PMC_NB; ++j) red[((size_t)warp * 64 + lane + 32 * h) * PMC_NB + j] = acc[h][j];
    __syncthreads();
    const int px = tid & 63;
    const int pr = r0 + (px & 7), pc = c0 + (px >> 3);
    if (pr >= nrb || pc >= ncb) return;
    const size_t q = (size_t)pc * nrb + pr;
    for (int g = tid >> 6; g < G; g += 4) {
        for (int j = 0; j < PMC_NB; ++j) {
            const int li = g * PMC_NB + j;
            if (li >= cnt) break;
            const int k = s_list[li];
            const int* b = bbox + 4 * k;
            if (pc < b[2] || pc > b[3] || pr < b[0] || pr > b[1]) continue;
            double v = 0.0;
            for (int sl = 0; sl < TS; ++sl) v += red[((size_t)(sl * Gp + g) * 64 + px) * PMC_NB + j];
            Mc[q * K + k] = v;
        }
    }
}

// Temporal projection, list version: one CTA = (neuron k, PBT_FRAMES frames).  The non-zero pixels of B(:,k) inside
// the neuron box are compacted (in box order) into shared memory, then every thread streams its 8 frames of those
// pixel rows with four 16-byte loads in flight.  Same summation order as proj_bt_kernel.
#define PBT_FRAMES 2048
#define PBT_SEG 2048
__global__ void __launch_bounds__(256)
proj_bt_list_kernel(const uint16_t* __restrict__ Yt, const double* __restrict__ Ymean, int nrb, int T, int Tpad,
                    const double* __restrict__ B, int K, const int* __restrict__ bbox, double* __restrict__ U) {
    __shared__ int s_q[PBT_SEG];
    __shared__ double s_w[PBT_SEG];
    __shared__ double s_m[PBT_SEG];
    __shared__ int s_wcnt[8];
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.y * PBT_FRAMES + tid * 8;
    const int* b = bbox + 4 * k;
    const int br0 = b[0], br1 = b[1], bc0 = b[2], bc1 = b[3];
    const int hgt = br1 - br0 + 1, npix = (br1 >= br0 && bc1 >= bc0) ? hgt * (bc1 - bc0 + 1) : 0;
    double acc[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) acc[f] = 0.0;
    const bool live = t < Tpad;
    for (int s0 = 0; s0 < npix; s0 += PBT_SEG) {
        const int seg = min(PBT_SEG, npix - s0);
        __syncthreads();
        int n = 0;
        for (int i0 = 0; i0 < seg; i0 += 256) {
            const int i = i0 + tid;
            double w = 0.0; int q = 0;
            if (i < seg) {
                const int x = s0 + i;
                q = (bc0 + x / hgt) * nrb + br0 + x % hgt;
                w = B[(size_t)q * K + k];
            }
            const unsigned m = __ballot_sync(0xffffffffu, w != 0.0);
            if (lane == 0) s_wcnt[warp] = __popc(m);
            __syncthreads();
            int pre = 0, tot = 0;
#pragma unroll
            for (int w2 = 0; w2 < 8; ++w2) { const int cw = s_wcnt[w2]; if (w2 < warp) pre += cw; tot += cw; }
            if (w != 0.0) {
                const int pos = n + pre + __popc(m & ((1u << lane) - 1));
                s_q[pos] = q; s_w[pos] = w; s_m[pos] = Ymean[q];
            }
            n += tot;
            __syncthreads();
        }
        if (live) {
            int e = 0;
            for (; e + 4 <= n; e += 4) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(Yt + (size_t)s_q[e + u] * Tpad + t));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double w = s_w[e + u], ym = s_m[e + u];
                    const unsigned ws[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        acc[2 * h] = fma(w, u16_to_f64(ws[h] & 0xffffu) - ym, acc[2 * h]);
                        acc[2 * h + 1] = fma(w, u16_to_f64(ws[h] >> 16) - ym, acc[2 * h + 1]);
                    }
                }
            }
            for (; e < n; ++e) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(Yt + (size_t)s_q[e] * Tpad + t));
                const double w = s_w[e], ym = s_m[e];
                const unsigned ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    acc[2 * h] = fma(w, u16_to_f64(ws[h] & 0xffffu) - ym, acc[2 * h]);
                    acc[2 * h + 1] = fma(w, u16_to_f64(ws[h] >> 16) - ym, acc[2 * h + 1]);
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int f = 0; f < 8; ++f) if (t + f < T) U[(size_t)k * T + t + f] = acc[f];
    }
}

// D[q][k] = Mc[q][kmap[k]] inside neuron k's box (block coordinates r0 r1 c0 c1), untouched (zero) elsewhere: the centred
// projections of the spatial update taken from the ones the background update of the same iteration computed
__global__ void proj_from_cache_kernel(const double* __restrict__ Mc, int Kb, const int* __restrict__ kmap, const int* __restrict__ bbox,
                                       int nrb, int Ks, double* __restrict__ D) {
    const size_t q = blockIdx.x;
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= Ks) return;
    const int r = (int)(q % nrb), c = (int)(q / nrb);
    const int* b = bbox + 4 * k;
    if (r < b[0] || r > b[1] || c < b[2] || c > b[3]) return;
    D[q * Ks + k] = Mc[q * Kb + kmap[k]];
}

// host-side dispatch: the tiled kernels cover the all-frames case; frame-subsampled fits (kf > 1) and odd T use the
// warp-per-pixel-group kernels above
inline void launch_proj_mc(cudaStream_t st, const uint16_t* Yt, const double* Ymean, int nrb, int ncb, int T, int Tpad,
                           int kf, const double* Cc, int K, const int* bbox, double* Mc) {
    if (kf == 1 && (T & 1) == 0) {
        cudaFuncSetAttribute(proj_mc_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PMC_STAGE);
        const unsigned tiles = (unsigned)(((nrb + 7) / 8) * ((ncb + 7) / 8));
        LAUNCH(proj_mc_tile_kernel, tiles, 256, 2 * PMC_STAGE, st, Yt, Ymean, nrb, ncb, T, Tpad, Cc, K, bbox, Mc);
    } else {
        long long nw = (long long)((nrb + PROJ_PQ - 1) / PROJ_PQ) * ncb;
        LAUNCH(proj_mc_kernel, (unsigned)((nw + 7) / 8), 256, 0, st, Yt, Ymean, nrb, ncb, T, Tpad, kf, Cc, K, bbox, Mc);
    }
}
inline void launch_proj_bt(cudaStream_t st, const uint16_t* Yt, const double* Ymean, int nrb, int T, int Tpad,
                           const double* B, int K, const int* bbox, double* U) {
    dim3 gg(K, (T + PBT_FRAMES - 1) / PBT_FRAMES);
    LAUNCH(proj_bt_list_kernel, gg, 256, 0, st, Yt, Ymean, nrb, T, Tpad, B, K, bbox, U);
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
