// kernels_ring.cuh -- ring-model background regression (endoscope/fit_ring_model.m:92-108) from exact integer
// second moments of the resident video.
//
//   S2[q][id(D)] = sum_{t in sel} Y[q,t] * Y[q+D,t]      (exact, < 2^53)   D in the canonical half plane
//   Cov_Bf(p,q)  = S2c(p,q) - N[p,:].A[q,:] - A[p,:].N[q,:]              
// so the (nnb+1)^2 Gram of every pixel is ASSEMBLED from the banded moment table instead of being recomputed
// (the reference gathers a 121 x T matrix per pixel and forms X*X').
#pragma once
#include "common.cuh"

namespace cnmfe {

struct RingGeom {
    int nnb;        // ring neighbours
    int rr;         // max |offset| component
    int nrb, ncb;   // block dims
    int nr, nc;     // patch dims
    int pr_off, pc_off;   // patch origin inside block (0-based)
    int br0, bc0;   // block origin in the FOV (0-based)
    int d1, d2;     // FOV
};

__host__ __device__ inline int ring_num_disp(int rr) { return 2 * rr * (4 * rr + 1) + (2 * rr + 1); }
// canonical displacement id; requires (dc > 0) || (dc == 0 && dr >= 0)
__host__ __device__ inline int ring_disp_id(int dr, int dc, int rr) {
    return dc == 0 ? dr : (2 * rr + 1) + (dc - 1) * (4 * rr + 1) + (dr + 2 * rr);
}

// ---- SIMT second-moment kernel (exact u64 accumulation).  One warp: 4 consecutive pixels of a column x 4
// consecutive dr at one dc (Toeplitz register tile: 16 products from 4 + 7 loads).
// groups: [ngroups][2] = (dc, dr_start).
__global__ void __launch_bounds__(256)
ring_s2_simt_kernel(const uint16_t* __restrict__ Yt, int nrb, int ncb, int T, int Tpad, int kf, int rr,
                    const int* __restrict__ groups, int ngroups, double* __restrict__ S2, size_t ND) {
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gpc = (nrb + 3) / 4;
    const long long wid = (long long)blockIdx.x * 8 + wib;
    if (wid >= (long long)gpc * ncb) return;
    const int c = (int)(wid / gpc), r0 = (int)(wid % gpc) * 4;
    const int dc = groups[2 * blockIdx.y], dr0 = groups[2 * blockIdx.y + 1];
    const int c2 = c + dc;
    if (c2 >= ncb) return;
    const uint16_t* yr[4];
    const uint16_t* zr[7];
    bool yv[4], zv[7];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        int r = r0 + p;
        yv[p] = r < nrb;
        yr[p] = Yt + ((size_t)c * nrb + (yv[p] ? r : 0)) * Tpad;
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        int r = r0 + dr0 + j;
        zv[j] = (r >= 0 && r < nrb);
        zr[j] = Yt + ((size_t)c2 * nrb + (zv[j] ? r : 0)) * Tpad;
    }
    unsigned long long acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) acc[p][dd] = 0ull;
    if (kf == 1) {
        for (int t = lane * 2; t < Tpad; t += 64) {
            unsigned y[4], z[7];
#pragma unroll
            for (int p = 0; p < 4; ++p) y[p] = yv[p] ? *reinterpret_cast<const unsigned*>(yr[p] + t) : 0u;
#pragma unroll
            for (int j = 0; j < 7; ++j) z[j] = zv[j] ? *reinterpret_cast<const unsigned*>(zr[j] + t) : 0u;
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int dd = 0; dd < 4; ++dd) {
                    acc[p][dd] += (unsigned long long)(y[p] & 0xffffu) * (unsigned long long)(z[p + dd] & 0xffffu);
                    acc[p][dd] += (unsigned long long)(y[p] >> 16) * (unsigned long long)(z[p + dd] >> 16);
                }
        }
    } else {
        for (int t = lane * kf; t < T; t += 32 * kf) {
            unsigned y[4], z[7];
#pragma unroll
            for (int p = 0; p < 4; ++p) y[p] = yv[p] ? yr[p][t] : 0u;
#pragma unroll
            for (int j = 0; j < 7; ++j) z[j] = zv[j] ? zr[j][t] : 0u;
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int dd = 0; dd < 4; ++dd) acc[p][dd] += (unsigned long long)y[p] * (unsigned long long)z[p + dd];
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
            unsigned long long v = acc[p][dd];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            int dr = dr0 + dd;
            if (lane == 0 && yv[p] && zv[p + dd] && dr <= 2 * rr && dr >= -2 * rr && (dc > 0 || dr >= 0)) {
                size_t q = (size_t)c * nrb + r0 + p;
                S2[q * ND + ring_disp_id(dr, dc, rr)] = (double)v;
            }
        }
}

// ---- explicit path for options.thresh_outlier (fit_ring_model.m:48-70): the outlier clamp is a per-element non-linearity of Bf, so
// the moments cannot come from the integer video; Bf is materialised in fp64 ([db][T]) and its banded second moments over the
// selected frames are accumulated in fp64.  Same displacement layout as the integer kernel, so the solver is shared.
// Bf[q][t] = (Y[q,t] - Ybar_q) - sum_k A(q,k) Cc[k][t]   (block pixels; grid = (ceil(T/256), rows))
__global__ void bf_rows_kernel(const uint16_t* __restrict__ Yt, const double* __restrict__ Ymean, int T, int Tpad,
                               const int* __restrict__ a_ptr, const int* __restrict__ a_col, const double* __restrict__ a_val,
                               const double* __restrict__ Cc, int q0, double* __restrict__ Bf) {
    const size_t q = (size_t)q0 + blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double v = (double)Yt[q * Tpad + t] - Ymean[q];
    for (int e = a_ptr[q]; e < a_ptr[q + 1]; ++e) v -= a_val[e] * Cc[(size_t)a_col[e] * T + t];
    Bf[q * T + t] = v;
}
// clamp of the patch rows: rowsY[i][t] = Y(p_i,t) - Bf_old(p_i,t) (ysig_rows_kernel with b0 = 0 and the CURRENT A, C), so
// Bf_old = Y - rowsY;  where Bf > Bf_old + thr * sn(p): Bf <- Bf_old and the frame's outlier count goes up (:50-54, :63)
__global__ void bf_clamp_kernel(RingGeom g, const double* __restrict__ rowsY, const int* __restrict__ rows, const uint16_t* __restrict__ Yt,
                                int T, int Tpad, const double* __restrict__ sn, double thr, double* __restrict__ Bf, int* __restrict__ counts) {
    const int i = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int p = rows[i];
    const size_t q = (size_t)(p / g.nr + g.pc_off) * g.nrb + (p % g.nr + g.pr_off);
    const double bold = (double)Yt[q * Tpad + t] - rowsY[(size_t)i * T + t];
    if (Bf[q * T + t] > bold + thr * sn[p]) { Bf[q * T + t] = bold; atomicAdd(&counts[t], 1); }
}
// S2[q][id(D)] = sum_{t: mask[t]} Bf[q][t] * Bf[q+D][t];  same warp tiling as ring_s2_simt_kernel (4 pixels x 4 dr at one dc)
__global__ void __launch_bounds__(256)
ring_s2_f64_kernel(const double* __restrict__ Bf, int nrb, int ncb, int T, const unsigned char* __restrict__ mask, int rr,
                   const int* __restrict__ groups, int ngroups, double* __restrict__ S2, size_t ND) {
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gpc = (nrb + 3) / 4;
    const long long wid = (long long)blockIdx.x * 8 + wib;
    if (wid >= (long long)gpc * ncb) return;
    const int c = (int)(wid / gpc), r0 = (int)(wid % gpc) * 4;
    const int dc = groups[2 * blockIdx.y], dr0 = groups[2 * blockIdx.y + 1];
    const int c2 = c + dc;
    if (c2 >= ncb) return;
    const double* yr[4];
    const double* zr[7];
    bool yv[4], zv[7];
#pragma unroll
    for (int p = 0; p < 4; ++p) { const int r = r0 + p; yv[p] = r < nrb; yr[p] = Bf + ((size_t)c * nrb + (yv[p] ? r : 0)) * T; }
#pragma unroll
    for (int j = 0; j < 7; ++j) { const int r = r0 + dr0 + j; zv[j] = (r >= 0 && r < nrb); zr[j] = Bf + ((size_t)c2 * nrb + (zv[j] ? r : 0)) * T; }
    double acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) acc[p][dd] = 0.0;
    for (int t = lane; t < T; t += 32) {
        if (!mask[t]) continue;
        double y[4], z[7];
#pragma unroll
        for (int p = 0; p < 4; ++p) y[p] = yv[p] ? yr[p][t] : 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j) z[j] = zv[j] ? zr[j][t] : 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) acc[p][dd] = fma(y[p], z[p + dd], acc[p][dd]);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
            const double v = warp_sum(acc[p][dd]);
            const int dr = dr0 + dd;
            if (lane == 0 && yv[p] && zv[p + dd] && dr <= 2 * rr && dr >= -2 * rr && (dc > 0 || dr >= 0))
                S2[((size_t)c * nrb + r0 + p) * ND + ring_disp_id(dr, dc, rr)] = v;
        }
}
// S1[q] = sum_{t: mask[t]} Bf[q][t]   (one warp per block pixel)
__global__ void bf_row_sum_kernel(const double* __restrict__ Bf, int db, int T, const unsigned char* __restrict__ mask, double* __restrict__ S1) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= db) return;
    double s = 0.0;
    for (int t = lane; t < T; t += 32) if (mask[t]) s += Bf[(size_t)q * T + t];
    s = warp_sum(s);
    if (lane == 0) S1[q] = s;
}

// ind_active (fit_ring_model.m:25-29) and b0 (:44).  One thread per patch pixel.
__global__ void ring_active_b0_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                                      const double* __restrict__ W, const double* __restrict__ sumA,
                                      const double* __restrict__ Ymean, const int* __restrict__ a_ptr,
                                      const int* __restrict__ a_col, const double* __restrict__ a_val,
                                      const double* __restrict__ Cmean, int first_run,
                                      unsigned char* __restrict__ active, double* __restrict__ b0) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int dp = g.nr * g.nc;
    if (p >= dp) return;
    int pr = p % g.nr, pc = p / g.nr;
    int r = pr + g.pr_off, c = pc + g.pc_off;
    size_t q = (size_t)c * g.nrb + r;
    double acc = 0.0;
    if (!first_run) {
        for (int i = 0; i < g.nnb; ++i) {
            int r2 = r + off_r[i], c2 = c + off_c[i];
            int fr = r2 + g.br0, fc = c2 + g.bc0;
            if (fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2) continue;
            acc += fabs(W[(size_t)p * g.nnb + i]) * sumA[(size_t)c2 * g.nrb + r2];
        }
    }
    active[p] = (first_run || acc > 0.0) ? 1 : 0;
    double s = 0.0;
    for (int e = a_ptr[q]; e < a_ptr[q + 1]; ++e) s += a_val[e] * Cmean[a_col[e]];
    b0[p] = Ymean[q] - s;
}

// max over rows of #(W > 0)  (fit_ring_model.m:61 pmax)
__global__ void ring_pmax_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                                 const double* __restrict__ W, int* __restrict__ pmax) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int dp = g.nr * g.nc;
    int cnt = 0;
    if (p < dp) {
        int r = p % g.nr + g.pr_off, c = p / g.nr + g.pc_off;
        for (int i = 0; i < g.nnb; ++i) {
            int fr = r + off_r[i] + g.br0, fc = c + off_c[i] + g.bc0;
            if (fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2) continue;
            if (W[(size_t)p * g.nnb + i] > 0.0) ++cnt;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o));
    if ((threadIdx.x & 31) == 0) atomicMax(pmax, cnt);
}

// N[q][k] = Mc[q][k] - 0.5 * sum_k' A[q,k'] * Vsel[k'][k]  for pixels with a non-empty A row (in place on Mc).
__global__ void ring_make_N_kernel(double* __restrict__ Mc, int K, const int* __restrict__ a_ptr,
                                   const int* __restrict__ a_col, const double* __restrict__ a_val,
                                   const double* __restrict__ Vsel, size_t db) {
    size_t q = blockIdx.x;
    int e0 = a_ptr[q], e1 = a_ptr[q + 1];
    if (e0 == e1) return;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double s = 0.0;
        for (int e = e0; e < e1; ++e) s += a_val[e] * Vsel[(size_t)a_col[e] * K + k];
        Mc[q * K + k] -= 0.5 * s;
    }
}

struct RingSolveArgs {
    RingGeom g;
    const int* off_r; const int* off_c;
    const double* S2; const double* S1; const double* Ymean;
    double nsel;
    const int* a_ptr; const int* a_col; const double* a_val;   // A rows by block pixel (local neuron ids)
    const double* N; int K; const double* Csum;
    const unsigned char* active;
    const int* active_list; int n_active;
    const int* n_active_dev;    // device-side count of active_list (the grid covers all patch pixels), or nullptr: n_active
    double* W;   // [dp][nnb]
    size_t db, ND;
    unsigned long long* prof;   // diagnostics (CNMFE_RING_PROFILE): 8 per-phase cycle counters of thread 0, else nullptr
};

#define RING_SOLVE_THREADS 256
#define RING_KSET 8
#define RING_KALL 128
#define RING_NIDX 128        // index space of the augmented system: 0..n-1 ring pixels, n ones row, n1 = n+1 rhs/centre
#define RING_XSTRIDE 10
#define RING_XBUF (16 * RING_XSTRIDE)
#define RING_TT (16 * 17)     // one 16 x 16 diagonal tile of the unit factor, row stride 17
// One CTA per active patch pixel: assemble the (n+1)x(n+1) normal equations, ridge, factorise, write weights.
// fit_ring_model.m:92-108:  X=[Bf(ring,:);1]; w=(X*X'+1e-5*trace(X*X')*I)\(X*y'); W(m,ring)=w(1:end-1)+1e-100
//
// The lower triangle of the AUGMENTED matrix [G; rhs'] (row n1 = right-hand side, so the forward substitution falls
// out of the factorisation) lives in REGISTERS: 256 threads form a 16 x 16 grid, thread (ti,tj) owns the elements
// (i, j) = (ti + 16a, tj + 16b), b <= a < 8.  Each elimination step broadcasts column k through shared memory and
// every thread updates its own 36 registers; the step is templated on k/16 so finished register blocks are skipped.
//
// Per-index vectors (means, sums, neuron rows, the published column) are stored PERMUTED, entry i = t + 16a at
// [t * RING_XSTRIDE + a], so that the 8 entries a thread needs for its rows (t = ti) or its columns (t = tj) are
// contiguous: 16-byte shared loads, conflict-free with the stride of 10 doubles.
//
// Everything is written so that NO per-element masking is needed:
//   index n  (ones row)  carries  Ybar = -1, S1 = 0, S1c = nsel  and no pixel (raw moment = 0),
//   index n1 (rhs row)   carries  the centre pixel's values,
//   indices > n1 (padding) carry zeros,
// and then  Cov(i,j) = raw(i,j) - Ybar_j*S1_i - Ybar_i*S1c_j  gives the ring covariances, the ones row/column
// (S1c_j, nsel) and the right-hand side in one formula; the neuron correction  A_j.N_i + A_i.N_j  works the same way
// with A_n = 0, N_n = Csum.  Strictly-upper entries of the diagonal register tiles and the (n1, n1) entry collect
// finite values that nothing reads.
struct RingRegs { double g[8][8]; };

__device__ __forceinline__ int ring_perm(int i) { return (i & 15) * RING_XSTRIDE + (i >> 4); }
__device__ const double ring_zero_moment = 0.0;

__host__ __device__ inline size_t ring_solve_smem_bytes(int NMAX) {
    (void)NMAX;
    static_assert(8 * RING_TT <= 2 * RING_KSET * RING_XBUF, "the diagonal tiles alias the neuron-correction buffers");
    size_t dbl = 4 * RING_XBUF + 128 + 3 * RING_XBUF + 2 * (size_t)RING_KSET * RING_XBUF /* XA, XN; later the diagonal tiles */ +
                 RING_KSET + 16 + RING_NIDX /* qoff */ + 2 * RING_NIDX /* zs, ws */;
    size_t ints = 5 * (size_t)RING_NIDX + RING_KALL;
    return dbl * 8 + ints * 4 + 64;
}

// Right-looking LDL' elimination, TWO columns per barrier.  Thread (ti, tj) = (tid & 15, tid >> 4): a warp holds the
// column residues tj = 2w (lanes 0-15) and 2w+1 (lanes 16-31), every row residue ti once per half -- so the columns
// (k, k+1), k even, belong to ONE warp.  That warp factorises the 2-column panel with shuffles (pivot d_k, the entry
// x_{k+1,k}, column k+1 after the update by column k, pivot d_{k+1}) and publishes both UNSCALED columns x_i (zeros for
// i <= column) plus the reciprocal pivots; after the barrier every thread applies the rank-2 update
// G(i,j) -= (x1_i / d_k) * x1_j + (x2_i / d_{k+1}) * x2_j to its registers.  Because the published column has a zero at
// its own index, the registers of a finished column keep the unscaled x_i; rd[] keeps 1/d, so the unit factor
// M(i,k) = x_i / d_k is formed when it is needed.
template <int KA>
__device__ __forceinline__ void ring_ldl_block(RingRegs& R, int n1, int ti, int tj, double* xbuf, double* rd) {
    const int kend = min(16 * KA + 15, n1 - 1);
    constexpr int A0 = KA & ~1;          // first (even) register index loaded: keeps the 16-byte alignment
    const int lane = threadIdx.x & 31, warp = warp_id_uniform()   /* no divergence guards around the panel shuffles */, half = lane >> 4;
    for (int k = 16 * KA; k <= kend; k += 2) {
        const int kr = k & 15;
        double* xb = xbuf + ((k >> 1) & 1) * (2 * RING_XBUF);
        if (warp == (kr >> 1)) {
            const bool second = (k + 1 <= kend);
            double x[8];
#pragma unroll
            for (int a = KA; a < 8; ++a) x[a] = R.g[a][KA];
            const double d1 = __shfl_sync(0xffffffffu, x[KA], kr);
            const double m = __shfl_sync(0xffffffffu, x[KA], kr + 1);     // x_{k+1,k}
            const double r1 = 1.0 / d1;
#pragma unroll
            for (int a = KA; a < 8; ++a) {
                const double lo = __shfl_sync(0xffffffffu, x[a], ti);     // column-k entry of this thread's row
                const bool live = (ti + 16 * a > k);
                if (half) x[a] = live ? fma(-(lo * r1), m, x[a]) : x[a];
            }
            const double d2 = __shfl_sync(0xffffffffu, x[KA], 16 + ((kr + 1) & 15));
            const double r2 = second ? 1.0 / d2 : 0.0;
            double* dst = xb + half * RING_XBUF + ti * RING_XSTRIDE;
            const int kk = k + half;
            const bool on = (half == 0) || second;
#pragma unroll
            for (int a = KA; a < 8; ++a) dst[a] = (on && ti + 16 * a > kk) ? x[a] : 0.0;
            if (lane == 0) { xb[8] = r1; xb[9] = r2; rd[k] = r1; rd[k + 1] = r2; }   // slots 8, 9 of thread-row 0: padding
        }
        __syncthreads();
        const double2 rv = *reinterpret_cast<const double2*>(xb + 8);
        double cj1[8], cj2[8];
        {
            const double2* s1 = reinterpret_cast<const double2*>(xb + tj * RING_XSTRIDE);
            const double2* s2 = reinterpret_cast<const double2*>(xb + RING_XBUF + tj * RING_XSTRIDE);
#pragma unroll
            for (int a = A0; a < 8; a += 2) {
                const double2 v1 = s1[a >> 1], v2 = s2[a >> 1];
                cj1[a] = v1.x; cj1[a + 1] = v1.y; cj2[a] = v2.x; cj2[a + 1] = v2.y;
            }
        }
        const double2* si1 = reinterpret_cast<const double2*>(xb + ti * RING_XSTRIDE);
        const double2* si2 = reinterpret_cast<const double2*>(xb + RING_XBUF + ti * RING_XSTRIDE);
#pragma unroll
        for (int a2 = A0; a2 < 8; a2 += 2) {
            const double2 v1 = si1[a2 >> 1], v2 = si2[a2 >> 1];
            const double c1[2] = {v1.x * rv.x, v1.y * rv.x}, c2[2] = {v2.x * rv.y, v2.y * rv.y};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int a = a2 + h;
                if (a < KA) continue;
#pragma unroll
                for (int b = KA; b <= a; ++b) R.g[a][b] = fma(-c2[h], cj2[b], fma(-c1[h], cj1[b], R.g[a][b]));
            }
        }
    }
}

// back substitution M' w = y, one 16-unknown block at a time (descending): warp 0 solves the diagonal tile with
// shuffles, then every thread subtracts its register entries M(16A+ti, tj+16b) * w from the unknowns below (b < A),
// reduced over ti inside the half-warp.
template <int A>
__device__ __forceinline__ void ring_backsub_block(const RingRegs& R, int n, int ti, int tj, const double* Tt,
                                                   const double* rd, double* zs, double* ws) {
    if (16 * A > n) return;                      // uniform: no unknown in this block
    const int lane = threadIdx.x & 31, warp = warp_id_uniform();
    if (warp == 0) {
        const int l = lane & 15;
        const double* tile = Tt + A * RING_TT;
        double col[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) col[k] = tile[k * 17 + l];
        double z = zs[16 * A + l];
#pragma unroll
        for (int k = 15; k >= 1; --k) {
            const double wk = __shfl_sync(0xffffffffu, z, k);
            if (l < k) z = fma(-col[k], wk, z);
        }
        if (lane < 16) ws[16 * A + l] = z;
    }
    __syncthreads();
    if (A > 0) {
        const double wv = ws[16 * A + ti];
#pragma unroll
        for (int b = 0; b < A; ++b) {
            double pz = R.g[A][b] * wv;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) pz += __shfl_xor_sync(0xffffffffu, pz, o);
            if (ti == 0) zs[tj + 16 * b] = fma(-pz, rd[tj + 16 * b], zs[tj + 16 * b]);
        }
        __syncthreads();
    }
}

template <bool PROF>
__global__ void __launch_bounds__(RING_SOLVE_THREADS, 2) ring_solve_kernel(RingSolveArgs a) {
    extern __shared__ double smem[];
    const RingGeom& g = a.g;
    if ((int)blockIdx.x >= (a.n_active_dev ? *a.n_active_dev : a.n_active)) return;
    const int p = a.active_list[blockIdx.x];
    const int tid = threadIdx.x, ti = tid & 15, tj = tid >> 4;
    const int pr = p % g.nr + g.pr_off, pc = p / g.nr + g.pc_off;
    const size_t qm = (size_t)pc * g.nrb + pr;
    long long pt0 = PROF ? clock64() : 0;
#define RING_PROF(i) do { if (PROF && tid == 0) { long long _t = clock64(); atomicAdd(a.prof + (i), (unsigned long long)(_t - pt0)); pt0 = _t; } } while (0)
    // shared layout
    double* colk = smem;                                      // 2 x (2*RING_XBUF): double-buffered column pair
    double* rd = colk + 4 * RING_XBUF;                        // 128: reciprocal pivots 1/d_k
    double* ymp = rd + 128;                                   // permuted per-index vectors (see above)
    double* S1p = ymp + RING_XBUF;
    double* s1cp = S1p + RING_XBUF;
    double* XA = s1cp + RING_XBUF;                            // RING_KSET x RING_XBUF: A rows of the indices
    double* XN = XA + RING_KSET * RING_XBUF;                  // RING_KSET x RING_XBUF: N rows of the indices
    double* cs = XN + RING_KSET * RING_XBUF;                  // RING_KSET (+ 16 partial traces)
    double* trs = cs + RING_KSET;
    long long* qoff = reinterpret_cast<long long*>(trs + 16); // RING_NIDX: block pixel index * ND
    double* Tt = XA;                                          // 8 diagonal tiles of the unit factor: written after the factorisation,
                                                              // when the correction buffers XA / XN are dead (less shared memory
                                                              // per CTA = more L1 for the moment gathers: 39 -> ms measured)
    double* zs = reinterpret_cast<double*>(qoff + RING_NIDX); // RING_NIDX: running right-hand side of the back substitution
    double* wsol = zs + RING_NIDX;                            // RING_NIDX: solution
    int* qi = reinterpret_cast<int*>(wsol + RING_NIDX);       // RING_NIDX block pixel index
    int* slot = qi + RING_NIDX;
    int* elin = slot + RING_NIDX;                             // dc*(4rr+1)+dr: displacement ids are differences of these
    int* ap0 = elin + RING_NIDX;                              // A-row extents of the indices
    int* ap1 = ap0 + RING_NIDX;
    int* kall = ap1 + RING_NIDX;                              // RING_KALL
    __shared__ int s_n, s_nk;
    // valid ring neighbours (inside the FOV), compacted in slot order: parallel ballot scan over <= 8 warps
    {
        const int lane = tid & 31, wid = tid >> 5;
        __shared__ int s_wcnt[8];
        int dr = 0, dc = 0;
        bool ok = false;
        if (tid < g.nnb) {
            dr = a.off_r[tid]; dc = a.off_c[tid];
            int fr = pr + dr + g.br0, fc = pc + dc + g.bc0;
            ok = !(fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_wcnt[wid] = __popc(m);
        if (tid < RING_NIDX) { qi[tid] = (int)qm; elin[tid] = 0; rd[tid] = 0.0; }   // rd beyond the last column stays 0 (read by the tiles)
        __syncthreads();
        int base = 0;
        for (int w = 0; w < wid; ++w) base += s_wcnt[w];
        if (ok) {
            const int pos = base + __popc(m & ((1u << lane) - 1));
            qi[pos] = (pc + dc) * g.nrb + (pr + dr); slot[pos] = tid; elin[pos] = dc * (4 * g.rr + 1) + dr;
        }
        if (tid == 0) {
            int n = 0;
            for (int w = 0; w < 8; ++w) n += s_wcnt[w];
            s_n = n;
        }
    }
    __syncthreads();
    const int n = s_n, n1 = n + 1;
    if (tid < RING_NIDX) {
        const int i = tid;
        double y = 0.0, s1 = 0.0, s1c = 0.0;
        int p0 = 0, p1 = 0;
        long long qo = 0;
        if (i < n || i == n1) {
            const int q = qi[i];
            y = a.Ymean[q]; s1 = a.S1[q]; s1c = s1 - a.nsel * y;
            p0 = a.a_ptr[q]; p1 = a.a_ptr[q + 1];
            qo = (long long)q * (long long)a.ND;
        } else if (i == n) {
            y = -1.0; s1 = 0.0; s1c = a.nsel;
        }
        const int pp = ring_perm(i);
        ymp[pp] = y; S1p[pp] = s1; s1cp[pp] = s1c;
        ap0[i] = p0; ap1[i] = p1; qoff[i] = qo;
    }
    __syncthreads();
    // --- assemble into registers: Cov(i,j) = raw(i,j) - Ybar_j*S1_i - Ybar_i*S1c_j, raw = second moment of the pixel
    //     pair (0 where an index carries no pixel).  The moment of a pair sits at S2[base*ND + |e_i - e_j|], base = the
    //     pixel the displacement starts from (canonical half plane <=> e_i - e_j >= 0).
    RING_PROF(0);
    RingRegs R;
    {
        double ymj[8], s1cj[8];
        int ej[8];
#pragma unroll
        for (int b = 0; b < 8; b += 2) {
            const double2 v = reinterpret_cast<const double2*>(ymp + tj * RING_XSTRIDE)[b >> 1];
            const double2 w = reinterpret_cast<const double2*>(s1cp + tj * RING_XSTRIDE)[b >> 1];
            ymj[b] = v.x; ymj[b + 1] = v.y; s1cj[b] = w.x; s1cj[b + 1] = w.y;
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) ej[b] = elin[tj + 16 * b];
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            const int i = ti + 16 * aa;
            const bool vi = (i < n) || (i == n1);
            const int ei = elin[i];
            const long long qoi = qoff[i];
            const double* ptr[8];
            double raw[8];
#pragma unroll
            for (int bb = 0; bb <= aa; ++bb) {
                const int j = tj + 16 * bb;
                const int lin = ei - ej[bb];
                const long long off = (lin >= 0 ? qoff[j] : qoi) + (long long)abs(lin);
                ptr[bb] = (vi && j < n) ? a.S2 + off : &ring_zero_moment;
            }
#pragma unroll
            for (int bb = 0; bb <= aa; ++bb) raw[bb] = __ldg(ptr[bb]);
            const double ymi = ymp[ti * RING_XSTRIDE + aa], s1i = S1p[ti * RING_XSTRIDE + aa];
#pragma unroll
            for (int bb = 0; bb <= aa; ++bb) R.g[aa][bb] = fma(-ymi, s1cj[bb], fma(-ymj[bb], s1i, raw[bb]));
        }
    }
    // --- neuron corrections: Cov_Bf = Cov_Y - N_x.A_y - A_x.N_y ; sum_sel Bf(x) = S1c_x - A_x.Csum
    // distinct neurons touching the ring pixels / the centre: bitmap over local neuron ids (K <= 4096), compacted in
    // ascending id order (deterministic), handled RING_KSET at a time
    RING_PROF(1);
    __shared__ unsigned s_bits[128];
    if (tid < 128) s_bits[tid] = 0u;
    __syncthreads();
    if (tid < RING_NIDX)
        for (int e = ap0[tid]; e < ap1[tid]; ++e) { int k = a.a_col[e]; atomicOr(&s_bits[(k >> 5) & 127], 1u << (k & 31)); }
    __syncthreads();
    if (warp_id_uniform() == 0) {
        int base = 0;
        for (int w0 = 0; w0 < 128; w0 += 32) {
            const unsigned bits = s_bits[w0 + tid];
            const int cnt = __popc(bits);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
            int pos = base + incl - cnt;
            unsigned b = bits;
            while (b) { int bit = __ffs(b) - 1; b &= b - 1; if (pos < RING_KALL) kall[pos] = ((w0 + tid) << 5) + bit; ++pos; }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (tid == 0) s_nk = min(base, RING_KALL);
    }
    __syncthreads();
    const int nall = s_nk;
    RING_PROF(2);
    if (PROF && tid == 0) atomicAdd(a.prof + 7, (unsigned long long)nall);
    for (int kbase = 0; kbase < nall; kbase += RING_KSET) {
        const int* kset = kall + kbase;
        const int nk = min(RING_KSET, nall - kbase);
        __syncthreads();
        for (int x = tid; x < 2 * RING_KSET * RING_XBUF; x += blockDim.x) XA[x] = 0.0;   // XA and XN are adjacent
        __syncthreads();
        for (int x = tid; x < RING_NIDX * nk; x += blockDim.x) {
            const int y = x & (RING_NIDX - 1), z = x >> 7;
            if (y < n || y == n1) XN[z * RING_XBUF + ring_perm(y)] = a.N[(size_t)qi[y] * a.K + kset[z]];
            else if (y == n) XN[z * RING_XBUF + ring_perm(y)] = a.Csum[kset[z]];
        }
        if (tid < RING_NIDX)
            for (int e = ap0[tid]; e < ap1[tid]; ++e) {
                const int k = a.a_col[e];
                for (int z = 0; z < nk; ++z) if (kset[z] == k) XA[z * RING_XBUF + ring_perm(tid)] = a.a_val[e];
            }
        __syncthreads();
        for (int z = 0; z < nk; ++z) {
            const double* xa = XA + z * RING_XBUF;
            const double* xn = XN + z * RING_XBUF;
            double aj[8], nj[8];
#pragma unroll
            for (int b = 0; b < 8; b += 2) {
                const double2 v = reinterpret_cast<const double2*>(xa + tj * RING_XSTRIDE)[b >> 1];
                const double2 w = reinterpret_cast<const double2*>(xn + tj * RING_XSTRIDE)[b >> 1];
                aj[b] = v.x; aj[b + 1] = v.y; nj[b] = w.x; nj[b + 1] = w.y;
            }
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const double ai = xa[ti * RING_XSTRIDE + aa], ni = xn[ti * RING_XSTRIDE + aa];
#pragma unroll
                for (int bb = 0; bb <= aa; ++bb) R.g[aa][bb] = fma(-aj[bb], ni, fma(-ai, nj[bb], R.g[aa][bb]));
            }
        }
    }
    RING_PROF(3);
    // --- ridge: trace over the n1 x n1 system (diagonal owners are the threads with ti == tj), summed in a fixed order
    {
        if (ti == tj) {
            double tr = 0.0;
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) if (ti + 16 * aa < n1) tr += R.g[aa][aa];
            trs[ti] = tr;
        }
        __syncthreads();
        double tot = 0.0;
#pragma unroll
        for (int t = 0; t < 16; ++t) tot += trs[t];
        const double lam = tot * 1e-5;
        if (ti == tj) {
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) if (ti + 16 * aa < n1) R.g[aa][aa] += lam;
        }
    }
    RING_PROF(4);
    // --- LDL' of the augmented matrix
    ring_ldl_block<0>(R, n1, ti, tj, colk, rd);
    if (n1 > 16) ring_ldl_block<1>(R, n1, ti, tj, colk, rd);
    if (n1 > 32) ring_ldl_block<2>(R, n1, ti, tj, colk, rd);
    if (n1 > 48) ring_ldl_block<3>(R, n1, ti, tj, colk, rd);
    if (n1 > 64) ring_ldl_block<4>(R, n1, ti, tj, colk, rd);
    if (n1 > 80) ring_ldl_block<5>(R, n1, ti, tj, colk, rd);
    if (n1 > 96) ring_ldl_block<6>(R, n1, ti, tj, colk, rd);
    if (n1 > 112) ring_ldl_block<7>(R, n1, ti, tj, colk, rd);
    __syncthreads();
    RING_PROF(5);
    // --- row n1 of the registers is now the unscaled forward-substituted right-hand side: y_j = x_{n1,j} / d_j.
    //     Diagonal tiles of the unit factor M(i,j) = x_ij / d_j go to shared memory; M' w = y is solved block-wise.
    if (tid < RING_NIDX) zs[tid] = 0.0;
    __syncthreads();
    {
        const int a0 = n1 >> 4, t0 = n1 & 15;
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            Tt[aa * RING_TT + ti * 17 + tj] = (ti > tj) ? R.g[aa][aa] * rd[16 * aa + tj] : 0.0;
            if (aa == a0 && ti == t0) {
#pragma unroll
                for (int bb = 0; bb <= aa; ++bb) { const int j = tj + 16 * bb; if (j < n1) zs[j] = R.g[aa][bb] * rd[j]; }
            }
        }
    }
    __syncthreads();
    ring_backsub_block<7>(R, n, ti, tj, Tt, rd, zs, wsol);
    ring_backsub_block<6>(R, n, ti, tj, Tt, rd, zs, wsol);
    ring_backsub_block<5>(R, n, ti, tj, Tt, rd, zs, wsol);
    ring_backsub_block<4>(R, n, ti, tj, Tt, rd, zs, wsol);
    ring_backsub_block<3>(R, n, ti, tj, Tt, rd, zs, wsol);
    ring_backsub_block<2>(R, n, ti, tj, Tt, rd, zs, wsol);
    ring_backsub_block<1>(R, n, ti, tj, Tt, rd, zs, wsol);
    ring_backsub_block<0>(R, n, ti, tj, Tt, rd, zs, wsol);
    for (int i = tid; i < n; i += blockDim.x) a.W[(size_t)p * g.nnb + slot[i]] = wsol[i] + 1e-100;
    RING_PROF(6);
#undef RING_PROF
}

// uniform ring initialisation (initComponents_parallel.m:213-236): W[i][p] = 1/#valid neighbours
__global__ void ring_uniform_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                                    double* __restrict__ W) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int dp = g.nr * g.nc;
    if (p >= dp) return;
    int r = p % g.nr + g.pr_off, c = p / g.nr + g.pc_off;
    int cnt = 0;
    for (int i = 0; i < g.nnb; ++i) {
        int fr = r + off_r[i] + g.br0, fc = c + off_c[i] + g.bc0;
        if (!(fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2)) ++cnt;
    }
    double v = cnt > 0 ? 1.0 / (double)cnt : 0.0;
    for (int i = 0; i < g.nnb; ++i) {
        int fr = r + off_r[i] + g.br0, fc = c + off_c[i] + g.bc0;
        bool ok = !(fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2);
        W[(size_t)p * g.nnb + i] = ok ? v : 0.0;
    }
}

}  // namespace cnmfe
