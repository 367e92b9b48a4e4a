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
            acc += fabs(W[(size_t)i * dp + p]) * sumA[(size_t)c2 * g.nrb + r2];
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
            if (W[(size_t)i * dp + p] > 0.0) ++cnt;
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
    double* W;   // [nnb][dp]
    size_t db, ND;
};

#define RING_SOLVE_THREADS 128
// One CTA per active patch pixel: assemble the (n+1)x(n+1) normal equations, ridge, Cholesky, write weights.
// fit_ring_model.m:92-108:  X=[Bf(ring,:);1]; w=(X*X'+1e-5*trace(X*X')*I)\(X*y'); W(m,ring)=w(1:end-1)+1e-100
__global__ void __launch_bounds__(RING_SOLVE_THREADS) ring_solve_kernel(RingSolveArgs a) {
    extern __shared__ double smem[];
    const RingGeom& g = a.g;
    const int dp = g.nr * g.nc;
    const int p = a.active_list[blockIdx.x];
    const int tid = threadIdx.x;
    const int pr = p % g.nr + g.pr_off, pc = p / g.nr + g.pc_off;
    const size_t qm = (size_t)pc * g.nrb + pr;
    // shared layout
    const int NMAX = g.nnb + 1;
    double* G = smem;                                   // packed lower, NMAX*(NMAX+1)/2
    double* rhs = G + (size_t)NMAX * (NMAX + 1) / 2;    // NMAX
    double* ym = rhs + NMAX;                            // NMAX  (Ymean of ring pixels)
    double* s1c = ym + NMAX;                            // NMAX  (centred S1)
    int* qi = reinterpret_cast<int*>(s1c + NMAX);       // NMAX  block pixel index
    int* slot = qi + NMAX;                              // NMAX  ring slot
    int* sdr = slot + NMAX;                             // NMAX
    int* sdc = sdr + NMAX;                              // NMAX
    __shared__ int s_n;
    __shared__ double s_tr;
    if (tid == 0) {
        int n = 0;
        for (int i = 0; i < g.nnb; ++i) {
            int r2 = pr + a.off_r[i], c2 = pc + a.off_c[i];
            int fr = r2 + g.br0, fc = c2 + g.bc0;
            if (fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2) continue;
            qi[n] = c2 * g.nrb + r2; slot[n] = i; sdr[n] = a.off_r[i]; sdc[n] = a.off_c[i];
            ++n;
        }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n, n1 = n + 1;
    const double ymm = a.Ymean[qm];
    const double s1cm = a.S1[qm] - a.nsel * ymm;
    for (int i = tid; i < n; i += blockDim.x) {
        double y = a.Ymean[qi[i]];
        ym[i] = y;
        s1c[i] = a.S1[qi[i]] - a.nsel * y;
    }
    __syncthreads();
    // --- moments of the centred video: S2c(p,q) = S2 - Ybar_q S1_p - Ybar_p S1_q + nsel Ybar_p Ybar_q
    //     = S2 - nsel*Ybar_p*Ybar_q - Ybar_q*S1c_p - Ybar_p*S1c_q   with S1c = S1 - nsel*Ybar
    const int npair = n * (n + 1) / 2;
    for (int e = tid; e < npair; e += blockDim.x) {
        int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        while (i * (i + 1) / 2 > e) --i;
        int j = e - i * (i + 1) / 2;   // j <= i
        int ddr = sdr[i] - sdr[j], ddc = sdc[i] - sdc[j];   // displacement from p_j to p_i
        double s2;
        if (ddc > 0 || (ddc == 0 && ddr >= 0)) s2 = a.S2[(size_t)qi[j] * a.ND + ring_disp_id(ddr, ddc, g.rr)];
        else s2 = a.S2[(size_t)qi[i] * a.ND + ring_disp_id(-ddr, -ddc, g.rr)];
        G[e] = s2 - a.nsel * ym[i] * ym[j] - ym[j] * s1c[i] - ym[i] * s1c[j];
    }
    for (int i = tid; i < n; i += blockDim.x) {
        // rhs_i = Cov(p_i, m): displacement from m to p_i is (sdr, sdc)
        int ddr = sdr[i], ddc = sdc[i];
        double s2;
        if (ddc > 0 || (ddc == 0 && ddr >= 0)) s2 = a.S2[qm * a.ND + ring_disp_id(ddr, ddc, g.rr)];
        else s2 = a.S2[(size_t)qi[i] * a.ND + ring_disp_id(-ddr, -ddc, g.rr)];
        rhs[i] = s2 - a.nsel * ym[i] * ymm - ymm * s1c[i] - ym[i] * s1cm;
        G[(size_t)n * (n + 1) / 2 + i] = s1c[i];   // ones row: sum_sel Bf(p_i)
    }
    if (tid == 0) {
        G[(size_t)n * (n + 1) / 2 + n] = a.nsel;
        rhs[n] = s1cm;
    }
    __syncthreads();
    // --- neuron corrections (sequential over sparse entries; threads over the other index)
    const int K = a.K;
    for (int x = 0; x < n; ++x) {
        int e0 = a.a_ptr[qi[x]], e1 = a.a_ptr[qi[x] + 1];
        for (int e = e0; e < e1; ++e) {
            int k = a.a_col[e];
            double av = a.a_val[e];
            for (int y = tid; y < n; y += blockDim.x) {
                double nv = a.N[(size_t)qi[y] * K + k];
                int hi = x > y ? x : y, lo = x > y ? y : x;
                double f = (x == y) ? 2.0 : 1.0;
                G[(size_t)hi * (hi + 1) / 2 + lo] -= f * av * nv;
            }
            if (tid == 0) {
                rhs[x] -= av * a.N[qm * K + k];                       // - A[p_x,:].N[m,:]
                G[(size_t)n * (n + 1) / 2 + x] -= av * a.Csum[k];     // ones row
            }
            __syncthreads();
        }
    }
    {
        int e0 = a.a_ptr[qm], e1 = a.a_ptr[qm + 1];
        for (int e = e0; e < e1; ++e) {
            int k = a.a_col[e];
            double av = a.a_val[e];
            for (int y = tid; y < n; y += blockDim.x) rhs[y] -= av * a.N[(size_t)qi[y] * K + k];   // - N[p_y,:].A[m,:]
            if (tid == 0) rhs[n] -= av * a.Csum[k];
            __syncthreads();
        }
    }
    // --- ridge
    if (tid == 0) {
        double tr = 0.0;
        for (int i = 0; i < n1; ++i) tr += G[(size_t)i * (i + 1) / 2 + i];
        s_tr = tr * 1e-5;
    }
    __syncthreads();
    for (int i = tid; i < n1; i += blockDim.x) G[(size_t)i * (i + 1) / 2 + i] += s_tr;
    __syncthreads();
    // --- Cholesky (packed lower, right-looking)
    for (int k = 0; k < n1; ++k) {
        double dkk = sqrt(G[(size_t)k * (k + 1) / 2 + k]);
        __syncthreads();
        if (tid == 0) G[(size_t)k * (k + 1) / 2 + k] = dkk;
        for (int i = k + 1 + tid; i < n1; i += blockDim.x) G[(size_t)i * (i + 1) / 2 + k] /= dkk;
        __syncthreads();
        const int m = n1 - k - 1;
        const int ne = m * (m + 1) / 2;
        for (int e = tid; e < ne; e += blockDim.x) {
            int ia = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
            while ((ia + 1) * (ia + 2) / 2 <= e) ++ia;
            while (ia * (ia + 1) / 2 > e) --ia;
            int jb = e - ia * (ia + 1) / 2;
            int i = k + 1 + ia, j = k + 1 + jb;
            G[(size_t)i * (i + 1) / 2 + j] -= G[(size_t)i * (i + 1) / 2 + k] * G[(size_t)j * (j + 1) / 2 + k];
        }
        __syncthreads();
    }
    // --- forward / backward substitution
    for (int k = 0; k < n1; ++k) {
        if (tid == 0) rhs[k] /= G[(size_t)k * (k + 1) / 2 + k];
        __syncthreads();
        double zk = rhs[k];
        for (int i = k + 1 + tid; i < n1; i += blockDim.x) rhs[i] -= G[(size_t)i * (i + 1) / 2 + k] * zk;
        __syncthreads();
    }
    for (int k = n1 - 1; k >= 0; --k) {
        if (tid == 0) rhs[k] /= G[(size_t)k * (k + 1) / 2 + k];
        __syncthreads();
        double wk = rhs[k];
        for (int i = tid; i < k; i += blockDim.x) rhs[i] -= G[(size_t)k * (k + 1) / 2 + i] * wk;
        __syncthreads();
    }
    for (int i = tid; i < n; i += blockDim.x) a.W[(size_t)slot[i] * dp + p] = rhs[i] + 1e-100;
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
        W[(size_t)i * dp + p] = ok ? v : 0.0;
    }
}

}  // namespace cnmfe
