// kernels_svd.cuh -- SVD background model (endoscope/fit_svd_model.m, svdsecon.m) as a matrix-free subspace iteration on
// X = (Y - Ybar) - A*(C - Cbar), streamed from the resident integer video; plus the svd branches of the BG subtraction
// (update_spatial_parallel.m:183-188, update_temporal_parallel.m:169-174).
#pragma once
#include "common.cuh"

namespace cnmfe {

// Z[q][j] -= sum_k A(q,k) * CV[k][j]     (CV = Cc * V', Kb x nb)
__global__ void svd_sub_ACV_kernel(double* __restrict__ Z, int nb, const int* __restrict__ a_ptr,
                                   const int* __restrict__ a_col, const double* __restrict__ a_val,
                                   const double* __restrict__ CV, int db) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= db) return;
    int e0 = a_ptr[q], e1 = a_ptr[q + 1];
    if (e0 == e1) return;
    for (int j = 0; j < nb; ++j) {
        double s = 0.0;
        for (int e = e0; e < e1; ++e) s += a_val[e] * CV[(size_t)a_col[e] * nb + j];
        Z[(size_t)q * nb + j] -= s;
    }
}

// partial[chunk][j][t] = sum_{q in chunk} Z[q][j] * (Y[q,t] - Ybar_q);  grid = (ceil(T/512), nchunks), block 256,
// each thread owns two consecutive frames.  Deterministic: fixed pixel order inside a chunk, chunks reduced in order.
#define SVD_MAXNB 4
__global__ void __launch_bounds__(256)
svd_colsum_partial_kernel(const uint16_t* __restrict__ Yt, const double* __restrict__ Ymean, int db, int T, int Tpad,
                          const double* __restrict__ Z, int nb, int qchunk, double* __restrict__ partial) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (t >= T) return;
    const int q0 = blockIdx.y * qchunk, q1 = min(db, q0 + qchunk);
    double a0[SVD_MAXNB], a1[SVD_MAXNB];
#pragma unroll
    for (int j = 0; j < SVD_MAXNB; ++j) { a0[j] = 0.0; a1[j] = 0.0; }
    for (int q = q0; q < q1; ++q) {
        unsigned v = *reinterpret_cast<const unsigned*>(Yt + (size_t)q * Tpad + t);
        double ym = Ymean[q];
        double y0 = (double)(v & 0xffffu) - ym, y1 = (double)(v >> 16) - ym;
#pragma unroll
        for (int j = 0; j < SVD_MAXNB; ++j)
            if (j < nb) { double z = Z[(size_t)q * nb + j]; a0[j] = fma(z, y0, a0[j]); a1[j] = fma(z, y1, a1[j]); }
    }
    for (int j = 0; j < nb; ++j) {
        double* o = partial + ((size_t)blockIdx.y * nb + j) * T;
        o[t] = a0[j];
        if (t + 1 < T) o[t + 1] = a1[j];
    }
}
__global__ void svd_colsum_reduce_kernel(const double* __restrict__ partial, int nchunks, int nb, int T,
                                         double* __restrict__ out) {
    int j = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double s = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) s += partial[((size_t)ch * nb + j) * T + t];
    out[(size_t)j * T + t] = s;
}

// AZ[k][j] = sum_q A(q,k) Z[q][j]  from the CSC columns of A (rows = block pixels)
__global__ void svd_AtZ_kernel(const int* __restrict__ c_ptr, const int* __restrict__ c_row,
                               const double* __restrict__ c_val, int K, const double* __restrict__ Z, int nb,
                               double* __restrict__ AZ) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    for (int j = 0; j < nb; ++j) {
        double s = 0.0;
        for (int e = c_ptr[k]; e < c_ptr[k + 1]; ++e) s += c_val[e] * Z[(size_t)c_row[e] * nb + j];
        AZ[(size_t)k * nb + j] = s;
    }
}

// dst[i][t] = sum_j M[i][j] * src[j][t]   (n, m <= SVD_MAXNB)
__global__ void rows_lincomb_kernel(const double* __restrict__ src, int m, const double* __restrict__ M, int n, int T,
                                    double* __restrict__ dst) {
    int i = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += M[i * m + j] * src[(size_t)j * T + t];
    dst[(size_t)i * T + t] = s;
}

// V0[j][t]: fixed, non-constant start vectors
__global__ void svd_init_V_kernel(double* __restrict__ V, int nb, int T) {
    int j = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    V[(size_t)j * T + t] = sinpi(2.0 * 3.3 * (double)(j + 1) * (double)t / (double)T + 0.37 * (double)j) +
                           0.25 * cospi(2.0 * 0.7 * (double)t / (double)T);
}

// ---- nmf background model (endoscope/fit_nmf_model.m): B = Y - A*C (not centred) factorised by alternating least squares
// murmur3 finaliser: the deterministic stand-in for nnmf's rand(n, k) start (the reference draws from MATLAB's global stream)
__host__ __device__ inline double nmf_hash_uniform(unsigned x) {
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return ((double)(x >> 8) + 0.5) / 16777216.0;
}
// rowen[q] = sum_t (Y[q,t] - sum_k A(q,k) C[k][t])^2.  One warp per block pixel.
__global__ void nmf_row_energy_kernel(const uint16_t* __restrict__ Yt, int db, int T, int Tpad, const int* __restrict__ a_ptr,
                                      const int* __restrict__ a_col, const double* __restrict__ a_val,
                                      const double* __restrict__ C, double* __restrict__ rowen) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= db) return;
    const uint16_t* row = Yt + (size_t)q * Tpad;
    const int e0 = a_ptr[q], e1 = a_ptr[q + 1];
    double s = 0.0;
    for (int t = lane; t < T; t += 32) {
        double v = (double)row[t];
        for (int e = e0; e < e1; ++e) v -= a_val[e] * C[(size_t)a_col[e] * T + t];
        s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) rowen[q] = s;
}
// b(p, j) = w[q(p)][j] * scale[j]  (patch rows of the block factor, columns re-ordered by `perm`)
__global__ void nmf_make_b_kernel(const double* __restrict__ Wf, int nb, const double* __restrict__ scale, const int* __restrict__ perm,
                                  int dp, int nr, int nrb, int pr_off, int pc_off, double* __restrict__ b) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dp) return;
    size_t q = (size_t)(p / nr + pc_off) * nrb + (p % nr + pr_off);
    for (int j = 0; j < nb; ++j) b[(size_t)j * dp + p] = Wf[q * nb + perm[j]] * scale[perm[j]];
}

// b(p, j) = sum_i u-part: b[p + j*dp] = Z[qp][i] * M[i][j]   (M = R * diag(1/s) * diag(s) = R: b = u*s = Z*R)
__global__ void svd_make_b_kernel(const double* __restrict__ Z, int nb, const double* __restrict__ M, int dp, int nr,
                                  int nrb, int pr_off, int pc_off, double* __restrict__ b) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dp) return;
    size_t q = (size_t)(p / nr + pc_off) * nrb + (p % nr + pr_off);
    for (int j = 0; j < nb; ++j) {
        double s = 0.0;
        for (int i = 0; i < nb; ++i) s += Z[q * nb + i] * M[i * nb + j];
        b[(size_t)j * dp + p] = s;
    }
}

// b0(p) = Ybar(p) - A(p,:)*Cmean - b(p,:)*mean(f,2)
__global__ void svd_b0_kernel(int dp, int nr, int nrb, int pr_off, int pc_off, const double* __restrict__ Ymean,
                              const int* __restrict__ a_ptr, const int* __restrict__ a_col,
                              const double* __restrict__ a_val, const double* __restrict__ Cmean,
                              const double* __restrict__ b, int nb, const double* __restrict__ fmean,
                              double* __restrict__ b0) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dp) return;
    size_t q = (size_t)(p / nr + pc_off) * nrb + (p % nr + pr_off);
    double s = Ymean[q];
    for (int e = a_ptr[q]; e < a_ptr[q + 1]; ++e) s -= a_val[e] * Cmean[a_col[e]];
    for (int j = 0; j < nb; ++j) s -= b[(size_t)j * dp + p] * fmean[j];
    b0[p] = s;
}

// U(p,k) = Mc(p,k) - sum_j b(p,j) * FC[j][k]   on the search pattern (svd BG subtraction projected on Cc)
__global__ void spatial_U_svd_kernel(int dp, int nr, int nrb, int pr_off, int pc_off, const double* __restrict__ Mc,
                                     int Ks, const int* __restrict__ ind_ptr, const int* __restrict__ ind_col,
                                     const double* __restrict__ b, int nb, const double* __restrict__ FC,
                                     double* __restrict__ U) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dp) return;
    size_t q = (size_t)(p / nr + pc_off) * nrb + (p % nr + pr_off);
    for (int e = ind_ptr[p]; e < ind_ptr[p + 1]; ++e) {
        int k = ind_col[e];
        double s = Mc[q * Ks + k];
        for (int j = 0; j < nb; ++j) s -= b[(size_t)j * dp + p] * FC[(size_t)j * Ks + k];
        U[e] = s;
    }
}

// B[q][k] = A(q,k) (patch rows);  cst[k] = sum_p A(p,k) (Ybar_p - b0_p);  Ab[k][j] = sum_p A(p,k) b(p,j)
__global__ void temporal_svd_setup_kernel(int nr, int nrb, int pr_off, int pc_off, int dp, const int* __restrict__ c_ptr,
                                          const int* __restrict__ c_row, const double* __restrict__ c_val, int Kt,
                                          const double* __restrict__ Ymean, const double* __restrict__ b0,
                                          const double* __restrict__ b, int nb, double* __restrict__ B,
                                          double* __restrict__ cst, double* __restrict__ negAb) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kt) return;
    double s = 0.0, ab[SVD_MAXNB] = {0, 0, 0, 0};
    for (int e = c_ptr[k]; e < c_ptr[k + 1]; ++e) {
        int q = c_row[e];
        int p = (q / nrb - pc_off) * nr + (q % nrb - pr_off);
        double a = c_val[e];
        B[(size_t)q * Kt + k] = a;
        s += a * (Ymean[q] - b0[p]);
        for (int j = 0; j < nb; ++j) ab[j] += a * b[(size_t)j * dp + p];
    }
    cst[k] = s;
    for (int j = 0; j < nb; ++j) negAb[(size_t)k * nb + j] = -ab[j];
}

}  // namespace cnmfe
