// kernels_ssub.cuh -- ring model with spatial down-sampling (bg_ssub > 1): data gather onto the ceil(block/ssub) grid,
// separable imresize operators as 1-D CSR tables, ring application on the coarse grid.
// Reference: update_background_parallel.m:70-118,220-227; update_spatial_parallel.m:167-177; update_temporal_parallel.m:154-163.
#pragma once
#include "common.cuh"
#include "kernels_ring.cuh"

namespace cnmfe {

// rows of the down-sampled video = rows of the block video at the 'nearest' source pixels (imresize(..,'nearest'))
__global__ void ssub_gather_rows_kernel(const uint16_t* __restrict__ Yt, const uint8_t* __restrict__ hi,
                                        const uint8_t* __restrict__ lo, const double* __restrict__ Ysum,
                                        const double* __restrict__ Ymean, const int* __restrict__ src, int Tpad,
                                        uint16_t* __restrict__ Yd, uint8_t* __restrict__ hid, uint8_t* __restrict__ lod,
                                        double* __restrict__ Ysumd, double* __restrict__ Ymeand) {
    const size_t i = blockIdx.x, q = (size_t)src[i];
    const uint4* s16 = reinterpret_cast<const uint4*>(Yt + q * Tpad);
    uint4* d16 = reinterpret_cast<uint4*>(Yd + i * Tpad);
    for (int x = threadIdx.x; x < Tpad / 8; x += blockDim.x) d16[x] = s16[x];
    const uint4* sh = reinterpret_cast<const uint4*>(hi + q * Tpad);
    const uint4* sl = reinterpret_cast<const uint4*>(lo + q * Tpad);
    uint4* dh = reinterpret_cast<uint4*>(hid + i * Tpad);
    uint4* dl = reinterpret_cast<uint4*>(lod + i * Tpad);
    for (int x = threadIdx.x; x < Tpad / 16; x += blockDim.x) { dh[x] = sh[x]; dl[x] = sl[x]; }
    if (threadIdx.x == 0) { Ysumd[i] = Ysum[q]; Ymeand[i] = Ymean[q]; }
}

// One dimension of imresize on a stack of K images stored as in[(c*nr + r)*K + k].
// along_r = 1: out(ro, c, k) = sum_e val[e] * in(idx[e], c, k), out has (n_out, nc_in) pixels;
// along_r = 0: out(r, co, k) = sum_e val[e] * in(r, idx[e], k), out has (nr_in, n_out) pixels.
__global__ void resize_dim_kernel(const double* __restrict__ in, int nr_in, int nc_in, int K,
                                  const int* __restrict__ ptr, const int* __restrict__ idx,
                                  const double* __restrict__ val, int n_out, int along_r, double* __restrict__ out) {
    const int nr_out = along_r ? n_out : nr_in, nc_out = along_r ? nc_in : n_out;
    const long long total = (long long)nr_out * nc_out * K;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int k = (int)(gid % K);
    const long long pix = gid / K;
    const int r = (int)(pix % nr_out), c = (int)(pix / nr_out);
    const int o = along_r ? r : c;
    double s = 0.0;
    for (int e = ptr[o]; e < ptr[o + 1]; ++e) {
        const size_t q = along_r ? ((size_t)c * nr_in + idx[e]) : ((size_t)idx[e] * nr_in + r);
        s += val[e] * in[q * K + k];
    }
    out[(size_t)pix * K + k] = s;
}

// Ring on the coarse grid: out[p][k] = sum_i W[p][i] * X[q_i][k]  (neighbours inside the coarse grid)
__global__ void ring_apply_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                                  const double* __restrict__ W, const double* __restrict__ X, int K,
                                  double* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int dp = g.nr * g.nc;
    if (gid >= (long long)dp * K) return;
    const int k = (int)(gid % K), p = (int)(gid / K);
    const int r = p % g.nr, c = p / g.nr;
    double s = 0.0;
    for (int i = 0; i < g.nnb; ++i) {
        int r2 = r + off_r[i], c2 = c + off_c[i];
        if (r2 < 0 || r2 >= g.nr || c2 < 0 || c2 >= g.nc) continue;
        s += W[(size_t)p * g.nnb + i] * X[((size_t)c2 * g.nr + r2) * K + k];
    }
    out[(size_t)p * K + k] = s;
}
// transpose: out[q][k] = sum_{p, i: q = p + off_i} W[p][i] * X[p][k]   (gather form, deterministic)
__global__ void ring_applyT_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                                   const double* __restrict__ W, const double* __restrict__ X, int K,
                                   double* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int dp = g.nr * g.nc;
    if (gid >= (long long)dp * K) return;
    const int k = (int)(gid % K), q = (int)(gid / K);
    const int r = q % g.nr, c = q / g.nr;
    double s = 0.0;
    for (int i = 0; i < g.nnb; ++i) {
        int r2 = r - off_r[i], c2 = c - off_c[i];
        if (r2 < 0 || r2 >= g.nr || c2 < 0 || c2 >= g.nc) continue;
        const size_t p = (size_t)c2 * g.nr + r2;
        s += W[p * g.nnb + i] * X[p * K + k];
    }
    out[(size_t)q * K + k] = s;
}

// U(p,k) = D(p,k) + R(p,k) - F(p,k) on the search pattern (F = up(W(down(D))) on the block grid)
__global__ void spatial_U_ssub_kernel(int dp, int nr, int nrb, int pr_off, int pc_off, const double* __restrict__ D,
                                      const double* __restrict__ F, int Ks, const int* __restrict__ ind_ptr,
                                      const int* __restrict__ ind_col, const int* __restrict__ ap_ptr,
                                      const int* __restrict__ ap_col, const double* __restrict__ ap_val,
                                      const double* __restrict__ P2, double* __restrict__ U) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dp) return;
    size_t q = (size_t)(p / nr + pc_off) * nrb + (p % nr + pr_off);
    for (int e = ind_ptr[p]; e < ind_ptr[p + 1]; ++e) {
        int k = ind_col[e];
        double rr_ = 0.0;
        for (int x = ap_ptr[q]; x < ap_ptr[q + 1]; ++x) rr_ += ap_val[x] * P2[(size_t)ap_col[x] * Ks + k];
        U[e] = (D[q * Ks + k] + rr_) - F[q * Ks + k];
    }
}

// dense image stack of the patch rows of A: Aimg[q][k] = A(q,k) (rows by block pixel; zero elsewhere)
__global__ void csr_to_dense_kernel(const int* __restrict__ a_ptr, const int* __restrict__ a_col,
                                    const double* __restrict__ a_val, int db, int K, double* __restrict__ out) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= db) return;
    for (int e = a_ptr[q]; e < a_ptr[q + 1]; ++e) out[(size_t)q * K + a_col[e]] = a_val[e];
}

__global__ void negate_kernel(double* __restrict__ x, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = -x[i];
}

// b0(p) = Ybar(p) - A(p,:)*Cmean on the patch pixels (mean residual, update_background_parallel.m:222-223)
__global__ void ssub_b0_kernel(int dp, int nr, int nrb, int pr_off, int pc_off, const double* __restrict__ Ymean,
                               const int* __restrict__ a_ptr, const int* __restrict__ a_col,
                               const double* __restrict__ a_val, const double* __restrict__ Cmean,
                               double* __restrict__ b0) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dp) return;
    size_t q = (size_t)(p / nr + pc_off) * nrb + (p % nr + pr_off);
    double s = Ymean[q];
    for (int e = a_ptr[q]; e < a_ptr[q + 1]; ++e) s -= a_val[e] * Cmean[a_col[e]];
    b0[p] = s;
}

}  // namespace cnmfe
