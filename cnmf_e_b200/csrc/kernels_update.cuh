// kernels_update.cuh -- per-pixel spatial solvers and the sparse algebra around the temporal projection.
// Reference: utilities/HALS_spatial.m:33-44, utilities/HALS_spatial_thresh.m:38-53, endoscope/nnls_spatial.m:33-109,
//            @Sources2D/update_spatial_parallel.m:157-166 (BG subtraction), update_temporal_parallel.m:144-181.
#pragma once
#include "common.cuh"
#include "kernels_ring.cuh"

namespace cnmfe {

#define SPATIAL_MAXROW 32   // max neurons whose search mask covers one pixel
#define NNLS_MAXP 32        // passive-set capacity (nnls_spatial maxN = 20; lars_spatial: maxIter = #masks <= 32)

// D[q][k] = Mc[q][k] - sum_k' Aprev[q,k'] * P2[k'][k]   (projection of the background residual on Cc), in place.
__global__ void spatial_make_D_kernel(double* __restrict__ Mc, int Ks, const int* __restrict__ ap_ptr,
                                      const int* __restrict__ ap_col, const double* __restrict__ ap_val,
                                      const double* __restrict__ P2) {
    size_t q = blockIdx.x;
    int e0 = ap_ptr[q], e1 = ap_ptr[q + 1];
    if (e0 == e1) return;
    for (int k = threadIdx.x; k < Ks; k += blockDim.x) {
        double s = 0.0;
        for (int e = e0; e < e1; ++e) s += ap_val[e] * P2[(size_t)ap_col[e] * Ks + k];
        Mc[q * Ks + k] -= s;
    }
}

// U(p,k) = Ysig(p,:) * Cc(k,:)' on the search pattern, from D:
//   U = [D(p,k) + R(p,k)] - sum_i W(p,i) * D(q_i,k),   R(p,k) = sum_k' Aprev(p,k') P2(k',k)
// One warp per patch pixel (rows of the pattern by patch pixel: ind_ptr / ind_col), lanes over ring slots.
__global__ void __launch_bounds__(256)
spatial_U_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                 const double* __restrict__ W, const double* __restrict__ D, int Ks,
                 const int* __restrict__ ind_ptr, const int* __restrict__ ind_col, const int* __restrict__ ap_ptr,
                 const int* __restrict__ ap_col, const double* __restrict__ ap_val, const double* __restrict__ P2,
                 double* __restrict__ U) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int dp = g.nr * g.nc;
    if (p >= dp) return;
    const int e0 = ind_ptr[p], e1 = ind_ptr[p + 1];
    if (e0 == e1) return;
    const int r = p % g.nr + g.pr_off, c = p / g.nr + g.pc_off;
    const size_t qp = (size_t)c * g.nrb + r;
    for (int e = e0; e < e1; ++e) {
        const int k = ind_col[e];
        double acc = 0.0;
        for (int i = lane; i < g.nnb; i += 32) {
            int r2 = r + off_r[i], c2 = c + off_c[i];
            int fr = r2 + g.br0, fc = c2 + g.bc0;
            if (fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2) continue;
            acc = fma(W[(size_t)p * g.nnb + i], D[((size_t)c2 * g.nrb + r2) * Ks + k], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            double rr_ = 0.0;
            for (int x = ap_ptr[qp]; x < ap_ptr[qp + 1]; ++x) rr_ += ap_val[x] * P2[(size_t)ap_col[x] * Ks + k];
            U[e] = (D[qp * Ks + k] + rr_) - acc;
        }
    }
}

__device__ inline bool chol_solve_small(double* M /*n*n row-major, overwritten*/, double* b, int n) {
    for (int k = 0; k < n; ++k) {
        double d = M[k * n + k];
        for (int j = 0; j < k; ++j) d -= M[k * n + j] * M[k * n + j];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        M[k * n + k] = d;
        for (int i = k + 1; i < n; ++i) {
            double s = M[i * n + k];
            for (int j = 0; j < k; ++j) s -= M[i * n + j] * M[k * n + j];
            M[i * n + k] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int j = 0; j < i; ++j) s -= M[i * n + j] * b[j];
        b[i] = s / M[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= M[j * n + i] * b[j];
        b[i] = s / M[i * n + i];
    }
    return true;
}

// method: 0 hals (3 sweeps, max(0,.)), 1 hals_thresh (3 sweeps, threshold 3*sn/sqrt(cc)), 2 nnls (maxN = 20),
// 3 lars (nnls with tol 1e-9, maxIter = #masks and the noise-constrained early exit of lars_spatial.m:105-109;
//   lars_thr[p] = the threshold the reference uses for pixel p -- INCLUDING its thresh(m) indexing quirk, :55).
// One thread per patch pixel.  a_io: initial A on the pattern (in), new A (out).  V = Cc*Cc' [Ks][Ks].
__global__ void __launch_bounds__(128)
spatial_solve_kernel(int dp, const int* __restrict__ ind_ptr, const int* __restrict__ ind_col,
                     const double* __restrict__ U, const double* __restrict__ V, int Ks,
                     const double* __restrict__ sn, int method, int maxIter, const double* __restrict__ lars_thr,
                     double* __restrict__ a_io, int* __restrict__ err) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dp) return;
    const int e0 = ind_ptr[p], n = ind_ptr[p + 1] - e0;
    if (n == 0) return;
    if (n > SPATIAL_MAXROW) { atomicExch(err, 1); return; }
    int col[SPATIAL_MAXROW];
    double a[SPATIAL_MAXROW], u[SPATIAL_MAXROW];
    for (int i = 0; i < n; ++i) { col[i] = ind_col[e0 + i]; a[i] = a_io[e0 + i]; u[i] = U[e0 + i]; }
    if (method == 0 || method == 1) {
        const double snp = (method == 1) ? sn[p] : 0.0;
        for (int it = 0; it < maxIter; ++it) {
            for (int i = 0; i < n; ++i) {
                const int k = col[i];
                const double cc = V[(size_t)k * Ks + k];
                if (cc == 0.0) continue;
                double s = 0.0;
                for (int j = 0; j < n; ++j) s += a[j] * V[(size_t)col[j] * Ks + k];
                double ak = a[i] + (u[i] - s) / cc;
                if (method == 0) ak = fmax(0.0, ak);
                else if (ak < snp * (3.0 / sqrt(cc))) ak = 0.0;
                a[i] = ak;
            }
        }
    } else {
        // nnls(CC(ind,ind), YC(ind,px), [], 1e-4, maxN)   (nnls_spatial.m:41-109)
        const double tol = (method == 3) ? 1e-9 : 1e-4;
        const int maxN = (method == 3) ? n : 20;
        const bool has_thr = (method == 3);
        const double thr = has_thr ? lars_thr[p] : 0.0;
        double s[SPATIAL_MAXROW], mu[NNLS_MAXP], M[NNLS_MAXP * NNLS_MAXP];
        int pidx[NNLS_MAXP];
        unsigned Pset = 0u;
        for (int i = 0; i < n; ++i) s[i] = 0.0;
        for (int miter = 0; miter < maxN; ++miter) {
            double lmax = -INFINITY;
            int imax = 0;
            for (int i = 0; i < n; ++i) {
                double l = u[i];
                for (int j = 0; j < n; ++j) l -= V[(size_t)col[i] * Ks + col[j]] * s[j];
                if (l > lmax) { lmax = l; imax = i; }
            }
            Pset = 0u;
            for (int i = 0; i < n; ++i) if (s[i] > 0.0) Pset |= (1u << i);
            if (lmax < tol) break;
            if (has_thr) {   // s'*A*s - 2*s'*b <= thresh  (lars_spatial.m:105-109)
                double q = 0.0;
                for (int i = 0; i < n; ++i) {
                    double as = 0.0;
                    for (int j = 0; j < n; ++j) as += V[(size_t)col[i] * Ks + col[j]] * s[j];
                    q += s[i] * as - 2.0 * s[i] * u[i];
                }
                if (q <= thr) break;
            }
            Pset |= (1u << imax);
            if (__popc(Pset) > maxN) break;
            int np = 0;
            bool have_mu = false;
            while (Pset) {
                np = 0;
                for (int i = 0; i < n; ++i) if (Pset & (1u << i)) pidx[np++] = i;
                for (int x = 0; x < np; ++x) {
                    mu[x] = u[pidx[x]];
                    for (int y = 0; y < np; ++y) M[x * np + y] = V[(size_t)col[pidx[x]] * Ks + col[pidx[y]]];
                }
                if (!chol_solve_small(M, mu, np)) {
                    // singular: regularised retry (nnls_spatial.m:92-94)
                    for (int x = 0; x < np; ++x) {
                        mu[x] = u[pidx[x]];
                        for (int y = 0; y < np; ++y)
                            M[x * np + y] = V[(size_t)col[pidx[x]] * Ks + col[pidx[y]]] + (x == y ? tol : 0.0);
                    }
                    chol_solve_small(M, mu, np);
                }
                have_mu = true;
                bool allpos = true;
                for (int x = 0; x < np; ++x) if (!(mu[x] > tol)) allpos = false;
                if (allpos) break;
                double amin = INFINITY;
                for (int x = 0; x < np; ++x)
                    if (!(mu[x] > tol)) amin = fmin(amin, s[pidx[x]] / (s[pidx[x]] - mu[x]));
                for (int x = 0; x < np; ++x) s[pidx[x]] = s[pidx[x]] + amin * (mu[x] - s[pidx[x]]);
                for (int i = 0; i < n; ++i) if (s[i] < tol) Pset &= ~(1u << i);
                have_mu = false;
            }
            if (have_mu) for (int x = 0; x < np; ++x) s[pidx[x]] = mu[x];
        }
        for (int i = 0; i < n; ++i) a[i] = s[i];
    }
    for (int i = 0; i < n; ++i) a_io[e0 + i] = a[i];
}

// ---- temporal side ---------------------------------------------------------------------------------------------
// B[q][k] = - sum_{patch pixels p with q = p + off_i} W(p,i) * A(p,k)     (the -W'A part; A(p,:) from rows by block
// pixel, non-empty only for patch pixels).  One thread per block pixel q (gather: deterministic order).
__global__ void temporal_build_negWtA_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                                             const double* __restrict__ W, const int* __restrict__ a_ptr,
                                             const int* __restrict__ a_col, const double* __restrict__ a_val,
                                             int Kt, double* __restrict__ B) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    int db = g.nrb * g.ncb;
    if (q >= db) return;
    int r = q % g.nrb, c = q / g.nrb;
    for (int i = 0; i < g.nnb; ++i) {
        int r2 = r - off_r[i], c2 = c - off_c[i];        // candidate centre pixel (block coords)
        int pr = r2 - g.pr_off, pc = c2 - g.pc_off;      // patch coords
        if (pr < 0 || pr >= g.nr || pc < 0 || pc >= g.nc) continue;
        size_t qc = (size_t)c2 * g.nrb + r2;
        int e0 = a_ptr[qc], e1 = a_ptr[qc + 1];
        if (e0 == e1) continue;
        double w = W[((size_t)pc * g.nr + pr) * g.nnb + i];
        for (int e = e0; e < e1; ++e) B[(size_t)q * Kt + a_col[e]] -= w * a_val[e];
    }
}

// AWA[k][k'] = sum_q (-B[q][k]) * Aprev[q,k']  with Aprev by column (CSC over local prev neurons, rows = block pixels)
__global__ void temporal_AWA_kernel(const double* __restrict__ B, int Kt, const int* __restrict__ pc_ptr,
                                    const int* __restrict__ pc_row, const double* __restrict__ pc_val, int Kp,
                                    double* __restrict__ AWA) {
    int k = blockIdx.x, kp = blockIdx.y * blockDim.x + threadIdx.x;
    if (kp >= Kp) return;
    double s = 0.0;
    for (int e = pc_ptr[kp]; e < pc_ptr[kp + 1]; ++e) s -= B[(size_t)pc_row[e] * Kt + k] * pc_val[e];
    AWA[(size_t)k * Kp + kp] = s;
}

// B[q][k] += A(q,k) on patch rows;  cst[k] = sum_p A(p,k) (Ymean_p - b0_p);  aa via V later.
__global__ void temporal_add_A_kernel(RingGeom g, const int* __restrict__ c_ptr, const int* __restrict__ c_row,
                                      const double* __restrict__ c_val, int Kt, const double* __restrict__ Ymean,
                                      const double* __restrict__ b0, double* __restrict__ B,
                                      double* __restrict__ cst) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kt) return;
    double s = 0.0;
    for (int e = c_ptr[k]; e < c_ptr[k + 1]; ++e) {
        int q = c_row[e];
        int r = q % g.nrb - g.pr_off, c = q / g.nrb - g.pc_off;
        B[(size_t)q * Kt + k] += c_val[e];
        s += c_val[e] * (Ymean[q] - b0[(size_t)c * g.nr + r]);
    }
    cst[k] = s;
}

// V = A'A from sorted CSC columns (deterministic sparse dot).  grid (Kt, ceil(Kt/128))
__global__ void temporal_V_kernel(const int* __restrict__ c_ptr, const int* __restrict__ c_row,
                                  const double* __restrict__ c_val, int Kt, double* __restrict__ V) {
    int k = blockIdx.x, j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= Kt) return;
    int a = c_ptr[k], a1 = c_ptr[k + 1], b = c_ptr[j], b1 = c_ptr[j + 1];
    double s = 0.0;
    if (a < a1 && b < b1 && c_row[a1 - 1] >= c_row[b] && c_row[b1 - 1] >= c_row[a]) {
        while (a < a1 && b < b1) {
            int ra = c_row[a], rb = c_row[b];
            if (ra == rb) { s += c_val[a] * c_val[b]; ++a; ++b; }
            else if (ra < rb) ++a;
            else ++b;
        }
    }
    V[(size_t)k * Kt + j] = s;
}

// num[ids[k]][t] += aa[k] * Craw[k][t]; den[ids[k]] += aa[k]   (update_temporal_parallel.m:269-280)
__global__ void temporal_merge_kernel(const double* __restrict__ Craw, const double* __restrict__ V, int Kt, int T,
                                      const int* __restrict__ ids, double* __restrict__ num,
                                      double* __restrict__ den) {
    int k = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    double aa = V[(size_t)k * Kt + k];
    if (t < T) num[(size_t)ids[k] * T + t] += Craw[(size_t)k * T + t] * aa;
    if (t == 0) den[ids[k]] += aa;
}

// same with the energies aa[k] = V[k][k] given as a vector
__global__ void temporal_merge_aa_kernel(const double* __restrict__ Craw, const double* __restrict__ aa_, int T,
                                         const int* __restrict__ ids, double* __restrict__ num, double* __restrict__ den) {
    int k = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    double aa = aa_[k];
    if (t < T) num[(size_t)ids[k] * T + t] += Craw[(size_t)k * T + t] * aa;
    if (t == 0) den[ids[k]] += aa;
}

// dst[k][t] = U[k][t] / V[k][k]  (0 when V[k][k] == 0): fast_temporal, update_temporal_parallel.m:329-335
__global__ void rows_div_diag_kernel(const double* __restrict__ U, const double* __restrict__ V, int K, int T,
                                     double* __restrict__ dst) {
    int k = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double aa = V[(size_t)k * K + k];
    dst[(size_t)k * T + t] = (aa == 0.0) ? 0.0 : U[(size_t)k * T + t] / aa;
}

__global__ void temporal_divide_kernel(double* __restrict__ num, const double* __restrict__ den, int K, int T) {
    int k = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double d = den[k];
    if (d == 0.0) d = 1.0;
    num[(size_t)k * T + t] = num[(size_t)k * T + t] * (1.0 / d);   // bsxfun(@times, C_new, 1./aa)
}

__global__ void rows_sub_min_kernel(double* __restrict__ X, int T) {
    __shared__ double red[32];
    double* x = X + (size_t)blockIdx.x * T;
    double mn = INFINITY;
    for (int t = threadIdx.x; t < T; t += blockDim.x) mn = fmin(mn, x[t]);
    mn = -block_max(-mn, red);
    for (int t = threadIdx.x; t < T; t += blockDim.x) x[t] -= mn;
}

__global__ void gather_rows_kernel(const double* __restrict__ src, const int* __restrict__ ids, int n, int T,
                                   double* __restrict__ dst) {
    int i = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) dst[(size_t)i * T + t] = src[(size_t)ids[i] * T + t];
}

// [K][T] <-> MATLAB column-major K x T
__global__ void kt_to_colmajor_kernel(const double* __restrict__ src, int K, int T, double* __restrict__ dst) {
    int k = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) dst[(size_t)t * K + k] = src[(size_t)k * T + t];
}
__global__ void colmajor_to_kt_kernel(const double* __restrict__ src, int K, int T, double* __restrict__ dst) {
    int k = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) dst[(size_t)k * T + t] = src[(size_t)t * K + k];
}

}  // namespace cnmfe

namespace cnmfe {

// Explicit background-subtracted rows (update_spatial_parallel.m:157-166, bg_ssub = 1) for a list of patch pixels:
//   Ysig(p,t) = Y(p,t) - b0(p) - sum_i W(p,i) * [ (Y(q_i,t) - Ybar(q_i)) - sum_k Aprev(q_i,k) * Cc_prev(k,t) ]
// Only the optional paths need the rows themselves (update_sn, lars_spatial); the main path works from projections.
// grid = (nrows, ceil(T / YSIG_TCHUNK)), block = 256.
#define YSIG_TCHUNK 2048
#define YSIG_MAXK 1024
__global__ void __launch_bounds__(256)
ysig_rows_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                 const double* __restrict__ W, const double* __restrict__ b0, const uint16_t* __restrict__ Yt,
                 const double* __restrict__ Ymean, int T, int Tpad, const int* __restrict__ ap_ptr,
                 const int* __restrict__ ap_col, const double* __restrict__ ap_val, int Kp,
                 const double* __restrict__ Ccp, const int* __restrict__ rows, double* __restrict__ out) {
    __shared__ double s_w[128], s_ym[128];
    __shared__ long long s_q[128];
    __shared__ int s_e0[128], s_e1[128];
    __shared__ double s_wa[YSIG_MAXK];
    __shared__ int s_list[64];
    __shared__ int s_nl;
    const int p = rows[blockIdx.x];
    const int r = p % g.nr + g.pr_off, c = p / g.nr + g.pc_off;
    const int tid = threadIdx.x;
    const size_t qp0 = (size_t)c * g.nrb + r;
    if (tid < 128) {
        double w = 0.0;
        long long q = (long long)qp0;
        if (tid < g.nnb) {
            int r2 = r + off_r[tid], c2 = c + off_c[tid];
            int fr = r2 + g.br0, fc = c2 + g.bc0;
            if (!(fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2)) {
                q = (long long)c2 * g.nrb + r2;
                w = W[(size_t)p * g.nnb + tid];
            }
        }
        s_w[tid] = w; s_q[tid] = q; s_ym[tid] = Ymean[q];
        s_e0[tid] = (w != 0.0) ? ap_ptr[q] : 0;
        s_e1[tid] = (w != 0.0) ? ap_ptr[q + 1] : 0;
    }
    __syncthreads();
    // WA(p,k) = sum_i W(p,i) * Aprev(q_i,k): one thread per neuron, slots in order (deterministic)
    for (int k = tid; k < Kp; k += blockDim.x) {
        double wa = 0.0;
        for (int i = 0; i < g.nnb; ++i)
            for (int e = s_e0[i]; e < s_e1[i]; ++e)
                if (ap_col[e] == k) wa += s_w[i] * ap_val[e];
        s_wa[k] = wa;
    }
    __syncthreads();
    if (tid == 0) {
        int nl = 0;
        for (int k = 0; k < Kp && nl < 64; ++k) if (s_wa[k] != 0.0) s_list[nl++] = k;
        s_nl = nl;
    }
    __syncthreads();
    const int n = g.nnb, nl = s_nl;
    const size_t qp = qp0;
    const double b0p = b0[p];
    const int t0 = blockIdx.y * YSIG_TCHUNK, t1 = min(T, t0 + YSIG_TCHUNK);
    for (int t = t0 + tid; t < t1; t += blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += s_w[i] * ((double)Yt[(size_t)s_q[i] * Tpad + t] - s_ym[i]);
        double corr = 0.0;
        for (int x = 0; x < nl; ++x) { int k = s_list[x]; corr += s_wa[k] * Ccp[(size_t)k * T + t]; }
        out[(size_t)blockIdx.x * T + t] = ((double)Yt[qp * Tpad + t] - b0p) - acc + corr;
    }
}

// ---- compute_RSS / reconstruct_background (Sources2D.m:1247-1510, ring model, bg_ssub = 1) from explicit rows of Ysig:
//   Bf(p,t)   = sum_i W(p,i) [Y(q_i,t) - b0_(q_i) - (Aprev Cprev)(q_i,t)]           (:1322-1325, :1475-1476)
//             = Y(p,t) - b0(p) - Ysig(p,t) + sum_i W(p,i) (meanR - b0_)(q_i),   meanR = Ybar - Aprev mean(Cprev)
//   Ybg(p,t)  = Bf(p,t) + b0_new(p) = Y(p,t) - Ysig(p,t) - cst(p)
//   resid     = (Y - A C)(p,t) - Ybg(p,t) = Ysig(p,t) - (A C)(p,t) + cst(p)
//   cst(p)    = b0(p) - b0_new(p) + sum_i W(p,i) (b0_ - meanR)(q_i)
// cst for every patch pixel.  b0blk / meanR: block vectors.  One thread per patch pixel.
__global__ void bg_cst_kernel(RingGeom g, const int* __restrict__ off_r, const int* __restrict__ off_c,
                              const double* __restrict__ W, const double* __restrict__ b0, const double* __restrict__ b0new,
                              const double* __restrict__ b0blk, const double* __restrict__ meanR, double* __restrict__ cst) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.nr * g.nc) return;
    const int r = p % g.nr + g.pr_off, c = p / g.nr + g.pc_off;
    double acc = 0.0;
    for (int i = 0; i < g.nnb; ++i) {
        const int r2 = r + off_r[i], c2 = c + off_c[i];
        const int fr = r2 + g.br0, fc = c2 + g.bc0;
        if (fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2) continue;
        const size_t q = (size_t)c2 * g.nrb + r2;
        acc += W[(size_t)p * g.nnb + i] * (b0blk[q] - meanR[q]);
    }
    cst[p] = b0[p] - b0new[p] + acc;
}
// meanR[q] = Ybar[q] - sum_k Aprev(q,k) * mean(Cprev_k)   (block pixels)
__global__ void bg_meanR_kernel(int db, const double* __restrict__ Ymean, const int* __restrict__ ap_ptr, const int* __restrict__ ap_col,
                                const double* __restrict__ ap_val, const double* __restrict__ Cpmean, double* __restrict__ meanR) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= db) return;
    double s = Ymean[q];
    for (int e = ap_ptr[q]; e < ap_ptr[q + 1]; ++e) s -= ap_val[e] * Cpmean[ap_col[e]];
    meanR[q] = s;
}
// rss[i] = sum_{t in [f0, f1)} (Ysig_i(t) - sum_k A(p_i,k) C(gid_k,t) + cst(p_i))^2, one CTA per row.  a_ptr indexed by BLOCK pixel.
__global__ void __launch_bounds__(256)
rss_rows_kernel(RingGeom g, const double* __restrict__ rowsY, const int* __restrict__ rows, int T, int f0, int f1,
                const int* __restrict__ a_ptr, const int* __restrict__ a_col, const double* __restrict__ a_val,
                const int* __restrict__ gid, const double* __restrict__ C, const double* __restrict__ cst, double* __restrict__ out) {
    __shared__ double red[32];
    const int p = rows[blockIdx.x];
    const size_t q = (size_t)(p / g.nr + g.pc_off) * g.nrb + (p % g.nr + g.pr_off);
    const int e0 = a_ptr[q], e1 = a_ptr[q + 1];
    const double cp = cst[p];
    const double* y = rowsY + (size_t)blockIdx.x * T;
    double acc = 0.0;
    for (int t = f0 + threadIdx.x; t < f1; t += blockDim.x) {
        double v = y[t] + cp;
        for (int e = e0; e < e1; ++e) v -= a_val[e] * C[(size_t)gid[a_col[e]] * T + t];
        acc = fma(v, v, acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}
// Ybg rows: out[i][t - f0] = Y(p_i,t) - Ysig_i(t) - cst(p_i)
__global__ void ybg_rows_kernel(RingGeom g, const double* __restrict__ rowsY, const int* __restrict__ rows, const uint16_t* __restrict__ Yt,
                                int T, int Tpad, int f0, int f1, const double* __restrict__ cst, double* __restrict__ out) {
    const int i = blockIdx.y, t = f0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f1) return;
    const int p = rows[i];
    const size_t q = (size_t)(p / g.nr + g.pc_off) * g.nrb + (p % g.nr + g.pr_off);
    out[(size_t)i * (f1 - f0) + (t - f0)] = (double)Yt[q * Tpad + t] - rowsY[(size_t)i * T + t] - cst[p];
}

// energy[i] = sum_t (X[i][t] - mean_t X[i])^2   (lars_spatial.m:42,50: Y centred, sum(Y.^2,2))
__global__ void rows_centered_energy_kernel(const double* __restrict__ X, int T, double* __restrict__ energy) {
    __shared__ double red[32];
    const double* x = X + (size_t)blockIdx.x * T;
    double a = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) a += x[t];
    const double m = block_sum(a, red) / (double)T;
    double e = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) { double d = x[t] - m; e += d * d; }
    e = block_sum(e, red);
    if (threadIdx.x == 0) energy[blockIdx.x] = e;
}

}  // namespace cnmfe
