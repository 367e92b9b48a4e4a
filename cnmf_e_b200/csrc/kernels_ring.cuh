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
    double* W;   // [dp][nnb]
    size_t db, ND;
};

#define RING_SOLVE_THREADS 256
#define RING_KSET 16
#define RING_KALL 128
// One CTA per active patch pixel: assemble the (n+1)x(n+1) normal equations, ridge, Cholesky, write weights.
// fit_ring_model.m:92-108:  X=[Bf(ring,:);1]; w=(X*X'+1e-5*trace(X*X')*I)\(X*y'); W(m,ring)=w(1:end-1)+1e-100
//
// The lower triangle of the AUGMENTED matrix [G; rhs'] (row n1 = right-hand side, so the forward substitution falls
// out of the factorisation) lives in REGISTERS: 256 threads form a 16 x 16 grid, thread (ti,tj) owns the elements
// (i, j) = (ti + 16a, tj + 16b), b <= a < 8.  Each elimination step broadcasts column k through shared memory and
// every thread updates its own 36 registers; the step is templated on k/16 so finished register blocks are skipped.
struct RingRegs { double g[8][8]; };

// Right-looking elimination with ONE barrier per column: the owners of column k publish its UNSCALED entries x_i and the
// pivot g_kk (double-buffered), then every thread applies G(i,j) -= (x_i / g_kk) * x_j to its registers.  The column of
// the Cholesky factor is L(i,k) = x_i * (1/sqrt(g_kk)) (LAPACK dpotf2 scales by the reciprocal pivot); the registers keep
// x_i and the pivots are stored in piv[] so that the scaling is applied once, when L is written out.
template <int KA>
__device__ __forceinline__ void ring_chol_block(RingRegs& R, int n1, int ti, int tj, double* xbuf, double* piv) {
    const int kend = min(16 * KA + 15, n1 - 1);
    for (int k = 16 * KA; k <= kend; ++k) {
        const int kr = k & 15;
        double* xb = xbuf + (k & 1) * 128;
        if (tj == kr) {
#pragma unroll
            for (int a = KA; a < 8; ++a) {
                const int i = ti + 16 * a;
                if (i > k && i <= n1) xb[i] = R.g[a][KA];
            }
            if (ti == kr) { piv[k] = R.g[KA][KA]; xb[127] = 1.0 / R.g[KA][KA]; }
        }
        __syncthreads();
        const double rinv = xb[127];
        double ci[8], cj[8];
#pragma unroll
        for (int a = KA; a < 8; ++a) {
            const int i = ti + 16 * a, j = tj + 16 * a;
            ci[a] = (i > k && i <= n1) ? xb[i] * rinv : 0.0;
            cj[a] = (j > k && j < n1) ? xb[j] : 0.0;
        }
#pragma unroll
        for (int a = KA; a < 8; ++a)
#pragma unroll
            for (int b = KA; b <= a; ++b) {
                // column k itself (b == KA && tj == kr) is final: keep x_i there
                R.g[a][b] = fma(-ci[a], cj[b], R.g[a][b]);
            }
    }
}

__global__ void __launch_bounds__(RING_SOLVE_THREADS, 2) ring_solve_kernel(RingSolveArgs a) {
    extern __shared__ double smem[];
    const RingGeom& g = a.g;
    const int dp = g.nr * g.nc;
    const int p = a.active_list[blockIdx.x];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    const int pr = p % g.nr + g.pr_off, pc = p / g.nr + g.pc_off;
    const size_t qm = (size_t)pc * g.nrb + pr;
    const int NMAX = g.nnb + 1;
    // shared layout
    double* L = smem;                                         // packed lower (after the factorisation), (NMAX+1)(NMAX+2)/2
    double* colk = L + (size_t)(NMAX + 1) * (NMAX + 2) / 2;   // 256: double-buffered column, later the solution
    double* piv = colk + 256;                                 // 128: pivots g_kk, later 1/sqrt(g_kk)
    double* ym = piv + 128;                                  // NMAX+1   (index n = the centre pixel m)
    double* s1c = ym + NMAX + 1;                              // NMAX+1
    double* Ar = s1c + NMAX + 1;                              // (NMAX+1) * RING_KSET
    double* Nr = Ar + (size_t)(NMAX + 1) * RING_KSET;         // (NMAX+1) * RING_KSET
    double* cs = Nr + (size_t)(NMAX + 1) * RING_KSET;         // RING_KSET
    int* qi = reinterpret_cast<int*>(cs + RING_KSET);         // NMAX+1  block pixel index (index n = m)
    int* slot = qi + NMAX + 1;
    int* sdr = slot + NMAX + 1;                               // index n: 0
    int* sdc = sdr + NMAX + 1;
    int* kall = sdc + NMAX + 1;                               // RING_KALL
    int* ap0 = kall + RING_KALL;                              // NMAX+1  A-row extents of the ring pixels / centre
    int* ap1 = ap0 + NMAX + 1;
    int* rows_with = ap1 + NMAX + 1;                          // NMAX+1  ring pixels that carry neuron entries
    __shared__ int s_n, s_nk;
    __shared__ double s_tr;
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
        __syncthreads();
        int base = 0;
        for (int w = 0; w < wid; ++w) base += s_wcnt[w];
        if (ok) {
            const int pos = base + __popc(m & ((1u << lane) - 1));
            qi[pos] = (pc + dc) * g.nrb + (pr + dr); slot[pos] = tid; sdr[pos] = dr; sdc[pos] = dc;
        }
        if (tid == 0) {
            int n = 0;
            for (int w = 0; w < 8; ++w) n += s_wcnt[w];
            qi[n] = (int)qm; sdr[n] = 0; sdc[n] = 0;
            s_n = n;
        }
    }
    __syncthreads();
    const int n = s_n, n1 = n + 1;
    for (int i = tid; i <= n; i += blockDim.x) {
        double y = a.Ymean[qi[i]];
        ym[i] = y;
        s1c[i] = a.S1[qi[i]] - a.nsel * y;
        ap0[i] = a.a_ptr[qi[i]];
        ap1[i] = a.a_ptr[qi[i] + 1];
    }
    __syncthreads();
    // --- assemble into registers.  Rows 0..n-1: ring pixels; row n: ones; row n1: right-hand side (pixel m = index n
    //     of the qi/ym/s1c arrays).  Cov(x, y) of the centred video for "pixel indices" x, y in [0, n]:
    //     S2c = S2 - nsel*Yb_x*Yb_y - Yb_y*S1c_x - Yb_x*S1c_y,  S1c = S1 - nsel*Yb
    RingRegs R;
    // address of the raw second moment of "pixel indices" (x, y) (canonical orientation chosen branch-free)
    auto s2ptr = [&](int x, int y) -> const double* {
        int ddr = sdr[x] - sdr[y], ddc = sdc[x] - sdc[y];   // displacement from pixel y to pixel x
        bool canon = (ddc > 0) || (ddc == 0 && ddr >= 0);
        int base = canon ? qi[y] : qi[x];
        int id = canon ? ring_disp_id(ddr, ddc, g.rr) : ring_disp_id(-ddr, -ddc, g.rr);
        return a.S2 + (size_t)base * a.ND + id;
    };
#pragma unroll
    for (int aa = 0; aa < 8; ++aa) {
        const int i = ti + 16 * aa;
        const double* ptr[8];
        double raw[8];
#pragma unroll
        for (int bb = 0; bb <= aa; ++bb) {
            const int j = tj + 16 * bb;
            const bool pair = (j <= i) && (j < n) && (i < n || i == n1);
            ptr[bb] = pair ? s2ptr(i < n ? i : n, j) : a.S2;
        }
#pragma unroll
        for (int bb = 0; bb <= aa; ++bb) raw[bb] = __ldg(ptr[bb]);
#pragma unroll
        for (int bb = 0; bb <= aa; ++bb) {
            const int j = tj + 16 * bb;
            double v = 0.0;
            if (j <= i && i <= n1 && j <= n) {
                const int x = (i < n) ? i : n;   // row n1 (rhs) pairs the centre pixel (index n) with p_j
                if ((i < n || i == n1) && j < n) v = raw[bb] - a.nsel * ym[x] * ym[j] - ym[j] * s1c[x] - ym[x] * s1c[j];
                else if (i == n) v = (j < n) ? s1c[j] : a.nsel;   // ones row
                else v = s1c[n];                                    // i == n1, j == n: sum Bf(m)
            }
            R.g[aa][bb] = v;
        }
    }
    // --- neuron corrections: Cov_Bf = Cov_Y - N_x.A_y - A_x.N_y ; sum_sel Bf(x) = S1c_x - A_x.Csum
    // distinct neurons touching the ring pixels / the centre (deterministic scan order), handled RING_KSET at a time
    // distinct neurons: bitmap over local neuron ids (K <= 4096), compacted in ascending id order (deterministic)
    __shared__ unsigned s_bits[128];
    if (tid < 128) s_bits[tid] = 0u;
    __syncthreads();
    for (int y = tid; y <= n; y += blockDim.x)
        for (int e = ap0[y]; e < ap1[y]; ++e) { int k = a.a_col[e]; atomicOr(&s_bits[(k >> 5) & 127], 1u << (k & 31)); }
    __syncthreads();
    if (tid < 32) {
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
    for (int kbase = 0; kbase < nall; kbase += RING_KSET) {
        const int* kset = kall + kbase;
        __syncthreads();
        const int nk = min(RING_KSET, nall - kbase);
        for (int x = tid; x < (n + 1) * RING_KSET; x += blockDim.x) { Ar[x] = 0.0; }
        for (int x = tid; x < (n + 1) * nk; x += blockDim.x) {
            int y = x / nk, z = x - y * nk;
            Nr[y * RING_KSET + z] = a.N[(size_t)qi[y] * a.K + kset[z]];
        }
        if (tid < nk) cs[tid] = a.Csum[kset[tid]];
        __syncthreads();
        for (int y = tid; y <= n; y += blockDim.x)
            for (int e = ap0[y]; e < ap1[y]; ++e) {
                int k = a.a_col[e];
                for (int z = 0; z < nk; ++z) if (kset[z] == k) Ar[y * RING_KSET + z] = a.a_val[e];
            }
        __syncthreads();
        // rows/columns of this thread that carry neuron entries (index n = centre pixel, needed by row n1)
        unsigned rmask = 0u, cmask = 0u;
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            int i = ti + 16 * aa, j = tj + 16 * aa;
            if (i == n1) i = n;
            if (i <= n && ap1[i] > ap0[i]) rmask |= 1u << aa;
            if (j <= n && ap1[j] > ap0[j]) cmask |= 1u << aa;
        }
        const bool centre_has = ap1[n] > ap0[n];
#pragma unroll
        for (int aa = 0; aa < 8; ++aa)
#pragma unroll
            for (int bb = 0; bb <= aa; ++bb) {
                const int i = ti + 16 * aa, j = tj + 16 * bb;
                const bool touch = ((rmask >> aa) & 1u) || ((cmask >> bb) & 1u) || (i == n1 && centre_has);
                if (touch && j <= i && i <= n1 && j <= n) {
                    double c = 0.0;
                    if (i < n) {
                        for (int z = 0; z < nk; ++z)
                            c += Ar[j * RING_KSET + z] * Nr[i * RING_KSET + z] + Ar[i * RING_KSET + z] * Nr[j * RING_KSET + z];
                    } else if (i == n) {
                        if (j < n) for (int z = 0; z < nk; ++z) c += Ar[j * RING_KSET + z] * cs[z];
                    } else {
                        if (j < n) {
                            for (int z = 0; z < nk; ++z)
                                c += Ar[n * RING_KSET + z] * Nr[j * RING_KSET + z] + Ar[j * RING_KSET + z] * Nr[n * RING_KSET + z];
                        } else {
                            for (int z = 0; z < nk; ++z) c += Ar[n * RING_KSET + z] * cs[z];
                        }
                    }
                    R.g[aa][bb] -= c;
                }
            }
    }
    // --- ridge: trace over the n1 x n1 system (diagonal owners are the threads with ti == tj)
    {
        double tr = 0.0;
        if (ti == tj) {
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) if (ti + 16 * aa < n1) tr += R.g[aa][aa];
        }
        __syncthreads();
        if (tid == 0) s_tr = 0.0;
        __syncthreads();
        if (ti == tj) atomicAdd(&s_tr, tr);
        __syncthreads();
        const double lam = s_tr * 1e-5;
        if (ti == tj) {
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) if (ti + 16 * aa < n1) R.g[aa][aa] += lam;
        }
    }
    // --- Cholesky of the augmented matrix
    ring_chol_block<0>(R, n1, ti, tj, colk, piv);
    if (n1 > 16) ring_chol_block<1>(R, n1, ti, tj, colk, piv);
    if (n1 > 32) ring_chol_block<2>(R, n1, ti, tj, colk, piv);
    if (n1 > 48) ring_chol_block<3>(R, n1, ti, tj, colk, piv);
    if (n1 > 64) ring_chol_block<4>(R, n1, ti, tj, colk, piv);
    if (n1 > 80) ring_chol_block<5>(R, n1, ti, tj, colk, piv);
    if (n1 > 96) ring_chol_block<6>(R, n1, ti, tj, colk, piv);
    if (n1 > 112) ring_chol_block<7>(R, n1, ti, tj, colk, piv);
    __syncthreads();
    // --- write out L (scaling column k by 1/sqrt(g_kk)), back substitution L' w = z (z = row n1) by one warp
    for (int k = tid; k < n1; k += blockDim.x) piv[k] = 1.0 / sqrt(piv[k]);
    __syncthreads();
#pragma unroll
    for (int aa = 0; aa < 8; ++aa)
#pragma unroll
        for (int bb = 0; bb <= aa; ++bb) {
            const int i = ti + 16 * aa, j = tj + 16 * bb;
            // diagonal: L(j,j) = sqrt(g_jj) = g_jj * (1/sqrt(g_jj));  below: x_i * (1/sqrt(g_jj))
            if (j <= i && i <= n1 && j < n1) L[(size_t)i * (i + 1) / 2 + j] = R.g[aa][bb] * piv[j];
        }
    __syncthreads();
    double* rdiag = piv;   // 1/L(k,k) = 1/sqrt(g_kk)
    if (tid < 32) {
        // lane l keeps z_i for i = l + 32 m (m < 4) in registers; step k broadcasts w_k by shuffle
        double z[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) { int i = tid + 32 * m; z[m] = (i < n1) ? L[(size_t)n1 * (n1 + 1) / 2 + i] : 0.0; }
        for (int k = n1 - 1; k >= 0; --k) {
            const double* row = L + (size_t)k * (k + 1) / 2;
            const int km = k >> 5;
            double zk = km == 0 ? z[0] : (km == 1 ? z[1] : (km == 2 ? z[2] : z[3]));
            const double wk = __shfl_sync(0xffffffffu, zk, k & 31) * rdiag[k];
            if (tid == (k & 31)) colk[k] = wk;
#pragma unroll
            for (int m = 0; m < 4; ++m) { int i = tid + 32 * m; if (i < k) z[m] -= row[i] * wk; }
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) a.W[(size_t)p * g.nnb + slot[i]] = colk[i] + 1e-100;
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
