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
#define RING_KSET 8
#define RING_KALL 128
#define RING_NIDX 128        // index space of the augmented system: 0..n-1 ring pixels, n ones row, n1 = n+1 rhs/centre
#define RING_XSTRIDE 10
#define RING_XBUF (16 * RING_XSTRIDE)
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
    size_t dbl = 2 * RING_XBUF + 128 + 3 * RING_XBUF + 2 * (size_t)RING_KSET * RING_XBUF + RING_KSET + 16 + RING_NIDX /* qoff */ +
                 (size_t)(NMAX + 1) * (NMAX + 2) / 2 + 2;
    size_t ints = 5 * (size_t)RING_NIDX + RING_KALL;
    return dbl * 8 + ints * 4 + 64;
}

// Right-looking LDL' elimination with ONE barrier per column: the owners of column k publish its UNSCALED entries x_i
// (zeros for i <= k) and the reciprocal pivot 1/d_k (double-buffered), then every thread applies
// G(i,j) -= (x_i / d_k) * x_j to its registers.  The registers keep x_i; rd[] keeps the reciprocal pivots, so the unit
// factor M(i,k) = x_i / d_k is formed once, when it is written out.
template <int KA>
__device__ __forceinline__ void ring_chol_block(RingRegs& R, int n1, int ti, int tj, double* xbuf, double* rd) {
    const int kend = min(16 * KA + 15, n1 - 1);
    constexpr int A0 = KA & ~1;          // first (even) register index loaded: keeps the 16-byte alignment
    for (int k = 16 * KA; k <= kend; ++k) {
        const int kr = k & 15;
        double* xb = xbuf + (k & 1) * RING_XBUF;
        if (tj == kr) {
            double* dst = xb + ti * RING_XSTRIDE;
#pragma unroll
            for (int a = KA; a < 8; ++a) dst[a] = (ti + 16 * a > k) ? R.g[a][KA] : 0.0;
            if (ti == kr) { const double r = 1.0 / R.g[KA][KA]; rd[k] = r; xb[8] = r; }   // slot 8 of thread-row 0: padding
        }
        __syncthreads();
        const double rinv = xb[8];
        double ci[8], cj[8];
        const double2* si = reinterpret_cast<const double2*>(xb + ti * RING_XSTRIDE);
        const double2* sj = reinterpret_cast<const double2*>(xb + tj * RING_XSTRIDE);
#pragma unroll
        for (int a = A0; a < 8; a += 2) {
            const double2 vi = si[a >> 1], vj = sj[a >> 1];
            ci[a] = vi.x * rinv; ci[a + 1] = vi.y * rinv;
            cj[a] = vj.x; cj[a + 1] = vj.y;
        }
#pragma unroll
        for (int a = KA; a < 8; ++a)
#pragma unroll
            for (int b = KA; b <= a; ++b) R.g[a][b] = fma(-ci[a], cj[b], R.g[a][b]);
    }
}

__global__ void __launch_bounds__(RING_SOLVE_THREADS, 2) ring_solve_kernel(RingSolveArgs a) {
    extern __shared__ double smem[];
    const RingGeom& g = a.g;
    const int p = a.active_list[blockIdx.x];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    const int pr = p % g.nr + g.pr_off, pc = p / g.nr + g.pc_off;
    const size_t qm = (size_t)pc * g.nrb + pr;
    const int NMAX = g.nnb + 1;
    // shared layout
    double* colk = smem;                                      // 2*RING_XBUF: double-buffered column, later the solution
    double* rd = colk + 2 * RING_XBUF;                        // 128: reciprocal pivots 1/d_k
    double* ymp = rd + 128;                                   // permuted per-index vectors (see above)
    double* S1p = ymp + RING_XBUF;
    double* s1cp = S1p + RING_XBUF;
    double* XA = s1cp + RING_XBUF;                            // RING_KSET x RING_XBUF: A rows of the indices
    double* XN = XA + RING_KSET * RING_XBUF;                  // RING_KSET x RING_XBUF: N rows of the indices
    double* cs = XN + RING_KSET * RING_XBUF;                  // RING_KSET (+ 16 partial traces)
    double* trs = cs + RING_KSET;
    long long* qoff = reinterpret_cast<long long*>(trs + 16); // RING_NIDX: block pixel index * ND
    double* L = reinterpret_cast<double*>(qoff + RING_NIDX);  // packed rows of the unit factor, (NMAX+1)(NMAX+2)/2
    int* qi = reinterpret_cast<int*>(L + (size_t)(NMAX + 1) * (NMAX + 2) / 2 + 2);   // RING_NIDX block pixel index
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
        if (tid < RING_NIDX) { qi[tid] = (int)qm; elin[tid] = 0; }
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
    __shared__ unsigned s_bits[128];
    if (tid < 128) s_bits[tid] = 0u;
    __syncthreads();
    if (tid < RING_NIDX)
        for (int e = ap0[tid]; e < ap1[tid]; ++e) { int k = a.a_col[e]; atomicOr(&s_bits[(k >> 5) & 127], 1u << (k & 31)); }
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
    // --- LDL' of the augmented matrix
    ring_chol_block<0>(R, n1, ti, tj, colk, rd);
    if (n1 > 16) ring_chol_block<1>(R, n1, ti, tj, colk, rd);
    if (n1 > 32) ring_chol_block<2>(R, n1, ti, tj, colk, rd);
    if (n1 > 48) ring_chol_block<3>(R, n1, ti, tj, colk, rd);
    if (n1 > 64) ring_chol_block<4>(R, n1, ti, tj, colk, rd);
    if (n1 > 80) ring_chol_block<5>(R, n1, ti, tj, colk, rd);
    if (n1 > 96) ring_chol_block<6>(R, n1, ti, tj, colk, rd);
    if (n1 > 112) ring_chol_block<7>(R, n1, ti, tj, colk, rd);
    __syncthreads();
    // --- write out the strictly-lower part of the unit factor M(i,j) = x_ij / d_j (rows packed); row n1 is then
    //     y = D^-1 M^-1 rhs, and M' w = y is solved by one warp
#pragma unroll
    for (int aa = 0; aa < 8; ++aa)
#pragma unroll
        for (int bb = 0; bb <= aa; ++bb) {
            const int i = ti + 16 * aa, j = tj + 16 * bb;
            if (j < i && i <= n1) L[(size_t)i * (i + 1) / 2 + j] = R.g[aa][bb] * rd[j];
        }
    __syncthreads();
    if (tid < 32) {
        // lane l keeps y_i for i = l + 32 m (m < 4) in registers; step k broadcasts w_k by shuffle and subtracts
        // M(k, i) * w_k from the entries i < k; row k-1 is fetched while step k runs
        double z[4], rc[4];
        const double* rown = L + (size_t)n1 * (n1 + 1) / 2;
#pragma unroll
        for (int m = 0; m < 4; ++m) { const int i = tid + 32 * m; z[m] = (i < n1) ? rown[i] : 0.0; }
        {
            const double* row = L + (size_t)n * (n + 1) / 2;
#pragma unroll
            for (int m = 0; m < 4; ++m) { const int i = tid + 32 * m; rc[m] = (i < n) ? row[i] : 0.0; }
        }
        for (int k = n; k >= 0; --k) {
            double rn[4] = {0.0, 0.0, 0.0, 0.0};
            if (k > 0) {
                const double* row = L + (size_t)(k - 1) * k / 2;
#pragma unroll
                for (int m = 0; m < 4; ++m) { const int i = tid + 32 * m; if (i < k - 1) rn[m] = row[i]; }
            }
            const int km = k >> 5;
            const double zk = km == 0 ? z[0] : (km == 1 ? z[1] : (km == 2 ? z[2] : z[3]));
            const double wk = __shfl_sync(0xffffffffu, zk, k & 31);
            if (tid == (k & 31)) colk[k] = wk;
#pragma unroll
            for (int m = 0; m < 4; ++m) { z[m] = fma(-rc[m], wk, z[m]); rc[m] = rn[m]; }
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
