// kernels_ring_mma.cuh -- per-pixel ring regression solve (endoscope/fit_ring_model.m:92-108) as a BLOCK LDL' factorisation
// whose trailing updates run on the fp64 tensor-core path (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4).
//
// Why: the register-tile SIMT solver (kernels_ring.cuh, kept as the checker) is bound by shared-memory OPERAND bandwidth: a
// rank-1 update of an 8 x 8 register tile needs 16 operand doubles for 36 useful DFMA, and the crossbar delivers 16 doubles per
// clock per SM against 64 DFMA per clock (measured: LDS wavefronts 50 % of peak over the launch, fp64 pipe 27 %).  A DMMA does
// 256 FMA for 2 operand doubles per thread, at the SAME 64 FMA/clk/SM the DFMA pipe has (scripts/micro/dmma.cu: 36.5 TFLOP/s
// either way on B200), so the factorisation becomes fp64-pipe bound instead of LDS bound.
//
// Layout.  Index space 0..127 = 16 blocks of 8: 0..n-1 ring pixels, n the ones row, n+1..126 padding (identity), 127 the
// right-hand side (centre pixel) -- the forward substitution falls out of the factorisation of the augmented matrix.
// 256 threads = 8 warps; warp w owns block rows w ("row A") and 15-w ("row B") = 17 lower-triangle blocks, each held as a DMMA
// C fragment (lane = 4r+q holds (r, 2q), (r, 2q+1)): 34 doubles per thread.  Slot s of a warp is block (15-w, s) for
// s <= 15-w and block (w, 16-s) above, so the blocks of block-column K sit in the STATIC slots K and 16-K of every warp and the
// whole factorisation indexes registers at compile time.
//
// Step K (block pivot D = G_KK, 8 x 8):   Minv = D^-1 (in-warp Gauss-Jordan on the fragment);  P_I = X_I Minv for the panel
// blocks X_I = G_IK (2 DMMA);  G_IJ -= P_I X_J' for K < J <= I (2 DMMA each).  P_I stays in the registers of G_IK: it is the
// block-unit-lower factor L(I,K), and its row 127 is the D-solved forward-substituted right-hand side.  Then x = L^-T w.
#pragma once
#include "kernels_ring.cuh"

namespace cnmfe {

// RM_PIX pixels per CTA (8 warps each).  2 = two pixels factorise in lock step behind CTA-wide barriers (see rm_factor_step);
// 1 = one pixel per CTA, two independent CTAs per SM.  RM_PIVOT2 = 2 x 2 block pivots in the 8 x 8 inversion.
#ifndef RM_PIX
#define RM_PIX 1
#endif
#ifndef RM_PIVOT2
#define RM_PIVOT2 0
#endif
#define RM_GROUP 256
#define RM_THREADS (RM_GROUP * RM_PIX)
#define RM_KSET 8
// barrier of one pixel group (named barrier 1 + group, 256 threads); __syncthreads() = both groups
#define RM_GSYNC() asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(RM_GROUP) : "memory")

__device__ __forceinline__ void rm_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// element (row r, col c) of an 8 x 8 block in "fragment order": lane 4r+q reads cols q and q+4 with one 16-byte load
__device__ __forceinline__ int rm_fo(int r, int c) { return r * 8 + (c & 3) * 2 + (c >> 2); }
__device__ __forceinline__ void rm_store_fo(double* blk, int r, int q, double v0, double v1) {
    blk[rm_fo(r, 2 * q)] = v0;
    blk[rm_fo(r, 2 * q + 1)] = v1;
}

struct alignas(16) RmSmem {
    double X[2][16 * 64];      // panel blocks G_IK, fragment order, double-buffered by step parity
    double NP[2][16 * 64];     // -P_I
    double Mi[2][64];          // D^-1
    double ym[128], S1[128], s1c[128];
    double xs[128], zs[128];
    double part[8][8];
    double dg[128];
    long long qoff[128];
    int qi[128], slot[128], elin[128], ap0[128], ap1[128], kall[128];
    unsigned bits[128];
    int wcnt[8];
    int n, nk;
};
__host__ __device__ inline size_t ring_solve_mma_smem_bytes() { return RM_PIX * sizeof(RmSmem); }

// in-warp inverse of an 8 x 8 SPD block held as a C fragment: block Gauss-Jordan with 2 x 2 pivots (principal 2 x 2 blocks of an
// SPD matrix are SPD, so no pivoting).  One reciprocal (of the 2 x 2 determinant) per TWO eliminated columns: the dependent
// chain -- what the whole factorisation waits for -- is 4 x (shuffle, det, rcp, 3 fma levels) instead of 8 x (shuffle, rcp, 2).
// Pivot columns 2m, 2m+1 are exactly the two elements of the lanes with q == m.
#if !RM_PIVOT2
__device__ __forceinline__ void rm_invert8(double& m0, double& m1, int r, int q) {      // scalar pivots
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double mk = (k & 1) ? m1 : m0;
        const double rk0 = __shfl_sync(0xffffffffu, m0, 4 * k + q), rk1 = __shfl_sync(0xffffffffu, m1, 4 * k + q);
        const double f = __shfl_sync(0xffffffffu, mk, 4 * r + (k >> 1));
        const double piv = __shfl_sync(0xffffffffu, mk, 4 * k + (k >> 1));
        const double rp = __drcp_rn(piv);
        double s0 = rk0 * rp, s1 = rk1 * rp;
        if (2 * q == k) s0 = rp;
        if (2 * q + 1 == k) s1 = rp;
        if (r == k) { m0 = s0; m1 = s1; }
        else {
            const double t0 = (2 * q == k) ? 0.0 : m0, t1 = (2 * q + 1 == k) ? 0.0 : m1;
            m0 = fma(-f, s0, t0); m1 = fma(-f, s1, t1);
        }
    }
}
#else
__device__ __forceinline__ void rm_invert8(double& m0, double& m1, int r, int q) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int k = 2 * m;
        // pivot block [[a, b], [b, c]]
        const double a = __shfl_sync(0xffffffffu, m0, 4 * k + m), b = __shfl_sync(0xffffffffu, m1, 4 * k + m);
        const double c = __shfl_sync(0xffffffffu, m1, 4 * (k + 1) + m);
        // my row's entries in the pivot columns, the pivot rows' entries in my columns
        const double f0 = __shfl_sync(0xffffffffu, m0, 4 * r + m), f1 = __shfl_sync(0xffffffffu, m1, 4 * r + m);
        double g00 = __shfl_sync(0xffffffffu, m0, 4 * k + q), g01 = __shfl_sync(0xffffffffu, m1, 4 * k + q);
        double g10 = __shfl_sync(0xffffffffu, m0, 4 * (k + 1) + q), g11 = __shfl_sync(0xffffffffu, m1, 4 * (k + 1) + q);
        if (q == m) { g00 = 1.0; g01 = 0.0; g10 = 0.0; g11 = 1.0; }      // the pivot columns receive Pinv / -f * Pinv
        const double rd = __drcp_rn(fma(a, c, -(b * b)));
        const double i00 = c * rd, i01 = -(b * rd), i11 = a * rd;
        const double s00 = fma(i01, g10, i00 * g00), s01 = fma(i01, g11, i00 * g01);     // Pinv * (pivot rows)
        const double s10 = fma(i11, g10, i01 * g00), s11 = fma(i11, g11, i01 * g01);
        if (r == k) { m0 = s00; m1 = s01; }
        else if (r == k + 1) { m0 = s10; m1 = s11; }
        else {
            const double t0 = (q == m) ? 0.0 : m0, t1 = (q == m) ? 0.0 : m1;
            m0 = fma(-f1, s10, fma(-f0, s00, t0));
            m1 = fma(-f1, s11, fma(-f0, s01, t1));
        }
    }
}
#endif

// pivot block of step K held in (m0, m1) by its owner warp: invert and publish (K < 15), or -- block 15, whose last row/column is
// the right-hand side -- pivot on [[A', 0], [0, 1]] and solve the row against it
template <int K>
__device__ __forceinline__ void rm_pivot(double m0, double m1, const int r, const int q, RmSmem& S) {
    if (K == 15) {
        const double z0 = m0, z1 = m1;                       // row 7 = rhs entries (lanes r == 7)
        if (r == 7) { m0 = 0.0; m1 = (q == 3) ? 1.0 : 0.0; }
        else if (q == 3) m1 = 0.0;
        rm_invert8(m0, m1, r, q);
        const double zv0 = __shfl_sync(0xffffffffu, z0, 28 + (r >> 1)), zv1 = __shfl_sync(0xffffffffu, z1, 28 + (r >> 1));
        const double zr = (r & 1) ? zv1 : zv0;               // z[r]
        double t0 = zr * m0, t1 = zr * m1;
        if (r == 7) { t0 = 0.0; t1 = 0.0; }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); }
        if (r == 0) { S.xs[120 + 2 * q] = t0; S.xs[120 + 2 * q + 1] = (q == 3) ? 0.0 : t1; }
    } else {
        rm_invert8(m0, m1, r, q);
        rm_store_fo(S.Mi[K & 1], r, q, m0, m1);
    }
}

// Step K of the factorisation.  TWO pixels share a CTA (warps 0-7 and 8-15) and run the factorisation in lock step behind
// CTA-wide barriers: the pivot inversion is a chain of dependent fp64 operations, and DMMAs issued by other warps of the same
// SM sub-partition occupy the fp64 pipe 16 cycles at a time -- with two independent CTAs per SM (or a look-ahead that overlaps
// the pivot with the trailing update) every link of the chain queued behind them (measured: 3.2 k cycles per step).  In lock
// step both pixels pivot while no DMMA is in flight on the SM, then both run their trailing updates at the pipe's full rate.
template <int K>
__device__ __forceinline__ void rm_factor_step(double (&acc)[17][2], const int w, const int lane, RmSmem& S) {
    const int r = lane >> 2, q = lane & 3;
    double* Xb = S.X[K & 1];
    double* NPb = S.NP[K & 1];
    const double* Mi = S.Mi[K & 1];
    constexpr int SB = K, SA = 16 - K;     // slots of the column-K blocks (row-B / row-A reading of the slot)
    // (1) pivot block: owner inverts and publishes
    const bool ownA = (K <= 7) && (w == K), ownB = (K >= 8) && (15 - w == K);
    if (ownA || ownB) rm_pivot<K>(ownA ? acc[SA][0] : acc[SB][0], ownA ? acc[SA][1] : acc[SB][1], r, q, S);
    if (K == 15) return;
    __syncthreads();
    // (2) panel: P_I = X_I * Minv for my blocks of column K
    const double2 mb = *reinterpret_cast<const double2*>(Mi + lane * 2);
    const bool panB = (15 - w > K), panA = (K <= 7) && (w > K);
    if (panB) {
        double* xblk = Xb + (15 - w) * 64;
        rm_store_fo(xblk, r, q, acc[SB][0], acc[SB][1]);
        __syncwarp();
        const double2 xa = *reinterpret_cast<const double2*>(xblk + lane * 2);
        double p0 = 0.0, p1 = 0.0;
        rm_dmma(p0, p1, xa.x, mb.x);
        rm_dmma(p0, p1, xa.y, mb.y);
        acc[SB][0] = p0; acc[SB][1] = p1;
        rm_store_fo(NPb + (15 - w) * 64, r, q, -p0, -p1);
        if (w == 0 && r == 7) { S.zs[8 * K + 2 * q] = p0; S.zs[8 * K + 2 * q + 1] = p1; }   // row 127 of P = w_K
    }
    if (K <= 7) {
        if (panA) {
            double* xblk = Xb + w * 64;
            rm_store_fo(xblk, r, q, acc[SA][0], acc[SA][1]);
            __syncwarp();
            const double2 xa = *reinterpret_cast<const double2*>(xblk + lane * 2);
            double p0 = 0.0, p1 = 0.0;
            rm_dmma(p0, p1, xa.x, mb.x);
            rm_dmma(p0, p1, xa.y, mb.y);
            acc[SA][0] = p0; acc[SA][1] = p1;
            rm_store_fo(NPb + w * 64, r, q, -p0, -p1);
        }
    }
    __syncthreads();
    // (3) trailing update of my blocks with J > K
    double2 npB = make_double2(0.0, 0.0), npA = make_double2(0.0, 0.0);
    if (panB) npB = *reinterpret_cast<const double2*>(NPb + (15 - w) * 64 + lane * 2);
    if (K <= 7) { if (panA) npA = *reinterpret_cast<const double2*>(NPb + w * 64 + lane * 2); }
    // Slots K < s < 16-K hold a block with J > K in EVERY warp (row B reading: J = s > K; row A reading: J = 16-s > K), so that
    // part of the loop is branch-free: the compiler hoists the operand loads and interleaves the DMMAs of different slots
#pragma unroll
    for (int s = 0; s < 17; ++s) {
        if (!(s > K && s < 16 - K)) continue;
        const bool isB = (s <= 15 - w);
        const int J = isB ? s : 16 - s;
        const double2 xb = *reinterpret_cast<const double2*>(Xb + J * 64 + lane * 2);
        const double2 np = isB ? npB : npA;
        rm_dmma(acc[s][0], acc[s][1], np.x, xb.x);
        rm_dmma(acc[s][0], acc[s][1], np.y, xb.y);
    }
    // the remaining slots (s >= 16-K, and s > K) belong to row B in the warps with s <= 15-w and are finished blocks elsewhere
#pragma unroll
    for (int s = 0; s < 17; ++s) {
        if (!(s > K && s >= 16 - K && s <= 15)) continue;
        if (s <= 15 - w) {
            const double2 xb = *reinterpret_cast<const double2*>(Xb + s * 64 + lane * 2);
            rm_dmma(acc[s][0], acc[s][1], npB.x, xb.x);
            rm_dmma(acc[s][0], acc[s][1], npB.y, xb.y);
        }
    }
}

template <int K>
__device__ __forceinline__ void rm_back_step(const double (&acc)[17][2], const int w, const int lane, const int grp, RmSmem& S) {
    // x_K = w_K - sum_{I > K} P_IK' x_I   (column K of the factor: slot K of the warps with 15-w > K, slot 16-K of those with w > K)
    const int r = lane >> 2, q = lane & 3;
    double t0 = 0.0, t1 = 0.0;
    if (15 - w > K) { const double xi = S.xs[8 * (15 - w) + r]; t0 = acc[K][0] * xi; t1 = acc[K][1] * xi; }
    if (K <= 7) {
        if (w > K) { const double xi = S.xs[8 * w + r]; t0 = fma(acc[16 - K][0], xi, t0); t1 = fma(acc[16 - K][1], xi, t1); }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) { t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); }
    if (r == 0) { S.part[w][2 * q] = t0; S.part[w][2 * q + 1] = t1; }
    RM_GSYNC();
    if (w == 0 && lane < 8) {
        double s = S.zs[8 * K + lane];
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) s -= S.part[ww][lane];
        S.xs[8 * K + lane] = s;
    }
    RM_GSYNC();
}

template <bool PROF>
__global__ void __launch_bounds__(RM_THREADS, 3 - RM_PIX) ring_solve_mma_kernel(RingSolveArgs a) {
    extern __shared__ __align__(16) unsigned char rm_smem_raw[];
    const RingGeom& g = a.g;
    const int n_act = a.n_active_dev ? *a.n_active_dev : a.n_active;
    if (RM_PIX * (int)blockIdx.x >= n_act) return;
    const int wall = warp_id_uniform();
    const int grp = wall >> 3, w = wall & 7;                      // pixel group of this warp, warp index inside the group
    RmSmem& S = reinterpret_cast<RmSmem*>(rm_smem_raw)[grp];
    // an odd pixel count leaves the last CTA's second group without a pixel: it recomputes the first one's (the lock-step
    // barriers need both groups) and does not store
    const int pidx = RM_PIX * (int)blockIdx.x + grp;
    const bool store = pidx < n_act;
    const int p = a.active_list[store ? pidx : pidx - 1];
    const int tid = threadIdx.x & (RM_GROUP - 1), lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int pr = p % g.nr + g.pr_off, pc = p / g.nr + g.pc_off;
    const size_t qm = (size_t)pc * g.nrb + pr;
    long long pt0 = PROF ? clock64() : 0;
#define RM_PROF(i) do { if (PROF && threadIdx.x == 0) { long long _t = clock64(); atomicAdd(a.prof + (i), (unsigned long long)(_t - pt0)); pt0 = _t; } } while (0)
    // ---- valid ring neighbours (inside the FOV), compacted in slot order
    {
        int dr = 0, dc = 0;
        bool ok = false;
        if (tid < g.nnb) {
            dr = a.off_r[tid]; dc = a.off_c[tid];
            const int fr = pr + dr + g.br0, fc = pc + dc + g.bc0;
            ok = !(fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) S.wcnt[w] = __popc(m);
        if (tid < 128) { S.qi[tid] = (int)qm; S.elin[tid] = 0; S.xs[tid] = 0.0; S.zs[tid] = 0.0; S.bits[tid] = 0u; }
        RM_GSYNC();
        int base = 0;
        for (int ww = 0; ww < w; ++ww) base += S.wcnt[ww];
        if (ok) {
            const int pos = base + __popc(m & ((1u << lane) - 1));
            S.qi[pos] = (pc + dc) * g.nrb + (pr + dr); S.slot[pos] = tid; S.elin[pos] = dc * (4 * g.rr + 1) + dr;
        }
        if (tid == 0) {
            int nn = 0;
            for (int ww = 0; ww < 8; ++ww) nn += S.wcnt[ww];
            S.n = nn;
        }
    }
    RM_GSYNC();
    const int n = S.n;
    if (tid < 128) {
        const int i = tid;
        double y = 0.0, s1 = 0.0, s1c = 0.0;
        int p0 = 0, p1 = 0;
        long long qo = 0;
        if (i < n || i == 127) {
            const int qq = S.qi[i];
            y = a.Ymean[qq]; s1 = a.S1[qq]; s1c = s1 - a.nsel * y;
            p0 = a.a_ptr[qq]; p1 = a.a_ptr[qq + 1];
            qo = (long long)qq * (long long)a.ND;
        } else if (i == n) {
            y = -1.0; s1 = 0.0; s1c = a.nsel;
        }
        S.ym[i] = y; S.S1[i] = s1; S.s1c[i] = s1c;
        S.ap0[i] = p0; S.ap1[i] = p1; S.qoff[i] = qo;
    }
    RM_GSYNC();
    RM_PROF(0);
    // ---- assemble: Cov(i,j) = raw(i,j) - Ybar_j*S1_i - Ybar_i*S1c_j  (see kernels_ring.cuh); diagonal blocks are assembled in
    //      full (the pivot inversion reads both triangles)
    double acc[17][2];
    const int iA = 8 * w + r, iB = 8 * (15 - w) + r;
    {
        const bool viA = (iA < n) || (iA == 127), viB = (iB < n) || (iB == 127);
        const int eiA = S.elin[iA], eiB = S.elin[iB];
        const long long qoA = S.qoff[iA], qoB = S.qoff[iB];
        // addresses of a batch of slots first, then all its loads: the gathers are L2 latency bound, so as many as the register
        // budget allows are kept in flight
#define RM_ASM_BATCH(S0, S1)                                                                                       \
        {                                                                                                          \
            const double* ptr[(S1) - (S0)][2];                                                                     \
            _Pragma("unroll") for (int s = (S0); s < (S1); ++s) {                                                  \
                const bool isB = (s <= 15 - w);                                                                    \
                const int J = isB ? s : 16 - s, j0 = 8 * J + 2 * q;                                                \
                const bool vi = isB ? viB : viA;                                                                   \
                const int ei = isB ? eiB : eiA;                                                                    \
                const long long qoi = isB ? qoB : qoA;                                                             \
                const int2 ej = *reinterpret_cast<const int2*>(S.elin + j0);                                       \
                const longlong2 qj = *reinterpret_cast<const longlong2*>(S.qoff + j0);                            \
                const int l0 = ei - ej.x, l1 = ei - ej.y;                                                          \
                ptr[s - (S0)][0] = (vi && j0 < n) ? a.S2 + ((l0 >= 0 ? qj.x : qoi) + (long long)abs(l0)) : &ring_zero_moment;     \
                ptr[s - (S0)][1] = (vi && j0 + 1 < n) ? a.S2 + ((l1 >= 0 ? qj.y : qoi) + (long long)abs(l1)) : &ring_zero_moment; \
            }                                                                                                      \
            _Pragma("unroll") for (int s = (S0); s < (S1); ++s) { acc[s][0] = __ldg(ptr[s - (S0)][0]); acc[s][1] = __ldg(ptr[s - (S0)][1]); } \
        }
        RM_ASM_BATCH(0, 9)
        RM_ASM_BATCH(9, 17)
#undef RM_ASM_BATCH
    }
    {
        const double ymA = S.ym[iA], s1A = S.S1[iA], ymB = S.ym[iB], s1B = S.S1[iB];
#pragma unroll
        for (int s = 0; s < 17; ++s) {
            const bool isB = (s <= 15 - w);
            const int J = isB ? s : 16 - s;
            const double ymi = isB ? ymB : ymA, s1i = isB ? s1B : s1A;
            const double2 ymj = *reinterpret_cast<const double2*>(S.ym + 8 * J + 2 * q);
            const double2 scj = *reinterpret_cast<const double2*>(S.s1c + 8 * J + 2 * q);
            acc[s][0] = fma(-ymi, scj.x, fma(-ymj.x, s1i, acc[s][0]));
            acc[s][1] = fma(-ymi, scj.y, fma(-ymj.y, s1i, acc[s][1]));
        }
    }
    RM_PROF(1);
    // ---- neuron corrections: Cov_Bf = Cov_Y - N_x.A_y - A_x.N_y ; sum_sel Bf(x) = S1c_x - A_x.Csum
    if (tid < 128)
        for (int e = S.ap0[tid]; e < S.ap1[tid]; ++e) { const int k = a.a_col[e]; atomicOr(&S.bits[(k >> 5) & 127], 1u << (k & 31)); }
    RM_GSYNC();
    if (w == 0) {
        int base = 0;
        for (int w0 = 0; w0 < 128; w0 += 32) {
            const unsigned bits = S.bits[w0 + lane];
            const int cnt = __popc(bits);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            int pos = base + incl - cnt;
            unsigned b = bits;
            while (b) { const int bit = __ffs(b) - 1; b &= b - 1; if (pos < RING_KALL) S.kall[pos] = ((w0 + lane) << 5) + bit; ++pos; }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) S.nk = min(base, RING_KALL);
    }
    RM_GSYNC();
    const int nall = S.nk;
    RM_PROF(2);
    if (PROF && threadIdx.x == 0) atomicAdd(a.prof + 7, (unsigned long long)nall);
    {
        double* XA = &S.X[0][0];      // RM_KSET x 128 A rows of the indices; the panel buffers are idle until the factorisation
        double* XN = &S.NP[0][0];     // RM_KSET x 128 N rows
        for (int kbase = 0; kbase < nall; kbase += RM_KSET) {
            const int* kset = S.kall + kbase;
            const int nk = min(RM_KSET, nall - kbase);
            RM_GSYNC();
            for (int x = tid; x < RM_KSET * 128; x += RM_GROUP) { XA[x] = 0.0; XN[x] = 0.0; }
            RM_GSYNC();
            for (int x = tid; x < 128 * nk; x += RM_GROUP) {
                const int y = x & 127, z = x >> 7;
                if (y < n || y == 127) XN[z * 128 + y] = a.N[(size_t)S.qi[y] * a.K + kset[z]];
                else if (y == n) XN[z * 128 + y] = a.Csum[kset[z]];
            }
            if (tid < 128)
                for (int e = S.ap0[tid]; e < S.ap1[tid]; ++e) {
                    const int k = a.a_col[e];
                    for (int z = 0; z < nk; ++z) if (kset[z] == k) XA[z * 128 + tid] = a.a_val[e];
                }
            RM_GSYNC();
            for (int z = 0; z < nk; ++z) {
                const double* xa = XA + z * 128;
                const double* xn = XN + z * 128;
                const double aA = xa[iA], nA = xn[iA], aB = xa[iB], nB = xn[iB];
#pragma unroll
                for (int s = 0; s < 17; ++s) {
                    const bool isB = (s <= 15 - w);
                    const int J = isB ? s : 16 - s;
                    const double ai = isB ? aB : aA, ni = isB ? nB : nA;
                    const double2 aj = *reinterpret_cast<const double2*>(xa + 8 * J + 2 * q);
                    const double2 nj = *reinterpret_cast<const double2*>(xn + 8 * J + 2 * q);
                    acc[s][0] = fma(-aj.x, ni, fma(-ai, nj.x, acc[s][0]));
                    acc[s][1] = fma(-aj.y, ni, fma(-ai, nj.y, acc[s][1]));
                }
            }
        }
        RM_GSYNC();
    }
    RM_PROF(3);
    // ---- ridge 1e-5 * trace over the indices 0..n (ring + ones row); padding gets a unit diagonal
#pragma unroll
    for (int s = 0; s < 17; ++s) {
        const bool isB = (s <= 15 - w);
        const int I = isB ? 15 - w : w, J = isB ? s : 16 - s;
        if (I == J) {
            if (2 * q == r) S.dg[8 * I + r] = acc[s][0];
            if (2 * q + 1 == r) S.dg[8 * I + r] = acc[s][1];
        }
    }
    RM_GSYNC();
    double lam;
    {
        // fixed-order sum: 16 groups of 8 consecutive indices, then a shuffle tree
        double tr = 0.0;
        if (lane < 16) {
#pragma unroll
            for (int x = 0; x < 8; ++x) { const int i = 8 * lane + x; if (i <= n) tr += S.dg[i]; }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
        lam = __shfl_sync(0xffffffffu, tr, 0) * 1e-5;
    }
#pragma unroll
    for (int s = 0; s < 17; ++s) {
        const bool isB = (s <= 15 - w);
        const int I = isB ? 15 - w : w, J = isB ? s : 16 - s;
        if (I == J) {
            const int i = 8 * I + r;
            if (2 * q == r) acc[s][0] = (i <= n) ? acc[s][0] + lam : 1.0;
            if (2 * q + 1 == r) acc[s][1] = (i <= n) ? acc[s][1] + lam : 1.0;
        }
    }
    RM_PROF(4);
    // ---- block LDL' of the augmented matrix
    __syncthreads();                                               // both pixels enter the factorisation together
    rm_factor_step<0>(acc, w, lane, S);
    rm_factor_step<1>(acc, w, lane, S);
    rm_factor_step<2>(acc, w, lane, S);
    rm_factor_step<3>(acc, w, lane, S);
    rm_factor_step<4>(acc, w, lane, S);
    rm_factor_step<5>(acc, w, lane, S);
    rm_factor_step<6>(acc, w, lane, S);
    rm_factor_step<7>(acc, w, lane, S);
    rm_factor_step<8>(acc, w, lane, S);
    rm_factor_step<9>(acc, w, lane, S);
    rm_factor_step<10>(acc, w, lane, S);
    rm_factor_step<11>(acc, w, lane, S);
    rm_factor_step<12>(acc, w, lane, S);
    rm_factor_step<13>(acc, w, lane, S);
    rm_factor_step<14>(acc, w, lane, S);
    rm_factor_step<15>(acc, w, lane, S);
    __syncthreads();
    RM_PROF(5);
    // ---- x = L^-T w, block columns 14 .. 0 (block 15 was solved with its pivot)
    rm_back_step<14>(acc, w, lane, grp, S);
    rm_back_step<13>(acc, w, lane, grp, S);
    rm_back_step<12>(acc, w, lane, grp, S);
    rm_back_step<11>(acc, w, lane, grp, S);
    rm_back_step<10>(acc, w, lane, grp, S);
    rm_back_step<9>(acc, w, lane, grp, S);
    rm_back_step<8>(acc, w, lane, grp, S);
    rm_back_step<7>(acc, w, lane, grp, S);
    rm_back_step<6>(acc, w, lane, grp, S);
    rm_back_step<5>(acc, w, lane, grp, S);
    rm_back_step<4>(acc, w, lane, grp, S);
    rm_back_step<3>(acc, w, lane, grp, S);
    rm_back_step<2>(acc, w, lane, grp, S);
    rm_back_step<1>(acc, w, lane, grp, S);
    rm_back_step<0>(acc, w, lane, grp, S);
    if (store)
        for (int i = tid; i < n; i += RM_GROUP) a.W[(size_t)p * g.nnb + S.slot[i]] = S.xs[i] + 1e-100;
    RM_PROF(6);
#undef RM_PROF
}

}  // namespace cnmfe
