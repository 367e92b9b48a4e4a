// oasis.cuh -- block-cooperative device routines for the per-trace part of the CNMF-E temporal update:
// Welch-PSD noise (GetSn), Yule-Walker time constant, order statistics, AR(1)/AR(2) pool-adjacent-violators
// (OASIS) and the FOOPSI hyper-parameter loop.  One CTA (CNMFE_BLOCK threads) owns one trace; all arithmetic
// is IEEE double, in the operation order of the reference wherever that order is observable.
//
// Reference behaviour being reproduced (file:line under /root/reference/OASIS_matlab):
//   GetSn                  functions/GetSn.m:33-41  (+ documented pwelch defaults)
//   estimate_time_constant functions/estimate_time_constant.m:36-67 (randn jitter -> 0; p>2 fallback -> "fail")
//   oasisAR1               packages/oasis/oasisAR1.m:57-109
//   oasisAR2               packages/oasis/oasisAR2.m:49-156
//   foopsi_oasisAR1        packages/oasis/foopsi_oasisAR1.m:78-179 (incl. the shared-`h` quirk of update_g)
//   constrained_oasisAR1   packages/oasis/constrained_oasisAR1.m:84-199
//   thresholded_oasisAR1   packages/oasis/thresholded_oasisAR1.m:104-213
//   fminbnd                MathWorks fminbnd = FMM golden section + parabolic interpolation, TolX 1e-4
#pragma once
#include "common.cuh"

namespace cnmfe {

// Per-CTA workspace in global memory (all arrays sized for T, see trace_ws_doubles()).
struct TraceWS {
    double* yb;    // T    working trace (y - b)
    double* c;     // T
    double* s;     // T
    double* pv;    // T    pools: v
    double* pw;    // T    pools: w
    int* pt;       // T    pools: start (0-based)
    int* pl;       // T    pools: length
    double* h;     // T+1  kernel table of update_g / AR2 g11
    double* hh;    // T+1  cumsum(h.^2)          / AR2 g12
    double* gp;    // 2T+2 powers g^l            / AR2 g11g11, g11g12 (T each)
    double* scr;   // scratch: max(3*nfft + nfft/4 + 8, 2T) doubles
    double* sv;    // T+1  saved pools v (constrained/thresholded trial copies)
    double* sw;    // T+1
    int* st;       // T
    int* sl;       // T
};

#define TRACE_PART 1280
struct BlockShared {
    double red[32];
    int hist[256];
    int ibc[4];
    double dbc[4];
    double part[TRACE_PART];    // block-wide scratch: scan partials, the two factor tables of block_pow_table (64 + 512), the
                                // partial sums of rss_g (one per (thread chunk, pool) pair while they fit, else global scratch)
    double2 stage[32];          // cold scan: (a_m, b_m) of the current window, read back as broadcast 16-byte loads
    double2 snap[32];           // cold scan: running (v, w) before element m
    double2* zfft;              // GetSn FFT buffer in dynamic shared memory (nfft complex), or nullptr -> global scratch
    double* ysm;                // update_g: shared-memory copy of the trace being fitted (aliases zfft), or nullptr
    double* hsm;                // update_g: kernel table h = g^(0..maxl) and its cumsum of squares when maxl < hcap,
    double* hhsm;               //           else ws.h / ws.hh are used
    int* ptsm;                  // update_g: pool starts / lengths when n <= pcap, else ws.pt / ws.pl
    int* plsm;
    int hcap, pcap;
    int dyn_bytes;              // dynamic shared memory actually allocated for this launch (trace_smem_bind)
    int nlong, long_thr;        // update_g: pools longer than long_thr, in ascending pool order (handled by the whole CTA)
    int longp[128];
    unsigned long long* prof;   // optional per-phase cycle counters (CNMFE_HALS_PROFILE diagnostics), else nullptr
    long long t0;
    unsigned long long pc[32];  // per-CTA accumulators, flushed to prof[] when the kernel ends
};

// diagnostics: add the cycles since the previous mark to counter i (thread 0 only; no-op unless profiling is on)
#define CNMFE_PROF(sh, i)                                                                     \
    do {                                                                                      \
        if ((sh)->prof) {                                                                     \
            if (threadIdx.x == 0) {                                                           \
                long long _t = clock64();                                                     \
                (sh)->pc[i] += (unsigned long long)(_t - (sh)->t0);                           \
                (sh)->t0 = _t;                                                                \
            }                                                                                 \
            __syncwarp();   /* re-converge warp 0: a divergent warp takes the slow shuffle path */ \
        }                                                                                     \
    } while (0)

__host__ __device__ inline int nextpow2_int(int L) {
    int n = 1;
    while (n < L) n <<= 1;
    return n;
}

__host__ __device__ inline int welch_nfft(int T) {
    int L = (int)floor((double)T / 4.5);
    int n = nextpow2_int(L);
    return n < 256 ? 256 : n;
}

// Dynamic shared memory of the per-trace kernels.  Two users that never overlap in time share it:
//   GetSn      the Welch FFT buffer (nfft complex doubles) when it fits 64 KB;
//   update_g   a copy of the trace (T doubles), the kernel table h and its cumsum (TRACE_HCAP doubles each) and the pool
//              starts/lengths (TRACE_PCAP ints each): fminbnd evaluates rss_g ~24 times per item and from L2 the dependent
//              look-ups were 85 % of rss_g;
//   cold scan  the same trace buffer + the head of the power table in the h / hh region (the scan is one warp's chain);
//   quantiles  (none: the radix select reads the trace through L1 -- keep the carve-out small).
#define TRACE_HCAP 1024
#define TRACE_PCAP 1024
struct TraceSmem { size_t total, y_bytes; int fft, stage; };
__host__ __device__ inline TraceSmem trace_smem_layout(int T, bool want_stage) {
    TraceSmem L;
    const size_t fb = (size_t)welch_nfft(T) * 16;
    L.fft = fb <= 65536 ? 1 : 0;
    L.y_bytes = (((size_t)T * 8) + 15) / 16 * 16;
    const size_t sb = L.y_bytes + 2 * (size_t)TRACE_HCAP * 8 + 2 * (size_t)TRACE_PCAP * 4;
    L.stage = (want_stage && sb <= 106496) ? 1 : 0;   // <= 104 KB: two CTAs per SM
    L.total = L.fft ? fb : 0;
    if (L.stage && sb > L.total) L.total = sb;
    return L;
}
// called by thread 0 of a per-trace kernel (mode bit 0: fft, bit 1: stage)
__device__ inline void trace_smem_bind(BlockShared* sh, unsigned char* dyn, int T, int mode) {
    const TraceSmem L = trace_smem_layout(T, true);
    sh->zfft = (mode & 1) ? reinterpret_cast<double2*>(dyn) : nullptr;
    sh->ysm = (mode & 2) ? reinterpret_cast<double*>(dyn) : nullptr;
    sh->hsm = (mode & 2) ? reinterpret_cast<double*>(dyn + L.y_bytes) : nullptr;
    sh->hhsm = (mode & 2) ? sh->hsm + TRACE_HCAP : nullptr;
    sh->ptsm = (mode & 2) ? reinterpret_cast<int*>(sh->hhsm + TRACE_HCAP) : nullptr;
    sh->plsm = (mode & 2) ? sh->ptsm + TRACE_PCAP : nullptr;
    sh->hcap = (mode & 2) ? TRACE_HCAP : 0;
    sh->pcap = (mode & 2) ? TRACE_PCAP : 0;
    {
        const size_t fb = (mode & 1) ? (size_t)welch_nfft(T) * 16 : 0;
        const size_t sb = (mode & 2) ? L.y_bytes + 2 * (size_t)TRACE_HCAP * 8 + 2 * (size_t)TRACE_PCAP * 4 : 0;
        sh->dyn_bytes = (int)(fb > sb ? fb : sb);
    }
}

__host__ __device__ inline size_t trace_scratch_doubles(int T) {
    size_t nfft = (size_t)welch_nfft(T);
    size_t a = 3 * nfft + nfft / 4 + 8;
    size_t b = 2 * (size_t)T + 8;
    return a > b ? a : b;
}

// ------------------------------------------------------------------------------------------------ order stats
__device__ __forceinline__ unsigned long long dkey(double x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// k-th smallest (0-based) of x[0..n) by 8-pass MSB radix select.  All threads call; result in all threads.
__device__ double select_kth(const double* __restrict__ x, int n, int k, BlockShared* sh) {
    unsigned long long prefix = 0, mask = 0;
    __syncthreads();
    if (threadIdx.x == 0) sh->ibc[0] = k;
    for (int pass = 7; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) sh->hist[i] = 0;
        __syncthreads();
        int shift = pass * 8;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            unsigned long long key = dkey(x[i]);
            if ((key & mask) == prefix) atomicAdd(&sh->hist[(int)((key >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (warp_id_uniform() == 0) {
            // warp 0: lane owns 8 consecutive bins; exactly one lane contains rank kk
            const int lane = threadIdx.x, kk = sh->ibc[0];
            int h[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = sh->hist[8 * lane + j]; tot += h[j]; }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            int cum = incl - tot;
            if (cum <= kk && kk < incl) {
                int b = 0;
#pragma unroll
                for (b = 0; b < 7; ++b) { if (cum + h[b] > kk) break; cum += h[b]; }
                sh->ibc[0] = kk - cum;
                sh->ibc[1] = 8 * lane + b;
            }
        }
        __syncthreads();
        unsigned long long b = (unsigned long long)sh->ibc[1];
        prefix |= (b << shift);
        mask |= (255ull << shift);
        __syncthreads();
    }
    return dkey_inv(prefix);
}

// (k+1)-th smallest given a = the k-th smallest (0-based): a again if it has a duplicate at rank k+1, else the
// smallest element above a.  One pass instead of a second 8-pass select.
__device__ double select_next(const double* __restrict__ x, int n, int k, double a, BlockShared* sh) {
    double le = 0.0, mn = INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = x[i];
        if (v <= a) le += 1.0; else mn = fmin(mn, v);
    }
    le = block_sum(le, sh->red);            // exact: integer counts < 2^53
    mn = -block_max(-mn, sh->red);
    return ((int)le >= k + 2) ? a : mn;
}

__device__ double block_median(const double* x, int n, BlockShared* sh) {
    if (n & 1) return select_kth(x, n, n / 2, sh);
    double a = select_kth(x, n, n / 2 - 1, sh);
    double b = select_next(x, n, n / 2 - 1, a, sh);
    return (a + b) / 2.0;
}

// MATLAB quantile(x,p): linear interpolation at plotting positions (i-0.5)/n.
__device__ double block_quantile(const double* x, int n, double p, BlockShared* sh) {
    double pos = p * (double)n + 0.5;   // 1-based fractional rank
    if (pos < 1.0) return select_kth(x, n, 0, sh);
    if (pos >= (double)n) return select_kth(x, n, n - 1, sh);
    int lo = (int)floor(pos);
    double fr = pos - (double)lo;
    double a = select_kth(x, n, lo - 1, sh);
    double b = select_next(x, n, lo - 1, a, sh);
    return a + fr * (b - a);
}

// HALS_temporal.m:78  b = mean(ck_raw(ck_raw<median(ck_raw)))
__device__ double block_mean_below(const double* x, int n, double thr, BlockShared* sh) {
    double s = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = x[i];
        if (v < thr) { s += v; cnt += 1.0; }
    }
    s = block_sum(s, sh->red);
    cnt = block_sum(cnt, sh->red);
    return s / cnt;   // NaN when empty, as in MATLAB
}

// ------------------------------------------------------------------------------------------------ GetSn
__device__ __forceinline__ double2 cmul2(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// In-place radix-4 decimation-in-frequency FFT of z[0..n) (interleaved re,im), twiddles tw[j] = (cos, -sin)(2 pi j / n), j < n/2.
// The input is in natural order and the output is left in BIT-REVERSED order (X[k] at z[brev(k)]): the caller indexes it
// through fft_brev, so there is no permutation pass (in shared memory a bit reversal is a 32-way bank conflict), and two
// radix-2 stages per barrier halve the latency-bound stage count.  One radix-4 butterfly = stage s on (x0,x2), (x1,x3) and
// stage s-1 on the results; w_m^(j+m/4) = -i w_m^j, w_(m/2)^j = w_m^(2j).
__device__ __forceinline__ int fft_brev(int i, int logn) { return (int)(__brev((unsigned)i) >> (32 - logn)); }
__device__ void block_fft(double2* z, const double2* tw, int n, int logn) {
    int s = logn;
    for (; s >= 2; s -= 2) {
        const int quarter = 1 << (s - 2), tshift = logn - s;
        for (int b = threadIdx.x; b < (n >> 2); b += blockDim.x) {
            const int g = b >> (s - 2), j = b & (quarter - 1);
            const int i0 = (g << s) + j;
            const double2 w1 = tw[j << tshift], w2 = tw[(2 * j) << tshift];
            const double2 x0 = z[i0], x1 = z[i0 + quarter], x2 = z[i0 + 2 * quarter], x3 = z[i0 + 3 * quarter];
            const double2 a0 = make_double2(x0.x + x2.x, x0.y + x2.y), d0 = make_double2(x0.x - x2.x, x0.y - x2.y);
            const double2 a1 = make_double2(x1.x + x3.x, x1.y + x3.y), d1 = make_double2(x1.y - x3.y, x3.x - x1.x);   // -i (x1 - x3)
            const double2 a2 = cmul2(d0, w1), a3 = cmul2(d1, w1);
            z[i0] = make_double2(a0.x + a1.x, a0.y + a1.y);
            z[i0 + quarter] = cmul2(make_double2(a0.x - a1.x, a0.y - a1.y), w2);
            z[i0 + 2 * quarter] = make_double2(a2.x + a3.x, a2.y + a3.y);
            z[i0 + 3 * quarter] = cmul2(make_double2(a2.x - a3.x, a2.y - a3.y), w2);
        }
        __syncthreads();
    }
    if (s == 1) {
        for (int b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
            const double2 u = z[2 * b], v = z[2 * b + 1];
            z[2 * b] = make_double2(u.x + v.x, u.y + v.y);
            z[2 * b + 1] = make_double2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

// sn = sqrt(exp(mean(log(Pxx(f)/2)))), 0.25 <= f <= 0.5, Pxx = pwelch(x,[],[],[],1).
// scr needs 3*nfft + nfft/4 + 8 doubles.
__device__ double block_getsn(const double* __restrict__ x, int N, double* scr, BlockShared* sh) {
    const int L = (int)floor((double)N / 4.5);
    const int nov = L / 2;
    const int nseg = (N - nov) / (L - nov);
    const int nfft = welch_nfft(N);
    int logn = 0;
    while ((1 << logn) < nfft) ++logn;
    double2* z = sh->zfft ? sh->zfft : reinterpret_cast<double2*>(scr);
    // twiddles in the global scratch (L1): the early stages read them with strides of 2^k entries, which in shared memory
    // would be 32-way bank conflicts (measured: GetSn 195 k -> 345 k cycles with a shared-memory table)
    double2* tw = reinterpret_cast<double2*>(scr + 2 * (size_t)nfft);
    double* acc = scr + 3 * (size_t)nfft;
    const int f0 = nfft / 4, nf = nfft / 4 + 1;
    __syncthreads();
    for (int j = threadIdx.x; j < nfft / 2; j += blockDim.x) {
        double sn_, cs_;
        sincospi(2.0 * (double)j / (double)nfft, &sn_, &cs_);
        tw[j] = make_double2(cs_, -sn_);
    }
    for (int j = threadIdx.x; j < nf; j += blockDim.x) acc[j] = 0.0;
    double usum = 0.0;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        double w = 0.54 - 0.46 * cospi(2.0 * (double)j / (double)(L - 1));
        usum += w * w;
    }
    usum = block_sum(usum, sh->red);
    const int step = L - nov;
    for (int sg = 0; sg < nseg; sg += 2) {
        const bool two = (sg + 1 < nseg);
        const double* xa = x + (size_t)sg * step;
        const double* xb = x + (size_t)(sg + 1) * step;
        for (int j = threadIdx.x; j < nfft; j += blockDim.x) {
            double re = 0.0, im = 0.0;
            if (j < L) {
                double w = 0.54 - 0.46 * cospi(2.0 * (double)j / (double)(L - 1));
                re = w * xa[j];
                if (two) im = w * xb[j];
            }
            z[j] = make_double2(re, im);
        }
        __syncthreads();
        block_fft(z, tw, nfft, logn);
        for (int j = threadIdx.x; j < nf; j += blockDim.x) {
            int f = f0 + j;
            double2 a = z[fft_brev(f, logn)], b = z[fft_brev((nfft - f) & (nfft - 1), logn)];
            double xr = 0.5 * (a.x + b.x), xi = 0.5 * (a.y - b.y);
            double p = xr * xr + xi * xi;
            if (two) {
                double yr = 0.5 * (a.y + b.y), yi = 0.5 * (b.x - a.x);
                p += yr * yr + yi * yi;
            }
            acc[j] += p;
        }
        __syncthreads();
    }
    double ls = 0.0;
    for (int j = threadIdx.x; j < nf; j += blockDim.x) {
        double p = acc[j] / ((double)nseg * usum);
        if (f0 + j != nfft / 2) p *= 2.0;
        ls += log(p / 2.0);
    }
    ls = block_sum(ls, sh->red);
    return sqrt(exp(ls / (double)nf));
}

// ------------------------------------------------------------------------------------------------ time constant
// estimate_time_constant(y, p, sn) for p in {1,2}.  Returns number of coefficients written (p) or 0 on "no stable
// AR(p) model" (the reference then escalates p and deconvolveCa.m:84-101 returns zeros for the trace).
__device__ int block_time_constant(const double* __restrict__ y, int T, int p, double sn, double* g,
                                   BlockShared* sh) {
    const int lags = 5 + p;
    double m = 0.0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) m += y[i];
    m = block_sum(m, sh->red) / (double)T;
    double xc[8];
    for (int k = 0; k <= lags; ++k) {
        double a = 0.0;
        for (int i = threadIdx.x; i + k < T; i += blockDim.x) a += (y[i + k] - m) * (y[i] - m);
        xc[k] = block_sum(a, sh->red) / (double)T;
    }
    const double s2 = sn * sn;
    if (p == 1) {
        double num = 0.0, den = 0.0;
        for (int i = 0; i < lags; ++i) {
            double a = xc[i] - (i == 0 ? s2 : 0.0);
            num += a * xc[i + 1];
            den += a * a;
        }
        double g1 = num / den;
        if (!(fabs(g1) <= 1.0)) return 0;
        if (g1 < 0.0) g1 = 0.15;
        g[0] = g1;
        return 1;
    }
    // p == 2: A(i,1) = xc(i) - s2*[i==0], A(i,2) = xc(|i-1|) - s2*[i==1]
    double a11 = 0, a12 = 0, a22 = 0, b1 = 0, b2 = 0;
    for (int i = 0; i < lags; ++i) {
        double c1 = xc[i] - (i == 0 ? s2 : 0.0);
        double c2 = xc[i == 0 ? 1 : i - 1] - (i == 1 ? s2 : 0.0);
        a11 += c1 * c1; a12 += c1 * c2; a22 += c2 * c2;
        b1 += c1 * xc[i + 1]; b2 += c2 * xc[i + 1];
    }
    double det = a11 * a22 - a12 * a12;
    double g1 = (a22 * b1 - a12 * b2) / det, g2 = (a11 * b2 - a12 * b1) / det;
    double disc = g1 * g1 + 4.0 * g2, r1, r2;
    if (disc < 0.0) {
        if (-g2 > 1.0) return 0;
        r1 = r2 = 0.5 * g1;
    } else {
        double sq = sqrt(disc);
        r1 = 0.5 * (g1 + sq); r2 = 0.5 * (g1 - sq);
        if (fmax(fabs(r1), fabs(r2)) > 1.0) return 0;
    }
    if (r1 > 1.0) r1 = 0.95;
    if (r2 > 1.0) r2 = 0.95;
    if (r1 < 0.0) r1 = 0.15;
    if (r2 < 0.0) r2 = 0.15;
    g[0] = r1 + r2;
    g[1] = -r1 * r2;
    return 2;
}

// ------------------------------------------------------------------------------------------------ AR(1) PAV
// gp[m] = g^m for m in [0, 2T+1]; pow() in the loop is replaced by this table.  Two levels: g^m = g^(64 a) * g^b with both
// factors from pow() (one call per thread instead of ~40: pow is ~1000 cycles of the per-item latency chain), i.e. within
// 2.5 ulp of the correctly rounded power instead of 1 ulp.  Both small tables live in shared memory (sh->part).
__device__ void block_pow_table(double g, int T, double* gp, BlockShared* sh) {
    const int M = 2 * T + 2;                       // entries
    const int na = (M + 63) >> 6;                  // coarse powers g^(64 a), a < na
    double* const pw_fine = sh->part;              // 64
    double* const pw_coarse = sh->part + 64;       // <= 512
    __syncthreads();
    if (na > 512) {                                // very long traces: one pow per entry
        for (int m = threadIdx.x; m < M; m += blockDim.x) gp[m] = pow(g, (double)m);
        __syncthreads();
        return;
    }
    for (int i = threadIdx.x; i < 64 + na; i += blockDim.x) {
        if (i < 64) pw_fine[i] = pow(g, (double)i);
        else pw_coarse[i - 64] = pow(g, (double)(64 * (i - 64)));
    }
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x) gp[m] = pw_coarse[m >> 6] * pw_fine[m & 63];
    __syncthreads();
}

__device__ void block_init_pools_ar1(const double* __restrict__ y, int T, double g, double lam, TraceWS& ws) {
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        ws.pv[i] = (i == T - 1) ? (y[i] - lam) : (y[i] - lam * (1.0 - g));
        ws.pw[i] = 1.0;
        ws.pt[i] = i;
        ws.pl[i] = 1;
    }
    __syncthreads();
}

// oasisAR1.m:57-98 on n pools held in ws (in place, stack form).  Thread 0 runs the scan; returns pool count.
__device__ int block_oasis_ar1_run(TraceWS& ws, int n, double smin, BlockShared* sh) {
    __syncthreads();
    if (threadIdx.x == 0) {
        double* v = ws.pv; double* w = ws.pw; int* t = ws.pt; int* l = ws.pl;
        const double* gp = ws.gp;
        int top = 0;
        double vt = v[0], wt = w[0];
        int lt = l[0];
        // incoming pool i + 1 is loaded while pool i is processed (stores only touch indices <= top < i + 1)
        double nv = (n > 1) ? v[1] : 0.0, nwt = (n > 1) ? w[1] : 1.0;
        int nl = (n > 1) ? l[1] : 0, ntt = (n > 1) ? t[1] : 0;
        for (int i = 1; i < n; ++i) {
            const double vi = nv, wi = nwt;
            const int li = nl, ti = ntt;
            if (i + 1 < n) { nv = v[i + 1]; nwt = w[i + 1]; nl = l[i + 1]; ntt = t[i + 1]; }
            // forward test  vi / wi >= vt / wt * g^lt + smin  (oasisAR1.m:63-64) multiplied through by wi wt > 0; the divisions
            // are only evaluated when the margin is within the rounding of either form (see warp_oasis_ar1_cold)
            bool fwd;
            {
                const double glt = gp[lt];
                const double a1 = vi * wt, a2 = vt * glt * wi, a3 = smin * wi * wt;
                const double mF = a1 - a2 - a3;
                if (fabs(mF) > 64.0 * 2.220446049250313e-16 * (fabs(a1) + fabs(a2) + fabs(a3))) fwd = mF >= 0.0;
                else fwd = vi / wi >= vt / wt * glt + smin;
            }
            if (fwd) {
                v[top] = vt; w[top] = wt; l[top] = lt;
                ++top;
                t[top] = ti;
                vt = vi; wt = wi; lt = li;
                continue;
            }
            vt = vt + vi * gp[lt];
            wt = wt + wi * gp[2 * lt];
            lt = lt + li;
            while (top > 0) {
                double vp = v[top - 1], wp = w[top - 1];
                int lp = l[top - 1];
                // back-track test  vt / wt < max(0, vp / wp * g^lp) + smin, same filter
                bool merge;
                {
                    const double glp = gp[lp];
                    const double pq = vp * glp;                     // sign of vp / wp * g^lp (wp > 0)
                    const double b1 = vt * wp, b2 = (pq > 0.0 ? pq : 0.0) * wt, b3 = smin * wp * wt;
                    const double mB = b1 - b2 - b3;
                    if (fabs(mB) > 64.0 * 2.220446049250313e-16 * (fabs(b1) + fabs(b2) + fabs(b3))) merge = mB < 0.0;
                    else merge = vt / wt < fmax(0.0, vp / wp * glp) + smin;
                }
                if (merge) {
                    vt = vp + vt * gp[lp];
                    wt = wp + wt * gp[2 * lp];
                    lt = lp + lt;
                    --top;
                } else break;
            }
        }
        v[top] = vt; w[top] = wt; l[top] = lt;
        sh->ibc[2] = (n > 0) ? top + 1 : 0;
    }
    __syncthreads();
    int r = sh->ibc[2];
    __syncthreads();
    return r;
}

// oasisAR1.m:101-109
__device__ void block_oasis_ar1_solution(const TraceWS& ws, int n, double g, int T, double* c, double* s) {
    for (int i = threadIdx.x; i < T; i += blockDim.x) s[i] = 0.0;
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int p = warp; p < n; p += nw) {
        double a = fmax(0.0, ws.pv[p] / ws.pw[p]);
        int t0 = ws.pt[p], l = ws.pl[p];
        for (int j = lane; j < l; j += 32) c[t0 + j] = a * ws.gp[j];
    }
    __syncthreads();
    for (int p = 1 + threadIdx.x; p < n; p += blockDim.x) {
        int t0 = ws.pt[p];
        s[t0] = c[t0] - g * c[t0 - 1];
    }
    __syncthreads();
}

// Cold-start scan (every incoming pool is a singleton, oasisAR1.m:46-50) by ONE WARP with exact speculation:
// the next 32 elements are tentatively absorbed into the top pool -- the running (v, w) are accumulated in the
// reference's sequential order -- then all 32 forward / back-track tests are evaluated in parallel and everything up
// to the first event (a new pool is accepted, or a back-track merge is needed) is committed.  Decisions and values are
// identical to the element-by-element loop; only the divisions/table look-ups are parallelised.
__device__ int warp_oasis_ar1_cold(const double* __restrict__ y, int T, double g, double lam, double smin,
                                   TraceWS& ws, double2* stage, double2* snap, const double* gsm, int gsm_n) {
    __syncwarp();   // enter converged (see common.cuh: divergent warps take the slow shuffle path)
    const int lane = threadIdx.x & 31;
    const double pen = lam * (1.0 - g);
    const double* gpg = ws.gp;
    // powers of g: the head of the table from shared memory (gsm, gsm_n entries; pools are rarely longer), the rest from global
    auto gpw = [&](int idx) { return idx < gsm_n ? gsm[idx] : gpg[idx]; };
    double* pv = ws.pv; double* pw = ws.pw; int* pt = ws.pt; int* pl = ws.pl;
    auto val = [&](int idx) { return (idx == T - 1) ? (y[idx] - lam) : (y[idx] - pen); };
    int top = 0, i = 1, lt = 1;
    double vt = val(0), wt = 1.0;
    if (lane == 0) pt[0] = 0;
    // the pool below the top is cached in registers (valid when top > 0): v, w, l, g^l, g^(2l) and its back-track
    // threshold rp = max(0, v/w * g^l) + smin
    double bv = 0.0, bw = 1.0, bg1 = 0.0, bg2 = 0.0, rp = 0.0;
    int bl = 0;
    // operands of the NEXT window under the no-event continuation (i + m, lt + m), loaded while this window's
    // sequential prefix runs
    int pf_i = -1, pf_lt = -1;
    double pf_y = 0.0, pf_e1 = 0.0, pf_e2 = 0.0;
    while (i < T) {
        const int m = min(32, T - i);
        const int idx = i + lane;
        const bool in = lane < m;
        double yj, e1, e2;
        if (pf_i == i && pf_lt == lt) { yj = pf_y; e1 = pf_e1; e2 = pf_e2; }
        else {
            yj = in ? val(idx) : 0.0;
            e1 = in ? gpw(lt + lane) : 0.0;
            e2 = in ? gpw(2 * (lt + lane)) : 0.0;
        }
        {
            const int ni = i + m, nlt = lt + m, nidx = ni + lane;
            const bool nin = nidx < T;
            pf_y = nin ? val(nidx) : 0.0;
            pf_e1 = nin ? gpw(nlt + lane) : 0.0;
            pf_e2 = nin ? gpw(2 * (nlt + lane)) : 0.0;
            pf_i = ni; pf_lt = nlt;
        }
        const double aj = yj * e1;
        // running (v, w) in the reference's sequential order; lanes >= m contribute exact zeros.  The window's operands
        // go through shared memory so every lane reads them back as one broadcast 16-byte load per element, and lane mm
        // stores the running pair it needs (4 instructions per element instead of 14 with shuffles + selects).
        __syncwarp();
        stage[lane] = make_double2(aj, e2);
        __syncwarp();
        double v = vt, w = wt;
#pragma unroll
        for (int mm = 0; mm < 32; ++mm) {
            const double2 ab = stage[mm];
            if (lane == mm) snap[mm] = make_double2(v, w);
            v = v + ab.x;
            w = w + ab.y;
        }
        __syncwarp();
        const double2 own = snap[lane];
        const double vj = own.x, wj = own.y;
        const double vj1 = vj + aj, wj1 = wj + e2;   // the same additions the chain performed for this element
        // The reference's tests are  yj >= vj / wj * e1 + smin  and  vj1 / wj1 < rp  (two divisions, ~110 cycles each, on the
        // chain).  Multiplied through by w > 0 they are  wj (yj - smin) - vj e1 >= 0  and  vj1 - rp wj1 < 0; evaluated that way
        // the sign is certain whenever the margin exceeds the rounding of both forms (a few ulp of the terms), and then it IS
        // the reference's decision.  Only when some lane's margin is inside that band (practically never) does the window
        // fall back to the divisions, so decisions stay identical to the element-by-element loop.
        bool fwd, back;
        {
            const double t1 = wj * (yj - smin), t2 = vj * e1;
            const double mF = t1 - t2, bF = 64.0 * 2.220446049250313e-16 * (fabs(t1) + fabs(t2) + fabs(wj * smin));
            const double t3 = rp * wj1;
            const double mB = vj1 - t3, bB = 64.0 * 2.220446049250313e-16 * (fabs(vj1) + fabs(t3));
            const bool unsureF = in && !(fabs(mF) > bF);
            fwd = in && (mF >= 0.0);
            const bool unsureB = in && !fwd && (top > 0) && !(fabs(mB) > bB);
            back = in && !fwd && (top > 0) && (mB < 0.0);
            if (__any_sync(0xffffffffu, unsureF || unsureB)) {
                fwd = in && (yj >= vj / wj * e1 + smin);
                back = in && !fwd && (top > 0) && (vj1 / wj1 < rp);
            }
        }
        const unsigned mf = __ballot_sync(0xffffffffu, fwd), mb = __ballot_sync(0xffffffffu, back);
        const unsigned ev = mf | mb;
        if (ev == 0u) {
            vt = v; wt = w; lt += m; i += m;
            continue;
        }
        const int e = __ffs(ev) - 1;
        if (mf & (1u << e)) {
            // elements 0..e-1 absorbed, element e starts a new pool; the finished pool becomes the cached one below
            bv = __shfl_sync(0xffffffffu, vj, e);
            bw = __shfl_sync(0xffffffffu, wj, e);
            bg1 = __shfl_sync(0xffffffffu, e1, e);      // g^(lt + e)
            bg2 = __shfl_sync(0xffffffffu, e2, e);      // g^(2 (lt + e))
            bl = lt + e;
            if (lane == 0) { pv[top] = bv; pw[top] = bw; pl[top] = bl; pt[top + 1] = i + e; }
            rp = fmax(0.0, bv / bw * bg1) + smin;
            ++top;
            vt = __shfl_sync(0xffffffffu, yj, e);
            wt = 1.0; lt = 1;
            i += e + 1;
        } else {
            // elements 0..e absorbed, then back-track (oasisAR1.m:82-95)
            vt = __shfl_sync(0xffffffffu, vj1, e);
            wt = __shfl_sync(0xffffffffu, wj1, e);
            lt += e + 1;
            i += e + 1;
            __syncwarp();
            while (top > 0) {
                if (vt / wt < rp) {
                    vt = bv + vt * bg1;
                    wt = bw + wt * bg2;
                    lt = bl + lt;
                    --top;
                    if (top > 0) {
                        bv = pv[top - 1]; bw = pw[top - 1]; bl = pl[top - 1];
                        bg1 = gpw(bl); bg2 = gpw(2 * bl);
                        rp = fmax(0.0, bv / bw * bg1) + smin;
                    }
                } else break;
            }
        }
        __syncwarp();
    }
    if (lane == 0) { pv[top] = vt; pw[top] = wt; pl[top] = lt; }
    __syncwarp();
    return top + 1;
}

// Cold oasisAR1(y, g, lam, smin): pools + solution into ws.c / ws.s.  Returns pool count.
__device__ int block_oasis_ar1(const double* y, int T, double g, double lam, double smin, TraceWS& ws,
                               BlockShared* sh) {
    CNMFE_PROF(sh, 10);
    block_pow_table(g, T, ws.gp, sh);
    __syncthreads();
    CNMFE_PROF(sh, 11);
    // the scan is one warp's dependency chain: its operands come from shared memory when the launch staged it (the trace
    // into the update_g buffer, the head of the power table into the h / hh tables -- both idle during the scan)
    const double* ysc = y;
    const double* gsm = nullptr;
    int gsm_n = 0;
    if (sh->ysm) {
        for (int i = threadIdx.x; i < T; i += blockDim.x) sh->ysm[i] = y[i];
        gsm_n = min(2 * TRACE_HCAP, 2 * T + 2);
        for (int i = threadIdx.x; i < gsm_n; i += blockDim.x) sh->hsm[i] = ws.gp[i];
        ysc = sh->ysm; gsm = sh->hsm;
        __syncthreads();
    }
    if (warp_id_uniform() == 0) {
        int n = warp_oasis_ar1_cold(ysc, T, g, lam, smin, ws, sh->stage, sh->snap, gsm, gsm_n);
        if (threadIdx.x == 0) sh->ibc[2] = n;
    }
    __syncthreads();
    const int n = sh->ibc[2];
    __syncthreads();
    CNMFE_PROF(sh, 5);
    block_oasis_ar1_solution(ws, n, g, T, ws.c, ws.s);
    CNMFE_PROF(sh, 6);
    return n;
}

#define RSS_SHORT_POOL 256
// Per-pool energy Q_p = sum_{t in pool} y_t^2 -> ws.sv (free during update_g).  Pools are fixed while fminbnd runs, so this
// is computed once per update_g; rss_g then needs ONE pass over the trace per evaluation.
__device__ void block_pool_energy(const double* y, int n, TraceWS& ws, BlockShared* sh) {
    if (sh->ysm) y = sh->ysm;
    const bool ps = (n <= sh->pcap);
    const int* const ptab = ps ? sh->ptsm : ws.pt;
    const int* const ltab = ps ? sh->plsm : ws.pl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int grp = lane >> 3, gl = lane & 7;
    for (int p0 = warp * 4; p0 < n; p0 += nw * 4) {
        const int p = p0 + grp;
        int l = 0, t0 = 0;
        if (p < n) { l = ltab[p]; t0 = ptab[p]; }
        const int le = (l <= RSS_SHORT_POOL) ? l : 0;
        double q = 0.0;
        for (int j = gl; j < le; j += 8) { const double v = y[t0 + j]; q = fma(v, v, q); }
        __syncwarp();
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        if (le > 0 && gl == 0) ws.sv[p] = q;
    }
    for (int p = warp; p < n; p += nw) {
        const int t0 = ptab[p], l = ltab[p];
        if (l <= RSS_SHORT_POOL) continue;
        double q = 0.0;
        for (int j = lane; j < l; j += 32) { const double v = y[t0 + j]; q = fma(v, v, q); }
        q = warp_sum(q);
        if (lane == 0) ws.sv[p] = q;
    }
    __syncthreads();
}

// Work split of rss_g: thread i owns the samples [ns i, ns (i+1)) of the trace (ns odd: the lanes of a warp then read
// shared memory 8 ns bytes apart without bank conflicts).  The pools are fixed while fminbnd runs, so the pool that
// contains a thread's first sample is found once per update_g.
struct RssPlan { int ns, p_start; };
__device__ RssPlan block_rss_plan(int T, int n, const int* ptab) {
    RssPlan pl;
    int ns = (T + (int)blockDim.x - 1) / (int)blockDim.x;
    if (!(ns & 1)) ++ns;
    pl.ns = ns;
    const int tA = ns * (int)threadIdx.x;
    int lo = 0, hi = n - 1;                    // largest p with ptab[p] <= tA
    if (tA < T) {
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (ptab[mid] <= tA) lo = mid; else hi = mid - 1; }
    }
    pl.p_start = lo;
    return pl;
}

// rss_g of update_g (foopsi_oasisAR1.m:165-178).  Leaves ws.h / ws.hh holding this g's tables.
// With h = g^(0..l-1), dy = sum y h over a pool and sh = sum h = (1 - g^l) / (1 - g):
//   dot = yp' h = dy - pen sh,  tv = max(dot / hh(l), 0),  sum (y - tv h)^2 = Q - tv (2 dy - tv hh(l))
// (the reference forms c = tv h and then res = y - c: the same number, summed in another order; Q from block_pool_energy).
// Two passes, no per-pool loops on the critical path: (1) every thread walks its ns consecutive samples and writes one partial
// dot product per (thread chunk, pool) pair into slot chunk + pool -- both indices only grow along the trace, so the slot is
// unique and a pool's partials are consecutive; (2) one thread per pool adds its partials in ascending order (deterministic)
// and forms the pool's term.  Per evaluation: ~ns + maxl/ns sequential steps instead of the longest pool's length.
__device__ double block_rss_g(const double* y, int T, int n, double g, double lam, int maxl, const RssPlan& plan, TraceWS& ws,
                              BlockShared* sh) {
    const double lg = log(g), pen = lam * (1.0 - g);
    __syncthreads();
    CNMFE_PROF(sh, 7);
    if (sh->ysm) y = sh->ysm;                                    // staged by block_update_g
    const bool hs = (maxl < sh->hcap), ps = (n <= sh->pcap);
    double* const htab = hs ? sh->hsm : ws.h;
    double* const hhtab = hs ? sh->hhsm : ws.hh;
    const int* const ptab = ps ? sh->ptsm : ws.pt;
    const int* const ltab = ps ? sh->plsm : ws.pl;
    const double* const Q = ws.sv;
    // <= chunks + n partial sums: in shared memory while they fit -- through the global scratch every evaluation paid a
    // store -> barrier -> load round trip to L2 (~2 k cycles)
    const int nchunks = (T + plan.ns - 1) / plan.ns;
    const bool slots_sm = nchunks + n <= TRACE_PART;
    double* const slots = ws.scr;
    double* const slots_s = sh->part;
    if (sh->prof) { if (threadIdx.x == 0) { sh->pc[20] += 1ull; sh->pc[21] += (unsigned long long)maxl; sh->pc[22] += (unsigned long long)n; } __syncwarp(); }
    // h = g^(0..maxl) and hh = cumsum(h.^2) in closed form, hh(j) = (1 - g^(2 (j+1))) / (1 - g^2): the geometric sum the
    // reference accumulates numerically (within ~1e-14 relative of the sequential cumsum for g <= 0.999; no scan, no barrier)
    {
        const double g2 = g * g, inv = 1.0 / (1.0 - g2);
        for (int j = threadIdx.x; j <= maxl; j += blockDim.x) {
            const double hv = exp(lg * (double)j);
            htab[j] = hv;
            hhtab[j] = fma(-g2 * hv, hv, 1.0) * inv;
        }
    }
    __syncthreads();
    CNMFE_PROF(sh, 16);
    // (1) partial dot products along the trace: the thread's samples split at the pool borders into segments, each a
    //     branch-free dot product
    {
        const int ns = plan.ns, c = (int)threadIdx.x;
        int t = ns * c;
        if (t < T) {
            const int te = min(T, t + ns);
            int p = plan.p_start;
            int t0 = ptab[p], tend = t0 + ltab[p];
            while (true) {
                const int se = min(te, tend);
                const double* yy = y + t;
                const double* hh_ = htab + (t - t0);
                const int len = se - t;
                double acc = 0.0;
                int k = 0;
                for (; k + 4 <= len; k += 4) {
                    const double y0 = yy[k], y1 = yy[k + 1], y2 = yy[k + 2], y3 = yy[k + 3];
                    const double h0 = hh_[k], h1 = hh_[k + 1], h2 = hh_[k + 2], h3 = hh_[k + 3];
                    acc = fma(y0, h0, acc); acc = fma(y1, h1, acc); acc = fma(y2, h2, acc); acc = fma(y3, h3, acc);
                }
                for (; k < len; ++k) acc = fma(yy[k], hh_[k], acc);
                if (slots_sm) slots_s[c + p] = acc; else slots[c + p] = acc;
                t = se;
                if (t >= te) break;
                ++p; t0 = tend; tend = t0 + ltab[p];
            }
        }
    }
    __syncthreads();
    CNMFE_PROF(sh, 17);
    // (2) four lanes per pool: each adds every fourth partial in ascending order, then a fixed two-level shuffle tree
    double rss = 0.0;
    {
        const int ns = plan.ns;
        const double inv1g = 1.0 / (1.0 - g);
        const int q = (int)threadIdx.x & 3;
        for (int pb = 0; pb < n; pb += (int)blockDim.x >> 2) {
            const int p = pb + ((int)threadIdx.x >> 2);
            double dy = 0.0;
            int l = 1;
            if (p < n) {
                const int t0 = ptab[p];
                l = ltab[p];
                const int c0 = t0 / ns, c1 = (t0 + l - 1) / ns;
                if (slots_sm) { for (int c = c0 + q; c <= c1; c += 4) dy += slots_s[c + p]; }
                else { for (int c = c0 + q; c <= c1; c += 4) dy += slots[c + p]; }
            }
            dy += __shfl_xor_sync(0xffffffffu, dy, 1);
            dy += __shfl_xor_sync(0xffffffffu, dy, 2);
            if (p < n && q == 0) {
                const double shs = (1.0 - g * htab[l - 1]) * inv1g;
                const double hhl = hhtab[l - 1];
                const double tv = fmax((dy - pen * shs) / hhl, 0.0);
                rss += Q[p] - tv * (2.0 * dy - tv * hhl);
            }
        }
    }
    CNMFE_PROF(sh, 24);
    CNMFE_PROF(sh, 18);
    rss = block_sum(rss, sh->red);
    CNMFE_PROF(sh, 19);
    return rss;
}

// MATLAB fminbnd on rss_g over [ax,bx]; returns xf (all threads run the scalar logic redundantly).
__device__ double block_fminbnd_rss(const double* y, int T, int n, double lam, int maxl, double ax, double bx,
                                    const RssPlan& plan, TraceWS& ws, BlockShared* sh) {
    const double tol = 1e-4, seps = 1.4901161193847656e-08, cgold = 0.3819660112501051;
    double a = ax, b = bx;
    double v = a + cgold * (b - a), w = v, xf = v, d = 0.0, e = 0.0, x = xf;
    double fx = block_rss_g(y, T, n, x, lam, maxl, plan, ws, sh);
    int funccount = 1, iter = 0;
    double fv = fx, fw = fx;
    double xm = 0.5 * (a + b);
    double tol1 = seps * fabs(xf) + tol / 3.0, tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        bool gs = true;
        if (fabs(e) > tol1) {
            gs = false;
            double r = (xf - w) * (fx - fv);
            double q = (xf - v) * (fx - fw);
            double p = (xf - v) * q - (xf - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = d;
            if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
                d = p / q;
                x = xf + d;
                if (((x - a) < tol2) || ((b - x) < tol2)) {
                    double df = xm - xf;
                    double si = (df > 0.0 ? 1.0 : (df < 0.0 ? -1.0 : 0.0)) + (df == 0.0 ? 1.0 : 0.0);
                    d = tol1 * si;
                }
            } else {
                gs = true;
            }
        }
        if (gs) {
            e = (xf >= xm) ? (a - xf) : (b - xf);
            d = cgold * e;
        }
        double si = (d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0)) + (d == 0.0 ? 1.0 : 0.0);
        x = xf + si * fmax(fabs(d), tol1);
        double fu = block_rss_g(y, T, n, x, lam, maxl, plan, ws, sh);
        ++funccount; ++iter;
        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            v = w; fv = fw;
            w = xf; fw = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if ((fu <= fw) || (w == xf)) {
                v = w; fv = fw;
                w = x; fw = fu;
            } else if ((fu <= fv) || (v == xf) || (v == w)) {
                v = x; fv = fu;
            }
        }
        xm = 0.5 * (a + b);
        tol1 = seps * fabs(xf) + tol / 3.0;
        tol2 = 2.0 * tol1;
        if (funccount >= 500 || iter >= 500) break;
    }
    return xf;
}

// update_g (foopsi_oasisAR1.m:124-163): returns new g; pools/solution updated (warm oasisAR1).
__device__ double block_update_g(const double* y, int T, int* n_io, double lam, double smin, double g_lo,
                                 double g_hi, TraceWS& ws, BlockShared* sh) {
    int n = *n_io;
    double ml = 0.0;
    for (int p = threadIdx.x; p < n; p += blockDim.x) ml = fmax(ml, (double)ws.pl[p]);
    int maxl = (int)block_max(ml, sh->red);
    if (sh->ysm) {
        for (int i = threadIdx.x; i < T; i += blockDim.x) sh->ysm[i] = y[i];
        if (n <= sh->pcap)
            for (int p = threadIdx.x; p < n; p += blockDim.x) { sh->ptsm[p] = ws.pt[p]; sh->plsm[p] = ws.pl[p]; }
        __syncthreads();
    }
    if (!sh->ysm) __syncthreads();
    block_pool_energy(y, n, ws, sh);       // Q_p of the (fixed) pools: rss_g is then one pass per evaluation
    const double* const hh_last = (maxl < sh->hcap) ? sh->hhsm : ws.hh;   // where the last rss_g evaluation left cumsum(h.^2)
    CNMFE_PROF(sh, 10);
    const RssPlan plan = block_rss_plan(T, n, (n <= sh->pcap && sh->ysm) ? sh->ptsm : ws.pt);
    double g = block_fminbnd_rss(y, T, n, lam, maxl, g_lo, g_hi, plan, ws, sh);
    CNMFE_PROF(sh, 7);
    // rebuild pools: v from the returned g, w from the LAST evaluated kernel (ws.hh), foopsi_oasisAR1.m:153-162
    const double lg = log(g), pen = lam * (1.0 - g);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int p = warp; p < n; p += nw) {
        int t0 = ws.pt[p], l = ws.pl[p];
        double dot = 0.0;
        for (int j = lane; j < l; j += 32) dot += (y[t0 + j] - pen) * exp(lg * (double)j);
        dot = warp_sum(dot);
        if (lane == 0) { ws.pv[p] = dot; ws.pw[p] = hh_last[l - 1]; }
    }
    __syncthreads();
    CNMFE_PROF(sh, 8);
    block_pow_table(g, T, ws.gp, sh);
    CNMFE_PROF(sh, 11);
    n = block_oasis_ar1_run(ws, n, smin, sh);
    CNMFE_PROF(sh, 9);
    block_oasis_ar1_solution(ws, n, g, T, ws.c, ws.s);
    CNMFE_PROF(sh, 6);
    *n_io = n;
    return g;
}

struct DeconvOut {
    double b, g1, g2, smin, lam, sn;
    int npars;   // 1 or 2; 0 => failed (outputs are zeros)
};

// foopsi_oasisAR1(y, g, lam, smin, optimize_b, optimize_g, [], maxIter, tau_range, gmax)  (foopsi_oasisAR1.m:78-122)
// y: input trace (read-only).  Outputs in ws.c / ws.s.
__device__ void block_foopsi_ar1(const double* __restrict__ y, int T, double g, double lam, double smin,
                                 bool optimize_b, bool optimize_g, int maxIter, double g_lo, double g_hi,
                                 bool has_tau_range, double gmax, TraceWS& ws, BlockShared* sh, DeconvOut* out) {
    if (has_tau_range) g = fmin(fmax(g, g_lo), g_hi);
    double b = 0.0;
    int n;
    if (!optimize_b) {
        n = block_oasis_ar1(y, T, g, lam, smin, ws, sh);
        if (optimize_g && n > 0) g = block_update_g(y, T, &n, lam, smin, g_lo, g_hi, ws, sh);
    } else {
        CNMFE_PROF(sh, 10);
        b = block_quantile(y, T, 0.15, sh);
        CNMFE_PROF(sh, 4);
        for (int i = threadIdx.x; i < T; i += blockDim.x) ws.yb[i] = y[i] - b;
        __syncthreads();
        n = block_oasis_ar1(ws.yb, T, g, lam, smin, ws, sh);
        for (int m = 0; m < maxIter; ++m) {
            if (sh->prof) { if (threadIdx.x == 0) sh->pc[15] += 1ull; __syncwarp(); }
            double a = 0.0;
            for (int i = threadIdx.x; i < T; i += blockDim.x) a += y[i] - ws.c[i];
            b = block_sum(a, sh->red) / (double)T;
            if (!optimize_g) break;
            if (n == 0) break;
            double g0 = g;
            for (int i = threadIdx.x; i < T; i += blockDim.x) ws.yb[i] = y[i] - b;
            __syncthreads();
            if (g > gmax) {
                // spike counts too small (foopsi_oasisAR1.m:104-108): g = estimate_time_constant(y,1) (sn from GetSn(y))
                double sn = block_getsn(y, T, ws.scr, sh);
                double gg[2];
                int np = block_time_constant(y, T, 1, sn, gg, sh);
                if (np != 1) {   // oasisAR1 with length(g)>1 returns zeros (oasisAR1.m:40-45)
                    for (int i = threadIdx.x; i < T; i += blockDim.x) { ws.c[i] = 0.0; ws.s[i] = 0.0; }
                    __syncthreads();
                    out->npars = 2;   // reference would carry a vector g; flagged to the caller
                    g = 0.0;
                } else {
                    g = gg[0];
                    n = block_oasis_ar1(ws.yb, T, g, lam, smin, ws, sh);
                }
                break;
            }
            g = block_update_g(ws.yb, T, &n, lam, smin, g_lo, g_hi, ws, sh);
            if (fabs(g - g0) / g0 < 1e-3) optimize_g = false;
        }
    }
    out->b = b;
    out->g1 = g;
    out->g2 = 0.0;
}

// ------------------------------------------------------------------------------------------------ AR(2) PAV
// oasisAR2.m:49-156 (cold start).  Tables: h=g11, hh=g12, gp[0..T)=g11g11, gp[T..2T)=g11g12.  yp in ws.yb.
// Warp 0 runs the scan (lanes share the merge dot product); outputs ws.c / ws.s; returns pool count.
__device__ int block_oasis_ar2(const double* __restrict__ y, int T, double g1, double g2, double lam, double smin,
                               TraceWS& ws, BlockShared* sh) {
    const double disc = g1 * g1 + 4.0 * g2;
    const double sq = sqrt(fmax(disc, 0.0));
    const double dd = 0.5 * (g1 + sq), rr = 0.5 * (g1 - sq);
    const double ld = log(dd), lr = log(rr);
    double* g11 = ws.h; double* g12 = ws.hh; double* g11g11 = ws.gp; double* g11g12 = ws.gp + T;
    double* yp = ws.yb;
    __syncthreads();
    for (int k = threadIdx.x; k < T; k += blockDim.x) {
        g11[k] = (exp(ld * (double)(k + 1)) - exp(lr * (double)(k + 1))) / (dd - rr);
        double v = y[k] - lam * (1.0 - g1 - g2);
        if (k == T - 2) v = y[k] - lam * (1.0 - g1);
        if (k == T - 1) v = y[k] - lam;
        yp[k] = v;
        ws.pv[k] = v; ws.pw[k] = v; ws.pt[k] = k; ws.pl[k] = 1;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < T; k += blockDim.x) g12[k] = (k == 0) ? 0.0 : g2 * g11[k - 1];
    __syncthreads();
    if (threadIdx.x == 0) {   // sequential cumsums (reference order, oasisAR2.m:75-76)
        double a1 = 0.0, a2 = 0.0;
        for (int k = 0; k < T; ++k) {
            a1 += g11[k] * g11[k]; g11g11[k] = a1;
            a2 += g11[k] * g12[k]; g11g12[k] = a2;
        }
    }
    __syncthreads();
    if (warp_id_uniform() == 0) {
        __syncwarp();   // enter converged
        const int lane = threadIdx.x;
        double* v = ws.pv; double* w = ws.pw; int* t = ws.pt; int* l = ws.pl;
        int top = 0;   // stack [0..top]; pool 0 is never merged (scan starts at the 2nd pool)
        if (T >= 3) {
            top = 1;
            for (int i = 2; i < T; ++i) {
                double vi = v[i];
                int li = l[i];
                // forward test against stack top (oasisAR2.m:83-84)
                if (g11[l[top]] * v[top] + g12[l[top]] * w[top - 1] + smin <= vi) {
                    ++top;
                    if (lane == 0) { v[top] = vi; w[top] = w[i]; t[top] = t[i]; l[top] = li; }
                    __syncwarp();
                    continue;
                }
                // merge incoming pool i into top (oasisAR2.m:93-104)
                int lnew = l[top] + li, ti = t[top];
                double dot = 0.0;
                for (int j = lane; j < lnew; j += 32) dot += g11[j] * yp[ti + j];
                dot = warp_sum(dot);
                double vn = (dot - g11g12[lnew - 1] * w[top - 1]) / g11g11[lnew - 1];
                double wn = g11[lnew - 1] * vn + g12[lnew - 1] * w[top - 1];
                __syncwarp();
                if (lane == 0) { l[top] = lnew; v[top] = vn; w[top] = wn; }
                __syncwarp();
                // backtrack (oasisAR2.m:109-128): needs prev-prev
                while (top >= 2 &&
                       (g11[l[top - 1]] * v[top - 1] + g12[l[top - 1]] * w[top - 2] + smin > v[top])) {
                    int lm = l[top - 1] + l[top], tm = t[top - 1];
                    double d2 = 0.0;
                    for (int j = lane; j < lm; j += 32) d2 += g11[j] * yp[tm + j];
                    d2 = warp_sum(d2);
                    double v2 = (d2 - g11g12[lm - 1] * w[top - 2]) / g11g11[lm - 1];
                    double w2 = g11[lm - 1] * v2 + g12[lm - 1] * w[top - 2];
                    __syncwarp();
                    if (lane == 0) { l[top - 1] = lm; v[top - 1] = v2; w[top - 1] = w2; }
                    --top;
                    __syncwarp();
                }
            }
        } else {
            top = T - 1;
        }
        if (lane == 0) sh->ibc[2] = top + 1;
    }
    __syncthreads();
    const int n = sh->ibc[2];
    // construct solution (oasisAR2.m:140-156): AR(2) recursion inside each pool needs the two previous samples,
    // which belong to the previous pool -> sequential over time; thread 0.
    if (threadIdx.x == 0) {
        double c1 = 0.0, c2 = 0.0;   // c(t-1), c(t-2) before clamping
        for (int p = 0; p < n; ++p) {
            int ti = ws.pt[p], li = ws.pl[p];
            double cur = ws.pv[p];
            ws.c[ti] = cur;
            c2 = c1; c1 = cur;
            for (int j = 1; j < li; ++j) {
                cur = g1 * c1 + g2 * c2;
                ws.c[ti + j] = cur;
                c2 = c1; c1 = cur;
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < T; k += blockDim.x) if (ws.c[k] < 0.0) ws.c[k] = 0.0;
    __syncthreads();
    for (int k = threadIdx.x; k < T; k += blockDim.x) {
        double sv = 0.0;
        if (k >= 3) {
            sv = ws.c[k] - g1 * ws.c[k - 1] - g2 * ws.c[k - 2];
            if (sv < smin) sv = 0.0;
        }
        ws.s[k] = sv;
    }
    __syncthreads();
    return n;
}

// max_ht(pars) for AR(2)  (functions/max_ht.m:11-29)
__device__ double ar2_max_ht(double g1, double g2) {
    double sq = sqrt(fmax(g1 * g1 + 4.0 * g2, 0.0));
    double d = 0.5 * (g1 + sq), r = 0.5 * (g1 - sq);
    double tau_d = -1.0 / log(d), tau_r = -1.0 / log(r);
    int n = (int)ceil(tau_d * 2.0);
    double dd = exp(-1.0 / tau_d), rr = exp(-1.0 / tau_r);
    double ld = log(dd), lr = log(rr), vmax = -INFINITY;
    for (int t = 1; t <= n; ++t) {
        double ht = (exp(ld * (double)t) - exp(lr * (double)t)) / (dd - rr);
        vmax = fmax(vmax, ht);
    }
    return vmax;
}

}  // namespace cnmfe
