// oasis_methods.cuh -- constrained / thresholded AR(1) drivers and the deconvolveCa dispatcher (device side).
// Reference: OASIS_matlab/packages/oasis/constrained_oasisAR1.m:84-199, thresholded_oasisAR1.m:78-213,
// OASIS_matlab/deconvolveCa.m:60-206, functions/choose_smin.m:28-42.
#pragma once
#include "oasis.cuh"
#include "../../include/cnmfe_b200.h"

namespace cnmfe {

__device__ __forceinline__ double block_rss(const double* y, const double* c, double b, int T, BlockShared* sh) {
    double a = 0.0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double r = y[i] - c[i] - b;
        a += r * r;
    }
    return block_sum(a, sh->red);
}

// update_phi (constrained_oasisAR1.m:151-187).  res = y - c - b is recomputed here.  Returns dphi.
__device__ double block_update_phi(const double* y, int T, int* n_io, double g, double lam, double b,
                                   bool optimize_b, double thresh, TraceWS& ws, BlockShared* sh) {
    int n = *n_io;
    double* zeta = ws.scr;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    for (int p = warp; p < n; p += nw) {
        int t0 = ws.pt[p], l = ws.pl[p];
        double f = (p < n - 1) ? (1.0 - ws.gp[l]) / ws.pw[p] : 1.0 / ws.pw[p];
        for (int j = lane; j < l; j += 32) zeta[t0 + j] = f * ws.gp[j];
    }
    __syncthreads();
    double zm = 0.0, rm = 0.0;
    if (optimize_b) {
        double a = 0.0, r = 0.0;
        for (int i = threadIdx.x; i < T; i += blockDim.x) { a += zeta[i]; r += y[i] - ws.c[i] - b; }
        zm = block_sum(a, sh->red) / (double)T;
        rm = block_sum(r, sh->red) / (double)T;
    }
    double aa = 0.0, bb = 0.0, cc = 0.0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double z = zeta[i] - zm, r = (y[i] - ws.c[i] - b) - rm;
        aa += z * z; bb += r * z; cc += r * r;
    }
    aa = block_sum(aa, sh->red);
    bb = block_sum(bb, sh->red);
    cc = block_sum(cc, sh->red) - thresh;
    double disc = bb * bb - aa * cc;
    if (disc < 0.0) {
        // complex dphi: the reference returns before touching the pools when imag(dphi) > 1e-9
        if (sqrt(-disc) / aa > 1e-9) return -bb / aa;
        disc = 0.0;
    }
    double dphi = (-bb + sqrt(disc)) / aa;
    for (int p = threadIdx.x; p < n; p += blockDim.x) ws.pv[p] = ws.pv[p] - dphi * (1.0 - ws.gp[ws.pl[p]]);
    __syncthreads();
    n = block_oasis_ar1_run(ws, n, 0.0, sh);
    block_oasis_ar1_solution(ws, n, g, T, ws.c, ws.s);
    *n_io = n;
    return dphi;
}

// update_lam_b (constrained_oasisAR1.m:189-199)
__device__ void block_update_lam_b(const double* y, int T, int n, double g, double* lam, double* b, TraceWS& ws,
                                   BlockShared* sh) {
    double a = 0.0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) a += y[i] - ws.c[i];
    double db = block_sum(a, sh->red) / (double)T - *b;
    *b = *b + db;
    double dlam = -db / (1.0 - g);
    *lam = fmax(0.0, *lam + dlam);
    __syncthreads();
    if (n > 0) {
        int l = ws.pl[n - 1], t0 = ws.pt[n - 1];
        double v = ws.pv[n - 1] - (*lam) * ws.gp[l];
        double w = ws.pw[n - 1];
        __syncthreads();
        if (threadIdx.x == 0) ws.pv[n - 1] = v;
        double amp = fmax(0.0, v / w);
        for (int j = threadIdx.x; j < l; j += blockDim.x) ws.c[t0 + j] = amp * ws.gp[j];
    }
    __syncthreads();
}

__device__ void block_constrained_ar1(const double* __restrict__ y, int T, double g, double sn, bool optimize_b,
                                      bool optimize_g, int maxIter, double g_lo, double g_hi, bool has_tau_range,
                                      TraceWS& ws, BlockShared* sh, DeconvOut* out) {
    if (has_tau_range) g = fmin(fmax(g, g_lo), g_hi);
    const double thresh = sn * sn * (double)T, tol = 1e-4;
    double lam = 0.0, b = 0.0;
    bool g_conv = false;
    int n;
    if (!optimize_b) {
        n = block_oasis_ar1(y, T, g, lam, 0.0, ws, sh);
        for (int it = 0; it < maxIter; ++it) {
            if (optimize_g && !g_conv && n > 0) {
                double g0 = g;
                g = block_update_g(y, T, &n, lam, 0.0, g_lo, g_hi, ws, sh);
                if (fabs(g - g0) / g0 < 1e-3) g_conv = true;
            }
            double RSS = block_rss(y, ws.c, 0.0, T, sh);
            if (RSS > thresh || n == 0) break;
            double dphi = block_update_phi(y, T, &n, g, lam, 0.0, false, thresh, ws, sh);
            lam = lam + dphi;
        }
    } else {
        b = block_quantile(y, T, 0.15, sh);
        for (int i = threadIdx.x; i < T; i += blockDim.x) ws.yb[i] = y[i] - b;
        __syncthreads();
        n = block_oasis_ar1(ws.yb, T, g, lam, 0.0, ws, sh);
        block_update_lam_b(y, T, n, g, &lam, &b, ws, sh);
        for (int it = 0; it < maxIter; ++it) {
            double RSS = block_rss(y, ws.c, b, T, sh);
            double sc = 0.0;
            for (int i = threadIdx.x; i < T; i += blockDim.x) sc += ws.c[i];
            sc = block_sum(sc, sh->red);
            if (fabs(RSS - thresh) < tol || sc < 1e-9 || n == 0) break;
            block_update_phi(y, T, &n, g, lam, b, true, thresh, ws, sh);
            block_update_lam_b(y, T, n, g, &lam, &b, ws, sh);
            if (optimize_g && !g_conv && n > 0) {
                double g0 = g;
                for (int i = threadIdx.x; i < T; i += blockDim.x) ws.yb[i] = y[i] - b;
                __syncthreads();
                g = block_update_g(ws.yb, T, &n, lam, 0.0, g_lo, g_hi, ws, sh);
                if (fabs(g - g0) / g0 < 1e-4) g_conv = true;
            }
        }
    }
    out->b = b; out->g1 = g; out->g2 = 0.0; out->lam = lam;
}

// choose_smin(g, sn, prob) for a scalar AR(1) coefficient (choose_smin.m:38-42)
__device__ double choose_smin_ar1(double g, double sn, double prob) {
    double h = 1.0, nrm = 0.0;
    for (int k = 0; k < 1000; ++k) { nrm += h * h; h *= g; }   // filter(1,[1,-g],impulse): h(k) = g*h(k-1)
    return sn / sqrt(nrm) * normcdfinv(prob);
}

// choose_smin for an AR(2) pair: h = filter(1,[1,-g1,-g2],impulse(1000))  (choose_smin.m:38-42)
__device__ double choose_smin_ar2(double g1, double g2, double sn, double prob) {
    double h1 = 0.0, h2 = 0.0, nrm = 0.0;
    for (int k = 0; k < 1000; ++k) {
        double h = (k == 0 ? 1.0 : 0.0) + g1 * h1 + g2 * h2;
        nrm += h * h;
        h2 = h1; h1 = h;
    }
    return sn / sqrt(nrm) * normcdfinv(prob);
}

// [b, sn] = estimate_baseline_noise(y) (OASIS_matlab/functions/estimate_baseline_noise.m:1-40): histogram of y on a grid derived
// from the deciles (hist.m semantics: bin centres, edges half-way, (edge_k, edge_k+1] bins, ends extended to min / max), then
// fit_gauss1(bins, nums, 0.3, 3) (functions/fit_gauss1.m: Guo's iteratively re-weighted log-parabola fit, 3 x 3 systems solved
// by Gaussian elimination with partial pivoting like mldivide).  Counts in ws.st (ints); the fit is serial work for one thread.
__device__ void block_estimate_baseline_noise(const double* __restrict__ y, int T, TraceWS& ws, BlockShared* sh, double* b_out,
                                              double* sn_out) {
    double temp[11];
    for (int i = 0; i <= 10; ++i) temp[i] = block_quantile(y, T, (double)i / 10.0, sh);
    double mind = INFINITY;
    for (int i = 0; i < 10; ++i) mind = fmin(mind, temp[i + 1] - temp[i]);
    const double dbin = fmax(mind / 3.0, (temp[10] - temp[0]) / 1000.0);
    const int nb = dbin > 0.0 ? (int)floor((temp[10] - temp[0]) / dbin + 1e-10) + 1 : 0;
    if (nb <= 0) {     // isempty(bins): b = mean(y), sn = 0
        double a = 0.0;
        for (int i = threadIdx.x; i < T; i += blockDim.x) a += y[i];
        *b_out = block_sum(a, sh->red) / (double)T; *sn_out = 0.0;
        return;
    }
    int* cnt = reinterpret_cast<int*>(ws.scr);          // nb <= 1001 ints; the scratch holds >= 3 * nfft (>= 768) doubles
    const double c0 = temp[0], ymin = temp[0], ymax = temp[10];
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    // shifted edge j (0..nb): first = min(c0 - dbin/2, min y), last = max(c_{nb-1}, max y), else c_{j-1} + (c_j - c_{j-1})/2
    auto edge = [&](int j) -> double {
        double e;
        if (j == 0) { const double w0 = (nb > 1) ? ((c0 + dbin * 1.0) - c0) : 0.0; e = fmin(c0 - w0 / 2.0, ymin); }
        else if (j == nb) e = fmax(c0 + dbin * (double)(nb - 1), ymax);
        else { const double ca = c0 + dbin * (double)(j - 1), cb = c0 + dbin * (double)j; e = ca + (cb - ca) / 2.0; }
        const double ae = fabs(e);                                   // xx + eps(xx): spacing to the next double above |xx|
        return e + (__longlong_as_double(__double_as_longlong(ae) + 1) - ae);
    };
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const double v = y[i];
        int k = (int)floor((v - c0) / dbin + 0.5);                  // nearest centre, then exact edge tests
        k = min(max(k, 0), nb - 1);
        while (k > 0 && v < edge(k)) --k;
        while (k < nb - 1 && v >= edge(k + 1)) ++k;
        atomicAdd(&cnt[k], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int mx = 0;
        for (int k = 0; k < nb; ++k) mx = max(mx, cnt[k]);
        const double thr = 0.3 * (double)mx;
        double p[3] = {0.0, 0.0, 0.0};
        for (int it = 0; it < 3; ++it) {
            double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, rhs[3] = {0, 0, 0};
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
            for (int k = 0; k < nb; ++k) {
                const double n0 = (double)cnt[k];
                if (!(n0 > thr)) continue;
                const double x = c0 + dbin * (double)k, x2 = x * x;
                const double ly = (it == 0) ? log(n0) : (p[0] + p[1] * x + p[2] * x2);
                const double yv = (it == 0) ? n0 : exp(ly), y2 = yv * yv, y2l = y2 * ly;
                s0 += y2; s1 += x * y2; s2 += x2 * y2; s3 += x2 * x * y2; s4 += x2 * x2 * y2;
                rhs[0] += y2l; rhs[1] += x * y2l; rhs[2] += x2 * y2l;
            }
            M[0][0] = s0; M[0][1] = s1; M[0][2] = s2; M[1][0] = s1; M[1][1] = s2; M[1][2] = s3; M[2][0] = s2; M[2][1] = s3; M[2][2] = s4;
            for (int c = 0; c < 3; ++c) {                 // elimination with partial pivoting
                int piv = c;
                for (int r = c + 1; r < 3; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
                if (piv != c) {
                    for (int k = 0; k < 3; ++k) { const double t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
                    const double t = rhs[c]; rhs[c] = rhs[piv]; rhs[piv] = t;
                }
                for (int r = c + 1; r < 3; ++r) {
                    const double f = M[r][c] / M[c][c];
                    for (int k = c; k < 3; ++k) M[r][k] = M[r][k] - f * M[c][k];
                    rhs[r] = rhs[r] - f * rhs[c];
                }
            }
            for (int r = 2; r >= 0; --r) {
                double acc = rhs[r];
                for (int k = r + 1; k < 3; ++k) acc -= M[r][k] * p[k];
                p[r] = acc / M[r][r];
            }
        }
        sh->dbc[0] = -p[1] / 2.0 / p[2];
        sh->dbc[1] = sqrt(fabs(-0.5 / p[2]));
    }
    __syncthreads();
    *b_out = sh->dbc[0]; *sn_out = sh->dbc[1];
    __syncthreads();
}

// thresholded_oasisAR1 (thresholded_oasisAR1.m:104-184,186-213), both branches: optimize_b fits y - b with b from
// estimate_baseline_noise (:142) and b = mean(y - solution) after every update_smin (:180)
__device__ void block_thresholded_ar1(const double* __restrict__ yraw, int T, double g, double sn, bool optimize_b, bool optimize_g,
                                      int maxIter, double thresh_factor, double p_noise, double g_lo, double g_hi,
                                      bool has_tau_range, TraceWS& ws, BlockShared* sh, DeconvOut* out) {
    double smin = choose_smin_ar1(g, sn, p_noise);
    const double thresh = thresh_factor * sn * sn * (double)T, tol = 1e-4;
    if (has_tau_range) g = fmin(fmax(g, g_lo), g_hi);
    bool g_conv = false;
    double b = 0.0;
    const double* y = yraw;                 // the trace the pools are fitted to: y - b (ws.yb) when the baseline is optimised
    if (optimize_b) {
        double sn_unused;
        block_estimate_baseline_noise(yraw, T, ws, sh, &b, &sn_unused);
        for (int i = threadIdx.x; i < T; i += blockDim.x) ws.yb[i] = yraw[i] - b;
        __syncthreads();
        y = ws.yb;
    }
    int n = block_oasis_ar1(y, T, g, 0.0, smin, ws, sh);
    double RSS0 = block_rss(y, ws.c, 0.0, T, sh);
    for (int it = 0; it < maxIter; ++it) {
        if (n == 0) break;
        if (optimize_g && !g_conv) {
            double g0 = g;
            g = block_update_g(y, T, &n, 0.0, smin, g_lo, g_hi, ws, sh);
            if (fabs(g - g0) / g0 < 1e-4) {
                g_conv = true;
                if (optimize_b) n = block_oasis_ar1(y, T, g, 0.0, smin, ws, sh);     // :156: cold re-run, optimize_b branch only
            }
        }
        double RSS = block_rss(y, ws.c, 0.0, T, sh);
        if (fabs(RSS - RSS0) < tol) break;
        double sc = 0.0;
        for (int i = threadIdx.x; i < T; i += blockDim.x) sc += ws.c[i];
        sc = block_sum(sc, sh->red);
        if (fabs(RSS - thresh) < tol || sc < 1e-9) break;
        RSS0 = RSS;
        // update_smin: bisection over <=9 candidates, warm PAV from the current pools
        double mx = -INFINITY;
        for (int p = threadIdx.x; p < n; p += blockDim.x) mx = fmax(mx, ws.pv[p] / ws.pw[p]);
        const double s_max = block_max(mx, sh->red);
        const int nsv = n < 9 ? n : 9;
        int ind_start = 1, ind_end = nsv;
        const double thr = sqrt(thresh);
        block_pow_table(g, T, ws.gp, sh);
        while (ind_end - ind_start > 1) {
            int ind = (ind_start + ind_end) / 2;
            double tmp_smin = smin + (double)(ind - 1) * ((s_max - smin) / (double)(nsv - 1));
            if (ind == nsv) tmp_smin = s_max;
            for (int p = threadIdx.x; p < n; p += blockDim.x) {
                ws.sv[p] = ws.pv[p]; ws.sw[p] = ws.pw[p]; ws.st[p] = ws.pt[p]; ws.sl[p] = ws.pl[p];
            }
            __syncthreads();
            int n2 = block_oasis_ar1_run(ws, n, tmp_smin, sh);
            block_oasis_ar1_solution(ws, n2, g, T, ws.c, ws.s);
            double sq = sqrt(block_rss(y, ws.c, 0.0, T, sh));
            if (sq < thr) {
                n = n2; smin = tmp_smin; ind_start = ind;
                // NB: the candidate grid `sv` is NOT recomputed (it was built from the smin at entry)
            } else {
                for (int p = threadIdx.x; p < n; p += blockDim.x) {
                    ws.pv[p] = ws.sv[p]; ws.pw[p] = ws.sw[p]; ws.pt[p] = ws.st[p]; ws.pl[p] = ws.sl[p];
                }
                __syncthreads();
                block_oasis_ar1_solution(ws, n, g, T, ws.c, ws.s);
                if (sq > thr) ind_end = ind; else break;
            }
        }
        if (optimize_b) {       // b = mean(y - solution) (:180); the next iteration fits y - b
            double a = 0.0;
            for (int i = threadIdx.x; i < T; i += blockDim.x) a += yraw[i] - ws.c[i];
            b = block_sum(a, sh->red) / (double)T;
            for (int i = threadIdx.x; i < T; i += blockDim.x) ws.yb[i] = yraw[i] - b;
            __syncthreads();
        }
    }
    out->b = b; out->g1 = g; out->g2 = 0.0; out->smin = smin;
}

// deconvolveCa (deconvolveCa.m:60-206).  y: T samples (read-only).  Results: ws.c, ws.s, *out.
// sn_in NaN => GetSn(y); pars_in both 0 => estimate_time_constant.  maxIter_override > 0 replaces opts.maxIter.
__device__ void block_deconvolveCa(const double* __restrict__ y, int T, const cnmfe_deconv_opts& o, double sn_in,
                                   double p1_in, double p2_in, int maxIter_override, double* ybuf, TraceWS& ws,
                                   BlockShared* sh, DeconvOut* out) {
    const int maxIter = maxIter_override > 0 ? maxIter_override : o.maxIter;
    double sn = isnan(sn_in) ? block_getsn(y, T, ws.scr, sh) : sn_in;
    out->sn = sn; out->lam = o.lambda; out->smin = o.smin; out->b = o.b; out->npars = o.type;
    double g1 = p1_in, g2 = p2_in;
    const bool nopars = (o.type == 1) ? (g1 == 0.0) : (g1 == 0.0 && g2 == 0.0);
    if (nopars) {
        double gg[2] = {0.0, 0.0};
        int np = block_time_constant(y, T, o.type, sn, gg, sh);
        if (np != o.type) {   // deconvolveCa.m:84-101: c = s = 0, pars = 0
            for (int i = threadIdx.x; i < T; i += blockDim.x) { ws.c[i] = 0.0; ws.s[i] = 0.0; }
            __syncthreads();
            out->g1 = 0.0; out->g2 = 0.0;
            return;
        }
        g1 = gg[0]; g2 = gg[1];
    }
    const double b0 = o.b;
    const double* yin = y;
    if (b0 != 0.0 && o.method == 0) {
        for (int i = threadIdx.x; i < T; i += blockDim.x) ybuf[i] = y[i] - b0;
        __syncthreads();
        yin = ybuf;
    }
    double g_lo = 0.0, g_hi = 1.0;
    if (o.has_tau_range) { g_lo = exp(-1.0 / o.tau_range[0]); g_hi = exp(-1.0 / o.tau_range[1]); }
    out->g1 = g1; out->g2 = g2;
    if (o.method == 0) {
        if (o.type == 1) {
            double smin = o.smin;
            if (smin < 0.0) smin = fabs(smin) * sn;
            out->smin = smin;
            double gmax = exp(-1.0 / o.max_tau);
            block_foopsi_ar1(yin, T, g1, o.lambda, smin, o.optimize_b != 0, o.optimize_pars != 0, maxIter, g_lo,
                             g_hi, o.has_tau_range != 0, gmax, ws, sh, out);
            out->b = out->b + b0;
        } else {
            double smin = o.smin;
            if (smin < 0.0) smin = fabs(smin) * sn / ar2_max_ht(g1, g2);
            out->smin = smin;
            block_oasis_ar2(yin, T, g1, g2, o.lambda, smin, ws, sh);
            out->b = 0.0 + b0;
        }
    } else if (o.method == 1) {
        block_constrained_ar1(y, T, g1, sn, o.optimize_b != 0, o.optimize_pars != 0, maxIter, g_lo, g_hi,
                              o.has_tau_range != 0, ws, sh, out);
    } else if (o.type == 1) {
        block_thresholded_ar1(y, T, g1, sn, o.optimize_b != 0, o.optimize_pars != 0, maxIter, o.thresh_factor, o.p_noise, g_lo, g_hi,
                              o.has_tau_range != 0, ws, sh, out);
    } else {
        // thresholded_oasisAR2 with optimize_g = false: both loops (:96-126, :133-163) exit at their first
        // abs(RSS-RSS0)<tol test (RSS is recomputed from the unchanged solution), so the result is one oasisAR2 pass with
        // smin = choose_smin(g, sn, 0.99999999) (:72) -- on y - b with b = estimate_baseline_noise(y) when optimize_b (:129-130)
        const double smin = choose_smin_ar2(g1, g2, sn, 0.99999999);
        double b = 0.0;
        const double* yfit = y;
        if (o.optimize_b) {
            double sn_unused;
            block_estimate_baseline_noise(y, T, ws, sh, &b, &sn_unused);
            for (int i = threadIdx.x; i < T; i += blockDim.x) ws.yb[i] = y[i] - b;
            __syncthreads();
            yfit = ws.yb;
        }
        block_oasis_ar2(yfit, T, g1, g2, 0.0, smin, ws, sh);
        out->b = b; out->smin = smin;
    }
    // avoid nan output (deconvolveCa.m:206)
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double v = ws.c[i];
        if (isnan(v) || isinf(v)) ws.c[i] = 0.0;
    }
    __syncthreads();
}

}  // namespace cnmfe
