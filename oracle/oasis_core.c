/*
 * oracle/oasis_core.c -- TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
 *
 * Plain-C float64 restatement of the pool-adjacent-violators loops of the reference's OASIS
 * deconvolution, so that the NumPy oracle (oracle/oasis.py) runs in seconds instead of hours.
 *
 *   oasis_ar1_pools   <- OASIS_matlab/packages/oasis/oasisAR1.m:57-109   (run + construct solution)
 *   oasis_ar2         <- OASIS_matlab/packages/oasis/oasisAR2.m:49-156
 *   rss_g_ar1         <- OASIS_matlab/packages/oasis/foopsi_oasisAR1.m:165-178 (nested rss_g)
 *   rebuild_pools_ar1 <- OASIS_matlab/packages/oasis/foopsi_oasisAR1.m:153-162
 *
 * PARITY UNPINNED: the reference holds no golden vectors for these routines and
 * MATLAB is not available here; the restatement is pinned only by solver-independent known-answer
 * tests (tests/test_oracle_*.py).
 *
 * Indices are 0-based here (the reference is 1-based); "nil" replaces NaN links.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NIL (-1)

/* Runs the AR(1) PAV on n pools given as arrays (v, w, t, l); t is 0-based start, l length.
 * Returns the number of surviving pools (compacted in place). oasisAR1.m:57-98 */
int oasis_ar1_pools(double *v, double *w, int *t, int *l, int n, double g, double smin)
{
    if (n <= 0) return 0;
    int *prev = (int *)malloc(sizeof(int) * (size_t)n);
    int *next = (int *)malloc(sizeof(int) * (size_t)n);
    char *alive = (char *)malloc((size_t)n);
    for (int i = 0; i < n; ++i) { prev[i] = i - 1; next[i] = i + 1; alive[i] = 1; }
    next[n - 1] = NIL; prev[0] = NIL;

    int ii = 0, ii_next = next[0], ii_prev;
    while (ii_next != NIL) {
        /* find the active set (oasisAR1.m:63-68) */
        while (ii_next != NIL &&
               (v[ii_next] / w[ii_next] >= v[ii] / w[ii] * pow(g, (double)l[ii]) + smin)) {
            prev[ii_next] = ii;
            ii = ii_next;
            ii_next = next[ii];
        }
        if (ii_next == NIL) break;
        /* merge pools (oasisAR1.m:73-79) */
        v[ii] = v[ii] + v[ii_next] * pow(g, (double)l[ii]);
        w[ii] = w[ii] + w[ii_next] * pow(g, 2.0 * (double)l[ii]);
        l[ii] = l[ii] + l[ii_next];
        next[ii] = next[ii_next];
        alive[ii_next] = 0;
        ii_next = next[ii];
        ii_prev = prev[ii];
        /* backtrack until violations fixed (oasisAR1.m:82-95) */
        while (ii_prev != NIL &&
               (v[ii] / w[ii] <
                fmax(0.0, v[ii_prev] / w[ii_prev] * pow(g, (double)l[ii_prev])) + smin)) {
            ii_next = ii;
            ii = ii_prev;
            v[ii] = v[ii] + v[ii_next] * pow(g, (double)l[ii]);
            w[ii] = w[ii] + w[ii_next] * pow(g, 2.0 * (double)l[ii]);
            l[ii] = l[ii] + l[ii_next];
            next[ii] = next[ii_next];
            alive[ii_next] = 0;
            ii_prev = prev[ii];
            ii_next = next[ii];
        }
    }
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (alive[i]) { v[m] = v[i]; w[m] = w[i]; t[m] = t[i]; l[m] = l[i]; ++m; }
    free(prev); free(next); free(alive);
    return m;
}

/* c, s from pools. oasisAR1.m:101-109 */
void oasis_ar1_solution(const double *v, const double *w, const int *t, const int *l, int n,
                        double g, int T, double *c, double *s)
{
    memset(c, 0, sizeof(double) * (size_t)T);
    memset(s, 0, sizeof(double) * (size_t)T);
    for (int i = 0; i < n; ++i) {
        double a = fmax(0.0, v[i] / w[i]);
        for (int j = 0; j < l[i]; ++j) c[t[i] + j] = a * pow(g, (double)j);
    }
    for (int i = 1; i < n; ++i) s[t[i]] = c[t[i]] - g * c[t[i] - 1];
}

/* rss_g nested function of update_g: foopsi_oasisAR1.m:165-178.
 * h = exp(log(g)*(0:maxl)); hh = cumsum(h.*h); yp = y - lam*(1-g);
 * per pool: tmp_v = max(yp(idx)'*h(1:li)/hh(li),0); c(idx)=tmp_v*h(1:li); rss = |y-c|^2.
 * h, hh are caller-provided scratch of length maxl+1 and are left holding the LAST evaluated kernel
 * (the reference shares `h` with the enclosing scope, foopsi_oasisAR1.m:151-157). */
double rss_g_ar1(const double *y, int T, const int *t, const int *l, int n, double g, double lam,
                 int maxl, double *h, double *hh, double *c)
{
    double lg = log(g), acc = 0.0;
    for (int j = 0; j <= maxl; ++j) { h[j] = exp(lg * (double)j); acc += h[j] * h[j]; hh[j] = acc; }
    double pen = lam * (1.0 - g);
    for (int i = 0; i < n; ++i) {
        double dot = 0.0;
        for (int j = 0; j < l[i]; ++j) dot += (y[t[i] + j] - pen) * h[j];
        double tv = fmax(dot / hh[l[i] - 1], 0.0);
        for (int j = 0; j < l[i]; ++j) c[t[i] + j] = tv * h[j];
    }
    double rss = 0.0;
    for (int k = 0; k < T; ++k) { double r = y[k] - c[k]; rss += r * r; }
    return rss;
}

/* foopsi_oasisAR1.m:153-162: v = yp(idx)'*tmp_h(1:li) with tmp_h from the RETURNED g,
 * w = tmp_hh(li) with tmp_hh = cumsum(h.*h) from the LAST-EVALUATED h (hh_last). */
void rebuild_pools_ar1(const double *y, const int *t, const int *l, int n, double g, double lam,
                       int maxl, const double *hh_last, double *v, double *w)
{
    double lg = log(g), pen = lam * (1.0 - g);
    double *th = (double *)malloc(sizeof(double) * (size_t)(maxl + 1));
    for (int j = 0; j <= maxl; ++j) th[j] = exp(lg * (double)j);
    for (int i = 0; i < n; ++i) {
        double dot = 0.0;
        for (int j = 0; j < l[i]; ++j) dot += (y[t[i] + j] - pen) * th[j];
        v[i] = dot;
        w[i] = hh_last[l[i] - 1];
    }
    free(th);
}

/* AR(2) PAV, cold start. oasisAR2.m:49-156. d >= r are the real roots of z^2 - g1 z - g2.
 * Returns number of pools; pool arrays (v = first value, w = last value, t, l) are outputs of size T. */
int oasis_ar2(const double *y, int T, double g1, double g2, double d, double r, double lam,
              double smin, double *c, double *s, double *pv, double *pw, int *pt, int *pl)
{
    double *yp = (double *)malloc(sizeof(double) * (size_t)T);
    for (int k = 0; k < T; ++k) yp[k] = y[k] - lam * (1.0 - g1 - g2);
    if (T >= 2) yp[T - 2] = y[T - 2] - lam * (1.0 - g1);
    if (T >= 1) yp[T - 1] = y[T - 1] - lam;

    int n = T;
    int *prev = (int *)malloc(sizeof(int) * (size_t)n);
    int *next = (int *)malloc(sizeof(int) * (size_t)n);
    char *alive = (char *)malloc((size_t)n);
    for (int i = 0; i < n; ++i) {
        pv[i] = yp[i]; pw[i] = yp[i]; pt[i] = i; pl[i] = 1;
        prev[i] = i - 1; next[i] = i + 1; alive[i] = 1;
    }
    prev[0] = NIL; next[n - 1] = NIL;

    /* precompute (oasisAR2.m:70-76); index k here = MATLAB k+1 */
    double *g11 = (double *)malloc(sizeof(double) * (size_t)T);
    double *g12 = (double *)malloc(sizeof(double) * (size_t)T);
    double *g11g11 = (double *)malloc(sizeof(double) * (size_t)T);
    double *g11g12 = (double *)malloc(sizeof(double) * (size_t)T);
    double ld = log(d), lr = log(r), a1 = 0.0, a2 = 0.0;
    for (int k = 0; k < T; ++k) {
        g11[k] = (exp(ld * (double)(k + 1)) - exp(lr * (double)(k + 1))) / (d - r);
        g12[k] = (k == 0) ? 0.0 : g2 * g11[k - 1];
        a1 += g11[k] * g11[k]; g11g11[k] = a1;
        a2 += g11[k] * g12[k]; g11g12[k] = a2;
    }

    if (T >= 3) {
        int ii = 1, ii_next = next[ii], ii_prev = prev[ii];
        while (ii_next != NIL) {
            /* find the active set (oasisAR2.m:83-89); g11(l+1) -> g11[l] */
            while (ii_next != NIL &&
                   (g11[pl[ii]] * pv[ii] + g12[pl[ii]] * pw[ii_prev] + smin <= pv[ii_next])) {
                prev[ii_next] = ii;
                ii = ii_next;
                ii_next = next[ii];
                ii_prev = prev[ii];
            }
            if (ii_next == NIL) break;
            /* merge pools (oasisAR2.m:93-104) */
            pl[ii] += pl[ii_next];
            {
                int ti = pt[ii], li = pl[ii];
                double dot = 0.0;
                for (int j = 0; j < li; ++j) dot += g11[j] * yp[ti + j];
                pv[ii] = (dot - g11g12[li - 1] * pw[ii_prev]) / g11g11[li - 1];
                pw[ii] = g11[li - 1] * pv[ii] + g12[li - 1] * pw[ii_prev];
            }
            alive[ii_next] = 0;
            next[ii] = next[ii_next];
            ii_next = next[ii];
            if (ii_next != NIL) prev[ii_next] = ii;

            int p1 = prev[ii];
            int p2 = (p1 != NIL) ? prev[p1] : NIL;
            /* backtrack (oasisAR2.m:109-128) */
            while (p2 != NIL &&
                   (g11[pl[p1]] * pv[p1] + g12[pl[p1]] * pw[p2] + smin > pv[ii])) {
                ii_next = ii;
                ii = p1;
                ii_prev = p2;
                pl[ii] += pl[ii_next];
                {
                    int ti = pt[ii], li = pl[ii];
                    double dot = 0.0;
                    for (int j = 0; j < li; ++j) dot += g11[j] * yp[ti + j];
                    pv[ii] = (dot - g11g12[li - 1] * pw[ii_prev]) / g11g11[li - 1];
                    pw[ii] = g11[li - 1] * pv[ii] + g12[li - 1] * pw[ii_prev];
                }
                alive[ii_next] = 0;
                next[ii] = next[ii_next];
                ii_next = next[ii];
                if (ii_next != NIL) prev[ii_next] = ii;
                p1 = prev[ii];
                p2 = (p1 != NIL) ? prev[p1] : NIL;
            }
        }
    }
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (alive[i]) { pv[m] = pv[i]; pw[m] = pw[i]; pt[m] = pt[i]; pl[m] = pl[i]; ++m; }

    /* construct solution (oasisAR2.m:140-156) */
    memset(c, 0, sizeof(double) * (size_t)T);
    for (int i = 0; i < m; ++i) {
        int ti = pt[i], li = pl[i];
        c[ti] = pv[i];
        for (int j = 1; j < li; ++j)
            c[ti + j] = g1 * c[ti + j - 1] + g2 * ((ti + j - 2 >= 0) ? c[ti + j - 2] : 0.0);
    }
    for (int k = 0; k < T; ++k) if (c[k] < 0) c[k] = 0;
    for (int k = 0; k < T; ++k) s[k] = 0.0;
    for (int k = 3; k < T; ++k) {
        double sv = c[k] - g1 * c[k - 1] - g2 * c[k - 2];
        s[k] = (sv < smin) ? 0.0 : sv;
    }
    free(yp); free(prev); free(next); free(alive);
    free(g11); free(g12); free(g11g11); free(g11g12);
    return m;
}
