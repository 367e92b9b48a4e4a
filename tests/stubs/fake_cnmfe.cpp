// Recording stand-in for libcnmfe_b200.so (test only): every entry point the MEX gateway binds logs its arguments to stdout and
// returns recognisable values, so tests/test_mex_gateway.py can check the gateway's marshalling without a GPU.
#include <cstdio>
#include <cstring>
#include "cnmfe_b200.h"

struct cnmfe_ctx { int d1, d2, T, np, nnb; };
static char g_err[256] = "";
static int g_created = 0, g_destroyed = 0;

extern "C" {
const char* cnmfe_last_error(void) { return g_err; }
void cnmfe_deconv_defaults(cnmfe_deconv_opts* o) {
    memset(o, 0, sizeof *o);
    o->type = 1; o->method = 1; o->maxIter = 10; o->max_tau = 100.0; o->thresh_factor = 1.0; o->p_noise = 0.9999;
}
void cnmfe_options_defaults(cnmfe_options* o) {
    memset(o, 0, sizeof *o);
    o->maxIter_temporal = 5; o->deconv_flag = 1; o->bg_acceleration = 1; o->replicate_spatial_aprev_quirk = 1; o->use_tensor_gram = 1;
    o->nb = 1; o->bg_ssub = 1; o->thresh_outlier = 0.0 / 0.0;
    cnmfe_deconv_defaults(&o->deconv);
}
int cnmfe_create(cnmfe_ctx** ctx, int d1, int d2, int T, int npatch, const int32_t* pp, const int32_t* bp, const uint8_t* owned,
                 int ring_radius, int num_neighbors, int device) {
    printf("create d1=%d d2=%d T=%d np=%d pp0=[%d %d %d %d] bp1=[%d %d %d %d] owned=%s rr=%d nn=%d dev=%d\n", d1, d2, T, npatch, pp[0], pp[1], pp[2],
           pp[3], bp[4], bp[5], bp[6], bp[7], owned ? "set" : "null", ring_radius, num_neighbors, device);
    if (device == 99) { snprintf(g_err, sizeof g_err, "cnmfe_create: device 99 out of range"); return -1; }
    *ctx = new cnmfe_ctx{d1, d2, T, npatch, 4};
    ++g_created;
    return 0;
}
void cnmfe_destroy(cnmfe_ctx* c) { ++g_destroyed; printf("destroy (%d of %d)\n", g_destroyed, g_created); delete c; }
int cnmfe_upload_block(cnmfe_ctx*, int ip, const void* Y, int dtype) {
    printf("upload_block ip=%d dtype=%d first=%u\n", ip, dtype, (unsigned)((const uint16_t*)Y)[0]);
    return 0;
}
int cnmfe_set_options(cnmfe_ctx*, const cnmfe_options* o) {
    printf("set_options alg=%d maxIter=%d deconv_flag=%d accel=%d quirk=%d tensor=%d model=%d nb=%d ssub=%d outlier=%g | type=%d method=%d optb=%d optp=%d maxIter=%d "
           "smin=%g lambda=%g max_tau=%g tau_range=%d\n", o->spatial_algorithm, o->maxIter_temporal, o->deconv_flag, o->bg_acceleration,
           o->replicate_spatial_aprev_quirk, o->use_tensor_gram, o->background_model, o->nb, o->bg_ssub, o->thresh_outlier, o->deconv.type, o->deconv.method,
           o->deconv.optimize_b, o->deconv.optimize_pars, o->deconv.maxIter, o->deconv.smin, o->deconv.lambda, o->deconv.max_tau, o->deconv.has_tau_range);
    return 0;
}
static int log_csc(const char* what, int K, const int64_t* jc, const int64_t* ir, const double* pr, const double* C) {
    printf("%s K=%d nnz=%lld ir0=%lld pr0=%g C0=%g\n", what, K, (long long)jc[K], (long long)ir[0], pr ? pr[0] : -1.0, C ? C[0] : -1.0);
    return 0;
}
int cnmfe_set_neurons(cnmfe_ctx*, int K, const int64_t* jc, const int64_t* ir, const double* pr, const double* C) { return log_csc("set_neurons", K, jc, ir, pr, C); }
int cnmfe_set_prev(cnmfe_ctx*, int K, const int64_t* jc, const int64_t* ir, const double* pr, const double* C) { return log_csc("set_prev", K, jc, ir, pr, C); }
int cnmfe_set_search(cnmfe_ctx*, int K, const int64_t* jc, const int64_t* ir) { return log_csc("set_search", K, jc, ir, nullptr, nullptr); }
int cnmfe_set_sn(cnmfe_ctx*, const double* sn) { printf("set_sn sn0=%g\n", sn[0]); return 0; }
int cnmfe_set_use_c_hat(cnmfe_ctx*, int f) { printf("set_use_c_hat %d\n", f); return 0; }
int cnmfe_ring_offsets(cnmfe_ctx* c, int* nnb, int32_t* r, int32_t* cs) {
    *nnb = c->nnb;
    if (r) for (int i = 0; i < c->nnb; ++i) { r[i] = i - 2; cs[i] = 2 - i; }
    return 0;
}
int cnmfe_ssub_dims(cnmfe_ctx* c, int ip, int* d1s, int* d2s, int* nnb, int32_t* r, int32_t* cs) {
    *d1s = 3; *d2s = 2; *nnb = c->nnb;
    if (r) { printf("ssub_dims ip=%d\n", ip); for (int i = 0; i < c->nnb; ++i) { r[i] = i; cs[i] = -i; } }
    return 0;
}
int cnmfe_set_ring(cnmfe_ctx*, int ip, const double* W, const double* b0) {
    printf("set_ring ip=%d W=%s b0=%s W1=%g b00=%g\n", ip, W ? "set" : "null", b0 ? "set" : "null", W ? W[1] : -1.0, b0 ? b0[0] : -1.0);
    return 0;
}
int cnmfe_get_ring(cnmfe_ctx* c, int ip, double* W, double* b0) {
    printf("get_ring ip=%d\n", ip);
    if (W) for (int i = 0; i < c->nnb * 6; ++i) W[i] = 100 + i;
    if (b0) b0[0] = 7.5;
    return 0;
}
int cnmfe_set_bf(cnmfe_ctx*, int ip, const double* b, const double* f, const double* b0) {
    printf("set_bf ip=%d b0=%g f1=%g b0vec=%s\n", ip, b[0], f[1], b0 ? "set" : "null");
    return 0;
}
int cnmfe_get_bf(cnmfe_ctx*, int ip, double* b, double* f, double* b0) { printf("get_bf ip=%d\n", ip); b[0] = 1; f[0] = 2; b0[0] = 3; return 0; }
int cnmfe_estimate_noise(cnmfe_ctx*, int f0, int f1, double* sn) { printf("estimate_noise %d %d\n", f0, f1); sn[0] = 9; return 0; }
int cnmfe_compute_rss(cnmfe_ctx* c, int f0, int f1, const double* b0, const double* b0n, double* rss) {
    printf("compute_rss %d %d b0=%g b0new=%g\n", f0, f1, b0[0], b0n[0]);
    for (int i = 0; i < c->np; ++i) rss[i] = 10.0 + i;
    return 0;
}
int cnmfe_reconstruct_background(cnmfe_ctx*, int ip, int f0, int f1, const double* b0, const double* b0n, double* Y) {
    printf("reconstruct_background ip=%d %d %d b0=%g b0new=%g\n", ip, f0, f1, b0[0], b0n[0]);
    Y[0] = 5.5;
    return 0;
}
int cnmfe_update_background(cnmfe_ctx*) { printf("update_background\n"); return 0; }
int cnmfe_update_spatial_ex(cnmfe_ctx*, int usn) { printf("update_spatial update_sn=%d\n", usn); return 0; }
int cnmfe_get_spatial(cnmfe_ctx*, double* v) { v[0] = 0.25; v[1] = 0.5; return 0; }
int cnmfe_get_sn_map(cnmfe_ctx*, double* sn) { sn[0] = 11; return 0; }
int cnmfe_set_spatial(cnmfe_ctx*, const double* v) { printf("set_spatial v0=%g\n", v[0]); return 0; }
int cnmfe_update_temporal(cnmfe_ctx*) { printf("update_temporal\n"); return 0; }
int cnmfe_get_temporal(cnmfe_ctx*, double* C, double* Cr, double* S, double* kp, double* nsn) {
    C[0] = 1; Cr[0] = 2; S[0] = 3; kp[0] = 0.95; nsn[0] = 0.1;
    return 0;
}
int cnmfe_deconvolve(const double* Y, int T, int N, const cnmfe_deconv_opts* o, const double* sn, const double* pars, double* c, double* s, double* b,
                     double* po, double* sno, double* smin, double* lam, int dev) {
    printf("deconvolve T=%d N=%d type=%d method=%d sn=%s pars=%s y0=%g dev=%d\n", T, N, o->type, o->method, sn ? "set" : "null", pars ? "set" : "null", Y[0], dev);
    c[0] = 1; s[0] = 2; b[0] = 3; po[0] = 4; sno[0] = 5; smin[0] = 6; lam[0] = 7;
    return 0;
}
int cnmfe_get_sn(const double* Y, int T, int N, double* sn, int) { printf("get_sn T=%d N=%d y0=%g\n", T, N, Y[0]); sn[0] = 0.3; return 0; }
int cnmfe_connectivity_constraint(int d1, int d2, int K, const int64_t*, const int64_t*, double* pr, double thr, int sz) {
    printf("connectivity_constraint d1=%d d2=%d K=%d thr=%g sz=%d\n", d1, d2, K, thr, sz); pr[0] = 0; return 0;
}
int cnmfe_circular_constraints(int d1, int d2, int K, const int64_t*, const int64_t*, const double*, int64_t* ojc, int64_t* oir, double* opr, int64_t cap) {
    printf("circular_constraints d1=%d d2=%d K=%d cap=%lld\n", d1, d2, K, (long long)cap);
    for (int k = 0; k <= K; ++k) ojc[k] = k ? 1 : 0;
    oir[0] = 5; opr[0] = 0.75;
    return 0;
}
int cnmfe_search_location_dilate(int d1, int d2, int K, const int64_t*, const int64_t*, const double*, double nrgthr, int nb, int bSiz, int64_t* ojc,
                                 int64_t* oir, int64_t cap) {
    printf("search_location_dilate d1=%d d2=%d K=%d nrgthr=%g nb=%d bSiz=%d cap=%lld\n", d1, d2, K, nrgthr, nb, bSiz, (long long)cap);
    for (int k = 0; k <= K; ++k) ojc[k] = k ? 1 : 0;
    oir[0] = 4;
    return 0;
}
int cnmfe_search_location_ellipse(int d1, int d2, int K, const int64_t*, const int64_t*, const double*, double mn, double mx, double dist, int64_t* ojc,
                                  int64_t* oir, int64_t cap) {
    printf("search_location d1=%d d2=%d K=%d min=%g max=%g dist=%g cap=%lld\n", d1, d2, K, mn, mx, dist, (long long)cap);
    for (int k = 0; k <= K; ++k) ojc[k] = k ? 2 : 0;
    oir[0] = 1; oir[1] = 2;
    return 0;
}
}
