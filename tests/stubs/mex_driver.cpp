// Drives matlab/cnmfe_b200_mex.cpp (built against the functional mex stub) through the command sequence the drop-in methods
// matlab/@Sources2D/update_*_parallel.m and the helpers matlab/cnmfe_b200_*.m issue.  Linked either with the recording fake
// library (argument marshalling) or with the real libcnmfe_b200.so (error path without a GPU: argument "real").
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "mex_test_api.h"

static std::vector<mxArray*> call(int nlhs, std::vector<mxArray*> in) {
    std::vector<mxArray*> out(8, nullptr);
    mexFunction(nlhs, out.data(), (int)in.size(), const_cast<const mxArray**>(in.data()));
    return out;
}

int main(int argc, char** argv) {
    const bool real = argc > 1 && !strcmp(argv[1], "real");
    const int d1 = 6, d2 = 4, T = 40, K = 2;
    int32_t pp[8] = {1, 3, 1, 4, 4, 6, 1, 4}, bp[8] = {1, 5, 1, 4, 2, 6, 1, 4};
    try {
        auto h = call(1, {t_string("create"), t_scalar(d1), t_scalar(d2), t_scalar(T), t_int32(4, 2, pp), t_int32(4, 2, bp), t_scalar(2), t_scalar(0), t_scalar(0)})[0];
        if (real) { printf("created on a real device\n"); call(0, {t_string("destroy"), h}); return 0; }
        std::vector<uint16_t> Y(5 * 4 * T, 0);
        Y[0] = 1234;
        call(0, {t_string("upload_block"), h, t_scalar(0), t_uint16_3d(5, 4, T, Y.data())});
        // cnmfe_b200_push: options struct
        mxArray* dopt = t_struct();
        t_setfield(dopt, "type", t_string("ar1")); t_setfield(dopt, "method", t_string("foopsi")); t_setfield(dopt, "smin", t_scalar(-5));
        t_setfield(dopt, "optimize_pars", t_logical(true)); t_setfield(dopt, "optimize_b", t_logical(true)); t_setfield(dopt, "max_tau", t_scalar(100));
        mxArray* o = t_struct();
        t_setfield(o, "spatial_algorithm", t_scalar(2)); t_setfield(o, "maxIter", t_scalar(5)); t_setfield(o, "deconv_flag", t_logical(true));
        t_setfield(o, "bg_acceleration", t_logical(true)); t_setfield(o, "background_model", t_string("ring")); t_setfield(o, "nb", t_scalar(1));
        t_setfield(o, "bg_ssub", t_scalar(2)); t_setfield(o, "deconv_options", dopt); t_setfield(o, "thresh_outlier", t_scalar(2.5));
        call(0, {t_string("set_options"), h, o});
        auto sd = call(5, {t_string("ssub_dims"), h, t_scalar(1)});
        printf("ssub_dims -> d1s=%g d2s=%g nnb=%g r_shift[3]=%d c_shift[3]=%d\n", mxGetScalar(sd[0]), mxGetScalar(sd[1]), mxGetScalar(sd[2]),
               ((int32_t*)mxGetData(sd[3]))[3], ((int32_t*)mxGetData(sd[4]))[3]);
        auto ro = call(2, {t_string("ring_offsets"), h});
        printf("ring_offsets -> n=%zu r_shift[0]=%d c_shift[0]=%d\n", mxGetN(ro[0]), ((int32_t*)mxGetData(ro[0]))[0], ((int32_t*)mxGetData(ro[1]))[0]);
        std::vector<double> Ws(4 * 6); for (size_t i = 0; i < Ws.size(); ++i) Ws[i] = 0.5 + i;
        std::vector<double> b0(12, 2.5);
        call(0, {t_string("set_ring"), h, t_scalar(1), t_double(4, 6, Ws.data()), t_double(12, 1, b0.data())});
        call(0, {t_string("set_ring"), h, t_scalar(0), t_double(0, 0, nullptr), t_double(12, 1, b0.data())});
        // A: d x K sparse with 3 entries, C: K x T
        std::vector<mwIndex> jc = {0, 2, 3}, ir = {3, 4, 10};
        std::vector<double> pr = {1.5, 2.5, 3.5}, C(K * T, 0.0);
        C[0] = 42;
        call(0, {t_string("set_neurons"), h, t_sparse(d1 * d2, K, jc, ir, pr, false), t_double(K, T, C.data())});
        call(0, {t_string("update_background"), h});
        auto gr = call(2, {t_string("get_ring"), h, t_scalar(1), t_scalar(4), t_scalar(6), t_scalar(12)});
        printf("get_ring -> W %zux%zu W[5]=%g b0 %zux%zu b0[0]=%g\n", mxGetM(gr[0]), mxGetN(gr[0]), mxGetPr(gr[0])[5], mxGetM(gr[1]), mxGetN(gr[1]), mxGetPr(gr[1])[0]);
        // spatial
        C[0] = 43;
        call(0, {t_string("set_prev"), h, t_sparse(d1 * d2, K, jc, ir, pr, false), t_double(K, T, C.data())});
        std::vector<double> sn(d1 * d2, 10.0);
        call(0, {t_string("set_sn"), h, t_double(d1, d2, sn.data())});
        auto sl = call(2, {t_string("search_location"), t_sparse(d1 * d2, K, jc, ir, pr, false), t_scalar(d1), t_scalar(d2), t_scalar(3), t_scalar(8), t_scalar(3)});
        printf("search_location -> jc[K]=%g ir[1]=%g\n", mxGetPr(sl[0])[K], mxGetPr(sl[1])[1]);
        call(0, {t_string("set_search"), h, t_sparse(d1 * d2, K, jc, ir, {}, true)});
        auto us = call(2, {t_string("update_spatial"), h, t_scalar(3), t_logical(true), t_scalar(d1), t_scalar(d2)});
        printf("update_spatial -> n=%zu v1=%g sn %zux%zu sn0=%g\n", mxGetM(us[0]), mxGetPr(us[0])[1], mxGetM(us[1]), mxGetN(us[1]), mxGetPr(us[1])[0]);
        auto pps = call(1, {t_string("post_process_spatial"), t_sparse(d1 * d2, K, jc, ir, pr, false), t_scalar(d1), t_scalar(d2)});
        printf("post_process_spatial -> sparse=%d pr0=%g pr1=%g\n", (int)mxIsSparse(pps[0]), mxGetPr(pps[0])[0], mxGetPr(pps[0])[1]);
        // temporal
        call(0, {t_string("set_use_c_hat"), h, t_logical(false)});
        auto ut = call(5, {t_string("update_temporal"), h, t_scalar(K), t_scalar(T)});
        printf("update_temporal -> C %zux%zu C0=%g Craw0=%g S0=%g kp %zux%zu kp0=%g nsn0=%g\n", mxGetM(ut[0]), mxGetN(ut[0]), mxGetPr(ut[0])[0], mxGetPr(ut[1])[0],
               mxGetPr(ut[2])[0], mxGetM(ut[3]), mxGetN(ut[3]), mxGetPr(ut[3])[0], mxGetPr(ut[4])[0]);
        // svd / nmf state, noise, stand-alone deconvolution
        std::vector<double> b(12, 0.5), f(T, 0.25);
        call(0, {t_string("set_bf"), h, t_scalar(0), t_double(12, 1, b.data()), t_double(1, T, f.data()), t_double(0, 0, nullptr)});
        auto gb = call(3, {t_string("get_bf"), h, t_scalar(0), t_scalar(12), t_scalar(1), t_scalar(T)});
        printf("get_bf -> b %zux%zu f %zux%zu b0 %zux%zu\n", mxGetM(gb[0]), mxGetN(gb[0]), mxGetM(gb[1]), mxGetN(gb[1]), mxGetM(gb[2]), mxGetN(gb[2]));
        auto en = call(1, {t_string("estimate_noise"), h, t_scalar(1), t_scalar(T), t_scalar(d1), t_scalar(d2)});
        printf("estimate_noise -> %zux%zu sn0=%g\n", mxGetM(en[0]), mxGetN(en[0]), mxGetPr(en[0])[0]);
        std::vector<double> bmap(d1 * d2, 3.25), bnew(d1 * d2, 4.5);
        auto rs = call(1, {t_string("compute_rss"), h, t_scalar(1), t_scalar(T), t_double(d1, d2, bmap.data()), t_double(d1, d2, bnew.data()), t_scalar(2)});
        printf("compute_rss -> %zux%zu rss1=%g\n", mxGetM(rs[0]), mxGetN(rs[0]), mxGetPr(rs[0])[1]);
        auto yb = call(1, {t_string("reconstruct_background"), h, t_scalar(1), t_scalar(3), t_scalar(12), t_double(d1, d2, bmap.data()), t_double(d1, d2, bnew.data()), t_scalar(12)});
        printf("reconstruct_background -> %zux%zu y0=%g\n", mxGetM(yb[0]), mxGetN(yb[0]), mxGetPr(yb[0])[0]);
        std::vector<double> y(T * 3, 0.125);
        auto dc = call(7, {t_string("deconvolve"), t_double(T, 3, y.data()), dopt, t_double(0, 0, nullptr), t_double(0, 0, nullptr)});
        printf("deconvolve -> c %zux%zu pars %zux%zu lam0=%g\n", mxGetM(dc[0]), mxGetN(dc[0]), mxGetM(dc[3]), mxGetN(dc[3]), mxGetPr(dc[6])[0]);
        // unknown model string must be rejected by the gateway itself
        mxArray* bad = t_struct();
        t_setfield(bad, "background_model", t_string("pca"));
        try { call(0, {t_string("set_options"), h, bad}); printf("ERROR: bad model accepted\n"); }
        catch (const std::runtime_error& e) { printf("rejected: %s\n", e.what()); }
        // lifetime: explicit destroy of one handle, the other released by the at-exit hook
        auto h2 = call(1, {t_string("create"), t_scalar(d1), t_scalar(d2), t_scalar(T), t_int32(4, 2, pp), t_int32(4, 2, bp), t_scalar(2), t_scalar(0), t_scalar(0)})[0];
        printf("locks=%d\n", g_mex_locks);
        call(0, {t_string("destroy"), h});
        printf("locks=%d\n", g_mex_locks);
        (void)h2;
        t_run_atexit();
        try { call(1, {t_string("create"), t_scalar(d1), t_scalar(d2), t_scalar(T), t_int32(4, 2, pp), t_int32(4, 2, bp), t_scalar(2), t_scalar(0), t_scalar(99)}); }
        catch (const std::runtime_error& e) { printf("error path: %s\n", e.what()); }
    } catch (const std::runtime_error& e) {
        printf("mex error: %s\n", e.what());
        return real ? 0 : 1;
    }
    return 0;
}
