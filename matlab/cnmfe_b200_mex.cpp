// cnmfe_b200_mex.cpp -- MEX gateway binding libcnmfe_b200.so (include/cnmfe_b200.h) into MATLAB.
//
// Build on a machine with MATLAB + the built library (cannot be compiled in the development image: no mex.h):
//   mex -O -largeArrayDims cnmfe_b200_mex.cpp -I../include -L../cnmf_e_b200 -lcnmfe_b200
// Conventions follow the only native file of the reference, utilities/graph_conn_comp_mex.cpp: outputs allocated with
// mxCreate* and owned by MATLAB (:58-65), inputs const (:38-56), errors through mexErrMsgIdAndTxt (:44-53).
//
// Usage:  varargout = cnmfe_b200_mex(command, args...)   -- see matlab/@Sources2D/*.m for the call sites.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include "mex.h"
#include "cnmfe_b200.h"

static void check(int rc, const char* where) {
    if (rc) mexErrMsgIdAndTxt("cnmfe:b200", "%s: %s", where, cnmfe_last_error());
}
static cnmfe_ctx* ctx_of(const mxArray* h) {
    if (!mxIsUint64(h) || mxGetNumberOfElements(h) != 1) mexErrMsgIdAndTxt("cnmfe:handle", "bad context handle");
    return reinterpret_cast<cnmfe_ctx*>(*static_cast<uint64_t*>(mxGetData(h)));
}
// MATLAB sparse -> 0-based int64 CSC (mwIndex is already 0-based)
static void csc_of(const mxArray* A, std::vector<int64_t>& jc, std::vector<int64_t>& ir) {
    if (!mxIsSparse(A)) mexErrMsgIdAndTxt("cnmfe:sparse", "sparse matrix expected");
    size_t K = mxGetN(A);
    const mwIndex* j = mxGetJc(A);
    const mwIndex* i = mxGetIr(A);
    jc.assign(j, j + K + 1);
    ir.assign(i, i + j[K]);
}
static double getfield_d(const mxArray* s, const char* f, double dflt) {
    const mxArray* v = mxGetField(s, 0, f);
    return (v && !mxIsEmpty(v)) ? mxGetScalar(v) : dflt;
}
static void deconv_opts_of(const mxArray* s, cnmfe_deconv_opts* o) {
    // struct produced by deconvolveCa's own parser fields (OASIS_matlab/deconvolveCa.m:212-230)
    cnmfe_deconv_defaults(o);
    if (!s || mxIsEmpty(s)) return;
    char buf[32];
    const mxArray* v;
    if ((v = mxGetField(s, 0, "type")) && !mxGetString(v, buf, sizeof buf)) o->type = strcmp(buf, "ar2") ? 1 : 2;
    if ((v = mxGetField(s, 0, "method")) && !mxGetString(v, buf, sizeof buf))
        o->method = !strcmp(buf, "foopsi") ? 0 : (!strcmp(buf, "thresholded") ? 2 : 1);
    o->optimize_b = (int)getfield_d(s, "optimize_b", 0);
    o->optimize_pars = (int)getfield_d(s, "optimize_pars", 0);
    o->maxIter = (int)getfield_d(s, "maxIter", 10);
    o->smin = getfield_d(s, "smin", 0);
    o->lambda = getfield_d(s, "lambda", 0);
    o->b = getfield_d(s, "b", 0);
    o->max_tau = getfield_d(s, "max_tau", 100);
    o->thresh_factor = getfield_d(s, "thresh_factor", 1.0);
    o->p_noise = getfield_d(s, "p_noise", 0.9999);
    if ((v = mxGetField(s, 0, "tau_range")) && mxGetNumberOfElements(v) == 2) {
        o->has_tau_range = 1; o->tau_range[0] = mxGetPr(v)[0]; o->tau_range[1] = mxGetPr(v)[1];
    }
}

// live contexts: released when MATLAB clears the MEX file or exits (SURVEY.md 8b ownership: ctx pointer handed to MATLAB as a
// uint64, mexLock'ed, released by 'destroy' / mexAtExit)
static std::vector<cnmfe_ctx*> g_live;
static void destroy_all(void) {
    for (cnmfe_ctx* h : g_live) cnmfe_destroy(h);
    g_live.clear();
}
static int field_int(const mxArray* s, const char* f, int dflt) {
    const mxArray* v = mxGetField(s, 0, f);
    return (v && !mxIsEmpty(v)) ? (int)mxGetScalar(v) : dflt;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("cnmfe:usage", "cnmfe_b200_mex(command, ...)");
    char cmd[64];
    mxGetString(prhs[0], cmd, sizeof cmd);
    const std::string c(cmd);
    static bool at_exit_registered = false;
    if (!at_exit_registered) { mexAtExit(destroy_all); at_exit_registered = true; }
    if (c == "create") {   // h = create(d1,d2,T,patch_pos(4 x np int32),block_pos,ring_radius,num_neighbors,device)
        cnmfe_ctx* h = nullptr;
        int np = (int)mxGetN(prhs[4]);
        check(cnmfe_create(&h, (int)mxGetScalar(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), np,
                           (const int32_t*)mxGetData(prhs[4]), (const int32_t*)mxGetData(prhs[5]), nullptr,
                           (int)mxGetScalar(prhs[6]), (int)mxGetScalar(prhs[7]), (int)mxGetScalar(prhs[8])), "create");
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *static_cast<uint64_t*>(mxGetData(plhs[0])) = reinterpret_cast<uint64_t>(h);
        g_live.push_back(h);
        mexLock();
    } else if (c == "destroy") {
        cnmfe_ctx* h = ctx_of(prhs[1]);
        for (size_t i = 0; i < g_live.size(); ++i)
            if (g_live[i] == h) { g_live.erase(g_live.begin() + i); cnmfe_destroy(h); mexUnlock(); break; }
    } else if (c == "upload_block") {   // upload_block(h, ipatch0, Yblock)  Yblock: nr_b x nc_b x T uint8/uint16
        int dt = mxIsUint8(prhs[3]) ? 0 : (mxIsUint16(prhs[3]) ? 1 : (mxIsSingle(prhs[3]) ? 2 : (mxIsDouble(prhs[3]) ? 3 : -1)));
        if (dt < 0) mexErrMsgIdAndTxt("cnmfe:dtype", "video must be uint8, uint16, or single/double holding integer counts");
        check(cnmfe_upload_block(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), mxGetData(prhs[3]), dt), "upload_block");
    } else if (c == "set_neurons" || c == "set_prev") {   // (h, A sparse d x K, C K x T)
        std::vector<int64_t> jc, ir;
        csc_of(prhs[2], jc, ir);
        int K = (int)mxGetN(prhs[2]);
        int rc = (c == "set_neurons")
                     ? cnmfe_set_neurons(ctx_of(prhs[1]), K, jc.data(), ir.data(), mxGetPr(prhs[2]), mxGetPr(prhs[3]))
                     : cnmfe_set_prev(ctx_of(prhs[1]), K, jc.data(), ir.data(), mxGetPr(prhs[2]), mxGetPr(prhs[3]));
        check(rc, cmd);
    } else if (c == "set_search") {   // (h, IND sparse logical d x K)
        std::vector<int64_t> jc, ir;
        csc_of(prhs[2], jc, ir);
        check(cnmfe_set_search(ctx_of(prhs[1]), (int)mxGetN(prhs[2]), jc.data(), ir.data()), cmd);
    } else if (c == "set_sn") {
        check(cnmfe_set_sn(ctx_of(prhs[1]), mxGetPr(prhs[2])), cmd);
    } else if (c == "set_options") {   // (h, struct: spatial_algorithm (code), maxIter, deconv_flag, bg_acceleration, background_model
                                       //  ('ring'|'svd'|'nmf'), nb, bg_ssub, deconv_options (struct) [, replicate_spatial_aprev_quirk, use_tensor_gram])
        if (nrhs < 3 || !mxIsStruct(prhs[2])) mexErrMsgIdAndTxt("cnmfe:usage", "set_options(h, options struct)");
        const mxArray* so = prhs[2];
        cnmfe_options o;
        cnmfe_options_defaults(&o);
        o.spatial_algorithm = field_int(so, "spatial_algorithm", o.spatial_algorithm);
        o.maxIter_temporal = field_int(so, "maxIter", o.maxIter_temporal);
        o.deconv_flag = field_int(so, "deconv_flag", o.deconv_flag);
        o.bg_acceleration = field_int(so, "bg_acceleration", o.bg_acceleration);
        o.nb = field_int(so, "nb", o.nb);
        o.bg_ssub = field_int(so, "bg_ssub", o.bg_ssub);
        o.replicate_spatial_aprev_quirk = field_int(so, "replicate_spatial_aprev_quirk", o.replicate_spatial_aprev_quirk);
        o.use_tensor_gram = field_int(so, "use_tensor_gram", o.use_tensor_gram);
        o.thresh_outlier = getfield_d(so, "thresh_outlier", o.thresh_outlier);     // NaN (CNMFSetParms default) = no outlier clamp
        const mxArray* bm = mxGetField(so, 0, "background_model");
        char buf[16];
        if (bm && !mxGetString(bm, buf, sizeof buf)) {
            if (!strcmp(buf, "ring")) o.background_model = 0;
            else if (!strcmp(buf, "svd")) o.background_model = 1;
            else if (!strcmp(buf, "nmf")) o.background_model = 2;
            else mexErrMsgIdAndTxt("cnmfe:options", "background_model '%s' unknown (ring, svd, nmf)", buf);
        }
        if (o.background_model != 0) o.bg_ssub = 1;
        deconv_opts_of(mxGetField(so, 0, "deconv_options"), &o.deconv);
        check(cnmfe_set_options(ctx_of(prhs[1]), &o), cmd);
    } else if (c == "set_use_c_hat") {   // (h, flag): third argument of update_temporal_parallel
        check(cnmfe_set_use_c_hat(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2])), cmd);
    } else if (c == "ssub_dims") {   // [d1s, d2s, nnb, r_shift, c_shift] = ssub_dims(h, ipatch0): coarse ring grid for bg_ssub > 1
        int d1s = 0, d2s = 0, nnb = 0;
        check(cnmfe_ssub_dims(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), &d1s, &d2s, &nnb, nullptr, nullptr), cmd);
        plhs[0] = mxCreateDoubleScalar(d1s);
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(d2s);
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(nnb);
        if (nlhs > 3) {
            plhs[3] = mxCreateNumericMatrix(1, nnb, mxINT32_CLASS, mxREAL);
            mxArray* cs = mxCreateNumericMatrix(1, nnb, mxINT32_CLASS, mxREAL);
            check(cnmfe_ssub_dims(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), &d1s, &d2s, &nnb, (int32_t*)mxGetData(plhs[3]), (int32_t*)mxGetData(cs)), cmd);
            if (nlhs > 4) plhs[4] = cs;
        }
    } else if (c == "set_bf") {   // (h, ipatch0, b (d_patch x nb), f (nb x T), b0 or []): svd / nmf background state
        const double* b = mxIsEmpty(prhs[3]) ? nullptr : mxGetPr(prhs[3]);
        const double* f = mxIsEmpty(prhs[4]) ? nullptr : mxGetPr(prhs[4]);
        const double* b0 = (nrhs > 5 && !mxIsEmpty(prhs[5])) ? mxGetPr(prhs[5]) : nullptr;
        check(cnmfe_set_bf(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), b, f, b0), cmd);
    } else if (c == "get_bf") {   // [b, f, b0] = get_bf(h, ipatch0, d_patch, nb, T)
        mwSize dp = (mwSize)mxGetScalar(prhs[3]), nb = (mwSize)mxGetScalar(prhs[4]), T = (mwSize)mxGetScalar(prhs[5]);
        plhs[0] = mxCreateDoubleMatrix(dp, nb, mxREAL);
        mxArray* f = mxCreateDoubleMatrix(nb, T, mxREAL);
        mxArray* b0 = mxCreateDoubleMatrix(dp, 1, mxREAL);
        check(cnmfe_get_bf(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), mxGetPr(plhs[0]), mxGetPr(f), mxGetPr(b0)), cmd);
        if (nlhs > 1) plhs[1] = f;
        if (nlhs > 2) plhs[2] = b0;
    } else if (c == "compute_rss") {   // rss = compute_rss(h, f0, f1, b0_map (d1 x d2), b0_new (d1 x d2), npatch): Sources2D.compute_RSS, per patch
        plhs[0] = mxCreateDoubleMatrix((mwSize)mxGetScalar(prhs[6]), 1, mxREAL);
        check(cnmfe_compute_rss(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), mxGetPr(prhs[4]), mxGetPr(prhs[5]),
                                mxGetPr(plhs[0])), cmd);
    } else if (c == "reconstruct_background") {   // Ybg = reconstruct_background(h, ipatch0, f0, f1, b0_map, b0_new, d_patch): d_patch x nframes
        const int f0 = (int)mxGetScalar(prhs[3]), f1 = (int)mxGetScalar(prhs[4]);
        plhs[0] = mxCreateDoubleMatrix((mwSize)mxGetScalar(prhs[7]), (mwSize)(f1 - f0 + 1), mxREAL);
        check(cnmfe_reconstruct_background(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), f0, f1, mxGetPr(prhs[5]), mxGetPr(prhs[6]),
                                           mxGetPr(plhs[0])), cmd);
    } else if (c == "estimate_noise") {   // sn = estimate_noise(h, f0, f1, d1, d2): per-pixel GetSn of the resident video (Sources2D.m:328-379)
        plhs[0] = mxCreateDoubleMatrix((mwSize)mxGetScalar(prhs[4]), (mwSize)mxGetScalar(prhs[5]), mxREAL);
        check(cnmfe_estimate_noise(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), mxGetPr(plhs[0])), cmd);
    } else if (c == "ring_offsets") {   // [r_shift, c_shift] = ring_offsets(h)
        int nnb = 0;
        check(cnmfe_ring_offsets(ctx_of(prhs[1]), &nnb, nullptr, nullptr), cmd);
        plhs[0] = mxCreateNumericMatrix(1, nnb, mxINT32_CLASS, mxREAL);
        plhs[1] = mxCreateNumericMatrix(1, nnb, mxINT32_CLASS, mxREAL);
        check(cnmfe_ring_offsets(ctx_of(prhs[1]), &nnb, (int32_t*)mxGetData(plhs[0]), (int32_t*)mxGetData(plhs[1])), cmd);
    } else if (c == "set_ring") {   // (h, ipatch0, Wslots (nnb x d_patch) or [], b0)
        const double* W = mxIsEmpty(prhs[3]) ? nullptr : mxGetPr(prhs[3]);
        const double* b0 = mxIsEmpty(prhs[4]) ? nullptr : mxGetPr(prhs[4]);
        check(cnmfe_set_ring(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), W, b0), cmd);
    } else if (c == "get_ring") {   // [Wslots, b0] = get_ring(h, ipatch0, nnb, nW, d_patch): nW = d_patch (bg_ssub = 1) or d1s*d2s (coarse grid)
        mwSize nnb = (mwSize)mxGetScalar(prhs[3]), nw = (mwSize)mxGetScalar(prhs[4]), dp = (mwSize)mxGetScalar(prhs[5]);
        plhs[0] = mxCreateDoubleMatrix(nnb, nw, mxREAL);
        mxArray* b0 = mxCreateDoubleMatrix(dp, 1, mxREAL);
        check(cnmfe_get_ring(ctx_of(prhs[1]), (int)mxGetScalar(prhs[2]), mxGetPr(plhs[0]), mxGetPr(b0)), cmd);
        if (nlhs > 1) plhs[1] = b0;
    } else if (c == "update_background") {
        check(cnmfe_update_background(ctx_of(prhs[1])), cmd);
    } else if (c == "update_spatial") {   // [vals, sn] = update_spatial(h, nnz(IND), update_sn, d1, d2)  -> values on find(IND) order
        const int usn = (nrhs > 3 && !mxIsEmpty(prhs[3])) ? (int)mxGetScalar(prhs[3]) : 0;
        check(cnmfe_update_spatial_ex(ctx_of(prhs[1]), usn), cmd);
        plhs[0] = mxCreateDoubleMatrix((mwSize)mxGetScalar(prhs[2]), 1, mxREAL);
        check(cnmfe_get_spatial(ctx_of(prhs[1]), mxGetPr(plhs[0])), "get_spatial");
        if (nlhs > 1) {                     // obj.P.sn after update_sn (update_spatial_parallel.m:191-194, :337-339)
            plhs[1] = mxCreateDoubleMatrix((mwSize)mxGetScalar(prhs[4]), (mwSize)mxGetScalar(prhs[5]), mxREAL);
            check(cnmfe_get_sn_map(ctx_of(prhs[1]), mxGetPr(plhs[1])), "get_sn_map");
        }
    } else if (c == "set_spatial") {
        check(cnmfe_set_spatial(ctx_of(prhs[1]), mxGetPr(prhs[2])), cmd);
    } else if (c == "update_temporal") {   // [C, C_raw, S, kernel_pars, neuron_sn] = update_temporal(h, K, T)
        check(cnmfe_update_temporal(ctx_of(prhs[1])), cmd);
        mwSize K = (mwSize)mxGetScalar(prhs[2]), T = (mwSize)mxGetScalar(prhs[3]);
        plhs[0] = mxCreateDoubleMatrix(K, T, mxREAL); plhs[1] = mxCreateDoubleMatrix(K, T, mxREAL);
        plhs[2] = mxCreateDoubleMatrix(K, T, mxREAL); plhs[3] = mxCreateDoubleMatrix(2, K, mxREAL);
        plhs[4] = mxCreateDoubleMatrix(K, 1, mxREAL);
        check(cnmfe_get_temporal(ctx_of(prhs[1]), mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]),
                                 mxGetPr(plhs[3]), mxGetPr(plhs[4])), "get_temporal");
    } else if (c == "deconvolve") {   // [c,s,b,pars,sn,smin,lam] = deconvolve(Y (T x N), opts struct, sn or [], pars (2 x N) or [])
        cnmfe_deconv_opts o;
        deconv_opts_of(prhs[2], &o);
        mwSize T = mxGetM(prhs[1]), N = mxGetN(prhs[1]);
        const double* sn = (nrhs > 3 && !mxIsEmpty(prhs[3])) ? mxGetPr(prhs[3]) : nullptr;
        const double* pars = (nrhs > 4 && !mxIsEmpty(prhs[4])) ? mxGetPr(prhs[4]) : nullptr;
        plhs[0] = mxCreateDoubleMatrix(T, N, mxREAL); plhs[1] = mxCreateDoubleMatrix(T, N, mxREAL);
        plhs[2] = mxCreateDoubleMatrix(N, 1, mxREAL); plhs[3] = mxCreateDoubleMatrix(2, N, mxREAL);
        plhs[4] = mxCreateDoubleMatrix(N, 1, mxREAL); plhs[5] = mxCreateDoubleMatrix(N, 1, mxREAL);
        plhs[6] = mxCreateDoubleMatrix(N, 1, mxREAL);
        check(cnmfe_deconvolve(mxGetPr(prhs[1]), (int)T, (int)N, &o, sn, pars, mxGetPr(plhs[0]), mxGetPr(plhs[1]),
                               mxGetPr(plhs[2]), mxGetPr(plhs[3]), mxGetPr(plhs[4]), mxGetPr(plhs[5]), mxGetPr(plhs[6]), 0), cmd);
    } else if (c == "get_sn") {   // sn = get_sn(Y (T x N))
        plhs[0] = mxCreateDoubleMatrix(mxGetN(prhs[1]), 1, mxREAL);
        check(cnmfe_get_sn(mxGetPr(prhs[1]), (int)mxGetM(prhs[1]), (int)mxGetN(prhs[1]), mxGetPr(plhs[0]), 0), cmd);
    } else if (c == "post_process_spatial") {   // A_ = post_process_spatial(A sparse d x K, d1, d2): connectivity_constraint per neuron
        const mwSize K = mxGetN(prhs[1]);
        const mwIndex* jcm = mxGetJc(prhs[1]); const mwIndex* irm = mxGetIr(prhs[1]);
        std::vector<int64_t> jc(jcm, jcm + K + 1), ir(irm, irm + jcm[K]);
        plhs[0] = mxDuplicateArray(prhs[1]);               // same pattern; removed entries become explicit zeros
        check(cnmfe_connectivity_constraint((int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), (int)K, jc.data(), ir.data(),
                                            mxGetPr(plhs[0]), 0.01, 5), cmd);
    } else if (c == "search_location_dilate") {   // [jc, ir] = search_location_dilate(A sparse d x K, d1, d2, nrgthr, nb, bSiz): 0-based CSC pattern
        const mwSize K = mxGetN(prhs[1]);
        const mwIndex* jcm = mxGetJc(prhs[1]); const mwIndex* irm = mxGetIr(prhs[1]);
        std::vector<int64_t> jc(jcm, jcm + K + 1), ir(irm, irm + jcm[K]);
        const int d1 = (int)mxGetScalar(prhs[2]), d2 = (int)mxGetScalar(prhs[3]);
        const double nrgthr = mxGetScalar(prhs[4]);
        const int nb = (int)mxGetScalar(prhs[5]), bSiz = (int)mxGetScalar(prhs[6]);
        int64_t cap = 1;
        for (mwSize k = 0; k < K; ++k) {
            int r0 = d1, r1 = -1, c0 = d2, c1 = -1;
            for (mwIndex e = jcm[k]; e < jcm[k + 1]; ++e) {
                const int r = (int)(irm[e] % d1), cc = (int)(irm[e] / d1);
                r0 = r < r0 ? r : r0; r1 = r > r1 ? r : r1; c0 = cc < c0 ? cc : c0; c1 = cc > c1 ? cc : c1;
            }
            const int64_t hgt = r1 >= 0 ? r1 - r0 + 1 : 1, wid = r1 >= 0 ? c1 - c0 + 1 : 1;
            cap += (hgt + 4 + 2 * bSiz) * (wid + 4 + 2 * bSiz);
        }
        std::vector<int64_t> ojc(K + 1), oir((size_t)cap);
        check(cnmfe_search_location_dilate(d1, d2, (int)K, jc.data(), ir.data(), mxGetPr(prhs[1]), nrgthr, nb, bSiz, ojc.data(), oir.data(), cap), cmd);
        plhs[0] = mxCreateDoubleMatrix(K + 1, 1, mxREAL);
        plhs[1] = mxCreateDoubleMatrix((mwSize)ojc[K], 1, mxREAL);
        for (mwSize k = 0; k <= K; ++k) mxGetPr(plhs[0])[k] = (double)ojc[k];
        for (int64_t e = 0; e < ojc[K]; ++e) mxGetPr(plhs[1])[e] = (double)oir[e];
    } else if (c == "circular_constraints") {   // A_ = circular_constraints(A sparse d x K, d1, d2): circular_constraints.m per neuron
        const mwSize K = mxGetN(prhs[1]);
        const mwIndex* jcm = mxGetJc(prhs[1]); const mwIndex* irm = mxGetIr(prhs[1]);
        std::vector<int64_t> jc(jcm, jcm + K + 1), ir(irm, irm + jcm[K]);
        const int d1 = (int)mxGetScalar(prhs[2]), d2 = (int)mxGetScalar(prhs[3]);
        int64_t cap = 1;                                   // the result lives on the bounding boxes of the footprints
        for (mwSize k = 0; k < K; ++k) {
            int r0 = d1, r1 = -1, c0 = d2, c1 = -1;
            for (mwIndex e = jcm[k]; e < jcm[k + 1]; ++e) {
                const int r = (int)(irm[e] % d1), cc = (int)(irm[e] / d1);
                r0 = r < r0 ? r : r0; r1 = r > r1 ? r : r1; c0 = cc < c0 ? cc : c0; c1 = cc > c1 ? cc : c1;
            }
            if (r1 >= 0) cap += (int64_t)(r1 - r0 + 1) * (c1 - c0 + 1);
        }
        std::vector<int64_t> ojc(K + 1), oir((size_t)cap);
        std::vector<double> opr((size_t)cap);
        check(cnmfe_circular_constraints(d1, d2, (int)K, jc.data(), ir.data(), mxGetPr(prhs[1]), ojc.data(), oir.data(), opr.data(), cap), cmd);
        plhs[0] = mxCreateDoubleMatrix(K + 1, 1, mxREAL);              // jc (0-based), ir (0-based), values
        plhs[1] = mxCreateDoubleMatrix((mwSize)ojc[K], 1, mxREAL);
        plhs[2] = mxCreateDoubleMatrix((mwSize)ojc[K], 1, mxREAL);
        for (mwSize k = 0; k <= K; ++k) mxGetPr(plhs[0])[k] = (double)ojc[k];
        for (int64_t e = 0; e < ojc[K]; ++e) { mxGetPr(plhs[1])[e] = (double)oir[e]; mxGetPr(plhs[2])[e] = opr[e]; }
    } else if (c == "search_location") {   // [jc, ir] = search_location(A sparse d x K, d1, d2, min_size, max_size, dist): 0-based CSC pattern
        const mwSize K = mxGetN(prhs[1]);
        const mwIndex* jcm = mxGetJc(prhs[1]); const mwIndex* irm = mxGetIr(prhs[1]);
        std::vector<int64_t> jc(jcm, jcm + K + 1), ir(irm, irm + jcm[K]);
        const double mn = mxGetScalar(prhs[4]), mx = mxGetScalar(prhs[5]), dist = mxGetScalar(prhs[6]);
        const int64_t reach = (int64_t)std::ceil(dist * (mx > mn ? mx : mn));
        const int64_t cap = (int64_t)K * (2 * reach + 2) * (2 * reach + 2) + 1;
        std::vector<int64_t> ojc(K + 1), oir((size_t)cap);
        check(cnmfe_search_location_ellipse((int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), (int)K, jc.data(), ir.data(),
                                            mxGetPr(prhs[1]), mn, mx, dist, ojc.data(), oir.data(), cap), cmd);
        plhs[0] = mxCreateDoubleMatrix(K + 1, 1, mxREAL);
        plhs[1] = mxCreateDoubleMatrix((mwSize)ojc[K], 1, mxREAL);
        for (mwSize k = 0; k <= K; ++k) mxGetPr(plhs[0])[k] = (double)ojc[k];
        for (int64_t e = 0; e < ojc[K]; ++e) mxGetPr(plhs[1])[e] = (double)oir[e];
    } else {
        mexErrMsgIdAndTxt("cnmfe:usage", "unknown command %s", cmd);
    }
}
