/*
 * cnmfe_b200.h -- C ABI of libcnmfe_b200.so: the B200-native (sm_100a) implementation of CNMF-E's
 * alternating-update hot path (ring background regression -> spatial update -> temporal update + OASIS).
 *
 * The reference (zhoupc/CNMF_E) is MATLAB and has no FFI for this path; its seam is the Sources2D method
 * surface.  Each entry point below names the reference interface it replaces (file:line under
 * /root/reference).  A MEX gateway (matlab/cnmfe_b200_mex.cpp) binds these 1:1; INTEGRATION.md shows the
 * MATLAB-side stubs.  Conventions:
 *   - plain pointers and sizes only; all HOST arrays unless the name ends in _dev;
 *   - dense matrices are MATLAB column-major doubles; sparse matrices are MATLAB CSC (jc, ir, pr) with
 *     0-based indices (exactly mxGetJc / mxGetIr / mxGetPr);
 *   - pixel linear index = r + c*d1 (0-based, MATLAB order); positions [r0 r1 c0 c1] are 1-based inclusive,
 *     exactly as stored in mat_data.patch_pos / block_pos (endoscope/distribute_data.m:163-173);
 *   - every function returns 0 on success, nonzero on error; cnmfe_last_error() gives the message
 *     (the MEX gateway forwards it through mexErrMsgIdAndTxt, cf. utilities/graph_conn_comp_mex.cpp:44-53);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CNMFE_B200_H
#define CNMFE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cnmfe_ctx cnmfe_ctx;

/* deconvolveCa options (OASIS_matlab/deconvolveCa.m:208-230 defaults in cnmfe_deconv_defaults). */
typedef struct cnmfe_deconv_opts {
    int type;            /* 1 = 'ar1', 2 = 'ar2' */
    int method;          /* 0 = 'foopsi', 1 = 'constrained', 2 = 'thresholded' */
    int optimize_b;
    int optimize_pars;
    int maxIter;         /* default 10 */
    int has_tau_range;
    double smin;         /* <0: multiple of the noise level (deconvolveCa.m:116-118,125-127) */
    double lambda;
    double b;            /* initial baseline */
    double max_tau;      /* default 100 */
    double tau_range[2];
    double thresh_factor;/* default 1.0 */
    double p_noise;      /* default 0.9999 */
} cnmfe_deconv_opts;

/* Options of the three updates (ca_source_extraction/CNMFSetParms.m:286-308 + demo overrides). */
typedef struct cnmfe_options {
    int spatial_algorithm;   /* 0 'hals', 1 'hals_thresh', 2 'nnls', 3 'lars' (update_spatial_parallel.m:202-212) */
    int maxIter_temporal;    /* options.maxIter, default 5 (update_temporal_parallel.m:60) */
    int deconv_flag;         /* options.deconv_flag */
    int bg_acceleration;     /* options.bg_acceleration -> fit_ring_model(..., with_projection) */
    int replicate_spatial_aprev_quirk; /* 1 = select A_prev neurons by HALO only in the spatial update, as
                                          update_spatial_parallel.m:82-98 does (mask==1 after the patch is set to 2) */
    int use_tensor_gram;     /* 1 = tcgen05 INT8 kernel for the ring second moments, 0 = SIMT reference kernel */
    cnmfe_deconv_opts deconv;
    int background_model;    /* 0 = 'ring' (1p, bg_ssub = 1), 1 = 'svd' (2p default, endoscope/fit_svd_model.m),
                                2 = 'nmf' (endoscope/fit_nmf_model.m: nnmf alternating least squares on Y - A*C from a fixed
                                hash start instead of MATLAB's random stream; BG subtraction Y - b*f,
                                update_spatial_parallel.m:179-182) */
    int nb;                  /* options.nb: number of svd background components (default 1) */
    int bg_ssub;             /* options.bg_ssub (ring model): 1, or > 1 = ring weights on the ceil(block/bg_ssub) grid
                                (demo_large_data_1p.m:30 uses 2); changing it re-initialises W (update_background_parallel.m:70-118) */
    double thresh_outlier;   /* options.thresh_outlier (CNMFSetParms default NaN = off): outlier clamp + frame selection of
                                fit_ring_model.m:48-70.  Non-NaN selects the explicit fp64 path (Bf materialised; ring model, bg_ssub = 1) */
} cnmfe_options;

const char* cnmfe_last_error(void);
void cnmfe_deconv_defaults(cnmfe_deconv_opts* o);
void cnmfe_options_defaults(cnmfe_options* o);
/* number of kernels of this library launched so far in this process */
unsigned long long cnmfe_launch_count(void);

/* ---- stand-alone trace routines (host arrays; batch of N traces, each trace contiguous: Y is T x N) --------- */

/* [c,s,options] = deconvolveCa(y, ...)  (OASIS_matlab/deconvolveCa.m:1; also @Sources2D/deconvTemporal.m:45-60).
 * sn_in / pars_in may be NULL (estimate: GetSn / estimate_time_constant); pars_in is 2 x N, zeros = estimate.
 * Outputs (any may be NULL): c,s (T x N), b,sn,smin,lam (N), pars (2 x N). */
int cnmfe_deconvolve(const double* Y, int T, int N, const cnmfe_deconv_opts* opts, const double* sn_in,
                     const double* pars_in, double* c, double* s, double* b, double* pars, double* sn,
                     double* smin, double* lam, int device);

/* same on DEVICE arrays (all pointers on `device`; Y, c, s are N traces of T contiguous doubles; sn_in, pars_in (2 per trace),
 * outs (6 per trace: b, g1, g2, smin, lambda, sn) may be NULL).  Asynchronous on the legacy default stream. */
int cnmfe_deconvolve_dev(const double* Y_dev, int T, int N, const cnmfe_deconv_opts* opts, const double* sn_dev,
                         const double* pars_dev, double* c_dev, double* s_dev, double* outs_dev, int device);

/* sn = GetSn(Y)  (OASIS_matlab/functions/GetSn.m:1; logmexp over [0.25,0.5]).  Y is T x N, sn has N entries. */
int cnmfe_get_sn(const double* Y, int T, int N, double* sn, int device);

/* [C,C_raw,results_deconv] = HALS_temporal(Y,A,C,maxIter,deconv_options) (utilities/HALS_temporal.m:1) given the
 * projections U = A'*Y (K x T col-major), V = A'*A (K x K dense col-major).  C is updated in place (K x T);
 * outputs C_raw, S (K x T), sn (K), kernel_pars (2 x K).  deconv may be NULL (no deconvolution). */
int cnmfe_hals_temporal_uv(const double* U, const double* V, int K, int T, double* C, int maxIter,
                           const cnmfe_deconv_opts* deconv, double* C_raw, double* S, double* sn,
                           double* kernel_pars, int device);

/* ---- resident-video context: Sources2D.update_{background,spatial,temporal}_parallel ------------------------- */

/* Geometry as produced by distribute_data.m:163-173.  patch_pos / block_pos: 4 x npatch int32 (col-major, 1-based
 * inclusive [r0 r1 c0 c1]), patches in MATLAB linear order.  owned: npatch flags (NULL = all) -- the patches this
 * process/GPU holds (multi-GPU sharding).  ring_radius, num_neighbors (0 = all) as in
 * CNMFSetParms.m:289-291 (bg_ssub = 1). */
int cnmfe_create(cnmfe_ctx** ctx, int d1, int d2, int T, int npatch, const int32_t* patch_pos,
                 const int32_t* block_pos, const uint8_t* owned, int ring_radius, int num_neighbors, int device);
void cnmfe_destroy(cnmfe_ctx* ctx);
int cnmfe_set_options(cnmfe_ctx* ctx, const cnmfe_options* opts);

/* get_patch_data(mat_data, patch_pos, frame_range, true) (endoscope/get_patch_data.m:1): hand the block of patch
 * `ipatch` (nr_block x nc_block x T, column-major, native dtype) to the device once; it stays resident.
 * dtype: 0 = uint8, 1 = uint16 (the dtypes distribute_data.m:144-147 writes for the demos), 2 = single, 3 = double -- accepted when
 * every value is an integer count in [0, 65535] (converted exactly; the source class of a TIFF may be 'single'), refused otherwise. */
int cnmfe_upload_block(cnmfe_ctx* ctx, int ipatch, const void* Y, int dtype);
/* same, Y already in device memory (frame-major nr_block*nc_block per frame, as MATLAB lays it out) */
int cnmfe_upload_block_dev(cnmfe_ctx* ctx, int ipatch, const void* Y_dev, int dtype);

/* obj.A (d x K sparse), obj.C (K x T)   (Sources2D.m:11-13).  For both setters: a NULL (jc, ir, pr) triple or a NULL C
 * keeps the part the context already holds (K must match) -- lets a caller re-send only what changed */
int cnmfe_set_neurons(cnmfe_ctx* ctx, int K, const int64_t* A_jc, const int64_t* A_ir, const double* A_pr,
                      const double* C);
/* obj.A_prev, obj.C_prev (Sources2D.m:14-15); cnmfe_update_background snapshots them itself (:316-317) */
int cnmfe_set_prev(cnmfe_ctx* ctx, int K, const int64_t* A_jc, const int64_t* A_ir, const double* A_pr,
                   const double* C);
/* IND = determine_search_location(...) (update_spatial_parallel.m:66): d x K logical sparse */
int cnmfe_set_search(cnmfe_ctx* ctx, int K, const int64_t* IND_jc, const int64_t* IND_ir);
/* obj.P.sn (d1 x d2) */
int cnmfe_set_sn(cnmfe_ctx* ctx, const double* sn);
/* obj.W{ipatch}, obj.b0{ipatch}: ring weights in slot form, W[i + p*nnb] = weight of patch pixel p for ring
 * offset i (cnmfe_ring_offsets), i.e. an nnb x d_patch column-major matrix; entries whose neighbour falls outside
 * the FOV are ignored.  NULL W and NULL b0 = uniform initialisation (initComponents_parallel.m:213-236); NULL W with a
 * b0 replaces only the offsets and keeps the weights. */
int cnmfe_ring_offsets(cnmfe_ctx* ctx, int* nnb, int32_t* r_shift, int32_t* c_shift);
int cnmfe_set_ring(cnmfe_ctx* ctx, int ipatch, const double* W_slots, const double* b0);
int cnmfe_get_ring(cnmfe_ctx* ctx, int ipatch, double* W_slots, double* b0);
/* with bg_ssub > 1 the ring weights handled by cnmfe_set_ring / cnmfe_get_ring live on the coarse grid: nnb x (d1s*d2s)
 * slots; this returns the coarse dimensions and ring offsets of patch ipatch */
int cnmfe_ssub_dims(cnmfe_ctx* ctx, int ipatch, int* d1s, int* d2s, int* nnb, int32_t* r_shift, int32_t* c_shift);
/* obj.b{ipatch} (d_patch x nb), obj.f{ipatch} (nb x T), obj.b0{ipatch} of the svd background model (column-major) */
int cnmfe_set_bf(cnmfe_ctx* ctx, int ipatch, const double* b, const double* f, const double* b0);
int cnmfe_get_bf(cnmfe_ctx* ctx, int ipatch, double* b, double* f, double* b0);

/* update_background_parallel(obj, use_parallel) (@Sources2D/update_background_parallel.m:1) for the configured background
 * model (ring with bg_ssub >= 1, svd, nmf).  Result stays on the device (W, b0 / b, f, b0; A_prev<-A, C_prev<-C); fetch with
 * cnmfe_get_ring / cnmfe_get_bf. */
int cnmfe_update_background(cnmfe_ctx* ctx);
/* update_spatial_parallel(obj, use_parallel, update_sn=false) (@Sources2D/update_spatial_parallel.m:1) up to
 * (not including) post_process_spatial (:341, host/MATLAB).  New A lives on the search pattern. */
int cnmfe_update_spatial(cnmfe_ctx* ctx);
/* same with update_sn = true: obj.P.sn is re-estimated per pixel from the BG-subtracted video (GetSn,
 * update_spatial_parallel.m:191-194) before the solve; read it back with cnmfe_get_sn_map (d1 x d2) */
int cnmfe_update_spatial_ex(cnmfe_ctx* ctx, int update_sn);
int cnmfe_get_sn_map(cnmfe_ctx* ctx, double* sn);
/* sn = estimate_noise(obj, frame_range, 'psd') (@Sources2D/Sources2D.m:328-379) from the RESIDENT video: per-pixel GetSn of
 * the raw frames [f0, f1] (1-based inclusive; the reference's default is [1, min(T, 3000)]) on the pixels of the owned patches;
 * sn is d1 x d2 column-major, other pixels untouched.  The reference's block-border quirk (:369-374) is applied by the caller. */
int cnmfe_estimate_noise(cnmfe_ctx* ctx, int f0, int f1, double* sn);
/* A on the search pattern: values aligned with (IND_jc, IND_ir) given to cnmfe_set_search */
int cnmfe_get_spatial(cnmfe_ctx* ctx, double* A_on_IND);
/* replace obj.A by values on the search pattern (e.g. after the cross-GPU exchange of the patches' rows, or after
 * post_process_spatial on the host); obj.C on the device is untouched */
int cnmfe_set_spatial(cnmfe_ctx* ctx, const double* A_on_IND);
/* update_temporal_parallel(obj, use_parallel, use_c_hat=true) (@Sources2D/update_temporal_parallel.m:1).
 * phase 1: per-patch HALS_temporal -> energy-weighted sums; phase 2 (after an optional cross-GPU all-reduce of the
 * buffers exposed by cnmfe_temporal_merge_buffers): C_raw = num/den, deconvTemporal (deconvTemporal.m:1). */
int cnmfe_update_temporal_patches(cnmfe_ctx* ctx);
int cnmfe_temporal_merge_buffers(cnmfe_ctx* ctx, double** num_dev /* K*T, [k][t] */, double** den_dev /* K */);
int cnmfe_update_temporal_finish(cnmfe_ctx* ctx);
int cnmfe_update_temporal(cnmfe_ctx* ctx);   /* = patches + finish (single process) */
/* use_c_hat argument of update_temporal_parallel (default 1); 0 = fast_temporal (update_temporal_parallel.m:314-337):
 * mean fluorescence over the pixels with A >= 0.5 max(A) instead of the HALS sweeps */
int cnmfe_set_use_c_hat(cnmfe_ctx* ctx, int use_c_hat);
/* Layout of the K x T arrays (C, C_prev, C_raw, S) at this boundary: 0 (default) = MATLAB column-major (k fastest);
 * 1 = trace-contiguous [K][T] (NumPy C order; what the device uses) -- saves two transpositions per array per call */
int cnmfe_set_trace_major(cnmfe_ctx* ctx, int on);
/* Page-lock / release a caller-owned host range (cudaHostRegister) so the big transfers (ring weights, video blocks)
 * run at PCIe rate; optional -- every entry point also takes pageable memory */
int cnmfe_host_register(void* p, size_t bytes);
int cnmfe_host_unregister(void* p);
/* obj.C, obj.C_raw, obj.S (K x T col-major), obj.P.kernel_pars (2 x K), obj.P.neuron_sn (K); any may be NULL */
int cnmfe_get_temporal(cnmfe_ctx* ctx, double* C, double* C_raw, double* S, double* kernel_pars,
                       double* neuron_sn);
/* Multi-GPU split of the final deconvTemporal (opt-in, see csrc/ctx.cu): this rank finishes only the traces
 * [k0, k1) and zeroes the other rows of C, C_raw, S and of the per-trace outputs; a SUM all-reduce of the four device buffers
 * over ranks with disjoint ranges then gives every rank the full result. */
int cnmfe_update_temporal_finish_part(cnmfe_ctx* ctx, int k0, int k1);
int cnmfe_temporal_state_buffers(cnmfe_ctx* ctx, double** C_dev, double** Craw_dev, double** S_dev, double** outs_dev);
/* ---- host-side brackets of the spatial update (plain C++, no device work, ctx-free) ----
 * post_process_spatial with spatial_constraints.connected (@Sources2D/post_process_spatial.m:19-32 ->
 * endoscope/connectivity_constraint.m): per column of A (d1*d2 x K CSC) 5x5 grey opening, threshold thr*max (0.01),
 * 4-connected labelling, keep the component of the arg-max pixel.  pr is edited in place: removed entries become 0. */
int cnmfe_connectivity_constraint(int d1, int d2, int K, const int64_t* jc, const int64_t* ir, double* pr,
                                  double thr, int sz);
/* post_process_spatial with spatial_constraints.circular (endoscope/circular_constraints.m): result as a new CSC matrix
 * (the median filter can create non-zeros next to the footprint); cap >= sum over neurons of the bounding-box areas
 * (<= d1*d2*K) -- the reference applies it after connectivity_constraint (post_process_spatial.m:24-30). */
int cnmfe_circular_constraints(int d1, int d2, int K, const int64_t* jc, const int64_t* ir, const double* pr,
                               int64_t* out_jc, int64_t* out_ir, double* out_pr, int64_t cap);
/* determine_search_location(A, 'ellipse', params) (utilities/determine_search_location.m:57-92): search mask as a CSC
 * pattern (out_jc K+1, out_ir sorted).  out_ir needs cap >= K * (2*ceil(dist*max_size) + 2)^2 entries.
 * Defaults of the reference: min_size 3, max_size 8, dist 3. */
int cnmfe_search_location_ellipse(int d1, int d2, int K, const int64_t* jc, const int64_t* ir, const double* pr,
                                  double min_size, double max_size, double dist, int64_t* out_jc, int64_t* out_ir,
                                  int64_t cap);
/* determine_search_location(A, 'dilate', params) (utilities/determine_search_location.m:93-99 -> threshold_components.m, then
 * imdilate with strel('disk', bSiz, 0)).  Reference defaults: nrgthr 0.9999, nb 1 (the last nb columns are not thresholded),
 * bSiz 3.  cap >= sum over neurons of (bbox height + 4 + 2 bSiz) * (bbox width + 4 + 2 bSiz). */
int cnmfe_search_location_dilate(int d1, int d2, int K, const int64_t* jc, const int64_t* ir, const double* pr,
                                 double nrgthr, int nb, int bSiz, int64_t* out_jc, int64_t* out_ir, int64_t cap);
/* [l, c] = graph_connected_comp(sA) (utilities/graph_connected_comp.m:26 -> utilities/graph_conn_comp_mex.cpp, the
 * reference's only native file; used by the merge routines): labels 1..c in the order of each component's smallest node.
 * Host-side, ctx-free. */
int cnmfe_graph_conn_comp(int n, const int64_t* jc, const int64_t* ir, uint32_t* labels, int* ncomp);
/* test hook (CPU, no device work): the block / patch-local view the host planning derives from a MATLAB CSC matrix
 * (patch_pos / block_pos 1-based inclusive [r0 r1 c0 c1]); see csrc/ctx.cu build_local */
int cnmfe_debug_local_view(int d1, int d2, const int* patch_pos, const int* block_pos, int K, const int64_t* jc,
                           const int64_t* ir, const double* pr, int sel, int rows, const int64_t* vjc, const int64_t* vir,
                           const double* vpr, int* n_local, int* n_kept, int* ids, int* ptr, int* col, double* val,
                           int64_t* entry_src, int* cptr, int* crow, double* cval, int* bbox);
/* [RSS_total, RSS] = compute_RSS(obj, frame_range) (@Sources2D/Sources2D.m:1358-1510) and Ybg = reconstruct_background(obj,
 * frame_range) (:1247-1356) for the ring model with bg_ssub = 1, from the resident video and the state the context holds (A, C,
 * A_prev, C_prev, W).  f0, f1: 1-based inclusive frames.  b0_map = obj.reconstruct_b0() and b0_new_map = obj.b0_new, both d1 x d2
 * column-major (the halo of a block reads b0 of neighbouring patches, which another rank may own).  rss: npatch entries, written for
 * the owned patches.  Ybg: d_patch x (f1-f0+1) in the boundary layout (column-major, or [pixel][frame] with cnmfe_set_trace_major). */
int cnmfe_compute_rss(cnmfe_ctx* ctx, int f0, int f1, const double* b0_map, const double* b0_new_map, double* rss);
int cnmfe_reconstruct_background(cnmfe_ctx* ctx, int ipatch, int f0, int f1, const double* b0_map, const double* b0_new_map,
                                 double* Ybg);
/* block the host until all queued device work of ctx is done */
int cnmfe_sync(cnmfe_ctx* ctx);
/* CUDA-event timing of the kernels on the ctx stream: begin/end bracket a region, returns milliseconds */
int cnmfe_timer_begin(cnmfe_ctx* ctx);
int cnmfe_timer_end(cnmfe_ctx* ctx, float* ms);
/* per-phase device time of the last update call (ms): [0] gram (second moments), [1] ring assemble+solve,
 * [2] projections, [3] spatial solve, [4] temporal sweeps, [5] deconvTemporal, [6] other */
int cnmfe_last_phase_ms(cnmfe_ctx* ctx, float* ms7);

/* ---- diagnostics (used by tests/ and bench.py) ---------------------------------------------------------------- */
/* banded second moments S2[q][id] (nr_block*nc_block x ND doubles, ND = 2rr(4rr+1)+2rr+1) of block ipatch computed by
 * the tcgen05 kernel (use_tensor=1) or the exact SIMT kernel (0); entries whose neighbour is outside the block are 0 */
int cnmfe_debug_second_moments(cnmfe_ctx* ctx, int ipatch, int use_tensor, double* out);
/* 1 if the last cnmfe_update_background used the tensor-core kernel for the second moments */
int cnmfe_last_gram_was_tensor(cnmfe_ctx* ctx);
/* number of frames the last ring fit used (T, or ceil(T/k) with the frame stride k of fit_ring_model.m:84-90) */
int cnmfe_last_gram_frames(cnmfe_ctx* ctx);
/* number of pixels whose ring weights the last cnmfe_update_background refitted (ind_active of fit_ring_model.m:25-29) */
long long cnmfe_last_active_pixels(cnmfe_ctx* ctx);
/* alternating-least-squares iterations the last nmf background fit took (nnmf stops on its TolX / TolFun tests), summed over patches */
int cnmfe_last_nmf_iterations(cnmfe_ctx* ctx);
/* rows of the resident video: out[i][0..T) = Y(block pixel idx[i], :) of block ipatch, idx = r + c*nr_block (0-based) */
int cnmfe_debug_video_rows(cnmfe_ctx* ctx, int ipatch, int n, const int32_t* idx, uint16_t* out);
/* the merged C_raw (update_temporal_parallel.m:269-280) as it entered the final deconvTemporal (before deconvTemporal.m:84
 * subtracts the baseline); K x T in the boundary layout.  A checker can re-run deconvolveCa on exactly this input. */
int cnmfe_get_merged_craw(cnmfe_ctx* ctx, double* Craw_in);

#ifdef __cplusplus
}
#endif
#endif /* CNMFE_B200_H */
