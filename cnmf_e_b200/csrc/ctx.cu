// ctx.cu -- resident-video context: Sources2D.update_{background,spatial,temporal}_parallel behind the C ABI.
// Host code here only plans (index lists, neuron selections -- the role of the reference's host gather at
// update_*_parallel.m:69-100) and launches; all arithmetic on the video runs in the kernels.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "internal.h"
#include "kernels_video.cuh"
#include "kernels_ring.cuh"
#include "kernels_ring_mma.cuh"
#include "kernels_update.cuh"
#include "kernels_svd.cuh"
#include "kernels_ssub.cuh"

namespace cnmfe {
int ring_s2_tensor(const uint8_t* hi, const uint8_t* lo, int nrb, int ncb, int T, int Tpad, int rr, double* S2,
                   cudaStream_t st);   // ring_tc.cu (tcgen05 INT8); returns 1 if unsupported for this shape
}

using namespace cnmfe;

namespace {

struct HostCsc {
    int K = 0;
    std::vector<int64_t> jc, ir;
    std::vector<double> pr;
    void set(int K_, const int64_t* jc_, const int64_t* ir_, const double* pr_) {
        K = K_;
        jc.assign(jc_, jc_ + K_ + 1);
        ir.assign(ir_, ir_ + jc_[K_]);
        if (pr_) pr.assign(pr_, pr_ + jc_[K_]); else pr.assign((size_t)jc_[K_], 1.0);
    }
};

struct Rect { int r0, r1, c0, c1; };   // 0-based inclusive, FOV coordinates

struct LocalSparse {
    std::vector<int> ids;                       // local -> global neuron
    std::vector<int> ptr, col; std::vector<double> val;      // rows by pixel (block or patch pixel index)
    std::vector<int> cptr, crow; std::vector<double> cval;   // columns by local neuron (rows sorted)
    std::vector<int> bbox;                      // [K][4] r0,r1,c0,c1 in block coords (tight)
    std::vector<int64_t> entry_src;             // csc entry index of each row-entry (for IND bookkeeping)
    int K() const { return (int)ids.size(); }
};

struct Patch {
    Rect patch, block;
    int nr, nc, nrb, ncb, dp, db;
    bool owned = false, uploaded = false, w_uniform = true;
    uint16_t* Yt = nullptr; uint8_t* hi = nullptr; uint8_t* lo = nullptr;
    uint8_t* hi_k = nullptr; uint8_t* lo_k = nullptr; int planes_kf = 0, Tpad_k = 0;   // byte planes of every kf-th frame (tensor path, kf > 1)
    double* Ysum = nullptr; double* Ymean = nullptr;
    double* W = nullptr; double* b0 = nullptr;
    double* bsvd = nullptr; double* fsvd = nullptr;   // svd background: b [nb][dp] (col-major dp x nb), f [nb][T]
    int nb_alloc = 0; bool b_zero = true;
    // spatial result bookkeeping
    std::vector<int64_t> ind_entry;   // for each pattern entry (patch CSR order): index into the user's IND csc
    RingGeom geom;
    // local view of A_prev on the block (SEL_SUM_BLOCK, ROWS_BLOCK): built by the ring BG update (A_prev = A there) and reused
    // by the spatial and temporal updates of the same iteration
    LocalSparse lp_cache; bool lp_valid = false;
    // centred projections M = (Y - Ybar) * Cc' computed by the ring BG update (all frames), kept for the spatial update of the same
    // iteration: C does not change in between (demo_large_data_1p.m:199-201), so the spatial update would stream the video again for
    // the very same numbers.  [db][mc_K] doubles; valid while c->C_version == mc_Cver.
    double* Mcache = nullptr; size_t Mcache_cap = 0; bool mc_valid = false; unsigned long long mc_Cver = 0;
    int mc_K = 0; std::vector<int> mc_ids, mc_bbox;
};

// Local (per patch) sparse view of a d x K matrix.

struct Scratch {
    char* base = nullptr; size_t cap = 0, off = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        // head-room: the need moves by a few bytes with nnz(A) from call to call, and re-allocating a multi-GB arena
        // costs tens of milliseconds
        size_t want = bytes + std::max<size_t>(bytes / 32, (size_t)64 << 20);
        if (cudaMalloc((void**)&base, want) != cudaSuccess) {
            (void)cudaGetLastError();
            want = bytes;
            CNMFE_CUDA_OK(cudaMalloc((void**)&base, want));
        }
        cap = want;
        return 0;
    }
    void reset() { off = 0; }
    template <class T> T* take(size_t n) {
        size_t b = (n * sizeof(T) + 255) / 256 * 256;
        if (off + b > cap) return nullptr;
        T* p = (T*)(base + off);
        off += b;
        return p;
    }
};

}  // namespace

struct cnmfe_ctx {
    int device = 0, d1 = 0, d2 = 0, T = 0, Tpad = 0, npatch = 0;
    int ring_radius = 0, rr = 0, nnb = 0;
    std::vector<int> off_r, off_c;
    int* d_off_r = nullptr; int* d_off_c = nullptr;
    int* d_groups = nullptr; int ngroups = 0;
    std::vector<Patch> patches;
    cnmfe_options opt;
    cudaStream_t st = nullptr, st2 = nullptr;   // st2: second-moment kernel, overlapped with the host planning of the BG update
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, pe0 = nullptr, pe1 = nullptr, ge0 = nullptr, ge1 = nullptr;
    float phase_ms[7] = {0, 0, 0, 0, 0, 0, 0};
    int last_gram_tensor = 0, last_gram_frames = 0;
    long long last_active_pixels = 0;   // pixels whose ring weights the last background update refitted
    unsigned long long C_version = 1;   // bumped whenever the device copy of obj.C changes (guards Patch::Mcache)
    int last_nmf_iters = 0;             // ALS iterations of the last nmf background fit (summed over the owned patches)
    int use_c_hat = 1;            // update_temporal_parallel(obj, use_parallel, use_c_hat)
    int trace_major = 0;          // 1: K x T arrays cross the ABI trace-contiguous ([K][T]) instead of MATLAB column-major
    void* ssub_state = nullptr;   // SsubCtx (ctx_ssub.inc): coarse-grid ring model for options.bg_ssub > 1
    int num_neighbors = 0;
    bool first_bg = true;          // flag_first of update_background_parallel.m:142-146 (W{1} still uniform)
    // neurons
    HostCsc A, Aprev, IND;
    int K = 0, Kprev = 0;
    double* C = nullptr; double* Cprev = nullptr;        // [K][T]
    double* Craw = nullptr; double* S = nullptr; double* num = nullptr; double* den = nullptr;
    double* kpars = nullptr; double* nsn = nullptr; double* outs = nullptr;
    size_t Kcap = 0, Kprev_cap = 0;
    std::vector<double> sn;        // d1*d2
    std::vector<double> A_on_IND;  // result of the spatial update on the IND pattern
    bool have_spatial = false;
    Scratch scr;
    TraceArena arena;
    unsigned int* d_ticket = nullptr; int* d_done = nullptr; int* d_order = nullptr; int* d_err = nullptr;
    int* d_pmax = nullptr;
};

namespace {

// host-side section timer: CNMFE_HOST_PROFILE=1 prints the wall time between ticks (diagnostics only)
struct HostTick {
    bool on; std::chrono::steady_clock::time_point t;
    HostTick() : on(getenv("CNMFE_HOST_PROFILE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void operator()(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[cnmfe host] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

void phase_begin(cnmfe_ctx* c) { cudaEventRecord(c->pe0, c->st); }
void phase_end(cnmfe_ctx* c, int which) {
    cudaEventRecord(c->pe1, c->st);
    cudaEventSynchronize(c->pe1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->pe0, c->pe1);
    c->phase_ms[which] += ms;
}

int ensure_K(cnmfe_ctx* c, int K) {
    if ((size_t)K <= c->Kcap) return 0;
    size_t n = (size_t)K * c->T;
    for (double** p : {&c->C, &c->Craw, &c->S, &c->num}) { if (*p) cudaFree(*p); *p = nullptr; }
    for (double** p : {&c->den, &c->kpars, &c->nsn, &c->outs}) { if (*p) cudaFree(*p); *p = nullptr; }
    if (c->d_done) cudaFree(c->d_done);
    if (c->d_order) cudaFree(c->d_order);
    ++c->C_version;
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->C, n * 8));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->Craw, n * 8));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->S, n * 8));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->num, n * 8));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->den, (size_t)K * 8));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->kpars, (size_t)K * 16));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->nsn, (size_t)K * 8));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->outs, (size_t)K * 48));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->d_done, (size_t)K * 4));
    CNMFE_CUDA_OK(cudaMalloc((void**)&c->d_order, (size_t)(K + 1) * 4));
    CNMFE_CUDA_OK(cudaMemset(c->Craw, 0, n * 8));
    CNMFE_CUDA_OK(cudaMemset(c->S, 0, n * 8));
    CNMFE_CUDA_OK(cudaMemset(c->kpars, 0, (size_t)K * 16));
    CNMFE_CUDA_OK(cudaMemset(c->nsn, 0, (size_t)K * 8));
    c->Kcap = K;
    return 0;
}

// MATLAB K x T column-major host -> device [K][T]
int upload_KT(cnmfe_ctx* c, const double* hostC, int K, double* dst) {
    if (K == 0) return 0;
    size_t n = (size_t)K * c->T;
    if (c->trace_major) {   // host array already trace-contiguous ([K][T], NumPy C order): no transposition
        CNMFE_CUDA_OK(cudaMemcpyAsync(dst, hostC, n * 8, cudaMemcpyHostToDevice, c->st));
        CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
        return 0;
    }
    if (c->scr.reserve(std::max(c->scr.cap, n * 8 + 4096))) return -1;
    c->scr.reset();
    double* tmp = c->scr.take<double>(n);
    CNMFE_CUDA_OK(cudaMemcpyAsync(tmp, hostC, n * 8, cudaMemcpyHostToDevice, c->st));
    dim3 g((c->T + 255) / 256, K);
    LAUNCH(colmajor_to_kt_kernel, g, 256, 0, c->st, tmp, K, c->T, dst);
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    return 0;
}
int download_KT(cnmfe_ctx* c, const double* src, int K, double* hostC) {
    if (K == 0 || !hostC) return 0;
    size_t n = (size_t)K * c->T;
    if (c->trace_major) {
        CNMFE_CUDA_OK(cudaMemcpyAsync(hostC, src, n * 8, cudaMemcpyDeviceToHost, c->st));
        CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
        return 0;
    }
    if (c->scr.reserve(std::max(c->scr.cap, n * 8 + 4096))) return -1;
    c->scr.reset();
    double* tmp = c->scr.take<double>(n);
    dim3 g((c->T + 255) / 256, K);
    LAUNCH(kt_to_colmajor_kernel, g, 256, 0, c->st, src, K, c->T, tmp);
    CNMFE_CUDA_OK(cudaMemcpyAsync(hostC, tmp, n * 8, cudaMemcpyDeviceToHost, c->st));
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    return 0;
}

// get_nhood (endoscope/get_nhood.m:7-24)
void get_nhood(int radius, int k, std::vector<int>& rs, std::vector<int>& cs) {
    rs.clear(); cs.clear();
    for (int c = -radius; c <= radius; ++c)
        for (int r = -radius; r <= radius; ++r) {
            double R = std::sqrt((double)(c * c + r * r));
            if (R >= radius && R < radius + 1) { rs.push_back(r); cs.push_back(c); }
        }
    int n = (int)rs.size();
    if (k <= 0 || k > n) return;
    std::vector<int> ids(n);
    for (int i = 0; i < n; ++i) ids[i] = i;
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) {
        return std::atan2((double)rs[a], (double)cs[a]) < std::atan2((double)rs[b], (double)cs[b]);
    });
    std::vector<int> r2, c2;
    for (int j = 0; j < k; ++j) {
        double v = (k == 1) ? (double)n : 1.0 + (double)j * (double)(n - 1) / (double)(k - 1);
        int idx = (int)std::floor(v + 0.5) - 1;
        idx = std::min(std::max(idx, 0), n - 1);
        r2.push_back(rs[ids[idx]]); c2.push_back(cs[ids[idx]]);
    }
    rs = r2; cs = c2;
}

enum SelMode { SEL_SUM_BLOCK, SEL_SUM_HALO, SEL_ANY_PATCH };
enum RowSpace { ROWS_BLOCK, ROWS_BLOCK_PATCHONLY, ROWS_PATCH };

// Build the local view of M (values from `vals`, pattern + selection from M).  rows: which pixel index space the
// row lists use.  vals == nullptr -> M's own values.
void build_local(const cnmfe_ctx* c, const Patch& P, const HostCsc& M, SelMode sel, RowSpace rows,
                 const HostCsc* vals, LocalSparse* L) {
    // Host planning runs between kernels on the critical path of every update call, so it is written flat: no per-neuron
    // containers, no 64-bit division per entry (the rows of a CSC column are sorted, so the FOV column of an entry only
    // ever moves forward), values of `vals` found by a sorted merge instead of a binary search per entry.
    const int64_t d1 = c->d1;
    L->ids.clear();
    const int nrows = (rows == ROWS_PATCH) ? P.dp : P.db;
    const size_t nnz = M.ir.size();
    static thread_local std::vector<int> lidx;          // local row index of each entry of the current column, -1 = dropped
    L->cptr.assign(1, 0); L->crow.clear(); L->cval.clear(); L->bbox.clear();
    L->crow.reserve(nnz); L->cval.reserve(nnz);
    static thread_local std::vector<int64_t> csrc;      // csc entry index of each kept column entry
    csrc.clear(); csrc.reserve(nnz);
    const int nrw = (rows == ROWS_PATCH) ? P.nr : P.nrb;
    for (int k = 0; k < M.K; ++k) {
        const int64_t e0 = M.jc[k], e1 = M.jc[k + 1];
        if (e1 == e0) continue;                          // no entries: never selected
        // sorted rows: the FOV columns of this neuron span [first / d1, last / d1]; if that misses the block, no entry can be
        // in the block or the patch -- skip without the per-entry pass (keeps the planning O(local nnz + K) when the FOV is
        // tiled by many patches / sharded over GPUs)
        if (M.ir[e0] <= M.ir[e1 - 1] && (M.ir[e1 - 1] / d1 < P.block.c0 || M.ir[e0] / d1 > P.block.c1)) continue;
        if ((size_t)(e1 - e0) > lidx.size()) lidx.resize((size_t)(e1 - e0));
        double sum = 0.0;
        bool any = false;
        int64_t cc = 0, base = 0;                       // FOV column of the current entry and its first linear index
        for (int64_t e = e0; e < e1; ++e) {
            const int64_t row = M.ir[e];
            if (row < base || row >= base + d1) { cc = row / d1; base = cc * d1; }
            const int r = (int)(row - base), ci = (int)cc;
            const bool inb = r >= P.block.r0 && r <= P.block.r1 && ci >= P.block.c0 && ci <= P.block.c1;
            const bool inp = r >= P.patch.r0 && r <= P.patch.r1 && ci >= P.patch.c0 && ci <= P.patch.c1;
            if (sel == SEL_SUM_BLOCK && inb) sum += M.pr[e];
            if (sel == SEL_SUM_HALO && inb && !inp) sum += M.pr[e];
            if (sel == SEL_ANY_PATCH && inp) any = true;
            int idx = -1;
            if (rows == ROWS_PATCH) { if (inp) idx = (ci - P.patch.c0) * P.nr + (r - P.patch.r0); }
            else if (rows == ROWS_BLOCK_PATCHONLY) { if (inp) idx = (ci - P.block.c0) * P.nrb + (r - P.block.r0); }
            else { if (inb) idx = (ci - P.block.c0) * P.nrb + (r - P.block.r0); }
            lidx[(size_t)(e - e0)] = idx;
        }
        const bool take = (sel == SEL_ANY_PATCH) ? any : (sum > 0.0);
        if (!take) continue;
        // entries are already sorted by local index because ir is sorted and the index map is monotone in (c, r)
        int r0 = 1 << 30, r1 = -1, c0 = 1 << 30, c1 = -1;
        int64_t vp = vals ? vals->jc[k] : 0;
        const int64_t vend = vals ? vals->jc[k + 1] : 0;
        for (int64_t e = e0; e < e1; ++e) {
            const int idx = lidx[(size_t)(e - e0)];
            if (idx < 0) continue;
            double v;
            if (vals) {
                const int64_t row = M.ir[e];
                while (vp < vend && vals->ir[vp] < row) ++vp;
                v = (vp < vend && vals->ir[vp] == row) ? vals->pr[vp] : 0.0;
            } else v = M.pr[e];
            L->crow.push_back(idx); L->cval.push_back(v); csrc.push_back(e);
            int r = idx % nrw, ci = idx / nrw;
            if (rows == ROWS_PATCH) { r += P.patch.r0 - P.block.r0; ci += P.patch.c0 - P.block.c0; }
            r0 = std::min(r0, r); r1 = std::max(r1, r); c0 = std::min(c0, ci); c1 = std::max(c1, ci);
        }
        L->ids.push_back(k);
        L->cptr.push_back((int)L->crow.size());
        if (r1 < 0) { r0 = 0; r1 = -1; c0 = 0; c1 = -1; }
        L->bbox.push_back(r0); L->bbox.push_back(r1); L->bbox.push_back(c0); L->bbox.push_back(c1);
    }
    const int K = L->K();
    // rows by pixel (ascending local neuron id inside each row)
    L->ptr.assign(nrows + 1, 0);
    const size_t nkept = L->crow.size();
    for (size_t x = 0; x < nkept; ++x) L->ptr[L->crow[x] + 1]++;
    for (int i = 0; i < nrows; ++i) L->ptr[i + 1] += L->ptr[i];
    L->col.assign(nkept, 0); L->val.assign(nkept, 0.0); L->entry_src.assign(nkept, 0);
    static thread_local std::vector<int> fill;
    fill.assign(L->ptr.begin(), L->ptr.end() - 1);
    for (int k = 0; k < K; ++k)
        for (int x = L->cptr[k]; x < L->cptr[k + 1]; ++x) {
            const int pos = fill[L->crow[x]]++;
            L->col[pos] = k; L->val[pos] = L->cval[x]; L->entry_src[pos] = csrc[x];
        }
}

std::vector<int> expand_bbox(const LocalSparse& L, int margin, int nrb, int ncb) {
    std::vector<int> b(L.bbox);
    for (int k = 0; k < L.K(); ++k) {
        if (b[4 * k + 1] < b[4 * k]) continue;
        b[4 * k] = std::max(0, b[4 * k] - margin); b[4 * k + 1] = std::min(nrb - 1, b[4 * k + 1] + margin);
        b[4 * k + 2] = std::max(0, b[4 * k + 2] - margin); b[4 * k + 3] = std::min(ncb - 1, b[4 * k + 3] + margin);
    }
    return b;
}

template <class T> T* to_dev(cnmfe_ctx* c, const std::vector<T>& v, size_t min_n = 1) {
    size_t n = std::max(v.size(), min_n);
    T* p = c->scr.take<T>(n);
    if (!p) return nullptr;
    if (!v.empty()) cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->st);
    return p;
}

size_t pad256(size_t b) { return (b + 255) / 256 * 256 + 256; }

#define TAKE_OR_FAIL(var, expr)                                   \
    auto var = (expr);                                            \
    if (!var) { set_error("internal: scratch arena exhausted at %s:%d", __FILE__, __LINE__); return -1; }

}  // namespace

static int ssub_configure(cnmfe_ctx* c, int ssub);
static int ssub_set_ring(cnmfe_ctx* c, int ip, const double* W, const double* b0);
static int ssub_get_ring(cnmfe_ctx* c, int ip, double* W, double* b0);
static void ssub_destroy(cnmfe_ctx* c);

// ===================================================================================================== lifetime
extern "C" int cnmfe_create(cnmfe_ctx** out, int d1, int d2, int T, int npatch, const int32_t* patch_pos,
                            const int32_t* block_pos, const uint8_t* owned, int ring_radius, int num_neighbors,
                            int device) {
    if (!out || d1 <= 0 || d2 <= 0 || T <= 0 || npatch <= 0 || !patch_pos || !block_pos || ring_radius <= 0) {
        set_error("cnmfe_create: bad arguments");
        return -1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("cnmfe_create: no CUDA device (libcnmfe_b200 has no CPU fallback)");
        return -1;
    }
    if (device < 0 || device >= ndev) { set_error("cnmfe_create: device %d out of range", device); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(device));
    cnmfe_ctx* c = new cnmfe_ctx();
    c->device = device; c->d1 = d1; c->d2 = d2; c->T = T; c->Tpad = (T + 127) / 128 * 128; c->npatch = npatch;
    c->ring_radius = ring_radius;
    c->num_neighbors = num_neighbors;
    cnmfe_options_defaults(&c->opt);
    get_nhood(ring_radius, num_neighbors, c->off_r, c->off_c);
    c->nnb = (int)c->off_r.size();
    c->rr = ring_radius;
    if (c->nnb + 2 > 128) {
        set_error("cnmfe_create: ring radius %d has %d neighbours; the register-resident solver handles <= 126 (pass num_neighbors)", ring_radius, c->nnb);
        delete c;
        return -1;
    }
    c->sn.assign((size_t)d1 * d2, 1.0);
    cudaStreamCreate(&c->st); cudaStreamCreate(&c->st2); cudaEventCreate(&c->ge0); cudaEventCreate(&c->ge1);
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1); cudaEventCreate(&c->pe0); cudaEventCreate(&c->pe1);
    cudaMalloc((void**)&c->d_off_r, c->nnb * 4); cudaMalloc((void**)&c->d_off_c, c->nnb * 4);
    cudaMemcpy(c->d_off_r, c->off_r.data(), c->nnb * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_off_c, c->off_c.data(), c->nnb * 4, cudaMemcpyHostToDevice);
    cudaMalloc((void**)&c->d_ticket, 64); cudaMalloc((void**)&c->d_err, 64); cudaMalloc((void**)&c->d_pmax, 64);
    // displacement groups for the SIMT moment kernel
    std::vector<int> groups;
    for (int dr = 0; dr <= 2 * c->rr; dr += 4) { groups.push_back(0); groups.push_back(dr); }
    for (int dc = 1; dc <= 2 * c->rr; ++dc)
        for (int dr = -2 * c->rr; dr <= 2 * c->rr; dr += 4) { groups.push_back(dc); groups.push_back(dr); }
    c->ngroups = (int)groups.size() / 2;
    cudaMalloc((void**)&c->d_groups, groups.size() * 4);
    cudaMemcpy(c->d_groups, groups.data(), groups.size() * 4, cudaMemcpyHostToDevice);
    c->patches.resize(npatch);
    for (int i = 0; i < npatch; ++i) {
        Patch& P = c->patches[i];
        const int32_t* pp = patch_pos + 4 * i; const int32_t* bp = block_pos + 4 * i;
        P.patch = {pp[0] - 1, pp[1] - 1, pp[2] - 1, pp[3] - 1};
        P.block = {bp[0] - 1, bp[1] - 1, bp[2] - 1, bp[3] - 1};
        P.nr = P.patch.r1 - P.patch.r0 + 1; P.nc = P.patch.c1 - P.patch.c0 + 1;
        P.nrb = P.block.r1 - P.block.r0 + 1; P.ncb = P.block.c1 - P.block.c0 + 1;
        P.dp = P.nr * P.nc; P.db = P.nrb * P.ncb;
        if (P.nr <= 0 || P.nc <= 0 || P.patch.r0 < P.block.r0 || P.patch.r1 > P.block.r1 || P.patch.c0 < P.block.c0 ||
            P.patch.c1 > P.block.c1 || P.block.r0 < 0 || P.block.r1 >= d1 || P.block.c0 < 0 || P.block.c1 >= d2) {
            set_error("cnmfe_create: patch %d geometry invalid", i);
            cnmfe_destroy(c);
            return -1;
        }
        P.owned = owned ? owned[i] != 0 : true;
        RingGeom& g = P.geom;
        g.nnb = c->nnb; g.rr = c->rr; g.nrb = P.nrb; g.ncb = P.ncb; g.nr = P.nr; g.nc = P.nc;
        g.pr_off = P.patch.r0 - P.block.r0; g.pc_off = P.patch.c0 - P.block.c0;
        g.br0 = P.block.r0; g.bc0 = P.block.c0; g.d1 = d1; g.d2 = d2;
        if (P.owned) {
            if (cudaMalloc((void**)&P.W, (size_t)c->nnb * P.dp * 8) != cudaSuccess ||
                cudaMalloc((void**)&P.b0, (size_t)P.dp * 8) != cudaSuccess) {
                set_error("cnmfe_create: out of device memory");
                cnmfe_destroy(c);
                return -1;
            }
            LAUNCH(ring_uniform_kernel, (P.dp + 255) / 256, 256, 0, c->st, g, c->d_off_r, c->d_off_c, P.W);
            cudaMemsetAsync(P.b0, 0, (size_t)P.dp * 8, c->st);
        }
    }
    cudaStreamSynchronize(c->st);
    *out = c;
    return 0;
}

extern "C" void cnmfe_destroy(cnmfe_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    ssub_destroy(c);
    for (Patch& P : c->patches) {
        for (void* p : {(void*)P.Yt, (void*)P.hi, (void*)P.lo, (void*)P.hi_k, (void*)P.lo_k, (void*)P.Ysum, (void*)P.Ymean, (void*)P.W, (void*)P.b0, (void*)P.bsvd, (void*)P.fsvd, (void*)P.Mcache})
            if (p) cudaFree(p);
    }
    for (void* p : {(void*)c->C, (void*)c->Cprev, (void*)c->Craw, (void*)c->S, (void*)c->num, (void*)c->den,
                    (void*)c->kpars, (void*)c->nsn, (void*)c->outs, (void*)c->d_off_r, (void*)c->d_off_c,
                    (void*)c->d_groups, (void*)c->d_ticket, (void*)c->d_done, (void*)c->d_order, (void*)c->d_err,
                    (void*)c->d_pmax, (void*)c->scr.base})
        if (p) cudaFree(p);
    trace_arena_free(&c->arena);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->pe0) cudaEventDestroy(c->pe0);
    if (c->pe1) cudaEventDestroy(c->pe1);
    if (c->ge0) cudaEventDestroy(c->ge0);
    if (c->ge1) cudaEventDestroy(c->ge1);
    if (c->st2) cudaStreamDestroy(c->st2);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

extern "C" int cnmfe_set_options(cnmfe_ctx* c, const cnmfe_options* o) {
    if (!c || !o) { set_error("cnmfe_set_options: null"); return -1; }
    if (o->spatial_algorithm < 0 || o->spatial_algorithm > 3) { set_error("spatial_algorithm out of range"); return -1; }
    if (o->background_model < 0 || o->background_model > 2) { set_error("background_model must be 0 (ring), 1 (svd) or 2 (nmf)"); return -1; }
    if (o->background_model >= 1 && (o->nb < 1 || o->nb > SVD_MAXNB)) { set_error("svd background: nb must be in 1..%d", SVD_MAXNB); return -1; }
    if (o->bg_ssub < 1 || o->bg_ssub > 8) { set_error("bg_ssub must be in 1..8"); return -1; }
    if (!std::isnan(o->thresh_outlier) && (o->background_model != 0 || o->bg_ssub != 1)) {
        set_error("thresh_outlier is built for the ring model with bg_ssub = 1 (leave it NaN, the CNMFSetParms default, otherwise)"); return -1;
    }
    if (o->background_model == 0) { CNMFE_CUDA_OK(cudaSetDevice(c->device)); if (ssub_configure(c, o->bg_ssub)) return -1; }
    c->opt = *o;
    return 0;
}

static int upload_common(cnmfe_ctx* c, int ip, const void* Y, int dtype, bool on_device) {
    if (!c || ip < 0 || ip >= c->npatch || !Y) { set_error("cnmfe_upload_block: bad arguments"); return -1; }
    if (dtype < 0 || dtype > 3) {
        set_error("cnmfe_upload_block: dtype %d unsupported (0 = uint8, 1 = uint16, 2 = single, 3 = double holding integer counts)", dtype);
        return -1;
    }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    Patch& P = c->patches[ip];
    if (!P.owned) { set_error("cnmfe_upload_block: patch %d is not owned by this context", ip); return -1; }
    size_t n = (size_t)P.db * c->Tpad;
    if (!P.Yt) {
        CNMFE_CUDA_OK(cudaMalloc((void**)&P.Yt, n * 2));
        CNMFE_CUDA_OK(cudaMalloc((void**)&P.hi, n));
        CNMFE_CUDA_OK(cudaMalloc((void**)&P.lo, n));
        CNMFE_CUDA_OK(cudaMalloc((void**)&P.Ysum, (size_t)P.db * 8));
        CNMFE_CUDA_OK(cudaMalloc((void**)&P.Ymean, (size_t)P.db * 8));
    }
    P.planes_kf = 0;   // sub-sampled planes belong to the previous video
    CNMFE_CUDA_OK(cudaMemsetAsync(P.Yt, 0, n * 2, c->st));
    CNMFE_CUDA_OK(cudaMemsetAsync(P.hi, 0, n, c->st));
    CNMFE_CUDA_OK(cudaMemsetAsync(P.lo, 0, n, c->st));
    const size_t esz = dtype == 0 ? 1 : (dtype == 1 ? 2 : (dtype == 2 ? 4 : 8));
    const int chunk = 256;
    void* stage = nullptr; uint16_t* conv = nullptr;
    if (!on_device) CNMFE_CUDA_OK(cudaMalloc(&stage, (size_t)chunk * P.db * esz));
    if (dtype >= 2) {      // integer counts stored as floating point: exact conversion, anything else is refused below
        CNMFE_CUDA_OK(cudaMalloc((void**)&conv, (size_t)chunk * P.db * 2));
        CNMFE_CUDA_OK(cudaMemsetAsync(c->d_err, 0, 4, c->st));
    }
    for (int t0 = 0; t0 < c->T; t0 += chunk) {
        int nf = std::min(chunk, c->T - t0);
        const char* src = (const char*)Y + (size_t)t0 * P.db * esz;
        const void* dsrc = src;
        if (!on_device) {
            CNMFE_CUDA_OK(cudaMemcpyAsync(stage, src, (size_t)nf * P.db * esz, cudaMemcpyHostToDevice, c->st));
            dsrc = stage;
        }
        dim3 g((P.db + 31) / 32, (nf + 31) / 32), b(32, 8);
        if (dtype >= 2) {
            const size_t n1 = (size_t)nf * P.db;
            if (dtype == 2) LAUNCH(float_to_u16_kernel<float>, (unsigned)((n1 + 255) / 256), 256, 0, c->st, (const float*)dsrc, n1, conv, c->d_err);
            else LAUNCH(float_to_u16_kernel<double>, (unsigned)((n1 + 255) / 256), 256, 0, c->st, (const double*)dsrc, n1, conv, c->d_err);
            dsrc = conv;
        }
        if (dtype == 0) LAUNCH(transpose_chunk_kernel<uint8_t>, g, b, 0, c->st, (const uint8_t*)dsrc, P.db, nf, t0, c->Tpad, P.Yt, P.hi, P.lo);
        else LAUNCH(transpose_chunk_kernel<uint16_t>, g, b, 0, c->st, (const uint16_t*)dsrc, P.db, nf, t0, c->Tpad, P.Yt, P.hi, P.lo);
        if (!on_device) CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    }
    if (dtype >= 2) {
        int bad = 0;
        CNMFE_CUDA_OK(cudaMemcpyAsync(&bad, c->d_err, 4, cudaMemcpyDeviceToHost, c->st));
        CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
        cudaFree(conv);
        if (bad) {
            if (stage) cudaFree(stage);
            set_error("cnmfe_upload_block: the floating-point video holds values that are not integers in [0, 65535]; the exact path "
                      "works on integer counts (convert or rescale the movie to uint16 first)");
            return -1;
        }
    }
    LAUNCH(row_sum_kernel, (P.db * 32 + 255) / 256, 256, 0, c->st, P.Yt, P.db, c->T, c->Tpad, P.Ysum);
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    if (stage) cudaFree(stage);
    std::vector<double> ys(P.db);
    CNMFE_CUDA_OK(cudaMemcpy(ys.data(), P.Ysum, (size_t)P.db * 8, cudaMemcpyDeviceToHost));
    for (double& v : ys) v = v / (double)c->T;     // mean(Y,2)
    CNMFE_CUDA_OK(cudaMemcpy(P.Ymean, ys.data(), (size_t)P.db * 8, cudaMemcpyHostToDevice));
    P.uploaded = true;
    return 0;
}
extern "C" int cnmfe_upload_block(cnmfe_ctx* c, int ip, const void* Y, int dtype) { return upload_common(c, ip, Y, dtype, false); }
extern "C" int cnmfe_upload_block_dev(cnmfe_ctx* c, int ip, const void* Y, int dtype) { return upload_common(c, ip, Y, dtype, true); }

extern "C" int cnmfe_set_neurons(cnmfe_ctx* c, int K, const int64_t* jc, const int64_t* ir, const double* pr,
                                 const double* C) {
    // a NULL (jc,ir,pr) triple or a NULL C keeps what the context already holds for that part (K must then match)
    const bool keepA = !jc && !ir && !pr, keepC = !C;
    if (!c || K < 0 || (K > 0 && !keepA && (!jc || !ir || !pr))) { set_error("cnmfe_set_neurons: bad arguments"); return -1; }
    if (K > 0 && (keepA || keepC) && K != c->K) {
        set_error("cnmfe_set_neurons: NULL part with K=%d but the context holds K=%d", K, c->K); return -1;
    }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    static const int64_t zero = 0;
    if (K == 0) { c->A.set(0, &zero, nullptr, nullptr); c->K = 0; return 0; }
    if (!keepA) c->A.set(K, jc, ir, pr);
    if (K != c->K) { c->have_spatial = false; }
    c->K = K;
    if (ensure_K(c, K)) return -1;
    if (!keepC) ++c->C_version;
    return keepC ? 0 : upload_KT(c, C, K, c->C);
}

extern "C" int cnmfe_set_prev(cnmfe_ctx* c, int K, const int64_t* jc, const int64_t* ir, const double* pr,
                              const double* C) {
    const bool keepA = !jc && !ir && !pr, keepC = !C;
    if (!c || K < 0 || (K > 0 && !keepA && (!jc || !ir || !pr))) { set_error("cnmfe_set_prev: bad arguments"); return -1; }
    if (K > 0 && (keepA || keepC) && K != c->Kprev) {
        set_error("cnmfe_set_prev: NULL part with K=%d but the context holds K=%d", K, c->Kprev); return -1;
    }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    static const int64_t zero = 0;
    for (Patch& Pq : c->patches) Pq.lp_valid = false;
    if (K == 0) { c->Aprev.set(0, &zero, nullptr, nullptr); c->Kprev = 0; return 0; }
    if (!keepA) c->Aprev.set(K, jc, ir, pr);
    c->Kprev = K;
    if ((size_t)K > c->Kprev_cap) {
        if (c->Cprev) cudaFree(c->Cprev);
        CNMFE_CUDA_OK(cudaMalloc((void**)&c->Cprev, (size_t)K * c->T * 8));
        c->Kprev_cap = K;
    }
    return keepC ? 0 : upload_KT(c, C, K, c->Cprev);
}

extern "C" int cnmfe_set_search(cnmfe_ctx* c, int K, const int64_t* jc, const int64_t* ir) {
    if (!c || K < 0 || (K > 0 && (!jc || !ir))) { set_error("cnmfe_set_search: bad arguments"); return -1; }
    static const int64_t zero = 0;
    if (K == 0) c->IND.set(0, &zero, nullptr, nullptr); else c->IND.set(K, jc, ir, nullptr);
    return 0;
}

extern "C" int cnmfe_set_sn(cnmfe_ctx* c, const double* sn) {
    if (!c || !sn) { set_error("cnmfe_set_sn: null"); return -1; }
    c->sn.assign(sn, sn + (size_t)c->d1 * c->d2);
    return 0;
}

extern "C" int cnmfe_ring_offsets(cnmfe_ctx* c, int* nnb, int32_t* r_shift, int32_t* c_shift) {
    if (!c || !nnb) { set_error("cnmfe_ring_offsets: null"); return -1; }
    *nnb = c->nnb;
    if (r_shift) for (int i = 0; i < c->nnb; ++i) r_shift[i] = c->off_r[i];
    if (c_shift) for (int i = 0; i < c->nnb; ++i) c_shift[i] = c->off_c[i];
    return 0;
}

extern "C" int cnmfe_set_ring(cnmfe_ctx* c, int ip, const double* W, const double* b0) {
    if (!c || ip < 0 || ip >= c->npatch) { set_error("cnmfe_set_ring: bad patch"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    if (c->opt.background_model == 0 && c->opt.bg_ssub > 1) return ssub_set_ring(c, ip, W, b0);
    Patch& P = c->patches[ip];
    if (!W && b0) {   // only the offsets change: the weights (and the first-run state that goes with them) stay
        if (P.owned) CNMFE_CUDA_OK(cudaMemcpy(P.b0, b0, (size_t)P.dp * 8, cudaMemcpyHostToDevice));
        return 0;
    }
    bool uniform = true;
    if (W) {
        // first-run test of fit_ring_model.m:25 / update_background_parallel.m:143 on row 1 (patch pixel 0)
        const RingGeom& g = P.geom;
        double v0 = 0.0; bool have = false;
        for (int i = 0; i < c->nnb; ++i) {
            int fr = g.pr_off + c->off_r[i] + g.br0, fc = g.pc_off + c->off_c[i] + g.bc0;
            if (fr < 0 || fr >= g.d1 || fc < 0 || fc >= g.d2) continue;
            double v = W[i];
            if (!have) { v0 = v; have = true; } else if (v != v0) uniform = false;
        }
    }
    P.w_uniform = uniform;
    if (ip == 0) c->first_bg = uniform;
    if (!P.owned) return 0;
    if (W) CNMFE_CUDA_OK(cudaMemcpy(P.W, W, (size_t)c->nnb * P.dp * 8, cudaMemcpyHostToDevice));
    else {
        LAUNCH(ring_uniform_kernel, (P.dp + 255) / 256, 256, 0, c->st, P.geom, c->d_off_r, c->d_off_c, P.W);
        CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    }
    if (b0) CNMFE_CUDA_OK(cudaMemcpy(P.b0, b0, (size_t)P.dp * 8, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int cnmfe_get_ring(cnmfe_ctx* c, int ip, double* W, double* b0) {
    if (!c || ip < 0 || ip >= c->npatch) { set_error("cnmfe_get_ring: bad patch"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    if (c->opt.background_model == 0 && c->opt.bg_ssub > 1) return ssub_get_ring(c, ip, W, b0);
    Patch& P = c->patches[ip];
    if (!P.owned) { set_error("cnmfe_get_ring: patch %d not owned", ip); return -1; }
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    if (W) CNMFE_CUDA_OK(cudaMemcpy(W, P.W, (size_t)c->nnb * P.dp * 8, cudaMemcpyDeviceToHost));
    if (b0) CNMFE_CUDA_OK(cudaMemcpy(b0, P.b0, (size_t)P.dp * 8, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int cnmfe_sync(cnmfe_ctx* c) {
    if (!c) return -1;
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    return 0;
}
extern "C" int cnmfe_timer_begin(cnmfe_ctx* c) {
    if (!c) return -1;
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    CNMFE_CUDA_OK(cudaEventRecord(c->ev0, c->st));
    return 0;
}
extern "C" int cnmfe_timer_end(cnmfe_ctx* c, float* ms) {
    if (!c || !ms) return -1;
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    CNMFE_CUDA_OK(cudaEventRecord(c->ev1, c->st));
    CNMFE_CUDA_OK(cudaEventSynchronize(c->ev1));
    CNMFE_CUDA_OK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}
extern "C" int cnmfe_last_phase_ms(cnmfe_ctx* c, float* ms7) {
    if (!c || !ms7) return -1;
    for (int i = 0; i < 7; ++i) ms7[i] = c->phase_ms[i];
    return 0;
}

// ===================================================================================================== background
static size_t bg_scratch_bytes(const cnmfe_ctx* c, const Patch& P, int Kb, size_t nnzA) {
    size_t ND = (size_t)ring_num_disp(c->rr);
    size_t b = 0;
    if (!std::isnan(c->opt.thresh_outlier))      // explicit path: Bf [db][T], a chunk of explicit rows, masks, zero vectors
        b += pad256((size_t)P.db * c->T * 8) + pad256((size_t)1024 * c->T * 8) + 4 * pad256((size_t)(P.db + 1) * 8) + 2 * pad256((size_t)c->T * 4) +
             pad256((size_t)P.dp * 8) + (1 << 20);
    b += pad256(ND * P.db * 8);                 // S2
    b += pad256((size_t)P.db * std::max(Kb, 1) * 8);   // Mc / N
    b += pad256((size_t)std::max(Kb, 1) * c->T * 8);   // Cc
    b += pad256((size_t)P.db * 8) * 3;          // sumA, S1, spare
    b += pad256((size_t)(P.db + 1) * 4) + pad256(nnzA * 12 + 64);
    b += pad256((size_t)Kb * Kb * 8 + 64) + pad256((size_t)Kb * 64 + 64) * 4;
    b += pad256((size_t)P.dp * 8);
    b += (1 << 20);
    return b;
}

static int update_background_svd(cnmfe_ctx* c);
static int update_background_nmf(cnmfe_ctx* c);
static int ssub_configure(cnmfe_ctx* c, int ssub);
static int ssub_ensure_patch(cnmfe_ctx* c, int ip, bool need_video);
static int update_background_ssub(cnmfe_ctx* c);
static int ssub_forward(cnmfe_ctx* c, int ip, const double* in, int K, double* out, double* t1, double* t2);
static int ssub_transpose(cnmfe_ctx* c, int ip, const double* in, int K, double* out, double* t1, double* t2);
static int ssub_set_ring(cnmfe_ctx* c, int ip, const double* W, const double* b0);
static int ssub_get_ring(cnmfe_ctx* c, int ip, double* W, double* b0);
static void ssub_destroy(cnmfe_ctx* c);
static int update_spatial_svd(cnmfe_ctx* c);
static int update_temporal_patches_svd(cnmfe_ctx* c);

extern "C" int cnmfe_update_background(cnmfe_ctx* c) {
    if (!c) { set_error("cnmfe_update_background: null ctx"); return -1; }
    for (Patch& Pq : c->patches) Pq.lp_valid = false;
    if (c->opt.background_model == 1) return update_background_svd(c);
    if (c->opt.background_model == 2) return update_background_nmf(c);
    if (c->opt.bg_ssub > 1) return update_background_ssub(c);
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    for (float& f : c->phase_ms) f = 0;
    const int T = c->T;
    const bool flag_first = c->first_bg;
    c->last_active_pixels = 0;
    for (int ip = 0; ip < c->npatch; ++ip) {
        Patch& P = c->patches[ip];
        if (!P.owned) continue;
        if (!P.uploaded) { set_error("update_background: block %d not uploaded", ip); return -1; }
        LocalSparse& L = P.lp_cache;     // = the local view of A_prev once this call returns (A_prev <- A below)
        HostTick tick;
        build_local(c, P, c->A, SEL_SUM_BLOCK, ROWS_BLOCK, nullptr, &L);
        P.lp_valid = true;
        tick("bg build_local");
        const int Kb = L.K();
        if (Kb == 0 && !flag_first) continue;   // update_background_parallel.m:188-199
        if (Kb > 4096) { set_error("update_background: %d neurons touch block %d; the solver's neuron bitmap handles <= 4096 per block", Kb, ip); return -1; }
        if (c->scr.reserve(bg_scratch_bytes(c, P, Kb, L.col.size()))) return -1;
        c->scr.reset();
        tick("bg reserve");
        const RingGeom& g = P.geom;
        // frames used for the weights (fit_ring_model.m:59-90)
        const bool first_run = P.w_uniform;
        CNMFE_CUDA_OK(cudaMemsetAsync(c->d_pmax, 0, 4, c->st));
        LAUNCH(ring_pmax_kernel, (P.dp + 255) / 256, 256, 0, c->st, g, c->d_off_r, c->d_off_c, P.W, c->d_pmax);
        int pmax = 0;
        CNMFE_CUDA_OK(cudaMemcpyAsync(&pmax, c->d_pmax, 4, cudaMemcpyDeviceToHost, c->st));
        CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
        int kf = 1;
        if (c->opt.bg_acceleration) {
            long long nmax = 100LL * pmax;
            long long nk = std::min<long long>(T, nmax);
            if (nk < 1) nk = 1;
            kf = (int)(T / nk);
            if (kf < 1) kf = 1;
        }
        // second moments: they depend only on the resident video and the frame stride, so the kernel (the longest of the
        // call) starts NOW on a side stream and the host planning below (uploads, active pixels, neuron grams) overlaps it
        const size_t ND = (size_t)ring_num_disp(c->rr);
        double* d_S2 = c->scr.take<double>(ND * P.db);
        if (!d_S2) { set_error("scratch exhausted"); return -1; }
        const bool outlier = !std::isnan(c->opt.thresh_outlier);    // explicit fp64 path (fit_ring_model.m:48-70), moments computed below
        if (outlier) {
            CNMFE_CUDA_OK(cudaEventRecord(c->ge0, c->st2));
            CNMFE_CUDA_OK(cudaEventRecord(c->ge1, c->st2));
            c->last_gram_tensor = 0;
        } else {
            CNMFE_CUDA_OK(cudaEventRecord(c->ge0, c->st2));
            int tc_rc = 1;
            if (c->opt.use_tensor_gram && kf == 1)
                tc_rc = ring_s2_tensor(P.hi, P.lo, P.nrb, P.ncb, T, c->Tpad, c->rr, d_S2, c->st2);
            else if (c->opt.use_tensor_gram) {
                // frame-subsampled fit (fit_ring_model.m:84-90): the tensor kernel runs on compacted byte planes of the kept
                // frames, rebuilt only when the stride changes (one strided pass over the resident video)
                const int Tk = (T - 1) / kf + 1, Tpadk = (Tk + 127) / 128 * 128;
                if (P.planes_kf != kf) {
                    if (P.hi_k) cudaFree(P.hi_k);
                    if (P.lo_k) cudaFree(P.lo_k);
                    P.hi_k = P.lo_k = nullptr; P.planes_kf = 0;
                    if (cudaMalloc((void**)&P.hi_k, (size_t)P.db * Tpadk) == cudaSuccess &&
                        cudaMalloc((void**)&P.lo_k, (size_t)P.db * Tpadk) == cudaSuccess) {
                        dim3 gs(P.db, (Tpadk + 255) / 256);
                        LAUNCH(subsample_planes_kernel, gs, 256, 0, c->st2, P.Yt, c->Tpad, kf, Tk, Tpadk, P.hi_k, P.lo_k);
                        P.planes_kf = kf; P.Tpad_k = Tpadk;
                    } else {
                        (void)cudaGetLastError();      // no room for the planes: the exact SIMT kernel takes over
                        if (P.hi_k) cudaFree(P.hi_k);
                        P.hi_k = nullptr;
                    }
                }
                if (P.planes_kf == kf)
                    tc_rc = ring_s2_tensor(P.hi_k, P.lo_k, P.nrb, P.ncb, Tk, P.Tpad_k, c->rr, d_S2, c->st2);
            }
            if (tc_rc < 0) return -1;
            c->last_gram_tensor = (tc_rc == 0);
            c->last_gram_frames = (T - 1) / kf + 1;
            if (tc_rc != 0) {
                long long nw = (long long)((P.nrb + 3) / 4) * P.ncb;
                dim3 gg((unsigned)((nw + 7) / 8), c->ngroups);
                LAUNCH(ring_s2_simt_kernel, gg, 256, 0, c->st2, P.Yt, P.nrb, P.ncb, T, c->Tpad, kf, c->rr, c->d_groups,
                       c->ngroups, d_S2, ND);
            }
            CNMFE_CUDA_OK(cudaEventRecord(c->ge1, c->st2));
        }
        tick("bg pmax + launch of the second moments");
        phase_begin(c);
        TAKE_OR_FAIL(d_ptr, to_dev(c, L.ptr));
        TAKE_OR_FAIL(d_col, to_dev(c, L.col));
        TAKE_OR_FAIL(d_val, to_dev(c, L.val));
        TAKE_OR_FAIL(d_ids, to_dev(c, L.ids));
        std::vector<int> bb = expand_bbox(L, 2 * c->rr, P.nrb, P.ncb);
        TAKE_OR_FAIL(d_bbox, to_dev(c, bb));
        std::vector<double> sumA(P.db, 0.0);
        for (int q = 0; q < P.db; ++q) for (int e = L.ptr[q]; e < L.ptr[q + 1]; ++e) sumA[q] += L.val[e];
        TAKE_OR_FAIL(d_sumA, to_dev(c, sumA));
        TAKE_OR_FAIL(d_Cc, c->scr.take<double>((size_t)std::max(Kb, 1) * T));
        TAKE_OR_FAIL(d_Cmean, c->scr.take<double>(std::max(Kb, 1)));
        TAKE_OR_FAIL(d_Csum, c->scr.take<double>(std::max(Kb, 1)));
        TAKE_OR_FAIL(d_Vsel, c->scr.take<double>((size_t)std::max(Kb, 1) * std::max(Kb, 1)));
        TAKE_OR_FAIL(d_active, c->scr.take<unsigned char>(P.dp));
        if (Kb > 0)
            LAUNCH(gather_center_rows_kernel, Kb, 256, 0, c->st, c->C, d_ids, Kb, T, d_Cc, d_Cmean);
        tick("bg uploads");
        const double nsel = (double)((T - 1) / kf + 1);
        double* d_S1 = P.Ysum;
        if (kf != 1) {
            d_S1 = c->scr.take<double>(P.db);
            if (!d_S1) { set_error("scratch exhausted"); return -1; }
            LAUNCH(row_sum_strided_kernel, (P.db * 32 + 255) / 256, 256, 0, c->st, P.Yt, P.db, T, c->Tpad, kf, d_S1);
        }
        LAUNCH(ring_active_b0_kernel, (P.dp + 255) / 256, 256, 0, c->st, g, c->d_off_r, c->d_off_c, P.W, d_sumA,
               P.Ymean, d_ptr, d_col, d_val, d_Cmean, first_run ? 1 : 0, d_active, P.b0);
        std::vector<unsigned char> act(P.dp);
        CNMFE_CUDA_OK(cudaMemcpyAsync(act.data(), d_active, P.dp, cudaMemcpyDeviceToHost, c->st));
        CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
        std::vector<int> alist;
        alist.reserve(P.dp);
        // 16 x 16 pixel tiles: concurrently solved pixels share their ring rows of the moment table in L2 (a device-side
        // warp-aggregated compaction was tried: its scrambled group order cost the solver 3 %)
        for (int tc0 = 0; tc0 < P.nc; tc0 += 16)
            for (int tr0 = 0; tr0 < P.nr; tr0 += 16)
                for (int cc = tc0; cc < std::min(P.nc, tc0 + 16); ++cc)
                    for (int rr_ = tr0; rr_ < std::min(P.nr, tr0 + 16); ++rr_) {
                        int p = cc * P.nr + rr_;
                        if (act[p]) alist.push_back(p);
                    }
        if (alist.empty()) { CNMFE_CUDA_OK(cudaStreamSynchronize(c->st2)); continue; }
        c->last_active_pixels += (long long)alist.size();
        TAKE_OR_FAIL(d_alist, to_dev(c, alist));
        phase_end(c, 6);
        tick("bg active list");
        CNMFE_CUDA_OK(cudaStreamWaitEvent(c->st, c->ge1, 0));   // what follows needs the SMs (and then the moments)
        // ---- options.thresh_outlier: explicit Bf, clamp against the previous fit, frame selection, fp64 moments
        const double* S2_use = d_S2; const double* S1_use = nullptr; const double* Ym_use = P.Ymean;
        const int* aptr_use = d_ptr; int K_use = Kb; double nsel_use = 0.0;
        if (outlier) {
            phase_begin(c);
            if (Kb > YSIG_MAXK) { set_error("update_background: %d neurons touch block %d (the thresh_outlier path handles <= %d)", Kb, ip, YSIG_MAXK); return -1; }
            const int CH = 1024;
            TAKE_OR_FAIL(d_Bf, c->scr.take<double>((size_t)P.db * T));
            TAKE_OR_FAIL(d_rowsY, c->scr.take<double>((size_t)CH * T));
            TAKE_OR_FAIL(d_rows, c->scr.take<int>(CH));
            TAKE_OR_FAIL(d_zero, c->scr.take<double>(P.db + 1));
            TAKE_OR_FAIL(d_zptr, c->scr.take<int>(P.db + 1));
            TAKE_OR_FAIL(d_S1x, c->scr.take<double>(P.db));
            TAKE_OR_FAIL(d_cnt, c->scr.take<int>(T));
            TAKE_OR_FAIL(d_mask, c->scr.take<unsigned char>(T));
            std::vector<double> snp(P.dp);
            for (int p = 0; p < P.dp; ++p) snp[p] = c->sn[(size_t)(p / P.nr + P.patch.c0) * c->d1 + (p % P.nr + P.patch.r0)];
            TAKE_OR_FAIL(d_snp, to_dev(c, snp));
            CNMFE_CUDA_OK(cudaMemsetAsync(d_zero, 0, (size_t)(P.db + 1) * 8, c->st));
            CNMFE_CUDA_OK(cudaMemsetAsync(d_zptr, 0, (size_t)(P.db + 1) * 4, c->st));
            CNMFE_CUDA_OK(cudaMemsetAsync(d_cnt, 0, (size_t)T * 4, c->st));
            CNMFE_CUDA_OK(cudaMemsetAsync(d_S2, 0, ND * P.db * 8, c->st));
            for (int q0 = 0; q0 < P.db; q0 += 32768) {
                dim3 gg((T + 255) / 256, std::min(32768, P.db - q0));
                LAUNCH(bf_rows_kernel, gg, 256, 0, c->st, P.Yt, P.Ymean, T, c->Tpad, d_ptr, d_col, d_val, d_Cc, q0, d_Bf);
            }
            std::vector<int> rows(CH);
            for (int b = 0; b < P.dp; b += CH) {
                const int nr = std::min(CH, P.dp - b);
                for (int i = 0; i < nr; ++i) rows[i] = b + i;
                CNMFE_CUDA_OK(cudaMemcpyAsync(d_rows, rows.data(), (size_t)nr * 4, cudaMemcpyHostToDevice, c->st));
                dim3 gg(nr, (T + YSIG_TCHUNK - 1) / YSIG_TCHUNK);
                // Y - W_old * Bf for the rows of this chunk (b0 = 0; "previous" neurons = the current A, C of fit_ring_model's Bf)
                LAUNCH(ysig_rows_kernel, gg, 256, 0, c->st, g, c->d_off_r, c->d_off_c, P.W, d_zero, P.Yt, P.Ymean, T, c->Tpad, d_ptr, d_col,
                       d_val, Kb, d_Cc, d_rows, d_rowsY);
                dim3 g2((T + 255) / 256, nr);
                LAUNCH(bf_clamp_kernel, g2, 256, 0, c->st, g, d_rowsY, d_rows, P.Yt, T, c->Tpad, d_snp, c->opt.thresh_outlier, d_Bf, d_cnt);
                CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));      // `rows` is reused by the next chunk
            }
            // frames (fit_ring_model.m:59-70): with nmax = 100 pmax < T keep the frames whose outlier count is <= quantile(counts, nmax/T)
            std::vector<int> cnt(T);
            CNMFE_CUDA_OK(cudaMemcpy(cnt.data(), d_cnt, (size_t)T * 4, cudaMemcpyDeviceToHost));
            std::vector<unsigned char> mask(T, 1);
            long long nmax = 100LL * pmax;
            int Tsel = T;
            if (nmax < T) {
                std::vector<double> srt(cnt.begin(), cnt.end());
                std::sort(srt.begin(), srt.end());
                const double pos = ((double)nmax / (double)T) * (double)T + 0.5;     // MATLAB quantile: sorted x(i) at (i - 0.5)/n
                double qv;
                if (pos < 1.0) qv = srt[0];
                else if (pos >= (double)T) qv = srt[T - 1];
                else { const int lo = (int)std::floor(pos); const double fr = pos - lo; qv = srt[lo - 1] + fr * (srt[lo] - srt[lo - 1]); }
                Tsel = 0;
                for (int t = 0; t < T; ++t) { mask[t] = ((double)cnt[t] <= qv) ? 1 : 0; Tsel += mask[t]; }
                nmax = Tsel;
            }
            if (c->opt.bg_acceleration) {            // :84-90 on the selected frames
                const long long nk = std::max<long long>(1, std::min<long long>(Tsel, nmax));
                const int k2 = (int)(Tsel / nk);
                if (k2 > 1) { int j = 0; for (int t = 0; t < T; ++t) if (mask[t]) { if (j % k2) mask[t] = 0; ++j; } }
            }
            int nsel_i = 0;
            for (int t = 0; t < T; ++t) nsel_i += mask[t];
            CNMFE_CUDA_OK(cudaMemcpyAsync(d_mask, mask.data(), (size_t)T, cudaMemcpyHostToDevice, c->st));
            {
                long long nw = (long long)((P.nrb + 3) / 4) * P.ncb;
                dim3 gg((unsigned)((nw + 7) / 8), c->ngroups);
                LAUNCH(ring_s2_f64_kernel, gg, 256, 0, c->st, d_Bf, P.nrb, P.ncb, T, d_mask, c->rr, c->d_groups, c->ngroups, d_S2, ND);
                LAUNCH(bf_row_sum_kernel, (unsigned)(((size_t)P.db * 32 + 255) / 256), 256, 0, c->st, d_Bf, P.db, T, d_mask, d_S1x);
            }
            CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
            c->last_gram_frames = nsel_i;
            S2_use = d_S2; S1_use = d_S1x; Ym_use = d_zero; aptr_use = d_zptr; K_use = 0; nsel_use = (double)nsel_i;
            phase_end(c, 0);
        }
        // projections needed by the neuron corrections
        phase_begin(c);
        double* d_N = c->scr.take<double>((size_t)P.db * std::max(Kb, 1));
        if (!d_N) { set_error("scratch exhausted"); return -1; }
        if (Kb > 0 && !outlier) {
            // (only the light kernels above share the SMs with the second-moment kernel; this Gram of the traces waits for it)
            dim3 gg(Kb, Kb);
            LAUNCH(small_gram_kernel, gg, 128, 0, c->st, d_Cc, Kb, d_Cc, Kb, T, kf, d_Vsel, d_Csum);
            CNMFE_CUDA_OK(cudaMemsetAsync(d_N, 0, (size_t)P.db * Kb * 8, c->st));
            launch_proj_mc(c->st, P.Yt, P.Ymean, P.nrb, P.ncb, T, c->Tpad,
                   kf, d_Cc, Kb, d_bbox, d_N);
            P.mc_valid = false;
            const size_t mc_n = (size_t)P.db * Kb;
            if (kf == 1 && mc_n * 8 <= ((size_t)6 << 30) && !getenv("CNMFE_NO_PROJ_CACHE")) {
                if (mc_n > P.Mcache_cap) {
                    if (P.Mcache) cudaFree(P.Mcache);
                    P.Mcache = nullptr; P.Mcache_cap = 0;
                    if (cudaMalloc((void**)&P.Mcache, (mc_n + mc_n / 16) * 8) == cudaSuccess) P.Mcache_cap = mc_n + mc_n / 16;
                    else (void)cudaGetLastError();
                }
                if (P.Mcache) {
                    CNMFE_CUDA_OK(cudaMemcpyAsync(P.Mcache, d_N, mc_n * 8, cudaMemcpyDeviceToDevice, c->st));
                    P.mc_valid = true; P.mc_Cver = c->C_version; P.mc_K = Kb; P.mc_ids = L.ids; P.mc_bbox = bb;
                }
            }
            LAUNCH(ring_make_N_kernel, P.db, 64, 0, c->st, d_N, Kb, d_ptr, d_col, d_val, d_Vsel, (size_t)P.db);
        }
        phase_end(c, 2);
        tick("bg projections");
        // second moments: done (see above); their time is read from the side stream's events
        {
            float gms = 0;
            CNMFE_CUDA_OK(cudaEventSynchronize(c->ge1));
            CNMFE_CUDA_OK(cudaEventElapsedTime(&gms, c->ge0, c->ge1));
            c->phase_ms[0] += gms;
        }
        // assemble + solve
        phase_begin(c);
        RingSolveArgs a;
        a.g = g; a.off_r = c->d_off_r; a.off_c = c->d_off_c; a.S2 = S2_use; a.S1 = outlier ? S1_use : d_S1; a.Ymean = Ym_use;
        a.nsel = outlier ? nsel_use : nsel; a.a_ptr = aptr_use; a.a_col = d_col; a.a_val = d_val; a.N = d_N; a.K = K_use; a.Csum = d_Csum;
        a.active = d_active; a.active_list = d_alist; a.n_active = (int)alist.size(); a.n_active_dev = nullptr; a.W = P.W; a.db = (size_t)P.db; a.ND = ND;
        a.prof = nullptr;
        static const bool ring_profile = getenv("CNMFE_RING_PROFILE") != nullptr;
        if (ring_profile) {
            CNMFE_CUDA_OK(cudaMalloc((void**)&a.prof, 64));
            CNMFE_CUDA_OK(cudaMemsetAsync(a.prof, 0, 64, c->st));
        }
        const int NMAX = c->nnb + 1;
        // solver: block LDL' on the fp64 tensor-core path (default) or the register-tile SIMT kernel (CNMFE_RING_SOLVER=simt: checker)
        const bool solver_simt = getenv("CNMFE_RING_SOLVER") && !strcmp(getenv("CNMFE_RING_SOLVER"), "simt");   // read per call: tests A/B it
        if (!solver_simt) {
            const size_t sm = ring_solve_mma_smem_bytes();
            if (ring_profile) {
                CNMFE_CUDA_OK(cudaFuncSetAttribute(ring_solve_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                int occ = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_solve_mma_kernel<true>, RM_THREADS, sm);
                fprintf(stderr, "[cnmfe ring profile] mma solver: dynamic smem %zu B, occupancy %d CTAs/SM\n", sm, occ);
                LAUNCH(ring_solve_mma_kernel<true>, (unsigned)((alist.size() + RM_PIX - 1) / RM_PIX), RM_THREADS, sm, c->st, a);
            } else {
                CNMFE_CUDA_OK(cudaFuncSetAttribute(ring_solve_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                LAUNCH(ring_solve_mma_kernel<false>, (unsigned)((alist.size() + RM_PIX - 1) / RM_PIX), RM_THREADS, sm, c->st, a);
            }
        } else {
        size_t smem = ring_solve_smem_bytes(NMAX);
        {
            // shared-memory carve-out: just enough for the two CTAs per SM the register file allows; the rest stays L1, which
            // the moment gathers of the assembly phase live on (measured at 54.5 KB/CTA: carve-out 50 % 39.3 ms, 60 % 41.4,
            // 75 % 43.3, 100 % 53 ms; below two CTAs' worth 61 ms).  CNMFE_RING_CARVEOUT=percent overrides (A/B knob).
            int pct = (int)((2 * (smem + 2048) * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (const char* e = getenv("CNMFE_RING_CARVEOUT")) pct = atoi(e);
            CNMFE_CUDA_OK(cudaFuncSetAttribute(ring_solve_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            CNMFE_CUDA_OK(cudaFuncSetAttribute(ring_solve_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        }
        if (ring_profile) {
            int occ = 0;
            cudaFuncSetAttribute(ring_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_solve_kernel<false>, RING_SOLVE_THREADS, smem);
            fprintf(stderr, "[cnmfe ring profile] dynamic smem %zu B, occupancy %d CTAs/SM\n", smem, occ);
            CNMFE_CUDA_OK(cudaFuncSetAttribute(ring_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LAUNCH(ring_solve_kernel<true>, (unsigned)alist.size(), RING_SOLVE_THREADS, smem, c->st, a);
        } else {
            CNMFE_CUDA_OK(cudaFuncSetAttribute(ring_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LAUNCH(ring_solve_kernel<false>, (unsigned)alist.size(), RING_SOLVE_THREADS, smem, c->st, a);
        }
        }
        CNMFE_CUDA_OK(cudaGetLastError());
        phase_end(c, 1);
        if (ring_profile) {
            unsigned long long h[8];
            CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
            CNMFE_CUDA_OK(cudaMemcpy(h, a.prof, 64, cudaMemcpyDeviceToHost));
            cudaFree(a.prof);
            const int nact = (int)alist.size();
            const double np_ = (double)std::max(nact, 1);
            fprintf(stderr, "[cnmfe ring profile] pixels=%d cycles/pixel: setup=%.0f assemble=%.0f neuron-scan=%.0f(+bitmap) compact=%.0f corrections=%.0f ridge=%.0f ldl=%.0f backsub+store=%.0f; neurons/pixel=%.2f\n",
                    nact, h[0] / np_, h[1] / np_, h[2] / np_, 0.0, h[3] / np_, h[4] / np_, h[5] / np_, h[6] / np_, h[7] / np_);
        }
        tick("bg ring solve");
        P.w_uniform = false;
    }
    c->first_bg = false;
    // obj.A_prev = obj.A; obj.C_prev = obj.C  (update_background_parallel.m:316-317)
    c->Aprev = c->A;
    c->Kprev = c->K;
    if (c->K > 0) {
        if ((size_t)c->K > c->Kprev_cap) {
            if (c->Cprev) cudaFree(c->Cprev);
            CNMFE_CUDA_OK(cudaMalloc((void**)&c->Cprev, (size_t)c->K * T * 8));
            c->Kprev_cap = c->K;
        }
        CNMFE_CUDA_OK(cudaMemcpyAsync(c->Cprev, c->C, (size_t)c->K * T * 8, cudaMemcpyDeviceToDevice, c->st));
    }
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    CNMFE_CUDA_OK(cudaGetLastError());
    return 0;
}

// ===================================================================================================== spatial
extern "C" int cnmfe_update_spatial(cnmfe_ctx* c) { return cnmfe_update_spatial_ex(c, 0); }

extern "C" int cnmfe_get_sn_map(cnmfe_ctx* c, double* sn) {
    if (!c || !sn) { set_error("cnmfe_get_sn_map: null"); return -1; }
    std::copy(c->sn.begin(), c->sn.end(), sn);
    return 0;
}

extern "C" int cnmfe_update_spatial_ex(cnmfe_ctx* c, int update_sn) {
    if (!c) { set_error("cnmfe_update_spatial: null ctx"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    for (float& f : c->phase_ms) f = 0;
    if (c->IND.K != c->K) { set_error("update_spatial: search mask has %d columns, A has %d", c->IND.K, c->K); return -1; }
    if (c->opt.background_model >= 1) {
        if (update_sn || c->opt.spatial_algorithm == 3) { set_error("update_spatial: update_sn / lars are built for the ring model only"); return -1; }
        return update_spatial_svd(c);
    }
    const bool lars = (c->opt.spatial_algorithm == 3);
    const bool ssub = (c->opt.bg_ssub > 1);
    if (ssub && (lars || update_sn)) { set_error("update_spatial: update_sn / lars are built for bg_ssub = 1 only"); return -1; }
    const int T = c->T;
    c->A_on_IND.assign(c->IND.ir.size(), 0.0);
    for (int ip = 0; ip < c->npatch; ++ip) {
        Patch& P = c->patches[ip];
        if (!P.owned) continue;
        if (!P.uploaded) { set_error("update_spatial: block %d not uploaded", ip); return -1; }
        LocalSparse LS, LPown;
        HostTick tick;
        build_local(c, P, c->IND, SEL_ANY_PATCH, ROWS_PATCH, &c->A, &LS);
        tick("sp build_local IND");
        const int Ks = LS.K();
        if (Ks == 0 && !update_sn) continue;   // update_spatial_parallel.m:121-124
        const bool lp_cached = P.lp_valid && !c->opt.replicate_spatial_aprev_quirk;
        if (!lp_cached)
            build_local(c, P, c->Aprev, c->opt.replicate_spatial_aprev_quirk ? SEL_SUM_HALO : SEL_SUM_BLOCK, ROWS_BLOCK,
                        nullptr, &LPown);
        const LocalSparse& LP = lp_cached ? P.lp_cache : LPown;
        tick("sp build_local Aprev");
        const int Kp = LP.K();
        const size_t nent = LS.col.size();
        size_t need = pad256((size_t)P.db * Ks * 8) + pad256((size_t)Ks * T * 8) + pad256((size_t)std::max(Kp, 1) * T * 8) +
                      pad256((size_t)Ks * Ks * 8) + pad256((size_t)std::max(Kp, 1) * Ks * 8) + pad256(nent * 32) +
                      pad256((size_t)(P.dp + 1) * 4) + pad256((size_t)(P.db + 1) * 4) + pad256(LP.col.size() * 12 + 64) +
                      3 * pad256((size_t)P.dp * 8) + pad256((size_t)(Ks + Kp) * 64 + 64) + (16 << 20);
        const int CH = 2048;   // rows of explicit Ysig per chunk (update_sn / lars only)
        if (update_sn || lars) need += pad256((size_t)CH * T * 8) + 3 * pad256((size_t)CH * 8) + pad256((size_t)P.dp * 8);
        if (ssub) need += 3 * pad256((size_t)P.db * Ks * 8);
        if (c->scr.reserve(need)) return -1;
        c->scr.reset();
        const RingGeom& g = P.geom;
        if ((update_sn || lars) && Kp > YSIG_MAXK) { set_error("update_spatial: %d previous neurons touch block %d (explicit rows handle <= %d)", Kp, ip, YSIG_MAXK); return -1; }
        phase_begin(c);
        TAKE_OR_FAIL(d_iptr, to_dev(c, LS.ptr));
        TAKE_OR_FAIL(d_icol, to_dev(c, LS.col));
        TAKE_OR_FAIL(d_a, to_dev(c, LS.val));
        TAKE_OR_FAIL(d_ids, to_dev(c, LS.ids));
        TAKE_OR_FAIL(d_pptr, to_dev(c, LP.ptr));
        TAKE_OR_FAIL(d_pcol, to_dev(c, LP.col));
        TAKE_OR_FAIL(d_pval, to_dev(c, LP.val));
        TAKE_OR_FAIL(d_pids, to_dev(c, LP.ids));
        // reach of the BG reconstruction operator: the ring (bg_ssub = 1) or up(4 taps) o ring o down(8/scale taps)
        const int reach_s = (c->opt.bg_ssub > 1) ? c->opt.bg_ssub * ((c->ring_radius + c->opt.bg_ssub - 1) / c->opt.bg_ssub + 6) + 4 : c->rr;
        std::vector<int> bb = expand_bbox(LS, reach_s, P.nrb, P.ncb);
        TAKE_OR_FAIL(d_bbox, to_dev(c, bb));
        std::vector<double> snp(P.dp);
        for (int p = 0; p < P.dp; ++p) {
            int r = p % P.nr + P.patch.r0, cc = p / P.nr + P.patch.c0;
            snp[p] = c->sn[(size_t)cc * c->d1 + r];
        }
        TAKE_OR_FAIL(d_sn, to_dev(c, snp));
        TAKE_OR_FAIL(d_Cc, c->scr.take<double>((size_t)std::max(Ks, 1) * T));
        TAKE_OR_FAIL(d_Ccp, c->scr.take<double>((size_t)std::max(Kp, 1) * T));
        TAKE_OR_FAIL(d_V, c->scr.take<double>((size_t)std::max(Ks, 1) * std::max(Ks, 1)));
        TAKE_OR_FAIL(d_P2, c->scr.take<double>((size_t)std::max(Kp, 1) * std::max(Ks, 1)));
        TAKE_OR_FAIL(d_U, c->scr.take<double>(std::max<size_t>(nent, 1)));
        TAKE_OR_FAIL(d_D, c->scr.take<double>((size_t)P.db * std::max(Ks, 1)));
        TAKE_OR_FAIL(d_thr, c->scr.take<double>(P.dp));
        if (Ks > 0) {
            LAUNCH(gather_center_rows_kernel, Ks, 256, 0, c->st, c->C, d_ids, Ks, T, d_Cc, (double*)nullptr);
            dim3 gg(Ks, Ks);
            LAUNCH(small_gram_kernel, gg, 128, 0, c->st, d_Cc, Ks, d_Cc, Ks, T, 1, d_V, (double*)nullptr);
        }
        if (Kp > 0) {
            LAUNCH(gather_center_rows_kernel, Kp, 256, 0, c->st, c->Cprev, d_pids, Kp, T, d_Ccp, (double*)nullptr);
            if (Ks > 0) {
                dim3 gg(Kp, Ks);
                LAUNCH(small_gram_kernel, gg, 128, 0, c->st, d_Ccp, Kp, d_Cc, Ks, T, 1, d_P2, (double*)nullptr);
            }
        }
        phase_end(c, 6);
        tick("sp uploads+small grams");
        // ---- optional explicit rows of the BG-subtracted video (update_sn: GetSn per pixel, :191-194; lars: energy)
        if (update_sn || lars) {
            phase_begin(c);
            double* d_rowsY = c->scr.take<double>((size_t)CH * T);
            int* d_rows = c->scr.take<int>(CH);
            double* d_tmp = c->scr.take<double>(CH);
            if (!d_rowsY || !d_rows || !d_tmp) { set_error("scratch exhausted"); return -1; }
            auto run_rows = [&](int nrows_total, bool want_sn, std::vector<double>& outv) -> int {
                outv.assign(nrows_total, 0.0);
                std::vector<int> rows(CH);
                for (int b = 0; b < nrows_total; b += CH) {
                    const int nr = std::min(CH, nrows_total - b);
                    for (int i = 0; i < nr; ++i) rows[i] = b + i;
                    CNMFE_CUDA_OK(cudaMemcpyAsync(d_rows, rows.data(), (size_t)nr * 4, cudaMemcpyHostToDevice, c->st));
                    dim3 gg(nr, (T + YSIG_TCHUNK - 1) / YSIG_TCHUNK);
                    LAUNCH(ysig_rows_kernel, gg, 256, 0, c->st, g, c->d_off_r, c->d_off_c, P.W, P.b0, P.Yt, P.Ymean, T,
                           c->Tpad, d_pptr, d_pcol, d_pval, Kp, d_Ccp, d_rows, d_rowsY);
                    if (want_sn) { if (getsn_batch_dev(d_rowsY, T, nr, d_tmp, &c->arena, c->st)) return -1; }
                    else LAUNCH(rows_centered_energy_kernel, nr, 256, 0, c->st, d_rowsY, T, d_tmp);
                    CNMFE_CUDA_OK(cudaMemcpyAsync(outv.data() + b, d_tmp, (size_t)nr * 8, cudaMemcpyDeviceToHost, c->st));
                    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
                }
                return 0;
            };
            if (update_sn) {
                std::vector<double> snew;
                if (run_rows(P.dp, true, snew)) return -1;
                for (int p = 0; p < P.dp; ++p) {
                    int r = p % P.nr + P.patch.r0, cc = p / P.nr + P.patch.c0;
                    c->sn[(size_t)cc * c->d1 + r] = snew[p];
                    snp[p] = snew[p];
                }
                CNMFE_CUDA_OK(cudaMemcpyAsync(d_sn, snp.data(), (size_t)P.dp * 8, cudaMemcpyHostToDevice, c->st));
            }
            if (lars && Ks > 0) {
                // thresh = sn.^2*T - sum(Y.^2,2) indexed by the LOOP COUNTER m over ind_fit (lars_spatial.m:50,55)
                std::vector<int> fit;
                for (int p = 0; p < P.dp; ++p) if (LS.ptr[p + 1] > LS.ptr[p]) fit.push_back(p);
                std::vector<double> en, thr(P.dp, 0.0);
                if (run_rows((int)fit.size(), false, en)) return -1;
                for (size_t m = 0; m < fit.size(); ++m) thr[fit[m]] = snp[m] * snp[m] * (double)T - en[m];
                CNMFE_CUDA_OK(cudaMemcpyAsync(d_thr, thr.data(), (size_t)P.dp * 8, cudaMemcpyHostToDevice, c->st));
                CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
            }
            phase_end(c, 6);
        }
        if (Ks == 0) continue;
        phase_begin(c);
        CNMFE_CUDA_OK(cudaMemsetAsync(d_D, 0, (size_t)P.db * Ks * 8, c->st));
        {
            // the BG update of this iteration projected the same centred traces on a larger box around every neuron: reuse it
            bool cached = P.mc_valid && P.mc_Cver == c->C_version && !ssub;
            std::vector<int> kmap(Ks, -1);
            for (int k = 0; cached && k < Ks; ++k) {
                const auto it = std::lower_bound(P.mc_ids.begin(), P.mc_ids.end(), LS.ids[k]);
                if (it == P.mc_ids.end() || *it != LS.ids[k]) { cached = false; break; }
                const int kb = (int)(it - P.mc_ids.begin());
                const int* sb = &bb[4 * k]; const int* cb = &P.mc_bbox[4 * kb];
                if (sb[1] >= sb[0] && (sb[0] < cb[0] || sb[1] > cb[1] || sb[2] < cb[2] || sb[3] > cb[3])) { cached = false; break; }
                kmap[k] = kb;
            }
            if (cached) {
                TAKE_OR_FAIL(d_kmap, to_dev(c, kmap));
                dim3 gg((unsigned)P.db, (Ks + 127) / 128);
                LAUNCH(proj_from_cache_kernel, gg, 128, 0, c->st, P.Mcache, P.mc_K, d_kmap, d_bbox, P.nrb, Ks, d_D);
            } else {
                launch_proj_mc(c->st, P.Yt, P.Ymean, P.nrb, P.ncb, T, c->Tpad, 1,
                       d_Cc, Ks, d_bbox, d_D);
            }
        }
        if (Kp > 0) LAUNCH(spatial_make_D_kernel, P.db, 64, 0, c->st, d_D, Ks, d_pptr, d_pcol, d_pval, d_P2);
        if (!ssub) {
            LAUNCH(spatial_U_kernel, (unsigned)(((size_t)P.dp * 32 + 255) / 256), 256, 0, c->st, g, c->d_off_r, c->d_off_c,
                   P.W, d_D, Ks, d_iptr, d_icol, d_pptr, d_pcol, d_pval, d_P2, d_U);
        } else {
            // Bf = imresize(W * imresize(R - mean(R), 1/ssub), [nr_block nc_block]) projected on Cc (:167-177)
            if (ssub_ensure_patch(c, ip, false)) return -1;
            double* d_F = c->scr.take<double>((size_t)P.db * Ks);
            double* d_t1 = c->scr.take<double>((size_t)P.db * Ks);
            double* d_t2 = c->scr.take<double>((size_t)P.db * Ks);
            if (!d_F || !d_t1 || !d_t2) { set_error("scratch exhausted"); return -1; }
            if (ssub_forward(c, ip, d_D, Ks, d_F, d_t1, d_t2)) return -1;
            LAUNCH(spatial_U_ssub_kernel, (P.dp + 127) / 128, 128, 0, c->st, P.dp, P.nr, P.nrb, g.pr_off, g.pc_off, d_D, d_F,
                   Ks, d_iptr, d_icol, d_pptr, d_pcol, d_pval, d_P2, d_U);
        }
        phase_end(c, 2);
        tick("sp projections");
        phase_begin(c);
        CNMFE_CUDA_OK(cudaMemsetAsync(c->d_err, 0, 4, c->st));
        LAUNCH(spatial_solve_kernel, (P.dp + 127) / 128, 128, 0, c->st, P.dp, d_iptr, d_icol, d_U, d_V, Ks, d_sn,
               c->opt.spatial_algorithm, 3, d_thr, d_a, c->d_err);
        std::vector<double> anew(nent);
        int err = 0;
        CNMFE_CUDA_OK(cudaMemcpyAsync(anew.data(), d_a, nent * 8, cudaMemcpyDeviceToHost, c->st));
        CNMFE_CUDA_OK(cudaMemcpyAsync(&err, c->d_err, 4, cudaMemcpyDeviceToHost, c->st));
        phase_end(c, 3);
        CNMFE_CUDA_OK(cudaGetLastError());
        if (err) { set_error("update_spatial: more than %d search masks overlap one pixel", SPATIAL_MAXROW); return -1; }
        for (size_t e = 0; e < nent; ++e) c->A_on_IND[LS.entry_src[e]] = anew[e];
        tick("sp solve+readback");
    }
    c->have_spatial = true;
    HostTick tick2;
    const int rc = cnmfe_set_spatial(c, c->A_on_IND.data());
    tick2("sp set_spatial (A <- IND)");
    return rc;
}

// obj.A = A_new on the pattern (zeros dropped), update_spatial_parallel.m:321-335
extern "C" int cnmfe_set_spatial(cnmfe_ctx* c, const double* vals) {
    if (!c || (!vals && !c->IND.ir.empty())) { set_error("cnmfe_set_spatial: null"); return -1; }   // empty pattern (K = 0): nothing to read
    if (c->IND.K != c->K) { set_error("cnmfe_set_spatial: search mask has %d columns, A has %d", c->IND.K, c->K); return -1; }
    if (vals && vals != c->A_on_IND.data()) c->A_on_IND.assign(vals, vals + c->IND.ir.size());
    HostCsc An;
    An.K = c->K;
    An.jc.assign(c->K + 1, 0);
    for (int k = 0; k < c->K; ++k) {
        for (int64_t e = c->IND.jc[k]; e < c->IND.jc[k + 1]; ++e)
            if (c->A_on_IND[e] != 0.0) { An.ir.push_back(c->IND.ir[e]); An.pr.push_back(c->A_on_IND[e]); }
        An.jc[k + 1] = (int64_t)An.ir.size();
    }
    c->A = An;
    c->have_spatial = true;
    return 0;
}

extern "C" int cnmfe_get_spatial(cnmfe_ctx* c, double* A_on_IND) {
    if (!c || (!A_on_IND && !c->A_on_IND.empty())) { set_error("cnmfe_get_spatial: null"); return -1; }
    if (!c->have_spatial) { set_error("cnmfe_get_spatial: no spatial update has run"); return -1; }
    std::copy(c->A_on_IND.begin(), c->A_on_IND.end(), A_on_IND);
    return 0;
}

// ===================================================================================================== temporal
extern "C" int cnmfe_update_temporal_patches(cnmfe_ctx* c) {
    if (!c) { set_error("cnmfe_update_temporal: null ctx"); return -1; }
    if (c->opt.background_model >= 1) return update_temporal_patches_svd(c);
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    for (float& f : c->phase_ms) f = 0;
    const int T = c->T, K = c->K;
    if (K == 0) return 0;
    CNMFE_CUDA_OK(cudaMemsetAsync(c->num, 0, (size_t)K * T * 8, c->st));
    CNMFE_CUDA_OK(cudaMemsetAsync(c->den, 0, (size_t)K * 8, c->st));
    // ---- host planning for every owned patch first: the Gauss-Seidel sweeps of ALL patches then run as ONE launch over the
    //      concatenated local neurons (block-diagonal V: neurons of different patches never wait for each other), so a rank
    //      holding several patches pays the dependency-chain latency of HALS_temporal once, not once per patch
    struct TPlan { int ip, Kt, Kp, koff; LocalSparse LA, LPown; const LocalSparse* LP; };
    std::vector<TPlan> plans;
    plans.reserve(c->npatch);
    int Ktot = 0;
    size_t per_patch_need = 0, vnnz_bound = 0;
    for (int ip = 0; ip < c->npatch; ++ip) {
        Patch& P = c->patches[ip];
        if (!P.owned) continue;
        if (!P.uploaded) { set_error("update_temporal: block %d not uploaded", ip); return -1; }
        plans.emplace_back();
        TPlan& pl = plans.back();
        pl.ip = ip;
        HostTick tick;
        build_local(c, P, c->A, SEL_SUM_BLOCK, ROWS_BLOCK_PATCHONLY, nullptr, &pl.LA);
        tick("tp build_local A");
        pl.Kt = pl.LA.K();
        if (pl.Kt == 0) { plans.pop_back(); continue; }   // update_temporal_parallel.m:123-126
        LocalSparse& LA = pl.LA;
        const int Kt = pl.Kt;
        if (!c->use_c_hat) {
            // fast_temporal (update_temporal_parallel.m:314-337): keep only the pixels with A >= 0.5*max(A) per neuron
            std::vector<double> amax(Kt, 0.0);
            for (int k = 0; k < Kt; ++k)
                for (int e = LA.cptr[k]; e < LA.cptr[k + 1]; ++e) amax[k] = std::max(amax[k], LA.cval[e]);
            for (int k = 0; k < Kt; ++k)
                for (int e = LA.cptr[k]; e < LA.cptr[k + 1]; ++e)
                    if (!(LA.cval[e] / amax[k] >= 0.5)) LA.cval[e] = 0.0;
            for (size_t e = 0; e < LA.col.size(); ++e)
                if (!(LA.val[e] / amax[LA.col[e]] >= 0.5)) LA.val[e] = 0.0;
        }
        if (!P.lp_valid) build_local(c, P, c->Aprev, SEL_SUM_BLOCK, ROWS_BLOCK, nullptr, &pl.LPown);
        pl.LP = P.lp_valid ? &P.lp_cache : &pl.LPown;
        tick("tp build_local Aprev");
        pl.Kp = pl.LP->K();
        pl.koff = Ktot;
        Ktot += Kt;
        vnnz_bound += (size_t)Kt * Kt;
        const int Kp = pl.Kp;
        size_t need = (c->opt.bg_ssub > 1 ? 4 : 1) * pad256((size_t)P.db * Kt * 8) + pad256((size_t)std::max(Kp, 1) * T * 8) +
                      2 * pad256((size_t)Kt * Kt * 12 + 64) + pad256((size_t)Kt * std::max(Kp, 1) * 8) +
                      2 * pad256((size_t)(P.db + 1) * 4) + pad256(LA.col.size() * 24 + 64) + pad256(pl.LP->col.size() * 24 + 64) +
                      pad256((size_t)(Kt + Kp) * 128 + 64) + (1 << 20);
        per_patch_need = std::max(per_patch_need, need);
    }
    if (Ktot == 0) { CNMFE_CUDA_OK(cudaStreamSynchronize(c->st)); return 0; }
    const size_t persistent = 4 * pad256((size_t)Ktot * T * 8) + 6 * pad256((size_t)(Ktot + 1) * 16) + 2 * pad256(vnnz_bound * 12 + 64) + (1 << 20);
    if (c->scr.reserve(persistent + per_patch_need)) return -1;
    c->scr.reset();
    TAKE_OR_FAIL(d_Uall, c->scr.take<double>((size_t)Ktot * T));
    TAKE_OR_FAIL(d_Clall, c->scr.take<double>((size_t)Ktot * T));
    TAKE_OR_FAIL(d_Crawall, c->scr.take<double>((size_t)Ktot * T));
    TAKE_OR_FAIL(d_Sall, c->scr.take<double>((size_t)Ktot * T));
    TAKE_OR_FAIL(d_snall, c->scr.take<double>(Ktot));
    TAKE_OR_FAIL(d_parsall, c->scr.take<double>((size_t)Ktot * 2));
    TAKE_OR_FAIL(d_aaall, c->scr.take<double>(Ktot));
    TAKE_OR_FAIL(d_idsall, c->scr.take<int>(Ktot));
    TAKE_OR_FAIL(d_doneall, c->scr.take<int>(Ktot));
    TAKE_OR_FAIL(d_orderall, c->scr.take<int>(Ktot + 1));
    const size_t base_off = c->scr.off;
    std::vector<int> vptr_all(1, 0), vidx_all;
    std::vector<double> vval_all, aa_all((size_t)Ktot);
    vidx_all.reserve(vnnz_bound / 4 + Ktot); vval_all.reserve(vnnz_bound / 4 + Ktot);
    for (TPlan& pl : plans) {
        const int ip = pl.ip, Kt = pl.Kt, Kp = pl.Kp;
        Patch& P = c->patches[ip];
        const LocalSparse& LA = pl.LA;
        const LocalSparse& LP = *pl.LP;
        HostTick tick;
        c->scr.off = base_off;
        const RingGeom& g = P.geom;
        phase_begin(c);
        TAKE_OR_FAIL(d_aptr, to_dev(c, LA.ptr));
        TAKE_OR_FAIL(d_acol, to_dev(c, LA.col));
        TAKE_OR_FAIL(d_aval, to_dev(c, LA.val));
        TAKE_OR_FAIL(d_cptr, to_dev(c, LA.cptr));
        TAKE_OR_FAIL(d_crow, to_dev(c, LA.crow));
        TAKE_OR_FAIL(d_cval, to_dev(c, LA.cval));
        int* d_ids = d_idsall + pl.koff;
        CNMFE_CUDA_OK(cudaMemcpyAsync(d_ids, LA.ids.data(), (size_t)Kt * 4, cudaMemcpyHostToDevice, c->st));
        TAKE_OR_FAIL(d_pcptr, to_dev(c, LP.cptr));
        TAKE_OR_FAIL(d_pcrow, to_dev(c, LP.crow));
        TAKE_OR_FAIL(d_pcval, to_dev(c, LP.cval));
        TAKE_OR_FAIL(d_pids, to_dev(c, LP.ids));
        const int reach_t = (c->opt.bg_ssub > 1) ? c->opt.bg_ssub * ((c->ring_radius + c->opt.bg_ssub - 1) / c->opt.bg_ssub + 6) + 4 : c->rr;
        std::vector<int> bb = expand_bbox(LA, reach_t, P.nrb, P.ncb);
        TAKE_OR_FAIL(d_bbox, to_dev(c, bb));
        TAKE_OR_FAIL(d_B, c->scr.take<double>((size_t)P.db * Kt));
        double* d_U = d_Uall + (size_t)pl.koff * T;
        double* d_Cl = d_Clall + (size_t)pl.koff * T;
        double* d_Crawl = d_Crawall + (size_t)pl.koff * T;
        TAKE_OR_FAIL(d_Ccp, c->scr.take<double>((size_t)std::max(Kp, 1) * T));
        TAKE_OR_FAIL(d_AWA, c->scr.take<double>((size_t)Kt * std::max(Kp, 1)));
        TAKE_OR_FAIL(d_cst, c->scr.take<double>(Kt));
        TAKE_OR_FAIL(d_V, c->scr.take<double>((size_t)Kt * Kt));
        CNMFE_CUDA_OK(cudaMemsetAsync(d_B, 0, (size_t)P.db * Kt * 8, c->st));
        if (c->opt.bg_ssub <= 1) {
            LAUNCH(temporal_build_negWtA_kernel, (P.db + 127) / 128, 128, 0, c->st, g, c->d_off_r, c->d_off_c, P.W, d_aptr,
                   d_acol, d_aval, Kt, d_B);
        } else {
            // B = -(down' * W' * up') * A_patch   (transpose of the BG reconstruction operator)
            if (ssub_ensure_patch(c, ip, false)) return -1;
            double* d_Ai = c->scr.take<double>((size_t)P.db * Kt);
            double* d_t1 = c->scr.take<double>((size_t)P.db * Kt);
            double* d_t2 = c->scr.take<double>((size_t)P.db * Kt);
            if (!d_Ai || !d_t1 || !d_t2) { set_error("scratch exhausted"); return -1; }
            CNMFE_CUDA_OK(cudaMemsetAsync(d_Ai, 0, (size_t)P.db * Kt * 8, c->st));
            LAUNCH(csr_to_dense_kernel, (P.db + 127) / 128, 128, 0, c->st, d_aptr, d_acol, d_aval, P.db, Kt, d_Ai);
            if (ssub_transpose(c, ip, d_Ai, Kt, d_B, d_t1, d_t2)) return -1;
            LAUNCH(negate_kernel, (unsigned)(((size_t)P.db * Kt + 255) / 256), 256, 0, c->st, d_B, (size_t)P.db * Kt);
        }
        if (Kp > 0) {
            dim3 gg(Kt, (Kp + 63) / 64);
            LAUNCH(temporal_AWA_kernel, gg, 64, 0, c->st, d_B, Kt, d_pcptr, d_pcrow, d_pcval, Kp, d_AWA);
            LAUNCH(gather_center_rows_kernel, Kp, 256, 0, c->st, c->Cprev, d_pids, Kp, T, d_Ccp, (double*)nullptr);
        }
        LAUNCH(temporal_add_A_kernel, (Kt + 63) / 64, 64, 0, c->st, g, d_cptr, d_crow, d_cval, Kt, P.Ymean, P.b0, d_B,
               d_cst);
        { dim3 gg(Kt, (Kt + 127) / 128); LAUNCH(temporal_V_kernel, gg, 128, 0, c->st, d_cptr, d_crow, d_cval, Kt, d_V); }
        phase_end(c, 6);
        tick("tp uploads+B build");
        phase_begin(c);
        launch_proj_bt(c->st, P.Yt, P.Ymean, P.nrb, T, c->Tpad, d_B, Kt, d_bbox, d_U);
        { dim3 gg((T + 255) / 256, Kt); LAUNCH(add_small_matmul_kernel, gg, 256, 0, c->st, d_U, Kt, T, d_cst, d_AWA, Kp, d_Ccp); }
        { dim3 gg((T + 255) / 256, Kt); LAUNCH(gather_rows_kernel, gg, 256, 0, c->st, c->C, d_ids, Kt, T, d_Cl); }
        if (!c->use_c_hat) {
            // C_raw = (tmp_A'*Y) ./ aa  (aa = 0 -> row of zeros, weight 0)
            dim3 gg((T + 255) / 256, Kt);
            LAUNCH(rows_div_diag_kernel, gg, 256, 0, c->st, d_U, d_V, Kt, T, d_Crawl);
        }
        phase_end(c, 2);
        tick("tp projection");
        // V -> CSR (host), appended to the block-diagonal system of all patches
        std::vector<double> V((size_t)Kt * Kt);
        CNMFE_CUDA_OK(cudaMemcpyAsync(V.data(), d_V, V.size() * 8, cudaMemcpyDeviceToHost, c->st));
        CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
        for (int k = 0; k < Kt; ++k) {
            for (int j = 0; j < Kt; ++j) {
                double v = V[(size_t)k * Kt + j];
                if (v != 0.0 || j == k) { vidx_all.push_back(pl.koff + j); vval_all.push_back(v); }
            }
            vptr_all.push_back((int)vidx_all.size());
            aa_all[(size_t)pl.koff + k] = V[(size_t)k * Kt + k];
        }
        tick("tp V csr");
    }
    // ---- the sweeps of all patches in one launch
    c->scr.off = base_off;
    phase_begin(c);
    TAKE_OR_FAIL(d_vptr, to_dev(c, vptr_all));
    TAKE_OR_FAIL(d_vidx, to_dev(c, vidx_all));
    TAKE_OR_FAIL(d_vval, to_dev(c, vval_all));
    CNMFE_CUDA_OK(cudaMemcpyAsync(d_aaall, aa_all.data(), (size_t)Ktot * 8, cudaMemcpyHostToDevice, c->st));
    if (c->use_c_hat) {
        CNMFE_CUDA_OK(cudaMemsetAsync(d_Crawall, 0, (size_t)Ktot * T * 8, c->st));
        CNMFE_CUDA_OK(cudaMemsetAsync(d_Sall, 0, (size_t)Ktot * T * 8, c->st));
        CNMFE_CUDA_OK(cudaMemsetAsync(d_parsall, 0, (size_t)Ktot * 16, c->st));
        if (hals_temporal_dev(d_Uall, d_vptr, d_vidx, d_vval, d_aaall, Ktot, T, c->opt.maxIter_temporal, c->opt.deconv_flag,
                              c->opt.deconv, d_Clall, d_Crawall, d_Sall, d_snall, d_parsall, d_doneall, c->d_ticket, d_orderall,
                              &c->arena, c->st)) return -1;
    }
    // energy-weighted accumulation, patch after patch (the order the reference adds them, update_temporal_parallel.m:269-280)
    for (TPlan& pl : plans) {
        dim3 gg((T + 255) / 256, pl.Kt);
        LAUNCH(temporal_merge_aa_kernel, gg, 256, 0, c->st, d_Crawall + (size_t)pl.koff * T, d_aaall + pl.koff, T, d_idsall + pl.koff,
               c->num, c->den);
    }
    phase_end(c, 4);
    CNMFE_CUDA_OK(cudaGetLastError());
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cnmfe_temporal_merge_buffers(cnmfe_ctx* c, double** num_dev, double** den_dev) {
    if (!c || !num_dev || !den_dev) { set_error("cnmfe_temporal_merge_buffers: null"); return -1; }
    *num_dev = c->num; *den_dev = c->den;
    return 0;
}

extern "C" int cnmfe_update_temporal_finish(cnmfe_ctx* c) {
    if (!c) { set_error("cnmfe_update_temporal_finish: null ctx"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    const int T = c->T, K = c->K;
    if (K == 0) return 0;
    ++c->C_version;
    phase_begin(c);
    { dim3 gg((T + 255) / 256, K); LAUNCH(temporal_divide_kernel, gg, 256, 0, c->st, c->num, c->den, K, T); }
    if (c->opt.deconv_flag) {
        // obj.C = obj.deconvTemporal()  (update_temporal_parallel.m:283)
        if (deconv_batch_dev(c->num, T, K, c->opt.deconv, nullptr, nullptr, 1, c->C, c->S, c->Craw, c->outs, &c->arena,
                             c->st)) return -1;
    } else {
        LAUNCH(rows_sub_min_kernel, K, 256, 0, c->st, c->num, T);
        CNMFE_CUDA_OK(cudaMemcpyAsync(c->Craw, c->num, (size_t)K * T * 8, cudaMemcpyDeviceToDevice, c->st));
        CNMFE_CUDA_OK(cudaMemcpyAsync(c->C, c->num, (size_t)K * T * 8, cudaMemcpyDeviceToDevice, c->st));
        CNMFE_CUDA_OK(cudaMemsetAsync(c->outs, 0, (size_t)K * 48, c->st));
    }
    phase_end(c, 5);
    CNMFE_CUDA_OK(cudaGetLastError());
    return 0;
}

// Multi-GPU variant of cnmfe_update_temporal_finish: this rank divides and deconvolves only the traces
// [k0, k1) of the merged C_raw; the other rows of C, C_raw, S and of the per-trace outputs are zeroed, so that a SUM
// all-reduce over ranks holding disjoint ranges assembles the full result (x + 0 is exact).  Buffers for that exchange:
// cnmfe_temporal_state_buffers.  OPT-IN (Sources2D option shard_deconv): written without GPU time left in round 1 --
// validate against the unsharded path on >= 2 GPUs before making it the default.
extern "C" int cnmfe_update_temporal_finish_part(cnmfe_ctx* c, int k0, int k1) {
    if (!c) { set_error("cnmfe_update_temporal_finish_part: null ctx"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    const int T = c->T, K = c->K;
    if (K == 0) return 0;
    if (k0 < 0 || k1 > K || k0 > k1) { set_error("cnmfe_update_temporal_finish_part: range [%d, %d) outside [0, %d)", k0, k1, K); return -1; }
    ++c->C_version;
    phase_begin(c);
    const int n = k1 - k0;
    for (double* buf : {c->C, c->Craw, c->S}) {
        if (k0 > 0) CNMFE_CUDA_OK(cudaMemsetAsync(buf, 0, (size_t)k0 * T * 8, c->st));
        if (k1 < K) CNMFE_CUDA_OK(cudaMemsetAsync(buf + (size_t)k1 * T, 0, (size_t)(K - k1) * T * 8, c->st));
    }
    CNMFE_CUDA_OK(cudaMemsetAsync(c->outs, 0, (size_t)K * 48, c->st));
    if (n > 0) {
        double* num = c->num + (size_t)k0 * T;
        { dim3 gg((T + 255) / 256, n); LAUNCH(temporal_divide_kernel, gg, 256, 0, c->st, num, c->den + k0, n, T); }
        if (c->opt.deconv_flag) {
            if (deconv_batch_dev(num, T, n, c->opt.deconv, nullptr, nullptr, 1, c->C + (size_t)k0 * T, c->S + (size_t)k0 * T,
                                 c->Craw + (size_t)k0 * T, c->outs + (size_t)k0 * 6, &c->arena, c->st)) return -1;
        } else {
            LAUNCH(rows_sub_min_kernel, n, 256, 0, c->st, num, T);
            CNMFE_CUDA_OK(cudaMemcpyAsync(c->Craw + (size_t)k0 * T, num, (size_t)n * T * 8, cudaMemcpyDeviceToDevice, c->st));
            CNMFE_CUDA_OK(cudaMemcpyAsync(c->C + (size_t)k0 * T, num, (size_t)n * T * 8, cudaMemcpyDeviceToDevice, c->st));
        }
    }
    phase_end(c, 5);
    CNMFE_CUDA_OK(cudaGetLastError());
    return 0;
}

// device buffers of obj.C, obj.C_raw, obj.S (K x T each, trace contiguous) and of the 6 per-trace outputs (K x 6)
extern "C" int cnmfe_temporal_state_buffers(cnmfe_ctx* c, double** C_dev, double** Craw_dev, double** S_dev, double** outs_dev) {
    if (!c || !C_dev || !Craw_dev || !S_dev || !outs_dev) { set_error("cnmfe_temporal_state_buffers: null"); return -1; }
    ++c->C_version;      // the caller is about to write into C (cross-rank exchange)
    *C_dev = c->C; *Craw_dev = c->Craw; *S_dev = c->S; *outs_dev = c->outs;
    return 0;
}

extern "C" int cnmfe_update_temporal(cnmfe_ctx* c) {
    if (cnmfe_update_temporal_patches(c)) return -1;
    return cnmfe_update_temporal_finish(c);
}

extern "C" int cnmfe_host_register(void* p, size_t bytes) {
    if (!p || !bytes) { set_error("cnmfe_host_register: null"); return -1; }
    CNMFE_CUDA_OK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return 0;
}
extern "C" int cnmfe_host_unregister(void* p) {
    if (!p) { set_error("cnmfe_host_unregister: null"); return -1; }
    CNMFE_CUDA_OK(cudaHostUnregister(p));
    return 0;
}

extern "C" int cnmfe_set_trace_major(cnmfe_ctx* c, int on) {
    if (!c) { set_error("cnmfe_set_trace_major: null ctx"); return -1; }
    c->trace_major = on ? 1 : 0;
    return 0;
}

extern "C" int cnmfe_set_use_c_hat(cnmfe_ctx* c, int use_c_hat) {
    if (!c) { set_error("cnmfe_set_use_c_hat: null ctx"); return -1; }
    if (!use_c_hat && c->opt.background_model != 0) { set_error("use_c_hat=false is built for the ring model only"); return -1; }
    c->use_c_hat = use_c_hat ? 1 : 0;
    return 0;
}

extern "C" int cnmfe_get_temporal(cnmfe_ctx* c, double* C, double* C_raw, double* S, double* kernel_pars,
                                  double* neuron_sn) {
    if (!c) { set_error("cnmfe_get_temporal: null ctx"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    const int K = c->K;
    if (K == 0) return 0;
    if (download_KT(c, c->C, K, C) || download_KT(c, c->Craw, K, C_raw) || download_KT(c, c->S, K, S)) return -1;
    if (kernel_pars || neuron_sn) {
        std::vector<double> outs((size_t)K * 6);
        CNMFE_CUDA_OK(cudaMemcpy(outs.data(), c->outs, (size_t)K * 48, cudaMemcpyDeviceToHost));
        for (int k = 0; k < K; ++k) {
            if (kernel_pars) { kernel_pars[2 * k] = outs[6 * k + 1]; kernel_pars[2 * k + 1] = outs[6 * k + 2]; }
            if (neuron_sn) neuron_sn[k] = outs[6 * k + 5];
        }
    }
    return 0;
}

// ===================================================================================================== test hooks
// Second moments S2[q][id] of block `ip` (all frames) by the tensor (use_tensor=1) or SIMT (0) kernel -> host.
extern "C" int cnmfe_debug_second_moments(cnmfe_ctx* c, int ip, int use_tensor, double* out) {
    if (!c || ip < 0 || ip >= c->npatch || !out) { set_error("cnmfe_debug_second_moments: bad arguments"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    Patch& P = c->patches[ip];
    if (!P.owned || !P.uploaded) { set_error("cnmfe_debug_second_moments: block not resident"); return -1; }
    const size_t ND = (size_t)ring_num_disp(c->rr);
    if (c->scr.reserve(ND * P.db * 8 + (1 << 20))) return -1;
    c->scr.reset();
    double* d_S2 = c->scr.take<double>(ND * P.db);
    CNMFE_CUDA_OK(cudaMemsetAsync(d_S2, 0, ND * P.db * 8, c->st));
    if (use_tensor) {
        int rc = ring_s2_tensor(P.hi, P.lo, P.nrb, P.ncb, c->T, c->Tpad, c->rr, d_S2, c->st);
        if (rc < 0) return -1;
        if (rc > 0) { set_error("tensor kernel does not support this shape"); return -1; }
    } else {
        long long nw = (long long)((P.nrb + 3) / 4) * P.ncb;
        dim3 gg((unsigned)((nw + 7) / 8), c->ngroups);
        LAUNCH(ring_s2_simt_kernel, gg, 256, 0, c->st, P.Yt, P.nrb, P.ncb, c->T, c->Tpad, 1, c->rr, c->d_groups,
               c->ngroups, d_S2, ND);
    }
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    CNMFE_CUDA_OK(cudaGetLastError());
    CNMFE_CUDA_OK(cudaMemcpy(out, d_S2, ND * P.db * 8, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int cnmfe_last_gram_was_tensor(cnmfe_ctx* c) { return c ? c->last_gram_tensor : 0; }
extern "C" int cnmfe_last_gram_frames(cnmfe_ctx* c) { return c ? c->last_gram_frames : 0; }
extern "C" long long cnmfe_last_active_pixels(cnmfe_ctx* c) { return c ? c->last_active_pixels : 0; }
extern "C" int cnmfe_last_nmf_iterations(cnmfe_ctx* c) { return c ? c->last_nmf_iters : 0; }

// sn = estimate_noise(obj, frame_range, 'psd') (@Sources2D/Sources2D.m:328-379) from the RESIDENT video: per-pixel GetSn
// (OASIS_matlab/functions/GetSn.m) of the raw frames [f0, f1] (1-based inclusive) for every pixel of the owned patches.
extern "C" int cnmfe_estimate_noise(cnmfe_ctx* c, int f0, int f1, double* sn) {
    if (!c || !sn) { set_error("cnmfe_estimate_noise: null"); return -1; }
    if (f0 < 1 || f1 > c->T || f1 - f0 + 1 < 32) { set_error("cnmfe_estimate_noise: frame range [%d, %d] outside [1, %d] or shorter than 32", f0, f1, c->T); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    const int n = f1 - f0 + 1, CH = 4096;
    if (c->scr.reserve(std::max(c->scr.cap, pad256((size_t)CH * n * 8) + pad256((size_t)CH * 8) + 4096))) return -1;
    std::vector<double> tmp(CH);
    for (int ip = 0; ip < c->npatch; ++ip) {
        Patch& P = c->patches[ip];
        if (!P.owned) continue;
        if (!P.uploaded) { set_error("cnmfe_estimate_noise: block %d not uploaded", ip); return -1; }
        c->scr.reset();
        TAKE_OR_FAIL(d_rows, c->scr.take<double>((size_t)CH * n));
        TAKE_OR_FAIL(d_sn, c->scr.take<double>(CH));
        for (int p0 = 0; p0 < P.dp; p0 += CH) {
            const int m = std::min(CH, P.dp - p0);
            dim3 gg((n + 255) / 256, m);   // m <= 4096 rows
            LAUNCH(rows_u16_to_f64_kernel, gg, 256, 0, c->st, P.Yt, c->Tpad, P.nrb, P.geom.pr_off, P.geom.pc_off, P.nr, p0,
                   f0 - 1, n, d_rows);
            if (getsn_batch_dev(d_rows, n, m, d_sn, &c->arena, c->st)) return -1;
            CNMFE_CUDA_OK(cudaMemcpyAsync(tmp.data(), d_sn, (size_t)m * 8, cudaMemcpyDeviceToHost, c->st));
            CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
            for (int i = 0; i < m; ++i) {
                const int p = p0 + i, r = p % P.nr + P.patch.r0, cc = p / P.nr + P.patch.c0;
                sn[(size_t)cc * c->d1 + r] = tmp[i];
            }
        }
    }
    CNMFE_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- compute_RSS / reconstruct_background (Sources2D.m:1247-1510), ring model with bg_ssub = 1.
// Both walk the patch in chunks of explicit BG-subtracted rows (ysig_rows_kernel, the kernel behind update_sn) -- see the algebra
// at bg_cst_kernel.  mode 0: rss_out[0] = RSS of patch ip over frames [f0, f1] (1-based inclusive).  mode 1: ybg_out =
// background of the patch for those frames, d_patch x nframes in the boundary layout.
static int bg_rows_pass(cnmfe_ctx* c, int ip, int f0, int f1, const double* b0_map, const double* b0_new_map, int mode,
                        double* rss_out, double* ybg_out) {
    if (c->opt.background_model != 0 || c->opt.bg_ssub != 1) { set_error("compute_RSS / reconstruct_background are built for the ring model with bg_ssub = 1"); return -1; }
    if (!b0_map || !b0_new_map) { set_error("compute_RSS / reconstruct_background: b0 (reconstruct_b0) and b0_new maps are required"); return -1; }
    if (f0 < 1 || f1 > c->T || f1 < f0) { set_error("frame range [%d, %d] outside [1, %d]", f0, f1, c->T); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    Patch& P = c->patches[ip];
    if (!P.owned || !P.uploaded) { set_error("block %d not resident", ip); return -1; }
    const int T = c->T, nf = f1 - f0 + 1, CH = 1024;
    LocalSparse LA, LP;
    build_local(c, P, c->A, SEL_SUM_BLOCK, ROWS_BLOCK, nullptr, &LA);          // A(logical(mask), ind): every neuron of the block
    build_local(c, P, c->Aprev, SEL_SUM_BLOCK, ROWS_BLOCK, nullptr, &LP);
    const int Ka = LA.K(), Kp = LP.K();
    if (Kp > YSIG_MAXK) { set_error("%d previous neurons touch block %d (explicit rows handle <= %d)", Kp, ip, YSIG_MAXK); return -1; }
    size_t need = pad256((size_t)CH * T * 8) + pad256((size_t)std::max(Kp, 1) * T * 8) + 2 * pad256((size_t)(P.db + 1) * 4) +
                  pad256(LA.col.size() * 12 + 64) + pad256(LP.col.size() * 12 + 64) + 4 * pad256((size_t)P.db * 8) +
                  3 * pad256((size_t)P.dp * 8) + pad256((size_t)CH * nf * 8) * (mode ? 2 : 0) + pad256((size_t)(Ka + Kp + CH) * 16) + (1 << 20);
    if (c->scr.reserve(need)) return -1;
    c->scr.reset();
    const RingGeom& g = P.geom;
    TAKE_OR_FAIL(d_aptr, to_dev(c, LA.ptr));
    TAKE_OR_FAIL(d_acol, to_dev(c, LA.col));
    TAKE_OR_FAIL(d_aval, to_dev(c, LA.val));
    TAKE_OR_FAIL(d_aids, to_dev(c, LA.ids));
    TAKE_OR_FAIL(d_pptr, to_dev(c, LP.ptr));
    TAKE_OR_FAIL(d_pcol, to_dev(c, LP.col));
    TAKE_OR_FAIL(d_pval, to_dev(c, LP.val));
    TAKE_OR_FAIL(d_pids, to_dev(c, LP.ids));
    TAKE_OR_FAIL(d_Ccp, c->scr.take<double>((size_t)std::max(Kp, 1) * T));
    TAKE_OR_FAIL(d_Cpmean, c->scr.take<double>(std::max(Kp, 1)));
    if (Kp > 0) LAUNCH(gather_center_rows_kernel, Kp, 256, 0, c->st, c->Cprev, d_pids, Kp, T, d_Ccp, d_Cpmean);
    // block / patch vectors of the two b0 maps
    std::vector<double> b0blk(P.db), b0new(P.dp), b0p(P.dp);
    for (int q = 0; q < P.db; ++q) b0blk[q] = b0_map[(size_t)(q / P.nrb + P.block.c0) * c->d1 + (q % P.nrb + P.block.r0)];
    for (int p = 0; p < P.dp; ++p) {
        const size_t f = (size_t)(p / P.nr + P.patch.c0) * c->d1 + (p % P.nr + P.patch.r0);
        b0new[p] = b0_new_map[f]; b0p[p] = b0_map[f];
    }
    TAKE_OR_FAIL(d_b0blk, to_dev(c, b0blk));
    TAKE_OR_FAIL(d_b0new, to_dev(c, b0new));
    TAKE_OR_FAIL(d_b0p, to_dev(c, b0p));
    TAKE_OR_FAIL(d_meanR, c->scr.take<double>(P.db));
    TAKE_OR_FAIL(d_cst, c->scr.take<double>(P.dp));
    TAKE_OR_FAIL(d_rowsY, c->scr.take<double>((size_t)CH * T));
    TAKE_OR_FAIL(d_rows, c->scr.take<int>(CH));
    TAKE_OR_FAIL(d_part, c->scr.take<double>(CH));
    LAUNCH(bg_meanR_kernel, (P.db + 255) / 256, 256, 0, c->st, P.db, P.Ymean, d_pptr, d_pcol, d_pval, d_Cpmean, d_meanR);
    // the explicit rows use the b0 the caller's map holds for the patch (obj.b0{m}), like the reference (b0_ = reconstruct_b0())
    LAUNCH(bg_cst_kernel, (P.dp + 255) / 256, 256, 0, c->st, g, c->d_off_r, c->d_off_c, P.W, d_b0p, d_b0new, d_b0blk, d_meanR, d_cst);
    double* d_out = nullptr; double* d_tr = nullptr;
    if (mode == 1) {
        d_out = c->scr.take<double>((size_t)CH * nf); d_tr = c->scr.take<double>((size_t)CH * nf);
        if (!d_out || !d_tr) { set_error("scratch exhausted"); return -1; }
    }
    std::vector<int> rows(CH);
    std::vector<double> part(CH), chunk;
    double total = 0.0;
    for (int b = 0; b < P.dp; b += CH) {
        const int nr = std::min(CH, P.dp - b);
        for (int i = 0; i < nr; ++i) rows[i] = b + i;
        CNMFE_CUDA_OK(cudaMemcpyAsync(d_rows, rows.data(), (size_t)nr * 4, cudaMemcpyHostToDevice, c->st));
        dim3 gg(nr, (T + YSIG_TCHUNK - 1) / YSIG_TCHUNK);
        LAUNCH(ysig_rows_kernel, gg, 256, 0, c->st, g, c->d_off_r, c->d_off_c, P.W, d_b0p, P.Yt, P.Ymean, T, c->Tpad, d_pptr, d_pcol,
               d_pval, Kp, d_Ccp, d_rows, d_rowsY);
        if (mode == 0) {
            LAUNCH(rss_rows_kernel, nr, 256, 0, c->st, g, d_rowsY, d_rows, T, f0 - 1, f1, d_aptr, d_acol, d_aval, d_aids, c->C, d_cst, d_part);
            CNMFE_CUDA_OK(cudaMemcpyAsync(part.data(), d_part, (size_t)nr * 8, cudaMemcpyDeviceToHost, c->st));
            CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
            for (int i = 0; i < nr; ++i) total += part[i];
        } else {
            dim3 g2((nf + 255) / 256, nr);
            LAUNCH(ybg_rows_kernel, g2, 256, 0, c->st, g, d_rowsY, d_rows, P.Yt, T, c->Tpad, f0 - 1, f1, d_cst, d_out);
            if (c->trace_major) {        // [p][t]: rows b .. b+nr of the output
                CNMFE_CUDA_OK(cudaMemcpyAsync(ybg_out + (size_t)b * nf, d_out, (size_t)nr * nf * 8, cudaMemcpyDeviceToHost, c->st));
                CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
            } else {                     // MATLAB d_patch x nframes column-major: element (p, t) at p + t * d_patch
                chunk.resize((size_t)nr * nf);
                CNMFE_CUDA_OK(cudaMemcpyAsync(chunk.data(), d_out, (size_t)nr * nf * 8, cudaMemcpyDeviceToHost, c->st));
                CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
                for (int i = 0; i < nr; ++i)
                    for (int t = 0; t < nf; ++t) ybg_out[(size_t)(b + i) + (size_t)t * P.dp] = chunk[(size_t)i * nf + t];
            }
        }
    }
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    CNMFE_CUDA_OK(cudaGetLastError());
    if (mode == 0) *rss_out = total;
    return 0;
}

// [RSS_total, RSS] = compute_RSS(obj, frame_range) (Sources2D.m:1358-1510): rss[ip] for the owned patches (others untouched)
extern "C" int cnmfe_compute_rss(cnmfe_ctx* c, int f0, int f1, const double* b0_map, const double* b0_new_map, double* rss) {
    if (!c || !rss) { set_error("cnmfe_compute_rss: null"); return -1; }
    for (int ip = 0; ip < c->npatch; ++ip) {
        if (!c->patches[ip].owned) continue;
        if (bg_rows_pass(c, ip, f0, f1, b0_map, b0_new_map, 0, rss + ip, nullptr)) return -1;
    }
    return 0;
}
// Ybg = reconstruct_background(obj, frame_range) (Sources2D.m:1247-1356), one patch at a time: d_patch x (f1-f0+1)
extern "C" int cnmfe_reconstruct_background(cnmfe_ctx* c, int ip, int f0, int f1, const double* b0_map, const double* b0_new_map,
                                            double* Ybg) {
    if (!c || ip < 0 || ip >= c->npatch || !Ybg) { set_error("cnmfe_reconstruct_background: bad arguments"); return -1; }
    return bg_rows_pass(c, ip, f0, f1, b0_map, b0_new_map, 1, nullptr, Ybg);
}

// merged C_raw = sum_p aa_p C_raw,p / sum_p aa_p (update_temporal_parallel.m:269-280) as it ENTERED the final deconvTemporal,
// i.e. before deconvTemporal.m:84 subtracts the baseline; valid after cnmfe_update_temporal_finish[_part] (rows this rank
// finished).  K x T in the boundary layout (cnmfe_set_trace_major).  Lets a checker re-run deconvolveCa on the same input.
extern "C" int cnmfe_get_merged_craw(cnmfe_ctx* c, double* Craw_in) {
    if (!c || !Craw_in) { set_error("cnmfe_get_merged_craw: null"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    return download_KT(c, c->num, c->K, Craw_in);
}

// rows of the resident video (diagnostics / spot checks against the oracle): out[i][0..T) = Y(block pixel idx[i], :) of block
// ipatch as uint16, idx = r + c * nr_block (0-based, block coordinates)
extern "C" int cnmfe_debug_video_rows(cnmfe_ctx* c, int ip, int n, const int32_t* idx, uint16_t* out) {
    if (!c || ip < 0 || ip >= c->npatch || n < 0 || (n > 0 && (!idx || !out))) { set_error("cnmfe_debug_video_rows: bad arguments"); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(c->device));
    Patch& P = c->patches[ip];
    if (!P.owned || !P.uploaded) { set_error("cnmfe_debug_video_rows: block not resident"); return -1; }
    CNMFE_CUDA_OK(cudaStreamSynchronize(c->st));
    for (int i = 0; i < n; ++i) {
        if (idx[i] < 0 || idx[i] >= P.db) { set_error("cnmfe_debug_video_rows: pixel %d outside the block", idx[i]); return -1; }
        CNMFE_CUDA_OK(cudaMemcpy(out + (size_t)i * c->T, P.Yt + (size_t)idx[i] * c->Tpad, (size_t)c->T * 2, cudaMemcpyDeviceToHost));
    }
    return 0;
}
// Host planning under test on CPU: the block / patch-local view build_local() derives from a MATLAB CSC matrix.  No device
// work.  sel: 0 sum-in-block > 0, 1 sum-in-halo > 0, 2 any-in-patch; rows: 0 block pixels, 1 block index of patch pixels only,
// 2 patch pixels.  Output arrays must hold K (+1) / nrows + 1 / nnz entries; *n_local and *n_kept receive the used sizes.
extern "C" int cnmfe_debug_local_view(int d1, int d2, const int* patch_pos, const int* block_pos, int K, const int64_t* jc,
                                      const int64_t* ir, const double* pr, int sel, int rows, const int64_t* vjc,
                                      const int64_t* vir, const double* vpr, int* n_local, int* n_kept, int* ids, int* ptr,
                                      int* col, double* val, int64_t* entry_src, int* cptr, int* crow, double* cval, int* bbox) {
    if (!patch_pos || !block_pos || !jc || !n_local || !n_kept || sel < 0 || sel > 2 || rows < 0 || rows > 2) { set_error("cnmfe_debug_local_view: bad arguments"); return -1; }
    cnmfe_ctx c;
    c.d1 = d1; c.d2 = d2;
    Patch P;
    P.patch = {patch_pos[0] - 1, patch_pos[1] - 1, patch_pos[2] - 1, patch_pos[3] - 1};
    P.block = {block_pos[0] - 1, block_pos[1] - 1, block_pos[2] - 1, block_pos[3] - 1};
    P.nr = P.patch.r1 - P.patch.r0 + 1; P.nc = P.patch.c1 - P.patch.c0 + 1;
    P.nrb = P.block.r1 - P.block.r0 + 1; P.ncb = P.block.c1 - P.block.c0 + 1;
    P.dp = P.nr * P.nc; P.db = P.nrb * P.ncb;
    HostCsc M, V;
    M.set(K, jc, ir, pr);
    if (vjc) V.set(K, vjc, vir, vpr);
    LocalSparse L;
    build_local(&c, P, M, (SelMode)sel, (RowSpace)rows, vjc ? &V : nullptr, &L);
    *n_local = L.K(); *n_kept = (int)L.col.size();
    std::copy(L.ids.begin(), L.ids.end(), ids);
    std::copy(L.ptr.begin(), L.ptr.end(), ptr);
    std::copy(L.col.begin(), L.col.end(), col);
    std::copy(L.val.begin(), L.val.end(), val);
    std::copy(L.entry_src.begin(), L.entry_src.end(), entry_src);
    std::copy(L.cptr.begin(), L.cptr.end(), cptr);
    std::copy(L.crow.begin(), L.crow.end(), crow);
    std::copy(L.cval.begin(), L.cval.end(), cval);
    std::copy(L.bbox.begin(), L.bbox.end(), bbox);
    return 0;
}


#include "ctx_svd.inc"
#include "ctx_ssub.inc"
