// deconv.cu -- per-trace kernels: batch deconvolveCa / GetSn and the HALS_temporal Gauss-Seidel sweeps.
// One CTA per trace (persistent CTAs pull work items); all math double.  See oasis.cuh for reference citations.
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <algorithm>
#include "oasis_methods.cuh"
#include "internal.h"

namespace cnmfe {

unsigned long long g_launch_count = 0;
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

// ------------------------------------------------------------------------------------------------ arena

struct SlotLayout {
    size_t yb, c, s, pv, pw, h, hh, gp, scr, sv, sw, x, ybuf, pt, pl, st, sl, total;
};
__host__ __device__ inline SlotLayout slot_layout(int T) {
    SlotLayout L;
    size_t o = 0, t = (size_t)T;
    auto take = [&](size_t n) { size_t r = o; o += ((n * 8 + 127) / 128) * 128; return r; };
    L.yb = take(t); L.c = take(t); L.s = take(t); L.pv = take(t + 1); L.pw = take(t + 1);
    L.h = take(t + 1); L.hh = take(t + 1); L.gp = take(2 * t + 2); L.scr = take(trace_scratch_doubles(T));
    L.sv = take(t + 1); L.sw = take(t + 1); L.x = take(t); L.ybuf = take(t);
    L.pt = take((t + 1) / 2 + 1); L.pl = take((t + 1) / 2 + 1); L.st = take((t + 1) / 2 + 1);
    L.sl = take((t + 1) / 2 + 1);
    L.total = o;
    return L;
}
size_t trace_slot_bytes(int T) { return slot_layout(T).total; }

__device__ inline void carve_ws(char* base, size_t slot_bytes, int slot, int T, TraceWS* ws, double** x,
                                double** ybuf) {
    SlotLayout L = slot_layout(T);
    char* p = base + (size_t)slot * slot_bytes;
    ws->yb = (double*)(p + L.yb); ws->c = (double*)(p + L.c); ws->s = (double*)(p + L.s);
    ws->pv = (double*)(p + L.pv); ws->pw = (double*)(p + L.pw); ws->h = (double*)(p + L.h);
    ws->hh = (double*)(p + L.hh); ws->gp = (double*)(p + L.gp); ws->scr = (double*)(p + L.scr);
    ws->sv = (double*)(p + L.sv); ws->sw = (double*)(p + L.sw);
    ws->pt = (int*)(p + L.pt); ws->pl = (int*)(p + L.pl); ws->st = (int*)(p + L.st); ws->sl = (int*)(p + L.sl);
    *x = (double*)(p + L.x);
    *ybuf = (double*)(p + L.ybuf);
}

int default_trace_slots(int device) {
    static int cached_dev = -1, cached = 148 * 4;
    if (device != cached_dev) {
        int sms = 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) sms = 148;
        cached = sms * 4;
        cached_dev = device;
    }
    return cached;
}

int trace_arena_reserve(TraceArena* a, int T, int nslots) {
    size_t sb = trace_slot_bytes(T);
    if (!a->ticket) CNMFE_CUDA_OK(cudaMalloc((void**)&a->ticket, 64));
    if (a->base && a->slot_bytes >= sb && a->nslots >= nslots && a->T == T) return 0;
    if (a->base) { cudaFree(a->base); a->base = nullptr; }
    CNMFE_CUDA_OK(cudaMalloc((void**)&a->base, sb * (size_t)nslots));
    a->slot_bytes = sb; a->nslots = nslots; a->T = T;
    return 0;
}
void trace_arena_free(TraceArena* a) {
    if (a->base) cudaFree(a->base);
    if (a->ticket) cudaFree(a->ticket);
    a->base = nullptr; a->nslots = 0; a->slot_bytes = 0; a->ticket = nullptr;
}

// ------------------------------------------------------------------------------------------------ kernels
extern __shared__ __align__(16) unsigned char cnmfe_dyn_smem[];

// launch shape of the per-trace kernels: dynamic shared memory (see trace_smem_layout) and the resident-CTA cap that goes
// with it; *mode is the bit mask the kernel passes to trace_smem_bind
template <typename Kern>
static int trace_launch_shape(Kern kern, int T, int device, int want, bool want_stage, int* slots, size_t* smem, int* mode) {
    const TraceSmem L = trace_smem_layout(T, want_stage);
    *smem = L.total;
    *mode = (L.fft ? 1 : 0) | (L.stage ? 2 : 0);
    int s = default_trace_slots(device);
    if (*smem) {
        CNMFE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
        int per_sm = (int)((227 * 1024) / (*smem + 6 * 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm < 4) s = s / 4 * per_sm;
    }
    *slots = s > want ? want : s;
    return 0;
}
__global__ void __launch_bounds__(CNMFE_BLOCK)
deconv_batch_kernel(const double* __restrict__ Y, int T, int N, cnmfe_deconv_opts o, const double* __restrict__ sn_in,
                    const double* __restrict__ pars_in, int mode, double* __restrict__ c_out,
                    double* __restrict__ s_out, double* __restrict__ craw_out, double* __restrict__ outs,
                    char* arena, size_t slot_bytes, unsigned int* ticket, int fft_smem) {
    __shared__ BlockShared sh;
    __shared__ unsigned int s_item;
    TraceWS ws;
    double *x, *ybuf;
    carve_ws(arena, slot_bytes, blockIdx.x, T, &ws, &x, &ybuf);
    if (threadIdx.x == 0) { sh.prof = nullptr; trace_smem_bind(&sh, cnmfe_dyn_smem, T, fft_smem); }
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned int n = s_item;
        if (n >= (unsigned)N) break;
        const double* y = Y + (size_t)n * T;
        DeconvOut out;
        out.b = 0; out.g1 = 0; out.g2 = 0; out.smin = 0; out.lam = 0; out.sn = 0; out.npars = o.type;
        bool skip = false;
        if (mode == 1) {   // deconvTemporal.m:64-68: any NaN -> zeros
            double bad = 0.0;
            for (int i = threadIdx.x; i < T; i += blockDim.x) if (isnan(y[i])) bad = 1.0;
            bad = block_sum(bad, sh.red);
            if (bad > 0.0) {
                skip = true;
                for (int i = threadIdx.x; i < T; i += blockDim.x) {
                    if (c_out) c_out[(size_t)n * T + i] = 0.0;
                    if (s_out) s_out[(size_t)n * T + i] = 0.0;
                    if (craw_out) craw_out[(size_t)n * T + i] = 0.0;
                }
            }
        }
        if (!skip) {
            double sn = sn_in ? sn_in[n] : NAN;
            double p1 = pars_in ? pars_in[2 * n] : 0.0, p2 = pars_in ? pars_in[2 * n + 1] : 0.0;
            block_deconvolveCa(y, T, o, sn, p1, p2, 0, ybuf, ws, &sh, &out);
            bool allzero = false;
            if (mode == 1) {
                double a = 0.0;
                for (int i = threadIdx.x; i < T; i += blockDim.x) a += fabs(ws.c[i]);
                allzero = (block_sum(a, sh.red) == 0.0);
            }
            for (int i = threadIdx.x; i < T; i += blockDim.x) {
                double yi = y[i];
                if (c_out) c_out[(size_t)n * T + i] = allzero ? yi : ws.c[i];
                if (s_out) s_out[(size_t)n * T + i] = ws.s[i];
                if (craw_out) craw_out[(size_t)n * T + i] = yi - out.b;
            }
        }
        if (threadIdx.x == 0 && outs) {
            double* q = outs + (size_t)n * 6;
            q[0] = out.b; q[1] = out.g1; q[2] = out.g2; q[3] = out.smin; q[4] = out.lam; q[5] = out.sn;
        }
    }
}

__global__ void __launch_bounds__(CNMFE_BLOCK)
getsn_batch_kernel(const double* __restrict__ Y, int T, int N, double* __restrict__ sn, char* arena,
                   size_t slot_bytes, unsigned int* ticket, int fft_smem) {
    __shared__ BlockShared sh;
    __shared__ unsigned int s_item;
    TraceWS ws;
    double *x, *ybuf;
    carve_ws(arena, slot_bytes, blockIdx.x, T, &ws, &x, &ybuf);
    if (threadIdx.x == 0) { sh.prof = nullptr; trace_smem_bind(&sh, cnmfe_dyn_smem, T, fft_smem); }
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned int n = s_item;
        if (n >= (unsigned)N) break;
        double v = block_getsn(Y + (size_t)n * T, T, ws.scr, &sh);
        if (threadIdx.x == 0) sn[n] = v;
    }
}

struct HalsArgs {
    const double* U; const int* Vptr; const int* Vidx; const double* Vval; const double* aa;
    int K, T, maxIter, deconv_flag, n_update;
    cnmfe_deconv_opts o;
    double* C; double* C_raw; double* S; double* sn; double* pars;
    int* done; unsigned int* ticket; const int* order;
    char* arena; size_t slot_bytes;
    int fft_smem;               // dynamic shared memory mode (trace_smem_bind): bit 0 Welch FFT buffer, bit 1 update_g staging
    unsigned long long* prof;   // 16 counters or nullptr
    unsigned long long* prof_items;   // [maxIter * n_update][4] = start, deps ready, end (globaltimer ns), foopsi iterations
};

// One work item = (sweep, neuron).  Items are handed out in the reference's sequential order; an item waits until
// the neurons it overlaps (V(k,j) != 0) have reached the state the sequential loop would have seen
// (HALS_temporal.m:59-62: neuron k reads rows j<k of THIS sweep and rows j>k of the PREVIOUS sweep).
__global__ void __launch_bounds__(CNMFE_HALS_BLOCK, 1) hals_temporal_kernel(HalsArgs a) {
    __shared__ BlockShared sh;
    __shared__ unsigned int s_item;
    TraceWS ws;
    double *x, *ybuf;
    carve_ws(a.arena, a.slot_bytes, blockIdx.x, a.T, &ws, &x, &ybuf);
    if (threadIdx.x == 0) { trace_smem_bind(&sh, cnmfe_dyn_smem, a.T, a.fft_smem); sh.prof = a.prof; sh.t0 = clock64(); for (int i = 0; i < 32; ++i) sh.pc[i] = 0ull; }
    long long k_c0 = clock64(); unsigned long long k_g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_g0));
    const int T = a.T;
    const unsigned int total = (unsigned)a.maxIter * (unsigned)a.n_update;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const unsigned int item = s_item;
        if (item >= total) break;
        const int sweep = item / a.n_update, k = a.order[item % a.n_update];
        const int r0 = a.Vptr[k], r1 = a.Vptr[k + 1];
        CNMFE_PROF(&sh, 10);
        if (sh.prof) {
            if (threadIdx.x == 0) {
                sh.pc[14] += 1ull;
                unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
                a.prof_items[4 * (size_t)item] = g; a.prof_items[4 * (size_t)item + 3] = sh.pc[15];
            }
            __syncwarp();
        }
        if (threadIdx.x == 0) {
            for (int e = r0; e < r1; ++e) {
                int j = a.Vidx[e];
                int need = (j < k) ? sweep + 1 : sweep;
                volatile int* dj = a.done + j;
                while (*dj < need) __nanosleep(200);
            }
            __threadfence();
        }
        __syncthreads();
        CNMFE_PROF(&sh, 0);
        if (sh.prof) {
            if (threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); a.prof_items[4 * (size_t)item + 1] = g; }
            __syncwarp();
        }
        const double aak = a.aa[k];
        // ck_raw = C(k,:) + (U(k,:) - V(k,:)*C)/aa(k)   (HALS_temporal.m:62)
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            double acc = 0.0;
            for (int e = r0; e < r1; ++e) acc += a.Vval[e] * __ldcg(a.C + (size_t)a.Vidx[e] * T + t);
            x[t] = __ldcg(a.C + (size_t)k * T + t) + (a.U[(size_t)k * T + t] - acc) / aak;
        }
        __syncthreads();
        CNMFE_PROF(&sh, 1);
        const bool last = (sweep == a.maxIter - 1);
        if (!a.deconv_flag) {
            double mn = INFINITY;
            for (int t = threadIdx.x; t < T; t += blockDim.x) mn = fmin(mn, x[t]);
            mn = -block_max(-mn, sh.red);
            for (int t = threadIdx.x; t < T; t += blockDim.x) {
                double v = x[t] - mn;
                a.C[(size_t)k * T + t] = v;
                a.C_raw[(size_t)k * T + t] = v;
            }
        } else {
            double med = block_median(x, T, &sh);
            double b = block_mean_below(x, T, med, &sh);       // HALS_temporal.m:78
            CNMFE_PROF(&sh, 2);
            double sn_psd = block_getsn(x, T, ws.scr, &sh);     // :79
            CNMFE_PROF(&sh, 3);
            for (int t = threadIdx.x; t < T; t += blockDim.x) x[t] -= b;
            __syncthreads();
            DeconvOut out;
            out.b = 0; out.g1 = 0; out.g2 = 0; out.smin = 0; out.lam = 0; out.sn = 0; out.npars = a.o.type;
            block_deconvolveCa(x, T, a.o, sn_psd, a.pars[2 * k], a.pars[2 * k + 1], 20, ybuf, ws, &sh, &out);  // :92
            double sa = 0.0;
            for (int t = threadIdx.x; t < T; t += blockDim.x) sa += fabs(ws.c[t]);
            const bool allzero = (block_sum(sa, sh.red) == 0.0);
            for (int t = threadIdx.x; t < T; t += blockDim.x) {
                double raw = x[t] - out.b;
                a.C[(size_t)k * T + t] = allzero ? raw : ws.c[t];
                if (last) {
                    a.S[(size_t)k * T + t] = ws.s[t];
                    a.C_raw[(size_t)k * T + t] = raw;
                }
            }
            if (threadIdx.x == 0) {
                a.sn[k] = sn_psd;
                a.pars[2 * k] = out.g1;
                a.pars[2 * k + 1] = out.g2;
            }
        }
        __threadfence();
        __syncthreads();
        CNMFE_PROF(&sh, 10);
        if (sh.prof) {
            if (threadIdx.x == 0) {
                unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
                a.prof_items[4 * (size_t)item + 2] = g; a.prof_items[4 * (size_t)item + 3] = sh.pc[15] - a.prof_items[4 * (size_t)item + 3];
            }
            __syncwarp();
        }
        if (threadIdx.x == 0) atomicAdd(a.done + k, 1);
    }
    if (threadIdx.x == 0 && a.prof) {
        for (int i = 0; i < 12; ++i) atomicAdd(&a.prof[i], sh.pc[i]);
        for (int i = 14; i < 32; ++i) atomicAdd(&a.prof[i], sh.pc[i]);
    }
    if (threadIdx.x == 0 && blockIdx.x == 0 && a.prof) {
        unsigned long long k_g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_g1));
        a.prof[12] = (unsigned long long)(clock64() - k_c0);
        a.prof[13] = k_g1 - k_g0;
    }
}

__global__ void hals_order_kernel(const double* aa, int K, int maxIter, int* order, int* done, unsigned int* ticket,
                                  int* n_update) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int m = 0;
        for (int k = 0; k < K; ++k) {
            if (aa[k] > 0.0) { order[m++] = k; done[k] = 0; }
            else done[k] = maxIter;   // never updated (HALS_temporal.m:50-51): treated as finished
        }
        *n_update = m;
        *ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------------ host wrappers

int deconv_batch_dev(const double* Y, int T, int N, const cnmfe_deconv_opts& o, const double* sn_in,
                     const double* pars_in, int mode, double* c, double* s, double* craw_out, double* outs,
                     TraceArena* arena, cudaStream_t st) {
    if (N <= 0) return 0;
    if (T < 32) { set_error("deconvolve: T=%d too short (need >= 32 frames)", T); return -1; }
    if (o.type != 1 && o.type != 2) { set_error("deconvolve: type must be 1 (ar1) or 2 (ar2)"); return -1; }
    if (o.type == 2 && o.method == 1) {
        set_error("deconvolve: ar2 + 'constrained' is the legacy constrained_foopsi (CVX/LARS) path: not built");
        return -1;
    }
    if (o.type == 2 && o.method == 2 && o.optimize_pars) {
        set_error("deconvolve: thresholded_oasisAR2 with optimize_pars (update_g of the AR(2) kernel: a dense spike-amplitude solve per fminbnd evaluation) is not built");
        return -1;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    int slots, smode; size_t smem;
    if (trace_launch_shape(deconv_batch_kernel, T, dev, N, o.optimize_pars != 0, &slots, &smem, &smode)) return -1;
    if (trace_arena_reserve(arena, T, slots > arena->nslots ? slots : arena->nslots)) return -1;
    CNMFE_CUDA_OK(cudaMemsetAsync(arena->ticket, 0, 4, st));
    LAUNCH(deconv_batch_kernel, slots, CNMFE_BLOCK, smem, st, Y, T, N, o, sn_in, pars_in, mode, c, s, craw_out, outs,
           arena->base, arena->slot_bytes, arena->ticket, smode);
    CNMFE_CUDA_OK(cudaGetLastError());
    return 0;
}

int getsn_batch_dev(const double* Y, int T, int N, double* sn, TraceArena* arena, cudaStream_t st) {
    if (N <= 0) return 0;
    if (T < 32) { set_error("GetSn: T=%d too short", T); return -1; }
    int dev = 0;
    cudaGetDevice(&dev);
    int slots, smode; size_t smem;
    if (trace_launch_shape(getsn_batch_kernel, T, dev, N, false, &slots, &smem, &smode)) return -1;
    if (trace_arena_reserve(arena, T, slots > arena->nslots ? slots : arena->nslots)) return -1;
    CNMFE_CUDA_OK(cudaMemsetAsync(arena->ticket, 0, 4, st));
    LAUNCH(getsn_batch_kernel, slots, CNMFE_BLOCK, smem, st, Y, T, N, sn, arena->base, arena->slot_bytes, arena->ticket, smode);
    CNMFE_CUDA_OK(cudaGetLastError());
    return 0;
}

int hals_temporal_dev(const double* U, const int* Vptr, const int* Vidx, const double* Vval, const double* aa,
                      int K, int T, int maxIter, int deconv_flag, const cnmfe_deconv_opts& o, double* C,
                      double* C_raw, double* S, double* sn, double* pars, int* done, unsigned int* ticket,
                      int* order_scratch, TraceArena* arena, cudaStream_t st) {
    if (K <= 0 || maxIter <= 0) return 0;
    if (T < 32) { set_error("HALS_temporal: T=%d too short", T); return -1; }
    int dev = 0;
    cudaGetDevice(&dev);
    int slots, smode; size_t smem;
    if (trace_launch_shape(hals_temporal_kernel, T, dev, K, deconv_flag && o.optimize_pars, &slots, &smem, &smode)) return -1;
    if (trace_arena_reserve(arena, T, slots > arena->nslots ? slots : arena->nslots)) return -1;
    // order_scratch: K ints + 1 (n_update at [K])
    LAUNCH(hals_order_kernel, 1, 32, 0, st, aa, K, maxIter, order_scratch, done, ticket, order_scratch + K);
    int n_update = 0;
    CNMFE_CUDA_OK(cudaMemcpyAsync(&n_update, order_scratch + K, sizeof(int), cudaMemcpyDeviceToHost, st));
    CNMFE_CUDA_OK(cudaStreamSynchronize(st));
    if (n_update == 0) return 0;
    HalsArgs a;
    a.U = U; a.Vptr = Vptr; a.Vidx = Vidx; a.Vval = Vval; a.aa = aa;
    a.K = K; a.T = T; a.maxIter = maxIter; a.deconv_flag = deconv_flag; a.n_update = n_update;
    a.o = o;
    a.C = C; a.C_raw = C_raw; a.S = S; a.sn = sn; a.pars = pars;
    a.done = done; a.ticket = ticket; a.order = order_scratch;
    a.arena = arena->base; a.slot_bytes = arena->slot_bytes;
    a.fft_smem = smode;
    a.prof = nullptr; a.prof_items = nullptr;
    static const bool profile = getenv("CNMFE_HALS_PROFILE") != nullptr;   // diagnostics: per-phase cycles of thread 0
    if (profile) {
        CNMFE_CUDA_OK(cudaMalloc((void**)&a.prof, 32 * 8));
        CNMFE_CUDA_OK(cudaMemsetAsync(a.prof, 0, 32 * 8, st));
        CNMFE_CUDA_OK(cudaMalloc((void**)&a.prof_items, (size_t)maxIter * n_update * 32));
        CNMFE_CUDA_OK(cudaMemsetAsync(a.prof_items, 0, (size_t)maxIter * n_update * 32, st));
    }
    int hals_threads = CNMFE_HALS_BLOCK;
    if (const char* e = getenv("CNMFE_HALS_THREADS")) { const int v = atoi(e); if (v >= 64 && v <= CNMFE_HALS_BLOCK && v % 32 == 0) hals_threads = v; }   // A/B knob
    LAUNCH(hals_temporal_kernel, slots, hals_threads, smem, st, a);
    CNMFE_CUDA_OK(cudaGetLastError());
    if (profile) {
        unsigned long long h[32];
        CNMFE_CUDA_OK(cudaStreamSynchronize(st));
        CNMFE_CUDA_OK(cudaMemcpy(h, a.prof, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(a.prof);
        static const char* nm[16] = {"wait deps", "gemv", "median+mean", "getsn", "quantile", "cold scan", "solution", "fminbnd",
                                     "rebuild pools", "warm run", "other", "pow table", "", "", "items", "foopsi iters"};
        fprintf(stderr, "[cnmfe hals profile] K=%d T=%d slots=%d:", K, T, slots);
        for (int i = 0; i < 12; ++i) fprintf(stderr, " %s=%.0fk", nm[i], h[14] ? (double)h[i] / (double)h[14] / 1e3 : 0.0);
        fprintf(stderr, " cycles/item; items=%llu iters/item=%.2f block0 clock64=%llu globaltimer_ns=%llu\n", h[14], h[14] ? (double)h[15] / (double)h[14] : 0.0, h[12], h[13]);
        if (h[20]) fprintf(stderr, "[cnmfe hals profile] rss_g: %.1f evals/item, per eval: h+hh tables %.0f, partial dots along the trace %.0f, per-pool terms %.0f (+%.0f +%.0f), sum %.0f cycles; mean maxl %.0f, mean pools %.0f, mean long pools %.2f\n",
                           (double)h[20] / h[14], (double)h[16] / h[20], (double)h[17] / h[20], (double)h[24] / h[20], (double)h[25] / h[20], (double)h[18] / h[20], (double)h[19] / h[20],
                           (double)h[21] / h[20], (double)h[22] / h[20], (double)h[26] / h[20]);
        // critical path: walk back from the item that finished last through the dependency that released it
        {
            const size_t total = (size_t)maxIter * n_update;
            std::vector<unsigned long long> it(total * 4);
            CNMFE_CUDA_OK(cudaMemcpy(it.data(), a.prof_items, total * 32, cudaMemcpyDeviceToHost));
            cudaFree(a.prof_items);
            std::vector<int> vptr(K + 1), order(n_update);
            CNMFE_CUDA_OK(cudaMemcpy(vptr.data(), Vptr, (K + 1) * 4, cudaMemcpyDeviceToHost));
            std::vector<int> vidx(vptr[K]);
            CNMFE_CUDA_OK(cudaMemcpy(vidx.data(), Vidx, vidx.size() * 4, cudaMemcpyDeviceToHost));
            CNMFE_CUDA_OK(cudaMemcpy(order.data(), order_scratch, n_update * 4, cudaMemcpyDeviceToHost));
            std::vector<int> pos(K, -1);
            for (int i = 0; i < n_update; ++i) pos[order[i]] = i;
            unsigned long long t0 = ~0ull, t1 = 0; size_t last = 0;
            double sum_exec = 0, max_exec = 0;
            for (size_t i = 0; i < total; ++i) {
                t0 = std::min(t0, it[4 * i]);
                if (it[4 * i + 2] > t1) { t1 = it[4 * i + 2]; last = i; }
                double ex = (double)(it[4 * i + 2] - it[4 * i + 1]);
                sum_exec += ex; max_exec = std::max(max_exec, ex);
            }
            fprintf(stderr, "[cnmfe hals profile] span %.3f ms, mean exec %.3f ms, max exec %.3f ms; critical chain (sweep,k: wait ms, exec ms, iters):", (t1 - t0) / 1e6, sum_exec / total / 1e6, max_exec / 1e6);
            size_t cur = last;
            for (int hop = 0; hop < 64; ++hop) {
                const int sweep = (int)(cur / n_update), k = order[cur % n_update];
                fprintf(stderr, " (%d,%d: %.2f %.2f %llu)", sweep, k, (it[4 * cur + 1] - it[4 * cur]) / 1e6, (it[4 * cur + 2] - it[4 * cur + 1]) / 1e6, it[4 * cur + 3]);
                // the dependency that finished last
                long long best = -1; unsigned long long bt = 0;
                for (int e = vptr[k]; e < vptr[k + 1]; ++e) {
                    const int j = vidx[e];
                    if (pos[j] < 0) continue;
                    const int need = (pos[j] < pos[k]) ? sweep + 1 : sweep;   // completions of j required
                    if (need <= 0) continue;
                    const size_t dep = (size_t)(need - 1) * n_update + pos[j];
                    if (it[4 * dep + 2] >= bt) { bt = it[4 * dep + 2]; best = (long long)dep; }
                }
                if (best < 0) break;
                cur = (size_t)best;
            }
            fprintf(stderr, "\n");
        }
    }
    return 0;
}

}  // namespace cnmfe

// ================================================================================================== C ABI
using namespace cnmfe;

extern "C" const char* cnmfe_last_error(void) { return cnmfe::get_error(); }
extern "C" unsigned long long cnmfe_launch_count(void) { return cnmfe::g_launch_count; }

extern "C" void cnmfe_deconv_defaults(cnmfe_deconv_opts* o) {
    memset(o, 0, sizeof(*o));
    o->type = 1; o->method = 1; o->optimize_b = 0; o->optimize_pars = 0; o->maxIter = 10; o->has_tau_range = 0;
    o->smin = 0.0; o->lambda = 0.0; o->b = 0.0; o->max_tau = 100.0; o->thresh_factor = 1.0; o->p_noise = 0.9999;
}

extern "C" void cnmfe_options_defaults(cnmfe_options* o) {
    memset(o, 0, sizeof(*o));
    o->spatial_algorithm = 0; o->maxIter_temporal = 5; o->deconv_flag = 1; o->bg_acceleration = 1;
    o->replicate_spatial_aprev_quirk = 1; o->use_tensor_gram = 1;
    o->background_model = 0; o->nb = 1; o->bg_ssub = 1; o->thresh_outlier = NAN;
    cnmfe_deconv_defaults(&o->deconv);
}

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { return cudaMalloc(&p, n ? n : 8) == cudaSuccess ? 0 : -1; }
    template <class T> T* as() { return (T*)p; }
};
int use_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        set_error("no CUDA device: libcnmfe_b200 has no CPU fallback");
        return -1;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range (0..%d)", device, n - 1); return -1; }
    CNMFE_CUDA_OK(cudaSetDevice(device));
    return 0;
}
}  // namespace

extern "C" int cnmfe_deconvolve(const double* Y, int T, int N, const cnmfe_deconv_opts* opts, const double* sn_in,
                                const double* pars_in, double* c, double* s, double* b, double* pars, double* sn,
                                double* smin, double* lam, int device) {
    if (!Y || !opts || T <= 0 || N < 0) { set_error("cnmfe_deconvolve: bad arguments"); return -1; }
    if (N == 0) return 0;
    if (use_device(device)) return -1;
    size_t nt = (size_t)T * N;
    DevBuf dY, dc, ds, dsn, dp, douts;
    TraceArena arena;
    int rc = -1;
    do {
        if (dY.alloc(nt * 8) || dc.alloc(nt * 8) || ds.alloc(nt * 8) || douts.alloc((size_t)N * 6 * 8)) {
            set_error("cnmfe_deconvolve: device allocation failed");
            break;
        }
        if (cudaMemcpy(dY.p, Y, nt * 8, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); break; }
        if (sn_in) { if (dsn.alloc((size_t)N * 8)) break; cudaMemcpy(dsn.p, sn_in, (size_t)N * 8, cudaMemcpyHostToDevice); }
        if (pars_in) { if (dp.alloc((size_t)N * 16)) break; cudaMemcpy(dp.p, pars_in, (size_t)N * 16, cudaMemcpyHostToDevice); }
        if (deconv_batch_dev(dY.as<double>(), T, N, *opts, sn_in ? dsn.as<double>() : nullptr,
                             pars_in ? dp.as<double>() : nullptr, 0, dc.as<double>(), ds.as<double>(), nullptr,
                             douts.as<double>(), &arena, 0)) break;
        if (cudaDeviceSynchronize() != cudaSuccess) {
            set_error("cnmfe_deconvolve: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (c) cudaMemcpy(c, dc.p, nt * 8, cudaMemcpyDeviceToHost);
        if (s) cudaMemcpy(s, ds.p, nt * 8, cudaMemcpyDeviceToHost);
        std::vector<double> outs((size_t)N * 6);
        cudaMemcpy(outs.data(), douts.p, (size_t)N * 48, cudaMemcpyDeviceToHost);
        for (int n = 0; n < N; ++n) {
            if (b) b[n] = outs[6 * n + 0];
            if (pars) { pars[2 * n] = outs[6 * n + 1]; pars[2 * n + 1] = outs[6 * n + 2]; }
            if (smin) smin[n] = outs[6 * n + 3];
            if (lam) lam[n] = outs[6 * n + 4];
            if (sn) sn[n] = outs[6 * n + 5];
        }
        rc = 0;
    } while (0);
    trace_arena_free(&arena);
    return rc;
}

// deconvolveCa on traces that already live on the device (BASELINE configs[4]: 5000 x 100000 AR2 is 4 GB in, 8 GB out -- the
// host entry point above would spend its time on the PCIe copies).  All pointers are device pointers on `device`; Y, c, s are
// [N][T] trace-contiguous; sn_in / pars_in ([N][2]) / outs ([N][6] = b, g1, g2, smin, lam, sn) may be NULL.  Asynchronous on the
// legacy default stream; the per-device workspace is cached, so the call is not re-entrant per device.
extern "C" int cnmfe_deconvolve_dev(const double* Y_dev, int T, int N, const cnmfe_deconv_opts* opts, const double* sn_dev,
                                    const double* pars_dev, double* c_dev, double* s_dev, double* outs_dev, int device) {
    if (!Y_dev || !opts || T <= 0 || N < 0 || device < 0 || device >= 64) { set_error("cnmfe_deconvolve_dev: bad arguments"); return -1; }
    if (N == 0) return 0;
    if (use_device(device)) return -1;
    static TraceArena arenas[64];
    if (deconv_batch_dev(Y_dev, T, N, *opts, sn_dev, pars_dev, 0, c_dev, s_dev, nullptr, outs_dev, &arenas[device], 0)) return -1;
    CNMFE_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int cnmfe_get_sn(const double* Y, int T, int N, double* sn, int device) {
    if (!Y || !sn || T <= 0 || N < 0) { set_error("cnmfe_get_sn: bad arguments"); return -1; }
    if (N == 0) return 0;
    if (use_device(device)) return -1;
    DevBuf dY, dsn;
    TraceArena arena;
    int rc = -1;
    do {
        if (dY.alloc((size_t)T * N * 8) || dsn.alloc((size_t)N * 8)) { set_error("device allocation failed"); break; }
        cudaMemcpy(dY.p, Y, (size_t)T * N * 8, cudaMemcpyHostToDevice);
        if (getsn_batch_dev(dY.as<double>(), T, N, dsn.as<double>(), &arena, 0)) break;
        if (cudaDeviceSynchronize() != cudaSuccess) {
            set_error("cnmfe_get_sn: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        cudaMemcpy(sn, dsn.p, (size_t)N * 8, cudaMemcpyDeviceToHost);
        rc = 0;
    } while (0);
    trace_arena_free(&arena);
    return rc;
}

extern "C" int cnmfe_hals_temporal_uv(const double* U, const double* V, int K, int T, double* C, int maxIter,
                                      const cnmfe_deconv_opts* deconv, double* C_raw, double* S, double* sn,
                                      double* kernel_pars, int device) {
    if (!U || !V || !C || K < 0 || T <= 0) { set_error("cnmfe_hals_temporal_uv: bad arguments"); return -1; }
    if (K == 0) return 0;
    if (use_device(device)) return -1;
    // MATLAB K x T column-major -> [k][t]
    size_t kt = (size_t)K * T;
    std::vector<double> Ut(kt), Ct(kt), aa(K);
    for (int k = 0; k < K; ++k)
        for (int t = 0; t < T; ++t) { Ut[(size_t)k * T + t] = U[(size_t)t * K + k]; Ct[(size_t)k * T + t] = C[(size_t)t * K + k]; }
    std::vector<int> ptr(K + 1, 0), idx;
    std::vector<double> val;
    for (int k = 0; k < K; ++k) {
        for (int j = 0; j < K; ++j) {
            double v = V[(size_t)j * K + k];   // V(k,j)
            if (v != 0.0 || j == k) { idx.push_back(j); val.push_back(v); }
        }
        ptr[k + 1] = (int)idx.size();
        aa[k] = V[(size_t)k * K + k];
    }
    DevBuf dU, dC, dCr, dS, dsn, dp, dptr, didx, dval, daa, ddone, dtick, dord;
    TraceArena arena;
    int rc = -1;
    do {
        if (dU.alloc(kt * 8) || dC.alloc(kt * 8) || dCr.alloc(kt * 8) || dS.alloc(kt * 8) || dsn.alloc(K * 8) ||
            dp.alloc(K * 16) || dptr.alloc((K + 1) * 4) || didx.alloc(idx.size() * 4) || dval.alloc(val.size() * 8) ||
            daa.alloc(K * 8) || ddone.alloc(K * 4) || dtick.alloc(64) || dord.alloc((K + 1) * 4)) {
            set_error("device allocation failed");
            break;
        }
        cudaMemcpy(dU.p, Ut.data(), kt * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dC.p, Ct.data(), kt * 8, cudaMemcpyHostToDevice);
        cudaMemset(dCr.p, 0, kt * 8); cudaMemset(dS.p, 0, kt * 8); cudaMemset(dsn.p, 0, K * 8);
        cudaMemset(dp.p, 0, K * 16);
        cudaMemcpy(dptr.p, ptr.data(), (K + 1) * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(didx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dval.p, val.data(), val.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(daa.p, aa.data(), K * 8, cudaMemcpyHostToDevice);
        cnmfe_deconv_opts o;
        cnmfe_deconv_defaults(&o);
        if (deconv) o = *deconv;
        if (hals_temporal_dev(dU.as<double>(), dptr.as<int>(), didx.as<int>(), dval.as<double>(), daa.as<double>(),
                              K, T, maxIter, deconv ? 1 : 0, o, dC.as<double>(), dCr.as<double>(), dS.as<double>(),
                              dsn.as<double>(), dp.as<double>(), ddone.as<int>(), dtick.as<unsigned int>(),
                              dord.as<int>(), &arena, 0)) break;
        if (cudaDeviceSynchronize() != cudaSuccess) {
            set_error("cnmfe_hals_temporal_uv: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        std::vector<double> tmp(kt);
        auto back = [&](void* d, double* h) {
            if (!h) return;
            cudaMemcpy(tmp.data(), d, kt * 8, cudaMemcpyDeviceToHost);
            for (int k = 0; k < K; ++k)
                for (int t = 0; t < T; ++t) h[(size_t)t * K + k] = tmp[(size_t)k * T + t];
        };
        back(dC.p, C); back(dCr.p, C_raw); back(dS.p, S);
        if (sn) cudaMemcpy(sn, dsn.p, K * 8, cudaMemcpyDeviceToHost);
        if (kernel_pars) cudaMemcpy(kernel_pars, dp.p, K * 16, cudaMemcpyDeviceToHost);
        rc = 0;
    } while (0);
    trace_arena_free(&arena);
    return rc;
}
