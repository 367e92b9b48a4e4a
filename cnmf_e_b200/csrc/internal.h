// internal.h -- device-pointer level interfaces between the translation units of libcnmfe_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "../../include/cnmfe_b200.h"

namespace cnmfe {

// ---- trace arena: per-CTA workspaces for the OASIS kernels ---------------------------------------------------
struct TraceArena {
    char* base = nullptr;
    size_t slot_bytes = 0;
    int nslots = 0;
    int T = 0;
    unsigned int* ticket = nullptr;   // work-item counter of the batch kernels launched on this arena (one arena = one ctx or one call)
};
size_t trace_slot_bytes(int T);
int trace_arena_reserve(TraceArena* a, int T, int nslots);   // (re)allocates when too small
void trace_arena_free(TraceArena* a);
int default_trace_slots(int device);

// Batch deconvolution, everything on the device.  Y: [N][T] (trace contiguous).  sn_in/pars_in may be null.
// mode 0: plain deconvolveCa.  mode 1: deconvTemporal semantics (deconvTemporal.m:62-84): NaN guard, sn = GetSn,
//         c = y when sum|c| == 0, and y_out = y - b written to craw_out.
// outs: [N][6] = (b, g1, g2, smin, lam, sn).
int deconv_batch_dev(const double* Y, int T, int N, const cnmfe_deconv_opts& o, const double* sn_in,
                     const double* pars_in, int mode, double* c, double* s, double* craw_out, double* outs,
                     TraceArena* arena, cudaStream_t st);

int getsn_batch_dev(const double* Y, int T, int N, double* sn, TraceArena* arena, cudaStream_t st);

// HALS_temporal sweeps (utilities/HALS_temporal.m:59-107) on projections.  U: [K][T]; V in CSR (diagonal included);
// C: [K][T] in/out; C_raw, S: [K][T]; sn: K; pars: [K][2] (in/out, zeros = estimate); done/ticket: scratch.
int hals_temporal_dev(const double* U, const int* Vptr, const int* Vidx, const double* Vval, const double* aa,
                      int K, int T, int maxIter, int deconv_flag, const cnmfe_deconv_opts& o, double* C,
                      double* C_raw, double* S, double* sn, double* pars, int* done, unsigned int* ticket,
                      int* order_scratch, TraceArena* arena, cudaStream_t st);

// ---- video / projections -------------------------------------------------------------------------------------
// Transpose a frame-major block (T frames of d pixels, u8 or u16) into pixel-major u16 rows Yt[d][Tpad]
// (pad = 0) and byte planes hi/lo [d][Tpad]; also per-pixel sums.
int transpose_block_dev(const void* Y_frames, int dtype, int d, int T, int Tpad, uint16_t* Yt, uint8_t* hi,
                        uint8_t* lo, double* Ysum, cudaStream_t st);

}  // namespace cnmfe
