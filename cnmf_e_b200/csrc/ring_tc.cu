// ring_tc.cu -- banded second moments of the resident video on the 5th-gen tensor cores (sm_100a only).
//
//   S2[q][id(D)] = sum_t Y[q,t] * Y[q+D,t]     for every block pixel q and canonical displacement D (|D|_inf <= 2*rr)
//
// The uint16 video is resident as two K-major byte planes hi/lo [pixel][Tpad]; Y = 256*hi + lo, so
//   Y_p*Y_q = 65536*hh + 256*(hl + lh) + ll   with hh = sum hi_p*hi_q, ... each an EXACT int32 sum
// computed by tcgen05.mma.kind::i8 (u8 x u8 -> s32 accumulators in TMEM).  Because every factor is an integer the
// result is bit-exact and identical to the SIMT kernel (kernels_ring.cuh) -- tests/test_gpu_ring_tc.py checks equality.
//
// Work item = (pixel tile of MR x MC = 32 x 4 = 128 pixels = MMA M, one neighbour column of NB = 112 rows = MMA N).
// Per item the K loop streams all frames: TMA (3-D boxes {128 B of t, rows, cols}, SWIZZLE_128B) -> 3-stage smem ring
// -> 16 MMAs per stage (4 K-steps of 32 B x 4 byte-plane products) into three accumulators (hh, hl+lh, ll), then the
// epilogue warps read TMEM (tcgen05.ld 32x32b), combine in int64 and write the (pixel, displacement) run.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue (one TMEM lane quadrant each).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "common.cuh"
#include "internal.h"

namespace cnmfe {

// same displacement indexing as kernels_ring.cuh
__host__ __device__ inline int tc_num_disp(int rr) { return 2 * rr * (4 * rr + 1) + (2 * rr + 1); }
__host__ __device__ inline int tc_disp_id(int dr, int dc, int rr) {
    return dc == 0 ? dr : (2 * rr + 1) + (dc - 1) * (4 * rr + 1) + (dr + 2 * rr);
}

namespace tc {
constexpr int MR = 32, MC = 4, NB = 112, BSHIFT = 40;     // tile rows/cols, B rows, B starts BSHIFT rows above the tile
constexpr int KSTAGE = 128;                               // bytes of t per stage (= swizzle span)
constexpr int STAGES = 3;
constexpr int A_BYTES = MR * MC * KSTAGE;                 // 16384
constexpr int B_BYTES = NB * KSTAGE;                      // 14336
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;    // 61440
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 512;
constexpr int COL_HH = 0, COL_MID = 128, COL_LL = 256;
constexpr int THREADS = 192;
constexpr int MAX_K_BYTES = 16384;                        // frames per pass so that hl+lh < 2^31
}  // namespace tc

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
}

struct TcParams {
    int nrb, ncb, rr, ND;
    int ntr, ntc, nbc;          // tiles along r, along c, B columns per tile
    int kb0, kb1;               // K-stage range of this pass
    int accumulate;             // epilogue: S2 += (second and later passes)
    long long nitems;
    double* S2;
};

__global__ void __launch_bounds__(tc::THREADS, 1)
ring_s2_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const TcParams P) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned tile area (SWIZZLE_128B atoms), barriers after it
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;   // [1]
    uint64_t* tmem_empty = tmem_full + 1;      // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
        mbar_init(smem_u32(tmem_full), 1);
        mbar_init(smem_u32(tmem_empty), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int nkb = P.kb1 - P.kb0;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long item = blockIdx.x; item < P.nitems; item += gridDim.x) {
                const int j = (int)(item % P.nbc);
                const long long tile = item / P.nbc;
                const int tr = (int)(tile % P.ntr), tcx = (int)(tile / P.ntr);
                const int r0 = tr * MR, c0 = tcx * MC, cB = c0 + j;
                const int rB0 = max(0, r0 - BSHIFT);
                for (int kb = P.kb0; kb < P.kb1; ++kb) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    uint8_t* st = tiles + stage * STAGE_BYTES;
                    mbar_expect_tx(fb, STAGE_BYTES);
                    tma_load_3d(smem_u32(st), &tmA_hi, fb, kb * KSTAGE, r0, c0);
                    tma_load_3d(smem_u32(st + A_BYTES), &tmA_lo, fb, kb * KSTAGE, r0, c0);
                    tma_load_3d(smem_u32(st + 2 * A_BYTES), &tmB_hi, fb, kb * KSTAGE, rB0, cB);
                    tma_load_3d(smem_u32(st + 2 * A_BYTES + B_BYTES), &tmB_lo, fb, kb * KSTAGE, rB0, cB);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one lane)
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) @4, a/b format U8 (0), K-major,
            // N>>3 @17, M>>4 @24
            const uint32_t idesc = (2u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long item = blockIdx.x; item < P.nitems; item += gridDim.x) {
                mbar_wait(smem_u32(tmem_empty), acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < KSTAGE / 32; ++ks) {
                        const uint64_t dAh = make_desc_sw128(sa + ks * 32);
                        const uint64_t dAl = make_desc_sw128(sa + A_BYTES + ks * 32);
                        const uint64_t dBh = make_desc_sw128(sa + 2 * A_BYTES + ks * 32);
                        const uint64_t dBl = make_desc_sw128(sa + 2 * A_BYTES + B_BYTES + ks * 32);
                        const uint32_t acc = (kb | ks) ? 1u : 0u;
                        mma_i8(tmem_base + COL_HH, dAh, dBh, idesc, acc);
                        mma_i8(tmem_base + COL_MID, dAh, dBl, idesc, acc);
                        mma_i8(tmem_base + COL_MID, dAl, dBh, idesc, 1u);
                        mma_i8(tmem_base + COL_LL, dAl, dBl, idesc, acc);
                    }
                    umma_commit(smem_u32(empty_bar + stage));     // frees the smem stage when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(smem_u32(tmem_full));                 // accumulators complete
                acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: warps 2..5 -> TMEM lane quadrant warp%4
        const int quad = warp & 3;
        uint32_t acc_phase = 0;
        for (long long item = blockIdx.x; item < P.nitems; item += gridDim.x) {
            const int j = (int)(item % P.nbc);
            const long long tile = item / P.nbc;
            const int tr = (int)(tile % P.ntr), tcx = (int)(tile / P.ntr);
            const int r0 = tr * MR, c0 = tcx * MC, cB = c0 + j;
            const int rB0 = max(0, r0 - BSHIFT);
            const int rp = r0 + lane, cp = c0 + quad;          // this thread's pixel (TMEM lane = lane + 32*quad)
            const int dc = cB - cp;
            const bool pix_ok = (rp < P.nrb) && (cp < P.ncb) && (cB < P.ncb) && (dc >= 0) && (dc <= 2 * P.rr);
            mbar_wait(smem_u32(tmem_full), acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
            double* out = P.S2 + ((size_t)cp * P.nrb + rp) * (size_t)P.ND;
#pragma unroll 1
            for (int cc = 0; cc < NB; cc += 16) {
                uint32_t hh[16], mid[16], ll[16];
                tmem_ld16(taddr + COL_HH + cc, hh);
                tmem_ld16(taddr + COL_MID + cc, mid);
                tmem_ld16(taddr + COL_LL + cc, ll);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (pix_ok) {
#pragma unroll
                    for (int x = 0; x < 16; ++x) {
                        const int rn = rB0 + cc + x;
                        const int dr = rn - rp;
                        if (rn < P.nrb && dr >= -2 * P.rr && dr <= 2 * P.rr && (dc > 0 || dr >= 0)) {
                            long long v = ((long long)(int)hh[x] << 16) + ((long long)(int)mid[x] << 8) + (long long)(int)ll[x];
                            double* o = out + tc_disp_id(dr, dc, P.rr);
                            if (P.accumulate) *o += (double)v; else *o = (double)v;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tmem_empty));
            acc_phase ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_map(CUtensorMap* m, const uint8_t* base, int Tpad, int nrb, int ncb, int box_r, int box_c) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return -1; }
    cuuint64_t dims[3] = {(cuuint64_t)Tpad, (cuuint64_t)nrb, (cuuint64_t)ncb};
    cuuint64_t strides[2] = {(cuuint64_t)Tpad, (cuuint64_t)Tpad * nrb};
    cuuint32_t box[3] = {(cuuint32_t)tc::KSTAGE, (cuuint32_t)box_r, (cuuint32_t)box_c};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
    return 0;
}

// returns 0 = done on the tensor cores, 1 = shape not supported (caller falls back to the SIMT kernel), <0 error
int ring_s2_tensor(const uint8_t* hi, const uint8_t* lo, int nrb, int ncb, int T, int Tpad, int rr, double* S2,
                   cudaStream_t st) {
    using namespace tc;
    (void)T;
    // neighbour rows needed: [r0-2rr, r0+MR-1+2rr] must lie inside [r0-BSHIFT, r0-BSHIFT+NB-1]
    if (2 * rr > BSHIFT || MR - 1 + 2 * rr > NB - 1 - BSHIFT) return 1;
    if (Tpad % KSTAGE != 0) return 1;
    CUtensorMap mAh, mAl, mBh, mBl;
    if (make_map(&mAh, hi, Tpad, nrb, ncb, MR, MC) || make_map(&mAl, lo, Tpad, nrb, ncb, MR, MC) ||
        make_map(&mBh, hi, Tpad, nrb, ncb, NB, 1) || make_map(&mBl, lo, Tpad, nrb, ncb, NB, 1))
        return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static bool attr_set = false;
    if (!attr_set) {
        CNMFE_CUDA_OK(cudaFuncSetAttribute(ring_s2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    TcParams P;
    P.nrb = nrb; P.ncb = ncb; P.rr = rr; P.ND = tc_num_disp(rr);
    P.ntr = (nrb + MR - 1) / MR; P.ntc = (ncb + MC - 1) / MC; P.nbc = MC + 2 * rr;
    P.nitems = (long long)P.ntr * P.ntc * P.nbc;
    P.S2 = S2;
    const int nkb_total = Tpad / KSTAGE, per_pass = MAX_K_BYTES / KSTAGE;
    for (int kb0 = 0, pass = 0; kb0 < nkb_total; kb0 += per_pass, ++pass) {
        P.kb0 = kb0; P.kb1 = kb0 + per_pass < nkb_total ? kb0 + per_pass : nkb_total;
        P.accumulate = pass > 0;
        long long grid = P.nitems < sms ? P.nitems : sms;
        ring_s2_tc_kernel<<<(unsigned)grid, THREADS, SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, P);
        ++g_launch_count;
        CNMFE_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace cnmfe
