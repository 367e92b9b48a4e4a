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
// -> 8 MMAs per stage (4 K-steps of 32 B x {A_hi, A_lo} x [B_hi; B_lo] as one N = 224 operand) into three accumulators
// (hh, hl+lh, ll), then the epilogue warps read TMEM (tcgen05.ld 32x32b), combine in int64 and write the (pixel,
// displacement) run.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..9 = epilogue (TMEM lane quadrant warp % 4,
// two warps per quadrant splitting the columns: the accumulators are single-buffered, so the epilogue is a pipeline bubble).
// Two kernels: ring_s2_tc_kernel (one CTA per item, cluster-of-2 multicast of the pixel tile; default) and
// ring_s2_tc_pair_kernel (cta_group::2: two SMs per M = 256 MMA, each staging half of the neighbour operand; opt-in) --
// what bounds each is measured in DESIGN.md section 5.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
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
constexpr int EPI_STRIDE = 17;                            // doubles per pixel row of the epilogue transpose buffer (16 + pad)
constexpr int EPI_WARPS = 8;                              // two per TMEM lane quadrant: they split the accumulator columns
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_STRIDE * 8; // one 32 x 16 buffer per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_BYTES;
constexpr int TMEM_COLS = 512;
constexpr int COL_HH = 0, COL_MID = NB, COL_LL = 2 * NB;   // [hh | hl+lh | ll]: adjacent, so one N = 2*NB MMA spans two of them
constexpr int THREADS = 64 + 32 * EPI_WARPS;
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
// multicast variant: the box lands at the same shared offset of every CTA in ctaMask and completes tx on the mbarrier at
// the same offset of each of them
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
        :
        : "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
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
// arrive on the mbarrier at this offset in every CTA of ctaMask once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
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
    long long ngroups;          // cluster work items (each = CL CTA items)
    double* S2;
};

// Work item of CTA `rank` in cluster work item g -> (r0, c0, cB): tile-major, neighbour column fastest, so the CL CTAs of
// a cluster share the pixel tile.  (A neighbour-column-major order -- every B column read from DRAM once per band of tile
// rows -- was measured and is slower: 32.7 vs 29.0 ms; the kernel is not DRAM/L2 bound.)
template <int CL>
__device__ __forceinline__ bool tc_decode(const TcParams& P, long long g, int rank, int* r0, int* c0, int* cB) {
    using namespace tc;
    const long long item = g * CL + rank;
    const int j = (int)(item % P.nbc);
    const long long tile = item / P.nbc;
    const int tr = (int)(tile % P.ntr), tcx = (int)(tile / P.ntr);
    *r0 = tr * MR; *c0 = tcx * MC; *cB = *c0 + j;
    return true;
}

// CL = CTAs per cluster.  The CL CTAs of a cluster work on the SAME pixel tile with CL consecutive neighbour columns, so
// the A tile (the larger operand) is fetched from L2 once per cluster: every CTA loads 1/CL of it and multicasts the slice
// to all of them.  The kernel is bound per SM by shared-memory port traffic (operand reads of the MMAs + TMA writes = 37 KB per
// 224 MMA cycles; its time scales with 1 / SMs), not by the MMAs or the chip-level L2.
template <int CL>
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
    double* epi_buf = reinterpret_cast<double*>(tiles + STAGES * STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), CL); }
        mbar_init(smem_u32(tmem_full), 1);
        mbar_init(smem_u32(tmem_empty), EPI_WARPS);
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
    if (CL > 1) cluster_sync_all();            // peers' barriers are initialised before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int nkb = P.kb1 - P.kb0;
    // work distribution: a cluster takes CL consecutive items (same tile, neighbour columns j .. j+CL-1; nbc % CL == 0)
    const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
    const long long g0 = blockIdx.x / CL, gstep = gridDim.x / CL;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
    constexpr int A_SLICE = A_BYTES / CL;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long g = g0; g < P.ngroups; g += gstep) {
                int r0, c0, cB;
                if (!tc_decode<CL>(P, g, crank, &r0, &c0, &cB)) continue;
                const int rB0 = max(0, r0 - BSHIFT);
                for (int kb = P.kb0; kb < P.kb1; ++kb) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    uint8_t* st = tiles + stage * STAGE_BYTES;
                    mbar_expect_tx(fb, STAGE_BYTES);
                    if (CL > 1) {
                        tma_load_3d_mc(smem_u32(st + crank * A_SLICE), &tmA_hi, fb, kb * KSTAGE, r0, c0 + crank * (MC / CL), CMASK);
                        tma_load_3d_mc(smem_u32(st + A_BYTES + crank * A_SLICE), &tmA_lo, fb, kb * KSTAGE, r0, c0 + crank * (MC / CL), CMASK);
                    } else {
                        tma_load_3d(smem_u32(st), &tmA_hi, fb, kb * KSTAGE, r0, c0);
                        tma_load_3d(smem_u32(st + A_BYTES), &tmA_lo, fb, kb * KSTAGE, r0, c0);
                    }
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
            // B_hi and B_lo are adjacent in the stage (NB rows = 14 swizzle atoms each), so a descriptor at B_hi with
            // N = 2*NB reads [B_hi; B_lo]:  A_hi x [B_hi;B_lo] -> [hh | hl]  and  A_lo x [B_hi;B_lo] -> [lh | ll], the
            // second one placed NB columns further so that lh lands on hl.  Two MMAs per K-step instead of four: each A
            // tile is fetched from shared memory once per K-step (operand reads 22 KB instead of 30 KB per 32 frames).
            const uint32_t idesc2 = (2u << 4) | ((uint32_t)((2 * NB) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long g = g0; g < P.ngroups; g += gstep) {
                { int r0, c0, cB; if (!tc_decode<CL>(P, g, crank, &r0, &c0, &cB)) continue; }
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
                        if ((kb | ks) == 0) {
                            // first K-step of the item initialises the three accumulators separately
                            const uint64_t dBl = make_desc_sw128(sa + 2 * A_BYTES + B_BYTES + ks * 32);
                            mma_i8(tmem_base + COL_HH, dAh, dBh, idesc, 0u);
                            mma_i8(tmem_base + COL_MID, dAh, dBl, idesc, 0u);
                            mma_i8(tmem_base + COL_MID, dAl, dBh, idesc, 1u);
                            mma_i8(tmem_base + COL_LL, dAl, dBl, idesc, 0u);
                        } else {
                            mma_i8(tmem_base + COL_HH, dAh, dBh, idesc2, 1u);
                            mma_i8(tmem_base + COL_MID, dAl, dBh, idesc2, 1u);
                        }
                    }
                    // frees the smem stage (in every CTA that multicasts into it) when these MMAs retire
                    if (CL > 1) umma_commit_mc(smem_u32(empty_bar + stage), CMASK); else umma_commit(smem_u32(empty_bar + stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(smem_u32(tmem_full));                 // accumulators complete
                acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: warps 2..9 -> TMEM lane quadrant warp%4
        const int quad = warp & 3;
        // the two warps of a quadrant split the NB accumulator columns (16-column steps): [0, 64) and [64, NB)
        const int cc_begin = ((warp - 2) >> 2) ? 64 : 0, cc_end = ((warp - 2) >> 2) ? NB : 64;
        uint32_t acc_phase = 0;
 for (long long g = g0; g < P.ngroups; g += gstep) {
            int r0, c0, cB;
            if (!tc_decode<CL>(P, g, crank, &r0, &c0, &cB)) continue;
            const int rB0 = max(0, r0 - BSHIFT);
            const int cp = c0 + quad;                          // this warp's pixel column; TMEM lane = pixel row r0 + lane
            const int dc = cB - cp;
            const bool warp_ok = (cp < P.ncb) && (cB < P.ncb) && (dc >= 0) && (dc <= 2 * P.rr);   // warp-uniform
            mbar_wait(smem_u32(tmem_full), acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
            // The 16 accumulator columns a thread reads per step are 16 consecutive displacements of ITS pixel, i.e.
            // 128 contiguous bytes of that pixel's S2 row -- but across the lanes the rows are ND*8 bytes apart.  The
            // values are therefore transposed through shared memory so that a half-warp writes one pixel's run:
            // every store instruction covers two 128-byte runs instead of 32 scattered 8-byte words.
            double* stg = epi_buf + (warp - 2) * (32 * EPI_STRIDE);
            const int hx = lane & 15, hp = lane >> 4;
            const int id0 = (dc == 0) ? 0 : (2 * P.rr + 1) + (dc - 1) * (4 * P.rr + 1) + 2 * P.rr;   // disp id = id0 + dr
            if (warp_ok)
#pragma unroll 1
            for (int cc = cc_begin; cc < cc_end; cc += 16) {
                uint32_t hh[16], mid[16], ll[16];
                tmem_ld16(taddr + COL_HH + cc, hh);
                tmem_ld16(taddr + COL_MID + cc, mid);
                tmem_ld16(taddr + COL_LL + cc, ll);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int x = 0; x < 16; ++x) {
                    const long long v = ((long long)(int)hh[x] << 16) + ((long long)(int)mid[x] << 8) + (long long)(int)ll[x];
                    stg[lane * EPI_STRIDE + x] = (double)v;
                }
                __syncwarp();
                const int rn = rB0 + cc + hx;                  // neighbour row of this lane's column
#pragma unroll 4
                for (int it = 0; it < 16; ++it) {
                    const int px = 2 * it + hp, rp = r0 + px;
                    const int dr = rn - rp;
                    if (rp < P.nrb && rn < P.nrb && dr >= -2 * P.rr && dr <= 2 * P.rr && (dc > 0 || dr >= 0)) {
                        double* o = P.S2 + ((size_t)cp * P.nrb + rp) * (size_t)P.ND + (id0 + dr);
                        const double val = stg[px * EPI_STRIDE + hx];
                        if (P.accumulate) *o += val; else *o = val;
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tmem_empty));
            acc_phase ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();            // no CTA leaves while a peer may still signal its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant
// cta_group::2: the two CTAs of a cluster (two SMs) work on two column-adjacent pixel tiles (32 x 8 pixels) against the SAME
// neighbour column, as ONE M = 256 MMA: each CTA stages its own A tiles (hi, lo) and only HALF of the B operand -- the leader
// the hi plane of the neighbour column, its peer the lo plane ([B_hi; B_lo] is the N = 224 operand, split by halves across
// the pair) -- and the tensor cores exchange the halves.  Why: the single-CTA kernel is bound by shared-memory bandwidth
// (per 32-byte K step the MMAs read 22 KB of operands and TMA writes 15 KB against 224 MMA cycles = 165 B/clk of a 128 B/clk
// port; measured 0.71 of the MMA rate, and the time scales with 1/SMs, so it is not the chip-level L2).  The pair reads
// 15 KB and writes 11.5 KB per K step and CTA (118 B/clk), and a stage is 46 KB instead of 60 KB, so four stages fit.
// Price: the pair covers 8 pixel columns, so 44 instead of 40 neighbour columns per tile (37 of them useful per pixel).
namespace tc2 {
constexpr int MR = tc::MR, MC = tc::MC, NB = tc::NB, BSHIFT = tc::BSHIFT, KSTAGE = tc::KSTAGE;
constexpr int STAGES = 4;
constexpr int A_BYTES = tc::A_BYTES;                      // one byte plane of this CTA's pixel tile
constexpr int B_BYTES = tc::B_BYTES;                      // ONE byte plane of the neighbour column (this CTA's half of B)
constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;        // 47104
constexpr int PAIR_TX = 2 * STAGE_BYTES;                  // bytes that land per stage in the two CTAs together
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + tc::EPI_BYTES;
constexpr int COL_HH = 0, COL_HL = NB, COL_LH = 2 * NB, COL_LL = 3 * NB;   // A_hi x [B_hi;B_lo] | A_lo x [B_hi;B_lo]
constexpr int PAIR_COLS = 2 * MC;
}  // namespace tc2

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// box -> this CTA's shared memory, transaction bytes -> the mbarrier `bar` (a shared::cluster address: the leader's barrier)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mma_i8_pair(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        :
        : "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// cluster work item g -> (r0, first column of the pair's 8, neighbour column)
__device__ __forceinline__ void tc2_decode(const TcParams& P, long long g, int* r0, int* c0, int* cB) {
    const int j = (int)(g % P.nbc);
    const long long tile = g / P.nbc;
    const int tr = (int)(tile % P.ntr), tcx = (int)(tile / P.ntr);
    *r0 = tr * tc2::MR; *c0 = tcx * tc2::PAIR_COLS; *cB = *c0 + j;
}

__global__ void __launch_bounds__(tc::THREADS, 1)
ring_s2_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                       const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                       const TcParams P) {
    using namespace tc2;
    constexpr int EPI_WARPS = tc::EPI_WARPS, EPI_STRIDE = tc::EPI_STRIDE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]  used in the leader only (both CTAs' TMA bytes land on it)
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]  one per CTA (multicast commit)
    uint64_t* tmem_full = bars + 2 * STAGES;   // [1]       one per CTA (multicast commit)
    uint64_t* tmem_empty = tmem_full + 1;      // [1]       used in the leader only (both CTAs' epilogue warps arrive on it)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    double* epi_buf = reinterpret_cast<double*>(tiles + STAGES * STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();   // 0 = leader (issues the MMAs), 1 = peer
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
        mbar_init(smem_u32(tmem_full), 1);
        mbar_init(smem_u32(tmem_empty), 2 * EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)tc::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int nkb = P.kb1 - P.kb0;
    const long long g0 = blockIdx.x / 2, gstep = gridDim.x / 2;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const CUtensorMap* mapB = crank == 0 ? &tmB_hi : &tmB_lo;
            for (long long g = g0; g < P.ngroups; g += gstep) {
                int r0, c0, cB;
                tc2_decode(P, g, &r0, &c0, &cB);
                c0 += crank * MC;
                const int rB0 = max(0, r0 - BSHIFT);
                for (int kb = P.kb0; kb < P.kb1; ++kb) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb_local = smem_u32(full_bar + stage);
                    const uint32_t fb = mapa_u32(fb_local, 0);            // the leader's barrier
                    uint8_t* st = tiles + stage * STAGE_BYTES;
                    if (crank == 0) mbar_expect_tx(fb_local, PAIR_TX);
                    tma_load_3d_pair(smem_u32(st), &tmA_hi, fb, kb * KSTAGE, r0, c0);
                    tma_load_3d_pair(smem_u32(st + A_BYTES), &tmA_lo, fb, kb * KSTAGE, r0, c0);
                    tma_load_3d_pair(smem_u32(st + 2 * A_BYTES), mapB, fb, kb * KSTAGE, rB0, cB);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one lane of the leader)
        if (lane == 0 && crank == 0) {
            // M = 256 (128 rows per CTA), N = 2 * NB = [B_hi; B_lo] (first half in the leader, second in the peer)
            const uint32_t idesc = (2u << 4) | ((uint32_t)((2 * NB) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long g = g0; g < P.ngroups; g += gstep) {
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
                        const uint64_t dB = make_desc_sw128(sa + 2 * A_BYTES + ks * 32);
                        const uint32_t acc = (kb | ks) != 0 ? 1u : 0u;
                        mma_i8_pair(tmem_base + COL_HH, dAh, dB, idesc, acc);      // [hh | hl]
                        mma_i8_pair(tmem_base + COL_LH, dAl, dB, idesc, acc);      // [lh | ll]
                    }
                    umma_commit_pair_mc(smem_u32(empty_bar + stage), (uint16_t)3);   // frees the stage in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair_mc(smem_u32(tmem_full), (uint16_t)3);               // accumulators complete, both CTAs
                acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (both CTAs, own pixel tile)
        const int quad = warp & 3;
        const int cc_begin = ((warp - 2) >> 2) ? 64 : 0, cc_end = ((warp - 2) >> 2) ? NB : 64;
        const uint32_t empty_leader = mapa_u32(smem_u32(tmem_empty), 0);
        uint32_t acc_phase = 0;
        for (long long g = g0; g < P.ngroups; g += gstep) {
            int r0, c0, cB;
            tc2_decode(P, g, &r0, &c0, &cB);
            c0 += crank * MC;
            const int rB0 = max(0, r0 - BSHIFT);
            const int cp = c0 + quad;
            const int dc = cB - cp;
            const bool warp_ok = (cp < P.ncb) && (cB < P.ncb) && (dc >= 0) && (dc <= 2 * P.rr);   // warp-uniform
            mbar_wait(smem_u32(tmem_full), acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
            double* stg = epi_buf + (warp - 2) * (32 * EPI_STRIDE);
            const int hx = lane & 15, hp = lane >> 4;
            const int id0 = (dc == 0) ? 0 : (2 * P.rr + 1) + (dc - 1) * (4 * P.rr + 1) + 2 * P.rr;
            if (warp_ok)
#pragma unroll 1
            for (int cc = cc_begin; cc < cc_end; cc += 16) {
                uint32_t hh[16], hl[16], lh[16], ll[16];
                tmem_ld16(taddr + COL_HH + cc, hh);
                tmem_ld16(taddr + COL_HL + cc, hl);
                tmem_ld16(taddr + COL_LH + cc, lh);
                tmem_ld16(taddr + COL_LL + cc, ll);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int x = 0; x < 16; ++x) {
                    const long long v = ((long long)(int)hh[x] << 16) + (((long long)(int)hl[x] + (long long)(int)lh[x]) << 8) +
                                        (long long)(int)ll[x];
                    stg[lane * EPI_STRIDE + x] = (double)v;
                }
                __syncwarp();
                const int rn = rB0 + cc + hx;
#pragma unroll 4
                for (int it = 0; it < 16; ++it) {
                    const int px = 2 * it + hp, rp = r0 + px;
                    const int dr = rn - rp;
                    if (rp < P.nrb && rn < P.nrb && dr >= -2 * P.rr && dr <= 2 * P.rr && (dc > 0 || dr >= 0)) {
                        double* o = P.S2 + ((size_t)cp * P.nrb + rp) * (size_t)P.ND + (id0 + dr);
                        const double val = stg[px * EPI_STRIDE + hx];
                        if (P.accumulate) *o += val; else *o = val;
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(empty_leader);
            acc_phase ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                        // no CTA leaves while its peer may still signal its barriers or read its B half
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tc::TMEM_COLS)
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
    TcParams P;
    P.nrb = nrb; P.ncb = ncb; P.rr = rr; P.ND = tc_num_disp(rr);
    P.ntr = (nrb + MR - 1) / MR; P.ntc = (ncb + MC - 1) / MC; P.nbc = MC + 2 * rr;
    P.nitems = (long long)P.ntr * P.ntc * P.nbc;
    P.S2 = S2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (const char* e = getenv("CNMFE_TC_SMS")) { int v = atoi(e); if (v >= 2 && v < sms) sms = v; }   // A/B knob: SMs the kernel occupies
    const int nkb_total = Tpad / KSTAGE, per_pass = MAX_K_BYTES / KSTAGE;
    // CNMFE_TC_MODE=pair selects the CTA-pair kernel (cta_group::2); default: the one-CTA kernel with the multicast A tile
    const char* mode = getenv("CNMFE_TC_MODE");
    if (mode && !strcmp(mode, "pair")) {
        P.nbc = tc2::PAIR_COLS + 2 * rr;
        P.ntc = (ncb + tc2::PAIR_COLS - 1) / tc2::PAIR_COLS;
        P.nitems = (long long)P.ntr * P.ntc * P.nbc;
        P.ngroups = P.nitems;
        CUtensorMap mAh, mAl, mBh, mBl;
        if (make_map(&mAh, hi, Tpad, nrb, ncb, MR, MC) || make_map(&mAl, lo, Tpad, nrb, ncb, MR, MC) ||
            make_map(&mBh, hi, Tpad, nrb, ncb, NB, 1) || make_map(&mBl, lo, Tpad, nrb, ncb, NB, 1))
            return -1;
        CNMFE_CUDA_OK(cudaFuncSetAttribute(ring_s2_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = tc2::SMEM_BYTES; cfg.stream = st;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = sms / 2;
        cfg.gridDim = dim3((unsigned)(nclusters * 2));
        int maxc = 0;
        if (cudaOccupancyMaxActiveClusters(&maxc, ring_s2_tc_pair_kernel, &cfg) == cudaSuccess && maxc > 0 && maxc < nclusters) nclusters = maxc;
        if (P.ngroups < nclusters) nclusters = (int)P.ngroups;
        cfg.gridDim = dim3((unsigned)(nclusters * 2));
        for (int kb0 = 0, pass = 0; kb0 < nkb_total; kb0 += per_pass, ++pass) {
            P.kb0 = kb0; P.kb1 = kb0 + per_pass < nkb_total ? kb0 + per_pass : nkb_total;
            P.accumulate = pass > 0;
            CNMFE_CUDA_OK(cudaLaunchKernelEx(&cfg, ring_s2_tc_pair_kernel, mAh, mAl, mBh, mBl, P));
            ++g_launch_count;
            CNMFE_CUDA_OK(cudaGetLastError());
        }
        return 0;
    }
    // cluster size: 2 when the neighbour-column count allows it (measured: 1 -> 29.2 ms, 2 -> 28.6 ms, 4 -> 32.9 ms: fewer
    // co-resident clusters).  CNMFE_TC_CLUSTER=1|2|4 overrides for A/B measurements.
    int CL = (P.nbc % 2 == 0) ? 2 : 1;
    if (const char* e = getenv("CNMFE_TC_CLUSTER")) { int v = atoi(e); if ((v == 1 || v == 2 || v == 4) && P.nbc % v == 0) CL = v; }
    CUtensorMap mAh, mAl, mBh, mBl;
    if (make_map(&mAh, hi, Tpad, nrb, ncb, MR, MC / CL) || make_map(&mAl, lo, Tpad, nrb, ncb, MR, MC / CL) ||
        make_map(&mBh, hi, Tpad, nrb, ncb, NB, 1) || make_map(&mBl, lo, Tpad, nrb, ncb, NB, 1))
        return -1;
    void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcParams) =
        CL == 4 ? ring_s2_tc_kernel<4> : (CL == 2 ? ring_s2_tc_kernel<2> : ring_s2_tc_kernel<1>);
    CNMFE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // persistent grid: as many clusters as can be co-resident (1 CTA per SM)
    int nclusters = sms / CL;
    if (CL > 1) {
        cfg.gridDim = dim3((unsigned)(nclusters * CL));
        int maxc = 0;
        if (cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg) == cudaSuccess && maxc > 0 && maxc < nclusters) nclusters = maxc;
    }
    P.ngroups = P.nitems / CL;
    if (P.ngroups < nclusters) nclusters = (int)P.ngroups;
    cfg.gridDim = dim3((unsigned)(nclusters * CL));
    for (int kb0 = 0, pass = 0; kb0 < nkb_total; kb0 += per_pass, ++pass) {
        P.kb0 = kb0; P.kb1 = kb0 + per_pass < nkb_total ? kb0 + per_pass : nkb_total;
        P.accumulate = pass > 0;
        CNMFE_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mAh, mAl, mBh, mBl, P));
        ++g_launch_count;
        CNMFE_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace cnmfe
