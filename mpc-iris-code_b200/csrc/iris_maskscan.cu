// Denominators-only scan (BASELINE config 3): popcount(rot(qmask, j-15) & dbmask_i) for ONE query mask over the
// resident masks, reference src/lib.rs:69-79 -> src/arch/generic.rs:4-9.
//
// Same arithmetic as the fused scan (mask bits expanded to bytes of value 2^t, query operand 2^(7-t), D >> 7),
// but the expanded A operand never touches shared memory: the expander threads write it straight into TENSOR
// MEMORY (tcgen05.st) and the UMMA reads A from TMEM (TS form) and only the 4 KiB query tile from smem.  That
// removes 32 KiB of shared-memory traffic per 2 KiB of database, which is what bounds the smem version
// (scan_kernel<false,true>) far below the HBM roofline.
//
// Per stage: 512 mask bits of 128 rows = 8 KiB of packed database (contiguous in HBM) + 16 KiB of query operand.
//   producer : bulk copies into an 8-deep smem ring
//   expanders: LDS packed bits -> 128 LOP3 -> 4 x tcgen05.st.32x32b.x32 into a 3-deep TMEM ring (128 columns each)
//   UMMA     : 16 x tcgen05.mma kind::i8 (A = TMEM, B = smem, M=128, N=32) per stage
//   epilogue : tcgen05.ld, >> 7, 62-byte rows, double-buffered accumulator
#include <cuda_runtime.h>

#include "iris_kernels.cuh"
#include "iris_ptx.cuh"

namespace iris {

void count_launch_external();

constexpr int kMsSub = 4;                                    // 128-bit chunks per stage
constexpr int kMsStagesPerTile = kChunks / kMsSub;           // 25
constexpr int kMsPkBytes = kMsSub * kMaskChunkBytes;         // 8 KiB
constexpr int kMsQmBytes = kMsSub * kQmChunkBytes;           // 16 KiB
constexpr int kMsStageBytes = kMsPkBytes + kMsQmBytes;       // 24 KiB
constexpr int kMsStages = 8;
constexpr int kMsARing = 2;                                  // TMEM A stages (128 columns each)
constexpr int kMsAccSplit = 4;                               // independent accumulators (one per K step): an N=32
                                                             // UMMA is ~16 cycles of work but ~100 cycles deep, so
                                                             // back-to-back accumulation into ONE tile serialises
constexpr int kMsOutStageBytes = 8192;
constexpr int kMsSmemBytes = 1024 + kMsStages * kMsStageBytes + kMsOutStageBytes + 512;
constexpr int kMsExpWarps = 8;                               // 2 per TMEM lane quadrant, each expands half a stage
constexpr int kMsThreads = (6 + kMsExpWarps) * 32;
constexpr uint32_t kMsAccCols = 2 * kMsAccSplit * 32;        // 2 accumulator buffers x 4 partial tiles x 32 columns
constexpr uint32_t kMsTmemCols = 512;
static_assert(kChunks % kMsSub == 0, "stages must tile the K dimension");
static_assert(kMsAccCols + kMsARing * 128 <= kMsTmemCols, "TMEM budget");
static_assert(kMsSmemBytes <= 232448, "exceeds 227 KiB of shared memory");

enum MsWatchdog { kWsProducer = 401, kWsMmaFull = 402, kWsMmaA = 403, kWsMmaTmem = 404, kWsExpFull = 405, kWsExpA = 406, kWsEpilogue = 407 };

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]   (TS form).  The B descriptor is passed as (lo, hi) words so that stepping
// through a stage is ONE uniform add per operand instead of rebuilding the 64-bit descriptor per instruction.
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
constexpr uint32_t kDescHiSw128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B

__device__ __forceinline__ void ms_copy_out(const uint8_t* stage, uint8_t* gbase, int b0, int b1, int tid) {
    if (b1 <= b0) return;
    int body0 = (b0 + 15) & ~15, body1 = b1 & ~15;
    if (body0 > body1) {
        for (int b = b0 + 2 * tid; b < b1; b += 2 * 128)
            *reinterpret_cast<uint16_t*>(gbase + b) = *reinterpret_cast<const uint16_t*>(stage + b);
        return;
    }
    for (int b = b0 + 2 * tid; b < body0; b += 2 * 128)
        *reinterpret_cast<uint16_t*>(gbase + b) = *reinterpret_cast<const uint16_t*>(stage + b);
    for (int b = body0 + 16 * tid; b < body1; b += 16 * 128)
        *reinterpret_cast<uint4*>(gbase + b) = *reinterpret_cast<const uint4*>(stage + b);
    for (int b = body1 + 2 * tid; b < b1; b += 2 * 128)
        *reinterpret_cast<uint16_t*>(gbase + b) = *reinterpret_cast<const uint16_t*>(stage + b);
}

__global__ void __launch_bounds__(kMsThreads, 1) mask_scan_kernel(const ScanParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw_addr);
    uint8_t* const out_stage_ptr = base_ptr + kMsStages * kMsStageBytes;
    const uint32_t bars = base + kMsStages * kMsStageBytes + kMsOutStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };                          // smem stage landed
    auto empty_bar = [&](int s) { return bars + 8u * (kMsStages + s); };           // count: expander warps + UMMA commit
    auto afull_bar = [&](int a) { return bars + 8u * (2 * kMsStages + a); };       // TMEM A stage written (all expander warps)
    auto aempty_bar = [&](int a) { return bars + 8u * (2 * kMsStages + kMsARing + a); };   // UMMA done reading it
    auto tfull_bar = [&](int b) { return bars + 8u * (2 * kMsStages + 2 * kMsARing + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (2 * kMsStages + 2 * kMsARing + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * kMsStages + 2 * kMsARing + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        out_stage_ptr + kMsOutStageBytes + 8 * (2 * kMsStages + 2 * kMsARing + 4));

    const int warp = ptx::warp_idx_sync();      // warp-uniform role index
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMsStages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), kMsExpWarps + 1);
        }
        for (int a = 0; a < kMsARing; ++a) {
            ptx::mbar_init(afull_bar(a), kMsExpWarps);
            ptx::mbar_init(aempty_bar(a), 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(tfull_bar(b), 1);
            ptx::mbar_init(tempty_bar(b), 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 5) ptx::tmem_alloc(tmem_slot, kMsTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const uint32_t tile0 = p.tile_begin + blockIdx.x;
    const uint32_t tile_step = gridDim.x;

    if (warp == 4) {
        // ------------------------------------------------------------------ producer (whole warp in uniform flow,
        // one elected lane issues: operands stay in uniform registers)
        const uint64_t pol_stream = ptx::policy_evict_first();
        const uint64_t pol_keep = ptx::policy_evict_last();
        int stage = 0;
        uint32_t phase = 0;
        for (uint32_t tile = tile0; tile < p.tile_end; tile += tile_step) {
            const uint8_t* mk = p.masks + (size_t)tile * kMaskTileBytes;
            for (int c = 0; c < kMsStagesPerTile; ++c) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.error, kWsProducer);
                const uint32_t sbase = base + stage * kMsStageBytes;
                const uint32_t fb = full_bar(stage);
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(fb, kMsStageBytes);
                    ptx::bulk_g2s_hint(sbase, mk + (size_t)c * kMsPkBytes, kMsPkBytes, fb, pol_stream);
                    ptx::bulk_g2s_hint(sbase + kMsPkBytes, p.qm + (size_t)c * kMsQmBytes, kMsQmBytes, fb, pol_keep);
                }
                __syncwarp();
                if (++stage == kMsStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ UMMA issuer (A from TMEM, B from smem)
        constexpr uint32_t kIdesc32 = ptx::umma_idesc_i8(32);
        int stage = 0, ar = 0;
        uint32_t phase = 0, aphase = 0, it = 0;
        for (uint32_t tile = tile0; tile < p.tile_end; tile += tile_step, ++it) {
            const uint32_t buf = it & 1u;
            ptx::mbar_wait(tempty_bar(buf), ((it >> 1) & 1u) ^ 1u, p.error, kWsMmaTmem);
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + buf * (kMsAccSplit * 32u);
            for (int c = 0; c < kMsStagesPerTile; ++c) {
                ptx::mbar_wait(full_bar(stage), phase, p.error, kWsMmaFull);
                ptx::mbar_wait(afull_bar(ar), aphase, p.error, kWsMmaA);
                ptx::tc_fence_after();
                const uint32_t qbase = base + stage * kMsStageBytes + kMsPkBytes;
                const uint32_t abase = tmem_base + kMsAccCols + ar * 128u;
                const uint32_t blo0 = ((qbase & 0x3FFFFu) >> 4) | (1u << 16);
                if (ptx::elect_one_sync()) {
#pragma unroll
                    for (int sub = 0; sub < kMsSub; ++sub) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // K step k accumulates into partial tile k
                            umma_i8_ts(d + k * 32, abase + sub * 32 + k * 8, blo0 + ((sub * kQmChunkBytes + 32 * k) >> 4),
                                       kDescHiSw128, kIdesc32, sub ? 1u : (c ? 1u : 0u));
                        }
                    }
                    ptx::umma_commit(aempty_bar(ar));
                    ptx::umma_commit(empty_bar(stage));
                    if (c == kMsStagesPerTile - 1) ptx::umma_commit(tfull_bar(buf));
                }
                __syncwarp();
                if (++stage == kMsStages) { stage = 0; phase ^= 1u; }
                if (++ar == kMsARing) { ar = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= 6) {
        // ------------------------------------------------------------------ expanders: packed bits -> TMEM A operand
        const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
        const int half = (warp - 6) >> 2;                 // which half of the stage's 128-bit chunks this warp expands
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        int stage = 0, ar = 0;
        uint32_t phase = 0, aphase = 0;
        for (uint32_t tile = tile0; tile < p.tile_end; tile += tile_step) {
            for (int c = 0; c < kMsStagesPerTile; ++c) {
                ptx::mbar_wait(full_bar(stage), phase, p.error, kWsExpFull);
                ptx::mbar_wait(aempty_bar(ar), aphase ^ 1u, p.error, kWsExpA);
                ptx::tc_fence_after();
                const uint8_t* pk = base_ptr + stage * kMsStageBytes;
                const uint32_t abase = tmem_base + lane_addr + kMsAccCols + ar * 128u;
#pragma unroll
                for (int s2 = 0; s2 < kMsSub / 2; ++s2) {
                    const int sub = half * (kMsSub / 2) + s2;
                    const uint4 x = *reinterpret_cast<const uint4*>(pk + sub * kMaskChunkBytes + row * 16);
                    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
                    uint32_t v[32];
#pragma unroll
                    for (int w = 0; w < 4; ++w)
#pragma unroll
                        for (int t = 0; t < 8; ++t) v[8 * w + t] = xs[w] & (0x01010101u << t);
                    tmem_st32(abase + sub * 32, v);
                }
                tmem_wait_st();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(afull_bar(ar));
                    ptx::mbar_arrive(empty_bar(stage));   // this warp no longer needs the packed bytes
                }
                if (++stage == kMsStages) { stage = 0; phase ^= 1u; }
                if (++ar == kMsARing) { ar = 0; aphase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0..3)
        const int row = threadIdx.x;
        uint32_t it = 0;
        for (uint32_t tile = tile0; tile < p.tile_end; tile += tile_step, ++it) {
            const uint32_t buf = it & 1u;
            ptx::mbar_wait(tfull_bar(buf), (it >> 1) & 1u, p.error, kWsEpilogue);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (kMsAccSplit * 32u);
            const int64_t trow0 = (int64_t)tile * kTileRows;
            int64_t lo = (int64_t)p.row_begin - trow0, hi = (int64_t)p.row_end - trow0;
            const int r0 = (int)(lo < 0 ? 0 : (lo > kTileRows ? kTileRows : lo));
            const int r1 = (int)(hi < 0 ? 0 : (hi > kTileRows ? kTileRows : hi));
            const int64_t tile_off = (trow0 - (int64_t)p.row_begin) * kOutRowBytes;
            uint32_t a[32];
            {
                uint32_t b[32], c2[32], d2[32];
                ptx::tmem_ld32(taddr, a);
                ptx::tmem_ld32(taddr + 32, b);
                ptx::tmem_ld32(taddr + 64, c2);
                ptx::tmem_ld32(taddr + 96, d2);
                ptx::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] += b[j] + c2[j] + d2[j];   // sum of the 4 partial tiles
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(buf));
            const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(p.den_out) + tile_off) & 15);
            uint8_t* st = out_stage_ptr + shift + row * kOutRowBytes;
#pragma unroll
            for (int j = 0; j < IRIS_ROTATIONS; ++j) *reinterpret_cast<uint16_t*>(st + 2 * j) = (uint16_t)(a[j] >> 7);
            ptx::named_bar_sync(1, 128);
            ms_copy_out(out_stage_ptr, reinterpret_cast<uint8_t*>(p.den_out) + tile_off - shift, (int)shift + r0 * kOutRowBytes,
                        (int)shift + r1 * kOutRowBytes, row);
            ptx::named_bar_sync(1, 128);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, kMsTmemCols);
}

cudaError_t launch_mask_scan(const ScanParams& p, int num_sms, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && !configured[dev]) {
        e = cudaFuncSetAttribute(mask_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMsSmemBytes);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const uint32_t tiles = p.tile_end - p.tile_begin;
    if (tiles == 0) return cudaSuccess;
    const uint32_t grid = tiles < (uint32_t)num_sms ? tiles : (uint32_t)num_sms;
    mask_scan_kernel<<<grid, kMsThreads, kMsSmemBytes, stream>>>(p);
    count_launch_external();
    return cudaGetLastError();
}

}  // namespace iris
