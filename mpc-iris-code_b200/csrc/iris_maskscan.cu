// Denominators-only scan (BASELINE config 3): popcount(rot(qmask, j-15) & dbmask_i) for ONE query mask over the
// resident masks, reference src/lib.rs:69-79 -> src/arch/generic.rs:4-9.
//
// Same arithmetic as the fused scan (mask bits expanded to bytes of value 2^t, query operand 2^(7-t), D >> 7),
// but the expanded A operand never touches shared memory: the expander threads write it straight into TENSOR
// MEMORY (tcgen05.st) and the UMMA reads A from TMEM (TS form) and only the 4 KiB query tile from smem.  That
// removes 32 KiB of shared-memory traffic per 2 KiB of database, which is what bounds the smem version
// (scan_kernel<false,true>) far below the HBM roofline.
//
// A CTA owns TWO consecutive 128-row tiles that share every query chunk (halves the operand traffic over the
// L2->SM fabric), and each tile has its OWN issuing warp and accumulators.  An expander warp's per-stage chain
// (LDS -> LOP -> tcgen05.st -> wait::st -> arrive) is ~600 cycles of LATENCY, so two sets of expander warps take
// alternate stages (splitting one stage over more warps does not help; measured).
//
// What bounds it is the tensor-core side, not HBM: with the expanders switched off (timing experiment) the UMMAs
// alone take 0.32-0.36 ms per 1 M rows -- one N=32, K=32 kind::i8 instruction per ~30 cycles per SM although its
// MACs need 16.  That cost per instruction does not move with M (64 or 128), with the number of accumulators
// (one or two per tile) or with the number of issuing warps (two or four), so it is a floor of the instruction,
// and K = 32 bytes and N = 31 rotations are fixed here.  The complete kernel takes 0.363-0.373 ms alone (0.70 of
// the HBM roofline).  Details and the other experiments (cluster multicast, expander sets) in DESIGN.md 5.3.
//
// Per stage: 256 mask bits of 2 x 128 rows = 2 x 4 KiB of packed database (contiguous per tile) + 8 KiB of operand.
//   producer (warp 4)     : 3 bulk copies into a 12-deep smem ring
//   expanders (warps 7-22): LDS packed bits -> 64 LOP3 -> 2 x tcgen05.st.32x32b.x32 into a 3-deep TMEM ring;
//                           two sets of 8 warps alternate stages to overlap their per-stage latency chain
//   issuers (warps 5, 6)  : 8 x tcgen05.mma kind::i8 (A = TMEM, B = smem) per stage, one warp per tile
//   epilogue (warps 0-3)  : tcgen05.ld, >> 7, 62-byte rows; accumulators double-buffered
#include <cuda_runtime.h>

#include <atomic>

#include "iris_epilogue.cuh"
#include "iris_kernels.cuh"
#include "iris_ptx.cuh"

namespace iris {

void count_launch_external();

constexpr int kMsTiles = 2;                                       // row tiles per CTA
constexpr int kMsSub = 2;                                         // 128-bit chunks per stage
constexpr int kMsStagesPerTile = kChunks / kMsSub;                // 50
constexpr int kMsPkBytes = kMsSub * kMaskChunkBytes;              // 4 KiB per tile
constexpr int kMsQmBytes = kMsSub * kQmChunkBytes;                // 8 KiB
constexpr int kMsOffQm = kMsTiles * kMsPkBytes;
constexpr int kMsStageBytes = kMsOffQm + kMsQmBytes;              // 16 KiB
constexpr int kMsStages = 12;
constexpr int kMsARing = 3;                                       // TMEM A slots: 2 tiles x 2 chunks x 32 columns
constexpr int kMsOutStageBytes = 8192;
constexpr int kMsSmemBytes = 1024 + kMsStages * kMsStageBytes + kMsOutStageBytes + 512;
constexpr int kMsIssuerWarp0 = 5;                                 // warps 5, 6
constexpr int kMsExpWarp0 = 7;                                    // expander warps: [set][tile][TMEM lane quadrant]
constexpr int kMsExpSets = 2;                                    // sets alternate stages: an expander's per-stage chain
                                                                  // (LDS -> LOP -> tcgen05.st -> wait::st -> arrive) is
                                                                  // ~600 cycles of latency, so two sets overlap it
constexpr int kMsThreads = (kMsExpWarp0 + kMsExpSets * 4 * kMsTiles) * 32;     // 736
constexpr uint32_t kMsAccCols = 2 * kMsTiles * 32;                // [buffer][tile] x 32 columns
constexpr uint32_t kMsASlotCols = kMsTiles * kMsSub * 32;         // 128
constexpr uint32_t kMsTmemCols = 512;
static_assert(kChunks % kMsSub == 0, "stages must tile the K dimension");
static_assert(kMsAccCols + kMsARing * kMsASlotCols <= kMsTmemCols, "TMEM budget");
static_assert(kMsSmemBytes <= 232448, "exceeds 227 KiB of shared memory");
static_assert(kMsStageBytes % 1024 == 0 && kMsOffQm % 1024 == 0, "operand tiles must stay 1024-byte aligned");

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


// p.tile_begin / p.tile_end are in units of 128-row tiles; this kernel walks PAIRS of tiles
// [tile_begin/2, ceil(tile_end/2)) -- the shard's capacity is a whole number of pairs and zero filled.
__global__ void __launch_bounds__(kMsThreads, 1) mask_scan_kernel(const ScanParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw_addr);
    uint8_t* const out_stage_ptr = base_ptr + kMsStages * kMsStageBytes;
    const uint32_t bars = base + kMsStages * kMsStageBytes + kMsOutStageBytes;
    // barrier table (8 bytes each)
    auto full_bar = [&](int s) { return bars + 8u * s; };                                   // stage landed (tx)
    auto empty_bar = [&](int s) { return bars + 8u * (kMsStages + s); };                    // 8 expander warps + 2 issuer commits
    auto afull_bar = [&](int a, int t) { return bars + 8u * (2 * kMsStages + 2 * a + t); };                 // 4 expander warps of tile t
    auto aempty_bar = [&](int a, int t) { return bars + 8u * (2 * kMsStages + 2 * kMsARing + 2 * a + t); }; // issuer t commit
    auto tfull_bar = [&](int b, int t) { return bars + 8u * (2 * kMsStages + 4 * kMsARing + 2 * b + t); };
    auto tempty_bar = [&](int b, int t) { return bars + 8u * (2 * kMsStages + 4 * kMsARing + 4 + 2 * b + t); };
    constexpr int kNumBars = 2 * kMsStages + 4 * kMsARing + 8;
    static_assert(8 * (kNumBars + 1) <= 512, "barrier table");
    const uint32_t tmem_slot = bars + 8u * kNumBars;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(out_stage_ptr + kMsOutStageBytes + 8 * kNumBars);

    const int warp = ptx::warp_idx_sync();      // warp-uniform role index
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMsStages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 4 * kMsTiles + kMsTiles);
        }
        for (int a = 0; a < kMsARing; ++a)
            for (int t = 0; t < kMsTiles; ++t) {
                ptx::mbar_init(afull_bar(a, t), 4);
                ptx::mbar_init(aempty_bar(a, t), 1);
            }
        for (int b = 0; b < 2; ++b)
            for (int t = 0; t < kMsTiles; ++t) {
                ptx::mbar_init(tfull_bar(b, t), 1);
                ptx::mbar_init(tempty_bar(b, t), 4);
            }
        ptx::fence_mbar_init();
    }
    if (warp == kMsIssuerWarp0) ptx::tmem_alloc(tmem_slot, kMsTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const uint32_t pair_begin = p.tile_begin / kMsTiles;
    const uint32_t pair_end = (p.tile_end + kMsTiles - 1) / kMsTiles;
    const uint32_t pair0 = pair_begin + blockIdx.x;
    const uint32_t pair_step = gridDim.x;

    if (warp == 4) {
        // ------------------------------------------------------------------ producer
        const uint64_t pol_stream = ptx::policy_evict_first();
        const uint64_t pol_keep = ptx::policy_evict_last();
        int stage = 0;
        uint32_t phase = 0;
        for (uint32_t pair = pair0; pair < pair_end; pair += pair_step) {
            const uint8_t* mk = p.masks + (size_t)pair * kMsTiles * kMaskTileBytes;
            for (int c = 0; c < kMsStagesPerTile; ++c) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.error, kWsProducer);
                const uint32_t sbase = base + stage * kMsStageBytes;
                const uint32_t fb = full_bar(stage);
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(fb, kMsStageBytes);
#pragma unroll
                    for (int t = 0; t < kMsTiles; ++t)
                        ptx::bulk_g2s_hint(sbase + t * kMsPkBytes, mk + (size_t)t * kMaskTileBytes + (size_t)c * kMsPkBytes,
                                           kMsPkBytes, fb, pol_stream);
                    ptx::bulk_g2s_hint(sbase + kMsOffQm, p.qm + (size_t)c * kMsQmBytes, kMsQmBytes, fb, pol_keep);
                }
                __syncwarp();
                if (++stage == kMsStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == kMsIssuerWarp0 || warp == kMsIssuerWarp0 + 1) {
        // ------------------------------------------------------------------ UMMA issuers: one warp per row tile
        const int t = warp - kMsIssuerWarp0;
        constexpr uint32_t kIdesc32 = ptx::umma_idesc_i8(32);
        int stage = 0, ar = 0;
        uint32_t phase = 0, aphase = 0, it = 0;
        for (uint32_t pair = pair0; pair < pair_end; pair += pair_step, ++it) {
            const uint32_t buf = it & 1u;
            ptx::mbar_wait(tempty_bar(buf, t), ((it >> 1) & 1u) ^ 1u, p.error, kWsMmaTmem);
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + buf * (kMsTiles * 32u) + t * 32u;
            for (int c = 0; c < kMsStagesPerTile; ++c) {
                ptx::mbar_wait(full_bar(stage), phase, p.error, kWsMmaFull);
                ptx::mbar_wait(afull_bar(ar, t), aphase, p.error, kWsMmaA);
                ptx::tc_fence_after();
                const uint32_t qbase = base + stage * kMsStageBytes + kMsOffQm;
                const uint32_t abase = tmem_base + kMsAccCols + ar * kMsASlotCols + t * (kMsSub * 32u);
                const uint32_t blo0 = ((qbase & 0x3FFFFu) >> 4) | (1u << 16);
                if (ptx::elect_one_sync()) {
#pragma unroll
                    for (int sub = 0; sub < kMsSub; ++sub) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_i8_ts(d, abase + sub * 32 + k * 8, blo0 + ((sub * kQmChunkBytes + 32 * k) >> 4), kDescHiSw128,
                                       kIdesc32, (sub | k) ? 1u : (c ? 1u : 0u));
                        }
                    }
                    ptx::umma_commit(aempty_bar(ar, t));
                    ptx::umma_commit(empty_bar(stage));
                    if (c == kMsStagesPerTile - 1) ptx::umma_commit(tfull_bar(buf, t));
                }
                __syncwarp();
                if (++stage == kMsStages) { stage = 0; phase ^= 1u; }
                if (++ar == kMsARing) { ar = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= kMsExpWarp0) {
        // ------------------------------------------------------------------ expanders: packed bits -> TMEM A operand
        const int set = (warp - kMsExpWarp0) / (4 * kMsTiles);       // which stages (g % kMsExpSets) this warp expands
        const int t = ((warp - kMsExpWarp0) >> 2) % kMsTiles;       // row tile of this warp
        const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        int stage = 0, ar = 0;
        uint32_t phase = 0, aphase = 0, g = 0;
        for (uint32_t pair = pair0; pair < pair_end; pair += pair_step) {
            for (int c = 0; c < kMsStagesPerTile; ++c, ++g) {
                if ((int)(g % kMsExpSets) != set) {       // the other set's stage: only keep the ring state in step
                    if (++stage == kMsStages) { stage = 0; phase ^= 1u; }
                    if (++ar == kMsARing) { ar = 0; aphase ^= 1u; }
                    continue;
                }
                ptx::mbar_wait(full_bar(stage), phase, p.error, kWsExpFull);
                ptx::mbar_wait(aempty_bar(ar, t), aphase ^ 1u, p.error, kWsExpA);
                ptx::tc_fence_after();
                const uint8_t* pk = base_ptr + stage * kMsStageBytes + t * kMsPkBytes;
                const uint32_t abase = tmem_base + lane_addr + kMsAccCols + ar * kMsASlotCols + t * (kMsSub * 32u);
#pragma unroll
                for (int sub = 0; sub < kMsSub; ++sub) {
                    const uint4 x = *reinterpret_cast<const uint4*>(pk + sub * kMaskChunkBytes + row * 16);
                    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
                    uint32_t v[32];
#pragma unroll
                    for (int w = 0; w < 4; ++w)
#pragma unroll
                        for (int b = 0; b < 8; ++b) v[8 * w + b] = xs[w] & (0x01010101u << b);
                    tmem_st32(abase + sub * 32, v);
                }
                tmem_wait_st();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(afull_bar(ar, t));
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
        for (uint32_t pair = pair0; pair < pair_end; pair += pair_step, ++it) {
            const uint32_t buf = it & 1u;
#pragma unroll 1
            for (int t = 0; t < kMsTiles; ++t) {
                ptx::mbar_wait(tfull_bar(buf, t), (it >> 1) & 1u, p.error, kWsEpilogue);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (kMsTiles * 32u) + t * 32u;
                const int64_t trow0 = ((int64_t)pair * kMsTiles + t) * kTileRows;
                int64_t lo = (int64_t)p.row_begin - trow0, hi = (int64_t)p.row_end - trow0;
                const int r0 = (int)(lo < 0 ? 0 : (lo > kTileRows ? kTileRows : lo));
                const int r1 = (int)(hi < 0 ? 0 : (hi > kTileRows ? kTileRows : hi));
                const int64_t tile_off = (trow0 - (int64_t)p.row_begin) * kOutRowBytes;
                uint32_t a[32];
                ptx::tmem_ld32(taddr, a);
                ptx::tmem_wait_ld();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(buf, t));
                const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(p.den_out) + tile_off) & 15);
                uint8_t* st = out_stage_ptr + shift + row * kOutRowBytes;
#pragma unroll
                for (int j = 0; j < IRIS_ROTATIONS; ++j) *reinterpret_cast<uint16_t*>(st + 2 * j) = (uint16_t)(a[j] >> 7);
                ptx::named_bar_sync(1, 128);
                if (r1 > r0)
                    copy_out_rows(out_stage_ptr, reinterpret_cast<uint8_t*>(p.den_out) + tile_off - shift,
                                (int)shift + r0 * kOutRowBytes, (int)shift + r1 * kOutRowBytes, row);
                ptx::named_bar_sync(1, 128);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kMsIssuerWarp0) ptx::tmem_dealloc(tmem_base, kMsTmemCols);
}

cudaError_t launch_mask_scan(const ScanParams& p, int num_sms, cudaStream_t stream) {
    static std::atomic<bool> configured[64];    // per device: opt-in shared memory size set
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(mask_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMsSmemBytes);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    if (p.tile_end <= p.tile_begin) return cudaSuccess;
    const uint32_t pairs = (p.tile_end + kMsTiles - 1) / kMsTiles - p.tile_begin / kMsTiles;
    const uint32_t grid = pairs < (uint32_t)num_sms ? pairs : (uint32_t)num_sms;
    mask_scan_kernel<<<grid, kMsThreads, kMsSmemBytes, stream>>>(p);
    count_launch_external();
    return cudaGetLastError();
}

}  // namespace iris
