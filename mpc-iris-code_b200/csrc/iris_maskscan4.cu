// Denominators-only scan with 4-BIT operands (BASELINE config 3): popcount(rot(qmask, j-15) & dbmask_i) for ONE query
// mask over the resident masks, reference src/lib.rs:69-79 -> src/arch/generic.rs:4-9.
//
// Same organisation as iris_maskscan.cu (packed bits -> expander warps -> A operand in TENSOR MEMORY -> TS-form
// UMMA), but the operands are e2m1 nibbles and the instruction is tcgen05.mma kind::mxf4 (block-scaled, K = 64):
// an N = 32 UMMA costs ~30 cycles whatever it computes, so covering 64 mask bits per instruction
// instead of 32 halves the tensor-side time, and the expanders store half the bytes with 5 instead of 8 logic
// operations per 32 bits.
//
// Arithmetic (exact): a database bit becomes the nibble  x & (1 << t)  of its 32-bit word, i.e. e2m1 0.5, 1.0 or 2.0
// for t = 0, 1, 2; bit 3 of a nibble would be the sign, so those bits are shifted down one place (2.0).  The query
// operand holds 2.0, 1.0, 0.5, 0.5 for t = 0..3 where the rotated query bit is set, so every coincidence adds exactly
// 1.0 to an f32 accumulator (sums <= 12 800 are exact).  All block scale factors are UE8M0 1.0 (0x7F): one TMEM
// region filled with 0x7F bytes serves as scale_A and scale_B whatever their layout.
//
// K order: output word o = 4 * w + t of a row's 256-bit stage slice (w = input word 0..7) holds, in nibble j, the bit
// 32 * w + 4 * j + t; prep_mask_query_fp4_kernel builds the query operand in the same order.
//
// Per stage: 256 mask bits of 2 x 128 rows = 2 x 4 KiB of packed database + 4 KiB of query operand.
//   producer (warp 4)     : 3 bulk copies into a 12-deep smem ring
//   expanders (warps 7-22): 2 x LDS.128 -> 40 logic ops -> 1 x tcgen05.st.32x32b.x32 into a 5-deep TMEM ring;
//                           two sets of 8 warps alternate stages
//   issuers (warps 5, 6)  : 4 x tcgen05.mma kind::mxf4 (A = TMEM, B = smem) per stage, one warp per tile
//   epilogue (warps 0-3)  : tcgen05.ld, f32 -> u16, 62-byte rows; accumulators double-buffered
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "iris_epilogue.cuh"
#include "iris_kernels.cuh"
#include "iris_ptx.cuh"

namespace iris {

void count_launch_external();

constexpr int kM4Tiles = 2;                                       // row tiles per CTA
constexpr int kM4StageBits = 256;
constexpr int kM4StagesPerTile = IRIS_BITS / kM4StageBits;        // 50
constexpr int kM4PkBytes = kM4StageBits / 8 * kTileRows;          // 4 KiB of packed bits per tile and stage
constexpr int kM4QBytes = kQm4StageBytes;                         // 4 KiB: 32 rotations x 256 nibbles
constexpr int kM4OffQ = kM4Tiles * kM4PkBytes;
constexpr int kM4StageBytes = kM4OffQ + kM4QBytes;                // 12 KiB
constexpr int kM4ARing = 5;                                       // TMEM A slots of 2 tiles x 32 columns
constexpr int kM4OutStageBytes = 8192;
constexpr int kM4BarBytes = 1024;
constexpr int m4_smem_bytes(int stages) { return 1024 + stages * kM4StageBytes + kM4OutStageBytes + kM4BarBytes; }
constexpr int kM4IssuerWarp0 = 5;                                 // issuer warps: [k-parity][tile]
constexpr int m4_exp_warp0(int iss) { return kM4IssuerWarp0 + iss * kM4Tiles; }   // expander warps: [set][tile][TMEM lane quadrant]
constexpr int m4_threads(int sets, int iss) { return (m4_exp_warp0(iss) + sets * 4 * kM4Tiles) * 32; }   // 736 for 2 sets, 1 issuer per tile
constexpr uint32_t kM4AccCols = 2 * kM4Tiles * 32;                // [buffer][tile] x 32 f32 columns
constexpr uint32_t kM4SfCol = kM4AccCols;                         // 32 columns of 0x7F scale-factor bytes
constexpr uint32_t kM4ACol = kM4SfCol + 32;
constexpr uint32_t kM4ASlotCols = kM4Tiles * 32;                  // 64
constexpr uint32_t kM4TmemCols = 512;
static_assert(IRIS_BITS % kM4StageBits == 0, "stages must tile the K dimension");
static_assert(kM4StageBits == 2 * 8 * kMaskChunkBytes / kTileRows, "a stage is two 128-bit mask chunks");
static_assert(kM4ACol + kM4ARing * kM4ASlotCols <= kM4TmemCols, "TMEM budget");
static_assert(kM4StageBytes % 1024 == 0 && kM4OffQ % 1024 == 0, "operand tiles must stay 1024-byte aligned");

enum M4Watchdog { kW4Producer = 501, kW4MmaFull = 502, kW4MmaA = 503, kW4MmaTmem = 504, kW4ExpFull = 505, kW4ExpA = 506, kW4Epilogue = 507 };

__device__ __forceinline__ void tmem_st32_m4(uint32_t taddr, const uint32_t (&v)[32]) {
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
// D[tmem] (+)= A[tmem] . B[smem], e2m1 x e2m1 -> f32, one UE8M0 scale per 32 elements of K (all 1.0 here).
__device__ __forceinline__ void umma_mxf4_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                             uint32_t idesc, uint32_t sf_tmem, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], bd, %4, [%5], [%5], p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(sf_tmem), "r"(accumulate)
        : "memory");
}
// Instruction descriptor, kind::mxf4 (cute::UMMA::InstrDescriptorBlockScaled): A = B = E2M1 (1), K-major both,
// N >> 3 at bit 17, scale format UE8M0 (1) at bit 23, M >> 4 at bit 24, scale-factor ids 0, K = 64.
constexpr uint32_t kM4Idesc = (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
constexpr uint32_t kM4DescHiSw128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B

// p.qm is the 4-bit operand image here (kQm4Bytes); p.tile_begin / p.tile_end as in mask_scan_kernel.
// kSets: expander warp sets taking alternate stages.  kIss: issuing warps per row tile; with two, each takes every
// second stage into its own accumulator (the epilogue adds them) and the accumulators are single-buffered.
// kFlags (A/B switches): 2 / 4 = timing only, no expansion / no UMMAs (wrong results); 32 = expand into registers
// before waiting for the TMEM slot; 256 = per-role wait-time profile of CTA 0 (device printf).
template <int kSets, int kIss, int kFlags, int kStages>
__global__ void __launch_bounds__(m4_threads(kSets, kIss), 1) mask_scan_fp4_kernel(const ScanParams p) {
    static_assert(kSets <= kM4ARing && kSets <= kStages && kIss <= 2, "ring positions advance with at most one wrap");
    static_assert(m4_smem_bytes(kStages) <= 232448, "exceeds 227 KiB of shared memory");
    static_assert(m4_threads(kSets, kIss) <= 1024, "too many warps");
    static_assert(kM4StagesPerTile % kIss == 0, "every issuer takes the same number of stages per tile");
    constexpr int kExpWarp0 = m4_exp_warp0(kIss);
    constexpr int kAccBufs = kIss == 1 ? 2 : 1;                     // accumulator sets: [buffer][tile][issuer] x 32 columns
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw_addr);
    uint8_t* const out_stage_ptr = base_ptr + kStages * kM4StageBytes;
    const uint32_t bars = base + kStages * kM4StageBytes + kM4OutStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };                                   // stage landed (tx)
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };                    // 8 expander warps + 2 issuer commits
    // The TMEM A ring has kM4ARing slots but 2 * kM4ARing barriers per direction: stage g uses slot g % kM4ARing and
    // barrier j = g % (2 * kM4ARing), so that each barrier always pairs the same expander set with the same issuing
    // warp (stages of one parity) and nobody ever waits on a phase two uses ahead.
    auto afull_bar = [&](int j, int t) { return bars + 8u * (2 * kStages + 2 * j + t); };                 // 4 expander warps of tile t
    auto aempty_bar = [&](int j, int t) { return bars + 8u * (2 * kStages + 4 * kM4ARing + 2 * j + t); }; // one commit
    auto tfull_bar = [&](int b, int t) { return bars + 8u * (2 * kStages + 8 * kM4ARing + 2 * b + t); };
    auto tempty_bar = [&](int b, int t) { return bars + 8u * (2 * kStages + 8 * kM4ARing + 4 + 2 * b + t); };
    constexpr int kNumBars = 2 * kStages + 8 * kM4ARing + 8;
    constexpr int kABars = 2 * kM4ARing;
    static_assert(8 * (kNumBars + 1) <= kM4BarBytes, "barrier table");
    const uint32_t tmem_slot = bars + 8u * kNumBars;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(out_stage_ptr + kM4OutStageBytes + 8 * kNumBars);

    const int warp = ptx::warp_idx_sync();      // warp-uniform role index
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 4 * kM4Tiles + kM4Tiles);     // expander warps + one commit per tile
        }
        for (int a = 0; a < kABars; ++a)
            for (int t = 0; t < kM4Tiles; ++t) {
                ptx::mbar_init(afull_bar(a, t), 4);
                ptx::mbar_init(aempty_bar(a, t), 1);
            }
        for (int b = 0; b < 2; ++b)
            for (int t = 0; t < kM4Tiles; ++t) {
                ptx::mbar_init(tfull_bar(b, t), kIss);
                ptx::mbar_init(tempty_bar(b, t), 4);
            }
        ptx::fence_mbar_init();
    }
    if (warp == kM4IssuerWarp0) ptx::tmem_alloc(tmem_slot, kM4TmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (warp < 4) {                             // scale factors: 1.0 everywhere (each warp fills its lane quadrant)
        uint32_t ones[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) ones[i] = 0x7F7F7F7Fu;
        tmem_st32_m4(tmem_base + ((uint32_t)(warp * 32) << 16) + kM4SfCol, ones);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    // kFlags & 256: per-role wait-time profile of CTA 0 (device printf at the end)
    constexpr bool kProf = (kFlags & 256) != 0;
    long long prof[4] = {0, 0, 0, 0};
    const long long prof_t0 = kProf ? clock64() : 0;
#define M4_TIMED(slot, stmt) do { if (kProf) { const long long t_ = clock64(); stmt; prof[slot] += clock64() - t_; } else { stmt; } } while (0)

    const uint32_t pair_begin = p.tile_begin / kM4Tiles;
    const uint32_t pair_end = (p.tile_end + kM4Tiles - 1) / kM4Tiles;
    const uint32_t pair0 = pair_begin + blockIdx.x;
    const uint32_t pair_step = gridDim.x;

    if (warp == 4) {
        // ------------------------------------------------------------------ producer
        const uint64_t pol_stream = ptx::policy_evict_first();
        const uint64_t pol_keep = ptx::policy_evict_last();
        int stage = 0;
        uint32_t phase = 0;
        for (uint32_t pair = pair0; pair < pair_end; pair += pair_step) {
            const uint8_t* mk = p.masks + (size_t)pair * kM4Tiles * kMaskTileBytes;
            for (int c = 0; c < kM4StagesPerTile; ++c) {
                M4_TIMED(0, ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.error, kW4Producer));
                const uint32_t sbase = base + stage * kM4StageBytes;
                const uint32_t fb = full_bar(stage);
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(fb, kM4StageBytes);
#pragma unroll
                    for (int t = 0; t < kM4Tiles; ++t)
                        ptx::bulk_g2s_hint(sbase + t * kM4PkBytes, mk + (size_t)t * kMaskTileBytes + (size_t)c * kM4PkBytes,
                                           kM4PkBytes, fb, pol_stream);
                    ptx::bulk_g2s_hint(sbase + kM4OffQ, p.qm + (size_t)c * kM4QBytes, kM4QBytes, fb, pol_keep);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= kM4IssuerWarp0 && warp < kExpWarp0) {
        // ------------------------------------------------------------------ UMMA issuers
        // warp (par, t) issues the stages g = par (mod kIss) of row tile t into its own accumulator
        const int t = (warp - kM4IssuerWarp0) % kM4Tiles;
        const int par = (warp - kM4IssuerWarp0) / kM4Tiles;
        const uint32_t sf = tmem_base + kM4SfCol;
        const uint32_t my_pairs = pair0 < pair_end ? (pair_end - pair0 + pair_step - 1) / pair_step : 0;
        const uint32_t total = my_pairs * kM4StagesPerTile;
        int stage = par, aj = par, c = par;         // ring positions and stage-within-tile of g
        uint32_t phase = 0, aphase = 0, it = 0, d = 0;
        for (uint32_t g = par; g < total; g += kIss) {
            if (c == par) {                          // first stage of a tile: the accumulator must have been drained
                const uint32_t buf = kAccBufs == 2 ? (it & 1u) : 0u;
                const uint32_t par_t = kAccBufs == 2 ? ((it >> 1) & 1u) : (it & 1u);
                M4_TIMED(2, ptx::mbar_wait(tempty_bar(buf, t), par_t ^ 1u, p.error, kW4MmaTmem));
                ptx::tc_fence_after();
                d = tmem_base + ((buf * kM4Tiles + t) * kIss + par) * 32u;
            }
            M4_TIMED(0, ptx::mbar_wait(full_bar(stage), phase, p.error, kW4MmaFull));
            M4_TIMED(1, ptx::mbar_wait(afull_bar(aj, t), aphase, p.error, kW4MmaA));
            const int ar = aj >= kM4ARing ? aj - kM4ARing : aj;
            ptx::tc_fence_after();
            const uint32_t qbase = base + stage * kM4StageBytes + kM4OffQ;
            const uint32_t abase = tmem_base + kM4ACol + ar * kM4ASlotCols + t * 32u;
            const uint32_t blo0 = ((qbase & 0x3FFFFu) >> 4) | (1u << 16);
            const bool last = c + kIss >= kM4StagesPerTile;
            if (ptx::elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < ((kFlags & 4) ? 0 : 4); ++k)      // 64 nibbles = 32 bytes = 8 TMEM columns per step
                    umma_mxf4_ts(d, abase + k * 8, blo0 + ((32 * k) >> 4), kM4DescHiSw128, kM4Idesc, sf,
                                 k ? 1u : (c != par ? 1u : 0u));
                ptx::umma_commit(aempty_bar(aj, t));
                ptx::umma_commit(empty_bar(stage));
                if (last) ptx::umma_commit(tfull_bar(kAccBufs == 2 ? (it & 1u) : 0u, t));
            }
            __syncwarp();
            stage += kIss;
            if (stage >= kStages) { stage -= kStages; phase ^= 1u; }
            aj += kIss;
            if (aj >= kABars) { aj -= kABars; aphase ^= 1u; }
            c += kIss;
            if (c >= kM4StagesPerTile) { c -= kM4StagesPerTile; ++it; }
        }
    } else if (warp >= kExpWarp0) {
        // ------------------------------------------------------------------ expanders: packed bits -> e2m1 A operand
        const int set = (warp - kExpWarp0) / (4 * kM4Tiles);      // this warp expands the stages g = set (mod kSets)
        const int t = ((warp - kExpWarp0) >> 2) % kM4Tiles;       // row tile of this warp
        const int quad = warp & 3;                                  // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const uint32_t my_pairs = pair0 < pair_end ? (pair_end - pair0 + pair_step - 1) / pair_step : 0;
        const uint32_t total = my_pairs * kM4StagesPerTile;         // stages of this CTA, all pairs
        int stage = set, aj = set;
        uint32_t phase = 0, aphase = 0;
        {
            for (uint32_t g = set; g < total; g += kSets) {
                // the slot was last used by stage g - kM4ARing: its commit went to barrier (aj + kM4ARing) % kABars
                const int ar = aj >= kM4ARing ? aj - kM4ARing : aj;
                const int jw = aj >= kM4ARing ? aj - kM4ARing : aj + kM4ARing;
                const uint32_t wpar = aj >= kM4ARing ? aphase : aphase ^ 1u;
                M4_TIMED(0, ptx::mbar_wait(full_bar(stage), phase, p.error, kW4ExpFull));
                if (!(kFlags & 32)) {
                    M4_TIMED(1, ptx::mbar_wait(aempty_bar(jw, t), wpar, p.error, kW4ExpA));
                    ptx::tc_fence_after();
                }
                uint32_t v[32];
                if (!(kFlags & 2)) {
                    const uint8_t* pk = base_ptr + stage * kM4StageBytes + t * kM4PkBytes;
                    const uint4 x0 = *reinterpret_cast<const uint4*>(pk + row * 16);                       // bits 0..127
                    const uint4 x1 = *reinterpret_cast<const uint4*>(pk + kMaskChunkBytes + row * 16);     // bits 128..255
                    const uint32_t xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        v[4 * w + 0] = xs[w] & 0x11111111u;             // 0.5
                        v[4 * w + 1] = xs[w] & 0x22222222u;             // 1.0
                        v[4 * w + 2] = xs[w] & 0x44444444u;             // 2.0
                        v[4 * w + 3] = (xs[w] >> 1) & 0x44444444u;      // bit 3 of each nibble, moved off the sign: 2.0
                    }
                }
                if (kFlags & 32) {
                    // the expanded words are in registers: hand the packed bytes back, then wait for the TMEM slot
#pragma unroll
                    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));     // keep the logic ops above the wait
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(empty_bar(stage));
                    M4_TIMED(1, ptx::mbar_wait(aempty_bar(jw, t), wpar, p.error, kW4ExpA));
                    ptx::tc_fence_after();
                }
                if (!(kFlags & 2)) {
                    tmem_st32_m4(tmem_base + lane_addr + kM4ACol + ar * kM4ASlotCols + t * 32u, v);
                    M4_TIMED(2, asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"));
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(afull_bar(aj, t));
                    if (!(kFlags & 32)) ptx::mbar_arrive(empty_bar(stage));   // this warp no longer needs the packed bytes
                }
                stage += kSets;
                if (stage >= kStages) { stage -= kStages; phase ^= 1u; }
                aj += kSets;
                if (aj >= kABars) { aj -= kABars; aphase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0..3)
        const int row = threadIdx.x;
        uint32_t it = 0;
        for (uint32_t pair = pair0; pair < pair_end; pair += pair_step, ++it) {
            const uint32_t buf = kAccBufs == 2 ? (it & 1u) : 0u;
            const uint32_t par_t = kAccBufs == 2 ? ((it >> 1) & 1u) : (it & 1u);
#pragma unroll 1
            for (int t = 0; t < kM4Tiles; ++t) {
                M4_TIMED(0, ptx::mbar_wait(tfull_bar(buf, t), par_t, p.error, kW4Epilogue));
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (buf * kM4Tiles + t) * kIss * 32u;
                const int64_t trow0 = ((int64_t)pair * kM4Tiles + t) * kTileRows;
                int64_t lo = (int64_t)p.row_begin - trow0, hi = (int64_t)p.row_end - trow0;
                const int r0 = (int)(lo < 0 ? 0 : (lo > kTileRows ? kTileRows : lo));
                const int r1 = (int)(hi < 0 ? 0 : (hi > kTileRows ? kTileRows : hi));
                const int64_t tile_off = (trow0 - (int64_t)p.row_begin) * kOutRowBytes;
                uint32_t a[32];
                ptx::tmem_ld32(taddr, a);
                ptx::tmem_wait_ld();
                if (kIss == 2) {                    // the other issuer's partial sums (f32, exact)
                    uint32_t b2[32];
                    ptx::tmem_ld32(taddr + 32u, b2);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = __float_as_uint(__uint_as_float(a[j]) + __uint_as_float(b2[j]));
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(buf, t));
                const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(p.den_out) + tile_off) & 15);
                uint8_t* st = out_stage_ptr + shift + row * kOutRowBytes;
#pragma unroll
                for (int j = 0; j < IRIS_ROTATIONS; ++j)
                    *reinterpret_cast<uint16_t*>(st + 2 * j) = (uint16_t)__float2uint_rn(__uint_as_float(a[j]));
                ptx::named_bar_sync(1, 128);
                if (r1 > r0)
                    copy_out_rows(out_stage_ptr, reinterpret_cast<uint8_t*>(p.den_out) + tile_off - shift,
                                  (int)shift + r0 * kOutRowBytes, (int)shift + r1 * kOutRowBytes, row);
                ptx::named_bar_sync(1, 128);
            }
        }
    }

    if (kProf && blockIdx.x == 0 && lane == 0)
        printf("m4prof warp %2d total %lld w0 %lld w1 %lld w2 %lld\n", warp, clock64() - prof_t0, prof[0], prof[1], prof[2]);
#undef M4_TIMED
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kM4IssuerWarp0) ptx::tmem_dealloc(tmem_base, kM4TmemCols);
}

template <int kSets, int kIss, int kFlags, int kStages>
static cudaError_t launch_m4_t(const ScanParams& p, int num_sms, cudaStream_t stream) {
    static std::atomic<bool> configured[64];    // per device: opt-in shared memory size set
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(mask_scan_fp4_kernel<kSets, kIss, kFlags, kStages>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, m4_smem_bytes(kStages));
        if (e != cudaSuccess) return e;
        if (dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    if (p.tile_end <= p.tile_begin) return cudaSuccess;
    const uint32_t pairs = (p.tile_end + kM4Tiles - 1) / kM4Tiles - p.tile_begin / kM4Tiles;
    const uint32_t grid = pairs < (uint32_t)num_sms ? pairs : (uint32_t)num_sms;
    mask_scan_fp4_kernel<kSets, kIss, kFlags, kStages>
        <<<grid, m4_threads(kSets, kIss), m4_smem_bytes(kStages), stream>>>(p);
    count_launch_external();
    return cudaGetLastError();
}

cudaError_t launch_mask_scan_fp4(const ScanParams& p, int num_sms, cudaStream_t stream) {
    // IRIS_M4_VARIANT = "<sets><issuers per tile><flags, 3 digits>" selects an A/B variant (diagnostics).
    static const int variant = [] {
        const char* e = getenv("IRIS_M4_VARIANT");
        return e ? atoi(e) : 21000;
    }();
    switch (variant) {
        case 22000: return launch_m4_t<2, 2, 0, 12>(p, num_sms, stream);
        case 22032: return launch_m4_t<2, 2, 32, 12>(p, num_sms, stream);
        case 21032: return launch_m4_t<2, 1, 32, 12>(p, num_sms, stream);
        case 22256: return launch_m4_t<2, 2, 256, 12>(p, num_sms, stream);
        case 22288: return launch_m4_t<2, 2, 288, 12>(p, num_sms, stream);
        case 22002: return launch_m4_t<2, 2, 2, 12>(p, num_sms, stream);
        case 22004: return launch_m4_t<2, 2, 4, 12>(p, num_sms, stream);
        default: return launch_m4_t<2, 1, 0, 12>(p, num_sms, stream);
    }
}

// Query operand image for mask_scan_fp4_kernel: [stage s < 50][rotation slot r < 32][128 B, SWIZZLE_128B]; the 16-byte
// chunk `ch` of a row holds the output words 4 * ch + t (t = 0..3) of input word ch; nibble j of word t stands for
// the bit 256 * s + 32 * ch + 4 * j + t of rot(qmask, r - 15) and holds 2.0, 1.0, 0.5, 0.5 (e2m1) for t = 0..3.
__global__ void prep_mask_query_fp4_kernel(const uint8_t* __restrict__ qmask, uint8_t* __restrict__ qm4) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (s, r, ch)
    if (idx >= kM4StagesPerTile * 32 * 8) return;
    const int ch = idx & 7, r = (idx >> 3) & 31, s = idx >> 8;
    uint32_t out[4] = {0, 0, 0, 0};
    if (r < IRIS_ROTATIONS) {
        const int rot = r - 15;
        for (int t = 0; t < 4; ++t) {
            const uint32_t weight = t == 0 ? 4u : (t == 1 ? 2u : 1u);     // e2m1 codes of 2.0, 1.0, 0.5
            for (int j = 0; j < 8; ++j) {
                const int bit_index = kM4StageBits * s + 32 * ch + 4 * j + t;
                const int row = bit_index / IRIS_COLS, col = bit_index % IRIS_COLS;
                const int src = row * IRIS_COLS + (col - rot + IRIS_COLS) % IRIS_COLS;
                const uint32_t bit = (qmask[src >> 3] >> (src & 7)) & 1u;
                out[t] |= (bit * weight) << (4 * j);
            }
        }
    }
    const size_t off = (size_t)s * kQm4StageBytes + r * 128 + ((ch ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(qm4 + off) = make_uint4(out[0], out[1], out[2], out[3]);
}
cudaError_t launch_prep_mask_query_fp4(const uint8_t* d_qmask, uint8_t* d_qm4, cudaStream_t stream) {
    prep_mask_query_fp4_kernel<<<(kM4StagesPerTile * 32 * 8 + 255) / 256, 256, 0, stream>>>(d_qmask, d_qm4);
    count_launch_external();
    return cudaGetLastError();
}

}  // namespace iris
