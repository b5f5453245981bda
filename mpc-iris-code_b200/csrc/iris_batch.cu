// Batched queries x rotations vs the resident database as a dense int8 GEMM (BASELINE config 4).
//
//   D[rows x (8 queries x 32 rotations)] = DB[rows x 12800] . Q[(8 x 32) x 12800]^T   per limb product
//
// Same arithmetic as the scan (iris_kernels.cu): S00 = d_lo.q_lo, S1 = d_lo.q_hi + d_hi.q_lo (two UMMAs
// accumulating into the same TMEM columns), dist = (S00 + (S1 << 8)) & 0xFFFF.  When every query of a
// batch is representable as a signed byte (always true for encode() output {0,1,0xFFFF}, src/lib.rs:16-26)
// the q_lo plane IS the s8 value and only two products are needed: S0 = d_lo.q_s, S1 = d_hi.q_s.
//
// One CTA PAIR (tcgen05 cta_group::2, UMMA M = 256, N = 256) works on 256 database rows x 8 queries:
// each CTA streams its own 128-row share tile (32 KiB per K-chunk) and HALF of the query operand
// (4 queries: 16 KiB q_lo [+ 16 KiB q_hi]).  The 2 x 256 s32 accumulator columns fill TMEM (512 columns).
// Query groups are the fastest-varying tile index so the clusters working on one row tile hit it in L2
// and HBM is read once per batch.
//
// Cross-CTA protocol (only async-proxy data crosses CTAs, so all mbarrier traffic is cta-scope/cheap):
//   full[s]   local   : this CTA's bulk copies of stage s landed (expect-tx)
//   ready[s]  leader  : count 2 -- each CTA's relay thread arrives once its full[s] completed
//   empty[s]  both    : tcgen05.commit multicast from the leader when the UMMAs reading stage s finished
//   tfull     both    : tcgen05.commit multicast after the last K-chunk of a tile
//   tempty    leader  : count 8 -- one arrive per epilogue warp of both CTAs once TMEM has been drained
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>

#include "iris_epilogue.cuh"
#include "iris_kernels.cuh"
#include "iris_ptx.cuh"

namespace iris {

void count_launch_external();

constexpr int kBatchQTile = 8;                         // queries per cluster tile
constexpr int kBatchBBytes = 4 * kQTileBytes;          // 4 queries x 4 KiB per plane per CTA
constexpr int kBatchOutStageBytes = 8192;
constexpr int kBatchThreads = 192;                     // warps 0-3 epilogue, 4 producer, 5 relay + UMMA issuer

template <bool SIGNED_Q>
struct BatchCfg {
    static constexpr int kStageBytes = kShareChunkBytes + (SIGNED_Q ? 1 : 2) * kBatchBBytes;   // 48 / 64 KiB
    static constexpr int kStages = SIGNED_Q ? 4 : 3;
    static constexpr int kOffAlo = 0;
    static constexpr int kOffAhi = kPlaneTileBytes;
    static constexpr int kOffBlo = kShareChunkBytes;
    static constexpr int kOffBhi = kShareChunkBytes + kBatchBBytes;
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * kBatchOutStageBytes + 512;
    static_assert(kSmemBytes <= 232448, "exceeds 227 KiB of shared memory");
};

enum BatchWatchdog { kWbProducer = 201, kWbFull = 202, kWbReady = 203, kWbTmemEmpty = 204, kWbEpilogue = 205 };


template <bool SIGNED_Q>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBatchThreads, 1)
    batch_distances_kernel(const BatchParams p) {
    using Cfg = BatchCfg<SIGNED_Q>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw_addr);
    uint8_t* const out_stage_ptr = base_ptr + kStages * Cfg::kStageBytes;
    const uint32_t bars = base + kStages * Cfg::kStageBytes + 2 * kBatchOutStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto ready_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
    const uint32_t tfull_bar = bars + 8u * (3 * kStages);
    const uint32_t tempty_bar = bars + 8u * (3 * kStages + 1);
    const uint32_t tmem_slot = bars + 8u * (3 * kStages + 2);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(out_stage_ptr + 2 * kBatchOutStageBytes + 8 * (3 * kStages + 2));

    const int warp = ptx::warp_idx_sync();                   // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const uint32_t rank = __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0);   // 0 = leader (issues the UMMAs)
    const uint32_t cluster_id = blockIdx.x >> 1;
    const uint32_t num_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
            ptx::mbar_init(ready_bar(s), 2);
        }
        ptx::mbar_init(tfull_bar, 1);
        ptx::mbar_init(tempty_bar, 8);
        ptx::fence_mbar_init();
    }
    if (warp == 5) ptx::tmem_alloc_2cta(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                                     // barriers of both CTAs are initialised
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const uint32_t num_groups = (p.num_queries + kBatchQTile - 1) / kBatchQTile;
    const uint32_t num_tiles = (p.pair_end - p.pair_begin) * num_groups;

    if (warp == 4) {
        // ------------------------------------------------------------------ producer (each CTA loads its own half)
        const uint64_t pol_keep = ptx::policy_evict_last();
        int stage = 0;
        uint32_t phase = 0;
        uint32_t wave = 0;
        bool pacing = p.wave_sync != nullptr && rank == 0;
        for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters, ++wave) {
            // Wave pacing: a row tile is shared by the num_groups clusters that work on it in the same wave, but only
            // through L2, and nothing else keeps those clusters in step -- they drift apart by whole tiles and the
            // tile is then fetched from DRAM again.  The leaders' producers therefore start a wave together (the consumers
            // follow within the depth of the stage ring, which also absorbs the wait).  The wait is bounded: a
            // cluster that is not resident yet (SMs taken by another kernel) ends the pacing, never the run.
            if (pacing) {
                if (ptx::elect_one_sync()) {
                    atomicAdd(p.wave_sync, 1u);                                   // this cluster has started wave `wave`
                    // everybody that has a tile in this wave has started it
                    const uint32_t left = num_tiles - wave * num_clusters;
                    const uint32_t want = num_clusters * wave + (left < num_clusters ? left : num_clusters);
                    const uint64_t t0 = ptx::globaltimer_ns();
                    for (;;) {
                        uint32_t seen, quit;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.wave_sync) : "memory");
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(quit) : "l"(p.wave_sync + 1) : "memory");
                        if (seen >= want || quit) break;
                        if (ptx::globaltimer_ns() - t0 > 200000ull) {
                            atomicExch(p.wave_sync + 1, 1u);                      // give up, for everybody
                            break;
                        }
                        __nanosleep(100);
                    }
                }
                __syncwarp();
                uint32_t quit;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(quit) : "l"(p.wave_sync + 1) : "memory");
                pacing = quit == 0;
            }
            const uint32_t pair = p.pair_begin + t / num_groups;
            const uint32_t group = t % num_groups;
            const uint8_t* sh = p.shares + (size_t)(2 * pair + rank) * kShareTileBytes;
            const uint8_t* q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t qi = group * kBatchQTile + 4 * rank + i;
                q[i] = p.qd[qi < p.num_queries ? qi : p.num_queries - 1];
            }
            for (int c = 0; c < kChunks; ++c) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.error, kWbProducer);
                const uint32_t sbase = base + stage * Cfg::kStageBytes;
                const uint32_t fb = full_bar(stage);
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
                    ptx::bulk_g2s(sbase + Cfg::kOffAlo, sh + (size_t)c * kShareChunkBytes, kShareChunkBytes, fb);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ptx::bulk_g2s_hint(sbase + Cfg::kOffBlo + i * kQTileBytes, q[i] + (size_t)c * kQdChunkBytes,
                                           kQTileBytes, fb, pol_keep);
                        if (!SIGNED_Q)
                            ptx::bulk_g2s_hint(sbase + Cfg::kOffBhi + i * kQTileBytes,
                                               q[i] + (size_t)c * kQdChunkBytes + kQTileBytes, kQTileBytes, fb, pol_keep);
                    }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ stage relay (both CTAs) + UMMA issue (leader)
        constexpr uint32_t kIdescU = ptx::umma_idesc_i8_m256(256, false, false);
        constexpr uint32_t kIdescS = ptx::umma_idesc_i8_m256(256, false, true);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
            if (rank == 0) {
                ptx::mbar_wait(tempty_bar, (it & 1u) ^ 1u, p.error, kWbTmemEmpty);
                ptx::tc_fence_after();
            }
            for (int c = 0; c < kChunks; ++c) {
                ptx::mbar_wait(full_bar(stage), phase, p.error, kWbFull);
                const uint32_t ready_leader = ptx::mapa(ready_bar(stage), 0);
                if (ptx::elect_one_sync()) ptx::mbar_arrive_cluster(ready_leader);           // this half landed
                __syncwarp();
                if (rank == 0) {
                    ptx::mbar_wait(ready_bar(stage), phase, p.error, kWbReady);
                    ptx::tc_fence_after();
                    const uint32_t sbase = base + stage * Cfg::kStageBytes;
                    if (ptx::elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < kChunkK / 32; ++k) {
                            const uint32_t acc = (c | k) ? 1u : 0u;
                            const uint64_t alo = ptx::umma_desc_sw128(sbase + Cfg::kOffAlo + 32 * k);
                            const uint64_t ahi = ptx::umma_desc_sw128(sbase + Cfg::kOffAhi + 32 * k);
                            const uint64_t blo = ptx::umma_desc_sw128(sbase + Cfg::kOffBlo + 32 * k);
                            if (SIGNED_Q) {
                                ptx::umma_i8_2cta(tmem_base + 0, alo, blo, kIdescS, acc);
                                ptx::umma_i8_2cta(tmem_base + 256, ahi, blo, kIdescS, acc);
                            } else {
                                const uint64_t bhi = ptx::umma_desc_sw128(sbase + Cfg::kOffBhi + 32 * k);
                                ptx::umma_i8_2cta(tmem_base + 0, alo, blo, kIdescU, acc);
                                ptx::umma_i8_2cta(tmem_base + 256, alo, bhi, kIdescU, acc);
                                ptx::umma_i8_2cta(tmem_base + 256, ahi, blo, kIdescU, 1u);
                            }
                        }
                        ptx::umma_commit_2cta(empty_bar(stage), 3);
                        if (c == kChunks - 1) ptx::umma_commit_2cta(tfull_bar, 3);
                    }
                    __syncwarp();
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0..3 of each CTA)
        const int row = threadIdx.x;
        const uint32_t tempty_leader = ptx::mapa(tempty_bar, 0);
        const uint64_t rows_out = p.row_end - p.row_begin;
        uint32_t it = 0;
        for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
            const uint32_t pair = p.pair_begin + t / num_groups;
            const uint32_t group = t % num_groups;
            ptx::mbar_wait(tfull_bar, it & 1u, p.error, kWbEpilogue);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            const int64_t trow0 = ((int64_t)2 * pair + rank) * kTileRows;
            int64_t lo = (int64_t)p.row_begin - trow0, hi = (int64_t)p.row_end - trow0;
            const int r0 = (int)(lo < 0 ? 0 : (lo > kTileRows ? kTileRows : lo));
            const int r1 = (int)(hi < 0 ? 0 : (hi > kTileRows ? kTileRows : hi));
            const int64_t tile_off = (trow0 - (int64_t)p.row_begin) * kOutRowBytes;
            // The accumulators fill all 512 TMEM columns, so the next tile's UMMAs cannot start before this tile has
            // been read out: the read-out is kept as short as possible.  Phase 1 drains the eight queries into
            // registers, recombining the limbs on the way (two u16 results per register: 128 registers), and releases
            // tensor memory; phase 2 -- staging the 62-byte rows and the coalesced stores -- then runs under the next
            // tile's UMMAs.
            uint32_t packed[kBatchQTile][16];
#pragma unroll
            for (int g = 0; g < kBatchQTile; ++g) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {                  // 16 columns at a time keeps the live registers down
                    uint32_t a[16], b[16];
                    ptx::tmem_ld16(taddr + 32 * g + 16 * h, a);
                    ptx::tmem_ld16(taddr + 256 + 32 * g + 16 * h, b);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t lo = (a[2 * j] + (b[2 * j] << 8)) & 0xFFFFu;
                        const uint32_t hi = (a[2 * j + 1] + (b[2 * j + 1] << 8)) << 16;     // slot 31 (padding) is never stored
                        packed[g][8 * h + j] = lo | hi;
                    }
                }
            }
            // every accumulator column of this warp's lanes has been read: one arrive per warp lets the leader start
            // the next tile's UMMAs
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader);
#pragma unroll
            for (int g = 0; g < kBatchQTile; ++g) {
                const uint32_t qi = group * kBatchQTile + g;
                if (qi < p.num_queries) {             // uniform over the CTA
                    uint8_t* outq = reinterpret_cast<uint8_t*>(p.out + (size_t)qi * rows_out * IRIS_ROTATIONS);
                    const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(outq) + tile_off) & 15);
                    uint8_t* stage_buf = out_stage_ptr + (g & 1) * kBatchOutStageBytes;
                    uint8_t* st = stage_buf + shift + row * kOutRowBytes;
#pragma unroll
                    for (int j = 0; j < IRIS_ROTATIONS; ++j)
                        *reinterpret_cast<uint16_t*>(st + 2 * j) = (uint16_t)(packed[g][j >> 1] >> (16 * (j & 1)));
                    // staging buffers alternate per query: a thread can only write buffer (g&1) again after
                    // passing the barrier of query g+1, which every thread reaches after its copy of query g.
                    ptx::named_bar_sync(1, 128);
                    copy_out_rows(stage_buf, outq + tile_off - shift, (int)shift + r0 * kOutRowBytes,
                                  (int)shift + r1 * kOutRowBytes, row);
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                                     // nobody may leave while the peer can still signal / read it
    if (warp == 5) ptx::tmem_dealloc_2cta(tmem_base, 512);
}

template <bool SIGNED_Q>
static cudaError_t launch_batch_t(const BatchParams& p, int num_sms, cudaStream_t stream) {
    using Cfg = BatchCfg<SIGNED_Q>;
    static std::atomic<bool> configured[64];    // per device: opt-in shared memory size set
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(batch_distances_kernel<SIGNED_Q>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const uint32_t num_groups = (p.num_queries + kBatchQTile - 1) / kBatchQTile;
    const uint32_t tiles = (p.pair_end - p.pair_begin) * num_groups;
    if (tiles == 0) return cudaSuccess;
    uint32_t clusters = (uint32_t)num_sms / 2;
#ifdef IRIS_DIAGNOSTICS
    static const int forced = [] {                      // diagnostics build only: A/B runs of the cluster count
        const char* e = getenv("IRIS_BATCH_CLUSTERS");
        return e ? atoi(e) : 0;
    }();
    if (forced > 0 && (uint32_t)forced < clusters) clusters = (uint32_t)forced;
#endif
    if (tiles < clusters) clusters = tiles;
    batch_distances_kernel<SIGNED_Q><<<2 * clusters, kBatchThreads, Cfg::kSmemBytes, stream>>>(p);
    count_launch_external();
    return cudaGetLastError();
}

cudaError_t launch_batch_distances(const BatchParams& p, bool signed_queries, int num_sms, cudaStream_t stream) {
    if (p.num_queries == 0 || p.num_queries > kMaxBatchQueries) return cudaErrorInvalidValue;
    return signed_queries ? launch_batch_t<true>(p, num_sms, stream) : launch_batch_t<false>(p, num_sms, stream);
}

#ifdef IRIS_DIAGNOSTICS   // superseded by mask_scan_fp4_multi_kernel; kept in the diagnostics build for A/B runs
// =====================================================================================
// batched denominators: popcount(rot(qmask_g, j-15) & dbmask_i) for 16 query masks per cluster tile.
// A = database mask bits expanded to bytes inside the SM (one LOP per 4 bytes, value 2^t, see
// iris_kernels.cu), B = prepared mask operand images (value 2^(7-t)), D >> 7 = popcount.
// 16 queries x 32 rotations = 512 s32 accumulator columns = two N=256 UMMAs per K step.
// =====================================================================================
constexpr int kMaskQTile = 16;
constexpr int kMaskStages = 4;
constexpr int kMaskOffAmx = 0;                                  // 16 KiB expanded operand (written by the SM)
constexpr int kMaskOffB = kPlaneTileBytes;                      // 8 query tiles of 4 KiB (this CTA's half)
constexpr int kMaskOffPk = kPlaneTileBytes + 8 * kQTileBytes;   // 2 KiB packed mask bytes
constexpr int kMaskStageBytes = kMaskOffPk + kMaskChunkBytes;   // 50 KiB
constexpr int kMaskSmemBytes = 1024 + kMaskStages * kMaskStageBytes + 2 * kBatchOutStageBytes + 512;
constexpr int kMaskThreads = 320;                               // + warps 6..9 expanders
static_assert(kMaskStageBytes % 1024 == 0, "operand tiles must stay 1024-byte aligned");
static_assert(kMaskSmemBytes <= 232448, "exceeds 227 KiB of shared memory");

enum MaskWatchdog { kWmProducer = 301, kWmFull = 302, kWmExp = 303, kWmReady = 304, kWmTmemEmpty = 305, kWmEpilogue = 306, kWmExpander = 307 };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMaskThreads, 1)
    batch_denominators_kernel(const BatchMaskParams p) {
    constexpr int kStages = kMaskStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw_addr);
    uint8_t* const out_stage_ptr = base_ptr + kStages * kMaskStageBytes;
    const uint32_t bars = base + kStages * kMaskStageBytes + 2 * kBatchOutStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto ready_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto expd_bar = [&](int s) { return bars + 8u * (3 * kStages + s); };
    const uint32_t tfull_bar = bars + 8u * (4 * kStages);
    const uint32_t tempty_bar = bars + 8u * (4 * kStages + 1);
    const uint32_t tmem_slot = bars + 8u * (4 * kStages + 2);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(out_stage_ptr + 2 * kBatchOutStageBytes + 8 * (4 * kStages + 2));

    const int warp = ptx::warp_idx_sync();
    const int lane = threadIdx.x & 31;
    const uint32_t rank = __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0);
    const uint32_t cluster_id = blockIdx.x >> 1;
    const uint32_t num_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
            ptx::mbar_init(ready_bar(s), 2);
            ptx::mbar_init(expd_bar(s), 128);
        }
        ptx::mbar_init(tfull_bar, 1);
        ptx::mbar_init(tempty_bar, 8);
        ptx::fence_mbar_init();
    }
    if (warp == 5) ptx::tmem_alloc_2cta(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const uint32_t num_groups = (p.num_queries + kMaskQTile - 1) / kMaskQTile;
    const uint32_t num_tiles = (p.pair_end - p.pair_begin) * num_groups;

    if (warp == 4) {
        // ------------------------------------------------------------------ producer
        const uint64_t pol_keep = ptx::policy_evict_last();
        int stage = 0;
        uint32_t phase = 0;
        for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters) {
            const uint32_t pair = p.pair_begin + t / num_groups;
            const uint32_t group = t % num_groups;
            const uint8_t* mk = p.masks + (size_t)(2 * pair + rank) * kMaskTileBytes;
            // this CTA's half of both N=256 operands: queries {0..3, 8..11} (rank 0) or {4..7, 12..15} (rank 1)
            const uint8_t* q[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t qi = group * kMaskQTile + (i < 4 ? 4 * rank + i : 8 + 4 * rank + (i - 4));
                q[i] = p.qm[qi < p.num_queries ? qi : p.num_queries - 1];
            }
            for (int c = 0; c < kChunks; ++c) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.error, kWmProducer);
                const uint32_t sbase = base + stage * kMaskStageBytes;
                const uint32_t fb = full_bar(stage);
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(fb, 8 * kQTileBytes + kMaskChunkBytes);
                    ptx::bulk_g2s(sbase + kMaskOffPk, mk + (size_t)c * kMaskChunkBytes, kMaskChunkBytes, fb);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        ptx::bulk_g2s_hint(sbase + kMaskOffB + i * kQTileBytes, q[i] + (size_t)c * kQmChunkBytes, kQTileBytes,
                                           fb, pol_keep);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ relay (both CTAs) + UMMA issue (leader)
        constexpr uint32_t kIdesc = ptx::umma_idesc_i8_m256(256, false, false);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
            if (rank == 0) {
                ptx::mbar_wait(tempty_bar, (it & 1u) ^ 1u, p.error, kWmTmemEmpty);
                ptx::tc_fence_after();
            }
            for (int c = 0; c < kChunks; ++c) {
                ptx::mbar_wait(full_bar(stage), phase, p.error, kWmFull);        // query operand landed
                ptx::mbar_wait(expd_bar(stage), phase, p.error, kWmExp);         // mask operand expanded
                const uint32_t ready_leader = ptx::mapa(ready_bar(stage), 0);
                if (ptx::elect_one_sync()) ptx::mbar_arrive_cluster(ready_leader);
                __syncwarp();
                if (rank == 0) {
                    ptx::mbar_wait(ready_bar(stage), phase, p.error, kWmReady);
                    ptx::tc_fence_after();
                    const uint32_t sbase = base + stage * kMaskStageBytes;
                    if (ptx::elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < kChunkK / 32; ++k) {
                            const uint32_t acc = (c | k) ? 1u : 0u;
                            const uint64_t a = ptx::umma_desc_sw128(sbase + kMaskOffAmx + 32 * k);
                            ptx::umma_i8_2cta(tmem_base + 0, a, ptx::umma_desc_sw128(sbase + kMaskOffB + 32 * k), kIdesc, acc);
                            ptx::umma_i8_2cta(tmem_base + 256, a,
                                              ptx::umma_desc_sw128(sbase + kMaskOffB + 4 * kQTileBytes + 32 * k), kIdesc, acc);
                        }
                        ptx::umma_commit_2cta(empty_bar(stage), 3);
                        if (c == kChunks - 1) ptx::umma_commit_2cta(tfull_bar, 3);
                    }
                    __syncwarp();
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 6) {
        // ------------------------------------------------------------------ mask bit -> byte expanders (own 128 rows)
        const int row = threadIdx.x - 6 * 32;
        const uint32_t sw = row & 7;
        int stage = 0;
        uint32_t phase = 0;
        for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters) {
            for (int c = 0; c < kChunks; ++c) {
                ptx::mbar_wait(full_bar(stage), phase, p.error, kWmExpander);
                uint8_t* sptr = base_ptr + stage * kMaskStageBytes;
                const uint4 x = *reinterpret_cast<const uint4*>(sptr + kMaskOffPk + row * 16);
                uint8_t* dst = sptr + kMaskOffAmx + row * 128;
                const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t v = xs[w];
                    uint4 lo4 = make_uint4(v & 0x01010101u, v & 0x02020202u, v & 0x04040404u, v & 0x08080808u);
                    uint4 hi4 = make_uint4(v & 0x10101010u, v & 0x20202020u, v & 0x40404040u, v & 0x80808080u);
                    *reinterpret_cast<uint4*>(dst + (((2 * w) ^ sw) << 4)) = lo4;
                    *reinterpret_cast<uint4*>(dst + (((2 * w + 1) ^ sw) << 4)) = hi4;
                }
                ptx::fence_proxy_async_smem();
                ptx::mbar_arrive(expd_bar(stage));
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0..3 of each CTA)
        const int row = threadIdx.x;
        const uint32_t tempty_leader = ptx::mapa(tempty_bar, 0);
        const uint64_t rows_out = p.row_end - p.row_begin;
        uint32_t it = 0;
        for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
            const uint32_t pair = p.pair_begin + t / num_groups;
            const uint32_t group = t % num_groups;
            ptx::mbar_wait(tfull_bar, it & 1u, p.error, kWmEpilogue);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            const int64_t trow0 = ((int64_t)2 * pair + rank) * kTileRows;
            int64_t lo = (int64_t)p.row_begin - trow0, hi = (int64_t)p.row_end - trow0;
            const int r0 = (int)(lo < 0 ? 0 : (lo > kTileRows ? kTileRows : lo));
            const int r1 = (int)(hi < 0 ? 0 : (hi > kTileRows ? kTileRows : hi));
            const int64_t tile_off = (trow0 - (int64_t)p.row_begin) * kOutRowBytes;
#pragma unroll 1
            for (int g = 0; g < kMaskQTile; ++g) {
                const uint32_t qi = group * kMaskQTile + g;
                uint32_t a[32];
                ptx::tmem_ld32(taddr + 32 * g, a);
                ptx::tmem_wait_ld();
                if (g == kMaskQTile - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader);
                }
                if (qi < p.num_queries) {
                    uint8_t* outq = reinterpret_cast<uint8_t*>(p.out + (size_t)qi * rows_out * IRIS_ROTATIONS);
                    const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(outq) + tile_off) & 15);
                    uint8_t* stage_buf = out_stage_ptr + (g & 1) * kBatchOutStageBytes;
                    uint8_t* st = stage_buf + shift + row * kOutRowBytes;
#pragma unroll
                    for (int j = 0; j < IRIS_ROTATIONS; ++j) *reinterpret_cast<uint16_t*>(st + 2 * j) = (uint16_t)(a[j] >> 7);
                    ptx::named_bar_sync(1, 128);
                    copy_out_rows(stage_buf, outq + tile_off - shift, (int)shift + r0 * kOutRowBytes,
                                  (int)shift + r1 * kOutRowBytes, row);
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 5) ptx::tmem_dealloc_2cta(tmem_base, 512);
}

cudaError_t launch_batch_denominators(const BatchMaskParams& p, int num_sms, cudaStream_t stream) {
    if (p.num_queries == 0 || p.num_queries > kMaxBatchQueries) return cudaErrorInvalidValue;
    static std::atomic<bool> configured[64];    // per device: opt-in shared memory size set
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(batch_denominators_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaskSmemBytes);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const uint32_t num_groups = (p.num_queries + kMaskQTile - 1) / kMaskQTile;
    const uint32_t tiles = (p.pair_end - p.pair_begin) * num_groups;
    if (tiles == 0) return cudaSuccess;
    uint32_t clusters = (uint32_t)num_sms / 2;
    if (tiles < clusters) clusters = tiles;
    batch_denominators_kernel<<<2 * clusters, kMaskThreads, kMaskSmemBytes, stream>>>(p);
    count_launch_external();
    return cudaGetLastError();
}

#endif  // IRIS_DIAGNOSTICS

}  // namespace iris
