// sm_100a kernels of the matching hot path.
//
//   scan_kernel<SHARES, MASKS>  -- the product: persistent, warp-specialised tcgen05 (UMMA kind::i8)
//       scan of the HBM-resident database against ONE prepared query at all 31 rotations.
//         distances    (reference src/lib.rs:42-52 -> src/arch/generic.rs:11-16)
//         denominators (reference src/lib.rs:69-79 -> src/arch/generic.rs:4-9)
//   prep_* / retile_* / untile_* / generate_*   -- query preparation and the loader
//   simt_* / dot_*  -- CUDA-core cross-checks and the per-pair arch entry points
//
// Arithmetic (bit-exact):
//   u16 dot  q.d mod 2^16 = S00 + 256*(S10 + S01) mod 2^16 with u8 limbs q = q_lo + 256 q_hi,
//   d = d_lo + 256 d_hi, S00 = sum q_lo d_lo, S10 = sum q_hi d_lo, S01 = sum q_lo d_hi (the hi.hi
//   term vanishes mod 2^16); each sum <= 12800*255*255 < 2^31 so s32 accumulation is exact.
//   popcount(qmask & dmask) = (1/128) * sum_k A[k] B[k] with A[k] = dbit << t, B[k] = qbit << (7-t),
//   t = bit position inside the source byte; the sum <= 12800*128 < 2^31.
#include <atomic>
#include <cstdlib>

#include "iris_epilogue.cuh"
#include "iris_kernels.cuh"
#include "iris_ptx.cuh"

namespace iris {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
static inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launch_external() { count_launch(); }

// =====================================================================================
// scan kernel
// =====================================================================================
// SQ: every query element is a sign-extended byte (true for encode() output): the q_lo plane read as s8 IS the
// value, so the q_hi plane is neither loaded nor multiplied (two limb products instead of three).
template <bool S, bool M, bool SQ>
struct ScanCfg {
    static constexpr int kOffAlo = 0;
    static constexpr int kOffAhi = kPlaneTileBytes;
    static constexpr int kOffQd = kShareChunkBytes;
    static constexpr int kQdLoadBytes = SQ ? kQTileBytes : kQdChunkBytes;
    static constexpr int kShareBytes = S ? kShareChunkBytes + kQdLoadBytes : 0;
    static constexpr int kOffAmx = kShareBytes;                  // expanded mask operand (written by SM)
    static constexpr int kOffQm = kOffAmx + kPlaneTileBytes;
    static constexpr int kOffPk = kOffQm + kQmChunkBytes;        // packed mask bytes (bulk-copied)
    static constexpr int kStageBytes = kShareBytes + (M ? kPlaneTileBytes + kQmChunkBytes + kMaskChunkBytes : 0);
    static constexpr int kStages = (S && M) ? 3 : (S ? 5 : 8);
    static constexpr uint32_t kTxBytes =
        (S ? kShareChunkBytes + kQdLoadBytes : 0) + (M ? kQmChunkBytes + kMaskChunkBytes : 0);
    static constexpr int kOutStageBytes = 8192;                  // 128*62 + alignment slack
    static constexpr int kBarBytes = 512;
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * kOutStageBytes + kBarBytes;
    static_assert(kStageBytes % 1024 == 0, "operand tiles must stay 1024-byte aligned");
    static_assert(kSmemBytes <= 232448, "exceeds 227 KiB of shared memory");
};

// warps 0..3: epilogue (TMEM lane quadrant = warp index)
constexpr int kProducerWarp = 4;
constexpr int kMmaWarp = 5;
constexpr int kExpanderWarp0 = 6;     // warps 6..9
constexpr int kScanThreads = 320;
constexpr uint32_t kTmemCols = 256;   // 2 accumulator buffers x 128 columns
// accumulator columns: [0,32) S00, [32,64) S10, [64,96) S01, [96,128) 128*popcount

enum WatchdogCode { kWdProducer = 101, kWdMmaFull = 102, kWdMmaExp = 103, kWdMmaTmem = 104, kWdExpander = 105, kWdEpilogue = 106 };


// R (search mode, fused scan only): decode + running minimum in the epilogue instead of per-row results.
template <bool S, bool M, bool SQ, bool R = false>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const ScanParams p) {
    static_assert(!R || (S && M), "search mode needs distances and denominators");
    using Cfg = ScanCfg<S, M, SQ>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;          // shared-window address, 1024-aligned
    uint8_t* const base_ptr = smem_raw + (base - raw_addr);
    const uint32_t out_stage = base + Cfg::kStages * Cfg::kStageBytes;
    uint8_t* const out_stage_ptr = base_ptr + Cfg::kStages * Cfg::kStageBytes;
    const uint32_t bars = out_stage + 2 * Cfg::kOutStageBytes;
    // barrier table (8 bytes each)
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::kStages + s); };
    auto expd_bar = [&](int s) { return bars + 8u * (2 * Cfg::kStages + s); };
    auto tfull_bar = [&](int b) { return bars + 8u * (3 * Cfg::kStages + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (3 * Cfg::kStages + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (3 * Cfg::kStages + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(out_stage_ptr + 2 * Cfg::kOutStageBytes + 8 * (3 * Cfg::kStages + 4));

    const int warp = ptx::warp_idx_sync();      // warp-uniform role index (see iris_ptx.cuh)

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
            ptx::mbar_init(expd_bar(s), 128);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(tfull_bar(b), 1);
            ptx::mbar_init(tempty_bar(b), 128);
        }
        ptx::fence_mbar_init();
    }
    if (warp == kMmaWarp) ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    ptx::pdl_launch_dependents();               // the next scan's CTAs may take the SMs this kernel leaves idle

    const uint32_t tile0 = p.tile_begin + blockIdx.x;
    const uint32_t tile_step = gridDim.x;

    if (warp == kProducerWarp) {
        // ------------------------------------------------------------------ bulk-copy producer (uniform flow, one
        // elected lane issues)
        const uint64_t pol_stream = ptx::policy_evict_first();
        const uint64_t pol_keep = ptx::policy_evict_last();
        int stage = 0;
        uint32_t phase = 0;
        for (uint32_t tile = tile0; tile < p.tile_end; tile += tile_step) {
            const uint8_t* sh = S ? p.shares + (size_t)tile * kShareTileBytes : nullptr;
            const uint8_t* mk = M ? p.masks + (size_t)tile * kMaskTileBytes : nullptr;
            for (int c = 0; c < kChunks; ++c) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.error, kWdProducer);
                const uint32_t sbase = base + stage * Cfg::kStageBytes;
                const uint32_t fb = full_bar(stage);
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(fb, Cfg::kTxBytes);
                    if (S) {
                        ptx::bulk_g2s_hint(sbase + Cfg::kOffAlo, sh + (size_t)c * kShareChunkBytes, kShareChunkBytes, fb,
                                           pol_stream);
                        ptx::bulk_g2s_hint(sbase + Cfg::kOffQd, p.qd + (size_t)c * kQdChunkBytes, Cfg::kQdLoadBytes, fb,
                                           pol_keep);
                    }
                    if (M) {
                        ptx::bulk_g2s_hint(sbase + Cfg::kOffPk, mk + (size_t)c * kMaskChunkBytes, kMaskChunkBytes, fb,
                                           pol_stream);
                        ptx::bulk_g2s_hint(sbase + Cfg::kOffQm, p.qm + (size_t)c * kQmChunkBytes, kQmChunkBytes, fb,
                                           pol_keep);
                    }
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ UMMA issuer
        constexpr uint32_t kIdesc64 = ptx::umma_idesc_i8(64);
        constexpr uint32_t kIdesc32 = ptx::umma_idesc_i8(32);
        constexpr uint32_t kIdesc32S = ptx::umma_idesc_i8(32, false, true);   // B read as s8
        int stage = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        for (uint32_t tile = tile0; tile < p.tile_end; tile += tile_step, ++it) {
            const uint32_t buf = it & 1u;
            ptx::mbar_wait(tempty_bar(buf), ((it >> 1) & 1u) ^ 1u, p.error, kWdMmaTmem);
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + buf * 128u;
            for (int c = 0; c < kChunks; ++c) {
                ptx::mbar_wait(full_bar(stage), phase, p.error, kWdMmaFull);
                if (M) ptx::mbar_wait(expd_bar(stage), phase, p.error, kWdMmaExp);
                ptx::tc_fence_after();
                const uint32_t sbase = base + stage * Cfg::kStageBytes;
                if (ptx::elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < kChunkK / 32; ++k) {
                        const uint32_t acc = (c | k) ? 1u : 0u;
                        if (S) {
                            const uint64_t bq = ptx::umma_desc_sw128(sbase + Cfg::kOffQd + 32 * k);
                            ptx::umma_i8(d + 0, ptx::umma_desc_sw128(sbase + Cfg::kOffAlo + 32 * k), bq,
                                         SQ ? kIdesc32S : kIdesc64, acc);
                            ptx::umma_i8(d + 64, ptx::umma_desc_sw128(sbase + Cfg::kOffAhi + 32 * k), bq,
                                         SQ ? kIdesc32S : kIdesc32, acc);
                        }
                        if (M) {
                            ptx::umma_i8(d + 96, ptx::umma_desc_sw128(sbase + Cfg::kOffAmx + 32 * k),
                                         ptx::umma_desc_sw128(sbase + Cfg::kOffQm + 32 * k), kIdesc32, acc);
                        }
                    }
                    ptx::umma_commit(empty_bar(stage));
                    if (c == kChunks - 1) ptx::umma_commit(tfull_bar(buf));
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= kExpanderWarp0) {
        // ------------------------------------------------------------------ mask bit -> byte expanders
        if (M) {
            const int row = threadIdx.x - kExpanderWarp0 * 32;
            const uint32_t sw = row & 7;
            int stage = 0;
            uint32_t phase = 0;
            for (uint32_t tile = tile0; tile < p.tile_end; tile += tile_step) {
                for (int c = 0; c < kChunks; ++c) {
                    ptx::mbar_wait(full_bar(stage), phase, p.error, kWdExpander);
                    uint8_t* sptr = base_ptr + stage * Cfg::kStageBytes;
                    const uint4 x = *reinterpret_cast<const uint4*>(sptr + Cfg::kOffPk + row * 16);
                    uint8_t* dst = sptr + Cfg::kOffAmx + row * 128;
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
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0..3)
        const int row = threadIdx.x;   // 0..127 == TMEM lane
        uint32_t it = 0;
        // search mode: this thread's best row so far as an exact fraction (best_d = 0: none yet) -- see
        // combine_decode_kernel (iris_reduce.cu) for why comparing fractions equals comparing the f64 quotients
        uint32_t best_n = 0, best_d = 0;
        unsigned long long best_row = ~0ull;
        for (uint32_t tile = tile0; tile < p.tile_end; tile += tile_step, ++it) {
            const uint32_t buf = it & 1u;
            ptx::mbar_wait(tfull_bar(buf), (it >> 1) & 1u, p.error, kWdEpilogue);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 128u;

            // rows of this tile inside [row_begin,row_end)
            const int64_t trow0 = (int64_t)tile * kTileRows;
            int r0 = (int)max((int64_t)0, (int64_t)p.row_begin - trow0);
            int r1 = (int)min((int64_t)kTileRows, (int64_t)p.row_end - trow0);
            if (R) {
                uint32_t a[32], b[32], c2[32], m[32];
                ptx::tmem_ld32(taddr + 0, a);
                if (SQ) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) b[j] = 0;
                } else {
                    ptx::tmem_ld32(taddr + 32, b);
                }
                ptx::tmem_ld32(taddr + 64, c2);
                ptx::tmem_ld32(taddr + 96, m);
                ptx::tmem_wait_ld();
                ptx::tc_fence_before();
                ptx::mbar_arrive(tempty_bar(buf));             // the accumulators are in registers
                if (row >= r0 && row < r1) {
                    uint32_t bn = 0, bd = 0;                   // this row: min over the rotations (src/lib.rs:97-107)
#pragma unroll
                    for (int j = 0; j < IRIS_ROTATIONS; ++j) {
                        const uint32_t dist = (a[j] + ((b[j] + c2[j]) << 8)) & 0xFFFFu;
                        const uint32_t den = (m[j] >> 7) & 0xFFFFu;
                        const uint32_t num = ((den - dist) & 0xFFFFu) >> 1;
                        if (den != 0 && (bd == 0 || num * bd < bn * den)) {
                            bn = num;
                            bd = den;
                        }
                    }
                    // strict `<` against the running minimum (src/main.rs:617): rows come in ascending order
                    if (bd != 0 && (best_d == 0 || (uint64_t)bn * best_d < (uint64_t)best_n * bd)) {
                        best_n = bn;
                        best_d = bd;
                        best_row = p.index_base + (uint64_t)(trow0 + row);
                    }
                }
                continue;
            }
            // byte offset (possibly negative) of tile row 0 in the packed output
            const int64_t tile_off = (trow0 - (int64_t)p.row_begin) * kOutRowBytes;

            uint32_t a[32];
            if (S) {
                uint32_t b[32], c2[32];
                ptx::tmem_ld32(taddr + 0, a);
                if (SQ) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) b[j] = 0;    // no q_hi product on the signed path
                } else {
                    ptx::tmem_ld32(taddr + 32, b);
                }
                ptx::tmem_ld32(taddr + 64, c2);
                ptx::tmem_wait_ld();
                if (p.raw_out) {
                    int32_t* ro = p.raw_out + ((size_t)(tile - p.tile_begin) * kTileRows + row) * 128;
#pragma unroll
                    for (int j = 0; j < 32; ++j) { ro[j] = (int32_t)a[j]; ro[32 + j] = (int32_t)b[j]; ro[64 + j] = (int32_t)c2[j]; }
                }
                const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(p.dist_out) + tile_off) & 15);
                uint8_t* st = out_stage_ptr + shift + row * kOutRowBytes;
#pragma unroll
                for (int j = 0; j < IRIS_ROTATIONS; ++j)
                    *reinterpret_cast<uint16_t*>(st + 2 * j) = (uint16_t)(a[j] + ((b[j] + c2[j]) << 8));
            }
            if (M) {
                ptx::tmem_ld32(taddr + 96, a);
                ptx::tmem_wait_ld();
                if (p.raw_out) {
                    int32_t* ro = p.raw_out + ((size_t)(tile - p.tile_begin) * kTileRows + row) * 128;
#pragma unroll
                    for (int j = 0; j < 32; ++j) ro[96 + j] = (int32_t)a[j];
                }
                const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(p.den_out) + tile_off) & 15);
                uint8_t* st = out_stage_ptr + Cfg::kOutStageBytes + shift + row * kOutRowBytes;
#pragma unroll
                for (int j = 0; j < IRIS_ROTATIONS; ++j) *reinterpret_cast<uint16_t*>(st + 2 * j) = (uint16_t)(a[j] >> 7);
            }
            // accumulator buffer may be overwritten by the next-but-one tile from here on
            ptx::tc_fence_before();
            ptx::mbar_arrive(tempty_bar(buf));
            ptx::named_bar_sync(1, 128);
            if (S) {
                const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(p.dist_out) + tile_off) & 15);
                uint8_t* g = reinterpret_cast<uint8_t*>(p.dist_out) + tile_off - shift;
                copy_out_rows(out_stage_ptr, g, (int)shift + r0 * kOutRowBytes, (int)shift + r1 * kOutRowBytes, row);
            }
            if (M) {
                const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(p.den_out) + tile_off) & 15);
                uint8_t* g = reinterpret_cast<uint8_t*>(p.den_out) + tile_off - shift;
                copy_out_rows(out_stage_ptr + Cfg::kOutStageBytes, g, (int)shift + r0 * kOutRowBytes,
                         (int)shift + r1 * kOutRowBytes, row);
            }
            ptx::named_bar_sync(1, 128);
        }
        if (R) {
            // CTA minimum: one division per thread, warp shuffles, then the four epilogue warps through shared memory
            double v = best_d ? (double)best_n / (double)best_d : __longlong_as_double(0x7FF0000000000000ll);
            unsigned long long i = best_row;
            for (int o = 16; o; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const unsigned long long oi = __shfl_xor_sync(0xffffffffu, i, o);
                if (ov < v || (ov == v && oi < i)) {
                    v = ov;
                    i = oi;
                }
            }
            double* sv = reinterpret_cast<double*>(out_stage_ptr);
            unsigned long long* si = reinterpret_cast<unsigned long long*>(out_stage_ptr + 64);
            if ((threadIdx.x & 31) == 0) {
                sv[warp] = v;
                si[warp] = i;
            }
            ptx::named_bar_sync(1, 128);
            if (threadIdx.x == 0) {
                for (int w = 1; w < 4; ++w)
                    if (sv[w] < v || (sv[w] == v && si[w] < i)) {
                        v = sv[w];
                        i = si[w];
                    }
                p.red_min[blockIdx.x] = v;
                p.red_idx[blockIdx.x] = i;
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, kTmemCols);
    ptx::pdl_wait();                            // complete in stream order
}

template <bool S, bool M, bool SQ, bool R = false>
static cudaError_t launch_scan_t(const ScanParams& p, int num_sms, cudaStream_t stream) {
    using Cfg = ScanCfg<S, M, SQ>;
    static std::atomic<bool> configured[64];    // per device: opt-in shared memory size set
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(scan_kernel<S, M, SQ, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const uint32_t tiles = p.tile_end - p.tile_begin;
    if (tiles == 0) return cudaSuccess;
    const uint32_t grid = tiles < (uint32_t)num_sms ? tiles : (uint32_t)num_sms;
    if (p.pdl) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kScanThreads);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        count_launch();
        return cudaLaunchKernelEx(&cfg, scan_kernel<S, M, SQ, R>, p);
    }
    scan_kernel<S, M, SQ, R><<<grid, kScanThreads, Cfg::kSmemBytes, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

uint32_t scan_grid(const ScanParams& p, int num_sms) {
    const uint32_t tiles = p.tile_end - p.tile_begin;
    return tiles < (uint32_t)num_sms ? tiles : (uint32_t)num_sms;
}

cudaError_t launch_scan(const ScanParams& p, int num_sms, cudaStream_t stream) {
    const bool s = p.shares != nullptr, m = p.masks != nullptr;
    if (p.red_min || p.red_idx) {
        if (!s || !m || !p.red_min || !p.red_idx || p.raw_out) return cudaErrorInvalidValue;
        return p.signed_query ? launch_scan_t<true, true, true, true>(p, num_sms, stream)
                              : launch_scan_t<true, true, false, true>(p, num_sms, stream);
    }
    if (s && m) return p.signed_query ? launch_scan_t<true, true, true>(p, num_sms, stream)
                                      : launch_scan_t<true, true, false>(p, num_sms, stream);
    if (s) return p.signed_query ? launch_scan_t<true, false, true>(p, num_sms, stream)
                                 : launch_scan_t<true, false, false>(p, num_sms, stream);
    if (m) {
        // denominators only: the 4-bit TMEM-operand kernel (iris_maskscan4.cu).  The shared-memory-operand variant of
        // this file serves the raw debug dump (and engines without the 4-bit image).
        if (p.raw_out || !p.qm4) return launch_scan_t<false, true, false>(p, num_sms, stream);
#ifdef IRIS_DIAGNOSTICS
        // Diagnostics build only: IRIS_MASKSCAN=i8 / smem select the int8 TMEM kernel (iris_maskscan.cu) / the
        // shared-memory-operand kernel for A/B measurements.
        static const char mode = [] {
            const char* e = getenv("IRIS_MASKSCAN");
            return e ? e[0] : 'f';
        }();
        if (mode == 's') return launch_scan_t<false, true, false>(p, num_sms, stream);
        if (mode == 'i') return launch_mask_scan(p, num_sms, stream);
#endif
        ScanParams p4 = p;
        p4.qm = p.qm4;
        return launch_mask_scan_fp4(p4, num_sms, stream);
    }
    return cudaErrorInvalidValue;
}

// =====================================================================================
// query preparation (reference: DistanceEngine::new / MasksEngine::new, src/lib.rs:33-40, 60-67;
// rotation semantics src/encoded_bits.rs:40-52 and src/bits.rs:178-205:
// rot(v, r)[row][col] = v[row][(col - r) mod 200], output slot j <-> r = j - 15)
// =====================================================================================
__global__ void prep_distance_query_kernel(const uint16_t* __restrict__ q, uint8_t* __restrict__ qd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (c, j, ch)
    if (idx >= kChunks * 32 * 8) return;
    const int ch = idx & 7, j = (idx >> 3) & 31, c = idx >> 8;
    uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    if (j < IRIS_ROTATIONS) {
        const int rot = j - 15;
        for (int b = 0; b < 16; ++b) {
            const int k = c * kChunkK + ch * 16 + b;
            const int row = k / IRIS_COLS, col = k % IRIS_COLS;
            const int src = row * IRIS_COLS + (col - rot + IRIS_COLS) % IRIS_COLS;
            const uint32_t v = q[src];
            lo[b >> 2] |= (v & 0xFFu) << (8 * (b & 3));
            hi[b >> 2] |= (v >> 8) << (8 * (b & 3));
        }
    }
    const size_t off = (size_t)c * kQdChunkBytes + j * 128 + ((ch ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(qd + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(qd + off + kQTileBytes) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

__global__ void prep_mask_query_kernel(const uint8_t* __restrict__ qmask, uint8_t* __restrict__ qm) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (c, j, ch)
    if (idx >= kChunks * 32 * 8) return;
    const int ch = idx & 7, j = (idx >> 3) & 31, c = idx >> 8;
    uint32_t out[4] = {0, 0, 0, 0};
    if (j < IRIS_ROTATIONS) {
        const int rot = j - 15;
        for (int b = 0; b < 16; ++b) {
            const int e = ch * 16 + b;
            const int w = e >> 5, t = (e >> 2) & 7, m = e & 3;
            const int s = c * kChunkK + 32 * w + 8 * m + t;
            const int row = s / IRIS_COLS, col = s % IRIS_COLS;
            const int src = row * IRIS_COLS + (col - rot + IRIS_COLS) % IRIS_COLS;
            const uint32_t bit = (qmask[src >> 3] >> (src & 7)) & 1u;
            out[b >> 2] |= (bit << (7 - t)) << (8 * (b & 3));
        }
    }
    const size_t off = (size_t)c * kQmChunkBytes + j * 128 + ((ch ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(qm + off) = make_uint4(out[0], out[1], out[2], out[3]);
}

// encode (src/lib.rs:16-26): mask - 2*(pattern & mask) in Z/2^16 -> 1 (mask & !pattern), 0 (!mask), 0xFFFF (mask & pattern)
__global__ void encode_kernel(const uint8_t* __restrict__ pattern, const uint8_t* __restrict__ mask,
                              uint16_t* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= IRIS_BITS) return;
    const uint32_t m = (mask[k >> 3] >> (k & 7)) & 1u;
    const uint32_t p = (pattern[k >> 3] >> (k & 7)) & m;
    out[k] = (uint16_t)(m - p - p);
}
cudaError_t launch_encode(const uint8_t* d_pattern, const uint8_t* d_mask, uint16_t* d_out, cudaStream_t stream) {
    encode_kernel<<<(IRIS_BITS + 255) / 256, 256, 0, stream>>>(d_pattern, d_mask, d_out);
    count_launch();
    return cudaGetLastError();
}

// Batched query preparation: Q wire Templates (3 200 B each: pattern, mask) -> Q encoded queries, Q distance
// operand images and Q mask operand images in three launches.
__global__ void encode_batch_kernel(const PrepBatchParams p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = blockIdx.y;
    if (k >= IRIS_BITS) return;
    const uint8_t* t = p.templates + (size_t)q * 2 * IRIS_MASK_BYTES;
    const uint32_t m = (t[IRIS_MASK_BYTES + (k >> 3)] >> (k & 7)) & 1u;
    const uint32_t v = (t[k >> 3] >> (k & 7)) & m;
    p.query[q][k] = (uint16_t)(m - v - v);
}
__global__ void prep_distance_batch_kernel(const PrepBatchParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (c, j, ch)
    const uint32_t qi = blockIdx.y;
    if (idx >= kChunks * 32 * 8) return;
    const uint16_t* __restrict__ q = p.query[qi];
    const int ch = idx & 7, j = (idx >> 3) & 31, c = idx >> 8;
    uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    if (j < IRIS_ROTATIONS) {
        const int rot = j - 15;
        for (int b = 0; b < 16; ++b) {
            const int k = c * kChunkK + ch * 16 + b;
            const int row = k / IRIS_COLS, col = k % IRIS_COLS;
            const uint32_t v = q[row * IRIS_COLS + (col - rot + IRIS_COLS) % IRIS_COLS];
            lo[b >> 2] |= (v & 0xFFu) << (8 * (b & 3));
            hi[b >> 2] |= (v >> 8) << (8 * (b & 3));
        }
    }
    const size_t off = (size_t)c * kQdChunkBytes + j * 128 + ((ch ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(p.qd[qi] + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(p.qd[qi] + off + kQTileBytes) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}
__global__ void prep_mask_batch_kernel(const PrepBatchParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (c, j, ch)
    const uint32_t qi = blockIdx.y;
    if (idx >= kChunks * 32 * 8) return;
    const uint8_t* __restrict__ qmask = p.templates + (size_t)qi * 2 * IRIS_MASK_BYTES + IRIS_MASK_BYTES;
    const int ch = idx & 7, j = (idx >> 3) & 31, c = idx >> 8;
    uint32_t out[4] = {0, 0, 0, 0};
    if (j < IRIS_ROTATIONS) {
        const int rot = j - 15;
        for (int b = 0; b < 16; ++b) {
            const int e = ch * 16 + b;
            const int w = e >> 5, t = (e >> 2) & 7, m = e & 3;
            const int s = c * kChunkK + 32 * w + 8 * m + t;
            const int row = s / IRIS_COLS, col = s % IRIS_COLS;
            const int src = row * IRIS_COLS + (col - rot + IRIS_COLS) % IRIS_COLS;
            const uint32_t bit = (qmask[src >> 3] >> (src & 7)) & 1u;
            out[b >> 2] |= (bit << (7 - t)) << (8 * (b & 3));
        }
    }
    const size_t off = (size_t)c * kQmChunkBytes + j * 128 + ((ch ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(p.qm[qi] + off) = make_uint4(out[0], out[1], out[2], out[3]);
}
cudaError_t launch_prep_batch(const PrepBatchParams& p, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    encode_batch_kernel<<<dim3((IRIS_BITS + 255) / 256, p.n), 256, 0, stream>>>(p);
    count_launch();
    prep_distance_batch_kernel<<<dim3((kChunks * 32 * 8 + 255) / 256, p.n), 256, 0, stream>>>(p);
    count_launch();
    if (p.qm[0]) {
        prep_mask_batch_kernel<<<dim3((kChunks * 32 * 8 + 255) / 256, p.n), 256, 0, stream>>>(p);
        count_launch();
        cudaError_t e = launch_prep_mask_fp4_batch(p, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

__global__ void classify_s8_kernel(const uint16_t* __restrict__ q, int* __restrict__ flag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < IRIS_BITS && (uint16_t)(q[k] + 0x80u) > 0xFFu) *flag = 0;
}
cudaError_t launch_classify_s8(const uint16_t* d_query, int* d_flag, cudaStream_t stream) {
    classify_s8_kernel<<<(IRIS_BITS + 255) / 256, 256, 0, stream>>>(d_query, d_flag);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_prep_distance_query(const uint16_t* d_query, uint8_t* d_qd, cudaStream_t stream) {
    prep_distance_query_kernel<<<(kChunks * 32 * 8 + 255) / 256, 256, 0, stream>>>(d_query, d_qd);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_prep_mask_query(const uint8_t* d_qmask, uint8_t* d_qm, uint8_t* d_qm4, cudaStream_t stream) {
    prep_mask_query_kernel<<<(kChunks * 32 * 8 + 255) / 256, 256, 0, stream>>>(d_qmask, d_qm);
    count_launch();
    return launch_prep_mask_query_fp4(d_qmask, d_qm4, stream);
}

// =====================================================================================
// loader: reference layouts (&[EncodedBits], &[Bits]; src/main.rs:389-391, 458-461) -> tiled image
// =====================================================================================
__global__ void retile_shares_kernel(const uint16_t* __restrict__ rows, uint64_t n, uint8_t* __restrict__ shares,
                                     uint64_t row0) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row i, group g of 16 elements)
    if (idx >= n * (IRIS_BITS / 16)) return;
    const uint64_t i = idx / (IRIS_BITS / 16);
    const uint32_t g = (uint32_t)(idx % (IRIS_BITS / 16));
    const uint4* src = reinterpret_cast<const uint4*>(rows + i * IRIS_BITS + g * 16);
    const uint4 v0 = src[0], v1 = src[1];
    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const uint32_t x = w[2 * t], y = w[2 * t + 1];   // elements 4t..4t+3
        lo[t] = __byte_perm(x, y, 0x6420);
        hi[t] = __byte_perm(x, y, 0x7531);
    }
    const size_t off = share_offset(row0 + i, g * 16, 0);
    *reinterpret_cast<uint4*>(shares + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(shares + off + kPlaneTileBytes) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

__global__ void retile_masks_kernel(const uint8_t* __restrict__ rows, uint64_t n, uint8_t* __restrict__ masks,
                                    uint64_t row0) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row i, chunk c)
    if (idx >= n * kChunks) return;
    const uint64_t i = idx / kChunks;
    const uint32_t c = (uint32_t)(idx % kChunks);
    const uint4 v = *reinterpret_cast<const uint4*>(rows + i * IRIS_MASK_BYTES + c * 16);
    *reinterpret_cast<uint4*>(masks + mask_offset(row0 + i, c * 16)) = v;
}

__global__ void untile_shares_kernel(const uint8_t* __restrict__ shares, uint64_t row0, uint64_t n,
                                     uint16_t* __restrict__ rows) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * IRIS_BITS) return;
    const uint64_t i = idx / IRIS_BITS;
    const uint32_t k = (uint32_t)(idx % IRIS_BITS);
    const size_t off = share_offset(row0 + i, k, 0);
    rows[idx] = (uint16_t)(shares[off] | ((uint16_t)shares[off + kPlaneTileBytes] << 8));
}

__global__ void untile_masks_kernel(const uint8_t* __restrict__ masks, uint64_t row0, uint64_t n,
                                    uint8_t* __restrict__ rows) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * IRIS_MASK_BYTES) return;
    const uint64_t i = idx / IRIS_MASK_BYTES;
    const uint32_t b = (uint32_t)(idx % IRIS_MASK_BYTES);
    rows[idx] = masks[mask_offset(row0 + i, b)];
}

static inline unsigned blocks_for(uint64_t threads, unsigned bs) { return (unsigned)((threads + bs - 1) / bs); }

cudaError_t launch_retile_shares(const uint16_t* d_rows, uint64_t n, uint8_t* d_shares, uint64_t row0,
                                 cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    retile_shares_kernel<<<blocks_for(n * (IRIS_BITS / 16), 256), 256, 0, stream>>>(d_rows, n, d_shares, row0);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_retile_masks(const uint8_t* d_rows, uint64_t n, uint8_t* d_masks, uint64_t row0,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    retile_masks_kernel<<<blocks_for(n * kChunks, 256), 256, 0, stream>>>(d_rows, n, d_masks, row0);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_untile_shares(const uint8_t* d_shares, uint64_t row0, uint64_t n, uint16_t* d_rows,
                                 cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    untile_shares_kernel<<<blocks_for(n * IRIS_BITS, 256), 256, 0, stream>>>(d_shares, row0, n, d_rows);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_untile_masks(const uint8_t* d_masks, uint64_t row0, uint64_t n, uint8_t* d_rows,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    untile_masks_kernel<<<blocks_for(n * IRIS_MASK_BYTES, 256), 256, 0, stream>>>(d_masks, row0, n, d_rows);
    count_launch();
    return cudaGetLastError();
}

// =====================================================================================
// synthetic database (uniform u16 shares = what share() yields, src/encoded_bits.rs:23-38;
// uniform mask bits = Standard for Bits, src/bits.rs:95-101).  Counter-based: element group
// g (4 u16) of row id R is mix64(seed ^ ((R*3200+g) * C)); limb l is mix64(seed' ^ ((R*200+l) * C)).
// =====================================================================================
__global__ void generate_shares_kernel(uint8_t* __restrict__ shares, uint64_t seed, uint64_t row_id0, uint64_t row0,
                                       uint64_t n) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row i, group of 16 elements)
    if (idx >= n * (IRIS_BITS / 16)) return;
    const uint64_t i = idx / (IRIS_BITS / 16);
    const uint32_t g16 = (uint32_t)(idx % (IRIS_BITS / 16));
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const uint64_t g = (uint64_t)g16 * 4 + t;
        const uint64_t h = gen_share_group(seed, row_id0 + i, g);
        const uint32_t x = (uint32_t)h, y = (uint32_t)(h >> 32);
        lo[t] = __byte_perm(x, y, 0x6420);
        hi[t] = __byte_perm(x, y, 0x7531);
    }
    const size_t off = share_offset(row0 + i, g16 * 16, 0);
    *reinterpret_cast<uint4*>(shares + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(shares + off + kPlaneTileBytes) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

__global__ void generate_masks_kernel(uint8_t* __restrict__ masks, uint64_t seed, uint64_t row_id0, uint64_t row0,
                                      uint64_t n) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row i, chunk c)
    if (idx >= n * kChunks) return;
    const uint64_t i = idx / kChunks;
    const uint32_t c = (uint32_t)(idx % kChunks);
    const uint64_t h0 = gen_bits_limb(seed, kGenMaskTag, row_id0 + i, 2 * c);
    const uint64_t h1 = gen_bits_limb(seed, kGenMaskTag, row_id0 + i, 2 * c + 1);
    *reinterpret_cast<uint4*>(masks + mask_offset(row0 + i, c * 16)) =
        make_uint4((uint32_t)h0, (uint32_t)(h0 >> 32), (uint32_t)h1, (uint32_t)(h1 >> 32));
}

// Party `party`'s additive share of encode(Template R) (see iris_kernels.cuh): uniform for every party but the last,
// which holds the encoding minus the sum of the others (src/encoded_bits.rs:23-38).
__global__ void generate_party_shares_kernel(uint8_t* __restrict__ shares, uint64_t seed, uint32_t party, uint32_t n_parties,
                                             uint64_t row_id0, uint64_t row0, uint64_t n) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row i, group of 16 elements)
    if (idx >= n * (IRIS_BITS / 16)) return;
    const uint64_t i = idx / (IRIS_BITS / 16);
    const uint32_t g16 = (uint32_t)(idx % (IRIS_BITS / 16));
    const uint64_t R = row_id0 + i;
    uint32_t w[8];                                   // 16 u16 elements, two per word
    if (party + 1 < n_parties) {
        const uint64_t ps = gen_party_seed(seed, party);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint64_t h = gen_share_group(ps, R, (uint64_t)g16 * 4 + t);
            w[2 * t] = (uint32_t)h;
            w[2 * t + 1] = (uint32_t)(h >> 32);
        }
    } else {
        // encode (src/lib.rs:16-26): mask - 2 (pattern & mask) -> 1, 0, 0xFFFF; 16 elements = a quarter of one limb
        const uint32_t k0 = g16 * 16;
        const uint64_t ml = gen_bits_limb(seed, kGenMaskTag, R, k0 / 64);
        const uint64_t pl = gen_bits_limb(seed, kGenPatternTag, R, k0 / 64);
        const uint32_t mb = (uint32_t)(ml >> (k0 % 64)) & 0xFFFFu;
        const uint32_t pb = (uint32_t)(pl >> (k0 % 64)) & 0xFFFFu & mb;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t m0 = (mb >> (2 * e)) & 1u, m1 = (mb >> (2 * e + 1)) & 1u;
            const uint32_t p0 = (pb >> (2 * e)) & 1u, p1 = (pb >> (2 * e + 1)) & 1u;
            w[e] = ((m0 - 2 * p0) & 0xFFFFu) | ((m1 - 2 * p1) << 16);
        }
        for (uint32_t q = 0; q + 1 < n_parties; ++q) {
            const uint64_t ps = gen_party_seed(seed, q);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint64_t h = gen_share_group(ps, R, (uint64_t)g16 * 4 + t);
                w[2 * t] = __vsub2(w[2 * t], (uint32_t)h);                 // per-halfword wrapping subtraction
                w[2 * t + 1] = __vsub2(w[2 * t + 1], (uint32_t)(h >> 32));
            }
        }
    }
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        lo[t] = __byte_perm(w[2 * t], w[2 * t + 1], 0x6420);
        hi[t] = __byte_perm(w[2 * t], w[2 * t + 1], 0x7531);
    }
    const size_t off = share_offset(row0 + i, g16 * 16, 0);
    *reinterpret_cast<uint4*>(shares + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(shares + off + kPlaneTileBytes) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

cudaError_t launch_generate_party_shares(uint8_t* d_shares, uint64_t seed, uint32_t party, uint32_t n_parties,
                                         uint64_t row_id0, uint64_t row0, uint64_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    generate_party_shares_kernel<<<blocks_for(n * (IRIS_BITS / 16), 256), 256, 0, stream>>>(d_shares, seed, party, n_parties,
                                                                                           row_id0, row0, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_generate(uint8_t* d_shares, uint8_t* d_masks, uint64_t seed, uint64_t row_id0, uint64_t row0,
                            uint64_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (d_shares) {
        generate_shares_kernel<<<blocks_for(n * (IRIS_BITS / 16), 256), 256, 0, stream>>>(d_shares, seed, row_id0, row0, n);
        count_launch();
    }
    if (d_masks) {
        generate_masks_kernel<<<blocks_for(n * kChunks, 256), 256, 0, stream>>>(d_masks, seed, row_id0, row0, n);
        count_launch();
    }
    return cudaGetLastError();
}

// =====================================================================================
// CUDA-core cross-check kernels (one block per database row)
// =====================================================================================
__global__ void __launch_bounds__(256) simt_distances_kernel(const uint8_t* __restrict__ shares,
                                                             const uint16_t* __restrict__ query, uint64_t row_begin,
                                                             uint64_t row_end, uint16_t* __restrict__ out) {
    __shared__ uint16_t q[IRIS_BITS];
    __shared__ uint32_t red[IRIS_ROTATIONS][8];
    const uint64_t R = row_begin + blockIdx.x;
    if (R >= row_end) return;
    for (int k = threadIdx.x; k < IRIS_BITS; k += 256) q[k] = query[k];
    __syncthreads();
    uint32_t acc[IRIS_ROTATIONS];
#pragma unroll
    for (int j = 0; j < IRIS_ROTATIONS; ++j) acc[j] = 0;
    for (int k = threadIdx.x; k < IRIS_BITS; k += 256) {
        const size_t off = share_offset(R, k, 0);
        const uint32_t d = shares[off] | ((uint32_t)shares[off + kPlaneTileBytes] << 8);
        const int row = k / IRIS_COLS, col = k % IRIS_COLS;
#pragma unroll
        for (int j = 0; j < IRIS_ROTATIONS; ++j) {
            int c = col - (j - 15);
            c += (c < 0) ? IRIS_COLS : 0;
            c -= (c >= IRIS_COLS) ? IRIS_COLS : 0;
            acc[j] += d * q[row * IRIS_COLS + c];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < IRIS_ROTATIONS; ++j) {
        uint32_t v = acc[j];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[j][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < IRIS_ROTATIONS) {
        uint32_t v = 0;
        for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
        out[(R - row_begin) * IRIS_ROTATIONS + threadIdx.x] = (uint16_t)v;
    }
}

__global__ void __launch_bounds__(128) simt_denominators_kernel(const uint8_t* __restrict__ masks,
                                                                const uint8_t* __restrict__ qmask, uint64_t row_begin,
                                                                uint64_t row_end, uint16_t* __restrict__ out) {
    __shared__ uint8_t qb[IRIS_BITS];     // query mask, one byte per bit
    __shared__ uint32_t red[IRIS_ROTATIONS][4];
    const uint64_t R = row_begin + blockIdx.x;
    if (R >= row_end) return;
    for (int k = threadIdx.x; k < IRIS_BITS; k += 128) qb[k] = (qmask[k >> 3] >> (k & 7)) & 1u;
    __syncthreads();
    uint32_t acc[IRIS_ROTATIONS];
#pragma unroll
    for (int j = 0; j < IRIS_ROTATIONS; ++j) acc[j] = 0;
    for (int k = threadIdx.x; k < IRIS_BITS; k += 128) {
        const uint32_t d = (masks[mask_offset(R, k >> 3)] >> (k & 7)) & 1u;
        const int row = k / IRIS_COLS, col = k % IRIS_COLS;
#pragma unroll
        for (int j = 0; j < IRIS_ROTATIONS; ++j) {
            int c = col - (j - 15);
            c += (c < 0) ? IRIS_COLS : 0;
            c -= (c >= IRIS_COLS) ? IRIS_COLS : 0;
            acc[j] += d & qb[row * IRIS_COLS + c];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < IRIS_ROTATIONS; ++j) {
        uint32_t v = acc[j];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[j][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < IRIS_ROTATIONS) {
        uint32_t v = 0;
        for (int w = 0; w < 4; ++w) v += red[threadIdx.x][w];
        out[(R - row_begin) * IRIS_ROTATIONS + threadIdx.x] = (uint16_t)v;
    }
}

cudaError_t launch_simt_distances(const uint8_t* d_shares, const uint16_t* d_query, uint64_t row_begin,
                                  uint64_t row_end, uint16_t* d_out, cudaStream_t stream) {
    if (row_end <= row_begin) return cudaSuccess;
    simt_distances_kernel<<<(unsigned)(row_end - row_begin), 256, 0, stream>>>(d_shares, d_query, row_begin, row_end, d_out);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_simt_denominators(const uint8_t* d_masks, const uint8_t* d_qmask, uint64_t row_begin,
                                     uint64_t row_end, uint16_t* d_out, cudaStream_t stream) {
    if (row_end <= row_begin) return cudaSuccess;
    simt_denominators_kernel<<<(unsigned)(row_end - row_begin), 128, 0, stream>>>(d_masks, d_qmask, row_begin, row_end, d_out);
    count_launch();
    return cudaGetLastError();
}

// per-pair arch entry points (reference src/arch/generic.rs:4-16), one block each
__global__ void __launch_bounds__(256) dot_u16_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b,
                                                      uint16_t* __restrict__ out) {
    __shared__ uint32_t red[8];
    uint32_t acc = 0;
    for (int k = threadIdx.x; k < IRIS_BITS; k += 256) acc += (uint32_t)a[k] * b[k];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t v = 0;
        for (int w = 0; w < 8; ++w) v += red[w];
        *out = (uint16_t)v;
    }
}
__global__ void __launch_bounds__(256) dot_bool_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                                                       uint16_t* __restrict__ out) {
    __shared__ uint32_t red[8];
    uint32_t acc = 0;
    for (int k = threadIdx.x; k < IRIS_LIMBS; k += 256) acc += __popcll(a[k] & b[k]);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t v = 0;
        for (int w = 0; w < 8; ++w) v += red[w];
        *out = (uint16_t)v;
    }
}
cudaError_t launch_dot_u16(const uint16_t* d_a, const uint16_t* d_b, uint16_t* d_out, cudaStream_t stream) {
    dot_u16_kernel<<<1, 256, 0, stream>>>(d_a, d_b, d_out);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_dot_bool(const uint64_t* d_a, const uint64_t* d_b, uint16_t* d_out, cudaStream_t stream) {
    dot_bool_kernel<<<1, 256, 0, stream>>>(d_a, d_b, d_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace iris
