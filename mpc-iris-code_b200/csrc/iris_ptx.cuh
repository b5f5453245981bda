// Inline-PTX wrappers for sm_100a: mbarrier, bulk async copies (TMA engine, 1-D), tcgen05
// (TMEM alloc, UMMA kind::i8, commit, ld) and proxy fences.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iris {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- warp-uniform role dispatch
// The single-thread instructions of this file (tcgen05.mma / commit, cp.async.bulk) take their operands from
// UNIFORM registers.  If they sit in a `lane == 0` branch the compiler must assume divergent values and wraps
// every one of them in an ELECT / R2UR.BROADCAST waterfall loop (~100 cycles per UMMA).  Keeping the whole warp
// in uniform control flow (warp index via shuffle) and predicating only the instruction itself with elect.sync
// lets the descriptors live in uniform registers.
__device__ __forceinline__ int warp_idx_sync() {
    return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
}
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait with a watchdog: a protocol bug must surface as a trapped launch with an error code,
// never as a hung GPU.  `code` identifies the waiting role.  The flag lives in MAPPED PINNED HOST memory
// (iris_db::h_error), so the code is still readable after the trap has poisoned the context.  The limit is
// wall-clock (globaltimer) and generous: a healthy wait is microseconds, but a time-sliced or replayed (ncu)
// kernel can be descheduled for long stretches.
constexpr unsigned long long kWatchdogNs = 20000000000ull;
__device__ __forceinline__ void watchdog_fire(int* err, int code) {
    if (err) {
        *reinterpret_cast<volatile int*>(err) = code;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 1024; ++i)          // the poll loop proper: no timer read in it
            if (mbar_try_wait(bar, parity)) return;
        const uint64_t now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > kWatchdogNs) watchdog_fire(err, code);
    }
}

// Same, for roles off the critical path (epilogue, producer): sleep between polls so the spinning warp
// leaves its scheduler's issue slots to the working warps.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, int* err, int code) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(64);
        if ((++spins & 0x3FF) == 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kWatchdogNs) watchdog_fire(err, code);
        }
    }
}

#ifdef IRIS_DIAGNOSTICS
// A/B only: try_wait with an explicit suspend-time hint (the hardware parks the warp until the phase completes or the
// hint expires) instead of the default time limit.
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity, int* err, int code, uint32_t hint_ns) {
    uint64_t t0 = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(hint_ns)
            : "memory");
        if (ok) return;
        const uint64_t now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > kWatchdogNs) watchdog_fire(err, code);
    }
}
#endif

// ---------------------------------------------------------------- programmatic dependent launch
// A scan launched with cudaLaunchAttributeProgrammaticStreamSerialization may start its CTAs as soon as every CTA of
// the previous kernel in the stream has called pdl_launch_dependents() -- i.e. on the SMs that kernel's one-tile tail
// leaves idle.  The scans are independent of each other (read-only database, disjoint outputs: the host checks), so
// the dependency wait sits at the END of the kernel: it only keeps completion in stream order.  Both are no-ops in a
// kernel launched the ordinary way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- bulk async copy (TMA engine, 1-D)
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            dst_smem),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- UMMA
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows of 128 bytes:
// start>>4 | LBO (unused for swizzled K-major) | SBO = 1024 B (8 rows) | version 1 | layout 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor, kind::i8: D = S32, A = B = U8 (or S8 for B when b_signed), K-major both, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_i8(uint32_t n, bool a_signed = false, bool b_signed = false) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((n >> 3) << 17) |
           ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued UMMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

}  // namespace ptx
}  // namespace iris

// ================================================================= cluster / 2-CTA (cta_group::2) helpers
namespace iris {
namespace ptx {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default .release.cta semantics: a bare SYNCS.ARRIVE.  (.release.cluster costs MEMBAR.ALL.GPU + ERRBAR per
    // arrive and is not needed here: only async-proxy data -- TMA writes, UMMA reads, TMEM -- crosses CTAs.)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int* err, int code) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if ((++spins & 0x3FF) == 0 && globaltimer_ns() - t0 > kWatchdogNs) watchdog_fire(err, code);
    }
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// Instruction descriptor for cta_group::2, M = 256 (128 rows per CTA).
__host__ __device__ constexpr uint32_t umma_idesc_i8_m256(uint32_t n, bool a_signed = false, bool b_signed = false) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((n >> 3) << 17) |
           ((256u >> 4) << 24);
}
__device__ __forceinline__ void umma_i8_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive (once all prior UMMAs of this thread completed) on the barrier at the same smem offset in
// every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(cta_mask)
        : "memory");
}

}  // namespace ptx
}  // namespace iris
