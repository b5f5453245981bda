// Micro-benchmark: cycles per tcgen05.mma as a function of kind, operand source (A in shared or tensor memory),
// M, N and the shared-memory swizzle of B.  One CTA per SM, one thread issues a chain of accumulating UMMAs on
// zero operands and the time to the commit is measured with clock64.  Diagnostic only:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_bench umma_bench.cu && ./umma_bench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

enum Kind { kI8 = 0, kMxf4 = 1, kF8f6f4 = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}

template <int KIND, bool A_TMEM>
__device__ __forceinline__ void umma(uint32_t d, uint32_t a_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t sf) {
    if (KIND == kI8) {
        if (A_TMEM)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
    } else if (KIND == kF8f6f4) {
        if (A_TMEM)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
    } else {
        if (A_TMEM)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%4], [%4], p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(sf) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%4], [%4], p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(sf) : "memory");
    }
}

// nacc: number of accumulators the chain cycles through (1 = every UMMA depends on the previous one's D)
template <int KIND, bool A_TMEM, int NCOMMIT>
__global__ void __launch_bounds__(128, 1) bench(int iters, uint32_t idesc, uint32_t bdesc_hi, uint32_t bdesc_lbo, int nacc, int ncols,
                                                long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    __shared__ uint64_t side_bar[4];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint4* z = reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&side_bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    {   // scale factors 1.0 (0x7F) in columns 448..479, zero A operand in columns 384..447
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 32; ++c) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + 448 + c), "r"(0x7F7F7F7Fu) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + 384 + c), "r"(0u) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + 416 + c), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 0) {
        const uint64_t adesc = (uint64_t)((base & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t bdesc = (uint64_t)(((base + 16384) & 0x3FFFFu) >> 4) | ((uint64_t)bdesc_lbo << 16) | ((uint64_t)bdesc_hi << 32);
        long long t0 = 0;
        if (elect_one()) {
            t0 = clock64();
            const uint32_t sb0 = smem_u32(&side_bar[0]), sb1 = smem_u32(&side_bar[1]), sb2 = smem_u32(&side_bar[2]);
            const uint32_t dstep = nacc > 1 ? (uint32_t)ncols : 0u;
#pragma unroll 1
            for (int it = 0; it < iters; it += 8) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t d = tmem + h * dstep;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        umma<KIND, A_TMEM>(d, tmem + 384 + u * 8, adesc + 2 * u, bdesc + 2 * u, idesc, tmem + 448);
                    if (NCOMMIT >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sb0) : "memory");
                    if (NCOMMIT >= 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sb1) : "memory");
                    if (NCOMMIT >= 3) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sb2) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (t0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Round trip: issue `group` UMMAs, commit, poll the barrier; repeated.  Gives issue -> completion-visible latency.
template <int KIND, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) latency(int rounds, int group, uint32_t idesc, uint32_t bdesc_hi, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint4* z = reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    {
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 32; ++c) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + 448 + c), "r"(0x7F7F7F7Fu) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + 384 + c), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 0) {
        const uint64_t adesc = (uint64_t)((base & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t bdesc = (uint64_t)(((base + 16384) & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)bdesc_hi << 32);
        const uint32_t b = smem_u32(&bar);
        const long long t0 = clock64();
        uint32_t parity = 0;
        for (int r = 0; r < rounds; ++r) {
            if (elect_one()) {
                for (int u = 0; u < group; ++u) umma<KIND, A_TMEM>(tmem, tmem + 384, adesc, bdesc, idesc, tmem + 448);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
            }
            __syncwarp();
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
            parity ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;");
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int KIND, bool A_TMEM>
static void run_latency(const char* name, int n, int group, long long* d_cycles) {
    const int m = 128;
    uint32_t idesc;
    if (KIND == kI8) idesc = (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    else idesc = (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const int rounds = 512, smem = 65 * 1024 + 1024;
    cudaFuncSetAttribute(latency<KIND, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) latency<KIND, A_TMEM><<<148, 128, smem>>>(rounds, group, idesc, hi, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("latency %s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    long long h[148];
    cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
    printf("latency %-5s %s N=%d group=%d UMMAs + commit + poll : %7.1f clk per round trip\n", name, A_TMEM ? "TS" : "SS", n, group,
           (double)h[0] / rounds);
    fflush(stdout);
}

struct Swz { const char* name; uint32_t layout, sbo, lbo; };

template <int KIND, bool A_TMEM, int NCOMMIT = 0>
static void run(const char* name, int m, int n, const Swz& sw, int nacc, long long* d_cycles) {
    const int ncommit = NCOMMIT;
    uint32_t idesc;
    if (KIND == kI8) idesc = (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    else if (KIND == kF8f6f4) idesc = (1u << 4) | (5u << 7) | (5u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);   // e2m1 x e2m1 -> f32
    else idesc = (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
    const uint32_t hi = (sw.sbo >> 4) | (1u << 14) | (sw.layout << 29);
    const int iters = 4096, smem = 65 * 1024 + 1024;
    cudaFuncSetAttribute(bench<KIND, A_TMEM, NCOMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int ncols = n < 32 ? 32 : n;
    if (nacc * ncols > 384) nacc = 384 / ncols;
    for (int rep = 0; rep < 2; ++rep) bench<KIND, A_TMEM, NCOMMIT><<<148, 128, smem>>>(iters, idesc, hi, sw.lbo >> 4, nacc, ncols, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-10s %s M=%d N=%d %s: %s\n", name, A_TMEM ? "TS" : "SS", m, n, sw.name, cudaGetErrorString(e)); exit(1); }
    long long h[148];
    cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1ll << 60;
    for (int i = 0; i < 148; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
    printf("%-7s %s M=%3d N=%3d B=%-5s acc=%d commits/4=%d : %6.1f clk/UMMA (min SM %.1f)\n", name, A_TMEM ? "TS" : "SS", m, n, sw.name, nacc, ncommit,
           (double)mx / iters, (double)mn / iters);
    fflush(stdout);
}

int main() {
    long long* d_cycles;
    cudaMalloc(&d_cycles, 148 * sizeof(long long));
    const Swz sw128{"sw128", 2, 1024, 16}, sw64{"sw64", 4, 512, 16}, sw32{"sw32", 6, 256, 16}, none{"none", 0, 256, 128};
    for (int n : {16, 32, 64, 128, 256}) run<kI8, true, 0>("i8", 128, n, sw128, 1, d_cycles);
    for (int n : {16, 32, 64, 128, 256}) run<kMxf4, true, 0>("mxf4", 128, n, sw128, 1, d_cycles);
    for (int n : {16, 32, 64}) run<kMxf4, true, 0>("mxf4", 128, n, sw128, 2, d_cycles);
    run<kI8, true, 0>("i8", 64, 32, sw128, 1, d_cycles);
    run<kI8, false, 0>("i8", 128, 32, sw128, 1, d_cycles);
    run<kMxf4, false, 0>("mxf4", 128, 32, sw128, 1, d_cycles);
    run<kMxf4, true, 1>("mxf4", 128, 32, sw128, 1, d_cycles);
    run<kMxf4, true, 2>("mxf4", 128, 32, sw128, 1, d_cycles);
    run<kMxf4, true, 3>("mxf4", 128, 32, sw128, 1, d_cycles);
    run<kMxf4, true, 3>("mxf4", 128, 32, sw128, 2, d_cycles);
    run<kI8, true, 3>("i8", 128, 32, sw128, 2, d_cycles);
    for (int g : {0, 1, 2, 4, 8}) run_latency<kMxf4, true>("mxf4", 32, g, d_cycles);
    for (int g : {1, 4}) run_latency<kI8, true>("i8", 32, g, d_cycles);
    for (int g : {1, 4}) run_latency<kI8, false>("i8", 32, g, d_cycles);
    (void)sw64; (void)sw32; (void)none;
    return 0;
}
