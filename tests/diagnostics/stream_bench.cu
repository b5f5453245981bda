// Micro-benchmark: what a pure HBM read stream of 1-D bulk copies (cp.async.bulk global -> shared, the TMA engine)
// delivers as a function of the copy size, of the number of sequential streams a CTA interleaves and of the bytes it
// keeps in flight.  One persistent CTA per SM: 1, 2 or 4 producer threads (one per warp, taking the stages round-robin) issue the
// copies into a ring of stages, a consumer thread waits for each stage and hands it back -- nothing touches the data.  Diagnostic only (is the
// floor of the denominators-only scan, two interleaved streams of 4 KiB copies per CTA, the access pattern or the
// barrier protocol?).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream_bench stream_bench.cu && ./stream_bench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
}

// Every CTA reads `streams` sequential streams of `per_stream` bytes each (stream s of CTA b starts at
// (b * streams + s) * per_stream), alternating between them copy by copy; a stage holds one copy of each stream.
__global__ void __launch_bounds__(160, 1) stream_kernel(const uint8_t* __restrict__ src, size_t per_stream, int streams,
                                                        int copy_bytes, int stages, int evict_first, int producers) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = (uint32_t)(streams * copy_bytes);
    const uint32_t bars = base + (uint32_t)stages * stage_bytes;      // full[stages], empty[stages]
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (stages + s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t n_copies = per_stream / copy_bytes;
    const uint8_t* mine = src + (size_t)blockIdx.x * streams * per_stream;
    const int pw = (int)(threadIdx.x >> 5) - 1;               // producer index: lane 0 of warps 1..producers
    if (pw >= 0 && pw < producers && (threadIdx.x & 31) == 0) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        // producer pw takes the stages pw, pw + producers, ... (stages is a multiple of producers)
        int st = pw;
        uint32_t ph = 0;
        for (size_t c = pw; c < n_copies; c += producers) {
            mbar_wait(bars + 8 * (stages + st), ph ^ 1u);
            const uint32_t fb = bars + 8 * st;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(stage_bytes) : "memory");
            for (int s = 0; s < streams; ++s) {
                const uint32_t dst = base + st * stage_bytes + s * copy_bytes;
                const uint8_t* g = mine + (size_t)s * per_stream + c * copy_bytes;
                if (evict_first)
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                        "l"(g), "r"(copy_bytes), "r"(fb), "l"(pol)
                        : "memory");
                else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                                 "l"(g), "r"(copy_bytes), "r"(fb)
                                 : "memory");
            }
            st += producers;
            if (st >= stages) { st -= stages; ph ^= 1u; }
        }
    } else if (threadIdx.x == 0) {
        int st = 0;
        uint32_t ph = 0;
        for (size_t c = 0; c < n_copies; ++c) {
            mbar_wait(bars + 8 * st, ph);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * (stages + st)) : "memory");
            if (++st == stages) { st = 0; ph ^= 1u; }
        }
    }
}

int main(int argc, char** argv) {
    const size_t total = (argc > 1 ? strtoull(argv[1], nullptr, 10) : 8ull) << 30;     // GiB read per launch
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint8_t* src = nullptr;
    if (cudaMalloc(&src, total) != cudaSuccess) return 1;
    cudaMemset(src, 1, total);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%d SMs, %.1f GiB per launch\n", sms, total / 1073741824.0);
    printf("%8s %8s %10s %7s %5s %9s %9s\n", "copy B", "streams", "in flight", "evict1", "prod", "ms", "GB/s");
    const int copies[] = {2048, 4096, 8192, 16384, 32768};
    for (int ef = 1; ef < 2; ++ef)
        for (int streams : {1, 2})
            for (int cb : copies)
              for (int producers : {1, 2, 4})
                for (int flight_kib : {128}) {
                    const int stages = flight_kib * 1024 / (streams * cb);
                    if (stages < 2 || stages > 64 || stages % producers) continue;
                    if (producers > 1 && cb > 8192) continue;
                    size_t per_stream = total / ((size_t)sms * streams);
                    per_stream -= per_stream % cb;
                    const size_t smem = 1024 + (size_t)stages * streams * cb + 16 * stages + 64;
                    float best = 1e30f;
                    for (int rep = 0; rep < 4; ++rep) {
                        cudaEventRecord(e0);
                        stream_kernel<<<sms, 160, smem>>>(src, per_stream, streams, cb, stages, ef, producers);
                        cudaEventRecord(e1);
                        if (cudaEventSynchronize(e1) != cudaSuccess) {
                            printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
                            return 1;
                        }
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        if (rep && ms < best) best = ms;
                    }
                    const double bytes = (double)per_stream * streams * sms;
                    printf("%8d %8d %7d KiB %7d %5d %9.3f %9.0f\n", cb, streams, flight_kib, ef, producers, best, bytes / best / 1e6);
                }
    return 0;
}
