// Micro-benchmark: register -> tensor-memory store throughput (tcgen05.st.32x32b.x32) per SM, as a function of the
// number of storing warps and of how often tcgen05.wait::st is issued.  Diagnostic only:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o sttm_bench sttm_bench.cu && ./sttm_bench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&v)[32]) {
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
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// mode 0: store, wait::st after every `per_wait` stores; mode 1: loads, wait::ld after every `per_wait` loads
template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(int iters, int per_wait, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            (uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 33 + i;
    __syncthreads();
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (uint32_t)(((warp >> 2) * 64 + (it & 1) * 32) & 511);
        if (MODE == 0) {
            st32(tbase + col, v);
            if ((it + 1) % per_wait == 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {
            ld32(tbase + col, v);
            if ((it + 1) % per_wait == 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[5];
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0xDEADBEEF) sink[0] = acc + v[3];
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

int main() {
    long long* d_cycles;
    uint32_t* d_sink;
    cudaMalloc(&d_cycles, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 64);
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {4, 8, 16, 32})
            for (int per_wait : {1, 2, 8}) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (mode == 0) bench<0><<<148, warps * 32>>>(iters, per_wait, d_cycles, d_sink);
                    else bench<1><<<148, warps * 32>>>(iters, per_wait, d_cycles, d_sink);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                }
                long long c[148];
                cudaMemcpy(c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
                const double bytes = (double)iters * warps * 4096.0;
                printf("%s warps=%2d per_wait=%d: %lld cycles, %.1f B/clk/SM, %.1f cycles per warp-instruction\n",
                       mode ? "tcgen05.ld" : "tcgen05.st", warps, per_wait, c[0], bytes / c[0], (double)c[0] / iters);
            }
    return 0;
}
