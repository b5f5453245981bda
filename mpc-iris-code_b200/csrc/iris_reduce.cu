// Coordinator-side reduction on the device:
//   numerator[r] = wrapping sum over parties of the distance shares          src/main.rs:603-608
//   decode_distance: min_r ((den - num) as u16 / 2) as f64 / den as f64       src/lib.rs:97-107
//                    (f64::min ignores NaN, so den = 0 with num = 0 is skipped; x/0 = +inf)
//   running min / argmin over rows with `distance < min_distance`             src/main.rs:611-621
//                    (strict <: the FIRST row attaining the minimum wins)
// IEEE-754 double division on the device is correctly rounded, so the result is bit-identical
// to the CPU's.  Collapses 124 bytes per row to 16 bytes per query, which is what makes the
// multi-GPU exchange "small vectors" (DESIGN.md section 7).
#include <cuda_runtime.h>
#include <math_constants.h>

#include "iris_kernels.cuh"

namespace iris {

void count_launch_external();

constexpr int kReduceThreads = 256;

__device__ __forceinline__ void min_pair(double& v, unsigned long long& i, double ov, unsigned long long oi) {
    if (ov < v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
    }
}

__global__ void __launch_bounds__(kReduceThreads) combine_decode_kernel(const CombineParams p, double* __restrict__ block_min,
                                                                        unsigned long long* __restrict__ block_idx) {
    __shared__ double s_v[kReduceThreads / 32];
    __shared__ unsigned long long s_i[kReduceThreads / 32];
    // The 62-byte rows are staged through shared memory with coalesced 16-byte loads (a thread reading its own
    // row straight from global memory touches a 2 KiB span per warp instruction and runs at ~1 TB/s).
    __shared__ __align__(16) uint16_t s_num[kReduceThreads * IRIS_ROTATIONS];
    __shared__ __align__(16) uint16_t s_den[kReduceThreads * IRIS_ROTATIONS];
    // blockIdx.y = query of a batch ([Q][n][31] arrays, query_stride elements apart); 0 for a single query
    const size_t qoff = (size_t)blockIdx.y * p.query_stride;
    const uint64_t row0 = (uint64_t)blockIdx.x * kReduceThreads;
    const uint64_t row = row0 + threadIdx.x;
    const uint32_t nrows = (uint32_t)(p.n - row0 < (uint64_t)kReduceThreads ? p.n - row0 : kReduceThreads);
    const uint32_t elems = nrows * IRIS_ROTATIONS;
    const uint32_t vecs = elems / 8;                           // whole 16-byte vectors of the block's slice
    {
        const uint16_t* g = p.denominators + qoff + row0 * IRIS_ROTATIONS;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            for (uint32_t i = threadIdx.x; i < vecs; i += kReduceThreads)
                reinterpret_cast<uint4*>(s_den)[i] = reinterpret_cast<const uint4*>(g)[i];
            for (uint32_t i = vecs * 8 + threadIdx.x; i < elems; i += kReduceThreads) s_den[i] = g[i];
        } else {
            for (uint32_t i = threadIdx.x; i < elems; i += kReduceThreads) s_den[i] = g[i];
        }
    }
    for (uint32_t q = 0; q < p.parties; ++q) {                 // numerator = wrapping sum of the parties' shares
        if (q) __syncthreads();                                // element ownership differs between the two paths
        const uint16_t* g = p.shares[q] + qoff + row0 * IRIS_ROTATIONS;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            for (uint32_t i = threadIdx.x; i < vecs; i += kReduceThreads) {
                uint4 v = reinterpret_cast<const uint4*>(g)[i];
                if (q) {
                    const uint4 a = reinterpret_cast<uint4*>(s_num)[i];
                    v.x = __vadd2(v.x, a.x);                   // per-halfword wrapping add
                    v.y = __vadd2(v.y, a.y);
                    v.z = __vadd2(v.z, a.z);
                    v.w = __vadd2(v.w, a.w);
                }
                reinterpret_cast<uint4*>(s_num)[i] = v;
            }
            for (uint32_t i = vecs * 8 + threadIdx.x; i < elems; i += kReduceThreads)
                s_num[i] = (uint16_t)((q ? s_num[i] : 0) + g[i]);
        } else {
            for (uint32_t i = threadIdx.x; i < elems; i += kReduceThreads) s_num[i] = (uint16_t)((q ? s_num[i] : 0) + g[i]);
        }
    }
    __syncthreads();
    double best = CUDART_INF;
    unsigned long long idx = ~0ull;
    if (row < p.n) {
        // decode_distance (src/lib.rs:97-107) folds f64::min over the 31 quotients num / den.  f64 division is the
        // slowest thing this GPU does, so the minimum is found on the exact fractions (num_a * den_b < num_b * den_a,
        // products below 2^32) and only the winner is divided.  Same result bit for bit: two different fractions with
        // denominators below 2^16 are more than 2^-32 apart and correctly rounded division is monotonic, so the
        // smallest fraction is the smallest quotient.  den = 0 gives NaN (num = 0: dropped by f64::min) or +inf
        // (never below the fold's initial INFINITY), so those rotations cannot win either way.
        uint32_t bn = 0, bd = 0;                               // bd = 0: no finite quotient yet
#pragma unroll
        for (int j = 0; j < IRIS_ROTATIONS; ++j) {
            const uint16_t s = s_num[threadIdx.x * IRIS_ROTATIONS + j];
            const uint32_t d = s_den[threadIdx.x * IRIS_ROTATIONS + j];
            const uint32_t num = (uint16_t)((uint16_t)(d - s) >> 1);
            if (d != 0 && (bd == 0 || num * bd < bn * d)) {
                bn = num;
                bd = d;
            }
        }
        if (bd) best = (double)bn / (double)bd;
        idx = p.index_base + row;
        if (p.distances_out) p.distances_out[(size_t)blockIdx.y * p.n + row] = best;
    }
    for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const unsigned long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
        min_pair(best, idx, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) {
        s_v[threadIdx.x >> 5] = best;
        s_i[threadIdx.x >> 5] = idx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kReduceThreads / 32; ++w) min_pair(best, idx, s_v[w], s_i[w]);
        block_min[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = best;
        block_idx[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = idx;
    }
}

// block q reduces the n per-block pairs of query q into result pair q ({f64 min, u64 index}, 16 bytes apart)
__global__ void __launch_bounds__(kReduceThreads) final_min_kernel(const double* __restrict__ block_min,
                                                                   const unsigned long long* __restrict__ block_idx, uint32_t n,
                                                                   ResultPair* __restrict__ out) {
    __shared__ double s_v[kReduceThreads / 32];
    __shared__ unsigned long long s_i[kReduceThreads / 32];
    double best = CUDART_INF;
    unsigned long long idx = ~0ull;
    block_min += (size_t)blockIdx.x * n;
    block_idx += (size_t)blockIdx.x * n;
    for (uint32_t k = threadIdx.x; k < n; k += kReduceThreads) min_pair(best, idx, block_min[k], block_idx[k]);
    for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const unsigned long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
        min_pair(best, idx, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) {
        s_v[threadIdx.x >> 5] = best;
        s_i[threadIdx.x >> 5] = idx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kReduceThreads / 32; ++w) min_pair(best, idx, s_v[w], s_i[w]);
        out[blockIdx.x].min_distance = best;
        out[blockIdx.x].min_index = best < CUDART_INF ? idx : ~0ull;   // nothing was `< INFINITY`: min_index stays usize::MAX (main.rs:582)
    }
}

// out[q] = best of in[s * stride + q], s < n_sets: the running min over the slices of one shard, and over the shards of a
// cluster (src/main.rs:611-621 continued across blocks of rows: strict `<`, so the lowest row wins a tie).  `out` may be
// peer memory of another GPU (NVLink stores) or mapped host memory; the pairs are tiny.
__global__ void merge_pairs_kernel(const ResultPair* __restrict__ in, uint32_t n_sets, uint32_t stride, uint32_t n_queries,
                                   ResultPair* __restrict__ out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    double best = CUDART_INF;
    unsigned long long idx = ~0ull;
    for (uint32_t s = 0; s < n_sets; ++s) {
        const ResultPair r = in[(size_t)s * stride + q];
        if (r.min_index != ~0ull) min_pair(best, idx, r.min_distance, r.min_index);
    }
    out[q].min_distance = best;
    out[q].min_index = best < CUDART_INF ? idx : ~0ull;
}

size_t combine_scratch_bytes(uint64_t n) {
    const uint64_t blocks = (n + kReduceThreads - 1) / kReduceThreads;
    return blocks * (sizeof(double) + sizeof(unsigned long long)) + 64;
}

// scratch: combine_scratch_bytes(n) device bytes per query; result: device ResultPair per query.  num_queries > 1 reads
// [Q][n][31] arrays (p.query_stride elements apart) and takes two launches for the whole batch.
cudaError_t launch_combine_min(const CombineParams& p, void* scratch, void* result, cudaStream_t stream, uint32_t num_queries) {
    if (num_queries == 0) return cudaSuccess;
    const uint32_t blocks = (uint32_t)((p.n + kReduceThreads - 1) / kReduceThreads);
    double* bmin = static_cast<double*>(scratch);
    unsigned long long* bidx = reinterpret_cast<unsigned long long*>(bmin + (size_t)blocks * num_queries);
    if (blocks) {
        combine_decode_kernel<<<dim3(blocks, num_queries), kReduceThreads, 0, stream>>>(p, bmin, bidx);
        count_launch_external();
    }
    final_min_kernel<<<num_queries, kReduceThreads, 0, stream>>>(bmin, bidx, blocks, static_cast<ResultPair*>(result));
    count_launch_external();
    return cudaGetLastError();
}

cudaError_t launch_final_min(const double* block_min, const unsigned long long* block_idx, uint32_t n, ResultPair* result,
                             cudaStream_t stream) {
    final_min_kernel<<<1, kReduceThreads, 0, stream>>>(block_min, block_idx, n, result);
    count_launch_external();
    return cudaGetLastError();
}

cudaError_t launch_merge_pairs(const ResultPair* in, uint32_t n_sets, uint32_t stride, uint32_t n_queries, ResultPair* out,
                               cudaStream_t stream) {
    if (n_queries == 0) return cudaSuccess;
    merge_pairs_kernel<<<(n_queries + 127) / 128, 128, 0, stream>>>(in, n_sets, stride, n_queries, out);
    count_launch_external();
    return cudaGetLastError();
}

}  // namespace iris
