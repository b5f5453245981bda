// Arch-level batched dots: the GPU form of the reference's criterion grid (src/arch/mod.rs:22-72), which calls
// dot_u16 / dot_bool (src/arch/generic.rs:4-16) on every pair of `a` INDEPENDENT vectors and `b` database vectors.
//
// A group of up to 31 arbitrary vectors takes the place of the 31 rotations of one query: the operand images below have
// exactly the layout the scan and the batched GEMM kernels consume (iris_layout.h), slot j holding vector j instead of
// rot(q, j - 15), so out[i][j] = dot(a[j], b[i]) comes out of the same tensor-core passes.  compact_columns_kernel
// gathers the [group][b][31] results into the caller's [b][a] array.
#include <cuda_runtime.h>

#include "iris_kernels.cuh"

namespace iris {

void count_launch_external();

// qd image ([chunk c][q_lo | q_hi][32 slots][128 B, SWIZZLE_128B]) of up to 31 u16 vectors; unused slots are zero.
__global__ void prep_distance_vectors_kernel(const uint16_t* __restrict__ a, uint32_t n_vec, uint8_t* __restrict__ qd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (c, j, ch)
    if (idx >= kChunks * 32 * 8) return;
    const int ch = idx & 7, j = (idx >> 3) & 31, c = idx >> 8;
    uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    if ((uint32_t)j < n_vec) {
        const uint16_t* v = a + (size_t)j * IRIS_BITS + c * kChunkK + ch * 16;
        for (int b = 0; b < 16; ++b) {
            const uint32_t x = v[b];
            lo[b >> 2] |= (x & 0xFFu) << (8 * (b & 3));
            hi[b >> 2] |= (x >> 8) << (8 * (b & 3));
        }
    }
    const size_t off = (size_t)c * kQdChunkBytes + j * 128 + ((ch ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(qd + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(qd + off + kQTileBytes) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

// *flag (preset to 1) is cleared unless every element of the n_vec vectors is a sign-extended byte.
__global__ void classify_vectors_s8_kernel(const uint16_t* __restrict__ a, uint32_t n, int* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && (uint16_t)(a[k] + 0x80u) > 0xFFu) *flag = 0;
}

// int8 mask image ([chunk c][32 slots][128 B], bit value 2^(7-t), permuted K -- see iris_layout.h) and the 4-bit image
// ([stage s][32 slots][128 B], e2m1 2.0 / 1.0 / 0.5 / 0.5 -- see iris_maskscan4.cu) of up to 31 Bits vectors.
__global__ void prep_mask_vectors_kernel(const uint8_t* __restrict__ a, uint32_t n_vec, uint8_t* __restrict__ qm) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (c, j, ch)
    if (idx >= kChunks * 32 * 8) return;
    const int ch = idx & 7, j = (idx >> 3) & 31, c = idx >> 8;
    uint32_t out[4] = {0, 0, 0, 0};
    if ((uint32_t)j < n_vec) {
        const uint8_t* v = a + (size_t)j * IRIS_MASK_BYTES;
        for (int b = 0; b < 16; ++b) {
            const int e = ch * 16 + b;
            const int w = e >> 5, t = (e >> 2) & 7, m = e & 3;
            const int s = c * kChunkK + 32 * w + 8 * m + t;
            const uint32_t bit = (v[s >> 3] >> (s & 7)) & 1u;
            out[b >> 2] |= (bit << (7 - t)) << (8 * (b & 3));
        }
    }
    const size_t off = (size_t)c * kQmChunkBytes + j * 128 + ((ch ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(qm + off) = make_uint4(out[0], out[1], out[2], out[3]);
}
__global__ void prep_mask_vectors_fp4_kernel(const uint8_t* __restrict__ a, uint32_t n_vec, uint8_t* __restrict__ qm4) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (s, r, ch)
    if (idx >= (IRIS_BITS / 256) * 32 * 8) return;
    const int ch = idx & 7, r = (idx >> 3) & 31, s = idx >> 8;
    uint32_t out[4] = {0, 0, 0, 0};
    if ((uint32_t)r < n_vec) {
        const uint8_t* v = a + (size_t)r * IRIS_MASK_BYTES;
        for (int t = 0; t < 4; ++t) {
            const uint32_t weight = t == 0 ? 4u : (t == 1 ? 2u : 1u);     // e2m1 codes of 2.0, 1.0, 0.5
            for (int j = 0; j < 8; ++j) {
                const int src = 256 * s + 32 * ch + 4 * j + t;
                const uint32_t bit = (v[src >> 3] >> (src & 7)) & 1u;
                out[t] |= (bit * weight) << (4 * j);
            }
        }
    }
    const size_t off = (size_t)s * kQm4StageBytes + r * 128 + ((ch ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(qm4 + off) = make_uint4(out[0], out[1], out[2], out[3]);
}

// in = [groups][n][31] u16 (group g holds vectors 31 g .. 31 g + 30), out = [n][n_vec] u16.
__global__ void compact_columns_kernel(const uint16_t* __restrict__ in, uint64_t n, uint32_t n_vec, uint16_t* __restrict__ out) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row i, vector j)
    if (idx >= n * n_vec) return;
    const uint64_t i = idx / n_vec;
    const uint32_t j = (uint32_t)(idx % n_vec);
    out[idx] = in[((uint64_t)(j / IRIS_ROTATIONS) * n + i) * IRIS_ROTATIONS + j % IRIS_ROTATIONS];
}

cudaError_t launch_prep_distance_vectors(const uint16_t* d_a, uint32_t n_vec, uint8_t* d_qd, int* d_flag, cudaStream_t stream) {
    if (n_vec == 0 || n_vec > IRIS_ROTATIONS) return cudaErrorInvalidValue;
    prep_distance_vectors_kernel<<<(kChunks * 32 * 8 + 255) / 256, 256, 0, stream>>>(d_a, n_vec, d_qd);
    count_launch_external();
    if (d_flag) {
        classify_vectors_s8_kernel<<<(n_vec * IRIS_BITS + 255) / 256, 256, 0, stream>>>(d_a, n_vec * IRIS_BITS, d_flag);
        count_launch_external();
    }
    return cudaGetLastError();
}

cudaError_t launch_prep_mask_vectors(const uint8_t* d_a, uint32_t n_vec, uint8_t* d_qm, uint8_t* d_qm4, cudaStream_t stream) {
    if (n_vec == 0 || n_vec > IRIS_ROTATIONS) return cudaErrorInvalidValue;
    prep_mask_vectors_kernel<<<(kChunks * 32 * 8 + 255) / 256, 256, 0, stream>>>(d_a, n_vec, d_qm);
    count_launch_external();
    prep_mask_vectors_fp4_kernel<<<((IRIS_BITS / 256) * 32 * 8 + 255) / 256, 256, 0, stream>>>(d_a, n_vec, d_qm4);
    count_launch_external();
    return cudaGetLastError();
}

cudaError_t launch_compact_columns(const uint16_t* d_in, uint64_t n, uint32_t n_vec, uint16_t* d_out, cudaStream_t stream) {
    if (n == 0 || n_vec == 0) return cudaSuccess;
    const uint64_t total = n * n_vec;
    compact_columns_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_in, n, n_vec, d_out);
    count_launch_external();
    return cudaGetLastError();
}

}  // namespace iris
