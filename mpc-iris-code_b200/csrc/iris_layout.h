// HBM / shared-memory layouts of the B200 matching path.  See DESIGN.md "Data layout".
//
// Reference types (byte-for-byte at the C ABI):
//   EncodedBits = [u16; 12800]  (src/encoded_bits.rs:13-15), index k = row*200 + col
//   Bits        = [u64; 200]    (src/bits.rs:13-15), bit k = byte k/8, bit k%8
// In HBM the database is re-tiled once at load time into the exact image the tensor-core
// pipeline consumes, so the hot loop is pure contiguous bulk copies:
//
//   shares : [tile][chunk c][plane lo|hi][128 rows][128 B, SWIZZLE_128B]   (32 KiB per (tile,c))
//            plane byte (r, b) = lo/hi byte of db[tile*128 + r][c*128 + b]
//   masks  : [tile][chunk c][128 rows][16 B]                               (2 KiB per (tile,c))
//            = bytes [16c, 16c+16) of each row's 1600 raw Bits bytes
//   qd     : [chunk c][plane q_lo|q_hi][32 rotations][128 B, SWIZZLE_128B]  (8 KiB per c)
//            plane byte (j, b) = lo/hi byte of rot(q, j-15)[c*128 + b]; j = 31 is zero padding
//   qm     : [chunk c][32 rotations][128 B, SWIZZLE_128B]                   (4 KiB per c)
//            byte (j, e): w=e>>5, t=(e>>2)&7, m=e&3, source bit s = c*128 + 32w + 8m + t,
//            value = rot(qmask, j-15)[s] << (7 - t)
//
// SWIZZLE_128B: within a tile of 128-byte rows whose base is 1024-byte aligned, the 16-byte
// chunk index of byte (r, b) is (b >> 4) ^ (r & 7).
#pragma once
#include <stdint.h>

#define IRIS_COLS 200
#define IRIS_ROWS 64
#define IRIS_BITS 12800
#define IRIS_LIMBS 200
#define IRIS_MASK_BYTES 1600
#define IRIS_ROTATIONS 31

namespace iris {

constexpr int kTileRows = 128;                 // database rows per MMA tile (UMMA M)
constexpr int kChunkK = 128;                   // K elements per pipeline stage (one swizzle span)
constexpr int kChunks = IRIS_BITS / kChunkK;   // 100
constexpr int kPlaneTileBytes = kTileRows * kChunkK;    // 16384
constexpr int kShareChunkBytes = 2 * kPlaneTileBytes;   // lo | hi
constexpr int kMaskChunkBytes = kTileRows * 16;         // 2048 packed mask bytes
constexpr int kQTileBytes = 32 * kChunkK;               // 4096
constexpr int kQdChunkBytes = 2 * kQTileBytes;          // q_lo | q_hi
constexpr int kQmChunkBytes = kQTileBytes;
constexpr size_t kShareTileBytes = (size_t)kChunks * kShareChunkBytes;   // 3,276,800 = 128 * 25600
constexpr size_t kMaskTileBytes = (size_t)kChunks * kMaskChunkBytes;     // 204,800   = 128 * 1600
constexpr size_t kQdBytes = (size_t)kChunks * kQdChunkBytes;             // 819,200
constexpr size_t kQmBytes = (size_t)kChunks * kQmChunkBytes;             // 409,600
constexpr int kQm4StageBytes = 32 * 128;                                 // 4-bit mask operand: 32 rotations x 256 nibbles
constexpr size_t kQm4Bytes = (size_t)(IRIS_BITS / 256) * kQm4StageBytes; // 204,800
constexpr int kOutRowBytes = IRIS_ROTATIONS * 2;                         // 62

__host__ __device__ inline uint32_t swz128(uint32_t r, uint32_t b) {
    return r * 128u + ((((b >> 4) ^ (r & 7u)) << 4) | (b & 15u));
}

// byte offset of share plane byte for global row R, flat element k
__host__ __device__ inline size_t share_offset(uint64_t R, uint32_t k, uint32_t plane) {
    uint64_t tile = R / kTileRows;
    uint32_t r = (uint32_t)(R % kTileRows);
    uint32_t c = k / kChunkK, b = k % kChunkK;
    return ((tile * kChunks + c) * 2 + plane) * (size_t)kPlaneTileBytes + swz128(r, b);
}

// byte offset of raw mask byte `byte` (0..1599) of global row R
__host__ __device__ inline size_t mask_offset(uint64_t R, uint32_t byte) {
    uint64_t tile = R / kTileRows;
    uint32_t r = (uint32_t)(R % kTileRows);
    uint32_t c = byte / 16, b = byte % 16;
    return (tile * kChunks + c) * (size_t)kMaskChunkBytes + r * 16u + b;
}

__host__ __device__ inline uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---- synthetic data spec (the device generators in iris_kernels.cu follow it; the test checker restates it) ----
constexpr uint64_t kGenMul = 0xD1342543DE82EF95ull;
constexpr uint64_t kGenMaskTag = 0xA5A5A5A55A5A5A5Aull;      // mask limb l of row R: mix64((seed ^ tag) ^ ((R*200+l) * kGenMul))
constexpr uint64_t kGenPatternTag = 0x5A5A5A5AA5A5A5A5ull;   // pattern limb, same form
// four u16 share elements 4g..4g+3 of row R: the 16-bit fields of mix64(seed ^ ((R*3200+g) * kGenMul))
__host__ __device__ inline uint64_t gen_share_group(uint64_t seed, uint64_t R, uint64_t g) {
    return mix64(seed ^ ((R * (IRIS_BITS / 4) + g) * kGenMul));
}
__host__ __device__ inline uint64_t gen_bits_limb(uint64_t seed, uint64_t tag, uint64_t R, uint64_t l) {
    return mix64((seed ^ tag) ^ ((R * IRIS_LIMBS + l) * kGenMul));
}
// seed of the uniform share vectors of party p (p < n_parties - 1)
__host__ __device__ inline uint64_t gen_party_seed(uint64_t seed, uint32_t p) {
    return mix64(seed ^ (0xC2B2AE3D27D4EB4Full * (uint64_t)(p + 1)));
}

}  // namespace iris
