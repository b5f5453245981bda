// Internal launch API between the C-ABI host runtime (iris_abi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "iris_layout.h"

namespace iris {

struct ScanParams {
    const uint8_t* shares;   // tiled share planes (nullptr when not scanning distances)
    const uint8_t* masks;    // tiled packed masks (nullptr when not scanning denominators)
    const uint8_t* qd;       // prepared distance operand image (kQdBytes)
    const uint8_t* qm;       // prepared mask operand image (kQmBytes)
    const uint8_t* qm4;      // 4-bit mask operand image (kQm4Bytes) for the denominators-only scan, or nullptr
    uint16_t* dist_out;      // [row_end-row_begin][31] u16, row-major packed (62 B rows)
    uint16_t* den_out;       // same
    int32_t* raw_out;        // optional debug dump: [tiles*128][128] raw s32 accumulators
    uint64_t row_begin;      // rows outside [row_begin,row_end) are computed but not stored
    uint64_t row_end;
    uint32_t tile_begin;     // = row_begin / 128
    uint32_t tile_end;       // = ceil(row_end / 128)
    int* error;              // device int, set by the watchdog
    bool signed_query;       // query elements are sign-extended bytes: two-product path (q_lo plane as s8)
    bool pdl;                // launch with programmatic stream serialization (may overlap the previous scan's tail)
    // Search mode (fused scan only, both red_* set): nothing is stored per row.  The epilogue decodes each row
    // (src/lib.rs:97-107) and keeps a running (min distance, row) per CTA (src/main.rs:611-621); CTA b writes its pair to
    // red_min[b] / red_idx[b] (gridDim.x entries, reduced afterwards by launch_final_min).  Row ids = index_base + row.
    double* red_min;
    unsigned long long* red_idx;
    uint64_t index_base;
};
// CTAs a scan launch uses (= entries of red_min / red_idx a search-mode launch writes).
uint32_t scan_grid(const ScanParams& p, int num_sms);

// Launches the persistent tcgen05 scan.  Mode is derived from which of shares/masks is non-null.
cudaError_t launch_scan(const ScanParams& p, int num_sms, cudaStream_t stream);
// Denominators-only scan with the expanded operand in tensor memory (iris_maskscan.cu).
cudaError_t launch_mask_scan(const ScanParams& p, int num_sms, cudaStream_t stream);
// The same scan with e2m1 operands (kind::mxf4, iris_maskscan4.cu); p.qm must be the 4-bit image.
cudaError_t launch_mask_scan_fp4(const ScanParams& p, int num_sms, cudaStream_t stream);
cudaError_t launch_prep_mask_query_fp4(const uint8_t* d_qmask, uint8_t* d_qm4, cudaStream_t stream);
// Four query masks per pass over the same expanded database operand (N = 128 UMMAs; iris_maskscan4.cu): the batched
// denominators path.  qm4[i] are 4-bit operand images, out[i] = [row_end-row_begin][31] u16 each.
constexpr int kMaskMultiQueries = 4;
struct MultiMaskScanParams {
    const uint8_t* masks;
    const uint8_t* qm4[kMaskMultiQueries];
    uint16_t* out[kMaskMultiQueries];
    uint64_t row_begin, row_end;
    uint32_t tile_begin, tile_end;
    int* error;
};
cudaError_t launch_mask_scan_fp4_multi(const MultiMaskScanParams& p, int num_sms, cudaStream_t stream);

// Query preparation (K3): reference DistanceEngine::new / MasksEngine::new (src/lib.rs:33-40, 60-67).
cudaError_t launch_encode(const uint8_t* d_pattern, const uint8_t* d_mask, uint16_t* d_out, cudaStream_t stream);
cudaError_t launch_prep_distance_query(const uint16_t* d_query, uint8_t* d_qd, cudaStream_t stream);
// Batched preparation from wire Templates ({pattern, mask}, 3 200 B each, device): encode + both operand images.
constexpr int kMaxPrepBatch = 64;
struct PrepBatchParams {
    const uint8_t* templates;            // [n][3200] device
    uint16_t* query[kMaxPrepBatch];      // encoded query of each engine
    uint8_t* qd[kMaxPrepBatch];          // distance operand images
    uint8_t* qm[kMaxPrepBatch];          // mask operand images (all null: skip)
    uint32_t n;
};
cudaError_t launch_prep_batch(const PrepBatchParams& p, cudaStream_t stream);
// 4-bit images of the same batch (iris_maskscan4.cu): p.qm[i] + kQmBytes receives the image of query i.
cudaError_t launch_prep_mask_fp4_batch(const PrepBatchParams& p, cudaStream_t stream);
// Both mask operand images of one query: the int8 image (fused scan) and the 4-bit image (denominators-only scan).
cudaError_t launch_prep_mask_query(const uint8_t* d_qmask, uint8_t* d_qm, uint8_t* d_qm4, cudaStream_t stream);
// *d_flag (preset to 1) is cleared unless every element of the query is a sign-extended byte.
cudaError_t launch_classify_s8(const uint16_t* d_query, int* d_flag, cudaStream_t stream);

// Loader: reference-layout rows (device staging) -> tiled HBM image, starting at global row row0.
cudaError_t launch_retile_shares(const uint16_t* d_rows, uint64_t n, uint8_t* d_shares, uint64_t row0,
                                 cudaStream_t stream);
cudaError_t launch_retile_masks(const uint8_t* d_rows, uint64_t n, uint8_t* d_masks, uint64_t row0,
                                cudaStream_t stream);
// Inverse of the loader (tests, debugging): tiled image -> reference-layout rows.
cudaError_t launch_untile_shares(const uint8_t* d_shares, uint64_t row0, uint64_t n, uint16_t* d_rows,
                                 cudaStream_t stream);
cudaError_t launch_untile_masks(const uint8_t* d_masks, uint64_t row0, uint64_t n, uint8_t* d_rows,
                                cudaStream_t stream);
// Synthetic database fill directly in the tiled layout; row ids row_id0.. stored at rows row0..
cudaError_t launch_generate(uint8_t* d_shares, uint8_t* d_masks, uint64_t seed, uint64_t row_id0, uint64_t row0,
                            uint64_t n, cudaStream_t stream);

// The same database as iris_db_generate's masks, but with shares that MEAN something: row id R is the synthetic
// Template (pattern_R, mask_R); its encoding (src/lib.rs:16-26) is split into n_parties additive shares as
// EncodedBits::share does (src/encoded_bits.rs:23-38: n-1 uniform vectors, the last = encoding - their sum).  This
// writes party `party`'s share rows; with n_parties = 1 the "share" is the plaintext encoding itself.
cudaError_t launch_generate_party_shares(uint8_t* d_shares, uint64_t seed, uint32_t party, uint32_t n_parties,
                                         uint64_t row_id0, uint64_t row0, uint64_t n, cudaStream_t stream);

// Arch-level batched dots (iris_dotbatch.cu): operand images of up to 31 arbitrary vectors in the slots of the 31
// rotations, and the gather of [groups][n][31] results into [n][n_vec].
cudaError_t launch_prep_distance_vectors(const uint16_t* d_a, uint32_t n_vec, uint8_t* d_qd, int* d_flag, cudaStream_t stream);
cudaError_t launch_prep_mask_vectors(const uint8_t* d_a, uint32_t n_vec, uint8_t* d_qm, uint8_t* d_qm4, cudaStream_t stream);
cudaError_t launch_compact_columns(const uint16_t* d_in, uint64_t n, uint32_t n_vec, uint16_t* d_out, cudaStream_t stream);

// CUDA-core cross-check kernels over the same tiled image (not the product path; used to verify
// the tensor path at full size on the GPU and to serve the per-pair arch entry points).
cudaError_t launch_simt_distances(const uint8_t* d_shares, const uint16_t* d_query, uint64_t row_begin,
                                  uint64_t row_end, uint16_t* d_out, cudaStream_t stream);
cudaError_t launch_simt_denominators(const uint8_t* d_masks, const uint8_t* d_qmask, uint64_t row_begin,
                                     uint64_t row_end, uint16_t* d_out, cudaStream_t stream);
cudaError_t launch_dot_u16(const uint16_t* d_a, const uint16_t* d_b, uint16_t* d_out, cudaStream_t stream);
cudaError_t launch_dot_bool(const uint64_t* d_a, const uint64_t* d_b, uint16_t* d_out, cudaStream_t stream);

// Batched distances (iris_batch.cu): up to kMaxBatchQueries prepared queries against rows
// [row_begin,row_end) in one dense int8 GEMM; out = [num_queries][row_end-row_begin][31] u16 (device).
constexpr int kMaxBatchQueries = 64;
struct BatchParams {
    const uint8_t* shares;
    const uint8_t* qd[kMaxBatchQueries];   // prepared distance operand image of each query
    uint16_t* out;
    uint64_t row_begin, row_end;
    uint32_t pair_begin, pair_end;         // 256-row tile range covering [row_begin,row_end)
    uint32_t num_queries;
    int* error;
    // Optional wave pacing (two device u32 {arrivals, gave-up}, zero before the launch; nullptr = off): the clusters that share a row tile stay
    // in step so the tile is fetched from DRAM once and served to the other query groups from L2 (see iris_batch.cu).
    uint32_t* wave_sync;
};
cudaError_t launch_batch_distances(const BatchParams& p, bool signed_queries, int num_sms, cudaStream_t stream);

// Batched denominators: same shape for prepared mask operand images.
struct BatchMaskParams {
    const uint8_t* masks;
    const uint8_t* qm[kMaxBatchQueries];
    uint16_t* out;
    uint64_t row_begin, row_end;
    uint32_t pair_begin, pair_end;
    uint32_t num_queries;
    int* error;
};
cudaError_t launch_batch_denominators(const BatchMaskParams& p, int num_sms, cudaStream_t stream);

// Coordinator reduction (iris_reduce.cu): wrapping sum of party shares, decode_distance, min / argmin.
constexpr int kMaxParties = 8;
struct CombineParams {
    const uint16_t* shares[kMaxParties];   // each [n][31] u16 (device)
    uint32_t parties;
    const uint16_t* denominators;          // [n][31] u16 (device)
    uint64_t n;
    uint64_t index_base;                   // added to the row number in the reported argmin
    double* distances_out;                 // optional [n] f64 (device); [Q][n] for a batch
    size_t query_stride;                   // batch: elements between the [n][31] arrays of consecutive queries
};
struct ResultPair {                        // what a search returns per query: 16 bytes
    double min_distance;
    unsigned long long min_index;          // ~0 = nothing below +inf
};
size_t combine_scratch_bytes(uint64_t n);  // per query
cudaError_t launch_combine_min(const CombineParams& p, void* scratch, void* result, cudaStream_t stream, uint32_t num_queries = 1);
// result pair = best of n per-block (min, index) entries (lowest index on ties); the second half of launch_combine_min
cudaError_t launch_final_min(const double* block_min, const unsigned long long* block_idx, uint32_t n, ResultPair* result,
                             cudaStream_t stream);
// out[q] = best over s < n_sets of in[s * stride + q] (lowest index on ties; ~0 indices ignored)
cudaError_t launch_merge_pairs(const ResultPair* in, uint32_t n_sets, uint32_t stride, uint32_t n_queries, ResultPair* out,
                               cudaStream_t stream);

// Number of kernels launched by this library since load (bench.py's gpu_launches).
uint64_t launch_count();

}  // namespace iris
