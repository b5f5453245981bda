/*
 * iris_b200.h -- C ABI of the B200-native matching backend for recmo/mpc-iris-code.
 *
 * This is the boundary a `src/arch/cuda.rs` backend would bind (see INTEGRATION.md for the
 * Rust `extern "C"` block).  It replaces, for the matching hot path only,
 *
 *   arch::dot_u16 / arch::dot_bool              src/arch/mod.rs:5, src/arch/generic.rs:4-16
 *   DistanceEngine::{new, batch_process}        src/lib.rs:28-52
 *   MasksEngine::{new, batch_process}           src/lib.rs:55-79
 *   distances / denominators                    src/lib.rs:82-94
 *
 * keeping the reference's byte layouts at the boundary:
 *
 *   EncodedBits = [u16; 12800]   25 600 B   src/encoded_bits.rs:13-15   (k = row*200 + col)
 *   Bits        = [u64; 200]      1 600 B   src/bits.rs:13-15           (bit k = byte k/8, bit k%8)
 *   Template    = {pattern, mask} 3 200 B   src/template.rs:11-29
 *   result row  = [u16; 31]          62 B   src/lib.rs:42, slot j <-> rotation j-15, rows packed
 *
 * All functions return IRIS_OK (0) or a negative iris_status; iris_last_error() gives the
 * message of the calling thread's last failure.  Nothing aborts or unwinds across the ABI:
 * the reference's `assert_eq!(out.len(), db.len())` panic (src/lib.rs:43,70) becomes
 * IRIS_ERR_INVALID.  There is no CPU fallback: without a CUDA device every compute entry
 * point fails with IRIS_ERR_CUDA.
 *
 * Threading: calls on one handle (database or engine) must be serialised by the caller
 * (the reference calls batch_process from one spawn_blocking thread at a time,
 * src/main.rs:425-431, 510-516); distinct handles may be used from distinct threads.
 * Ownership: the library owns device memory behind the opaque handles; the caller owns
 * every buffer it passes in.  Host inputs (queries, templates, rows) are consumed before the call
 * returns.  Calls with DEVICE outputs are asynchronous on the shard's stream: the output array must
 * stay valid until that stream has reached the call (iris_db_synchronize, or the caller's own
 * stream order after iris_db_set_stream); engines may be freed right away (their operand images
 * are released in stream order).  Calls with host outputs return when the outputs are complete.
 */
#ifndef IRIS_B200_H
#define IRIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRIS_COLS 200      /* src/lib.rs:10 */
#define IRIS_ROWS 64       /* src/lib.rs:11 */
#define IRIS_BITS 12800    /* src/lib.rs:12 */
#define IRIS_LIMBS 200     /* src/bits.rs:10 */
#define IRIS_ROTATIONS 31  /* src/lib.rs:34-35: -15..=15 */

typedef enum iris_status {
    IRIS_OK = 0,
    IRIS_ERR_INVALID = -1, /* bad argument / length mismatch (reference: assert_eq! panic) */
    IRIS_ERR_CUDA = -2,    /* CUDA runtime or kernel failure, or no device */
    IRIS_ERR_NOMEM = -3,   /* host or device allocation failed */
    IRIS_ERR_STATE = -4    /* handle does not hold what the call needs (e.g. no masks loaded) */
} iris_status;

typedef struct iris_db iris_db;                           /* HBM-resident database shard */
typedef struct iris_distance_engine iris_distance_engine; /* DistanceEngine  (src/lib.rs:28-31) */
typedef struct iris_masks_engine iris_masks_engine;       /* MasksEngine     (src/lib.rs:55-58) */

const char *iris_last_error(void);
int iris_device_count(int *count);
/* Kernels launched by the library since it was loaded (benchmark bookkeeping). */
uint64_t iris_launch_count(void);

/* ---- per-pair arch entry points: src/arch/generic.rs:11-16 and :4-9 (host pointers).  API
 * parity only -- one launch per pair cannot amortise; use the engines for throughput. ---- */
int iris_dot_u16(int device, const uint16_t a[IRIS_BITS], const uint16_t b[IRIS_BITS], uint16_t *out);
int iris_dot_bool(int device, const uint64_t a[IRIS_LIMBS], const uint64_t b[IRIS_LIMBS], uint16_t *out);
/* ---- arch-level grids: the reference's criterion benchmark (src/arch/mod.rs:22-72) calls dot_u16 / dot_bool on
 * every pair of n_a INDEPENDENT vectors `a` and n_b vectors `b` (31 x 100 000 for dot_u16).  Here the whole grid is
 * one call: out[i][j] = dot(a[j], b[i]), out = [n_b][n_a] u16.  Groups of 31 vectors of `a` take the place of the 31
 * rotations of a query in the same tensor-core kernels the engines use.  a, b, out: host or device memory. ---- */
int iris_dot_u16_batch(int device, const uint16_t *a /* [n_a][12800] */, uint32_t n_a, const uint16_t *b /* [n_b][12800] */,
                       uint64_t n_b, uint16_t *out);
int iris_dot_bool_batch(int device, const uint64_t *a /* [n_a][200] */, uint32_t n_a, const uint64_t *b /* [n_b][200] */,
                        uint64_t n_b, uint16_t *out);

/* ---- database shard: what the reference mmaps as &[EncodedBits] (src/main.rs:389-391) and
 * &[Bits] (src/main.rs:458-461), loaded ONCE into HBM and re-tiled for the tensor pipeline. ---- */
#define IRIS_DB_SHARES 1u
#define IRIS_DB_MASKS 2u
int iris_db_create(int device, uint64_t capacity_rows, uint32_t flags, iris_db **out);
int iris_db_destroy(iris_db *db);
int iris_db_clear(iris_db *db); /* forget all rows, keep the allocation */
int iris_db_len(const iris_db *db, uint64_t *n_shares, uint64_t *n_masks);
/* Append rows given in the reference's flat-file layouts.  `rows` is host memory; device memory is accepted too, but then
 * the caller must have synchronised the stream that produced it (the library reads it on the shard's stream). */
int iris_db_append_shares(iris_db *db, const uint16_t *rows /* [n][12800] */, uint64_t n);
int iris_db_append_masks(iris_db *db, const uint64_t *rows /* [n][200] */, uint64_t n);
/* Append rows [first_row, first_row+n_rows) (n_rows = 0: to the end) of a file in the reference's on-disk
 * formats: `mpc.share-i` = raw EncodedBits rows, `mpc.masks` = raw Bits rows (written by `prepare`,
 * src/main.rs:337-371; mmapped at src/main.rs:386-400, 458-461).  A size that is not a whole number of rows
 * is an error (the reference's try_cast_slice failure, src/main.rs:391-392). */
int iris_db_load_shares_file(iris_db *db, const char *path, uint64_t first_row, uint64_t n_rows);
int iris_db_load_masks_file(iris_db *db, const char *path, uint64_t first_row, uint64_t n_rows);
/* Append n synthetic rows (uniform u16 shares and/or uniform mask bits, per `flags` of the
 * shard) produced on the device by a counter-based generator keyed by (seed, row id);
 * row ids are first_row_id, first_row_id+1, ...  (oracle/iris_oracle.c restates the generator). */
int iris_db_generate(iris_db *db, uint64_t seed, uint64_t first_row_id, uint64_t n);
/* The same synthetic database, but with shares that MEAN something: row id R is the synthetic Template
 * (pattern_R, mask_R), mask_R being the mask row iris_db_generate produces; its encoding (src/lib.rs:16-26) is split into
 * n_parties additive shares as EncodedBits::share does (src/encoded_bits.rs:23-38: n-1 uniform vectors, the last one
 * the encoding minus their sum), and this shard receives party `party`'s rows.  n_parties = 1 stores the plaintext
 * encodings (the n = 1 sharing).  Masks are generated too when the shard holds masks. */
int iris_db_generate_shares(iris_db *db, uint64_t seed, uint32_t party, uint32_t n_parties, uint64_t first_row_id,
                            uint64_t n);
/* Overwrite rows [row, row+n) that are already loaded (an enrolment update; reference layouts, host pointers). */
int iris_db_write_shares(iris_db *db, uint64_t row, const uint16_t *rows /* [n][12800] */, uint64_t n);
int iris_db_write_masks(iris_db *db, uint64_t row, const uint64_t *rows /* [n][200] */, uint64_t n);
/* Read rows back in the reference layouts (inverse of the loader; tests and debugging). */
int iris_db_read_shares(iris_db *db, uint64_t row_begin, uint64_t n, uint16_t *out /* [n][12800] */);
int iris_db_read_masks(iris_db *db, uint64_t row_begin, uint64_t n, uint64_t *out /* [n][200] */);
/* Run all work of this shard on the given cudaStream_t (NULL = the library's own stream). */
int iris_db_set_stream(iris_db *db, void *cuda_stream);
int iris_db_get_stream(iris_db *db, void **cuda_stream);
int iris_db_device(const iris_db *db, int *device);
int iris_db_synchronize(iris_db *db);
/* Watchdog state of a shard whose stream the caller has synchronised itself. */
int iris_db_check(iris_db *db);

/* ---- page-locked host buffers for result slices.  The reference allocates a fresh Vec per chunk
 * (src/main.rs:429, 514); any host memory works here too, but device->host copies into page-locked
 * memory run at the full PCIe rate and overlap the scan. ---- */
int iris_host_alloc(uint64_t bytes, void **out);
int iris_host_free(void *ptr);
/* Device buffers for results that stay in HBM between calls (e.g. the coordinator's denominators, which
 * iris_combine_min consumes), for callers that do not link the CUDA runtime themselves. */
int iris_device_alloc(int device, uint64_t bytes, void **out);
int iris_device_free(int device, void *ptr);

/* ---- DistanceEngine (src/lib.rs:28-52) ---- */
/* new: prepares the 31 rotations (-15..=15) of `query` as the tensor-core operand image. */
int iris_distance_engine_new(int device, const uint16_t query[IRIS_BITS], iris_distance_engine **out);
/* DistanceEngine::new(&encode(&template)) as the participant does per request (src/main.rs:427): encode
 * (src/lib.rs:16-26) and the rotation preparation both run on the device from the wire Template. */
int iris_distance_engine_new_from_template(int device, const uint64_t pattern[IRIS_LIMBS],
                                           const uint64_t mask[IRIS_LIMBS], iris_distance_engine **out);
/* The same for num_queries Templates at once (templates = [num_queries] x {pattern[200], mask[200]} u64, the wire
 * layout of src/template.rs:26-29): fills distance_engines[i] = DistanceEngine::new(&encode(&t_i)) and, when
 * masks_engines is not NULL, masks_engines[i] = MasksEngine::new(&t_i.mask), with one copy and one synchronisation. */
int iris_engines_new_from_templates(int device, const uint64_t *templates, uint32_t num_queries,
                                    iris_distance_engine **distance_engines, iris_masks_engine **masks_engines);
/* encode(&Template) -> EncodedBits (src/lib.rs:16-26), computed on the device; out host or device. */
int iris_encode(int device, const uint64_t pattern[IRIS_LIMBS], const uint64_t mask[IRIS_LIMBS],
                uint16_t out[IRIS_BITS]);
int iris_distance_engine_free(iris_distance_engine *e);
/* batch_process(&self, out, db) with the reference's exact shape: `db` is a HOST slice of
 * db_len EncodedBits, `out` a HOST slice of out_len [u16;31]; out_len != db_len is an error.
 * Rows are streamed to the GPU, so this call is PCIe-bound; it exists for drop-in parity. */
int iris_distance_engine_batch_process(iris_distance_engine *e, uint16_t *out, uint64_t out_len,
                                       const uint16_t *db, uint64_t db_len);
/* batch_process against rows [row_begin,row_end) of an HBM-resident shard -- the production
 * path (the reference calls batch_process on 20 000-row chunks of its mmap, src/main.rs:428).
 * `out` ([row_end-row_begin][31]) may be host memory (pageable or pinned) or device memory. */
int iris_distance_engine_batch_process_resident(iris_distance_engine *e, uint16_t *out, uint64_t out_len,
                                                iris_db *db, uint64_t row_begin, uint64_t row_end);

/* ---- MasksEngine (src/lib.rs:55-79) ---- */
int iris_masks_engine_new(int device, const uint64_t query_mask[IRIS_LIMBS], iris_masks_engine **out);
int iris_masks_engine_free(iris_masks_engine *e);
int iris_masks_engine_batch_process(iris_masks_engine *e, uint16_t *out, uint64_t out_len, const uint64_t *db,
                                    uint64_t db_len);
int iris_masks_engine_batch_process_resident(iris_masks_engine *e, uint16_t *out, uint64_t out_len, iris_db *db,
                                             uint64_t row_begin, uint64_t row_end);

/* ---- fused scan: both engines over the same rows in ONE pass over HBM (shares + masks read
 * once; BASELINE config 2).  Either engine/out pair may be NULL. ---- */
int iris_match_resident(iris_distance_engine *de, iris_masks_engine *me, iris_db *db, uint64_t row_begin,
                        uint64_t row_end, uint16_t *distances_out, uint16_t *denominators_out);

/* The same fused scan with HOST result arrays, reporting progress: `progress(user, b, e)` is called on the calling thread
 * each time the rows [b, e) (database rows, in ascending order, together covering [row_begin,row_end)) are complete in
 * host memory, while the scan of the following rows is already running -- so a caller can stream the reply out (the
 * participant's socket loop, src/main.rs:437-443) while the GPU is still scanning. */
typedef void (*iris_progress_fn)(void *user, uint64_t row_begin, uint64_t row_end);
int iris_match_resident_streamed(iris_distance_engine *de, iris_masks_engine *me, iris_db *db, uint64_t row_begin,
                                 uint64_t row_end, uint16_t *distances_out, uint16_t *denominators_out,
                                 iris_progress_fn progress, void *user);
/* Consecutive scans of one shard with device outputs that do not overlap in memory may overlap in time: the next scan
 * starts on the SMs the previous scan's last partial wave leaves idle (the reference calls batch_process on 20 000-row
 * chunks, src/main.rs:427-430: 157 tiles on 148 SMs).  On by default, on the library's stream and on a caller's stream
 * alike; a scan never starts early behind a kernel that is not a scan of this shard.  allow = 0 turns it off. */
int iris_db_set_overlap(iris_db *db, int allow);

/* ---- batched queries (BASELINE config 4): num_queries DistanceEngines against the same rows as ONE dense
 * int8 GEMM on the tensor cores; num_queries MasksEngines four at a time over the 4-bit operand expanded into
 * tensor memory.  Equivalent to calling batch_process_resident once per engine;
 * out = [num_queries][row_end-row_begin][31] u16, host or device memory. ---- */
int iris_distances_batch_resident(iris_distance_engine *const *engines, uint32_t num_queries, iris_db *db,
                                  uint64_t row_begin, uint64_t row_end, uint16_t *out);
int iris_denominators_batch_resident(iris_masks_engine *const *engines, uint32_t num_queries, iris_db *db,
                                     uint64_t row_begin, uint64_t row_end, uint16_t *out);

/* The arch-level grids (see iris_dot_u16_batch) against rows [row_begin,row_end) of a resident shard taking the place of
 * `b`: out = [row_end-row_begin][n_a] u16, host or device; asynchronous on the shard's stream for a device `out`. */
int iris_dot_u16_batch_resident(const uint16_t *a, uint32_t n_a, iris_db *db, uint64_t row_begin, uint64_t row_end,
                                uint16_t *out);
int iris_dot_bool_batch_resident(const uint64_t *a, uint32_t n_a, iris_db *db, uint64_t row_begin, uint64_t row_end,
                                 uint16_t *out);

/* ---- coordinator reduction on the device (src/main.rs:597-621 with decode_distance, src/lib.rs:97-107):
 * numerator = wrapping sum of the parties' distance shares; distance = min over rotations of
 * ((den - num) as u16 / 2) / den in f64 (NaN ignored); running min with `<` (first minimum wins;
 * min_index = UINT64_MAX when nothing is below +inf).  Inputs are [n][31] u16 arrays, host or device;
 * distances_out (optional, [n] f64) receives the per-row decoded distances.  Bit-identical to the CPU. ---- */
int iris_combine_min(int device, const uint16_t *const *distance_shares, uint32_t parties,
                     const uint16_t *denominators, uint64_t n, uint64_t index_base, double *distances_out,
                     double *min_distance, uint64_t *min_index);
/* The same reduction for a whole batch: distances / denominators are [Q][n][31] DEVICE arrays as written by the
 * batched kernels (synchronise the producing shard first); min_distance / min_index receive Q entries. */
int iris_combine_min_batch(int device, const uint16_t *distances, const uint16_t *denominators, uint32_t num_queries,
                           uint64_t n, uint64_t index_base, double *min_distance, uint64_t *min_index);
/* Fused scan + reduction for a shard that holds the whole (1-share) encodings: both engines over rows
 * [row_begin,row_end), decode and min/argmin on the device; only 16 bytes return to the host.  This is
 * the per-shard step of the multi-GPU path (each rank reduces its rows, the pairs are all-gathered). */
int iris_match_min_resident(iris_distance_engine *de, iris_masks_engine *me, iris_db *db, uint64_t row_begin,
                            uint64_t row_end, uint64_t index_base, double *min_distance, uint64_t *min_index);

/* Asynchronous form: the {f64 min_distance, u64 min_index} pair (16 bytes) is written to `result` -- device memory of
 * this GPU, of a peer GPU (NVLink store), or mapped host memory -- in stream order on the shard's stream. */
int iris_match_min_resident_async(iris_distance_engine *de, iris_masks_engine *me, iris_db *db, uint64_t row_begin,
                                  uint64_t row_end, uint64_t index_base, void *result);
/* Batched search on one shard (BASELINE configs[3]/[4] per GPU): num_queries (<= 64) engine pairs against rows
 * [row_begin,row_end): batched tensor-core distances + denominators slice by slice into scratch owned by the shard,
 * decode + min/argmin on the device, running min over the slices; results = [num_queries] x {f64, u64} written in
 * stream order like above. */
int iris_search_batch_resident_async(iris_distance_engine *const *des, iris_masks_engine *const *mes,
                                     uint32_t num_queries, iris_db *db, uint64_t row_begin, uint64_t row_end,
                                     uint64_t index_base, void *results);

/* ---- cluster: ONE database row-sharded over several GPUs of the box (BASELINE configs[4]; the reference's mmap of
 * the whole share file, src/main.rs:386-400, becomes contiguous row blocks in the HBM of each GPU).  The library runs
 * one host thread and one stream per GPU.  Rows are independent, so the scan needs no collective; what is exchanged
 * are the small per-query vectors:
 *   - search: each shard's (min, argmin) pairs are stored by its reduction kernel straight into the root GPU's memory
 *     over NVLink (peer stores; mapped host memory when the GPUs cannot reach each other) and merged there;
 *   - match: every shard's scan kernel stores its [rows][31] result block at its row offset of ONE caller array, which
 *     may live on any GPU of the cluster (peer stores) or in (pinned) host memory (one PCIe link per GPU in parallel);
 *   - several processes (one per GPU, e.g. under torchrun) join one cluster with iris_cluster_join: the merged pairs
 *     of each process are all-gathered over NCCL (libnccl.so.2, loaded at run time) and merged again.  The copy used
 *     is the one the process has already loaded, else the file the environment variable IRIS_NCCL_LIB names, else the
 *     system's; a process can hold only one libnccl.so.2, so a host that will load a different copy later (Python
 *     importing torch after its first join) names that copy in IRIS_NCCL_LIB before the first NCCL call.
 * `devices` may name a GPU more than once (several shards on one GPU; used by the tests on one-GPU boxes). ---- */
typedef struct iris_cluster iris_cluster;
int iris_cluster_create(const int *devices, uint32_t n_devices, uint64_t capacity_rows, uint32_t flags,
                        iris_cluster **out);
int iris_cluster_destroy(iris_cluster *c);
/* Contiguous block of shard `shard` when n_total rows are spread over n_shards (sizes differ by at most one row). */
int iris_cluster_partition(uint64_t n_total, uint32_t n_shards, uint32_t shard, uint64_t *row_begin, uint64_t *row_end);
/* Shard i of the cluster: its handle (usable with every iris_db_* / engine call), device and block of cluster rows. */
int iris_cluster_shard(iris_cluster *c, uint32_t shard, iris_db **db, int *device, uint64_t *row_begin, uint64_t *row_end);
int iris_cluster_len(const iris_cluster *c, uint32_t *n_shards, uint64_t *n_shares, uint64_t *n_masks);
/* Populate (each replaces what the cluster held): n rows spread evenly, every GPU loading its block in parallel.
 * n_parties = 0: uniform shares (iris_db_generate); otherwise iris_db_generate_shares.  Files are the reference's
 * formats (see iris_db_load_*_file); either path may be NULL. */
int iris_cluster_generate(iris_cluster *c, uint64_t seed, uint32_t party, uint32_t n_parties, uint64_t first_row_id,
                          uint64_t n);
int iris_cluster_load_files(iris_cluster *c, const char *shares_path, const char *masks_path);
int iris_cluster_load_rows(iris_cluster *c, const uint16_t *shares /* [n][12800] or NULL */,
                           const uint64_t *masks /* [n][200] or NULL */, uint64_t n);
/* Row ids reported by searches are index_base + cluster row (a process of a multi-process cluster sets its offset). */
int iris_cluster_set_index_base(iris_cluster *c, uint64_t index_base);
/* One query against every row: DistanceEngine::new(query).batch_process and/or MasksEngine::new(mask).batch_process
 * over the whole cluster (src/main.rs:425-431, 510-516).  query / query_mask: host memory (either may be NULL with its
 * output); outputs = [n][31] u16 in host memory or on any GPU of the cluster.  Returns when the outputs are complete. */
int iris_cluster_match(iris_cluster *c, const uint16_t *query, const uint64_t *query_mask, uint16_t *distances_out,
                       uint16_t *denominators_out);
/* The participant's request (src/main.rs:419-431): encode(&template) on each GPU, distances of the whole cluster. */
int iris_cluster_match_template(iris_cluster *c, const uint64_t *pattern, const uint64_t *mask, uint16_t *distances_out,
                                uint16_t *denominators_out);
/* The same with HOST outputs and progress reports (see iris_match_resident_streamed): `progress` is called with ranges of
 * CLUSTER rows, from the library's per-GPU threads (several at a time, each GPU's ranges in ascending order). */
int iris_cluster_match_template_streamed(iris_cluster *c, const uint64_t *pattern, const uint64_t *mask,
                                         uint16_t *distances_out, uint16_t *denominators_out, iris_progress_fn progress,
                                         void *user);
/* Search: num_queries wire Templates ([num_queries][400] u64 = {pattern[200], mask[200]}) against a cluster that
 * holds whole encodings (n_parties = 1): per query the minimum decoded distance over all rows and rotations and the
 * row attaining it (lowest row on ties; UINT64_MAX when nothing is below +inf) -- the coordinator's loop
 * (src/main.rs:597-621) over every shard.  One query takes the fused HBM-bound scan, several the batched tensor-core
 * path.  Only num_queries x 16 bytes leave the GPUs. */
int iris_cluster_search(iris_cluster *c, const uint64_t *templates, uint32_t num_queries, double *min_distance,
                        uint64_t *min_index);
/* Multi-process clusters: rank 0 draws an id (128 bytes), every process receives it by any channel and joins.
 * Afterwards iris_cluster_search is a collective call (same num_queries everywhere) and returns the global result in
 * every process. */
#define IRIS_UNIQUE_ID_BYTES 128
int iris_comm_unique_id(void *id_out /* [128] */);
int iris_cluster_join(iris_cluster *c, const void *unique_id, int rank, int world_size);
/* Multi-process clusters: iris_cluster_match with the full result vectors of ALL processes gathered on every process
 * (collective).  The outputs are DEVICE arrays of [rows of all processes][31] u16; process p's rows occupy
 * [index_base_p, index_base_p + rows_p).  Every process scans into its slot, then the blocks travel over NVLink with one
 * grouped ncclBroadcast per process.  (Inside one process no second step is needed: see iris_cluster_match.) */
int iris_cluster_match_allgather(iris_cluster *c, const uint16_t *query, const uint64_t *query_mask,
                                 uint16_t *distances_out, uint16_t *denominators_out);

/* ---- single-pair wrappers: src/lib.rs:82-87 and :89-94 ---- */
int iris_distances(int device, const uint16_t query[IRIS_BITS], const uint16_t entry[IRIS_BITS],
                   uint16_t out[IRIS_ROTATIONS]);
int iris_denominators(int device, const uint64_t query[IRIS_LIMBS], const uint64_t entry[IRIS_LIMBS],
                      uint16_t out[IRIS_ROTATIONS]);

/* ---- verification helpers (CUDA-core kernels over the same HBM image; NOT the product path):
 * used by the GPU tests to cross-check the tensor-core scan at full database size. ---- */
int iris_check_distances_simt(iris_db *db, const uint16_t query[IRIS_BITS], uint64_t row_begin, uint64_t row_end,
                              uint16_t *out);
int iris_check_denominators_simt(iris_db *db, const uint64_t query_mask[IRIS_LIMBS], uint64_t row_begin,
                                 uint64_t row_end, uint16_t *out);
/* Debug: run the fused scan over [row_begin,row_end) (row_begin multiple of 128) and dump the raw s32
 * accumulators, [tiles*128][128] = {S00[32], S10[32], S01[32], 128*popcount[32]} per row, to host. */
int iris_debug_raw_accumulators(iris_distance_engine *de, iris_masks_engine *me, iris_db *db, uint64_t row_begin,
                                uint64_t row_end, int32_t *raw_out);

#ifdef __cplusplus
}
#endif
#endif /* IRIS_B200_H */
