// Host runtime behind include/iris_b200.h: handles, loader, stream pipeline, error channel.
// Mirrors the reference's engine API (src/lib.rs:28-94); see the header for the mapping.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/iris_b200.h"
#include "iris_kernels.cuh"

using namespace iris;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CK(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            cudaGetLastError(); /* a non-sticky error must not resurface in a later, unrelated call */ \
            return fail(e_ == cudaErrorMemoryAllocation ? IRIS_ERR_NOMEM : IRIS_ERR_CUDA, "%s failed: %s", #expr, \
                        cudaGetErrorString(e_));                                                     \
        }                                                                                            \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static bool is_device_pointer(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

extern "C" const char* iris_last_error(void) { return g_last_error.c_str(); }

extern "C" int iris_device_count(int* count) {
    if (!count) return fail(IRIS_ERR_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(IRIS_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *count = n;
    return IRIS_OK;
}

extern "C" uint64_t iris_launch_count(void) { return launch_count(); }

// ------------------------------------------------------------------------------------ handles
static constexpr uint64_t kStageRows = 2048;                      // loader staging: 52 MB of shares
static constexpr int kBatchWavePacingDefault = 1;                 // see batch_wave_pacing()
static constexpr uint64_t kResultChunkTilesPerSm = 8;             // host-output pipeline granularity

struct iris_db {
    int device = 0;
    int num_sms = 148;
    uint64_t capacity = 0;   // rows, multiple of 256
    uint32_t flags = 0;
    uint64_t n_shares = 0, n_masks = 0;
    uint8_t* d_shares = nullptr;
    uint8_t* d_masks = nullptr;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_scan[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    void* d_stage = nullptr;         // loader staging (reference-layout rows)
    uint16_t* d_res[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [buffer][dist|den]
    uint64_t res_rows = 0;
    uint8_t* d_red = nullptr;        // match_min: results + reduction scratch
    uint64_t red_rows = 0;
    uint8_t* d_batch = nullptr;      // batched search: [Q][slice][31] distances + denominators + reduction scratch
    size_t batch_bytes = 0;
    uint32_t* d_wave_sync = nullptr; // pacing words of the batched GEMM
    ResultPair* d_pairs = nullptr;   // per-slice result pairs of a search
    size_t pairs_cap = 0;
    // Watchdog flag: mapped page-locked HOST memory, so the code a trapping kernel leaves behind can still be read
    // after the trap has poisoned the context.  h_error is the host view, d_error the device alias.
    int* h_error = nullptr;
    int* d_error = nullptr;
    // device-output scans still possibly in flight on own_stream: [begin, end) byte ranges of their outputs.  A new scan
    // whose outputs are disjoint from all of them may overlap their tails (programmatic dependent launch).
    static constexpr int kChain = 6;
    uintptr_t chain_lo[kChain][2] = {}, chain_hi[kChain][2] = {};
    int chain_len = 0;
    bool chain_all_wide = true;      // every scan of the current chain has at least num_sms / 2 CTAs
    bool overlap = true;             // consecutive device-output scans of this shard may overlap (iris_db_set_overlap)
};

// Shards that are alive (engines remember the shard they last scanned so that their buffers are released in stream
// order; a shard that has been destroyed was synchronised on the way out and needs no ordering any more).
static std::mutex g_db_mu;
static std::vector<iris_db*> g_live_dbs;
static bool db_is_live(iris_db* db) {
    std::lock_guard<std::mutex> g(g_db_mu);
    return std::find(g_live_dbs.begin(), g_live_dbs.end(), db) != g_live_dbs.end();
}

// Everything one prepared query owns, pooled per (kind, device) so that engine construction per request performs no
// cudaMalloc.  Release is STREAM-ORDERED: `used` is recorded behind the last scan that reads the operand images when
// the engine is freed, and the next owner's preparation waits for it on the device -- iris_*_engine_free never
// blocks and a scan that is still running keeps reading intact data.  `h` is page-locked staging owned by the slot:
// the caller's query is copied into it before the call returns, so no caller pointer is retained.
struct EngineSlot {
    int device = -1;
    uint8_t* d = nullptr;
    uint8_t* h = nullptr;
    cudaEvent_t ready = nullptr, used = nullptr;
};
enum SlotKind { kSlotDistance = 0, kSlotMasks = 1 };
// distance slot, device: [qd 819 200][query 25 600][template 3 200]; host: [query 25 600][s8 flag]
constexpr size_t kDSlotQuery = kQdBytes, kDSlotTemplate = kQdBytes + IRIS_BITS * 2;
constexpr size_t kDSlotDevBytes = kDSlotTemplate + 2 * IRIS_MASK_BYTES, kDSlotHostBytes = IRIS_BITS * 2 + 64;
// masks slot, device: [qm 409 600 | qm4 204 800][mask 1 600]; host: [mask 1 600]
constexpr size_t kMSlotMask = kQmBytes + kQm4Bytes;
constexpr size_t kMSlotDevBytes = kMSlotMask + 1664, kMSlotHostBytes = 1664;

struct SlotPool {
    std::mutex mu;
    std::vector<EngineSlot> free_slots[2];
    bool take(SlotKind kind, int dev, EngineSlot* out) {
        std::lock_guard<std::mutex> g(mu);
        auto& v = free_slots[kind];
        for (size_t i = v.size(); i-- > 0;)
            if (v[i].device == dev) {
                *out = v[i];
                v.erase(v.begin() + i);
                return true;
            }
        return false;
    }
    bool give(SlotKind kind, const EngineSlot& s) {
        std::lock_guard<std::mutex> g(mu);
        if (free_slots[kind].size() >= 512) return false;
        free_slots[kind].push_back(s);
        return true;
    }
};
static SlotPool g_slots;

static void slot_destroy(EngineSlot& s) {
    if (s.used) cudaEventSynchronize(s.used);
    if (s.ready) cudaEventSynchronize(s.ready);
    if (s.d) cudaFree(s.d);
    if (s.h) cudaFreeHost(s.h);
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.used) cudaEventDestroy(s.used);
    cudaGetLastError();
    s = EngineSlot();
}

// A slot for a new engine; `stream` (the preparation stream) is made to wait for the previous owner's last scan.
static int slot_acquire(SlotKind kind, int device, cudaStream_t stream, EngineSlot* out) {
    if (g_slots.take(kind, device, out)) {
        CK(cudaStreamWaitEvent(stream, out->used, 0));          // a never-recorded event is complete
        return IRIS_OK;
    }
    EngineSlot s;
    s.device = device;
    auto body = [&]() -> int {
        CK(cudaMalloc(&s.d, kind == kSlotDistance ? kDSlotDevBytes : kMSlotDevBytes));
        CK(cudaHostAlloc(&s.h, kind == kSlotDistance ? kDSlotHostBytes : kMSlotHostBytes, cudaHostAllocPortable | cudaHostAllocMapped));
        CK(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.used, cudaEventDisableTiming));
        return IRIS_OK;
    };
    int rc = body();
    if (rc) {
        std::string keep = g_last_error;
        slot_destroy(s);
        g_last_error = keep;
        return rc;
    }
    *out = s;
    return IRIS_OK;
}

struct EngineUse {                   // which shard's stream last read the operand images of an engine
    iris_db* last_db = nullptr;
    cudaStream_t ready_seen = nullptr;   // stream that already waits behind `ready`
    bool ready_seen_valid = false;
};

struct iris_distance_engine {
    int device = 0;
    EngineSlot slot;
    uint16_t* d_query = nullptr;
    uint8_t* d_qd = nullptr;
    bool fits_s8 = false;            // every element is a sign-extended byte (true for encode() output)
    bool classified = true;          // false: the s8 flag is still being computed on the device (device-pointer query)
    EngineUse use;
    iris_db* scratch = nullptr;      // for the host-slice batch_process
};

struct iris_masks_engine {
    int device = 0;
    EngineSlot slot;
    uint8_t* d_qmask = nullptr;
    uint8_t* d_qm = nullptr;
    EngineUse use;
    iris_db* scratch = nullptr;
};

// Orders a scan on `db` behind the engine's preparation and remembers the shard for the stream-ordered release.
static int engine_begin_use(EngineSlot& slot, EngineUse& use, iris_db* db) {
    if (use.last_db && use.last_db != db && db_is_live(use.last_db)) {
        // rare: one engine scanning several shards.  Make the new shard's stream follow the old one's, so that the
        // single `used` record at release time covers both.
        CK(cudaEventRecord(slot.used, use.last_db->stream));
        CK(cudaStreamWaitEvent(db->stream, slot.used, 0));
    }
    use.last_db = db;
    if (!use.ready_seen_valid || use.ready_seen != db->stream) {
        CK(cudaStreamWaitEvent(db->stream, slot.ready, 0));
        use.ready_seen = db->stream;
        use.ready_seen_valid = true;
    }
    return IRIS_OK;
}

static void engine_release(SlotKind kind, EngineSlot& slot, EngineUse& use) {
    if (!slot.d) return;
    if (use.last_db && db_is_live(use.last_db)) cudaEventRecord(slot.used, use.last_db->stream);
    else cudaEventRecord(slot.used, cudaStreamPerThread);      // behind the preparation itself
    cudaGetLastError();
    if (!g_slots.give(kind, slot)) slot_destroy(slot);
    slot = EngineSlot();
}

// Page-locked staging of the calling thread for multi-query uploads (per device; grown on demand, reused once the
// previous upload from it has completed; lives as long as the thread -- never freed, it is one small buffer).
struct ThreadStage {
    void* p = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
};
static int thread_stage(int device, size_t bytes, ThreadStage** out) {
    static thread_local ThreadStage stages[64];
    if (device < 0 || device >= 64) return fail(IRIS_ERR_INVALID, "device %d out of range", device);
    ThreadStage& st = stages[device];
    if (!st.done) CK(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    CK(cudaEventSynchronize(st.done));
    if (st.cap < bytes) {
        if (st.p) cudaFreeHost(st.p);
        st.p = nullptr;
        st.cap = 0;
        CK(cudaHostAlloc(&st.p, bytes, cudaHostAllocPortable));
        st.cap = bytes;
    }
    *out = &st;
    return IRIS_OK;
}

// Temporary device memory that lives for one ABI call: stream-ordered allocations (cudaStreamPerThread) from the
// device's default pool, which is told to keep what it has been given.  After the first call of a kind no
// cudaMalloc / cudaFree -- both of which synchronise the whole device -- is left on the per-query path.
static int temp_alloc(int device, void** out, size_t bytes, cudaStream_t stream = cudaStreamPerThread) {
    static std::atomic<bool> tuned[64];
    if (device >= 0 && device < 64 && !tuned[device].load(std::memory_order_acquire)) {
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        tuned[device].store(true, std::memory_order_release);
    }
    CK(cudaMallocAsync(out, bytes ? bytes : 1, stream));
    return IRIS_OK;
}
static void temp_free(void* p, cudaStream_t stream = cudaStreamPerThread) {
    if (p) cudaFreeAsync(p, stream);
}

// ------------------------------------------------------------------------------------ database
extern "C" int iris_db_create(int device, uint64_t capacity_rows, uint32_t flags, iris_db** out) {
    if (!out) return fail(IRIS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!(flags & (IRIS_DB_SHARES | IRIS_DB_MASKS)) || (flags & ~(IRIS_DB_SHARES | IRIS_DB_MASKS)))
        return fail(IRIS_ERR_INVALID, "flags must be a combination of IRIS_DB_SHARES and IRIS_DB_MASKS");
    if (capacity_rows == 0) return fail(IRIS_ERR_INVALID, "capacity_rows must be > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(IRIS_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(IRIS_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return fail(IRIS_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    iris_db* db = new (std::nothrow) iris_db();
    if (!db) return fail(IRIS_ERR_NOMEM, "host allocation failed");
    db->device = device;
    db->flags = flags;
    // whole 256-row pair tiles (the batched kernel works on pairs of 128-row tiles), zero filled
    db->capacity = (capacity_rows + 2 * kTileRows - 1) / (2 * kTileRows) * (2 * kTileRows);
    const uint64_t tiles = db->capacity / kTileRows;
    int rc = IRIS_OK;
    auto body = [&]() -> int {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) return fail(IRIS_ERR_CUDA, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
        db->num_sms = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&db->own_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&db->copy_stream, cudaStreamNonBlocking));
        db->stream = db->own_stream;
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&db->ev_scan[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&db->ev_copy[i], cudaEventDisableTiming));
        }
        CK(cudaHostAlloc(reinterpret_cast<void**>(&db->h_error), 64, cudaHostAllocPortable | cudaHostAllocMapped));
        *db->h_error = 0;
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&db->d_error), db->h_error, 0));
        if (flags & IRIS_DB_SHARES) {
            CK(cudaMalloc(&db->d_shares, tiles * kShareTileBytes));
            CK(cudaMemsetAsync(db->d_shares, 0, tiles * kShareTileBytes, db->stream));
        }
        if (flags & IRIS_DB_MASKS) {
            CK(cudaMalloc(&db->d_masks, tiles * kMaskTileBytes));
            CK(cudaMemsetAsync(db->d_masks, 0, tiles * kMaskTileBytes, db->stream));
        }
        CK(cudaStreamSynchronize(db->stream));
        return IRIS_OK;
    };
    rc = body();
    if (rc != IRIS_OK) {
        std::string keep = g_last_error;
        iris_db_destroy(db);
        g_last_error = keep;
        return rc;
    }
    {
        std::lock_guard<std::mutex> lk(g_db_mu);
        g_live_dbs.push_back(db);
    }
    *out = db;
    return IRIS_OK;
}

extern "C" int iris_db_destroy(iris_db* db) {
    if (!db) return IRIS_OK;
    {
        std::lock_guard<std::mutex> lk(g_db_mu);
        g_live_dbs.erase(std::remove(g_live_dbs.begin(), g_live_dbs.end(), db), g_live_dbs.end());
    }
    DeviceGuard g(db->device);
    if (db->own_stream) cudaStreamSynchronize(db->own_stream);
    if (db->copy_stream) cudaStreamSynchronize(db->copy_stream);
    cudaFree(db->d_shares);
    cudaFree(db->d_masks);
    cudaFree(db->d_stage);
    cudaFree(db->d_red);
    cudaFree(db->d_batch);
    cudaFree(db->d_pairs);
    cudaFree(db->d_wave_sync);
    if (db->h_error) cudaFreeHost(db->h_error);
    for (int b = 0; b < 2; ++b)
        for (int k = 0; k < 2; ++k) cudaFree(db->d_res[b][k]);
    for (int i = 0; i < 2; ++i) {
        if (db->ev_scan[i]) cudaEventDestroy(db->ev_scan[i]);
        if (db->ev_copy[i]) cudaEventDestroy(db->ev_copy[i]);
    }
    if (db->own_stream) cudaStreamDestroy(db->own_stream);
    if (db->copy_stream) cudaStreamDestroy(db->copy_stream);
    cudaGetLastError();
    delete db;
    return IRIS_OK;
}

extern "C" int iris_db_clear(iris_db* db) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    db->n_shares = db->n_masks = 0;   // stale rows beyond n_* are never stored by a scan
    return IRIS_OK;
}

extern "C" int iris_db_len(const iris_db* db, uint64_t* n_shares, uint64_t* n_masks) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    if (n_shares) *n_shares = db->n_shares;
    if (n_masks) *n_masks = db->n_masks;
    return IRIS_OK;
}

extern "C" int iris_db_set_stream(iris_db* db, void* cuda_stream) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    DeviceGuard g(db->device);
    CK(cudaStreamSynchronize(db->stream));
    db->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : db->own_stream;
    db->chain_len = 0;               // everything launched so far has completed
    return IRIS_OK;
}

// The flag is host memory: reading it costs nothing and works after a trap has poisoned the context.
static int check_error_flag(iris_db* db) {
    const int h = *reinterpret_cast<volatile int*>(db->h_error);
    if (h != 0) return fail(IRIS_ERR_CUDA, "scan kernel watchdog fired (code %d)", h);
    return IRIS_OK;
}

// synchronise a stream of this shard; a failure (e.g. the launch failure a watchdog trap leaves behind) reports the
// watchdog code when there is one
static int sync_checked(iris_db* db, cudaStream_t s) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        cudaGetLastError();
        const int h = *reinterpret_cast<volatile int*>(db->h_error);
        if (h != 0) return fail(IRIS_ERR_CUDA, "scan kernel watchdog fired (code %d): %s", h, cudaGetErrorString(e));
        return fail(IRIS_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
    }
    return IRIS_OK;
}

extern "C" int iris_db_synchronize(iris_db* db) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    DeviceGuard g(db->device);
    int rc = sync_checked(db, db->stream);
    if (rc) return rc;
    rc = sync_checked(db, db->copy_stream);
    if (rc) return rc;
    return check_error_flag(db);
}

static int require_device(int device);

extern "C" int iris_host_alloc(uint64_t bytes, void** out) {
    if (!out) return fail(IRIS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(IRIS_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    if (cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(IRIS_ERR_NOMEM, "cannot page-lock %llu bytes of host memory", (unsigned long long)bytes);
    }
    return IRIS_OK;
}

extern "C" int iris_host_free(void* ptr) {
    if (ptr) CK(cudaFreeHost(ptr));
    return IRIS_OK;
}

extern "C" int iris_device_alloc(int device, uint64_t bytes, void** out) {
    if (!out) return fail(IRIS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    if (cudaMalloc(out, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(IRIS_ERR_NOMEM, "cannot allocate %llu bytes on device %d", (unsigned long long)bytes, device);
    }
    return IRIS_OK;
}

extern "C" int iris_device_free(int device, void* ptr) {
    if (!ptr) return IRIS_OK;
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    CK(cudaFree(ptr));
    return IRIS_OK;
}

static int ensure_stage(iris_db* db) {
    if (!db->d_stage) CK(cudaMalloc(&db->d_stage, kStageRows * IRIS_BITS * sizeof(uint16_t)));
    return IRIS_OK;
}

// rows (reference layout; host, or device after the producing stream has been synchronised by the caller) ->
// tiled image at rows [row0, row0 + n)
static int upload_shares(iris_db* db, const uint16_t* rows, uint64_t n, uint64_t row0) {
    DeviceGuard g(db->device);
    if (is_device_pointer(rows)) {
        CK(launch_retile_shares(rows, n, db->d_shares, row0, db->stream));
    } else {
        int rc = ensure_stage(db);
        if (rc) return rc;
        for (uint64_t off = 0; off < n; off += kStageRows) {
            const uint64_t m = std::min(kStageRows, n - off);
            CK(cudaMemcpyAsync(db->d_stage, rows + off * IRIS_BITS, m * IRIS_BITS * sizeof(uint16_t), cudaMemcpyHostToDevice, db->stream));
            CK(launch_retile_shares(static_cast<const uint16_t*>(db->d_stage), m, db->d_shares, row0 + off, db->stream));
        }
    }
    return sync_checked(db, db->stream);
}

static int upload_masks(iris_db* db, const uint64_t* rows, uint64_t n, uint64_t row0) {
    DeviceGuard g(db->device);
    if (is_device_pointer(rows)) {
        CK(launch_retile_masks(reinterpret_cast<const uint8_t*>(rows), n, db->d_masks, row0, db->stream));
    } else {
        int rc = ensure_stage(db);
        if (rc) return rc;
        const uint64_t step = kStageRows * 16;   // same staging buffer, 1600-byte rows
        for (uint64_t off = 0; off < n; off += step) {
            const uint64_t m = std::min(step, n - off);
            CK(cudaMemcpyAsync(db->d_stage, rows + off * IRIS_LIMBS, m * IRIS_MASK_BYTES, cudaMemcpyHostToDevice, db->stream));
            CK(launch_retile_masks(static_cast<const uint8_t*>(db->d_stage), m, db->d_masks, row0 + off, db->stream));
        }
    }
    return sync_checked(db, db->stream);
}

extern "C" int iris_db_append_shares(iris_db* db, const uint16_t* rows, uint64_t n) {
    if (!db || (!rows && n)) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (!(db->flags & IRIS_DB_SHARES)) return fail(IRIS_ERR_STATE, "shard was created without IRIS_DB_SHARES");
    if (db->n_shares + n > db->capacity)
        return fail(IRIS_ERR_INVALID, "append of %llu rows exceeds capacity %llu", (unsigned long long)n, (unsigned long long)db->capacity);
    int rc = upload_shares(db, rows, n, db->n_shares);
    if (rc) return rc;
    db->n_shares += n;
    return IRIS_OK;
}

extern "C" int iris_db_append_masks(iris_db* db, const uint64_t* rows, uint64_t n) {
    if (!db || (!rows && n)) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (!(db->flags & IRIS_DB_MASKS)) return fail(IRIS_ERR_STATE, "shard was created without IRIS_DB_MASKS");
    if (db->n_masks + n > db->capacity)
        return fail(IRIS_ERR_INVALID, "append of %llu rows exceeds capacity %llu", (unsigned long long)n, (unsigned long long)db->capacity);
    int rc = upload_masks(db, rows, n, db->n_masks);
    if (rc) return rc;
    db->n_masks += n;
    return IRIS_OK;
}

// Overwrite rows that are already loaded (an enrolment update; the tests and the bench plant known templates).
extern "C" int iris_db_write_shares(iris_db* db, uint64_t row, const uint16_t* rows, uint64_t n) {
    if (!db || (!rows && n)) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (!(db->flags & IRIS_DB_SHARES)) return fail(IRIS_ERR_STATE, "shard was created without IRIS_DB_SHARES");
    if (row + n > db->n_shares) return fail(IRIS_ERR_INVALID, "rows [%llu,%llu) beyond the %llu loaded shares", (unsigned long long)row, (unsigned long long)(row + n), (unsigned long long)db->n_shares);
    return upload_shares(db, rows, n, row);
}

extern "C" int iris_db_write_masks(iris_db* db, uint64_t row, const uint64_t* rows, uint64_t n) {
    if (!db || (!rows && n)) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (!(db->flags & IRIS_DB_MASKS)) return fail(IRIS_ERR_STATE, "shard was created without IRIS_DB_MASKS");
    if (row + n > db->n_masks) return fail(IRIS_ERR_INVALID, "rows [%llu,%llu) beyond the %llu loaded masks", (unsigned long long)row, (unsigned long long)(row + n), (unsigned long long)db->n_masks);
    return upload_masks(db, rows, n, row);
}

extern "C" int iris_db_generate(iris_db* db, uint64_t seed, uint64_t first_row_id, uint64_t n) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    const bool s = db->flags & IRIS_DB_SHARES, m = db->flags & IRIS_DB_MASKS;
    if (s && m && db->n_shares != db->n_masks) return fail(IRIS_ERR_STATE, "shares and masks lengths differ");
    const uint64_t row0 = s ? db->n_shares : db->n_masks;
    if (row0 + n > db->capacity) return fail(IRIS_ERR_INVALID, "generate exceeds capacity");
    DeviceGuard g(db->device);
    // bounded launches (grid dimension limits)
    const uint64_t step = 1u << 18;
    for (uint64_t off = 0; off < n; off += step) {
        const uint64_t cnt = std::min(step, n - off);
        CK(launch_generate(s ? db->d_shares : nullptr, m ? db->d_masks : nullptr, seed, first_row_id + off, row0 + off, cnt, db->stream));
    }
    CK(cudaStreamSynchronize(db->stream));
    if (s) db->n_shares += n;
    if (m) db->n_masks += n;
    return IRIS_OK;
}

extern "C" int iris_db_generate_shares(iris_db* db, uint64_t seed, uint32_t party, uint32_t n_parties, uint64_t first_row_id,
                                       uint64_t n) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    if (n_parties == 0 || party >= n_parties) return fail(IRIS_ERR_INVALID, "party %u of %u", party, n_parties);
    const bool s = db->flags & IRIS_DB_SHARES, m = db->flags & IRIS_DB_MASKS;
    if (s && m && db->n_shares != db->n_masks) return fail(IRIS_ERR_STATE, "shares and masks lengths differ");
    const uint64_t row0 = s ? db->n_shares : db->n_masks;
    if (row0 + n > db->capacity) return fail(IRIS_ERR_INVALID, "generate exceeds capacity");
    DeviceGuard g(db->device);
    const uint64_t step = 1u << 18;     // bounded launches (grid dimension limits)
    for (uint64_t off = 0; off < n; off += step) {
        const uint64_t cnt = std::min(step, n - off);
        if (s) CK(launch_generate_party_shares(db->d_shares, seed, party, n_parties, first_row_id + off, row0 + off, cnt, db->stream));
        if (m) CK(launch_generate(nullptr, db->d_masks, seed, first_row_id + off, row0 + off, cnt, db->stream));
    }
    int rc = sync_checked(db, db->stream);
    if (rc) return rc;
    if (s) db->n_shares += n;
    if (m) db->n_masks += n;
    return IRIS_OK;
}

extern "C" int iris_db_read_shares(iris_db* db, uint64_t row_begin, uint64_t n, uint16_t* out) {
    if (!db || (!out && n)) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (!(db->flags & IRIS_DB_SHARES)) return fail(IRIS_ERR_STATE, "no shares in this shard");
    if (row_begin + n > db->n_shares) return fail(IRIS_ERR_INVALID, "row range beyond loaded shares");
    DeviceGuard g(db->device);
    int rc = ensure_stage(db);
    if (rc) return rc;
    for (uint64_t off = 0; off < n; off += kStageRows) {
        const uint64_t m = std::min(kStageRows, n - off);
        CK(launch_untile_shares(db->d_shares, row_begin + off, m, static_cast<uint16_t*>(db->d_stage), db->stream));
        CK(cudaMemcpyAsync(out + off * IRIS_BITS, db->d_stage, m * IRIS_BITS * sizeof(uint16_t), cudaMemcpyDeviceToHost, db->stream));
        CK(cudaStreamSynchronize(db->stream));
    }
    return IRIS_OK;
}

extern "C" int iris_db_read_masks(iris_db* db, uint64_t row_begin, uint64_t n, uint64_t* out) {
    if (!db || (!out && n)) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (!(db->flags & IRIS_DB_MASKS)) return fail(IRIS_ERR_STATE, "no masks in this shard");
    if (row_begin + n > db->n_masks) return fail(IRIS_ERR_INVALID, "row range beyond loaded masks");
    DeviceGuard g(db->device);
    int rc = ensure_stage(db);
    if (rc) return rc;
    const uint64_t step = kStageRows * 16;
    for (uint64_t off = 0; off < n; off += step) {
        const uint64_t m = std::min(step, n - off);
        CK(launch_untile_masks(db->d_masks, row_begin + off, m, static_cast<uint8_t*>(db->d_stage), db->stream));
        CK(cudaMemcpyAsync(out + off * IRIS_LIMBS, db->d_stage, m * IRIS_MASK_BYTES, cudaMemcpyDeviceToHost, db->stream));
        CK(cudaStreamSynchronize(db->stream));
    }
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------ scan core
static int ensure_result_buffers(iris_db* db, uint64_t rows) {
    if (db->res_rows >= rows) return IRIS_OK;
    for (int b = 0; b < 2; ++b)
        for (int k = 0; k < 2; ++k) {
            cudaFree(db->d_res[b][k]);
            db->d_res[b][k] = nullptr;
        }
    db->res_rows = 0;
    for (int b = 0; b < 2; ++b)
        for (int k = 0; k < 2; ++k) CK(cudaMalloc(&db->d_res[b][k], rows * kOutRowBytes + 64));
    db->res_rows = rows;
    return IRIS_OK;
}

// The s8 flag of an engine built from a DEVICE query is computed by a kernel; the first scan needs it on the host.
static int engine_resolve(iris_distance_engine* e) {
    if (e->classified) return IRIS_OK;
    CK(cudaEventSynchronize(e->slot.ready));
    e->fits_s8 = *reinterpret_cast<volatile int*>(e->slot.h + IRIS_BITS * 2) != 0;
    e->classified = true;
    return IRIS_OK;
}

struct Progress {                 // host outputs only: called on the calling thread as blocks of rows become complete
    iris_progress_fn fn = nullptr;
    void* user = nullptr;
};

// de / me: the engines whose operand images are scanned (nullptr = that half is not computed).
static int scan_core(iris_db* db, iris_distance_engine* de, iris_masks_engine* me, uint64_t row_begin, uint64_t row_end,
                     uint16_t* dist_out, uint16_t* den_out, int32_t* raw_dev, const Progress* progress = nullptr) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    if (!de && !me) return fail(IRIS_ERR_INVALID, "no engine given");
    const uint8_t* qd = de ? de->d_qd : nullptr;
    const uint8_t* qm = me ? me->d_qm : nullptr;
    if (row_begin > row_end) return fail(IRIS_ERR_INVALID, "row_begin > row_end");
    if (qd) {
        if (!db->d_shares) return fail(IRIS_ERR_STATE, "shard holds no shares");
        if (row_end > db->n_shares) return fail(IRIS_ERR_INVALID, "row_end %llu beyond %llu loaded shares", (unsigned long long)row_end, (unsigned long long)db->n_shares);
        if (!dist_out && row_end > row_begin) return fail(IRIS_ERR_INVALID, "distances output is NULL");
    }
    if (qm) {
        if (!db->d_masks) return fail(IRIS_ERR_STATE, "shard holds no masks");
        if (row_end > db->n_masks) return fail(IRIS_ERR_INVALID, "row_end %llu beyond %llu loaded masks", (unsigned long long)row_end, (unsigned long long)db->n_masks);
        if (!den_out && row_end > row_begin) return fail(IRIS_ERR_INVALID, "denominators output is NULL");
    }
    if (row_begin == row_end) return IRIS_OK;
    DeviceGuard g(db->device);
    bool signed_query = false;
    if (de) {
        int rc = engine_resolve(de);
        if (rc) return rc;
        signed_query = de->fits_s8;
        rc = engine_begin_use(de->slot, de->use, db);
        if (rc) return rc;
    }
    if (me) {
        int rc = engine_begin_use(me->slot, me->use, db);
        if (rc) return rc;
    }

    ScanParams p{};
    p.shares = qd ? db->d_shares : nullptr;
    p.masks = qm ? db->d_masks : nullptr;
    p.qd = qd;
    p.qm = qm;
    p.qm4 = qm ? qm + kQmBytes : nullptr;        // every mask operand buffer is [int8 image | 4-bit image]
    p.raw_out = raw_dev;
    p.error = db->d_error;
    p.signed_query = signed_query && !raw_dev;   // the raw dump shows the three-product accumulators

    const bool dist_dev = !qd || is_device_pointer(dist_out);
    const bool den_dev = !qm || is_device_pointer(den_out);
    if (dist_dev && den_dev) {
        // results stay in HBM: one persistent launch, asynchronous on the shard's stream
        p.dist_out = dist_out;
        p.den_out = den_out;
        p.row_begin = row_begin;
        p.row_end = row_end;
        p.tile_begin = (uint32_t)(row_begin / kTileRows);
        p.tile_end = (uint32_t)((row_end + kTileRows - 1) / kTileRows);
        // The reference calls batch_process chunk after chunk (src/main.rs:427-430): consecutive scans on the library's
        // own stream with disjoint outputs may overlap (the next one starts on the SMs the previous one's tail leaves
        // idle).  The ranges below are all a running scan can belong to: after kChain chained launches an ordinary launch
        // drains the chain, unless every scan of the chain is wide (see below).  The same holds on a caller-supplied stream: a scan only ever starts early behind
        // another scan of this shard (a foreign kernel in between never triggers the programmatic launch, so the scan
        // behind it starts when that kernel has completed); iris_db_set_overlap(db, 0) turns the overlap off.
        const size_t bytes = (size_t)(row_end - row_begin) * kOutRowBytes;
        const uintptr_t lo[2] = {reinterpret_cast<uintptr_t>(dist_out), reinterpret_cast<uintptr_t>(den_out)};
        const uintptr_t hi[2] = {dist_out ? lo[0] + bytes : 0, den_out ? lo[1] + bytes : 0};
        // A scan with at least num_sms / 2 CTAs (one CTA per SM) can start only when its predecessor is fully resident,
        // i.e. when at most one older scan still holds SMs: three scans in flight at the very most, so for such scans
        // the last kChain ranges are a sliding window and the chain never has to be drained.
        const uint32_t grid_ctas = std::min<uint32_t>(p.tile_end - p.tile_begin, (uint32_t)db->num_sms);
        if (db->chain_len == iris_db::kChain && db->chain_all_wide && grid_ctas >= (uint32_t)db->num_sms / 2) {
            for (int i = 1; i < iris_db::kChain; ++i)
                for (int a = 0; a < 2; ++a) {
                    db->chain_lo[i - 1][a] = db->chain_lo[i][a];
                    db->chain_hi[i - 1][a] = db->chain_hi[i][a];
                }
            --db->chain_len;
        }
        bool chain = db->overlap && !raw_dev && db->chain_len < iris_db::kChain;
        for (int i = 0; chain && i < db->chain_len; ++i)
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b)
                    if (lo[a] < db->chain_hi[i][b] && db->chain_lo[i][b] < hi[a]) chain = false;
        if (!chain) {                           // an ordinary launch waits for everything before it
            db->chain_len = 0;
            db->chain_all_wide = true;
        }
        if (grid_ctas < (uint32_t)db->num_sms / 2) db->chain_all_wide = false;
        p.pdl = chain && db->chain_len > 0;
        for (int a = 0; a < 2; ++a) {
            db->chain_lo[db->chain_len][a] = lo[a];
            db->chain_hi[db->chain_len][a] = hi[a];
        }
        ++db->chain_len;
        CK(launch_scan(p, db->num_sms, db->stream));
        return IRIS_OK;
    }
    if (raw_dev) return fail(IRIS_ERR_INVALID, "raw dump needs device outputs");

    // host outputs: chunked scan on `stream`, D2H on `copy_stream`, double-buffered
    const uint64_t chunk_rows = (uint64_t)db->num_sms * kResultChunkTilesPerSm * kTileRows;
    int rc = ensure_result_buffers(db, chunk_rows + kTileRows);
    if (rc) return rc;
    const uint64_t aligned0 = row_begin / kTileRows * kTileRows;
    uint64_t cb = row_begin;
    // Chunks of 8 tiles per SM, halving towards the end (never below one tile per SM): the last device->host copy is
    // the only one that is not hidden behind a scan, so it should be short.
    const uint64_t wave_rows = (uint64_t)db->num_sms * kTileRows;
    uint64_t next_edge = aligned0;
    uint64_t prev_b = row_begin, prev_e = row_begin;           // the chunk enqueued one iteration ago (progress reports)
    for (uint64_t i = 0; cb < row_end; ++i) {
        const uint64_t remaining = row_end > next_edge ? row_end - next_edge : 0;
        uint64_t step = chunk_rows;
        if (remaining < 2 * chunk_rows) step = std::max(wave_rows, remaining / 2 / wave_rows * wave_rows);
        next_edge += step;
        const uint64_t ce = std::min(row_end, next_edge);
        const int b = (int)(i & 1);
        if (i >= 2) CK(cudaStreamWaitEvent(db->stream, db->ev_copy[b], 0));
        p.dist_out = qd ? (dist_dev ? dist_out + (cb - row_begin) * IRIS_ROTATIONS : db->d_res[b][0]) : nullptr;
        p.den_out = qm ? (den_dev ? den_out + (cb - row_begin) * IRIS_ROTATIONS : db->d_res[b][1]) : nullptr;
        p.row_begin = cb;
        p.row_end = ce;
        p.tile_begin = (uint32_t)(cb / kTileRows);
        p.tile_end = (uint32_t)((ce + kTileRows - 1) / kTileRows);
        CK(launch_scan(p, db->num_sms, db->stream));
        CK(cudaEventRecord(db->ev_scan[b], db->stream));
        CK(cudaStreamWaitEvent(db->copy_stream, db->ev_scan[b], 0));
        const size_t bytes = (ce - cb) * kOutRowBytes;
        if (qd && !dist_dev)
            CK(cudaMemcpyAsync(dist_out + (cb - row_begin) * IRIS_ROTATIONS, db->d_res[b][0], bytes, cudaMemcpyDeviceToHost, db->copy_stream));
        if (qm && !den_dev)
            CK(cudaMemcpyAsync(den_out + (cb - row_begin) * IRIS_ROTATIONS, db->d_res[b][1], bytes, cudaMemcpyDeviceToHost, db->copy_stream));
        CK(cudaEventRecord(db->ev_copy[b], db->copy_stream));
        if (progress && progress->fn) {
            // the device now has chunk i queued behind chunk i-1: wait for i-1 to be in host memory and report it
            if (prev_e > prev_b) {
                CK(cudaEventSynchronize(db->ev_copy[b ^ 1]));
                int rce = check_error_flag(db);
                if (rce) return rce;
                progress->fn(progress->user, prev_b, prev_e);
            }
            prev_b = cb;
            prev_e = ce;
        }
        cb = ce;
    }
    int rc2 = sync_checked(db, db->copy_stream);
    if (rc2) return rc2;
    rc2 = sync_checked(db, db->stream);
    if (rc2) return rc2;
    rc2 = check_error_flag(db);
    if (rc2) return rc2;
    if (progress && progress->fn && prev_e > prev_b) progress->fn(progress->user, prev_b, prev_e);
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------ engines
static int require_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(IRIS_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(IRIS_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    return IRIS_OK;
}

static int new_distance_engine(int device, cudaStream_t s, iris_distance_engine** out) {
    iris_distance_engine* e = new (std::nothrow) iris_distance_engine();
    if (!e) return fail(IRIS_ERR_NOMEM, "host allocation failed");
    e->device = device;
    int rc = slot_acquire(kSlotDistance, device, s, &e->slot);
    if (rc) {
        delete e;
        return rc;
    }
    e->d_qd = e->slot.d;
    e->d_query = reinterpret_cast<uint16_t*>(e->slot.d + kDSlotQuery);
    *out = e;
    return IRIS_OK;
}

static int new_masks_engine(int device, cudaStream_t s, iris_masks_engine** out) {
    iris_masks_engine* e = new (std::nothrow) iris_masks_engine();
    if (!e) return fail(IRIS_ERR_NOMEM, "host allocation failed");
    e->device = device;
    int rc = slot_acquire(kSlotMasks, device, s, &e->slot);
    if (rc) {
        delete e;
        return rc;
    }
    e->d_qm = e->slot.d;
    e->d_qmask = e->slot.d + kMSlotMask;
    *out = e;
    return IRIS_OK;
}

// Engine construction is ASYNCHRONOUS: the query is copied into the slot's page-locked staging (so the caller's
// buffer is free when the call returns), the upload and the preparation kernels are queued on the calling thread's
// stream, and `ready` is recorded behind them; the first scan waits for it on the device.  No host synchronisation.
extern "C" int iris_distance_engine_new(int device, const uint16_t* query, iris_distance_engine** out) {
    if (!query || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    cudaStream_t s = cudaStreamPerThread;
    iris_distance_engine* e = nullptr;
    rc = new_distance_engine(device, s, &e);
    if (rc) return rc;
    auto body = [&]() -> int {
        if (is_device_pointer(query)) {
            // sign-extended bytes need only two limb products: classified on the device, read at the first scan
            int* h_flag = reinterpret_cast<int*>(e->slot.h + IRIS_BITS * 2);
            int* d_flag = nullptr;
            *h_flag = 1;
            CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_flag), h_flag, 0));
            CK(cudaMemcpyAsync(e->d_query, query, IRIS_BITS * sizeof(uint16_t), cudaMemcpyDeviceToDevice, s));
            CK(launch_classify_s8(e->d_query, d_flag, s));
            e->classified = false;
        } else {
            uint16_t* st = reinterpret_cast<uint16_t*>(e->slot.h);
            std::memcpy(st, query, IRIS_BITS * sizeof(uint16_t));
            uint32_t bad = 0;
            for (int k = 0; k < IRIS_BITS; ++k) bad |= (uint32_t)((uint16_t)(st[k] + 0x80u) > 0xFFu);
            e->fits_s8 = bad == 0;
            CK(cudaMemcpyAsync(e->d_query, st, IRIS_BITS * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
        }
        CK(launch_prep_distance_query(e->d_query, e->d_qd, s));
        CK(cudaEventRecord(e->slot.ready, s));
        return IRIS_OK;
    };
    rc = body();
    if (rc) {
        std::string keep = g_last_error;
        iris_distance_engine_free(e);
        g_last_error = keep;
        return rc;
    }
    *out = e;
    return IRIS_OK;
}

// encode(&Template) (src/lib.rs:16-26) on the device; `out` host or device.
extern "C" int iris_encode(int device, const uint64_t* pattern, const uint64_t* mask, uint16_t* out) {
    if (!pattern || !mask || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    uint8_t* d = nullptr;
    rc = temp_alloc(device, reinterpret_cast<void**>(&d), 2 * IRIS_MASK_BYTES + IRIS_BITS * sizeof(uint16_t));
    if (rc) return rc;
    auto body = [&]() -> int {
        uint16_t* d_out = reinterpret_cast<uint16_t*>(d + 2 * IRIS_MASK_BYTES);
        CK(cudaMemcpyAsync(d, pattern, IRIS_MASK_BYTES, cudaMemcpyDefault, cudaStreamPerThread));
        CK(cudaMemcpyAsync(d + IRIS_MASK_BYTES, mask, IRIS_MASK_BYTES, cudaMemcpyDefault, cudaStreamPerThread));
        CK(launch_encode(d, d + IRIS_MASK_BYTES, d_out, cudaStreamPerThread));
        CK(cudaMemcpyAsync(out, d_out, IRIS_BITS * sizeof(uint16_t), cudaMemcpyDefault, cudaStreamPerThread));
        CK(cudaStreamSynchronize(cudaStreamPerThread));
        return IRIS_OK;
    };
    rc = body();
    temp_free(d);
    return rc;
}

// DistanceEngine::new(&encode(&template)) as the participant does per request (src/main.rs:427): the 3 200-byte
// wire Template goes to the device, encode + rotation preparation run there.
extern "C" int iris_distance_engine_new_from_template(int device, const uint64_t* pattern, const uint64_t* mask,
                                                      iris_distance_engine** out) {
    if (!pattern || !mask || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    cudaStream_t s = cudaStreamPerThread;
    iris_distance_engine* e = nullptr;
    rc = new_distance_engine(device, s, &e);
    if (rc) return rc;
    auto body = [&]() -> int {
        uint8_t* d_t = e->slot.d + kDSlotTemplate;
        if (is_device_pointer(pattern) || is_device_pointer(mask)) {
            CK(cudaMemcpyAsync(d_t, pattern, IRIS_MASK_BYTES, cudaMemcpyDefault, s));
            CK(cudaMemcpyAsync(d_t + IRIS_MASK_BYTES, mask, IRIS_MASK_BYTES, cudaMemcpyDefault, s));
        } else {
            std::memcpy(e->slot.h, pattern, IRIS_MASK_BYTES);
            std::memcpy(e->slot.h + IRIS_MASK_BYTES, mask, IRIS_MASK_BYTES);
            CK(cudaMemcpyAsync(d_t, e->slot.h, 2 * IRIS_MASK_BYTES, cudaMemcpyHostToDevice, s));
        }
        CK(launch_encode(d_t, d_t + IRIS_MASK_BYTES, e->d_query, s));
        CK(launch_prep_distance_query(e->d_query, e->d_qd, s));
        CK(cudaEventRecord(e->slot.ready, s));
        e->fits_s8 = true;                           // encode() only yields 0, 1, 0xFFFF
        return IRIS_OK;
    };
    rc = body();
    if (rc) {
        std::string keep = g_last_error;
        iris_distance_engine_free(e);
        g_last_error = keep;
        return rc;
    }
    *out = e;
    return IRIS_OK;
}

// Q engines of each kind from Q wire Templates in one go: one H2D copy (3 200 B per query) and three or four
// launches per 64 queries, no synchronisation.
extern "C" int iris_engines_new_from_templates(int device, const uint64_t* templates, uint32_t num_queries,
                                               iris_distance_engine** distance_engines, iris_masks_engine** masks_engines) {
    if (!templates || !distance_engines) return fail(IRIS_ERR_INVALID, "NULL argument");
    for (uint32_t i = 0; i < num_queries; ++i) {
        distance_engines[i] = nullptr;
        if (masks_engines) masks_engines[i] = nullptr;
    }
    if (num_queries == 0) return IRIS_OK;
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    cudaStream_t s = cudaStreamPerThread;
    const size_t tbytes = (size_t)num_queries * 2 * IRIS_MASK_BYTES;
    uint8_t* d_t = nullptr;
    rc = temp_alloc(device, reinterpret_cast<void**>(&d_t), tbytes);
    if (rc) return rc;
    auto body = [&]() -> int {
        if (is_device_pointer(templates)) {
            CK(cudaMemcpyAsync(d_t, templates, tbytes, cudaMemcpyDeviceToDevice, s));
        } else {
            ThreadStage* st = nullptr;
            int r = thread_stage(device, tbytes, &st);
            if (r) return r;
            std::memcpy(st->p, templates, tbytes);
            CK(cudaMemcpyAsync(d_t, st->p, tbytes, cudaMemcpyHostToDevice, s));
            CK(cudaEventRecord(st->done, s));
        }
        for (uint32_t q0 = 0; q0 < num_queries; q0 += kMaxPrepBatch) {
            const uint32_t nq = std::min<uint32_t>(kMaxPrepBatch, num_queries - q0);
            PrepBatchParams p{};
            p.templates = d_t + (size_t)q0 * 2 * IRIS_MASK_BYTES;
            p.n = nq;
            for (uint32_t i = 0; i < nq; ++i) {
                iris_distance_engine* e = nullptr;
                int r = new_distance_engine(device, s, &e);
                if (r) return r;
                distance_engines[q0 + i] = e;
                e->fits_s8 = true;                           // encode() only yields 0, 1, 0xFFFF
                p.query[i] = e->d_query;
                p.qd[i] = e->d_qd;
                if (masks_engines) {
                    iris_masks_engine* m = nullptr;
                    r = new_masks_engine(device, s, &m);
                    if (r) return r;
                    masks_engines[q0 + i] = m;
                    p.qm[i] = m->d_qm;
                }
            }
            CK(launch_prep_batch(p, s));
            for (uint32_t i = 0; i < nq; ++i) {
                CK(cudaEventRecord(distance_engines[q0 + i]->slot.ready, s));
                if (masks_engines) CK(cudaEventRecord(masks_engines[q0 + i]->slot.ready, s));
            }
        }
        return IRIS_OK;
    };
    rc = body();
    temp_free(d_t);
    if (rc) {
        std::string keep = g_last_error;
        for (uint32_t i = 0; i < num_queries; ++i) {
            iris_distance_engine_free(distance_engines[i]);
            distance_engines[i] = nullptr;
            if (masks_engines) {
                iris_masks_engine_free(masks_engines[i]);
                masks_engines[i] = nullptr;
            }
        }
        g_last_error = keep;
    }
    return rc;
}

// Never blocks: the buffers go back to the pool with an event recorded behind the last scan that reads them, and
// the next engine built from them waits for that event on the device.
extern "C" int iris_distance_engine_free(iris_distance_engine* e) {
    if (!e) return IRIS_OK;
    DeviceGuard g(e->device);
    if (e->scratch) iris_db_destroy(e->scratch);
    engine_release(kSlotDistance, e->slot, e->use);
    delete e;
    return IRIS_OK;
}

extern "C" int iris_masks_engine_new(int device, const uint64_t* query_mask, iris_masks_engine** out) {
    if (!query_mask || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    cudaStream_t s = cudaStreamPerThread;
    iris_masks_engine* e = nullptr;
    rc = new_masks_engine(device, s, &e);
    if (rc) return rc;
    auto body = [&]() -> int {
        if (is_device_pointer(query_mask)) {
            CK(cudaMemcpyAsync(e->d_qmask, query_mask, IRIS_MASK_BYTES, cudaMemcpyDeviceToDevice, s));
        } else {
            std::memcpy(e->slot.h, query_mask, IRIS_MASK_BYTES);
            CK(cudaMemcpyAsync(e->d_qmask, e->slot.h, IRIS_MASK_BYTES, cudaMemcpyHostToDevice, s));
        }
        CK(launch_prep_mask_query(e->d_qmask, e->d_qm, e->d_qm + kQmBytes, s));
        CK(cudaEventRecord(e->slot.ready, s));
        return IRIS_OK;
    };
    rc = body();
    if (rc) {
        std::string keep = g_last_error;
        iris_masks_engine_free(e);
        g_last_error = keep;
        return rc;
    }
    *out = e;
    return IRIS_OK;
}

extern "C" int iris_masks_engine_free(iris_masks_engine* e) {
    if (!e) return IRIS_OK;
    DeviceGuard g(e->device);
    if (e->scratch) iris_db_destroy(e->scratch);
    engine_release(kSlotMasks, e->slot, e->use);
    delete e;
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------ literal host-slice batch_process
// The reference's exact shape: `db` is a HOST slice.  Rows travel host -> staging -> re-tiled scratch shard -> scan ->
// results -> host in chunks, double buffered: the upload and re-tiling of chunk i+1 (copy stream) overlap the scan and
// the result download of chunk i (scan stream); one synchronisation at the end.  PCIe-bound by construction.
static constexpr uint64_t kSliceChunkShares = 2048;        // 52 MB of EncodedBits per chunk
static constexpr uint64_t kSliceChunkMasks = 32768;        // 52 MB of Bits per chunk

struct SlicePipe {                                         // lives in the scratch shard of an engine
    cudaEvent_t loaded[2] = {nullptr, nullptr}, scanned[2] = {nullptr, nullptr};
};

template <bool SHARES>
static int slice_batch_process(iris_distance_engine* de, iris_masks_engine* me, iris_db*& scratch, int device,
                               uint16_t* out, const void* rows, uint64_t n) {
    const uint64_t chunk = SHARES ? kSliceChunkShares : kSliceChunkMasks;
    const size_t row_bytes = SHARES ? IRIS_BITS * sizeof(uint16_t) : IRIS_MASK_BYTES;
    // scratch shard: two halves of `cap` rows each, sized to the call (rounded to whole 256-row pair tiles)
    const uint64_t cap = (std::min(n, chunk) + 255) / 256 * 256;
    if (scratch && scratch->capacity < 2 * cap) {
        iris_db_destroy(scratch);
        scratch = nullptr;
    }
    if (!scratch) {
        int rc = iris_db_create(device, 2 * cap, SHARES ? IRIS_DB_SHARES : IRIS_DB_MASKS, &scratch);
        if (rc) return rc;
    }
    iris_db* db = scratch;
    DeviceGuard g(device);
    const uint64_t half = db->capacity / 2;
    if (SHARES) db->n_shares = db->capacity; else db->n_masks = db->capacity;   // every row of the scratch is addressable
    // device staging for two chunks of reference-layout rows, device results for two chunks
    if (!db->d_stage) CK(cudaMalloc(&db->d_stage, 2 * half * row_bytes));
    int rc = ensure_result_buffers(db, half);
    if (rc) return rc;
    SlicePipe pipe;
    auto body = [&]() -> int {
        for (int b = 0; b < 2; ++b) {
            CK(cudaEventCreateWithFlags(&pipe.loaded[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&pipe.scanned[b], cudaEventDisableTiming));
        }
        uint64_t i = 0;
        for (uint64_t off = 0; off < n; off += half, ++i) {
            const int b = (int)(i & 1);
            const uint64_t m = std::min(half, n - off);
            uint8_t* stage = static_cast<uint8_t*>(db->d_stage) + (size_t)b * half * row_bytes;
            if (i >= 2) CK(cudaStreamWaitEvent(db->copy_stream, pipe.scanned[b], 0));    // half b free again
            CK(cudaMemcpyAsync(stage, static_cast<const uint8_t*>(rows) + off * row_bytes, m * row_bytes, cudaMemcpyHostToDevice, db->copy_stream));
            if (SHARES) CK(launch_retile_shares(reinterpret_cast<const uint16_t*>(stage), m, db->d_shares, (uint64_t)b * half, db->copy_stream));
            else CK(launch_retile_masks(stage, m, db->d_masks, (uint64_t)b * half, db->copy_stream));
            CK(cudaEventRecord(pipe.loaded[b], db->copy_stream));
            CK(cudaStreamWaitEvent(db->stream, pipe.loaded[b], 0));
            uint16_t* d_out = db->d_res[b][SHARES ? 0 : 1];
            int r = scan_core(db, de, me, (uint64_t)b * half, (uint64_t)b * half + m, SHARES ? d_out : nullptr,
                              SHARES ? nullptr : d_out, nullptr);
            if (r) return r;
            CK(cudaMemcpyAsync(out + off * IRIS_ROTATIONS, d_out, m * kOutRowBytes, cudaMemcpyDeviceToHost, db->stream));
            CK(cudaEventRecord(pipe.scanned[b], db->stream));
        }
        return IRIS_OK;
    };
    rc = body();
    int rc2 = sync_checked(db, db->copy_stream);
    int rc3 = sync_checked(db, db->stream);
    for (int b = 0; b < 2; ++b) {
        if (pipe.loaded[b]) cudaEventDestroy(pipe.loaded[b]);
        if (pipe.scanned[b]) cudaEventDestroy(pipe.scanned[b]);
    }
    if (rc) return rc;
    if (rc2) return rc2;
    if (rc3) return rc3;
    return check_error_flag(db);
}

extern "C" int iris_distance_engine_batch_process(iris_distance_engine* e, uint16_t* out, uint64_t out_len,
                                                  const uint16_t* db, uint64_t db_len) {
    if (!e) return fail(IRIS_ERR_INVALID, "engine is NULL");
    // reference: assert_eq!(out.len(), db.len())  (src/lib.rs:43)
    if (out_len != db_len) return fail(IRIS_ERR_INVALID, "out.len() (%llu) != db.len() (%llu)", (unsigned long long)out_len, (unsigned long long)db_len);
    if (db_len == 0) return IRIS_OK;
    if (!out || !db) return fail(IRIS_ERR_INVALID, "NULL slice");
    if (is_device_pointer(db) || is_device_pointer(out)) return fail(IRIS_ERR_INVALID, "batch_process takes HOST slices (use a resident shard for device data)");
    return slice_batch_process<true>(e, nullptr, e->scratch, e->device, out, db, db_len);
}

extern "C" int iris_masks_engine_batch_process(iris_masks_engine* e, uint16_t* out, uint64_t out_len,
                                               const uint64_t* db, uint64_t db_len) {
    if (!e) return fail(IRIS_ERR_INVALID, "engine is NULL");
    // reference: assert_eq!(out.len(), db.len())  (src/lib.rs:70)
    if (out_len != db_len) return fail(IRIS_ERR_INVALID, "out.len() (%llu) != db.len() (%llu)", (unsigned long long)out_len, (unsigned long long)db_len);
    if (db_len == 0) return IRIS_OK;
    if (!out || !db) return fail(IRIS_ERR_INVALID, "NULL slice");
    if (is_device_pointer(db) || is_device_pointer(out)) return fail(IRIS_ERR_INVALID, "batch_process takes HOST slices (use a resident shard for device data)");
    return slice_batch_process<false>(nullptr, e, e->scratch, e->device, out, db, db_len);
}

extern "C" int iris_distance_engine_batch_process_resident(iris_distance_engine* e, uint16_t* out, uint64_t out_len,
                                                           iris_db* db, uint64_t row_begin, uint64_t row_end) {
    if (!e || !db) return fail(IRIS_ERR_INVALID, "NULL handle");
    if (e->device != db->device) return fail(IRIS_ERR_INVALID, "engine and shard live on different devices");
    if (row_end < row_begin || out_len != row_end - row_begin)
        return fail(IRIS_ERR_INVALID, "out.len() (%llu) != db.len() (%llu)", (unsigned long long)out_len, (unsigned long long)(row_end - row_begin));
    return scan_core(db, e, nullptr, row_begin, row_end, out, nullptr, nullptr);
}

extern "C" int iris_masks_engine_batch_process_resident(iris_masks_engine* e, uint16_t* out, uint64_t out_len,
                                                        iris_db* db, uint64_t row_begin, uint64_t row_end) {
    if (!e || !db) return fail(IRIS_ERR_INVALID, "NULL handle");
    if (e->device != db->device) return fail(IRIS_ERR_INVALID, "engine and shard live on different devices");
    if (row_end < row_begin || out_len != row_end - row_begin)
        return fail(IRIS_ERR_INVALID, "out.len() (%llu) != db.len() (%llu)", (unsigned long long)out_len, (unsigned long long)(row_end - row_begin));
    return scan_core(db, nullptr, e, row_begin, row_end, nullptr, out, nullptr);
}

extern "C" int iris_match_resident(iris_distance_engine* de, iris_masks_engine* me, iris_db* db, uint64_t row_begin,
                                   uint64_t row_end, uint16_t* distances_out, uint16_t* denominators_out) {
    if (!db || (!de && !me)) return fail(IRIS_ERR_INVALID, "NULL handle");
    if ((de && de->device != db->device) || (me && me->device != db->device))
        return fail(IRIS_ERR_INVALID, "engine and shard live on different devices");
    return scan_core(db, de, me, row_begin, row_end, distances_out, denominators_out, nullptr);
}

extern "C" int iris_match_resident_streamed(iris_distance_engine* de, iris_masks_engine* me, iris_db* db, uint64_t row_begin,
                                            uint64_t row_end, uint16_t* distances_out, uint16_t* denominators_out,
                                            iris_progress_fn progress, void* user) {
    if (!db || (!de && !me)) return fail(IRIS_ERR_INVALID, "NULL handle");
    if ((de && de->device != db->device) || (me && me->device != db->device))
        return fail(IRIS_ERR_INVALID, "engine and shard live on different devices");
    if ((de && is_device_pointer(distances_out)) || (me && is_device_pointer(denominators_out)))
        return fail(IRIS_ERR_INVALID, "the streamed form takes HOST result arrays");
    Progress pr;
    pr.fn = progress;
    pr.user = user;
    return scan_core(db, de, me, row_begin, row_end, distances_out, denominators_out, nullptr, &pr);
}

extern "C" int iris_db_set_overlap(iris_db* db, int allow) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    DeviceGuard g(db->device);
    CK(cudaStreamSynchronize(db->stream));
    db->overlap = allow != 0;
    db->chain_len = 0;
    return IRIS_OK;
}

extern "C" int iris_distances(int device, const uint16_t* query, const uint16_t* entry, uint16_t* out) {
    if (!query || !entry || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    iris_distance_engine* e = nullptr;
    int rc = iris_distance_engine_new(device, query, &e);
    if (rc) return rc;
    rc = iris_distance_engine_batch_process(e, out, 1, entry, 1);
    std::string keep = g_last_error;
    iris_distance_engine_free(e);
    g_last_error = keep;
    return rc;
}

extern "C" int iris_denominators(int device, const uint64_t* query, const uint64_t* entry, uint16_t* out) {
    if (!query || !entry || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    iris_masks_engine* e = nullptr;
    int rc = iris_masks_engine_new(device, query, &e);
    if (rc) return rc;
    rc = iris_masks_engine_batch_process(e, out, 1, entry, 1);
    std::string keep = g_last_error;
    iris_masks_engine_free(e);
    g_last_error = keep;
    return rc;
}

// ------------------------------------------------------------------------------------ batched queries (dense GEMM)
constexpr uint32_t kBatchDistanceGroup = 8;     // queries per accumulator tile of batch_distances_kernel
// Wave pacing of the batched GEMM (iris_batch.cu).  Diagnostics build: IRIS_BATCH_SYNC=0/1 for A/B runs.
static bool batch_wave_pacing() {
#ifdef IRIS_DIAGNOSTICS
    static const int v = [] {
        const char* e = getenv("IRIS_BATCH_SYNC");
        return e ? atoi(e) : kBatchWavePacingDefault;
    }();
    return v != 0;
#else
    return kBatchWavePacingDefault != 0;
#endif
}
#ifdef IRIS_DIAGNOSTICS
constexpr uint32_t kBatchMaskGroup = 16;        // query masks per accumulator tile of batch_denominators_kernel
constexpr uint32_t kBatchMaskTailLoop = 10;     // left-over masks handled by the single-query scan instead
#endif
extern "C" int iris_distances_batch_resident(iris_distance_engine* const* engines, uint32_t num_queries, iris_db* db,
                                             uint64_t row_begin, uint64_t row_end, uint16_t* out) {
    if (!engines || !db) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (num_queries == 0) return IRIS_OK;
    if (row_begin > row_end) return fail(IRIS_ERR_INVALID, "row_begin > row_end");
    if (!db->d_shares) return fail(IRIS_ERR_STATE, "shard holds no shares");
    if (row_end > db->n_shares) return fail(IRIS_ERR_INVALID, "row_end beyond loaded shares");
    if (row_begin == row_end) return IRIS_OK;
    if (!out) return fail(IRIS_ERR_INVALID, "out is NULL");
    for (uint32_t i = 0; i < num_queries; ++i) {
        if (!engines[i]) return fail(IRIS_ERR_INVALID, "engine %u is NULL", i);
        if (engines[i]->device != db->device) return fail(IRIS_ERR_INVALID, "engine %u lives on another device", i);
    }
    DeviceGuard g(db->device);
    const uint64_t rows = row_end - row_begin;
    const bool out_dev = is_device_pointer(out);
    uint16_t* d_out = out;
    const size_t out_bytes = (size_t)num_queries * rows * kOutRowBytes;
    for (uint32_t i = 0; i < num_queries; ++i) {
        int rc0 = engine_resolve(engines[i]);
        if (!rc0) rc0 = engine_begin_use(engines[i]->slot, engines[i]->use, db);
        if (rc0) return rc0;
    }
    if (!out_dev) {                      // host result: device staging from the stream-ordered pool (no cudaMalloc per call)
        int rc0 = temp_alloc(db->device, reinterpret_cast<void**>(&d_out), out_bytes + 64, db->stream);
        if (rc0) return rc0;
    }
    auto body = [&]() -> int {
        for (uint32_t q0 = 0; q0 < num_queries; q0 += kMaxBatchQueries) {
            const uint32_t nq = std::min<uint32_t>(kMaxBatchQueries, num_queries - q0);
            // A group of 8 queries costs one pass of the int8 GEMM (~4.5 ms per 1 M rows) however few of its slots
            // are used; a single left-over query is cheaper on the HBM-bound single-query scan (~3.6 ms).
            if (nq % kBatchDistanceGroup == 1) {
                iris_distance_engine* e = engines[q0 + nq - 1];
                int rc1 = scan_core(db, e, nullptr, row_begin, row_end,
                                    d_out + (size_t)(q0 + nq - 1) * rows * IRIS_ROTATIONS, nullptr, nullptr);
                if (rc1) return rc1;
                if (nq == 1) continue;
            }
            const uint32_t nb = nq % kBatchDistanceGroup == 1 ? nq - 1 : nq;
            BatchParams p{};
            p.shares = db->d_shares;
            bool all_s8 = true;
            for (uint32_t i = 0; i < nb; ++i) {
                p.qd[i] = engines[q0 + i]->d_qd;
                all_s8 &= engines[q0 + i]->fits_s8;
            }
            p.out = d_out + (size_t)q0 * rows * IRIS_ROTATIONS;
            p.row_begin = row_begin;
            p.row_end = row_end;
            p.pair_begin = (uint32_t)(row_begin / (2 * kTileRows));
            p.pair_end = (uint32_t)((row_end + 2 * kTileRows - 1) / (2 * kTileRows));
            p.num_queries = nb;
            p.error = db->d_error;
            if (batch_wave_pacing()) {
                if (!db->d_wave_sync) CK(cudaMalloc(reinterpret_cast<void**>(&db->d_wave_sync), 64));
                CK(cudaMemsetAsync(db->d_wave_sync, 0, 8, db->stream));
                p.wave_sync = db->d_wave_sync;
            }
            CK(launch_batch_distances(p, all_s8, db->num_sms, db->stream));
        }
        if (!out_dev) {
            CK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, db->stream));
            temp_free(d_out, db->stream);
            d_out = nullptr;
            int rc1 = sync_checked(db, db->stream);
            if (rc1) return rc1;
            return check_error_flag(db);
        }
        return IRIS_OK;
    };
    int rc = body();
    if (!out_dev && d_out) temp_free(d_out, db->stream);
    return rc;
}

extern "C" int iris_denominators_batch_resident(iris_masks_engine* const* engines, uint32_t num_queries, iris_db* db,
                                                uint64_t row_begin, uint64_t row_end, uint16_t* out) {
    if (!engines || !db) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (num_queries == 0) return IRIS_OK;
    if (row_begin > row_end) return fail(IRIS_ERR_INVALID, "row_begin > row_end");
    if (!db->d_masks) return fail(IRIS_ERR_STATE, "shard holds no masks");
    if (row_end > db->n_masks) return fail(IRIS_ERR_INVALID, "row_end beyond loaded masks");
    if (row_begin == row_end) return IRIS_OK;
    if (!out) return fail(IRIS_ERR_INVALID, "out is NULL");
    for (uint32_t i = 0; i < num_queries; ++i) {
        if (!engines[i]) return fail(IRIS_ERR_INVALID, "engine %u is NULL", i);
        if (engines[i]->device != db->device) return fail(IRIS_ERR_INVALID, "engine %u lives on another device", i);
    }
    DeviceGuard g(db->device);
    const uint64_t rows = row_end - row_begin;
    const bool out_dev = is_device_pointer(out);
    uint16_t* d_out = out;
    const size_t out_bytes = (size_t)num_queries * rows * kOutRowBytes;
    for (uint32_t i = 0; i < num_queries; ++i) {
        int rc0 = engine_begin_use(engines[i]->slot, engines[i]->use, db);
        if (rc0) return rc0;
    }
    if (!out_dev) {
        int rc0 = temp_alloc(db->device, reinterpret_cast<void**>(&d_out), out_bytes + 64, db->stream);
        if (rc0) return rc0;
    }
    // Four query masks per pass of the 4-bit TMEM-operand scan (mask_scan_fp4_multi_kernel); a single left-over mask
    // goes to the single-query scan.  (Diagnostics build only: IRIS_BATCHDEN=i8 selects the int8 GEMM kernel.)
#ifdef IRIS_DIAGNOSTICS
    static const bool use_i8_gemm = [] {
        const char* e = getenv("IRIS_BATCHDEN");
        return e && e[0] == 'i';
    }();
#else
    constexpr bool use_i8_gemm = false;
#endif
    auto body = [&]() -> int {
        if (!use_i8_gemm) {
            uint32_t q = 0;
            // a pass costs the same for two, three or four masks (0.49 ms per 1 M rows against 0.28 ms for a single-query
            // scan), so two or three left-over masks still take one pass: the last mask fills the unused slots and is
            // simply stored again
            for (; q + 2 <= num_queries; q += kMaskMultiQueries) {
                MultiMaskScanParams p{};
                p.masks = db->d_masks;
                for (int i = 0; i < kMaskMultiQueries; ++i) {
                    const uint32_t qi = std::min<uint32_t>(q + i, num_queries - 1);
                    p.qm4[i] = engines[qi]->d_qm + kQmBytes;         // every mask operand buffer is [int8 image | 4-bit image]
                    p.out[i] = d_out + (size_t)qi * rows * IRIS_ROTATIONS;
                }
                p.row_begin = row_begin;
                p.row_end = row_end;
                p.tile_begin = (uint32_t)(row_begin / kTileRows);
                p.tile_end = (uint32_t)((row_end + kTileRows - 1) / kTileRows);
                p.error = db->d_error;
                CK(launch_mask_scan_fp4_multi(p, db->num_sms, db->stream));
            }
            for (; q < num_queries; ++q) {      // at most one mask is left
                int rc1 = scan_core(db, nullptr, engines[q], row_begin, row_end, nullptr,
                                    d_out + (size_t)q * rows * IRIS_ROTATIONS, nullptr);
                if (rc1) return rc1;
            }
        }
#ifdef IRIS_DIAGNOSTICS
        else
        for (uint32_t q0 = 0; q0 < num_queries; q0 += kMaxBatchQueries) {
            const uint32_t nq = std::min<uint32_t>(kMaxBatchQueries, num_queries - q0);
            // A group of 16 query masks costs one pass of the int8 GEMM (~3.2 ms per 1 M rows) however few of its
            // slots are used; up to kBatchMaskTailLoop left-over masks are cheaper one by one on the 4-bit
            // single-query scan (~0.28 ms each).
            const uint32_t tail = nq % kBatchMaskGroup;
            const uint32_t nb = tail <= kBatchMaskTailLoop ? nq - tail : nq;
            for (uint32_t i = nb; i < nq; ++i) {
                int rc1 = scan_core(db, nullptr, engines[q0 + i], row_begin, row_end, nullptr,
                                    d_out + (size_t)(q0 + i) * rows * IRIS_ROTATIONS, nullptr);
                if (rc1) return rc1;
            }
            if (nb == 0) continue;
            BatchMaskParams p{};
            p.masks = db->d_masks;
            for (uint32_t i = 0; i < nb; ++i) p.qm[i] = engines[q0 + i]->d_qm;
            p.out = d_out + (size_t)q0 * rows * IRIS_ROTATIONS;
            p.row_begin = row_begin;
            p.row_end = row_end;
            p.pair_begin = (uint32_t)(row_begin / (2 * kTileRows));
            p.pair_end = (uint32_t)((row_end + 2 * kTileRows - 1) / (2 * kTileRows));
            p.num_queries = nb;
            p.error = db->d_error;
            CK(launch_batch_denominators(p, db->num_sms, db->stream));
        }
#endif
        if (!out_dev) {
            CK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, db->stream));
            temp_free(d_out, db->stream);
            d_out = nullptr;
            int rc1 = sync_checked(db, db->stream);
            if (rc1) return rc1;
            return check_error_flag(db);
        }
        return IRIS_OK;
    };
    int rc = body();
    if (!out_dev && d_out) temp_free(d_out, db->stream);
    return rc;
}

// ------------------------------------------------------------------------------------ flat-file loader (f-2)
// The reference's on-disk formats (written by `prepare`, src/main.rs:337-371; mapped by the participant and the
// coordinator, src/main.rs:386-400, 458-461): raw EncodedBits rows (25 600 B) / raw Bits rows (1 600 B), the row
// index being the join key.  file -> pinned staging (pread) -> H2D -> retile, double buffered.
static int load_file(iris_db* db, const char* path, uint64_t first_row, uint64_t n_rows, bool shares) {
    if (!db || !path) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (shares && !(db->flags & IRIS_DB_SHARES)) return fail(IRIS_ERR_STATE, "shard was created without IRIS_DB_SHARES");
    if (!shares && !(db->flags & IRIS_DB_MASKS)) return fail(IRIS_ERR_STATE, "shard was created without IRIS_DB_MASKS");
    const size_t row_bytes = shares ? IRIS_BITS * sizeof(uint16_t) : IRIS_MASK_BYTES;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(IRIS_ERR_INVALID, "cannot open %s", path);
    int rc = IRIS_OK;
    void* pinned[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    DeviceGuard g(db->device);
    auto body = [&]() -> int {
        if (fseeko(f, 0, SEEK_END) != 0) return fail(IRIS_ERR_INVALID, "cannot seek %s", path);
        const uint64_t size = (uint64_t)ftello(f);
        // reference: try_cast_slice fails -> "Share file invalid" / "Masks file invalid" (src/main.rs:391-392, 460-461)
        if (size % row_bytes) return fail(IRIS_ERR_INVALID, "%s: size %llu is not a multiple of %zu", path, (unsigned long long)size, row_bytes);
        const uint64_t file_rows = size / row_bytes;
        if (first_row > file_rows) return fail(IRIS_ERR_INVALID, "%s: first_row beyond the %llu rows of the file", path, (unsigned long long)file_rows);
        const uint64_t n = n_rows ? n_rows : file_rows - first_row;
        if (first_row + n > file_rows) return fail(IRIS_ERR_INVALID, "%s: row range beyond the %llu rows of the file", path, (unsigned long long)file_rows);
        uint64_t& have = shares ? db->n_shares : db->n_masks;
        if (have + n > db->capacity) return fail(IRIS_ERR_INVALID, "load of %llu rows exceeds capacity %llu", (unsigned long long)n, (unsigned long long)db->capacity);
        const uint64_t chunk_rows = shares ? kStageRows / 2 : kStageRows * 8;      // 26 MB per buffer
        const size_t chunk_bytes = chunk_rows * row_bytes;
        int r = ensure_stage(db);                                                 // device staging holds two chunks
        if (r) return r;
        for (int i = 0; i < 2; ++i) {
            CK(cudaHostAlloc(&pinned[i], chunk_bytes, cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        if (fseeko(f, (off_t)(first_row * row_bytes), SEEK_SET) != 0) return fail(IRIS_ERR_INVALID, "cannot seek %s", path);
        uint64_t off = 0;
        for (uint64_t i = 0; off < n; ++i) {
            const int b = (int)(i & 1);
            const uint64_t m = std::min(chunk_rows, n - off);
            if (i >= 2) CK(cudaEventSynchronize(done[b]));                        // buffer b free again
            if (fread(pinned[b], row_bytes, m, f) != m) return fail(IRIS_ERR_INVALID, "%s: short read", path);
            uint8_t* d_stage = static_cast<uint8_t*>(db->d_stage) + (size_t)b * chunk_bytes;
            CK(cudaMemcpyAsync(d_stage, pinned[b], m * row_bytes, cudaMemcpyHostToDevice, db->stream));
            if (shares) CK(launch_retile_shares(reinterpret_cast<const uint16_t*>(d_stage), m, db->d_shares, have + off, db->stream));
            else CK(launch_retile_masks(d_stage, m, db->d_masks, have + off, db->stream));
            CK(cudaEventRecord(done[b], db->stream));
            off += m;
        }
        CK(cudaStreamSynchronize(db->stream));
        have += n;
        return IRIS_OK;
    };
    rc = body();
    cudaStreamSynchronize(db->stream);
    for (int i = 0; i < 2; ++i) {
        if (pinned[i]) cudaFreeHost(pinned[i]);
        if (done[i]) cudaEventDestroy(done[i]);
    }
    fclose(f);
    return rc;
}

extern "C" int iris_db_load_shares_file(iris_db* db, const char* path, uint64_t first_row, uint64_t n_rows) {
    return load_file(db, path, first_row, n_rows, true);
}
extern "C" int iris_db_load_masks_file(iris_db* db, const char* path, uint64_t first_row, uint64_t n_rows) {
    return load_file(db, path, first_row, n_rows, false);
}

// ------------------------------------------------------------------------------------ coordinator reduction (f-1)
extern "C" int iris_combine_min(int device, const uint16_t* const* distance_shares, uint32_t parties,
                                const uint16_t* denominators, uint64_t n, uint64_t index_base, double* distances_out,
                                double* min_distance, uint64_t* min_index) {
    if (!distance_shares || !denominators || !min_distance || !min_index) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (parties == 0 || parties > (uint32_t)kMaxParties) return fail(IRIS_ERR_INVALID, "parties must be in [1,%d]", kMaxParties);
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    std::vector<void*> owned;
    auto cleanup = [&]() {
        for (void* q : owned) temp_free(q);
    };
    auto to_device = [&](const void* src, size_t bytes, const void** out) -> int {
        if (is_device_pointer(src) || bytes == 0) {
            *out = src;
            return IRIS_OK;
        }
        void* d = nullptr;
        int r = temp_alloc(device, &d, bytes);
        if (r) return r;
        owned.push_back(d);
        CK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
        *out = d;
        return IRIS_OK;
    };
    auto body = [&]() -> int {
        CombineParams p{};
        p.parties = parties;
        p.n = n;
        p.index_base = index_base;
        const size_t bytes = n * kOutRowBytes;
        for (uint32_t i = 0; i < parties; ++i) {
            if (!distance_shares[i]) return fail(IRIS_ERR_INVALID, "share %u is NULL", i);
            int r = to_device(distance_shares[i], bytes, reinterpret_cast<const void**>(&p.shares[i]));
            if (r) return r;
        }
        int r = to_device(denominators, bytes, reinterpret_cast<const void**>(&p.denominators));
        if (r) return r;
        const bool dist_dev = distances_out && is_device_pointer(distances_out);
        double* d_dist = nullptr;
        if (distances_out) {
            if (dist_dev) d_dist = distances_out;
            else {
                r = temp_alloc(device, reinterpret_cast<void**>(&d_dist), n * sizeof(double) + 8);
                if (r) return r;
                owned.push_back(d_dist);
            }
        }
        p.distances_out = d_dist;
        void* scratch = nullptr;
        r = temp_alloc(device, &scratch, combine_scratch_bytes(n) + 16);
        if (r) return r;
        owned.push_back(scratch);
        void* result = static_cast<uint8_t*>(scratch) + (combine_scratch_bytes(n) / 16) * 16;
        CK(launch_combine_min(p, scratch, result, cudaStreamPerThread));
        struct { double v; unsigned long long i; } h;
        CK(cudaMemcpyAsync(&h, result, sizeof h, cudaMemcpyDeviceToHost, cudaStreamPerThread));
        if (distances_out && !dist_dev)
            CK(cudaMemcpyAsync(distances_out, d_dist, n * sizeof(double), cudaMemcpyDeviceToHost, cudaStreamPerThread));
        CK(cudaStreamSynchronize(cudaStreamPerThread));
        *min_distance = h.v;
        *min_index = h.i;
        return IRIS_OK;
    };
    rc = body();
    cudaStreamSynchronize(cudaStreamPerThread);
    cleanup();
    return rc;
}

// Reduction of a whole batch: distances / denominators are [Q][n][31] DEVICE arrays (the outputs of the batched
// kernels); one (min, argmin) pair per query comes back.  One scratch allocation, 2Q launches, one sync.
extern "C" int iris_combine_min_batch(int device, const uint16_t* distances, const uint16_t* denominators,
                                      uint32_t num_queries, uint64_t n, uint64_t index_base, double* min_distance,
                                      uint64_t* min_index) {
    if (!distances || !denominators || !min_distance || !min_index) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (num_queries == 0) return IRIS_OK;
    int rc = require_device(device);
    if (rc) return rc;
    if (!is_device_pointer(distances) || !is_device_pointer(denominators))
        return fail(IRIS_ERR_INVALID, "iris_combine_min_batch takes device arrays");
    DeviceGuard g(device);
    const size_t sbytes = ((size_t)num_queries * combine_scratch_bytes(n) + 15) / 16 * 16;
    uint8_t* scratch = nullptr;
    rc = temp_alloc(device, reinterpret_cast<void**>(&scratch), sbytes + 16 * (size_t)num_queries);
    if (rc) return rc;
    auto body = [&]() -> int {
        CombineParams p{};
        p.shares[0] = distances;
        p.parties = 1;
        p.denominators = denominators;
        p.n = n;
        p.index_base = index_base;
        p.query_stride = (size_t)n * IRIS_ROTATIONS;
        CK(launch_combine_min(p, scratch, scratch + sbytes, cudaStreamPerThread, num_queries));
        std::vector<uint64_t> host(2 * (size_t)num_queries);
        CK(cudaMemcpyAsync(host.data(), scratch + sbytes, 16 * (size_t)num_queries, cudaMemcpyDeviceToHost, cudaStreamPerThread));
        CK(cudaStreamSynchronize(cudaStreamPerThread));
        for (uint32_t q = 0; q < num_queries; ++q) {
            std::memcpy(&min_distance[q], &host[2 * q], sizeof(double));
            min_index[q] = host[2 * q + 1];
        }
        return IRIS_OK;
    };
    rc = body();
    temp_free(scratch);
    cudaStreamSynchronize(cudaStreamPerThread);
    return rc;
}

// Fused scan + reduction on a resident shard holding the full (n = 1 share) encodings: both engines over
// rows [row_begin,row_end), then decode + min/argmin on the device; only 16 bytes come back.
// Asynchronous core: the pair is written to `result` -- device memory of this or (with peer access enabled) another
// GPU, or mapped host memory -- in stream order on the shard's stream.  No host synchronisation.
extern "C" int iris_match_min_resident_async(iris_distance_engine* de, iris_masks_engine* me, iris_db* db, uint64_t row_begin,
                                             uint64_t row_end, uint64_t index_base, void* result) {
    if (!de || !me || !db || !result) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (row_end < row_begin) return fail(IRIS_ERR_INVALID, "row_begin > row_end");
    if (de->device != db->device || me->device != db->device) return fail(IRIS_ERR_INVALID, "engine and shard live on different devices");
    if (!db->d_shares || !db->d_masks) return fail(IRIS_ERR_STATE, "a search needs shares and masks in the shard");
    if (row_end > db->n_shares || row_end > db->n_masks) return fail(IRIS_ERR_INVALID, "row_end beyond the loaded rows");
    DeviceGuard g(db->device);
    // one (min, row) pair per CTA of the scan: the scan's epilogue decodes and reduces (search mode), nothing per row
    // is written or read back
    if (!db->d_red) {
        CK(cudaMalloc(&db->d_red, 2 * 1024 * sizeof(double)));
        db->red_rows = 1024;
    }
    int rc = engine_resolve(de);
    if (!rc) rc = engine_begin_use(de->slot, de->use, db);
    if (!rc) rc = engine_begin_use(me->slot, me->use, db);
    if (rc) return rc;
    ScanParams p{};
    p.shares = db->d_shares;
    p.masks = db->d_masks;
    p.qd = de->d_qd;
    p.qm = me->d_qm;
    p.error = db->d_error;
    p.signed_query = de->fits_s8;
    p.row_begin = row_begin;
    p.row_end = row_end;
    p.tile_begin = (uint32_t)(row_begin / kTileRows);
    p.tile_end = (uint32_t)((row_end + kTileRows - 1) / kTileRows);
    p.red_min = reinterpret_cast<double*>(db->d_red);
    p.red_idx = reinterpret_cast<unsigned long long*>(db->d_red + 1024 * sizeof(double));
    p.index_base = index_base;
    const uint32_t blocks = row_end > row_begin ? scan_grid(p, db->num_sms) : 0;
    if (blocks > 1024) return fail(IRIS_ERR_STATE, "scan grid of %u CTAs", blocks);
    // an ordinary launch: it starts when every earlier scan of this shard has completed (the per-CTA pairs are reused
    // from search to search), and it ends the chain of overlapping scans
    db->chain_len = 0;
    db->chain_all_wide = true;
    if (blocks) CK(launch_scan(p, db->num_sms, db->stream));
    CK(launch_final_min(p.red_min, p.red_idx, blocks, static_cast<ResultPair*>(result), db->stream));
    return IRIS_OK;
}

static int ensure_pairs(iris_db* db, size_t n_pairs) {
    if (db->pairs_cap >= n_pairs) return IRIS_OK;
    CK(cudaStreamSynchronize(db->stream));
    cudaFree(db->d_pairs);
    db->d_pairs = nullptr;
    db->pairs_cap = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&db->d_pairs), n_pairs * sizeof(ResultPair)));
    db->pairs_cap = n_pairs;
    return IRIS_OK;
}

extern "C" int iris_match_min_resident(iris_distance_engine* de, iris_masks_engine* me, iris_db* db, uint64_t row_begin,
                                       uint64_t row_end, uint64_t index_base, double* min_distance, uint64_t* min_index) {
    if (!de || !me || !db || !min_distance || !min_index) return fail(IRIS_ERR_INVALID, "NULL argument");
    DeviceGuard g(db->device);
    int rc = ensure_pairs(db, 64);
    if (rc) return rc;
    rc = iris_match_min_resident_async(de, me, db, row_begin, row_end, index_base, db->d_pairs);
    if (rc) return rc;
    ResultPair h;
    CK(cudaMemcpyAsync(&h, db->d_pairs, sizeof h, cudaMemcpyDeviceToHost, db->stream));
    rc = sync_checked(db, db->stream);
    if (rc) return rc;
    rc = check_error_flag(db);
    if (rc) return rc;
    *min_distance = h.min_distance;
    *min_index = h.min_index;
    return IRIS_OK;
}

// Batched search on one shard (BASELINE configs[3]/[4] per GPU): num_queries engine pairs against rows
// [row_begin,row_end) -- batched tensor-core distances and denominators slice by slice into scratch arrays owned by
// the shard, decode + min/argmin per query and slice on the device, running min over the slices -- and the
// num_queries result pairs written to `results` (device / peer / mapped host memory) in stream order.
static constexpr uint64_t kSearchSliceElems = 1ull << 25;        // rows x queries per slice: 2 x 2.08 GB of scratch
extern "C" int iris_search_batch_resident_async(iris_distance_engine* const* des, iris_masks_engine* const* mes,
                                                uint32_t num_queries, iris_db* db, uint64_t row_begin, uint64_t row_end,
                                                uint64_t index_base, void* results) {
    if (!des || !mes || !db || !results) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (num_queries == 0) return IRIS_OK;
    if (num_queries > (uint32_t)kMaxBatchQueries) return fail(IRIS_ERR_INVALID, "at most %d queries per search", kMaxBatchQueries);
    if (row_end < row_begin) return fail(IRIS_ERR_INVALID, "row_begin > row_end");
    DeviceGuard g(db->device);
    const uint64_t n = row_end - row_begin;
    const uint64_t slice = std::max<uint64_t>(2 * kTileRows, kSearchSliceElems / num_queries / (2 * kTileRows) * (2 * kTileRows));
    const uint64_t srows = std::min(n, slice);
    const uint32_t n_slices = (uint32_t)std::max<uint64_t>(1, (n + slice - 1) / slice);
    const size_t arr_bytes = ((size_t)num_queries * srows * kOutRowBytes + 255) / 256 * 256;
    const size_t scratch_bytes = ((size_t)num_queries * combine_scratch_bytes(srows) + 255) / 256 * 256;
    const size_t need = 2 * arr_bytes + scratch_bytes + 256;
    if (db->batch_bytes < need) {
        CK(cudaStreamSynchronize(db->stream));
        cudaFree(db->d_batch);
        db->d_batch = nullptr;
        db->batch_bytes = 0;
        CK(cudaMalloc(reinterpret_cast<void**>(&db->d_batch), need));
        db->batch_bytes = need;
    }
    int rc = ensure_pairs(db, (size_t)n_slices * num_queries);
    if (rc) return rc;
    uint16_t* d_bd = reinterpret_cast<uint16_t*>(db->d_batch);
    uint16_t* d_bn = reinterpret_cast<uint16_t*>(db->d_batch + arr_bytes);
    uint8_t* scratch = db->d_batch + 2 * arr_bytes;
    for (uint32_t sidx = 0; sidx < n_slices; ++sidx) {
        const uint64_t b = row_begin + (uint64_t)sidx * slice, e = std::min(row_end, b + slice);
        CombineParams p{};
        p.shares[0] = d_bd;
        p.parties = 1;
        p.denominators = d_bn;
        p.n = e - b;
        p.index_base = index_base + b;
        p.query_stride = (size_t)(e - b) * IRIS_ROTATIONS;
        if (e > b) {
            rc = iris_distances_batch_resident(des, num_queries, db, b, e, d_bd);
            if (rc) return rc;
            rc = iris_denominators_batch_resident(mes, num_queries, db, b, e, d_bn);
            if (rc) return rc;
        }
        CK(launch_combine_min(p, scratch, db->d_pairs + (size_t)sidx * num_queries, db->stream, num_queries));
    }
    CK(launch_merge_pairs(db->d_pairs, n_slices, num_queries, num_queries, static_cast<ResultPair*>(results), db->stream));
    return IRIS_OK;
}

extern "C" int iris_db_get_stream(iris_db* db, void** cuda_stream) {
    if (!db || !cuda_stream) return fail(IRIS_ERR_INVALID, "NULL argument");
    *cuda_stream = db->stream;
    return IRIS_OK;
}

extern "C" int iris_db_device(const iris_db* db, int* device) {
    if (!db || !device) return fail(IRIS_ERR_INVALID, "NULL argument");
    *device = db->device;
    return IRIS_OK;
}

// Watchdog state of a shard after the caller has synchronised its stream itself.
extern "C" int iris_db_check(iris_db* db) {
    if (!db) return fail(IRIS_ERR_INVALID, "db is NULL");
    return check_error_flag(db);
}

namespace iris {
void set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }
}

// ------------------------------------------------------------------------------------ arch-level batched dots
// The reference's criterion grid (src/arch/mod.rs:22-72) calls dot_u16 / dot_bool on every pair of `a` independent
// vectors and `b` database vectors.  Up to 31 vectors take the slots of the 31 rotations of one "engine"
// (iris_dotbatch.cu), so the same scan / batched-GEMM kernels compute out[i][j] = dot(a[j], b[i]).
static int vectors_to_device(int device, const void* a, size_t bytes, cudaStream_t s, uint8_t** d_out) {
    int rc = temp_alloc(device, reinterpret_cast<void**>(d_out), bytes, s);
    if (rc) return rc;
    if (is_device_pointer(a)) {
        CK(cudaMemcpyAsync(*d_out, a, bytes, cudaMemcpyDeviceToDevice, s));
    } else {
        ThreadStage* st = nullptr;
        rc = thread_stage(device, bytes, &st);
        if (rc) return rc;
        std::memcpy(st->p, a, bytes);
        CK(cudaMemcpyAsync(*d_out, st->p, bytes, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(st->done, s));
    }
    return IRIS_OK;
}

static int distance_engines_from_vectors(int device, const uint16_t* a, uint32_t n_vec, std::vector<iris_distance_engine*>& out) {
    cudaStream_t s = cudaStreamPerThread;
    uint8_t* d_a = nullptr;
    int rc = vectors_to_device(device, a, (size_t)n_vec * IRIS_BITS * 2, s, &d_a);
    if (rc) {
        temp_free(d_a, s);
        return rc;
    }
    auto body = [&]() -> int {
        for (uint32_t v0 = 0; v0 < n_vec; v0 += IRIS_ROTATIONS) {
            iris_distance_engine* e = nullptr;
            int r = new_distance_engine(device, s, &e);
            if (r) return r;
            out.push_back(e);
            int* h_flag = reinterpret_cast<int*>(e->slot.h + IRIS_BITS * 2);
            int* d_flag = nullptr;
            *h_flag = 1;
            CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_flag), h_flag, 0));
            CK(launch_prep_distance_vectors(reinterpret_cast<const uint16_t*>(d_a) + (size_t)v0 * IRIS_BITS,
                                            std::min<uint32_t>(IRIS_ROTATIONS, n_vec - v0), e->d_qd, d_flag, s));
            CK(cudaEventRecord(e->slot.ready, s));
            e->classified = false;
        }
        return IRIS_OK;
    };
    rc = body();
    temp_free(d_a, s);
    return rc;
}

static int masks_engines_from_vectors(int device, const uint64_t* a, uint32_t n_vec, std::vector<iris_masks_engine*>& out) {
    cudaStream_t s = cudaStreamPerThread;
    uint8_t* d_a = nullptr;
    int rc = vectors_to_device(device, a, (size_t)n_vec * IRIS_MASK_BYTES, s, &d_a);
    if (rc) {
        temp_free(d_a, s);
        return rc;
    }
    auto body = [&]() -> int {
        for (uint32_t v0 = 0; v0 < n_vec; v0 += IRIS_ROTATIONS) {
            iris_masks_engine* e = nullptr;
            int r = new_masks_engine(device, s, &e);
            if (r) return r;
            out.push_back(e);
            CK(launch_prep_mask_vectors(d_a + (size_t)v0 * IRIS_MASK_BYTES, std::min<uint32_t>(IRIS_ROTATIONS, n_vec - v0), e->d_qm,
                                        e->d_qm + kQmBytes, s));
            CK(cudaEventRecord(e->slot.ready, s));
        }
        return IRIS_OK;
    };
    rc = body();
    temp_free(d_a, s);
    return rc;
}

template <bool U16>
static int dot_batch_resident(const void* a, uint32_t n_a, iris_db* db, uint64_t row_begin, uint64_t row_end, uint16_t* out) {
    if (!a || !db) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (row_begin > row_end) return fail(IRIS_ERR_INVALID, "row_begin > row_end");
    if (n_a == 0 || row_begin == row_end) return IRIS_OK;
    if (!out) return fail(IRIS_ERR_INVALID, "out is NULL");
    if (n_a > 31u * kMaxBatchQueries) return fail(IRIS_ERR_INVALID, "at most %u vectors per call", 31u * kMaxBatchQueries);
    DeviceGuard g(db->device);
    const uint64_t n = row_end - row_begin;
    const uint32_t groups = (n_a + IRIS_ROTATIONS - 1) / IRIS_ROTATIONS;
    std::vector<iris_distance_engine*> des;
    std::vector<iris_masks_engine*> mes;
    int rc = U16 ? distance_engines_from_vectors(db->device, static_cast<const uint16_t*>(a), n_a, des)
                 : masks_engines_from_vectors(db->device, static_cast<const uint64_t*>(a), n_a, mes);
    uint16_t* d_tmp = nullptr;
    uint16_t* d_res = nullptr;
    const bool out_dev = is_device_pointer(out);
    auto body = [&]() -> int {
        if (rc) return rc;
        // [groups][n][31] intermediate, unless the caller's array already has that shape (31 vectors, device memory)
        const bool direct = n_a == IRIS_ROTATIONS && out_dev;
        if (!direct) {
            int r = temp_alloc(db->device, reinterpret_cast<void**>(&d_tmp), (size_t)groups * n * kOutRowBytes + 64, db->stream);
            if (r) return r;
        }
        uint16_t* grid = direct ? out : d_tmp;
        int r = U16 ? iris_distances_batch_resident(des.data(), groups, db, row_begin, row_end, grid)
                    : iris_denominators_batch_resident(mes.data(), groups, db, row_begin, row_end, grid);
        if (r) return r;
        if (direct) return IRIS_OK;
        uint16_t* dst = out;
        if (!out_dev) {
            r = temp_alloc(db->device, reinterpret_cast<void**>(&d_res), (size_t)n * n_a * 2 + 64, db->stream);
            if (r) return r;
            dst = d_res;
        }
        CK(launch_compact_columns(d_tmp, n, n_a, dst, db->stream));
        if (!out_dev) {
            CK(cudaMemcpyAsync(out, d_res, (size_t)n * n_a * 2, cudaMemcpyDeviceToHost, db->stream));
            int r2 = sync_checked(db, db->stream);
            if (r2) return r2;
            return check_error_flag(db);
        }
        return IRIS_OK;
    };
    rc = body();
    std::string keep = g_last_error;
    temp_free(d_tmp, db->stream);
    temp_free(d_res, db->stream);
    for (auto* e : des) iris_distance_engine_free(e);
    for (auto* e : mes) iris_masks_engine_free(e);
    g_last_error = keep;
    return rc;
}

extern "C" int iris_dot_u16_batch_resident(const uint16_t* a, uint32_t n_a, iris_db* db, uint64_t row_begin,
                                           uint64_t row_end, uint16_t* out) {
    return dot_batch_resident<true>(a, n_a, db, row_begin, row_end, out);
}
extern "C" int iris_dot_bool_batch_resident(const uint64_t* a, uint32_t n_a, iris_db* db, uint64_t row_begin,
                                            uint64_t row_end, uint16_t* out) {
    return dot_batch_resident<false>(a, n_a, db, row_begin, row_end, out);
}

// Host-array forms: `b` is uploaded into a temporary shard first (PCIe-bound; for throughput keep the vectors resident).
template <bool U16>
static int dot_batch_host(int device, const void* a, uint32_t n_a, const void* b, uint64_t n_b, uint16_t* out) {
    if (!a || !b || !out) return (n_a == 0 || n_b == 0) ? IRIS_OK : fail(IRIS_ERR_INVALID, "NULL argument");
    if (n_a == 0 || n_b == 0) return IRIS_OK;
    iris_db* db = nullptr;
    int rc = iris_db_create(device, n_b, U16 ? IRIS_DB_SHARES : IRIS_DB_MASKS, &db);
    if (rc) return rc;
    rc = U16 ? iris_db_append_shares(db, static_cast<const uint16_t*>(b), n_b) : iris_db_append_masks(db, static_cast<const uint64_t*>(b), n_b);
    if (!rc) rc = dot_batch_resident<U16>(a, n_a, db, 0, n_b, out);
    if (!rc) rc = iris_db_synchronize(db);
    std::string keep = g_last_error;
    iris_db_destroy(db);
    g_last_error = keep;
    return rc;
}
extern "C" int iris_dot_u16_batch(int device, const uint16_t* a, uint32_t n_a, const uint16_t* b, uint64_t n_b, uint16_t* out) {
    return dot_batch_host<true>(device, a, n_a, b, n_b, out);
}
extern "C" int iris_dot_bool_batch(int device, const uint64_t* a, uint32_t n_a, const uint64_t* b, uint64_t n_b, uint16_t* out) {
    return dot_batch_host<false>(device, a, n_a, b, n_b, out);
}

// ------------------------------------------------------------------------------------ per-pair arch entry points
template <typename T, typename F>
static int dot_pair(int device, const T* a, const T* b, size_t bytes, uint16_t* out, F launch) {
    if (!a || !b || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    int rc = require_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    uint8_t* d = nullptr;
    rc = temp_alloc(device, reinterpret_cast<void**>(&d), 2 * bytes + 16);
    if (rc) return rc;
    int result = IRIS_OK;
    auto body = [&]() -> int {
        CK(cudaMemcpyAsync(d, a, bytes, cudaMemcpyDefault, cudaStreamPerThread));
        CK(cudaMemcpyAsync(d + bytes, b, bytes, cudaMemcpyDefault, cudaStreamPerThread));
        CK(launch(reinterpret_cast<const T*>(d), reinterpret_cast<const T*>(d + bytes), reinterpret_cast<uint16_t*>(d + 2 * bytes), cudaStreamPerThread));
        CK(cudaMemcpyAsync(out, d + 2 * bytes, sizeof(uint16_t), cudaMemcpyDeviceToHost, cudaStreamPerThread));
        CK(cudaStreamSynchronize(cudaStreamPerThread));
        return IRIS_OK;
    };
    result = body();
    temp_free(d);
    return result;
}

extern "C" int iris_dot_u16(int device, const uint16_t* a, const uint16_t* b, uint16_t* out) {
    return dot_pair<uint16_t>(device, a, b, IRIS_BITS * sizeof(uint16_t), out, launch_dot_u16);
}
extern "C" int iris_dot_bool(int device, const uint64_t* a, const uint64_t* b, uint16_t* out) {
    return dot_pair<uint64_t>(device, a, b, IRIS_LIMBS * sizeof(uint64_t), out, launch_dot_bool);
}

// ------------------------------------------------------------------------------------ verification helpers
extern "C" int iris_check_distances_simt(iris_db* db, const uint16_t* query, uint64_t row_begin, uint64_t row_end,
                                         uint16_t* out) {
    if (!db || !query || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (!db->d_shares || row_end > db->n_shares || row_begin > row_end) return fail(IRIS_ERR_INVALID, "bad row range");
    DeviceGuard g(db->device);
    uint16_t* d_q = nullptr;
    uint16_t* d_o = nullptr;
    const uint64_t n = row_end - row_begin;
    const bool out_dev = is_device_pointer(out);
    CK(cudaMalloc(&d_q, IRIS_BITS * sizeof(uint16_t)));
    if (!out_dev) CK(cudaMalloc(&d_o, n * kOutRowBytes + 16));
    auto body = [&]() -> int {
        CK(cudaMemcpyAsync(d_q, query, IRIS_BITS * sizeof(uint16_t), cudaMemcpyDefault, db->stream));
        CK(launch_simt_distances(db->d_shares, d_q, row_begin, row_end, out_dev ? out : d_o, db->stream));
        if (!out_dev) CK(cudaMemcpyAsync(out, d_o, n * kOutRowBytes, cudaMemcpyDeviceToHost, db->stream));
        CK(cudaStreamSynchronize(db->stream));
        return IRIS_OK;
    };
    int rc = body();
    cudaFree(d_q);
    cudaFree(d_o);
    return rc;
}

extern "C" int iris_check_denominators_simt(iris_db* db, const uint64_t* query_mask, uint64_t row_begin,
                                            uint64_t row_end, uint16_t* out) {
    if (!db || !query_mask || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (!db->d_masks || row_end > db->n_masks || row_begin > row_end) return fail(IRIS_ERR_INVALID, "bad row range");
    DeviceGuard g(db->device);
    uint8_t* d_q = nullptr;
    uint16_t* d_o = nullptr;
    const uint64_t n = row_end - row_begin;
    const bool out_dev = is_device_pointer(out);
    CK(cudaMalloc(&d_q, IRIS_MASK_BYTES));
    if (!out_dev) CK(cudaMalloc(&d_o, n * kOutRowBytes + 16));
    auto body = [&]() -> int {
        CK(cudaMemcpyAsync(d_q, query_mask, IRIS_MASK_BYTES, cudaMemcpyDefault, db->stream));
        CK(launch_simt_denominators(db->d_masks, d_q, row_begin, row_end, out_dev ? out : d_o, db->stream));
        if (!out_dev) CK(cudaMemcpyAsync(out, d_o, n * kOutRowBytes, cudaMemcpyDeviceToHost, db->stream));
        CK(cudaStreamSynchronize(db->stream));
        return IRIS_OK;
    };
    int rc = body();
    cudaFree(d_q);
    cudaFree(d_o);
    return rc;
}

extern "C" int iris_debug_raw_accumulators(iris_distance_engine* de, iris_masks_engine* me, iris_db* db,
                                           uint64_t row_begin, uint64_t row_end, int32_t* raw_out) {
    if (!db || !raw_out || (!de && !me)) return fail(IRIS_ERR_INVALID, "NULL argument");
    if (row_begin % kTileRows) return fail(IRIS_ERR_INVALID, "row_begin must be a multiple of 128");
    if (row_end <= row_begin) return fail(IRIS_ERR_INVALID, "empty range");
    DeviceGuard g(db->device);
    const uint64_t tiles = (row_end - row_begin + kTileRows - 1) / kTileRows;
    const uint64_t n = row_end - row_begin;
    int32_t* d_raw = nullptr;
    uint16_t* d_o = nullptr;
    CK(cudaMalloc(&d_raw, tiles * kTileRows * 128 * sizeof(int32_t)));
    CK(cudaMalloc(&d_o, 2 * (n * kOutRowBytes + 64)));
    auto body = [&]() -> int {
        CK(cudaMemsetAsync(d_raw, 0xEE, tiles * kTileRows * 128 * sizeof(int32_t), db->stream));
        uint16_t* d_den = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(d_o) + (n * kOutRowBytes + 63) / 64 * 64);
        int rc = scan_core(db, de, me, row_begin, row_end, d_o, d_den, d_raw);
        if (rc) return rc;
        CK(cudaMemcpyAsync(raw_out, d_raw, tiles * kTileRows * 128 * sizeof(int32_t), cudaMemcpyDeviceToHost, db->stream));
        CK(cudaStreamSynchronize(db->stream));
        return check_error_flag(db);
    };
    int rc = body();
    cudaFree(d_raw);
    cudaFree(d_o);
    return rc;
}
