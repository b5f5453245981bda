// One database row-sharded over several GPUs of a box, behind one handle (include/iris_b200.h, "cluster").
//
// The reference mmaps the whole share / mask file and walks it in 20 000-row chunks on one host (src/main.rs:386-400,
// 425-431, 510-516); rows never interact (src/lib.rs:44-51), so here every GPU owns a contiguous block of rows in its
// own HBM and scans it with the same kernels a single shard uses.  One host thread per GPU submits that GPU's work, so
// the eight submissions (and the eight PCIe links, for host outputs) run side by side.  What crosses GPUs is small:
//   search  the (min, argmin) pairs -- the last kernel of each shard's reduction stores them straight into the root
//           GPU's memory (NVLink peer stores), the root merges them after waiting for one event per shard;
//   match   nothing is exchanged afterwards: each scan kernel's epilogue stores its [rows][31] block at its row offset
//           of the caller's single array, which may live on another GPU (peer stores) or in host memory;
//   several processes (torchrun: one per GPU) all-gather their merged pairs over NCCL and merge again.
// Built on the public C ABI plus the CUDA runtime; NCCL is looked up at run time (no link-time dependency).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; the functions are resolved with dlsym
#include <sys/stat.h>

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/iris_b200.h"
#include "iris_kernels.cuh"

namespace iris {
void set_last_error(const char* msg);
}
using namespace iris;

namespace {

int cfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    set_last_error(buf);
    return code;
}

#define CCK(expr)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            cudaGetLastError();                                                                         \
            return cfail(e_ == cudaErrorMemoryAllocation ? IRIS_ERR_NOMEM : IRIS_ERR_CUDA, "%s failed: %s", #expr, \
                         cudaGetErrorString(e_));                                                       \
        }                                                                                               \
    } while (0)

struct SetDevice {
    int prev = -1;
    explicit SetDevice(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        cudaSetDevice(dev);
    }
    ~SetDevice() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- one submission thread per GPU
class Worker {
public:
    explicit Worker(int device) : device_(device), thread_([this] { loop(); }) {}
    ~Worker() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        thread_.join();
    }
    void post(std::function<int()> job) {
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = std::move(job);
            has_job_ = true;
            done_ = false;
        }
        cv_.notify_all();
    }
    int wait(std::string* err) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return done_; });
        if (rc_ && err) *err = err_;
        return rc_;
    }

private:
    void loop() {
        cudaSetDevice(device_);
        for (;;) {
            std::function<int()> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return has_job_ || stop_; });
                if (stop_) return;
                job = std::move(job_);
                has_job_ = false;
            }
            const int rc = job();
            std::string err = rc ? iris_last_error() : "";
            {
                std::lock_guard<std::mutex> g(mu_);
                rc_ = rc;
                err_ = std::move(err);
                done_ = true;
            }
            cv_.notify_all();
        }
    }
    int device_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::function<int()> job_;
    bool has_job_ = false, done_ = true, stop_ = false;
    int rc_ = 0;
    std::string err_;
    std::thread thread_;   // last: the members above must exist before the thread starts
};

struct Shard {
    int device = 0;
    iris_db* db = nullptr;
    uint64_t begin = 0, end = 0;        // block of cluster rows held at local rows [0, end - begin)
    cudaEvent_t done = nullptr;
    Worker* worker = nullptr;
};

// ---- NCCL, resolved at run time: the copy the process already has (torch's), else the file IRIS_NCCL_LIB names, else
// the system's libnccl.so.2.  A process may hold only one library with that soname, so a host that loads another copy
// LATER (a Python process importing torch after its first join) should name that copy in IRIS_NCCL_LIB up front.
struct Nccl {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};

Nccl* nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) {
            const char* path = getenv("IRIS_NCCL_LIB");
            if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            const char* e = dlerror();
            n.error = std::string("libnccl.so.2 cannot be loaded: ") + (e ? e : "?");
            return;
        }
        n.handle = h;
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(h, "ncclAllGather"));
        n.Broadcast = reinterpret_cast<decltype(n.Broadcast)>(dlsym(h, "ncclBroadcast"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(dlsym(h, "ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        if (!n.GetUniqueId || !n.CommInitRank || !n.AllGather || !n.CommDestroy || !n.GetErrorString || !n.Broadcast ||
            !n.GroupStart || !n.GroupEnd) {
            n.error = "libnccl.so.2 lacks a required symbol";
            n.handle = nullptr;
        }
    });
    return &n;
}

constexpr uint32_t kMaxSearchQueries = 64;        // per pass (the batched kernels' limit); larger batches loop

}  // namespace

struct iris_cluster {
    std::vector<Shard> shards;
    std::vector<Worker*> workers;
    uint32_t flags = 0;
    uint64_t shard_capacity = 0;
    uint64_t n_shares = 0, n_masks = 0;
    uint64_t index_base = 0;
    // search: gather[shard][64] pairs written by the shards, merged[64] by the root, all[world][64] after the all-gather
    ResultPair* gather = nullptr;
    bool gather_on_host = false;
    ResultPair* d_merged = nullptr;
    ResultPair* d_all = nullptr;
    ResultPair* d_final = nullptr;
    ResultPair* h_result = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    uint64_t* d_layout = nullptr;        // [world][2] = {first global row, rows} of every process, all-gathered per gather
    uint64_t* h_layout = nullptr;
    std::mutex mu;                       // one operation at a time
    std::vector<std::pair<int, int>> peers;   // (from, to) pairs with peer access enabled
    std::vector<std::pair<int, int>> no_peers;
};

namespace {

// Can kernels running on `from` store into memory of `to`?  Enables peer access on first use.
bool peer_ok(iris_cluster* c, int from, int to) {
    if (from == to) return true;
    const auto key = std::make_pair(from, to);
    if (std::find(c->peers.begin(), c->peers.end(), key) != c->peers.end()) return true;
    if (std::find(c->no_peers.begin(), c->no_peers.end(), key) != c->no_peers.end()) return false;
    int can = 0;
    bool ok = cudaDeviceCanAccessPeer(&can, from, to) == cudaSuccess && can;
    if (ok) {
        SetDevice g(from);
        const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
        ok = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
    }
    cudaGetLastError();
    (ok ? c->peers : c->no_peers).push_back(key);
    return ok;
}

// Runs fn(shard index) on every shard's thread; first failure wins (its message becomes the caller's last error).
int run_all(iris_cluster* c, const std::function<int(uint32_t)>& fn) {
    const uint32_t n = (uint32_t)c->shards.size();
    if (n == 1) return fn(0);                      // one shard: the calling thread submits (no hand-over latency)
    for (uint32_t i = 0; i < n; ++i) c->shards[i].worker->post([&fn, i] { return fn(i); });
    int rc = IRIS_OK;
    std::string err;
    for (uint32_t i = 0; i < n; ++i) {
        std::string e;
        const int r = c->shards[i].worker->wait(&e);
        if (r && !rc) {
            rc = r;
            err = e;
        }
    }
    if (rc) set_last_error(err.c_str());
    return rc;
}

int partition(uint64_t n_total, uint32_t n_shards, uint32_t shard, uint64_t* b, uint64_t* e) {
    if (n_shards == 0 || shard >= n_shards) return cfail(IRIS_ERR_INVALID, "shard %u of %u", shard, n_shards);
    const uint64_t base = n_total / n_shards, extra = n_total % n_shards;
    *b = shard * base + std::min<uint64_t>(shard, extra);
    *e = *b + base + (shard < extra ? 1 : 0);
    return IRIS_OK;
}

int set_blocks(iris_cluster* c, uint64_t n) {
    const uint32_t ns = (uint32_t)c->shards.size();
    for (uint32_t i = 0; i < ns; ++i) {
        int rc = partition(n, ns, i, &c->shards[i].begin, &c->shards[i].end);
        if (rc) return rc;
        if (c->shards[i].end - c->shards[i].begin > c->shard_capacity)
            return cfail(IRIS_ERR_INVALID, "%llu rows exceed the cluster's capacity", (unsigned long long)n);
    }
    return IRIS_OK;
}

int file_rows(const char* path, size_t row_bytes, uint64_t* rows) {
    struct stat st;
    if (stat(path, &st) != 0) return cfail(IRIS_ERR_INVALID, "cannot open %s", path);
    // reference: try_cast_slice fails -> "Share file invalid" / "Masks file invalid" (src/main.rs:391-392, 460-461)
    if ((uint64_t)st.st_size % row_bytes)
        return cfail(IRIS_ERR_INVALID, "%s: size %llu is not a multiple of %zu", path, (unsigned long long)st.st_size, row_bytes);
    *rows = (uint64_t)st.st_size / row_bytes;
    return IRIS_OK;
}

}  // namespace

extern "C" int iris_cluster_partition(uint64_t n_total, uint32_t n_shards, uint32_t shard, uint64_t* row_begin,
                                      uint64_t* row_end) {
    if (!row_begin || !row_end) return cfail(IRIS_ERR_INVALID, "NULL argument");
    return partition(n_total, n_shards, shard, row_begin, row_end);
}

extern "C" int iris_cluster_destroy(iris_cluster* c) {
    if (!c) return IRIS_OK;
    for (auto& s : c->shards) {
        if (s.db) {
            iris_db_synchronize(s.db);
            iris_db_destroy(s.db);
        }
        if (s.done) {
            SetDevice g(s.device);
            cudaEventDestroy(s.done);
        }
    }
    for (Worker* w : c->workers) delete w;
    if (!c->shards.empty()) {
        SetDevice g(c->shards[0].device);
        if (c->comm && nccl()->handle) nccl()->CommDestroy(c->comm);
        if (c->gather) {
            if (c->gather_on_host) cudaFreeHost(c->gather);
            else cudaFree(c->gather);
        }
        cudaFree(c->d_merged);
        cudaFree(c->d_all);
        cudaFree(c->d_layout);
        if (c->h_layout) cudaFreeHost(c->h_layout);
        if (c->h_result) cudaFreeHost(c->h_result);
    }
    cudaGetLastError();
    delete c;
    return IRIS_OK;
}

extern "C" int iris_cluster_create(const int* devices, uint32_t n_devices, uint64_t capacity_rows, uint32_t flags,
                                   iris_cluster** out) {
    if (!out) return cfail(IRIS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices == 0) return cfail(IRIS_ERR_INVALID, "a cluster needs at least one device");
    if (n_devices > 64) return cfail(IRIS_ERR_INVALID, "at most 64 shards");
    if (capacity_rows == 0) return cfail(IRIS_ERR_INVALID, "capacity_rows must be > 0");
    iris_cluster* c = new (std::nothrow) iris_cluster();
    if (!c) return cfail(IRIS_ERR_NOMEM, "host allocation failed");
    c->flags = flags;
    c->shard_capacity = (capacity_rows + n_devices - 1) / n_devices;
    c->shards.resize(n_devices);
    for (uint32_t i = 0; i < n_devices; ++i) c->shards[i].device = devices[i];
    auto body = [&]() -> int {
        for (uint32_t i = 0; i < n_devices; ++i) {
            Worker* w = new (std::nothrow) Worker(devices[i]);
            if (!w) return cfail(IRIS_ERR_NOMEM, "host allocation failed");
            c->workers.push_back(w);
            c->shards[i].worker = w;
        }
        // every GPU allocates and clears its block at the same time
        int rc = run_all(c, [&](uint32_t i) -> int {
            Shard& s = c->shards[i];
            int r = iris_db_create(s.device, c->shard_capacity, flags, &s.db);
            if (r) return r;
            SetDevice g(s.device);
            CCK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            return IRIS_OK;
        });
        if (rc) return rc;
        const int root = c->shards[0].device;
        SetDevice g(root);
        bool all_peer = true;
        for (auto& s : c->shards) all_peer &= peer_ok(c, s.device, root);
        const size_t gbytes = (size_t)n_devices * kMaxSearchQueries * sizeof(ResultPair);
        if (all_peer) {
            CCK(cudaMalloc(reinterpret_cast<void**>(&c->gather), gbytes));
        } else {   // GPUs that cannot reach each other: the pairs go through mapped host memory instead
            CCK(cudaHostAlloc(reinterpret_cast<void**>(&c->gather), gbytes, cudaHostAllocPortable | cudaHostAllocMapped));
            c->gather_on_host = true;
        }
        CCK(cudaMalloc(reinterpret_cast<void**>(&c->d_merged), 2 * kMaxSearchQueries * sizeof(ResultPair)));
        c->d_final = c->d_merged + kMaxSearchQueries;
        CCK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_result), kMaxSearchQueries * sizeof(ResultPair), cudaHostAllocPortable));
        return IRIS_OK;
    };
    const int rc = body();
    if (rc) {
        std::string keep = iris_last_error();
        iris_cluster_destroy(c);
        set_last_error(keep.c_str());
        return rc;
    }
    *out = c;
    return IRIS_OK;
}

extern "C" int iris_cluster_shard(iris_cluster* c, uint32_t shard, iris_db** db, int* device, uint64_t* row_begin,
                                  uint64_t* row_end) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    if (shard >= c->shards.size()) return cfail(IRIS_ERR_INVALID, "shard %u of %zu", shard, c->shards.size());
    const Shard& s = c->shards[shard];
    if (db) *db = s.db;
    if (device) *device = s.device;
    if (row_begin) *row_begin = s.begin;
    if (row_end) *row_end = s.end;
    return IRIS_OK;
}

extern "C" int iris_cluster_len(const iris_cluster* c, uint32_t* n_shards, uint64_t* n_shares, uint64_t* n_masks) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    if (n_shards) *n_shards = (uint32_t)c->shards.size();
    if (n_shares) *n_shares = c->n_shares;
    if (n_masks) *n_masks = c->n_masks;
    return IRIS_OK;
}

extern "C" int iris_cluster_set_index_base(iris_cluster* c, uint64_t index_base) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    c->index_base = index_base;
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------ populate
extern "C" int iris_cluster_generate(iris_cluster* c, uint64_t seed, uint32_t party, uint32_t n_parties, uint64_t first_row_id,
                                     uint64_t n) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = set_blocks(c, n);
    if (rc) return rc;
    rc = run_all(c, [&](uint32_t i) -> int {
        Shard& s = c->shards[i];
        int r = iris_db_clear(s.db);
        if (r) return r;
        const uint64_t cnt = s.end - s.begin;
        if (cnt == 0) return IRIS_OK;
        return n_parties == 0 ? iris_db_generate(s.db, seed, first_row_id + s.begin, cnt)
                              : iris_db_generate_shares(s.db, seed, party, n_parties, first_row_id + s.begin, cnt);
    });
    c->n_shares = !rc && (c->flags & IRIS_DB_SHARES) ? n : 0;
    c->n_masks = !rc && (c->flags & IRIS_DB_MASKS) ? n : 0;
    return rc;
}

extern "C" int iris_cluster_load_files(iris_cluster* c, const char* shares_path, const char* masks_path) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    if (!shares_path && !masks_path) return cfail(IRIS_ERR_INVALID, "no file given");
    if (shares_path && !(c->flags & IRIS_DB_SHARES)) return cfail(IRIS_ERR_STATE, "cluster was created without IRIS_DB_SHARES");
    if (masks_path && !(c->flags & IRIS_DB_MASKS)) return cfail(IRIS_ERR_STATE, "cluster was created without IRIS_DB_MASKS");
    std::lock_guard<std::mutex> lk(c->mu);
    uint64_t ns = 0, nm = 0;
    int rc = IRIS_OK;
    if (shares_path && (rc = file_rows(shares_path, IRIS_BITS * sizeof(uint16_t), &ns))) return rc;
    if (masks_path && (rc = file_rows(masks_path, IRIS_LIMBS * sizeof(uint64_t), &nm))) return rc;
    if (shares_path && masks_path && ns != nm)
        return cfail(IRIS_ERR_INVALID, "%llu share rows but %llu mask rows", (unsigned long long)ns, (unsigned long long)nm);
    const uint64_t n = shares_path ? ns : nm;
    rc = set_blocks(c, n);
    if (rc) return rc;
    rc = run_all(c, [&](uint32_t i) -> int {
        Shard& s = c->shards[i];
        int r = iris_db_clear(s.db);
        if (r) return r;
        const uint64_t cnt = s.end - s.begin;
        if (cnt == 0) return IRIS_OK;
        if (shares_path && (r = iris_db_load_shares_file(s.db, shares_path, s.begin, cnt))) return r;
        if (masks_path && (r = iris_db_load_masks_file(s.db, masks_path, s.begin, cnt))) return r;
        return IRIS_OK;
    });
    c->n_shares = !rc && shares_path ? n : 0;
    c->n_masks = !rc && masks_path ? n : 0;
    return rc;
}

extern "C" int iris_cluster_load_rows(iris_cluster* c, const uint16_t* shares, const uint64_t* masks, uint64_t n) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    if (!shares && !masks && n) return cfail(IRIS_ERR_INVALID, "no rows given");
    if (shares && !(c->flags & IRIS_DB_SHARES)) return cfail(IRIS_ERR_STATE, "cluster was created without IRIS_DB_SHARES");
    if (masks && !(c->flags & IRIS_DB_MASKS)) return cfail(IRIS_ERR_STATE, "cluster was created without IRIS_DB_MASKS");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = set_blocks(c, n);
    if (rc) return rc;
    rc = run_all(c, [&](uint32_t i) -> int {
        Shard& s = c->shards[i];
        int r = iris_db_clear(s.db);
        if (r) return r;
        const uint64_t cnt = s.end - s.begin;
        if (cnt == 0) return IRIS_OK;
        if (shares && (r = iris_db_append_shares(s.db, shares + s.begin * IRIS_BITS, cnt))) return r;
        if (masks && (r = iris_db_append_masks(s.db, masks + s.begin * IRIS_LIMBS, cnt))) return r;
        return IRIS_OK;
    });
    c->n_shares = !rc && shares ? n : 0;
    c->n_masks = !rc && masks ? n : 0;
    return rc;
}

// ------------------------------------------------------------------------------------ match: full result vectors
namespace {

// Where shard `s` stores the block of rows [s.begin, s.end) of the caller's array `out` ([n][31] u16).  Host memory and
// reachable device memory are written in place; for an unreachable GPU the block is staged locally and copied.
struct BlockOut {
    uint16_t* direct = nullptr;   // pointer the scan writes to
    uint16_t* staged = nullptr;   // local staging to copy from afterwards (rare: no peer access)
    uint16_t* final_dst = nullptr;
    int dst_device = -1;
};

int block_out(iris_cluster* c, const Shard& s, uint16_t* out, BlockOut* b) {
    *b = BlockOut();
    if (!out) return IRIS_OK;
    uint16_t* dst = out + s.begin * IRIS_ROTATIONS;
    cudaPointerAttributes attr;
    const bool dev = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (!dev || peer_ok(c, s.device, attr.device)) {
        b->direct = dst;
        return IRIS_OK;
    }
    void* p = nullptr;
    int rc = iris_device_alloc(s.device, (s.end - s.begin) * IRIS_ROTATIONS * sizeof(uint16_t) + 64, &p);
    if (rc) return rc;
    b->direct = b->staged = static_cast<uint16_t*>(p);
    b->final_dst = dst;
    b->dst_device = attr.device;
    return IRIS_OK;
}

int block_finish(const Shard& s, BlockOut& b) {
    if (!b.staged) return IRIS_OK;
    void* stream = nullptr;
    int rc = iris_db_get_stream(s.db, &stream);
    if (!rc) {
        SetDevice g(s.device);
        const cudaError_t e = cudaMemcpyPeerAsync(b.final_dst, b.dst_device, b.staged, s.device,
                                                  (s.end - s.begin) * IRIS_ROTATIONS * sizeof(uint16_t), static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) rc = cfail(IRIS_ERR_CUDA, "cudaMemcpyPeerAsync failed: %s", cudaGetErrorString(e));
    }
    if (!rc) rc = iris_db_synchronize(s.db);
    iris_device_free(s.device, b.staged);
    b.staged = nullptr;
    return rc;
}

struct ShardProgress {            // forwards a shard's progress reports as ranges of cluster rows
    iris_progress_fn fn;
    void* user;
    uint64_t offset;
};
void shard_progress(void* u, uint64_t b, uint64_t e) {
    const ShardProgress* sp = static_cast<const ShardProgress*>(u);
    sp->fn(sp->user, sp->offset + b, sp->offset + e);
}

int cluster_match(iris_cluster* c, const uint16_t* query, const uint64_t* pattern, const uint64_t* mask, bool want_den,
                  uint16_t* distances_out, uint16_t* denominators_out, iris_progress_fn progress = nullptr, void* user = nullptr) {
    const bool want_dist = query || pattern;
    if (!want_dist && !want_den) return cfail(IRIS_ERR_INVALID, "no query given");
    if (want_dist && !distances_out && c->n_shares) return cfail(IRIS_ERR_INVALID, "distances output is NULL");
    if (want_den && !denominators_out && c->n_masks) return cfail(IRIS_ERR_INVALID, "denominators output is NULL");
    if (want_dist && !(c->flags & IRIS_DB_SHARES)) return cfail(IRIS_ERR_STATE, "cluster holds no shares");
    if (want_den && !(c->flags & IRIS_DB_MASKS)) return cfail(IRIS_ERR_STATE, "cluster holds no masks");
    if (want_dist && want_den && c->n_shares != c->n_masks) return cfail(IRIS_ERR_STATE, "shares and masks lengths differ");
    std::lock_guard<std::mutex> lk(c->mu);
    {   // peer access is enabled from this thread before the workers need it (the bookkeeping is not thread safe)
        for (auto& s : c->shards)
            for (uint16_t* o : {distances_out, denominators_out}) {
                cudaPointerAttributes attr;
                if (o && cudaPointerGetAttributes(&attr, o) == cudaSuccess && attr.type == cudaMemoryTypeDevice)
                    peer_ok(c, s.device, attr.device);
                cudaGetLastError();
            }
    }
    return run_all(c, [&](uint32_t i) -> int {
        Shard& s = c->shards[i];
        const uint64_t cnt = s.end - s.begin;
        if (cnt == 0) return IRIS_OK;
        iris_distance_engine* de = nullptr;
        iris_masks_engine* me = nullptr;
        BlockOut bd, bn;
        auto body = [&]() -> int {
            int r = IRIS_OK;
            if (query) r = iris_distance_engine_new(s.device, query, &de);
            else if (pattern) r = iris_distance_engine_new_from_template(s.device, pattern, mask, &de);
            if (r) return r;
            if (want_den && (r = iris_masks_engine_new(s.device, mask, &me))) return r;
            if (de && (r = block_out(c, s, distances_out, &bd))) return r;
            if (me && (r = block_out(c, s, denominators_out, &bn))) return r;
            if (progress) {
                ShardProgress sp{progress, user, s.begin};
                r = iris_match_resident_streamed(de, me, s.db, 0, cnt, bd.direct, bn.direct, shard_progress, &sp);
            } else {
                r = iris_match_resident(de, me, s.db, 0, cnt, bd.direct, bn.direct);
            }
            if (r) return r;
            if ((r = block_finish(s, bd))) return r;
            if ((r = block_finish(s, bn))) return r;
            return iris_db_synchronize(s.db);
        };
        const int rc = body();
        std::string keep = rc ? iris_last_error() : "";
        if (bd.staged) iris_device_free(s.device, bd.staged);
        if (bn.staged) iris_device_free(s.device, bn.staged);
        iris_distance_engine_free(de);
        iris_masks_engine_free(me);
        if (rc) set_last_error(keep.c_str());
        return rc;
    });
}

}  // namespace

extern "C" int iris_cluster_match(iris_cluster* c, const uint16_t* query, const uint64_t* query_mask, uint16_t* distances_out,
                                  uint16_t* denominators_out) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    return cluster_match(c, query, nullptr, query_mask, query_mask != nullptr, distances_out, denominators_out);
}

extern "C" int iris_cluster_match_template(iris_cluster* c, const uint64_t* pattern, const uint64_t* mask, uint16_t* distances_out,
                                           uint16_t* denominators_out) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    if (!pattern || !mask) return cfail(IRIS_ERR_INVALID, "NULL template");
    return cluster_match(c, nullptr, pattern, mask, denominators_out != nullptr, distances_out, denominators_out);
}

extern "C" int iris_cluster_match_template_streamed(iris_cluster* c, const uint64_t* pattern, const uint64_t* mask,
                                                    uint16_t* distances_out, uint16_t* denominators_out,
                                                    iris_progress_fn progress, void* user) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    if (!pattern || !mask) return cfail(IRIS_ERR_INVALID, "NULL template");
    return cluster_match(c, nullptr, pattern, mask, denominators_out != nullptr, distances_out, denominators_out, progress, user);
}

// ------------------------------------------------------------------------------------ search: (min, argmin) per query
extern "C" int iris_cluster_search(iris_cluster* c, const uint64_t* templates, uint32_t num_queries, double* min_distance,
                                   uint64_t* min_index) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    if (num_queries == 0) return IRIS_OK;
    if (!templates || !min_distance || !min_index) return cfail(IRIS_ERR_INVALID, "NULL argument");
    if ((c->flags & (IRIS_DB_SHARES | IRIS_DB_MASKS)) != (IRIS_DB_SHARES | IRIS_DB_MASKS))
        return cfail(IRIS_ERR_STATE, "a search needs shares and masks");
    if (c->n_shares != c->n_masks) return cfail(IRIS_ERR_STATE, "shares and masks lengths differ");
    std::lock_guard<std::mutex> lk(c->mu);
    const uint32_t ns = (uint32_t)c->shards.size();
    const int root = c->shards[0].device;
    for (uint32_t q0 = 0; q0 < num_queries; q0 += kMaxSearchQueries) {
        const uint32_t nq = std::min(kMaxSearchQueries, num_queries - q0);
        const uint64_t* tq = templates + (size_t)q0 * 2 * IRIS_LIMBS;
        // every shard: engines from the wire Templates, scan + reduce, pairs stored into the root's gather rows
        int rc = run_all(c, [&](uint32_t i) -> int {
            Shard& s = c->shards[i];
            std::vector<iris_distance_engine*> des(nq, nullptr);
            std::vector<iris_masks_engine*> mes(nq, nullptr);
            ResultPair* dst = c->gather + (size_t)i * kMaxSearchQueries;
            auto body = [&]() -> int {
                int r = iris_engines_new_from_templates(s.device, tq, nq, des.data(), mes.data());
                if (r) return r;
                const uint64_t cnt = s.end - s.begin;
                if (nq == 1) r = iris_match_min_resident_async(des[0], mes[0], s.db, 0, cnt, c->index_base + s.begin, dst);
                else r = iris_search_batch_resident_async(des.data(), mes.data(), nq, s.db, 0, cnt, c->index_base + s.begin, dst);
                if (r) return r;
                void* stream = nullptr;
                if ((r = iris_db_get_stream(s.db, &stream))) return r;
                SetDevice g(s.device);
                CCK(cudaEventRecord(s.done, static_cast<cudaStream_t>(stream)));
                return IRIS_OK;
            };
            const int r = body();
            std::string keep = r ? iris_last_error() : "";
            for (auto* e : des) iris_distance_engine_free(e);     // stream-ordered release: never blocks
            for (auto* e : mes) iris_masks_engine_free(e);
            if (r) set_last_error(keep.c_str());
            return r;
        });
        // root: wait for one event per shard, merge, (all-gather + merge again), 16 bytes per query to the host
        void* rs = nullptr;
        int rc2 = iris_db_get_stream(c->shards[0].db, &rs);
        if (rc2) return rc2;
        cudaStream_t root_stream = static_cast<cudaStream_t>(rs);
        SetDevice g(root);
        if (rc) {
            for (auto& s : c->shards) iris_db_synchronize(s.db);
            return rc;
        }
        for (auto& s : c->shards) CCK(cudaStreamWaitEvent(root_stream, s.done, 0));
        CCK(launch_merge_pairs(c->gather, ns, kMaxSearchQueries, nq, c->d_merged, root_stream));
        const ResultPair* final_pairs = c->d_merged;
        if (c->world > 1) {
            Nccl* n = nccl();
            const ncclResult_t r = n->AllGather(c->d_merged, c->d_all, (size_t)nq * 2, ncclUint64, c->comm, root_stream);
            if (r != ncclSuccess) return cfail(IRIS_ERR_CUDA, "ncclAllGather failed: %s", n->GetErrorString(r));
            CCK(launch_merge_pairs(c->d_all, (uint32_t)c->world, nq, nq, c->d_final, root_stream));
            final_pairs = c->d_final;
        }
        CCK(cudaMemcpyAsync(c->h_result, final_pairs, nq * sizeof(ResultPair), cudaMemcpyDeviceToHost, root_stream));
        const cudaError_t e = cudaStreamSynchronize(root_stream);
        for (auto& s : c->shards) {
            const int r = iris_db_check(s.db);
            if (r) return r;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return cfail(IRIS_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
        }
        for (uint32_t q = 0; q < nq; ++q) {
            min_distance[q0 + q] = c->h_result[q].min_distance;
            min_index[q0 + q] = c->h_result[q].min_index;
        }
    }
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------ several processes: NCCL
extern "C" int iris_comm_unique_id(void* id_out) {
    if (!id_out) return cfail(IRIS_ERR_INVALID, "id_out is NULL");
    static_assert(sizeof(ncclUniqueId) == IRIS_UNIQUE_ID_BYTES, "ncclUniqueId size");
    Nccl* n = nccl();
    if (!n->handle) return cfail(IRIS_ERR_STATE, "%s", n->error.c_str());
    ncclUniqueId id;
    const ncclResult_t r = n->GetUniqueId(&id);
    if (r != ncclSuccess) return cfail(IRIS_ERR_CUDA, "ncclGetUniqueId failed: %s", n->GetErrorString(r));
    std::memcpy(id_out, &id, sizeof id);
    return IRIS_OK;
}

extern "C" int iris_cluster_join(iris_cluster* c, const void* unique_id, int rank, int world_size) {
    if (!c || !unique_id) return cfail(IRIS_ERR_INVALID, "NULL argument");
    if (world_size < 1 || rank < 0 || rank >= world_size) return cfail(IRIS_ERR_INVALID, "rank %d of %d", rank, world_size);
    if (c->comm) return cfail(IRIS_ERR_STATE, "cluster has already joined a communicator");
    Nccl* n = nccl();
    if (!n->handle) return cfail(IRIS_ERR_STATE, "%s", n->error.c_str());
    std::lock_guard<std::mutex> lk(c->mu);
    SetDevice g(c->shards[0].device);
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof id);
    const ncclResult_t r = n->CommInitRank(&c->comm, world_size, id, rank);
    if (r != ncclSuccess) {
        c->comm = nullptr;
        return cfail(IRIS_ERR_CUDA, "ncclCommInitRank failed: %s", n->GetErrorString(r));
    }
    CCK(cudaMalloc(reinterpret_cast<void**>(&c->d_all), (size_t)world_size * kMaxSearchQueries * sizeof(ResultPair)));
    CCK(cudaMalloc(reinterpret_cast<void**>(&c->d_layout), (size_t)(world_size + 1) * 2 * sizeof(uint64_t)));
    CCK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_layout), (size_t)(world_size + 1) * 2 * sizeof(uint64_t), cudaHostAllocPortable));
    c->rank = rank;
    c->world = world_size;
    return IRIS_OK;
}

// Multi-process clusters: the full per-row result vectors of one query on EVERY process.  Each process scans its rows
// into its slot of the caller's device array(s) ([rows of all processes][31] u16, slot = index_base .. index_base +
// local rows) and the blocks are exchanged over NVLink with one grouped ncclBroadcast per process (blocks may differ
// in size).  Collective: every process calls it with the same kind of query.  This is the "compute, then collective"
// baseline; inside ONE process the scan epilogues store straight into the destination GPU (iris_cluster_match), which
// needs no second step.
extern "C" int iris_cluster_match_allgather(iris_cluster* c, const uint16_t* query, const uint64_t* query_mask,
                                            uint16_t* distances_out, uint16_t* denominators_out) {
    if (!c) return cfail(IRIS_ERR_INVALID, "cluster is NULL");
    if (c->world <= 1 || !c->comm) return cfail(IRIS_ERR_STATE, "the cluster has not joined a multi-process communicator");
    if ((query && !distances_out) || (query_mask && !denominators_out) || (!query && !query_mask))
        return cfail(IRIS_ERR_INVALID, "query / output mismatch");
    for (uint16_t* o : {distances_out, denominators_out}) {
        cudaPointerAttributes attr;
        if (o && (cudaPointerGetAttributes(&attr, o) != cudaSuccess || attr.type != cudaMemoryTypeDevice)) {
            cudaGetLastError();
            return cfail(IRIS_ERR_INVALID, "the gathered arrays must be device memory");
        }
    }
    Nccl* n = nccl();
    const uint64_t local = std::max(c->n_shares, c->n_masks);
    void* rs = nullptr;
    int rc = iris_db_get_stream(c->shards[0].db, &rs);
    if (rc) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(rs);
    {   // who holds which rows: {index_base, rows} of every process
        std::lock_guard<std::mutex> lk(c->mu);
        SetDevice g(c->shards[0].device);
        c->h_layout[0] = c->index_base;
        c->h_layout[1] = local;
        uint64_t* mine = c->d_layout + 2 * (size_t)c->world;
        CCK(cudaMemcpyAsync(mine, c->h_layout, 16, cudaMemcpyHostToDevice, stream));
        const ncclResult_t r = n->AllGather(mine, c->d_layout, 2, ncclUint64, c->comm, stream);
        if (r != ncclSuccess) return cfail(IRIS_ERR_CUDA, "ncclAllGather failed: %s", n->GetErrorString(r));
        CCK(cudaMemcpyAsync(c->h_layout, c->d_layout, (size_t)c->world * 16, cudaMemcpyDeviceToHost, stream));
        CCK(cudaStreamSynchronize(stream));
    }
    std::vector<uint64_t> layout(c->h_layout, c->h_layout + 2 * (size_t)c->world);
    // local scan into this process's slot (the shards of this process store at their own offsets inside it)
    rc = cluster_match(c, query, nullptr, query_mask, query_mask != nullptr,
                       distances_out ? distances_out + c->index_base * IRIS_ROTATIONS : nullptr,
                       denominators_out ? denominators_out + c->index_base * IRIS_ROTATIONS : nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    SetDevice g(c->shards[0].device);
    ncclResult_t r = n->GroupStart();
    for (int p = 0; r == ncclSuccess && p < c->world; ++p) {
        const size_t off = (size_t)layout[2 * p] * IRIS_ROTATIONS, bytes = (size_t)layout[2 * p + 1] * IRIS_ROTATIONS * sizeof(uint16_t);
        if (bytes == 0) continue;
        for (uint16_t* o : {distances_out, denominators_out})
            if (o && r == ncclSuccess) r = n->Broadcast(o + off, o + off, bytes, ncclUint8, p, c->comm, stream);
    }
    if (r == ncclSuccess) r = n->GroupEnd();
    if (r != ncclSuccess) return cfail(IRIS_ERR_CUDA, "NCCL broadcast of the result blocks failed: %s", n->GetErrorString(r));
    CCK(cudaStreamSynchronize(stream));
    return IRIS_OK;
}
