// iris_engine.hpp -- header-only C++17 mirror of the reference crate's hot-path API over the C ABI
// (include/iris_b200.h).  Same names, argument meaning and failure behaviour as the Rust:
//
//   Bits, EncodedBits, Template                     src/bits.rs:13-15, src/encoded_bits.rs:13-15, src/template.rs:11-29
//   encode(&Template) -> EncodedBits                src/lib.rs:16-26
//   DistanceEngine::new / batch_process             src/lib.rs:28-52
//   MasksEngine::new / batch_process                src/lib.rs:55-79
//   distances(), denominators()                     src/lib.rs:82-94
//   decode_distance(&[u16;31], &[u16;31]) -> f64    src/lib.rs:97-107
//   arch::dot_u16 / arch::dot_bool                  src/arch/generic.rs:4-16
//   arch::dot_u16_grid / dot_bool_grid              the criterion grids, src/arch/mod.rs:22-72
//   Cluster                                         the whole mmapped file (src/main.rs:386-400) over several GPUs
//
// The reference panics on a length mismatch (assert_eq!, src/lib.rs:43,70); here every failure of the library,
// including that one, throws iris::Error carrying the status code.  There is no CPU fallback.
#pragma once
#include <array>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "iris_b200.h"

namespace iris {

constexpr std::size_t COLS = IRIS_COLS;
constexpr std::size_t ROWS = IRIS_ROWS;
constexpr std::size_t BITS = IRIS_BITS;
constexpr std::size_t LIMBS = IRIS_LIMBS;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) {
    if (rc != IRIS_OK) throw Error(rc, iris_last_error());
}

// #[repr(transparent)] [u64; 200]
struct Bits {
    std::array<uint64_t, LIMBS> limbs{};
    bool operator[](std::size_t i) const { return (limbs[i / 64] >> (i % 64)) & 1u; }   // src/bits.rs:44-57
};
// #[repr(transparent)] [u16; 12800]
struct EncodedBits {
    std::array<uint16_t, BITS> v{};
};
// #[repr(C)] { pattern, mask }
struct Template {
    Bits pattern, mask;
};
static_assert(sizeof(Bits) == 1600 && sizeof(EncodedBits) == 25600 && sizeof(Template) == 3200, "Pod layouts");
using Row31 = std::array<uint16_t, IRIS_ROTATIONS>;

namespace arch {
inline uint16_t dot_u16(const std::array<uint16_t, BITS>& a, const std::array<uint16_t, BITS>& b, int device = 0) {
    uint16_t out = 0;
    check(iris_dot_u16(device, a.data(), b.data(), &out));
    return out;
}
inline uint16_t dot_bool(const std::array<uint64_t, LIMBS>& a, const std::array<uint64_t, LIMBS>& b, int device = 0) {
    uint16_t out = 0;
    check(iris_dot_bool(device, a.data(), b.data(), &out));
    return out;
}
// The criterion grids (src/arch/mod.rs:22-72): out[i][j] = dot(a[j], b[i]) for every pair, one call.
inline std::vector<uint16_t> dot_u16_grid(const std::vector<std::array<uint16_t, BITS>>& a,
                                          const std::vector<std::array<uint16_t, BITS>>& b, int device = 0) {
    std::vector<uint16_t> out(a.size() * b.size());
    if (!out.empty()) check(iris_dot_u16_batch(device, a[0].data(), (uint32_t)a.size(), b[0].data(), b.size(), out.data()));
    return out;
}
inline std::vector<uint16_t> dot_bool_grid(const std::vector<std::array<uint64_t, LIMBS>>& a,
                                           const std::vector<std::array<uint64_t, LIMBS>>& b, int device = 0) {
    std::vector<uint16_t> out(a.size() * b.size());
    if (!out.empty()) check(iris_dot_bool_batch(device, a[0].data(), (uint32_t)a.size(), b[0].data(), b.size(), out.data()));
    return out;
}
}  // namespace arch

inline EncodedBits encode(const Template& t, int device = 0) {
    EncodedBits out;
    check(iris_encode(device, t.pattern.limbs.data(), t.mask.limbs.data(), out.v.data()));
    return out;
}

// HBM-resident shard: what the reference mmaps as &[EncodedBits] / &[Bits] (src/main.rs:389-391, 458-461).
class Database {
  public:
    Database(uint64_t capacity_rows, int device = 0, bool shares = true, bool masks = true) {
        check(iris_db_create(device, capacity_rows, (shares ? IRIS_DB_SHARES : 0u) | (masks ? IRIS_DB_MASKS : 0u), &h_));
    }
    ~Database() { iris_db_destroy(h_); }
    Database(const Database&) = delete;
    Database& operator=(const Database&) = delete;
    void append(const std::vector<EncodedBits>& rows) {
        check(iris_db_append_shares(h_, rows.empty() ? nullptr : rows[0].v.data(), rows.size()));
    }
    void append(const std::vector<Bits>& rows) {
        check(iris_db_append_masks(h_, rows.empty() ? nullptr : rows[0].limbs.data(), rows.size()));
    }
    void load_shares_file(const std::string& path) { check(iris_db_load_shares_file(h_, path.c_str(), 0, 0)); }
    void load_masks_file(const std::string& path) { check(iris_db_load_masks_file(h_, path.c_str(), 0, 0)); }
    void generate(uint64_t seed, uint64_t first_row_id, uint64_t n) { check(iris_db_generate(h_, seed, first_row_id, n)); }
    uint64_t len_shares() const {
        uint64_t a = 0, b = 0;
        check(iris_db_len(h_, &a, &b));
        return a;
    }
    uint64_t len_masks() const {
        uint64_t a = 0, b = 0;
        check(iris_db_len(h_, &a, &b));
        return b;
    }
    iris_db* handle() const { return h_; }

  private:
    iris_db* h_ = nullptr;
};

class DistanceEngine {
  public:
    explicit DistanceEngine(const EncodedBits& query, int device = 0) {        // DistanceEngine::new
        check(iris_distance_engine_new(device, query.v.data(), &h_));
    }
    explicit DistanceEngine(const Template& t, int device = 0) {               // new(&encode(&template)), src/main.rs:427
        check(iris_distance_engine_new_from_template(device, t.pattern.limbs.data(), t.mask.limbs.data(), &h_));
    }
    ~DistanceEngine() { iris_distance_engine_free(h_); }
    DistanceEngine(const DistanceEngine&) = delete;
    DistanceEngine& operator=(const DistanceEngine&) = delete;
    // batch_process(&self, out: &mut [[u16;31]], db: &[EncodedBits])  -- literal signature, host slices
    void batch_process(std::vector<Row31>& out, const std::vector<EncodedBits>& db) const {
        check(iris_distance_engine_batch_process(h_, out.empty() ? nullptr : out[0].data(), out.size(),
                                                 db.empty() ? nullptr : db[0].v.data(), db.size()));
    }
    // the participant's chunk loop (src/main.rs:428-431) against the resident shard
    void batch_process(std::vector<Row31>& out, const Database& db, uint64_t row_begin, uint64_t row_end) const {
        check(iris_distance_engine_batch_process_resident(h_, out.empty() ? nullptr : out[0].data(), out.size(), db.handle(),
                                                          row_begin, row_end));
    }
    iris_distance_engine* handle() const { return h_; }

  private:
    iris_distance_engine* h_ = nullptr;
};

class MasksEngine {
  public:
    explicit MasksEngine(const Bits& query, int device = 0) { check(iris_masks_engine_new(device, query.limbs.data(), &h_)); }
    ~MasksEngine() { iris_masks_engine_free(h_); }
    MasksEngine(const MasksEngine&) = delete;
    MasksEngine& operator=(const MasksEngine&) = delete;
    void batch_process(std::vector<Row31>& out, const std::vector<Bits>& db) const {
        check(iris_masks_engine_batch_process(h_, out.empty() ? nullptr : out[0].data(), out.size(),
                                              db.empty() ? nullptr : db[0].limbs.data(), db.size()));
    }
    void batch_process(std::vector<Row31>& out, const Database& db, uint64_t row_begin, uint64_t row_end) const {
        check(iris_masks_engine_batch_process_resident(h_, out.empty() ? nullptr : out[0].data(), out.size(), db.handle(),
                                                       row_begin, row_end));
    }
    iris_masks_engine* handle() const { return h_; }

  private:
    iris_masks_engine* h_ = nullptr;
};

inline Row31 distances(const EncodedBits& query, const EncodedBits& entry, int device = 0) {     // src/lib.rs:82-87
    Row31 out{};
    check(iris_distances(device, query.v.data(), entry.v.data(), out.data()));
    return out;
}
inline Row31 denominators(const Bits& query, const Bits& entry, int device = 0) {                // src/lib.rs:89-94
    Row31 out{};
    check(iris_denominators(device, query.limbs.data(), entry.limbs.data(), out.data()));
    return out;
}
// decode_distance (src/lib.rs:97-107), evaluated by the device reduction kernel (bit-identical f64).
inline double decode_distance(const Row31& dist, const Row31& den, int device = 0) {
    const uint16_t* share = dist.data();
    double per_row = 0, m = 0;
    uint64_t idx = 0;
    check(iris_combine_min(device, &share, 1, den.data(), 1, 0, &per_row, &m, &idx));
    return per_row;
}
// fused single pass over the resident shard
inline void match(const DistanceEngine& de, const MasksEngine& me, const Database& db, uint64_t row_begin, uint64_t row_end,
                  std::vector<Row31>& dist, std::vector<Row31>& den) {
    if (dist.size() != row_end - row_begin || den.size() != row_end - row_begin)
        throw Error(IRIS_ERR_INVALID, "out.len() != db.len()");
    check(iris_match_resident(de.handle(), me.handle(), db.handle(), row_begin, row_end, dist.empty() ? nullptr : dist[0].data(),
                              den.empty() ? nullptr : den[0].data()));
}
// coordinator reduction: (min distance, argmin) with the reference's first-minimum rule (src/main.rs:611-621)
inline std::pair<double, uint64_t> match_min(const DistanceEngine& de, const MasksEngine& me, const Database& db,
                                             uint64_t row_begin, uint64_t row_end, uint64_t index_base = 0) {
    double m = std::numeric_limits<double>::infinity();
    uint64_t i = 0;
    check(iris_match_min_resident(de.handle(), me.handle(), db.handle(), row_begin, row_end, index_base, &m, &i));
    return {m, i};
}


// One database row-sharded over several GPUs behind one handle (iris_cluster_*): what the participant / coordinator
// mmap as one file (src/main.rs:386-400, 458-461), with the chunk loops of src/main.rs:425-431, 510-516 run by all GPUs
// at once and the coordinator's running minimum (src/main.rs:597-621) continued across them.
class Cluster {
  public:
    Cluster(const std::vector<int>& devices, uint64_t capacity_rows, bool shares = true, bool masks = true) {
        check(iris_cluster_create(devices.data(), (uint32_t)devices.size(), capacity_rows,
                                  (shares ? IRIS_DB_SHARES : 0u) | (masks ? IRIS_DB_MASKS : 0u), &h_));
    }
    ~Cluster() { iris_cluster_destroy(h_); }
    Cluster(const Cluster&) = delete;
    Cluster& operator=(const Cluster&) = delete;
    void load_files(const char* shares_path, const char* masks_path) { check(iris_cluster_load_files(h_, shares_path, masks_path)); }
    void load(const std::vector<EncodedBits>& shares, const std::vector<Bits>& masks) {
        check(iris_cluster_load_rows(h_, shares.empty() ? nullptr : shares[0].v.data(), masks.empty() ? nullptr : masks[0].limbs.data(),
                                     shares.empty() ? masks.size() : shares.size()));
    }
    uint64_t len() const {
        uint64_t a = 0, b = 0;
        check(iris_cluster_len(h_, nullptr, &a, &b));
        return a > b ? a : b;
    }
    // the participant's request: distances of every row for one template
    void distances(const Template& t, std::vector<Row31>& out) const {
        if (out.size() != len()) throw Error(IRIS_ERR_INVALID, "out.len() != db.len()");
        check(iris_cluster_match_template(h_, t.pattern.limbs.data(), t.mask.limbs.data(), out.empty() ? nullptr : out[0].data(), nullptr));
    }
    // the coordinator's side: denominators of every row for one query mask
    void denominators(const Bits& mask, std::vector<Row31>& out) const {
        if (out.size() != len()) throw Error(IRIS_ERR_INVALID, "out.len() != db.len()");
        check(iris_cluster_match(h_, nullptr, mask.limbs.data(), nullptr, out.empty() ? nullptr : out[0].data()));
    }
    // whole search for a cluster holding plain encodings: (min distance, row) per template
    std::vector<std::pair<double, uint64_t>> search(const std::vector<Template>& queries) const {
        std::vector<double> d(queries.size());
        std::vector<uint64_t> i(queries.size());
        if (!queries.empty())
            check(iris_cluster_search(h_, queries[0].pattern.limbs.data(), (uint32_t)queries.size(), d.data(), i.data()));
        std::vector<std::pair<double, uint64_t>> out(queries.size());
        for (std::size_t k = 0; k < queries.size(); ++k) out[k] = {d[k], i[k]};
        return out;
    }
    iris_cluster* handle() const { return h_; }

  private:
    iris_cluster* h_ = nullptr;
};

}  // namespace iris
