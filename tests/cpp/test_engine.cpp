// C++ rendition of the reference's unit tests (src/lib.rs:117-193) against include/iris_engine.hpp, with the GPU
// library in place of arch::generic and the C oracle (oracle/iris_oracle.c) as the checker.
// Built and run by tests/test_cpp_mirror_gpu.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "iris_engine.hpp"

extern "C" {
void oracle_encode(const uint64_t* pattern, const uint64_t* mask, uint16_t* out);
void oracle_distances(const uint16_t* query, const uint16_t* entry, uint16_t* out31);
void oracle_denominators(const uint64_t* query, const uint64_t* entry, uint16_t* out31);
double oracle_decode_distance(const uint16_t* distances, const uint16_t* denominators);
double oracle_template_distance(const uint64_t* ap, const uint64_t* am, const uint64_t* bp, const uint64_t* bm);
uint16_t oracle_dot_u16(const uint16_t* a, const uint16_t* b);
uint16_t oracle_dot_bool(const uint64_t* a, const uint64_t* b);
}

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);    \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

static std::mt19937_64 rng(0x1715C0DE);
static iris::Bits random_bits() {
    iris::Bits b;
    for (auto& l : b.limbs) l = rng();
    return b;
}
static iris::Template random_template() { return {random_bits(), random_bits()}; }
static iris::EncodedBits random_encoded() {
    iris::EncodedBits e;
    for (auto& x : e.v) x = (uint16_t)rng();
    return e;
}

static void test_preprocess() {   // src/lib.rs:117-132
    for (int it = 0; it < 10; ++it) {
        iris::Template t = random_template();
        iris::EncodedBits enc = iris::encode(t);
        for (std::size_t i = 0; i < iris::BITS; ++i) {
            const uint16_t v = enc.v[i];
            if (v == 0xFFFF) CHECK(t.mask[i] && t.pattern[i]);
            else if (v == 0) CHECK(!t.mask[i]);
            else if (v == 1) CHECK(t.mask[i] && !t.pattern[i]);
            else CHECK(false);
        }
        iris::EncodedBits ref;
        oracle_encode(t.pattern.limbs.data(), t.mask.limbs.data(), ref.v.data());
        CHECK(enc.v == ref.v);
    }
}

static void test_dotproduct() {   // src/lib.rs:134-163
    for (int it = 0; it < 10; ++it) {
        iris::Template a = random_template(), b = random_template();
        iris::EncodedBits pa = iris::encode(a), pb = iris::encode(b);
        int equal = 0, uneq = 0, denominator = 0;
        for (std::size_t i = 0; i < iris::BITS; ++i)
            if (a.mask[i] && b.mask[i]) {
                ++denominator;
                (a.pattern[i] == b.pattern[i] ? equal : uneq)++;
            }
        const int16_t sum = (int16_t)iris::arch::dot_u16(pa.v, pb.v);
        CHECK(equal - uneq == sum);
        CHECK(equal + uneq == denominator);
        CHECK((denominator - sum) % 2 == 0);
        CHECK(uneq == (denominator - sum) / 2);
        CHECK(iris::arch::dot_bool(a.mask.limbs, b.mask.limbs) == denominator);
        CHECK(iris::arch::dot_u16(pa.v, pb.v) == oracle_dot_u16(pa.v.data(), pb.v.data()));
    }
}

static void test_encrypted_distances() {   // src/lib.rs:165-193 (stored distances absent upstream: plaintext path instead)
    for (int it = 0; it < 10; ++it) {
        iris::Template query = random_template(), entry = random_template();
        iris::EncodedBits encrypted = iris::encode(entry), preprocessed = iris::encode(query);
        iris::Row31 d = iris::distances(preprocessed, encrypted);
        iris::Row31 n = iris::denominators(query.mask, entry.mask);
        iris::Row31 rd{}, rn{};
        oracle_distances(preprocessed.v.data(), encrypted.v.data(), rd.data());
        oracle_denominators(query.mask.limbs.data(), entry.mask.limbs.data(), rn.data());
        CHECK(d == rd);
        CHECK(n == rn);
        const double actual = iris::decode_distance(d, n);
        CHECK(actual == oracle_decode_distance(d.data(), n.data()));
        const double expected = oracle_template_distance(query.pattern.limbs.data(), query.mask.limbs.data(),
                                                         entry.pattern.limbs.data(), entry.mask.limbs.data());
        CHECK(std::fabs(actual - expected) <= std::nextafter(expected, INFINITY) - expected);
    }
}

static void test_engines() {   // src/lib.rs:28-79
    const std::size_t n = 300;
    std::vector<iris::EncodedBits> shares(n);
    std::vector<iris::Bits> masks(n);
    for (auto& s : shares) s = random_encoded();
    for (auto& m : masks) m = random_bits();
    iris::Template q = random_template();
    iris::EncodedBits uq = random_encoded();
    iris::DistanceEngine de(q), du(uq);
    iris::MasksEngine me(q.mask);
    std::vector<iris::Row31> out(n), out2(n), exp(n);
    iris::EncodedBits enc = iris::encode(q);
    // literal host-slice batch_process
    de.batch_process(out, shares);
    for (std::size_t i = 0; i < n; ++i) oracle_distances(enc.v.data(), shares[i].v.data(), exp[i].data());
    CHECK(out == exp);
    du.batch_process(out, shares);
    for (std::size_t i = 0; i < n; ++i) oracle_distances(uq.v.data(), shares[i].v.data(), exp[i].data());
    CHECK(out == exp);
    me.batch_process(out, masks);
    for (std::size_t i = 0; i < n; ++i) oracle_denominators(q.mask.limbs.data(), masks[i].limbs.data(), exp[i].data());
    CHECK(out == exp);
    // resident shard, chunked like the participant (src/main.rs:428)
    iris::Database db(n);
    db.append(shares);
    db.append(masks);
    CHECK(db.len_shares() == n && db.len_masks() == n);
    std::vector<iris::Row31> part(100);
    for (std::size_t b = 0; b < n; b += 100) {
        me.batch_process(part, db, b, b + 100);
        for (std::size_t i = 0; i < 100; ++i) CHECK(part[i] == exp[b + i]);
    }
    iris::match(du, me, db, 0, n, out, out2);
    CHECK(out2 == exp);
    // assert_eq!(out.len(), db.len()) (src/lib.rs:43) -> Error(IRIS_ERR_INVALID)
    std::vector<iris::Row31> wrong(n - 1);
    bool threw = false;
    try {
        de.batch_process(wrong, shares);
    } catch (const iris::Error& e) {
        threw = e.code == IRIS_ERR_INVALID;
    }
    CHECK(threw);
    // coordinator reduction against a brute-force scan of decode_distance
    std::vector<iris::EncodedBits> encs(n);
    for (std::size_t i = 0; i < n; ++i) {
        iris::Template t{random_bits(), masks[i]};
        if (i == 77) t.pattern = q.pattern, t.mask = q.mask, masks[i] = q.mask;
        encs[i] = iris::encode(t);
    }
    iris::Database db2(n);
    db2.append(encs);
    db2.append(masks);
    auto best = iris::match_min(de, me, db2, 0, n, 1000);
    CHECK(best.first == 0.0 && best.second == 1077);
}

static void test_cluster_and_grids() {
    // one database over three shards (on as many GPUs as the box has): the participant's and the coordinator's chunk
    // loops (src/main.rs:425-431, 510-516) and the coordinator's running minimum (src/main.rs:597-621)
    int gpus = 1;
    iris::check(iris_device_count(&gpus));
    const std::size_t n = 777;
    std::vector<iris::Bits> masks(n);
    std::vector<iris::EncodedBits> encs(n);
    std::vector<iris::Template> tmpl(n);
    for (std::size_t i = 0; i < n; ++i) {
        tmpl[i] = random_template();
        masks[i] = tmpl[i].mask;
        encs[i] = iris::encode(tmpl[i]);
    }
    iris::Cluster c({0, 1 % gpus, 2 % gpus}, n);
    c.load(encs, masks);
    CHECK(c.len() == n);
    iris::Template q = tmpl[n - 1];                    // a noisy copy of the last row: it lives in the last shard
    q.pattern.limbs[3] ^= 0x00FF00FF00FF00FFull;
    std::vector<iris::Row31> dist(n), den(n), exp(n);
    c.distances(q, dist);
    iris::EncodedBits enc = iris::encode(q);
    for (std::size_t i = 0; i < n; ++i) oracle_distances(enc.v.data(), encs[i].v.data(), exp[i].data());
    CHECK(dist == exp);
    c.denominators(q.mask, den);
    for (std::size_t i = 0; i < n; ++i) oracle_denominators(q.mask.limbs.data(), masks[i].limbs.data(), exp[i].data());
    CHECK(den == exp);
    double best = std::numeric_limits<double>::infinity();
    uint64_t arg = ~0ull;
    for (std::size_t i = 0; i < n; ++i) {
        const double d = iris::decode_distance(dist[i], den[i]);
        if (d < best) best = d, arg = i;
    }
    auto r = c.search({q, tmpl[5]});
    CHECK(r[0].first == best && r[0].second == arg && arg == n - 1);
    CHECK(r[1].first == 0.0 && r[1].second == 5);
    // the criterion grids (src/arch/mod.rs:22-72): every pair of independent vectors
    std::vector<std::array<uint16_t, iris::BITS>> a(5), b(40);
    for (auto& v : a) v = random_encoded().v;
    for (auto& v : b) v = random_encoded().v;
    auto grid = iris::arch::dot_u16_grid(a, b);
    bool ok = true;
    for (std::size_t i = 0; i < b.size(); ++i)
        for (std::size_t j = 0; j < a.size(); ++j) ok &= grid[i * a.size() + j] == oracle_dot_u16(a[j].data(), b[i].data());
    CHECK(ok);
}

int main() {
    try {
        test_preprocess();
        test_dotproduct();
        test_encrypted_distances();
        test_engines();
        test_cluster_and_grids();
    } catch (const iris::Error& e) {
        std::printf("FAIL exception %d: %s\n", e.code, e.what());
        return 2;
    }
    std::printf(failures ? "FAILED %d checks\n" : "ALL PASS\n", failures);
    return failures ? 1 : 0;
}
