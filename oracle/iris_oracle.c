/*
 * iris_oracle.c -- CPU restatement of the recmo/mpc-iris-code matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mpc-iris-code_b200/ may include, link or
 * call this file; it is the checker for the CUDA path (tests/, __graft_entry__.smoke(),
 * and bench.py's cpu_baseline / --impl reference legs), never the product.
 *
 * Each function is a plain-C restatement of one reference function and cites the
 * reference file:line it follows (paths relative to the reference crate root).
 * The reference is Rust and there is no Rust toolchain in this image, so the real
 * crate cannot be compiled here (no oracle/_ref).  The arithmetic is all in-tree in
 * the reference (src/arch/generic.rs); nothing comes from a third-party dependency.
 *
 * Pinning: the reference holds no stored vector for dot_u16 / dot_bool values (its only
 * file-backed golden test needs data/templates.json + data/distances.json, which are
 * git-ignored upstream and absent).  This oracle is pinned against every test of the
 * reference that can run without those files -- test_rotated_number (deterministic
 * known answer), test_rotated_inverse (x2), test_rotated_bits, test_index,
 * test_preprocess, test_dotproduct, sve::test_dot_u16, and the
 * encrypted-path == plaintext-path relation that test_encrypted_distances asserts --
 * see tests/test_oracle.py and tests/golden/.  Dot VALUES are therefore pinned by
 * definition + identities, not by an upstream stored vector ("parity unpinned" in the
 * strict sense for stored vectors; stated in DESIGN.md too).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define IRIS_COLS 200              /* src/lib.rs:10 */
#define IRIS_ROWS 64               /* src/lib.rs:11 */
#define IRIS_BITS 12800            /* src/lib.rs:12 */
#define IRIS_LIMBS 200             /* src/bits.rs:10 */
#define IRIS_BYTES_PER_COL 25      /* src/bits.rs:11 */
#define IRIS_ROTATIONS 31          /* src/lib.rs:34 (-15..=15) */

/* src/arch/generic.rs:4-9 -- sum of popcount(a&b) folded with u16 wrapping_add. */
uint16_t oracle_dot_bool(const uint64_t *a, const uint64_t *b) {
    uint16_t acc = 0;
    for (int i = 0; i < IRIS_LIMBS; ++i)
        acc = (uint16_t)(acc + (uint16_t)__builtin_popcountll(a[i] & b[i]));
    return acc;
}

/* src/arch/generic.rs:11-16 -- sum of wrapping_mul folded with wrapping_add, in u16. */
uint16_t oracle_dot_u16(const uint16_t *a, const uint16_t *b) {
    uint16_t acc = 0;
    for (int i = 0; i < IRIS_BITS; ++i)
        acc = (uint16_t)(acc + (uint16_t)((uint32_t)a[i] * (uint32_t)b[i])); /* u32: u16*u16 promoted to int would overflow */
    return acc;
}

/* slice::rotate_left / rotate_right on a 25-byte row (used by src/bits.rs:179-185). */
static void bytes_rotate_left(uint8_t *a, int n, int mid) {
    uint8_t tmp[IRIS_BYTES_PER_COL];
    for (int i = 0; i < n; ++i) tmp[i] = a[(i + mid) % n];
    memcpy(a, tmp, (size_t)n);
}
static void bytes_rotate_right(uint8_t *a, int n, int k) {
    uint8_t tmp[IRIS_BYTES_PER_COL];
    for (int i = 0; i < n; ++i) tmp[(i + k) % n] = a[i];
    memcpy(a, tmp, (size_t)n);
}

/* src/bits.rs:178-205 rotate_row, statement by statement (C '%' truncates like Rust's). */
static void rotate_row(uint8_t *a, int amount) {
    if (amount <= -8) {
        bytes_rotate_left(a, IRIS_BYTES_PER_COL, (-amount) / 8);
        amount %= 8;
    } else if (amount >= 8) {
        bytes_rotate_right(a, IRIS_BYTES_PER_COL, amount / 8);
        amount %= 8;
    }
    if (amount < 0) {
        int r = -amount, l = 8 - r;
        uint8_t carry = (uint8_t)(a[0] << l);
        for (int i = IRIS_BYTES_PER_COL - 1; i >= 0; --i) {
            uint8_t old = a[i];
            a[i] = (uint8_t)((old >> r) | carry);
            carry = (uint8_t)(old << l);
        }
    } else if (amount > 0) {
        int l = amount, r = 8 - l;
        uint8_t carry = (uint8_t)(a[24] >> r);
        for (int i = 0; i < IRIS_BYTES_PER_COL; ++i) {
            uint8_t old = a[i];
            a[i] = (uint8_t)((old << l) | carry);
            carry = (uint8_t)(old >> r);
        }
    }
}

/* src/bits.rs:18-23 Bits::rotate -- 64 chunks of 25 bytes over the 1600 raw bytes. */
void oracle_bits_rotate(uint64_t *bits, int amount) {
    uint8_t *bytes = (uint8_t *)bits;
    for (int row = 0; row < IRIS_ROWS; ++row) rotate_row(bytes + row * IRIS_BYTES_PER_COL, amount);
}

/* src/bits.rs:44-57 Index<usize>: bit i = limb i/64, bit i%64. */
int oracle_bits_get(const uint64_t *bits, int index) {
    return (int)((bits[index / 64] >> (index % 64)) & 1u);
}

/* src/bits.rs:31-33 count_ones. */
uint16_t oracle_bits_count_ones(const uint64_t *bits) {
    uint16_t acc = 0;
    for (int i = 0; i < IRIS_LIMBS; ++i) acc = (uint16_t)(acc + __builtin_popcountll(bits[i]));
    return acc;
}

/* src/encoded_bits.rs:40-52 EncodedBits::rotate -- rotate_left(|a|) for a<0, rotate_right(a) for a>0
 * on each 200-element chunk. */
void oracle_encoded_rotate(uint16_t *v, int amount) {
    uint16_t tmp[IRIS_COLS];
    if (amount == 0) return;
    for (int row = 0; row < IRIS_ROWS; ++row) {
        uint16_t *r = v + row * IRIS_COLS;
        if (amount < 0) {
            int mid = -amount;
            for (int i = 0; i < IRIS_COLS; ++i) tmp[i] = r[(i + mid) % IRIS_COLS];
        } else {
            for (int i = 0; i < IRIS_COLS; ++i) tmp[(i + amount) % IRIS_COLS] = r[i];
        }
        memcpy(r, tmp, sizeof tmp);
    }
}

/* src/encoded_bits.rs:75-79 From<&Bits>: 0/1 embedding. */
void oracle_encoded_from_bits(const uint64_t *bits, uint16_t *out) {
    for (int i = 0; i < IRIS_BITS; ++i) out[i] = (uint16_t)oracle_bits_get(bits, i);
}

/* src/encoded_bits.rs:60-62 sum (wrapping). */
uint16_t oracle_encoded_sum(const uint16_t *v) {
    uint16_t acc = 0;
    for (int i = 0; i < IRIS_BITS; ++i) acc = (uint16_t)(acc + v[i]);
    return acc;
}

/* src/lib.rs:16-26 encode: pattern&=mask; mask - pattern - pattern in Z/2^16. */
void oracle_encode(const uint64_t *pattern, const uint64_t *mask, uint16_t *out) {
    for (int i = 0; i < IRIS_BITS; ++i) {
        uint16_t m = (uint16_t)oracle_bits_get(mask, i);
        uint16_t p = (uint16_t)(oracle_bits_get(pattern, i) & (int)m);
        out[i] = (uint16_t)(m - p - p);
    }
}

/* src/encoded_bits.rs:23-38 share(n): the first n-1 shares are caller-supplied uniform
 * vectors (the reference draws them from thread_rng); last = self - sum(rest). */
void oracle_share_last(const uint16_t *self, const uint16_t *rest, size_t n_rest, uint16_t *last) {
    for (int i = 0; i < IRIS_BITS; ++i) {
        uint16_t s = 0;
        for (size_t j = 0; j < n_rest; ++j) s = (uint16_t)(s + rest[j * IRIS_BITS + i]);
        last[i] = (uint16_t)(self[i] - s);
    }
}

/* src/lib.rs:33-40 DistanceEngine::new: rotations[j] = query.rotated(j-15). */
void oracle_distance_rotations(const uint16_t *query, uint16_t *rotations /* [31][12800] */) {
    for (int j = 0; j < IRIS_ROTATIONS; ++j) {
        uint16_t *dst = rotations + (size_t)j * IRIS_BITS;
        memcpy(dst, query, IRIS_BITS * sizeof(uint16_t));
        oracle_encoded_rotate(dst, j - 15);
    }
}

/* src/lib.rs:60-67 MasksEngine::new. */
void oracle_mask_rotations(const uint64_t *query, uint64_t *rotations /* [31][200] */) {
    for (int j = 0; j < IRIS_ROTATIONS; ++j) {
        uint64_t *dst = rotations + (size_t)j * IRIS_LIMBS;
        memcpy(dst, query, IRIS_LIMBS * sizeof(uint64_t));
        oracle_bits_rotate(dst, j - 15);
    }
}

/* src/lib.rs:42-52 DistanceEngine::batch_process.  rayon par_iter over rows -> OpenMP static. */
void oracle_distance_batch(const uint16_t *rotations, const uint16_t *db, size_t n,
                           uint16_t *out /* [n][31] */, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long i = 0; i < (long)n; ++i) {
        const uint16_t *entry = db + (size_t)i * IRIS_BITS;
        for (int j = 0; j < IRIS_ROTATIONS; ++j)
            out[(size_t)i * IRIS_ROTATIONS + j] = oracle_dot_u16(rotations + (size_t)j * IRIS_BITS, entry);
    }
}

/* src/lib.rs:69-79 MasksEngine::batch_process. */
void oracle_masks_batch(const uint64_t *rotations, const uint64_t *db, size_t n,
                        uint16_t *out /* [n][31] */, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long i = 0; i < (long)n; ++i) {
        const uint64_t *entry = db + (size_t)i * IRIS_LIMBS;
        for (int j = 0; j < IRIS_ROTATIONS; ++j)
            out[(size_t)i * IRIS_ROTATIONS + j] = oracle_dot_bool(rotations + (size_t)j * IRIS_LIMBS, entry);
    }
}

/* src/lib.rs:82-87 distances(query, entry). */
void oracle_distances(const uint16_t *query, const uint16_t *entry, uint16_t *out31) {
    static _Thread_local uint16_t rot[IRIS_ROTATIONS * IRIS_BITS];
    oracle_distance_rotations(query, rot);
    oracle_distance_batch(rot, entry, 1, out31, 1);
}

/* src/lib.rs:89-94 denominators(query, entry). */
void oracle_denominators(const uint64_t *query, const uint64_t *entry, uint16_t *out31) {
    uint64_t rot[IRIS_ROTATIONS * IRIS_LIMBS];
    oracle_mask_rotations(query, rot);
    oracle_masks_batch(rot, entry, 1, out31, 1);
}

/* src/lib.rs:97-107 decode_distance: ((d - n) wrapping / 2) / d as f64, fold(INFINITY, f64::min).
 * f64::min ignores a NaN operand; C fmin has the same rule. */
double oracle_decode_distance(const uint16_t *distances, const uint16_t *denominators) {
    double best = INFINITY;
    for (int j = 0; j < IRIS_ROTATIONS; ++j) {
        uint16_t n = (uint16_t)((uint16_t)(denominators[j] - distances[j]) / 2);
        double f = (double)n / (double)denominators[j];
        best = fmin(best, f);
    }
    return best;
}

/* src/template.rs:49-64 fraction_hamming on (pattern, mask) pairs. */
double oracle_fraction_hamming(const uint64_t *ap, const uint64_t *am, const uint64_t *bp, const uint64_t *bm) {
    uint32_t num = 0, den = 0;
    for (int i = 0; i < IRIS_LIMBS; ++i) {
        uint64_t m = am[i] & bm[i];
        uint64_t p = (ap[i] ^ bp[i]) & m;
        num += (uint32_t)__builtin_popcountll(p);
        den += (uint32_t)__builtin_popcountll(m);
    }
    return (double)num / (double)den;
}

/* src/template.rs:43-47 Template::distance: min over self.rotated(r).fraction_hamming(other);
 * the fold uses |a,b| a.min(b) (NaN-ignoring). */
double oracle_template_distance(const uint64_t *ap, const uint64_t *am, const uint64_t *bp, const uint64_t *bm) {
    double best = INFINITY;
    for (int r = -15; r <= 15; ++r) {
        uint64_t rp[IRIS_LIMBS], rm[IRIS_LIMBS];
        memcpy(rp, ap, sizeof rp);
        memcpy(rm, am, sizeof rm);
        oracle_bits_rotate(rp, r); /* src/template.rs:32-35 rotate mask and pattern */
        oracle_bits_rotate(rm, r);
        best = fmin(best, oracle_fraction_hamming(rp, rm, bp, bm));
    }
    return best;
}

/* ---- synthetic database rows (not from the reference: the repo's own data spec, restated
 * here independently of the device generator in csrc/ so sampled rows can be re-derived on
 * the CPU).  Uniform u16 shares are what share() produces (src/encoded_bits.rs:23-38,81-87);
 * uniform mask bits are what Standard for Bits produces (src/bits.rs:95-101). ---- */
static inline uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_gen_share_row(uint64_t seed, uint64_t row, uint16_t *out /* [12800] */) {
    for (uint64_t g = 0; g < IRIS_BITS / 4; ++g) {
        uint64_t h = mix64(seed ^ ((row * (IRIS_BITS / 4) + g) * 0xD1342543DE82EF95ull));
        out[4 * g + 0] = (uint16_t)h;
        out[4 * g + 1] = (uint16_t)(h >> 16);
        out[4 * g + 2] = (uint16_t)(h >> 32);
        out[4 * g + 3] = (uint16_t)(h >> 48);
    }
}

void oracle_gen_mask_row(uint64_t seed, uint64_t row, uint64_t *out /* [200] */) {
    for (uint64_t l = 0; l < IRIS_LIMBS; ++l)
        out[l] = mix64((seed ^ 0xA5A5A5A55A5A5A5Aull) ^ ((row * IRIS_LIMBS + l) * 0xD1342543DE82EF95ull));
}

/* Synthetic Templates and their additive shares (the data spec in csrc/iris_layout.h, restated): row R is the Template
 * (pattern_R, mask_R) with mask_R = oracle_gen_mask_row and pattern_R drawn the same way under another tag; party p <
 * n-1 holds a uniform vector (oracle_gen_share_row under the party's seed), the last party holds
 * encode(Template R) - sum of the others -- EncodedBits::share, src/encoded_bits.rs:23-38, via oracle_share_last. */
void oracle_gen_pattern_row(uint64_t seed, uint64_t row, uint64_t *out /* [200] */) {
    for (uint64_t l = 0; l < IRIS_LIMBS; ++l)
        out[l] = mix64((seed ^ 0x5A5A5A5AA5A5A5A5ull) ^ ((row * IRIS_LIMBS + l) * 0xD1342543DE82EF95ull));
}

static uint64_t party_seed(uint64_t seed, uint32_t p) { return mix64(seed ^ (0xC2B2AE3D27D4EB4Full * (uint64_t)(p + 1))); }

void oracle_gen_party_share_row(uint64_t seed, uint32_t party, uint32_t n_parties, uint64_t row, uint16_t *out /* [12800] */) {
    if (party + 1 < n_parties) {
        oracle_gen_share_row(party_seed(seed, party), row, out);
        return;
    }
    static _Thread_local uint16_t enc[IRIS_BITS], rest[IRIS_BITS], acc[IRIS_BITS];
    uint64_t pattern[IRIS_LIMBS], mask[IRIS_LIMBS];
    oracle_gen_pattern_row(seed, row, pattern);
    oracle_gen_mask_row(seed, row, mask);
    oracle_encode(pattern, mask, enc);
    memset(acc, 0, sizeof acc);
    for (uint32_t q = 0; q + 1 < n_parties; ++q) {
        oracle_gen_share_row(party_seed(seed, q), row, rest);
        for (int i = 0; i < IRIS_BITS; ++i) acc[i] = (uint16_t)(acc[i] + rest[i]);
    }
    oracle_share_last(enc, acc, 1, out);
}

void oracle_gen_party_share_rows(uint64_t seed, uint32_t party, uint32_t n_parties, uint64_t row0, size_t n, uint16_t *out, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long i = 0; i < (long)n; ++i)
        oracle_gen_party_share_row(seed, party, n_parties, row0 + (uint64_t)i, out + (size_t)i * IRIS_BITS);
}

void oracle_gen_pattern_rows(uint64_t seed, uint64_t row0, size_t n, uint64_t *out, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long i = 0; i < (long)n; ++i) oracle_gen_pattern_row(seed, row0 + (uint64_t)i, out + (size_t)i * IRIS_LIMBS);
}

void oracle_gen_share_rows(uint64_t seed, uint64_t row0, size_t n, uint16_t *out, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long i = 0; i < (long)n; ++i) oracle_gen_share_row(seed, row0 + (uint64_t)i, out + (size_t)i * IRIS_BITS);
}

void oracle_gen_mask_rows(uint64_t seed, uint64_t row0, size_t n, uint64_t *out, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long i = 0; i < (long)n; ++i) oracle_gen_mask_row(seed, row0 + (uint64_t)i, out + (size_t)i * IRIS_LIMBS);
}

/* Combine step of the coordinator (src/main.rs:597-621): numerator[r] = wrapping sum of the
 * parties' distance shares; decode_distance; running min / argmin over rows (first minimum
 * wins: the reference uses `if distance < min_distance`). */
void oracle_combine_min(const uint16_t *dist_shares /* [parties][n][31] */, size_t parties,
                        const uint16_t *denoms /* [n][31] */, size_t n, double *min_distance, int64_t *min_index) {
    double best = INFINITY;
    int64_t best_i = -1;
    for (size_t i = 0; i < n; ++i) {
        uint16_t num[IRIS_ROTATIONS];
        for (int j = 0; j < IRIS_ROTATIONS; ++j) {
            uint16_t s = 0;
            for (size_t p = 0; p < parties; ++p) s = (uint16_t)(s + dist_shares[(p * n + i) * IRIS_ROTATIONS + j]);
            num[j] = s;
        }
        double d = oracle_decode_distance(num, denoms + i * IRIS_ROTATIONS);
        if (d < best) {
            best = d;
            best_i = (int64_t)i;
        }
    }
    *min_distance = best;
    *min_index = best_i;
}
