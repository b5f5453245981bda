"""Pins the CPU oracle (oracle/) against every reference test that runs without data/.

Each test names the reference test it re-creates (file:line relative to the reference crate).
The reference uses thread_rng(); here every input is seeded.
"""
import json
import os

import numpy as np
import pytest

import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rng(seed):
    return np.random.default_rng(seed)


def rand_bits(r):
    return r.integers(0, 2**64, size=O.LIMBS, dtype=np.uint64)


def rand_encoded(r):
    return r.integers(0, 2**16, size=O.BITS, dtype=np.uint16)


def test_limbs_exact():
    # src/bits.rs:213-216
    assert O.LIMBS * 64 == O.BITS
    assert 25 * 8 == O.COLS


def test_index():
    # src/bits.rs:218-232: bit i <-> byte i/8, bit i%8
    r = rng(1)
    lib = O.load()
    for _ in range(5):
        bits = rand_bits(r)
        raw = bits.view(np.uint8)
        got = np.array([lib.oracle_bits_get(bits.ctypes.data_as(O.ctypes.POINTER(O.ctypes.c_uint64)), i) for i in range(O.BITS)])
        exp = (raw[np.arange(O.BITS) // 8] >> (np.arange(O.BITS) % 8)) & 1
        assert np.array_equal(got, exp)
        assert np.array_equal(O.np_bits_to_bool(bits), exp)


def test_rotated_inverse_bits():
    # src/bits.rs:234-247
    r = rng(2)
    for _ in range(20):
        bits = rand_bits(r)
        for a in range(-15, 16):
            assert np.array_equal(O.bits_rotated(O.bits_rotated(bits, a), -a), bits), a


def test_rotated_inverse_encoded():
    # src/encoded_bits.rs:189-203
    r = rng(3)
    for _ in range(20):
        v = rand_encoded(r)
        for a in range(-15, 16):
            assert np.array_equal(O.encoded_rotated(O.encoded_rotated(v, a), -a), v), a


def test_rotated_number():
    # src/encoded_bits.rs:205-219 -- the reference's only deterministic known-answer vector.
    i = np.arange(O.BITS)
    row, col = i // O.COLS, i % O.COLS
    secret = ((row << 8) | col).astype(np.uint16)
    for a in range(-15, 16):
        rotated = O.encoded_rotated(secret, a)
        exp_col = ((O.COLS + col) - a) % O.COLS
        assert np.array_equal(rotated, ((row << 8) | exp_col).astype(np.uint16)), a
        assert np.array_equal(O.np_encoded_rotated(secret, a), rotated), a


def test_rotated_bits():
    # src/encoded_bits.rs:221-236: Bits and EncodedBits rotate identically.
    r = rng(4)
    for _ in range(20):
        bits = rand_bits(r)
        secret = O.encoded_from_bits(bits)
        for a in range(-15, 16):
            assert np.array_equal(O.encoded_from_bits(O.bits_rotated(bits, a)), O.encoded_rotated(secret, a)), a
            assert np.array_equal(O.np_bits_rotated(bits, a), O.bits_rotated(bits, a)), a


def test_preprocess():
    # src/lib.rs:117-132
    r = rng(5)
    for _ in range(20):
        pattern, mask = rand_bits(r), rand_bits(r)
        enc = O.encode(pattern, mask)
        p, m = O.np_bits_to_bool(pattern).astype(bool), O.np_bits_to_bool(mask).astype(bool)
        assert set(np.unique(enc)) <= {0, 1, 0xFFFF}
        assert np.array_equal(enc == 0xFFFF, m & p)
        assert np.array_equal(enc == 0, ~m)
        assert np.array_equal(enc == 1, m & ~p)
        assert np.array_equal(enc, O.np_encode(pattern, mask))


def test_dotproduct():
    # src/lib.rs:134-163
    r = rng(6)
    for _ in range(20):
        ap, am, bp, bm = (rand_bits(r) for _ in range(4))
        pre_a, pre_b = O.encode(ap, am), O.encode(bp, bm)
        A_p, A_m = O.np_bits_to_bool(ap).astype(bool), O.np_bits_to_bool(am).astype(bool)
        B_p, B_m = O.np_bits_to_bool(bp).astype(bool), O.np_bits_to_bool(bm).astype(bool)
        joint = A_m & B_m
        denominator = int(joint.sum())
        equal = int((joint & (A_p == B_p)).sum())
        uneq = int((joint & (A_p != B_p)).sum())
        s = np.int16(np.uint16(O.dot_u16(pre_a, pre_b)))  # (pre_a * pre_b).sum() as i16
        assert equal - uneq == int(s)
        assert equal + uneq == denominator
        assert (denominator - int(s)) % 2 == 0
        assert uneq == (denominator - int(s)) // 2
        assert O.dot_bool(am, bm) == denominator


def test_dot_u16_is_truncated_u64_sum():
    # src/arch/sve.rs:84-108: dot == (sum of a*b as u64) as u16
    r = rng(7)
    for _ in range(20):
        a, b = rand_encoded(r), rand_encoded(r)
        exp = int((a.astype(np.uint64) * b.astype(np.uint64)).sum() & np.uint64(0xFFFF))
        assert O.dot_u16(a, b) == exp


def test_encrypted_equals_plaintext_distance():
    # src/lib.rs:165-193 asserts decode_distance(distances, denominators) == stored distance and
    # src/template.rs:101-112 asserts Template::distance == the same stored distance (<= 1 ulp);
    # the stored file is absent upstream, so check the two paths against each other.
    r = rng(8)
    for _ in range(25):
        qp, qm, ep, em = (rand_bits(r) for _ in range(4))
        d = O.distances(O.encode(qp, qm), O.encode(ep, em))
        n = O.denominators(qm, em)
        actual = O.decode_distance(d, n)
        expected = O.template_distance(qp, qm, ep, em)
        assert abs(actual - expected) <= np.spacing(expected)
        assert actual == O.np_decode_distance(d, n)


def test_decode_distance_ignores_nan():
    # den = 0 -> 0/0 = NaN, ignored by f64::min (src/lib.rs:106)
    d = np.zeros(31, np.uint16)
    n = np.zeros(31, np.uint16)
    assert O.decode_distance(d, n) == float("inf")
    n[3] = 100
    d[3] = 40
    assert O.decode_distance(d, n) == 30 / 100


def test_c_vs_numpy_batches():
    r = rng(9)
    db = r.integers(0, 2**16, size=(40, O.BITS), dtype=np.uint16)
    masks = r.integers(0, 2**64, size=(40, O.LIMBS), dtype=np.uint64)
    for q in (rand_encoded(r), O.encode(rand_bits(r), rand_bits(r))):
        assert np.array_equal(O.distance_batch(q, db, threads=2), O.np_distance_batch(q, db))
    qm = rand_bits(r)
    assert np.array_equal(O.masks_batch(qm, masks, threads=2), O.np_masks_batch(qm, masks))
    # engine rotations are -15..=15 in order (src/lib.rs:34-35, 61-62)
    rot = O.distance_rotations(db[0])
    for j in range(31):
        assert np.array_equal(rot[j], O.np_encoded_rotated(db[0], j - 15))


def test_share_reconstructs_and_distances_are_linear():
    # src/encoded_bits.rs:23-38 + the coordinator's wrapping sum (src/main.rs:603-608)
    r = rng(10)
    enc = O.encode(rand_bits(r), rand_bits(r))
    rest = r.integers(0, 2**16, size=(2, O.BITS), dtype=np.uint16)
    last = O.share_last(enc, rest)
    shares = np.concatenate([rest, last[None]])
    assert np.array_equal(shares.sum(axis=0, dtype=np.uint64).astype(np.uint16), enc)
    q = O.encode(rand_bits(r), rand_bits(r))
    parts = np.stack([O.distances(q, s) for s in shares])
    assert np.array_equal(parts.sum(axis=0, dtype=np.uint64).astype(np.uint16), O.distances(q, enc))


def test_combine_min_matches_plaintext_argmin():
    r = rng(11)
    n = 12
    qp, qm = rand_bits(r), rand_bits(r)
    q = O.encode(qp, qm)
    ep = r.integers(0, 2**64, size=(n, O.LIMBS), dtype=np.uint64)
    em = r.integers(0, 2**64, size=(n, O.LIMBS), dtype=np.uint64)
    ep[5] = qp  # plant a near match
    em[5] = qm
    enc = np.stack([O.encode(ep[i], em[i]) for i in range(n)])
    s0 = r.integers(0, 2**16, size=enc.shape, dtype=np.uint16)
    s1 = (enc - s0).astype(np.uint16)
    dist = np.stack([O.distance_batch(q, s0), O.distance_batch(q, s1)])
    den = O.masks_batch(qm, em)
    md, mi = O.combine_min(dist, den)
    plain = [O.template_distance(qp, qm, ep[i], em[i]) for i in range(n)]
    assert mi == int(np.argmin(plain)) == 5
    assert md == min(plain) == 0.0


def test_synthetic_rows_are_deterministic_and_uniformish():
    a = O.gen_share_rows(0x1715C0DE, 7, 3, threads=2)
    b = O.gen_share_rows(0x1715C0DE, 8, 1)
    assert np.array_equal(a[1], b[0])
    m = O.gen_mask_rows(0x1715C0DE, 7, 3)
    assert np.array_equal(m[2], O.gen_mask_rows(0x1715C0DE, 9, 1)[0])
    assert 0.49 < O.np_bits_to_bool(m).mean() < 0.51
    assert 32000 < a.mean() < 33500


def test_golden_fixture():
    """tests/golden/golden_small.json was produced by tests/golden/make_golden.py from the numpy
    definition; the C oracle must reproduce it bit for bit."""
    with open(os.path.join(GOLDEN, "golden_small.json")) as f:
        g = json.load(f)
    seed, n = g["seed"], g["n_rows"]
    db = O.gen_share_rows(seed, g["row0"], n)
    masks = O.gen_mask_rows(seed, g["row0"], n)
    for case in g["cases"]:
        if case["kind"] == "ternary":
            qp = O.gen_mask_rows(case["qseed"], 0, 1)[0]
            qm = O.gen_mask_rows(case["qseed"], 1, 1)[0]
            q = O.encode(qp, qm)
        else:
            q = O.gen_share_rows(case["qseed"], 0, 1)[0]
            qm = O.gen_mask_rows(case["qseed"], 1, 1)[0]
        assert np.array_equal(O.distance_batch(q, db), np.array(case["distances"], np.uint16))
        assert np.array_equal(O.masks_batch(qm, masks), np.array(case["denominators"], np.uint16))
    kn = g["rotated_number"]
    i = np.arange(O.BITS)
    secret = (((i // O.COLS) << 8) | (i % O.COLS)).astype(np.uint16)
    for a_str, head in kn.items():
        assert O.encoded_rotated(secret, int(a_str))[: len(head)].tolist() == head


def test_synthetic_party_shares_are_a_secret_sharing_of_the_encodings():
    """The synthetic database spec (restated in oracle/iris_oracle.c): row R is encode(pattern_R, mask_R) split like
    EncodedBits::share (src/encoded_bits.rs:23-38).  Shares sum to the encoding, single shares look uniform, and the
    coordinator's combine over the three parties' distances (src/main.rs:597-621) equals Template::distance in the
    clear (src/template.rs:43-64) -- on every row, so a planted near-duplicate is what a search returns."""
    seed, n, parties = 0x1715C0DE, 40, 3
    pats, masks = O.gen_pattern_rows(seed, 100, n), O.gen_mask_rows(seed, 100, n)
    enc = np.stack([O.encode(pats[i], masks[i]) for i in range(n)])
    shares = [O.gen_party_share_rows(seed, p, parties, 100, n, threads=2) for p in range(parties)]
    total = np.zeros_like(enc)
    for s in shares:
        total = (total + s).astype(np.uint16)
    assert np.array_equal(total, enc)
    assert np.array_equal(O.gen_party_share_rows(seed, 0, 1, 100, n), enc)
    for s in shares:
        assert 32000 < s.mean() < 33500 and len(np.unique(s[0])) > 5000
    # query: row 117 with 1 000 pattern bits flipped, rotated by +3 columns
    rng = np.random.default_rng(1)
    bits = O.np_bits_to_bool(pats[17]).copy()
    bits[rng.choice(O.BITS, size=1000, replace=False)] ^= 1
    qp, qm = O.bits_rotated(O.np_bool_to_bits(bits), 3), O.bits_rotated(masks[17], 3)
    q = O.encode(qp, qm)
    dist = np.stack([O.distance_batch(q, s) for s in shares])
    md, mi = O.combine_min(dist, O.masks_batch(qm, masks))
    plain = [O.template_distance(qp, qm, pats[i], masks[i]) for i in range(n)]
    assert mi == 17 == int(np.argmin(plain)) and md == plain[17] and md < 0.2
    for i in (0, 5, 39):
        one = O.decode_distance(((dist[0][i].astype(np.uint32) + dist[1][i] + dist[2][i]) & 0xFFFF).astype(np.uint16),
                                O.masks_batch(qm, masks[i:i + 1])[0])
        assert one == plain[i]
