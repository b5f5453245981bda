"""GPU parity tests: the CUDA path, called through the C ABI (ctypes wrapper), against the CPU oracle.

Bit-exact for everything (integer / bit work).  Tests re-create the reference's own tests
(file:line cited) with the GPU in place of arch::generic, plus the edge cases of the boundary.
"""
import json
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def iris():
    import mpc_iris_code_b200 as iris

    assert iris.device_count() >= 1
    return iris


@pytest.fixture(scope="module")
def small(iris):
    """1000-row resident shard loaded through the loader from reference-layout host rows."""
    n = 1000
    shares = O.gen_share_rows(SEED, 0, n, threads=8)
    masks = O.gen_mask_rows(SEED, 0, n, threads=8)
    db = iris.Database(n + 50)
    db.append_shares(shares[:300])       # ragged appends exercise partially filled tiles
    db.append_shares(shares[300:])
    db.append_masks(masks[:1])
    db.append_masks(masks[1:])
    yield db, shares, masks
    db.close()


def ternary_query(seed):
    p, m = O.gen_mask_rows(seed, 0, 1)[0], O.gen_mask_rows(seed, 1, 1)[0]
    return O.encode(p, m), p, m


def uniform_query(seed):
    return O.gen_share_rows(seed, 0, 1)[0]


def test_loader_roundtrip(small):
    db, shares, masks = small
    assert db.len_shares == 1000 and db.len_masks == 1000
    assert np.array_equal(db.read_shares(0, 1000), shares)
    assert np.array_equal(db.read_masks(0, 1000), masks)
    assert np.array_equal(db.read_shares(123, 7), shares[123:130])


def test_device_generator_matches_oracle(iris):
    with iris.Database(700) as g:
        g.generate(SEED, 5000, 300)
        g.generate(SEED, 5300, 400)
        assert np.array_equal(g.read_shares(0, 700), O.gen_share_rows(SEED, 5000, 700, threads=8))
        assert np.array_equal(g.read_masks(0, 700), O.gen_mask_rows(SEED, 5000, 700, threads=8))


def test_arch_dot_entry_points(iris, small):
    # src/arch/generic.rs:4-16; sve::test_dot_u16 (src/arch/sve.rs:84-108): dot == truncated u64 sum
    _, shares, masks = small
    for i in range(4):
        a, b = shares[i], shares[i + 10]
        assert iris.dot_u16(a, b) == O.dot_u16(a, b) == int((a.astype(np.uint64) * b).sum() & 0xFFFF)
        assert iris.dot_bool(masks[i], masks[i + 10]) == O.dot_bool(masks[i], masks[i + 10])
    full = np.full(O.BITS, 0xFFFF, np.uint16)
    assert iris.dot_u16(full, full) == O.dot_u16(full, full)
    ones = np.full(O.LIMBS, 2**64 - 1, np.uint64)
    assert iris.dot_bool(ones, ones) == 12800


@pytest.mark.parametrize("kind", ["uniform", "ternary"])
def test_distance_engine_resident(iris, small, kind):
    # DistanceEngine::batch_process (src/lib.rs:42-52)
    db, shares, _ = small
    q = uniform_query(31) if kind == "uniform" else ternary_query(32)[0]
    out = np.zeros((1000, 31), np.uint16)
    iris.DistanceEngine(q).batch_process(out, db)
    assert np.array_equal(out, O.distance_batch(q, shares, threads=8))


def test_masks_engine_resident(iris, small):
    # MasksEngine::batch_process (src/lib.rs:69-79)
    db, _, masks = small
    qm = O.gen_mask_rows(33, 1, 1)[0]
    out = np.zeros((1000, 31), np.uint16)
    iris.MasksEngine(qm).batch_process(out, db)
    assert np.array_equal(out, O.masks_batch(qm, masks, threads=8))


def test_fused_match_device_and_host_outputs(iris, small):
    import torch

    db, shares, masks = small
    q, _, qm = ternary_query(34)
    de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
    exp_d, exp_n = O.distance_batch(q, shares, threads=8), O.masks_batch(qm, masks, threads=8)
    hd, hn = np.zeros((1000, 31), np.uint16), np.zeros((1000, 31), np.uint16)
    iris.match(de, me, db, 0, 1000, hd, hn)
    assert np.array_equal(hd, exp_d) and np.array_equal(hn, exp_n)
    dd = torch.zeros((1000, 31), dtype=torch.int16, device="cuda")
    dn = torch.zeros((1000, 31), dtype=torch.int16, device="cuda")
    iris.match(de, me, db, 0, 1000, dd, dn)
    db.synchronize()
    assert np.array_equal(dd.cpu().numpy().view(np.uint16), exp_d) and np.array_equal(dn.cpu().numpy().view(np.uint16), exp_n)


@pytest.mark.parametrize("rb,re", [(0, 1), (0, 128), (0, 129), (127, 129), (1, 1000), (128, 256), (333, 777), (999, 1000)])
def test_row_ranges(iris, small, rb, re):
    # the reference calls batch_process on arbitrary chunks of its mmap (src/main.rs:428)
    db, shares, masks = small
    q, _, qm = ternary_query(35)
    de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
    guard = 0xABCD
    hd = np.full((re - rb + 2, 31), guard, np.uint16)
    hn = np.full((re - rb + 2, 31), guard, np.uint16)
    iris.match(de, me, db, rb, re, hd[1:-1], hn[1:-1])
    assert np.array_equal(hd[1:-1], O.distance_batch(q, shares[rb:re]))
    assert np.array_equal(hn[1:-1], O.masks_batch(qm, masks[rb:re]))
    assert (hd[0] == guard).all() and (hd[-1] == guard).all() and (hn[0] == guard).all() and (hn[-1] == guard).all()


def test_row_ranges_device_unaligned_output(iris, small):
    import torch

    db, shares, _ = small
    q = uniform_query(36)
    de = iris.DistanceEngine(q)
    exp = O.distance_batch(q, shares, threads=8)
    buf = torch.full((1000 * 31 + 64,), 0x7777, dtype=torch.int16, device="cuda")
    for off, rb, re in ((1, 5, 300), (3, 128, 1000), (7, 0, 1000)):
        buf.fill_(0x7777)
        view = buf[off : off + (re - rb) * 31]
        de.batch_process(view, db, rb, re)
        db.synchronize()
        got = buf.cpu().numpy().view(np.uint16)
        assert np.array_equal(got[off : off + (re - rb) * 31].reshape(-1, 31), exp[rb:re])
        assert (got[:off] == 0x7777).all() and (got[off + (re - rb) * 31 :] == 0x7777).all()


def test_reference_chunked_calling_pattern(iris):
    # participant: patterns.chunks(20_000) -> batch_process per chunk (src/main.rs:428-431)
    n = 45_000
    with iris.Database(n) as db:
        db.generate(SEED, 0, n)
        q, _, qm = ternary_query(37)
        de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
        whole_d, whole_n = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
        iris.match(de, me, db, 0, n, whole_d, whole_n)
        parts_d, parts_n = [], []
        for b in range(0, n, 20_000):
            e = min(n, b + 20_000)
            od, on = np.zeros((e - b, 31), np.uint16), np.zeros((e - b, 31), np.uint16)
            de.batch_process(od, db, b, e)
            me.batch_process(on, db, b, e)
            parts_d.append(od)
            parts_n.append(on)
        assert np.array_equal(np.concatenate(parts_d), whole_d)
        assert np.array_equal(np.concatenate(parts_n), whole_n)
        idx = np.array([0, 19_999, 20_000, 20_001, 44_999])
        for i in idx:
            assert np.array_equal(whole_d[i], O.distance_batch(q, O.gen_share_rows(SEED, int(i), 1))[0])
            assert np.array_equal(whole_n[i], O.masks_batch(qm, O.gen_mask_rows(SEED, int(i), 1))[0])


def test_literal_host_slice_api_and_single_pair(iris, small):
    # batch_process(&self, out, db: &[EncodedBits]) with host slices; distances()/denominators() (src/lib.rs:82-94)
    _, shares, masks = small
    q, _, qm = ternary_query(38)
    out = np.zeros((1000, 31), np.uint16)
    iris.DistanceEngine(q).batch_process(out, shares)
    assert np.array_equal(out, O.distance_batch(q, shares, threads=8))
    out = np.zeros((1000, 31), np.uint16)
    iris.MasksEngine(qm).batch_process(out, masks)
    assert np.array_equal(out, O.masks_batch(qm, masks, threads=8))
    assert np.array_equal(iris.distances(q, shares[5]), O.distances(q, shares[5]))
    assert np.array_equal(iris.denominators(qm, masks[5]), O.denominators(qm, masks[5]))


def test_errors_mirror_reference_asserts(iris, small):
    db, shares, masks = small
    q, _, qm = ternary_query(39)
    de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
    # assert_eq!(out.len(), db.len()) (src/lib.rs:43, 70) -> IRIS_ERR_INVALID
    with pytest.raises(iris.IrisError) as ei:
        de.batch_process(np.zeros((999, 31), np.uint16), shares)
    assert ei.value.code == -1
    with pytest.raises(iris.IrisError) as ei:
        me.batch_process(np.zeros((10, 31), np.uint16), db, 0, 11)
    assert ei.value.code == -1
    with pytest.raises(iris.IrisError):
        de.batch_process(np.zeros((10, 31), np.uint16), db, 995, 1005)   # beyond loaded rows
    with iris.Database(10, shares=True, masks=False) as only_shares:
        only_shares.generate(1, 0, 10)
        with pytest.raises(iris.IrisError) as ei:
            me.batch_process(np.zeros((10, 31), np.uint16), only_shares, 0, 10)
        assert ei.value.code == -4
    # empty slices are fine (rayon over zero rows)
    de.batch_process(np.zeros((0, 31), np.uint16), np.zeros((0, O.BITS), np.uint16))
    de.batch_process(np.zeros((0, 31), np.uint16), db, 5, 5)


def test_extreme_values(iris):
    rows = np.zeros((4, O.BITS), np.uint16)
    rows[0] = 0xFFFF
    rows[1] = 0x00FF
    rows[2] = 0xFF00
    rows[3] = np.arange(O.BITS, dtype=np.uint16) * 5
    masks = np.zeros((4, O.LIMBS), np.uint64)
    masks[0] = 2**64 - 1
    masks[2] = 0xAAAAAAAAAAAAAAAA
    masks[3] = 1
    with iris.Database(4) as db:
        db.append_shares(rows)
        db.append_masks(masks)
        for q in (np.full(O.BITS, 0xFFFF, np.uint16), np.full(O.BITS, 0x8000, np.uint16), np.zeros(O.BITS, np.uint16), np.full(O.BITS, 0x0101, np.uint16)):
            out = np.zeros((4, 31), np.uint16)
            iris.DistanceEngine(q).batch_process(out, db)
            assert np.array_equal(out, O.distance_batch(q, rows))
        for qm in (np.full(O.LIMBS, 2**64 - 1, np.uint64), np.zeros(O.LIMBS, np.uint64), np.full(O.LIMBS, 0x5555555555555555, np.uint64)):
            out = np.zeros((4, 31), np.uint16)
            iris.MasksEngine(qm).batch_process(out, db)
            assert np.array_equal(out, O.masks_batch(qm, masks))
        out = np.zeros((4, 31), np.uint16)
        iris.MasksEngine(np.full(O.LIMBS, 2**64 - 1, np.uint64)).batch_process(out, db)
        assert (out[0] == 12800).all() and (out[1] == 0).all()


def test_rotated_number_through_the_gpu(iris):
    # src/encoded_bits.rs:205-219: with q[row*200+col] = row<<8|col, rotated(a)[row][col] = row<<8|(col-a) mod 200.
    # A one-hot database row e_k makes distances[j] = rot(q, j-15)[k], so the GPU's rotation is read out directly.
    i = np.arange(O.BITS)
    secret = (((i // O.COLS) << 8) | (i % O.COLS)).astype(np.uint16)
    ks = [0, 1, 14, 15, 16, 199, 200, 215, 12799, 12800 - 200, 6407]
    rows = np.zeros((len(ks), O.BITS), np.uint16)
    for r, k in enumerate(ks):
        rows[r, k] = 1
    with iris.Database(len(ks), masks=False) as db:
        db.append_shares(rows)
        out = np.zeros((len(ks), 31), np.uint16)
        iris.DistanceEngine(secret).batch_process(out, db)
    for r, k in enumerate(ks):
        row, col = k // O.COLS, k % O.COLS
        for j in range(31):
            a = j - 15
            assert out[r, j] == ((row << 8) | ((O.COLS + col - a) % O.COLS)), (k, a)


def test_rotated_bits_through_the_gpu(iris):
    # src/encoded_bits.rs:221-236: Bits and EncodedBits rotate identically -> with a 0/1-embedded query and a
    # 0/1-embedded database row, distances == denominators of the underlying bit vectors.
    qbits = O.gen_mask_rows(41, 0, 1)[0]
    dbits = O.gen_mask_rows(41, 1, 6)
    with iris.Database(6) as db:
        db.append_shares(np.stack([O.encoded_from_bits(b) for b in dbits]))
        db.append_masks(dbits)
        d, n = np.zeros((6, 31), np.uint16), np.zeros((6, 31), np.uint16)
        iris.match(iris.DistanceEngine(O.encoded_from_bits(qbits)), iris.MasksEngine(qbits), db, 0, 6, d, n)
    assert np.array_equal(d, n)
    assert np.array_equal(n, O.masks_batch(qbits, dbits))


def test_dotproduct_identity_and_decode_through_the_gpu(iris):
    # src/lib.rs:134-163 (identity) and :165-193 (encrypted == plaintext distance, <= 1 ulp)
    n = 64
    ep = O.gen_mask_rows(42, 0, n)
    em = O.gen_mask_rows(43, 0, n)
    qp, qm = O.gen_mask_rows(44, 0, 1)[0], O.gen_mask_rows(44, 1, 1)[0]
    enc = np.stack([O.encode(ep[i], em[i]) for i in range(n)])
    with iris.Database(n) as db:
        db.append_shares(enc)
        db.append_masks(em)
        d, den = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
        iris.match(iris.DistanceEngine(O.encode(qp, qm)), iris.MasksEngine(qm), db, 0, n, d, den)
    diff = (den.astype(np.int32) - d.astype(np.int16).astype(np.int32))
    assert (diff % 2 == 0).all() and (diff >= 0).all()
    for i in range(n):
        expected = O.template_distance(qp, qm, ep[i], em[i])
        actual = O.decode_distance(d[i], den[i])
        assert abs(actual - expected) <= np.spacing(expected)


def test_share_linearity_through_the_gpu(iris):
    # additive sharing (src/encoded_bits.rs:23-38): distances of the shares sum (wrapping) to distances of the secret
    r = np.random.default_rng(45)
    n = 40
    enc = np.stack([O.encode(O.gen_mask_rows(46, i, 1)[0], O.gen_mask_rows(47, i, 1)[0]) for i in range(n)])
    s0 = r.integers(0, 2**16, size=enc.shape, dtype=np.uint16)
    s1 = r.integers(0, 2**16, size=enc.shape, dtype=np.uint16)
    s2 = (enc - s0 - s1).astype(np.uint16)
    q = ternary_query(48)[0]
    outs = []
    for s in (s0, s1, s2, enc):
        with iris.Database(n, masks=False) as db:
            db.append_shares(s)
            o = np.zeros((n, 31), np.uint16)
            iris.DistanceEngine(q).batch_process(o, db)
            outs.append(o)
    assert np.array_equal((outs[0] + outs[1] + outs[2]).astype(np.uint16), outs[3])
    assert np.array_equal(outs[3], O.distance_batch(q, enc))


def test_golden_fixture_through_the_gpu(iris):
    with open(os.path.join(GOLDEN, "golden_small.json")) as f:
        g = json.load(f)
    n = g["n_rows"]
    with iris.Database(n) as db:
        db.generate(g["seed"], g["row0"], n)
        for case in g["cases"]:
            qm = O.gen_mask_rows(case["qseed"], 1, 1)[0]
            q = O.encode(O.gen_mask_rows(case["qseed"], 0, 1)[0], qm) if case["kind"] == "ternary" else O.gen_share_rows(case["qseed"], 0, 1)[0]
            d, den = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
            iris.match(iris.DistanceEngine(q), iris.MasksEngine(qm), db, 0, n, d, den)
            assert np.array_equal(d, np.array(case["distances"], np.uint16))
            assert np.array_equal(den, np.array(case["denominators"], np.uint16))


def test_full_size_scan_properties(iris):
    """BASELINE config 2 size (1 M rows, 27.2 GB): the tensor-core scan must equal (i) the independent
    CUDA-core kernel on every row (compared on the device) and (ii) the CPU oracle on sampled rows
    regenerated from the counter-based generator; host (chunk-pipelined) and device outputs must agree."""
    import torch

    n = 1_000_000
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("not enough free HBM for the 1 M-row shard")
    with iris.Database(n) as db:
        db.generate(SEED, 0, n)
        q, _, qm = ternary_query(50)
        de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
        dd = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        dn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        iris.match(de, me, db, 0, n, dd, dn)
        db.synchronize()
        cd = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        cn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        db.check_distances_simt(q, 0, n, cd)
        db.check_denominators_simt(qm, 0, n, cn)
        assert torch.equal(dd, cd)
        assert torch.equal(dn, cn)
        hd, hn = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
        iris.match(de, me, db, 0, n, hd, hn)
        assert np.array_equal(hd, dd.cpu().numpy().view(np.uint16)) and np.array_equal(hn, dn.cpu().numpy().view(np.uint16))
        r = np.random.default_rng(51)
        idx = np.concatenate([[0, 127, 128, n - 129, n - 1], r.integers(0, n, 300)])
        for i in idx:
            assert np.array_equal(hd[i], O.distance_batch(q, O.gen_share_rows(SEED, int(i), 1))[0]), i
            assert np.array_equal(hn[i], O.masks_batch(qm, O.gen_mask_rows(SEED, int(i), 1))[0]), i
        # uniform-u16 query (what the reference's criterion bench feeds dot_u16, src/arch/mod.rs:56-58)
        qu = uniform_query(52)
        iris.DistanceEngine(qu).batch_process(dd, db)
        db.synchronize()
        db.check_distances_simt(qu, 0, n, cd)
        assert torch.equal(dd, cd)


# ----------------------------------------------------------------------------------- batched queries (config 4)
@pytest.mark.parametrize("kind,nq", [("ternary", 8), ("uniform", 8), ("ternary", 5), ("mixed", 20), ("ternary", 1), ("uniform", 13)])
def test_distances_batch_matches_per_engine_and_oracle(iris, small, kind, nq):
    """iris_distances_batch_resident (2-CTA tcgen05 GEMM) == batch_process per engine == oracle.
    Ternary batches take the signed two-product path, any non-s8 query forces the three-product path."""
    db, shares, _ = small
    qs = []
    for i in range(nq):
        tern = kind == "ternary" or (kind == "mixed" and i % 3)
        qs.append(ternary_query(700 + i)[0] if tern else uniform_query(800 + i))
    engines = [iris.DistanceEngine(q) for q in qs]
    exp = np.stack([O.distance_batch(q, shares, threads=8) for q in qs])
    for rb, re in ((0, 1000), (3, 997), (256, 512), (511, 513)):
        out = np.full((nq, re - rb, 31), 0x5A5A, np.uint16)
        iris.distances_batch(engines, db, rb, re, out)
        assert np.array_equal(out, exp[:, rb:re]), (kind, nq, rb, re)
    single = np.zeros((1000, 31), np.uint16)
    engines[-1].batch_process(single, db)
    assert np.array_equal(single, exp[-1])


def test_distances_batch_device_output_and_s8_edge_values(iris, small):
    import torch

    db, shares, _ = small
    # values at the edge of the signed-byte fast path: 0x7F / 0xFF80 are s8, 0x80 / 0xFF7F are not
    q_s8 = np.tile(np.array([0x7F, 0xFF80, 0, 1, 0xFFFF], np.uint16), O.BITS // 5)
    q_u = q_s8.copy()
    q_u[7] = 0x80
    q_v = q_s8.copy()
    q_v[9] = 0xFF7F
    for qs in ([q_s8] * 3, [q_s8, q_u], [q_v]):
        engines = [iris.DistanceEngine(q) for q in qs]
        out = torch.zeros((len(qs), 1000, 31), dtype=torch.int16, device="cuda")
        iris.distances_batch(engines, db, 0, 1000, out)
        db.synchronize()
        got = out.cpu().numpy().view(np.uint16)
        for i, q in enumerate(qs):
            assert np.array_equal(got[i], O.distance_batch(q, shares, threads=8))


def test_distances_batch_errors(iris, small):
    db, _, _ = small
    e = iris.DistanceEngine(uniform_query(900))
    with pytest.raises(iris.IrisError):
        iris.distances_batch([e], db, 0, 2000, np.zeros((1, 2000, 31), np.uint16))     # beyond loaded rows
    with pytest.raises(ValueError):
        iris.distances_batch([e, e], db, 0, 10, np.zeros((1, 10, 31), np.uint16))      # out too small
    iris.distances_batch([e], db, 5, 5, np.zeros((1, 0, 31), np.uint16))               # empty range is fine


@pytest.mark.parametrize("nq", [16, 1, 5, 37])
def test_denominators_batch_matches_per_engine_and_oracle(iris, small, nq):
    """iris_denominators_batch_resident (2-CTA tcgen05 GEMM over in-SM expanded mask bits) == oracle."""
    db, _, masks = small
    qms = [O.gen_mask_rows(1000 + i, 1, 1)[0] for i in range(nq)]
    qms[0] = np.full(O.LIMBS, 2**64 - 1, np.uint64)          # all-ones mask: denominators = popcount of the row
    engines = [iris.MasksEngine(q) for q in qms]
    exp = np.stack([O.masks_batch(q, masks, threads=8) for q in qms])
    for rb, re in ((0, 1000), (3, 997), (256, 512), (511, 513)):
        out = np.full((nq, re - rb, 31), 0x5A5A, np.uint16)
        iris.denominators_batch(engines, db, rb, re, out)
        assert np.array_equal(out, exp[:, rb:re]), (nq, rb, re)
    single = np.zeros((1000, 31), np.uint16)
    engines[-1].batch_process(single, db)
    assert np.array_equal(single, exp[-1])


# ----------------------------------------------------------------------------------- coordinator reduction (f-1)
def test_combine_min_matches_reference_coordinator(iris, small):
    """iris_combine_min == wrapping sum of party shares + decode_distance + running min (src/main.rs:597-621,
    src/lib.rs:97-107), bit-exact in f64, first minimum wins."""
    r = np.random.default_rng(60)
    n = 600
    qp, qm = O.gen_mask_rows(61, 0, 1)[0], O.gen_mask_rows(61, 1, 1)[0]
    ep, em = O.gen_mask_rows(62, 0, n), O.gen_mask_rows(63, 0, n)
    ep[417], em[417] = qp, qm                      # exact match
    ep[99], em[99] = qp, qm                        # equal distance earlier: must win
    em[5] = 0                                      # fully masked row: den = 0 -> NaN / inf handling
    enc = np.stack([O.encode(ep[i], em[i]) for i in range(n)])
    s0 = r.integers(0, 2**16, size=enc.shape, dtype=np.uint16)
    s1 = r.integers(0, 2**16, size=enc.shape, dtype=np.uint16)
    s2 = (enc - s0 - s1).astype(np.uint16)
    q = O.encode(qp, qm)
    shares = [O.distance_batch(q, s, threads=8) for s in (s0, s1, s2)]
    den = O.masks_batch(qm, em, threads=8)
    exp_min, exp_idx = O.combine_min(np.stack(shares), den)
    md, mi, dist = iris.combine_min(shares, den, want_distances=True)
    assert (md, mi) == (exp_min, exp_idx) == (0.0, 99)
    exp_dist = np.array([O.decode_distance((shares[0][i] + shares[1][i] + shares[2][i]).astype(np.uint16), den[i]) for i in range(n)])
    assert np.array_equal(dist, exp_dist)          # bit-exact f64, including the +inf of the masked-out row
    assert dist[5] == np.inf
    assert iris.combine_min(shares, den, index_base=1000) == (0.0, 1099)
    # nothing below +inf: the reference leaves min_index = usize::MAX
    z = np.zeros((4, 31), np.uint16)
    assert iris.combine_min([z], z) == (np.inf, -1)


def test_decode_on_arbitrary_u16_pairs_is_bit_identical(iris):
    """decode_distance (src/lib.rs:97-107) on the device picks the smallest FRACTION exactly and divides once; the
    reference divides 31 times and folds f64::min.  Same bits for any u16 inputs: uniform pairs, denominators of zero
    (NaN / +inf rotations), near-equal fractions with large denominators, and rows where every rotation is NaN."""
    r = np.random.default_rng(2026)
    n = 200_000
    dist = r.integers(0, 2**16, size=(n, 31), dtype=np.uint16)
    den = r.integers(0, 2**16, size=(n, 31), dtype=np.uint16)
    den[r.random((n, 31)) < 0.05] = 0                                   # scattered zero denominators
    den[:50] = 0                                                         # rows with no finite quotient at all
    dist[:25] = 0                                                        # ... 0/0 = NaN everywhere -> +inf
    big = r.integers(65000, 65536, size=(1000, 31), dtype=np.uint16)    # neighbouring fractions near 1/2
    den[1000:2000] = big
    dist[1000:2000] = (big.astype(np.int64) - 2 * (big.astype(np.int64) // 2 - r.integers(0, 3, size=big.shape))).astype(np.uint16)
    den[3000:3100] = 1                                                   # quotients far above 1
    md, mi, got = iris.combine_min([dist], den, want_distances=True)
    exp = np.array([O.decode_distance(dist[i], den[i]) for i in range(0, n, 7)])
    assert np.array_equal(got[::7], exp)
    exp_head = np.array([O.decode_distance(dist[i], den[i]) for i in range(3200)])
    assert np.array_equal(got[:3200], exp_head) and np.isinf(got[:50]).all()
    assert (md, mi) == O.combine_min(dist[None], den)


def test_match_min_fused_scan_and_reduction(iris):
    n = 3000
    qp, qm = O.gen_mask_rows(71, 0, 1)[0], O.gen_mask_rows(71, 1, 1)[0]
    ep, em = O.gen_mask_rows(72, 0, n), O.gen_mask_rows(73, 0, n)
    noisy = qp.copy()
    noisy[:3] ^= np.uint64(0xFFFF)                 # near match at row 2222
    ep[2222], em[2222] = noisy, qm
    enc = np.stack([O.encode(ep[i], em[i]) for i in range(n)])
    q = O.encode(qp, qm)
    with iris.Database(n) as db:
        db.append_shares(enc)
        db.append_masks(em)
        de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
        exp = O.combine_min(O.distance_batch(q, enc, threads=8)[None], O.masks_batch(qm, em, threads=8))
        assert iris.match_min(de, me, db, 0, n) == exp
        assert exp[1] == 2222
        exp_sub = O.combine_min(O.distance_batch(q, enc[100:2000], threads=8)[None], O.masks_batch(qm, em[100:2000], threads=8))
        assert iris.match_min(de, me, db, 100, 2000, index_base=50_000) == (exp_sub[0], 50_000 + 100 + exp_sub[1])
        plain = min(O.template_distance(qp, qm, ep[i], em[i]) for i in (2222, 7, 1234))
        assert abs(exp[0] - plain) <= np.spacing(plain)


# ----------------------------------------------------------------------------------- flat-file loader (f-2)
def test_flat_file_loader_reads_reference_formats(iris, tmp_path):
    """mpc.share-i = raw EncodedBits rows, mpc.masks = raw Bits rows (src/main.rs:337-371, 386-400, 458-461)."""
    n = 2500                                        # > one 1024-row staging chunk: exercises the double buffer
    shares = O.gen_share_rows(SEED, 0, n, threads=8)
    masks = O.gen_mask_rows(SEED, 0, n, threads=8)
    sp, mp_ = tmp_path / "mpc.share-0", tmp_path / "mpc.masks"
    shares.tofile(sp)
    masks.tofile(mp_)
    with iris.Database(n + 10) as db:
        db.load_shares_file(str(sp))
        db.load_masks_file(str(mp_), 0, 1000)
        db.load_masks_file(str(mp_), 1000)           # the rest
        assert (db.len_shares, db.len_masks) == (n, n)
        assert np.array_equal(db.read_shares(0, n), shares)
        assert np.array_equal(db.read_masks(0, n), masks)
        q, _, qm = ternary_query(77)
        d, den = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
        iris.match(iris.DistanceEngine(q), iris.MasksEngine(qm), db, 0, n, d, den)
        assert np.array_equal(d, O.distance_batch(q, shares, threads=8))
        assert np.array_equal(den, O.masks_batch(qm, masks, threads=8))
    with iris.Database(100) as db:
        db.load_shares_file(str(sp), 2400, 100)      # a row window of the file
        assert np.array_equal(db.read_shares(0, 100), shares[2400:])
        with pytest.raises(iris.IrisError):
            db.load_shares_file(str(sp), 0, 200)     # exceeds capacity (100 rounds up to one 256-row pair tile)
        with pytest.raises(iris.IrisError):
            db.load_shares_file(str(tmp_path / "missing"))
    bad = tmp_path / "truncated"
    bad.write_bytes(shares.tobytes()[:-7])           # not a whole number of rows: "Share file invalid"
    with iris.Database(n) as db:
        with pytest.raises(iris.IrisError) as ei:
            db.load_shares_file(str(bad))
        assert ei.value.code == -1


# ----------------------------------------------------------------------------------- device-side encode (f-5) + wire shim (f-3)
def test_encode_on_device_and_engine_from_template(iris, small):
    # src/lib.rs:16-26 + test_preprocess (src/lib.rs:117-132)
    db, shares, _ = small
    for seed in range(80, 84):
        p, m = O.gen_mask_rows(seed, 0, 1)[0], O.gen_mask_rows(seed, 1, 1)[0]
        enc = iris.encode(p, m)
        assert np.array_equal(enc, O.encode(p, m))
        assert set(np.unique(enc)) <= {0, 1, 0xFFFF}
    out = np.zeros((1000, 31), np.uint16)
    iris.DistanceEngine.from_template(p, m).batch_process(out, db)
    assert np.array_equal(out, O.distance_batch(O.encode(p, m), shares, threads=8))


def test_participant_wire_protocol(iris, tmp_path):
    """The C++ front-end speaks the reference participant protocol (src/main.rs:411-446): 3 200-byte Template in,
    62 bytes per row out in batches, EOF at the end.  Played here by a Python 'coordinator'."""
    import socket
    import subprocess
    import time

    from mpc_iris_code_b200 import build

    n = 2300
    shares = O.gen_share_rows(SEED, 0, n, threads=8)
    path = tmp_path / "mpc.share-0"
    shares.tofile(path)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    exe = build.PARTICIPANT_PATH
    if not os.path.exists(exe):
        build.build_participant()
    proc = subprocess.Popen([exe, "--input", str(path), "--bind", f"127.0.0.1:{port}", "--batch-size", "1000", "--max-requests", "2"],
                            stderr=subprocess.PIPE, text=True)
    try:
        deadline = time.time() + 120
        while True:                                   # wait for "Listening on"
            line = proc.stderr.readline()
            if "Listening on" in line:
                break
            assert proc.poll() is None and time.time() < deadline, line
        for seed in (90, 91):
            p, m = O.gen_mask_rows(seed, 0, 1)[0], O.gen_mask_rows(seed, 1, 1)[0]
            with socket.create_connection(("127.0.0.1", port), timeout=60) as c:
                c.sendall(p.tobytes() + m.tobytes())  # Template {pattern, mask} (src/template.rs:26-29)
                chunks = []
                while True:
                    b = c.recv(1 << 20)
                    if not b:
                        break
                    chunks.append(b)
            got = np.frombuffer(b"".join(chunks), np.uint16).reshape(-1, 31)
            assert got.shape == (n, 31)
            assert np.array_equal(got, O.distance_batch(O.encode(p, m), shares, threads=8))
        assert proc.wait(timeout=60) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
