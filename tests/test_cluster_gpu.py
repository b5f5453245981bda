"""GPU tests of the row-sharded cluster, the secret-shared synthetic database, and the arch-level grids.

The cluster is exercised on ONE GPU with several shards on device 0 (every code path but the NVLink hop), and on two
or more GPUs when the box has them (peer stores, result arrays on another GPU, the NCCL join of several processes).
Everything is compared with the CPU oracle, bit for bit.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def iris():
    import mpc_iris_code_b200 as iris

    assert iris.device_count() >= 1
    return iris


def synthetic_template(row):
    """(pattern, mask) of synthetic row id `row` of the generated database (oracle restatement of the data spec)."""
    return O.gen_pattern_rows(SEED, row, 1)[0], O.gen_mask_rows(SEED, row, 1)[0]


def noisy_query(row, flips=1300, rotation=2, rng_seed=5):
    """A wire Template close to synthetic row `row`: `flips` pattern bits toggled, then rotated by `rotation` columns."""
    p, m = synthetic_template(row)
    rng = np.random.default_rng(rng_seed + row)
    bits = O.np_bits_to_bool(p).copy()
    idx = rng.choice(O.BITS, size=flips, replace=False)
    bits[idx] ^= 1
    p = O.np_bool_to_bits(bits)
    return O.bits_rotated(p, rotation), O.bits_rotated(m, rotation)


def plaintext_min(qp, qm, rows):
    """The coordinator's answer over synthetic rows `rows` computed in the clear: Template::distance
    (src/template.rs:43-64) per row, running min with `<` (src/main.rs:611-621)."""
    best, arg = np.inf, -1
    for r in rows:
        p, m = synthetic_template(int(r))
        d = O.template_distance(qp, qm, p, m)
        if d < best:
            best, arg = d, int(r)
    return best, arg


# ---------------------------------------------------------------------------------- secret-shared synthetic rows
def test_generated_party_shares_match_oracle_and_sum_to_the_encoding(iris):
    # EncodedBits::share (src/encoded_bits.rs:23-38), encode (src/lib.rs:16-26)
    n, row0, parties = 300, 7000, 3
    total = np.zeros((n, O.BITS), np.uint16)
    for p in range(parties):
        with iris.Database(n) as db:
            db.generate_shares(SEED, p, parties, row0, n)
            got = db.read_shares(0, n)
            assert np.array_equal(got, O.gen_party_share_rows(SEED, p, parties, row0, n, threads=8))
            assert np.array_equal(db.read_masks(0, n), O.gen_mask_rows(SEED, row0, n, threads=8))
            total += got
    pats, masks = O.gen_pattern_rows(SEED, row0, n), O.gen_mask_rows(SEED, row0, n)
    enc = np.stack([O.encode(pats[i], masks[i]) for i in range(n)])
    assert np.array_equal(total, enc)
    with iris.Database(n) as db:          # n = 1: the plaintext encodings
        db.generate_shares(SEED, 0, 1, row0, n)
        assert np.array_equal(db.read_shares(0, n), enc)
    with iris.Database(n) as db:
        with pytest.raises(iris.IrisError):
            db.generate_shares(SEED, 3, 3, row0, n)


def test_rows_can_be_overwritten(iris):
    n = 600
    with iris.Database(n) as db:
        db.generate(SEED, 0, n)
        new_s = O.gen_share_rows(99, 0, 200)
        new_m = O.gen_mask_rows(99, 0, 200)
        db.write_shares(250, new_s)
        db.write_masks(399, new_m)
        exp_s, exp_m = O.gen_share_rows(SEED, 0, n), O.gen_mask_rows(SEED, 0, n)
        exp_s[250:450] = new_s
        exp_m[399:599] = new_m
        assert np.array_equal(db.read_shares(0, n), exp_s) and np.array_equal(db.read_masks(0, n), exp_m)
        with pytest.raises(iris.IrisError):
            db.write_shares(500, new_s)


def test_three_party_shares_of_a_large_database_find_the_planted_template(iris):
    # the protocol of src/main.rs:419-431 (participants) and :597-621 (coordinator) on 200 000 shared rows
    import torch

    n, parties, target = 200_000, 3, 199_999
    qp, qm = noisy_query(target)
    q = O.encode(qp, qm)
    outs = []
    for p in range(parties):
        with iris.Database(n, masks=(p == 0)) as db:
            db.generate_shares(SEED, p, parties, 0, n)
            d = torch.empty((n, 31), dtype=torch.int16, device="cuda")
            iris.DistanceEngine(q).batch_process(d, db)
            if p == 0:
                den = torch.empty((n, 31), dtype=torch.int16, device="cuda")
                iris.MasksEngine(qm).batch_process(den, db)
            db.synchronize()
            outs.append(d)
    md, mi = iris.combine_min(outs, den)
    p_t, m_t = synthetic_template(target)
    assert (md, mi) == (O.template_distance(qp, qm, p_t, m_t), target)
    assert md < 0.2
    # one party alone learns nothing: its "distances" do not point at the target
    md1, mi1 = iris.combine_min(outs[:1], den)
    assert mi1 != target


# ---------------------------------------------------------------------------------- arch-level grids
@pytest.mark.parametrize("n_a", [1, 31, 40])
def test_dot_grids_match_oracle(iris, n_a):
    # the criterion grid of src/arch/mod.rs:22-72: every pair of independent vectors
    n_b = 700
    a16, b16 = O.gen_share_rows(41, 0, n_a), O.gen_share_rows(42, 0, n_b)
    exp = np.array([[O.dot_u16(a16[j], b16[i]) for j in range(n_a)] for i in range(0, n_b, 37)], np.uint16)
    got = iris.dot_u16_batch(a16, b16)
    assert got.shape == (n_b, n_a) and np.array_equal(got[::37], exp)
    # all rows through the engine identity: column j = rotation-0 distances of query a[j]
    assert np.array_equal(got[:, 0], O.distance_batch(a16[0], b16, threads=8)[:, 15])
    am, bm = O.gen_mask_rows(43, 0, n_a), O.gen_mask_rows(44, 0, n_b)
    gotm = iris.dot_bool_batch(am, bm)
    expm = np.array([[O.dot_bool(am[j], bm[i]) for j in range(n_a)] for i in range(n_b)], np.uint16)
    assert np.array_equal(gotm, expm)


def test_dot_grid_ternary_vectors_and_resident_rows(iris):
    import torch

    n_a, n_b = 31, 1000
    pats, masks = O.gen_pattern_rows(51, 0, n_a), O.gen_mask_rows(51, 0, n_a)
    a = np.stack([O.encode(pats[j], masks[j]) for j in range(n_a)])       # sign-extended bytes: the two-product path
    with iris.Database(n_b) as db:
        db.generate(SEED, 0, n_b)
        b = O.gen_share_rows(SEED, 0, n_b, threads=8)
        exp = np.stack([O.distance_batch(a[j], b, threads=8)[:, 15] for j in range(n_a)], axis=1)
        assert np.array_equal(iris.dot_u16_batch(a, db), exp)
        out = torch.zeros((n_b, n_a), dtype=torch.int16, device="cuda")
        iris.dot_u16_batch(a, db, out=out)
        db.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint16), exp)
        bm = O.gen_mask_rows(SEED, 0, n_b, threads=8)
        expm = np.stack([O.masks_batch(masks[j], bm, threads=8)[:, 15] for j in range(n_a)], axis=1)
        assert np.array_equal(iris.dot_bool_batch(masks, db), expm)


# ---------------------------------------------------------------------------------- cluster
def cluster_devices(iris, want):
    n = iris.device_count()
    return [i % n for i in range(want)]


@pytest.mark.parametrize("shards", [1, 3])
def test_cluster_blocks_and_match_equal_one_shard(iris, shards):
    import torch

    n = 5003
    q, _, qm = O.gen_share_rows(61, 0, 1)[0], None, O.gen_mask_rows(62, 0, 1)[0]
    with iris.Cluster(cluster_devices(iris, shards), n) as c, iris.Database(n) as ref:
        c.generate(SEED, n, first_row_id=100, n_parties=0)
        ref.generate(SEED, 100, n)
        assert len(c) == n
        covered = 0
        for i in range(shards):
            db, dev, b, e = c.shard(i)
            assert (b, e) == iris.cluster_partition(n, shards, i) and b == covered
            assert np.array_equal(db.read_shares(0, min(5, e - b)), ref.read_shares(b, min(5, e - b)))
            covered = e
        assert covered == n
        exp_d, exp_n = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
        iris.match(iris.DistanceEngine(q), iris.MasksEngine(qm), ref, 0, n, exp_d, exp_n)
        sample = [0, 1, n // 3, n - 1]
        assert np.array_equal(exp_d[sample], O.distance_batch(q, O.gen_share_rows(SEED, 100, n)[sample]))
        hd, hn = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
        c.match(q, qm, hd, hn)                                   # host outputs
        assert np.array_equal(hd, exp_d) and np.array_equal(hn, exp_n)
        for dev in sorted(set(c.devices)):                       # one array on each GPU of the cluster in turn
            dd = torch.zeros((n, 31), dtype=torch.int16, device=f"cuda:{dev}")
            dn = torch.zeros((n, 31), dtype=torch.int16, device=f"cuda:{dev}")
            c.match(q, qm, dd, dn)
            assert np.array_equal(dd.cpu().numpy().view(np.uint16), exp_d)
            assert np.array_equal(dn.cpu().numpy().view(np.uint16), exp_n)
        hd[:] = 0
        c.match(q, None, hd, None)                               # distances only (the participant)
        assert np.array_equal(hd, exp_d)
        hn[:] = 0
        c.match(None, qm, None, hn)                              # denominators only (the coordinator)
        assert np.array_equal(hn, exp_n)
        with pytest.raises(iris.IrisError):
            c.match(q, qm, None, hn)


def test_cluster_match_template_is_the_participants_request(iris):
    n = 2000
    qp, qm = noisy_query(1234)
    with iris.Cluster(cluster_devices(iris, 2), n) as c:
        c.generate(SEED, n, n_parties=2, party=1)
        hd = np.zeros((n, 31), np.uint16)
        c.match_template(qp, qm, hd)
        assert np.array_equal(hd, O.distance_batch(O.encode(qp, qm), O.gen_party_share_rows(SEED, 1, 2, 0, n, threads=8), threads=8))


@pytest.mark.parametrize("shards,queries", [(1, 1), (3, 1), (3, 5), (2, 9)])
def test_cluster_search_finds_planted_templates(iris, shards, queries):
    n = 3001
    targets = [n - 1, 0, 1700, 1000, 2999, 5, 2000, 1001, 77][:queries]       # the first winner lives in the LAST shard
    tq = np.zeros((queries, 400), np.uint64)
    for k, t in enumerate(targets):
        tq[k, :200], tq[k, 200:] = noisy_query(t, flips=900 + 100 * k, rotation=(k % 7) - 3)
    with iris.Cluster(cluster_devices(iris, shards), n) as c:
        c.generate(SEED, n, n_parties=1)
        md, mi = c.search(tq)
        for k, t in enumerate(targets):
            p_t, m_t = synthetic_template(t)
            assert mi[k] == t
            assert md[k] == O.template_distance(tq[k, :200].copy(), tq[k, 200:].copy(), p_t, m_t)
        # a query that resembles nothing: the full scan in the clear agrees (ties and all)
        rnd = np.random.default_rng(3).integers(0, 2**64, size=(1, 400), dtype=np.uint64)
        md, mi = c.search(rnd)
        assert (md[0], mi[0]) == plaintext_min(rnd[0, :200].copy(), rnd[0, 200:].copy(), range(n))
        c.set_index_base(10**12)
        md2, mi2 = c.search(rnd)
        assert md2[0] == md[0] and mi2[0] == mi[0] + 10**12


def test_cluster_search_ties_go_to_the_lowest_row(iris):
    # the reference keeps the first minimum (`distance < min_distance`, src/main.rs:617)
    n = 1500
    shares, masks = O.gen_party_share_rows(SEED, 0, 1, 0, n, threads=8), O.gen_mask_rows(SEED, 0, n, threads=8)
    for dup in (1499, 900, 300):           # the same template in three shards
        shares[dup], masks[dup] = shares[20], masks[20]
    qp, qm = noisy_query(20)
    with iris.Cluster(cluster_devices(iris, 3), n) as c:
        c.load_rows(shares, masks)
        md, mi = c.search(np.concatenate([qp, qm])[None, :].copy())
        assert mi[0] == 20
        # an empty mask: every denominator is zero, nothing is below +inf (min_index stays usize::MAX, src/main.rs:582)
        md, mi = c.search(np.zeros((1, 400), np.uint64))
        assert mi[0] == -1 and md[0] == np.inf


def test_search_mode_ties_inside_one_cta_go_to_the_lowest_row(iris):
    # one shard, 313 tiles on 148 CTAs: CTA 0 sees tiles 0, 148 and 296, and lane 5 of its epilogue sees row 5 of each.
    # The same template at those three rows (and, for good measure, in a neighbouring CTA): the first one must win
    # (`distance < min_distance`, src/main.rs:617), and when it is removed the next one.
    n = 40_000
    rows = [5, 5 + 148 * 128, 5 + 296 * 128, 149 * 128 + 77]
    tp, tm = synthetic_template(123_456)
    enc = O.encode(tp, tm)[None, :].copy()
    q = np.concatenate([tp, tm])[None, :].copy()
    with iris.Cluster([0], n) as c:
        c.generate(SEED, n, n_parties=1)
        db = c.shard(0)[0]
        for r in rows[1:]:
            db.write_shares(r, enc)
            db.write_masks(r, tm[None, :].copy())
        md, mi = c.search(q)
        assert (md[0], mi[0]) == (0.0, rows[1])
        db.write_shares(rows[0], enc)
        db.write_masks(rows[0], tm[None, :].copy())
        md, mi = c.search(q)
        assert (md[0], mi[0]) == (0.0, rows[0])
        de, me = iris.DistanceEngine.from_template(tp, tm), iris.MasksEngine(tm)
        assert iris.match_min(de, me, db, 6, n) == (0.0, rows[1])             # a range that starts behind the first copy
        assert iris.match_min(de, me, db, rows[1] + 1, n, index_base=7) == (0.0, 7 + rows[3])


def test_cluster_edge_shapes(iris):
    # fewer rows than shards (empty blocks), an empty database, more queries than one pass holds, zero queries
    with iris.Cluster(cluster_devices(iris, 3), 900) as c:
        c.generate(SEED, 2, n_parties=1)                      # blocks of 1, 1 and 0 rows
        assert [c.shard(i)[3] - c.shard(i)[2] for i in range(3)] == [1, 1, 0]
        qp, qm = noisy_query(1, flips=500, rotation=0)
        md, mi = c.search(np.concatenate([qp, qm])[None, :].copy())
        p1, m1 = synthetic_template(1)
        assert (md[0], mi[0]) == (O.template_distance(qp, qm, p1, m1), 1)
        out = np.zeros((2, 31), np.uint16)
        c.match_template(qp, qm, out)
        assert np.array_equal(out, O.distance_batch(O.encode(qp, qm), O.gen_party_share_rows(SEED, 0, 1, 0, 2)))
        c.generate(SEED, 0, n_parties=1)                      # nothing at all
        md, mi = c.search(np.concatenate([qp, qm])[None, :].copy())
        assert mi[0] == -1 and md[0] == np.inf
        md, mi = c.search(np.zeros((5, 400), np.uint64))
        assert (mi == -1).all() and np.isinf(md).all()
        c.match_template(qp, qm, np.zeros((0, 31), np.uint16))
        md, mi = c.search(np.zeros((0, 400), np.uint64))
        assert len(md) == 0
        n = 700
        c.generate(SEED, n, n_parties=1)
        tq = np.random.default_rng(21).integers(0, 2**64, size=(70, 400), dtype=np.uint64)   # 64 + 6 queries
        tq[66, :200], tq[66, 200:] = noisy_query(n - 1)
        md, mi = c.search(tq)
        assert mi[66] == n - 1
        for k in (0, 63, 64, 69):
            assert (md[k], mi[k]) == plaintext_min(tq[k, :200].copy(), tq[k, 200:].copy(), range(n))


def test_cluster_loads_the_reference_files(iris, tmp_path):
    n = 1000
    shares, masks = O.gen_share_rows(SEED, 0, n, threads=8), O.gen_mask_rows(SEED, 0, n, threads=8)
    shares.tofile(tmp_path / "mpc.share-0")
    masks.tofile(tmp_path / "mpc.masks")
    q, qm = O.gen_share_rows(71, 0, 1)[0], O.gen_mask_rows(72, 0, 1)[0]
    with iris.Cluster(cluster_devices(iris, 3), n) as c:
        c.load_files(str(tmp_path / "mpc.share-0"), str(tmp_path / "mpc.masks"))
        hd, hn = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
        c.match(q, qm, hd, hn)
        assert np.array_equal(hd, O.distance_batch(q, shares, threads=8)) and np.array_equal(hn, O.masks_batch(qm, masks, threads=8))
        with open(tmp_path / "bad", "wb") as f:
            f.write(b"x" * 1601)
        with pytest.raises(iris.IrisError):
            c.load_files(None, str(tmp_path / "bad"))
        with pytest.raises(iris.IrisError):
            c.load_files(str(tmp_path / "mpc.share-0"), str(tmp_path / "bad"))
    with iris.Cluster(cluster_devices(iris, 2), 10) as c:
        with pytest.raises(iris.IrisError):
            c.generate(SEED, 5000)                                # beyond capacity


def test_batched_search_over_several_slices_equals_per_query_search(iris):
    # 64 queries x 600 000 rows: two slices of the batched path; per-query fused scan + reduction as the yardstick
    n, nq = 600_000, 64
    tq = np.random.default_rng(11).integers(0, 2**64, size=(nq, 400), dtype=np.uint64)
    tq[3, :200], tq[3, 200:] = noisy_query(599_999)
    tq[40, :200], tq[40, 200:] = noisy_query(524_288, rotation=-5)
    with iris.Cluster([0], n) as c:
        c.generate(SEED, n, n_parties=1)
        md, mi = c.search(tq)
        db = c.shard(0)[0]
        for k in range(nq):
            de = iris.DistanceEngine.from_template(tq[k, :200].copy(), tq[k, 200:].copy())
            me = iris.MasksEngine(tq[k, 200:].copy())
            assert iris.match_min(de, me, db, 0, n) == (md[k], mi[k])
        assert mi[3] == 599_999 and mi[40] == 524_288


def test_streamed_match_reports_rows_in_order_while_scanning(iris):
    # the participant's producer (src/main.rs:425-434): results become available block by block
    n = 400_000                       # three chunks of the host-output pipeline
    qp, qm = noisy_query(5)
    with iris.Cluster([0], n) as c:
        c.generate(SEED, n, n_parties=2, party=0)
        db = c.shard(0)[0]
        ref = np.zeros((n, 31), np.uint16)
        de = iris.DistanceEngine.from_template(qp, qm)
        iris.match(de, None, db, 0, n, ref, None)
        out = np.zeros((n, 31), np.uint16)
        seen = []

        def progress(b, e):
            assert np.array_equal(out[b:e], ref[b:e])        # complete in host memory when reported
            seen.append((b, e))

        iris.match_streamed(de, None, db, 1000, n - 7, out[1000:n - 7], None, progress)
        assert len(seen) >= 3 and seen[0][0] == 1000 and seen[-1][1] == n - 7
        assert all(seen[i][1] == seen[i + 1][0] for i in range(len(seen) - 1))
        assert np.array_equal(out[1000:n - 7], ref[1000:n - 7]) and not out[:1000].any() and not out[n - 7:].any()
    with iris.Cluster(cluster_devices(iris, 3), n) as c:
        c.generate(SEED, n, n_parties=2, party=0)
        out = np.zeros((n, 31), np.uint16)
        import threading

        lock, covered = threading.Lock(), []

        def progress2(b, e):
            with lock:
                covered.append((b, e))

        c.match_template_streamed(qp, qm, out, progress2)
        assert np.array_equal(out, ref)
        covered.sort()
        assert covered[0][0] == 0 and covered[-1][1] == n and all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))


def test_scans_on_a_caller_stream_overlap_safely(iris):
    # the reference's calling pattern (src/main.rs:427-430) on a caller-supplied stream: chunk after chunk, device outputs
    import torch

    n, chunk = 300_000, 20_000
    q, qm = O.gen_share_rows(81, 0, 1)[0], O.gen_mask_rows(82, 0, 1)[0]
    with iris.Database(n) as db:
        db.generate(SEED, 0, n)
        de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
        ref_d = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        ref_n = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        db.set_overlap(False)
        iris.match(de, me, db, 0, n, ref_d, ref_n)
        db.synchronize()
        db.set_overlap(True)
        stream = torch.cuda.Stream()
        db.set_stream(stream.cuda_stream)
        for trial in range(3):
            dd = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
            dn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
            with torch.cuda.stream(stream):
                for b in range(0, n, chunk):
                    iris.match(de, me, db, b, b + chunk, dd[b:b + chunk], dn[b:b + chunk])
                    if trial == 1 and b % (3 * chunk) == 0:
                        dd[b:b + chunk].add_(0)               # foreign kernels between the scans, reading their output
                same_d, same_n = torch.equal(dd, ref_d), torch.equal(dn, ref_n)   # stream-ordered consumers
            stream.synchronize()
            assert same_d and same_n
        db.set_stream(None)


# ---------------------------------------------------------------------------------- two or more GPUs
def test_cluster_on_distinct_gpus_planted_in_the_last(iris):
    g = iris.device_count()
    if g < 2:
        pytest.skip("needs two GPUs")
    n = 40_000
    target = n - 3
    qp, qm = noisy_query(target)
    with iris.Cluster(list(range(g)), n) as c:
        c.generate(SEED, n, n_parties=1)
        assert c.shard(g - 1)[2] <= target
        md, mi = c.search(np.concatenate([qp, qm])[None, :].copy())
        p_t, m_t = synthetic_template(target)
        assert (md[0], mi[0]) == (O.template_distance(qp, qm, p_t, m_t), target)
        tq = np.random.default_rng(12).integers(0, 2**64, size=(8, 400), dtype=np.uint64)
        tq[5, :200], tq[5, 200:] = qp, qm
        md8, mi8 = c.search(tq)
        assert (md8[5], mi8[5]) == (md[0], target)


WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import mpc_iris_code_b200 as iris
rank, world, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
uid = bytes.fromhex(open(sys.argv[4]).read())
tq = np.load(sys.argv[5])
b, e = iris.cluster_partition(n, world, rank)
c = iris.Cluster([rank], e - b)
c.generate({seed}, e - b, first_row_id=b, n_parties=1)
c.set_index_base(b)
c.join(uid, rank, world)
md, mi = c.search(tq)
np.save(sys.argv[6] + f".{{rank}}.npy", np.stack([md, mi.astype(np.float64)]))
# the full result vectors of all processes, gathered on every process over NCCL
import torch
torch.cuda.set_device(rank)
q = iris.encode(tq[1, :200].copy(), tq[1, 200:].copy(), device=rank)
dd = torch.zeros((n, 31), dtype=torch.int16, device=f"cuda:{{rank}}")
dn = torch.zeros((n, 31), dtype=torch.int16, device=f"cuda:{{rank}}")
c.match_allgather(q, tq[1, 200:].copy(), dd, dn)
np.save(sys.argv[6] + f".vec{{rank}}.npy", np.stack([dd.cpu().numpy(), dn.cpu().numpy()]))
c.close()
"""


def test_processes_join_one_cluster_over_nccl(iris, tmp_path):
    g = iris.device_count()
    if g < 2:
        pytest.skip("needs two GPUs")
    world, n = min(g, 4), 30_000
    (tmp_path / "uid").write_text(iris.comm_unique_id().hex())
    tq = np.random.default_rng(13).integers(0, 2**64, size=(3, 400), dtype=np.uint64)
    tq[1, :200], tq[1, 200:] = noisy_query(n - 1)
    np.save(tmp_path / "tq.npy", tq)
    (tmp_path / "worker.py").write_text(WORKER.format(root=ROOT, seed=SEED))
    procs = [subprocess.Popen([sys.executable, str(tmp_path / "worker.py"), str(r), str(world), str(n), str(tmp_path / "uid"),
                               str(tmp_path / "tq.npy"), str(tmp_path / "out")]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    res = [np.load(str(tmp_path / "out") + f".{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(res[0], res[r])                    # every process holds the global answer
    p_t, m_t = synthetic_template(n - 1)
    assert res[0][1][1] == n - 1 and res[0][0][1] == O.template_distance(tq[1, :200].copy(), tq[1, 200:].copy(), p_t, m_t)
    with iris.Cluster([0], n) as c:                              # one shard holding everything gives the same three answers
        c.generate(SEED, n, n_parties=1)
        md, mi = c.search(tq)
        assert np.array_equal(res[0][0], md) and np.array_equal(res[0][1], mi.astype(np.float64))
        hd, hn = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
        c.match_template(tq[1, :200].copy(), tq[1, 200:].copy(), hd, hn)
    for r in range(world):                                       # ... and every process holds every row's vectors
        vec = np.load(str(tmp_path / "out") + f".vec{r}.npy").view(np.uint16)
        assert np.array_equal(vec[0], hd) and np.array_equal(vec[1], hn)
