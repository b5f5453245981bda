"""Out-of-bounds guards (compute-sanitizer is not available on the pool): every kernel writes its results into the
middle of a poisoned device buffer at an odd element offset; bytes outside the result must stay untouched and the
result must still be exact.  Covers the 16-byte-phase logic of the staged 62-byte-row stores."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE
POISON = 0x6B6B


@pytest.fixture(scope="module")
def ctx():
    import torch

    import mpc_iris_code_b200 as iris

    n = 777
    shares = O.gen_share_rows(SEED, 0, n, threads=8)
    masks = O.gen_mask_rows(SEED, 0, n, threads=8)
    db = iris.Database(n)
    db.append_shares(shares)
    db.append_masks(masks)
    yield iris, torch, db, shares, masks, n
    db.close()


def _region(torch, elems, off):
    buf = torch.full((elems + 128,), POISON, dtype=torch.int16, device="cuda")
    return buf, buf[off : off + elems]


def _check(buf, off, elems, expected):
    got = buf.cpu().numpy().view(np.uint16)
    assert np.array_equal(got[off : off + elems], expected.reshape(-1))
    assert (got[:off] == POISON).all() and (got[off + elems :] == POISON).all()


@pytest.mark.parametrize("off", [0, 1, 5, 8, 13])
@pytest.mark.parametrize("rb,re", [(0, 777), (129, 640), (255, 257)])
def test_scan_kernels_stay_inside_their_output(ctx, off, rb, re):
    iris, torch, db, shares, masks, n = ctx
    qm = O.gen_mask_rows(21, 1, 1)[0]
    qt = O.encode(O.gen_mask_rows(21, 0, 1)[0], qm)          # signed two-product path
    qu = O.gen_share_rows(22, 0, 1)[0]                       # three-product path
    m = (re - rb) * 31
    me = iris.MasksEngine(qm)
    exp_n = O.masks_batch(qm, masks[rb:re])
    for q in (qt, qu):
        de = iris.DistanceEngine(q)
        exp_d = O.distance_batch(q, shares[rb:re])
        bd, vd = _region(torch, m, off)
        bn, vn = _region(torch, m, off + 3)
        iris.match(de, me, db, rb, re, vd, vn)               # fused scan
        db.synchronize()
        _check(bd, off, m, exp_d)
        _check(bn, off + 3, m, exp_n)
        bd, vd = _region(torch, m, off)
        de.batch_process(vd, db, rb, re)                     # distances only
        db.synchronize()
        _check(bd, off, m, exp_d)
    bn, vn = _region(torch, m, off)
    me.batch_process(vn, db, rb, re)                         # mask_scan_kernel (TMEM operand, tile pairs)
    db.synchronize()
    _check(bn, off, m, exp_n)


def test_batched_kernels_more_queries_than_one_launch_and_tiny_shard(ctx):
    """70 queries span two launches (64 + 6); a 1-row shard exercises the zero-filled pair tile."""
    iris, torch, db, shares, masks, n = ctx
    nq = 70
    qs = [O.gen_share_rows(500 + i, 0, 1)[0] if i % 9 == 0 else O.encode(O.gen_mask_rows(500 + i, 0, 1)[0], O.gen_mask_rows(500 + i, 1, 1)[0])
          for i in range(nq)]
    qms = [O.gen_mask_rows(500 + i, 1, 1)[0] for i in range(nq)]
    rb, re = 100, 420
    out = np.zeros((nq, re - rb, 31), np.uint16)
    iris.distances_batch([iris.DistanceEngine(q) for q in qs], db, rb, re, out)
    assert np.array_equal(out, np.stack([O.distance_batch(q, shares[rb:re], threads=8) for q in qs]))
    iris.denominators_batch([iris.MasksEngine(q) for q in qms], db, rb, re, out)
    assert np.array_equal(out, np.stack([O.masks_batch(q, masks[rb:re], threads=8) for q in qms]))
    with iris.Database(1) as tiny:
        tiny.append_shares(shares[:1])
        tiny.append_masks(masks[:1])
        o1 = np.zeros((3, 1, 31), np.uint16)
        iris.distances_batch([iris.DistanceEngine(q) for q in qs[:3]], tiny, 0, 1, o1)
        assert np.array_equal(o1[:, 0], np.stack([O.distances(q, shares[0]) for q in qs[:3]]))
        iris.denominators_batch([iris.MasksEngine(q) for q in qms[:3]], tiny, 0, 1, o1)
        assert np.array_equal(o1[:, 0], np.stack([O.denominators(q, masks[0]) for q in qms[:3]]))
        d, dn = np.zeros((1, 31), np.uint16), np.zeros((1, 31), np.uint16)
        iris.match(iris.DistanceEngine(qs[1]), iris.MasksEngine(qms[1]), tiny, 0, 1, d, dn)
        assert np.array_equal(d[0], O.distances(qs[1], shares[0])) and np.array_equal(dn[0], O.denominators(qms[1], masks[0]))


@pytest.mark.parametrize("off", [0, 3, 8])
@pytest.mark.parametrize("rb,re", [(0, 777), (130, 700)])
def test_batched_kernels_stay_inside_their_output(ctx, off, rb, re):
    iris, torch, db, shares, masks, n = ctx
    nq = 11
    qs = [O.encode(O.gen_mask_rows(300 + i, 0, 1)[0], O.gen_mask_rows(300 + i, 1, 1)[0]) for i in range(nq)]
    qms = [O.gen_mask_rows(300 + i, 1, 1)[0] for i in range(nq)]
    m = nq * (re - rb) * 31
    buf, view = _region(torch, m, off)
    iris.distances_batch([iris.DistanceEngine(q) for q in qs], db, rb, re, view)
    db.synchronize()
    _check(buf, off, m, np.stack([O.distance_batch(q, shares[rb:re]) for q in qs]))
    buf, view = _region(torch, m, off)
    iris.denominators_batch([iris.MasksEngine(q) for q in qms], db, rb, re, view)
    db.synchronize()
    _check(buf, off, m, np.stack([O.masks_batch(q, masks[rb:re]) for q in qms]))
