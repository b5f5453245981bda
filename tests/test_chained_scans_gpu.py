"""Consecutive device-output scans on the library's own stream may overlap (programmatic dependent launch: the next scan
starts on the SMs the previous one's tail leaves idle, iris_abi.cu scan_core).  The reference's calling pattern --
batch_process chunk after chunk, src/main.rs:427-430, 512-515 -- must give the same bytes as one call, and calls that
reuse an output buffer must still complete in order."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE


@pytest.fixture(scope="module")
def ctx():
    import torch

    import mpc_iris_code_b200 as iris

    n = 61_000                                       # 477 tiles: every 20 000-row call ends with a partial wave
    db = iris.Database(n)                            # no set_stream: the library's own stream, chaining allowed
    db.generate(SEED, 0, n)
    yield iris, torch, db, n
    db.close()


def _queries(iris, k):
    t = np.random.default_rng(300 + k).integers(0, 2**64, size=400, dtype=np.uint64)
    q = iris.encode(t[:200].copy(), t[200:].copy())
    return q, t[200:].copy()


@pytest.mark.parametrize("chunk", [20_000, 7_001, 640])
def test_chunked_calls_equal_one_call(ctx, chunk):
    iris, torch, db, n = ctx
    q, qm = _queries(iris, chunk)
    de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
    ref_d = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
    ref_n = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
    db.check_distances_simt(q, 0, n, ref_d)
    db.check_denominators_simt(qm, 0, n, ref_n)
    for mode in ("fused", "distances", "denominators"):
        got_d = torch.full((n, 31), 0x5A5A, dtype=torch.int16, device="cuda")
        got_n = torch.full((n, 31), 0x5A5A, dtype=torch.int16, device="cuda")
        for c in range(0, n, chunk):                 # back-to-back asynchronous calls, disjoint output slices
            e = min(n, c + chunk)
            iris.match(de if mode != "denominators" else None, me if mode != "distances" else None, db, c, e,
                       got_d[c:e] if mode != "denominators" else None, got_n[c:e] if mode != "distances" else None)
        db.synchronize()
        if mode != "denominators":
            assert torch.equal(got_d, ref_d), (mode, chunk)
        if mode != "distances":
            assert torch.equal(got_n, ref_n), (mode, chunk)
    host = ref_d.cpu().numpy().view(np.uint16)
    for i in (0, 19_999, 20_000, n - 1):
        assert np.array_equal(host[i], O.distance_batch(q, O.gen_share_rows(SEED, int(i), 1))[0]), i


def test_reused_output_buffer_keeps_stream_order(ctx):
    """Different queries into the SAME buffer without a synchronize in between: the last call wins, row for row."""
    iris, torch, db, n = ctx
    rows = 30_000
    out_d = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    out_n = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    engines = [_queries(iris, 40 + k) for k in range(5)]
    des = [iris.DistanceEngine(q) for q, _ in engines]
    mes = [iris.MasksEngine(m) for _, m in engines]
    for _ in range(3):
        for k in range(5):
            iris.match(des[k], mes[k], db, 1_000, 1_000 + rows, out_d, out_n)
    db.synchronize()
    ref_d = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    ref_n = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    db.check_distances_simt(engines[4][0], 1_000, 1_000 + rows, ref_d)
    db.check_denominators_simt(engines[4][1], 1_000, 1_000 + rows, ref_n)
    assert torch.equal(out_d, ref_d) and torch.equal(out_n, ref_n)
    # partially overlapping slices of one buffer
    buf = torch.zeros((rows + 5_000, 31), dtype=torch.int16, device="cuda")
    iris.match(des[0], None, db, 0, rows, buf[:rows], None)
    iris.match(des[1], None, db, 0, rows, buf[5_000:], None)          # overlaps rows 5 000 .. 30 000 of the first call
    db.synchronize()
    r0 = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    r1 = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    db.check_distances_simt(engines[0][0], 0, rows, r0)
    db.check_distances_simt(engines[1][0], 0, rows, r1)
    assert torch.equal(buf[:5_000], r0[:5_000]) and torch.equal(buf[5_000:], r1)


def test_engine_freed_and_recreated_while_its_scan_is_still_running():
    """Device-output scans are asynchronous and engines come from a pool: freeing an engine right behind its scan and
    building the next one (which takes over the same operand buffers) must not disturb the scan that is still
    reading them -- the release is ordered on the shard's stream, not on the host."""
    import torch

    import mpc_iris_code_b200 as iris

    rows = 300_000
    rng = np.random.default_rng(77)
    q1, q2 = (rng.integers(0, 2**16, size=12800, dtype=np.uint16) for _ in range(2))
    m1, m2 = (rng.integers(0, 2**64, size=200, dtype=np.uint64) for _ in range(2))
    with iris.Database(rows) as db:
        db.generate(0x1715C0DE, 0, rows)
        refs = []
        for q, m in ((q1, m1), (q2, m2)):
            d = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
            n = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
            e, me = iris.DistanceEngine(q), iris.MasksEngine(m)
            iris.match(e, me, db, 0, rows, d, n)
            db.synchronize()
            refs.append((d, n))
        outs = [(torch.zeros_like(refs[0][0]), torch.zeros_like(refs[0][1])) for _ in range(2)]
        for trial in range(15):
            for k, (q, m) in enumerate(((q1, m1), (q2, m2))):
                e, me = iris.DistanceEngine(q), iris.MasksEngine(m)
                iris.match(e, me, db, 0, rows, outs[k][0], outs[k][1])      # returns while the scan runs
                e.close()                                                   # freed at once; the next pair reuses the slots
                me.close()
            db.synchronize()
            for k in range(2):
                assert torch.equal(outs[k][0], refs[k][0]) and torch.equal(outs[k][1], refs[k][1]), (trial, k)
