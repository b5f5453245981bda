"""BASELINE configs[3]-sized properties of the batched int8-GEMM path (64 queries, hundreds of thousands of rows):
the batched kernels must agree, on every row, with the single-query scan kernels (an independent code path that the
other tests pin to the oracle), and on sampled rows with the oracle itself."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE


@pytest.fixture(scope="module")
def setup():
    import torch

    import mpc_iris_code_b200 as iris

    n = 300_000
    free, _ = torch.cuda.mem_get_info()
    if free < 20e9:
        pytest.skip("not enough free HBM")
    db = iris.Database(n)
    db.generate(SEED, 0, n)
    yield iris, torch, db, n
    db.close()


@pytest.mark.parametrize("kind", ["ternary", "uniform"])
def test_batched_distances_equal_single_query_scan_everywhere(setup, kind):
    iris, torch, db, n = setup
    nq = 64
    if kind == "ternary":
        qs = [O.encode(O.gen_mask_rows(3000 + i, 0, 1)[0], O.gen_mask_rows(3000 + i, 1, 1)[0]) for i in range(nq)]
    else:
        qs = [O.gen_share_rows(4000 + i, 0, 1)[0] for i in range(nq)]
    engines = [iris.DistanceEngine(q) for q in qs]
    out = torch.zeros((nq, n, 31), dtype=torch.int16, device="cuda")
    iris.distances_batch(engines, db, 0, n, out)
    db.synchronize()
    single = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
    for qi in (0, 7, 8, 31, 63):
        engines[qi].batch_process(single, db)
        db.synchronize()
        assert torch.equal(out[qi], single), (kind, qi)
    got = out[:, [0, 255, 256, n // 2, n - 1]].cpu().numpy().view(np.uint16)
    for qi in (0, 40, 63):
        for k, i in enumerate((0, 255, 256, n // 2, n - 1)):
            assert np.array_equal(got[qi, k], O.distance_batch(qs[qi], O.gen_share_rows(SEED, i, 1))[0])


def test_batched_denominators_equal_single_query_scan_everywhere(setup):
    iris, torch, db, n = setup
    nq = 64
    qms = [O.gen_mask_rows(5000 + i, 1, 1)[0] for i in range(nq)]
    engines = [iris.MasksEngine(q) for q in qms]
    out = torch.zeros((nq, n, 31), dtype=torch.int16, device="cuda")
    iris.denominators_batch(engines, db, 0, n, out)
    db.synchronize()
    single = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
    for qi in (0, 15, 16, 47, 63):
        engines[qi].batch_process(single, db)      # mask_scan_kernel (TMEM operand)
        db.synchronize()
        assert torch.equal(out[qi], single), qi
    check = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
    db.check_denominators_simt(qms[5], 0, n, check)  # CUDA-core kernel
    assert torch.equal(out[5], check)
    assert int(out.max()) <= 12800


def test_combine_min_batch_matches_per_query_reduction(setup):
    iris, torch, db, n = setup
    nq, m = 9, 50_000
    qms = [O.gen_mask_rows(7000 + i, 1, 1)[0] for i in range(nq)]
    qs = [O.encode(O.gen_mask_rows(7000 + i, 0, 1)[0], qms[i]) for i in range(nq)]
    d = torch.zeros((nq, m, 31), dtype=torch.int16, device="cuda")
    dn = torch.zeros((nq, m, 31), dtype=torch.int16, device="cuda")
    iris.distances_batch([iris.DistanceEngine(q) for q in qs], db, 1000, 1000 + m, d)
    iris.denominators_batch([iris.MasksEngine(q) for q in qms], db, 1000, 1000 + m, dn)
    db.synchronize()
    mins, idxs = iris.combine_min_batch(d, dn, nq, index_base=1000)
    hd, hn = d.cpu().numpy().view(np.uint16), dn.cpu().numpy().view(np.uint16)
    for i in range(nq):
        em, ei = O.combine_min(hd[i][None], hn[i])
        assert (mins[i], idxs[i]) == (em, 1000 + ei)


def test_match_min_at_scale_agrees_with_host_reduction(setup):
    iris, torch, db, n = setup
    qm = O.gen_mask_rows(6000, 1, 1)[0]
    q = O.encode(O.gen_mask_rows(6000, 0, 1)[0], qm)
    de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
    d, den = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
    iris.match(de, me, db, 0, n, d, den)
    assert iris.match_min(de, me, db, 0, n) == O.combine_min(d[None], den)
