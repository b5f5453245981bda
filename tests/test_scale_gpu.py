"""Sizes beyond BASELINE config 2: the largest point of the denominators sweep (config 3: 4 M masks, 6.4 GB, byte
offsets past 2^32) and a multi-million-row shard of shares (config 5's 2 M rows per GPU: 54 GB), plus the
out-of-memory error path.  Every row is compared on the device with the independent CUDA-core kernels; sampled
rows are regenerated from the counter-based generator and compared with the CPU oracle."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE


@pytest.fixture(scope="module")
def iris():
    import mpc_iris_code_b200 as iris_mod

    return iris_mod


def test_denominators_sweep_top_size_every_row(iris):
    import torch

    n = 4_000_000
    free, _ = torch.cuda.mem_get_info()
    if free < 12e9:
        pytest.skip("not enough free HBM")
    with iris.Database(n, shares=False) as db:
        db.generate(SEED, 0, n)
        qm = O.gen_mask_rows(60, 1, 1)[0]
        me = iris.MasksEngine(qm)
        dn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        me.batch_process(dn, db)                               # mask_scan_kernel (TMEM-operand path)
        db.synchronize()
        cn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        db.check_denominators_simt(qm, 0, n, cn)
        assert torch.equal(dn, cn)
        host = dn.cpu().numpy().view(np.uint16)
        r = np.random.default_rng(61)
        for i in np.concatenate([[0, 255, 256, n - 257, n - 1, 2_684_355, 2_684_356], r.integers(0, n, 100)]):
            assert np.array_equal(host[i], O.masks_batch(qm, O.gen_mask_rows(SEED, int(i), 1))[0]), i
        # an odd row range in the middle, straight into an unaligned slice of a larger device buffer
        rb, re = 1_234_567, 3_210_987
        part = torch.zeros(((re - rb) * 31 + 7,), dtype=torch.int16, device="cuda")
        me.batch_process(part[5 : 5 + (re - rb) * 31], db, rb, re)
        db.synchronize()
        assert torch.equal(part[5 : 5 + (re - rb) * 31].view(-1, 31), cn[rb:re])
        assert int(part[:5].abs().sum()) == 0 and int(part[5 + (re - rb) * 31 :].abs().sum()) == 0


def test_two_million_row_shard_fused_and_batched(iris):
    import torch

    n = 2_000_000
    free, _ = torch.cuda.mem_get_info()
    if free < 70e9:
        pytest.skip("not enough free HBM for a 2 M-row shard")
    with iris.Database(n) as db:
        db.generate(SEED, 7_000_000, n)                         # row ids of a shard in the middle of a 16 M-row database
        pattern, qm = O.gen_mask_rows(62, 0, 1)[0], O.gen_mask_rows(62, 1, 1)[0]
        q = O.encode(pattern, qm)
        de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
        dd = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        dn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        iris.match(de, me, db, 0, n, dd, dn)
        db.synchronize()
        cd = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        cn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        db.check_distances_simt(q, 0, n, cd)
        db.check_denominators_simt(qm, 0, n, cn)
        assert torch.equal(dd, cd) and torch.equal(dn, cn)
        hd = dd.cpu().numpy().view(np.uint16)
        r = np.random.default_rng(63)
        for i in np.concatenate([[0, n - 1, 1_999_871, 1_999_872], r.integers(0, n, 60)]):
            assert np.array_equal(hd[i], O.distance_batch(q, O.gen_share_rows(SEED, 7_000_000 + int(i), 1))[0]), i
        # the reduction over the whole shard equals the reduction of the per-row results
        md, mi = iris.match_min(de, me, db, 0, n, index_base=7_000_000)
        want = iris.combine_min([dd], dn, index_base=7_000_000)
        assert (md, mi) == want
        # ... and the running minimum continued over ragged blocks of rows (src/main.rs:611-621) gives the same answer
        parts = [iris.match_min(de, me, db, b, e, index_base=7_000_000) for b, e in ((0, 777_001), (777_001, 1_500_000), (1_500_000, n))]
        assert min(parts, key=lambda t: (t[0], t[1])) == (md, mi)
        # batched kernels on the last 300 k rows of the shard (tile indices past 2^32 / 32 KiB)
        rb = n - 300_000
        tmpl = np.random.default_rng(64).integers(0, 2**64, size=(3, 400), dtype=np.uint64)
        des, mes = iris.engines_from_templates(tmpl)
        bd = torch.zeros((3, n - rb, 31), dtype=torch.int16, device="cuda")
        bn = torch.zeros((3, n - rb, 31), dtype=torch.int16, device="cuda")
        iris.distances_batch(des, db, rb, n, bd)
        iris.denominators_batch(mes, db, rb, n, bn)
        db.synchronize()
        for k in range(3):
            db.check_distances_simt(O.encode(tmpl[k, :200], tmpl[k, 200:]), rb, n, cd[: n - rb])
            db.check_denominators_simt(tmpl[k, 200:], rb, n, cn[: n - rb])
            assert torch.equal(bd[k], cd[: n - rb]) and torch.equal(bn[k], cn[: n - rb])


def test_allocation_failure_is_an_error_code_not_a_crash(iris):
    with pytest.raises(iris.IrisError) as ei:
        iris.Database(12_000_000)                               # 326 GB of shares + masks
    assert ei.value.code == -3                                  # IRIS_ERR_NOMEM
    with iris.Database(256) as db:                              # the device is still usable
        db.generate(SEED, 0, 256)
        out = np.zeros((256, 31), np.uint16)
        qm = O.gen_mask_rows(65, 0, 1)[0]
        iris.MasksEngine(qm).batch_process(out, db)
        assert np.array_equal(out, O.masks_batch(qm, O.gen_mask_rows(SEED, 0, 256)))
