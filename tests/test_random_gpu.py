"""Seeded randomised sweep (the reference's property tests use thread_rng x100; here every case is reproducible):
random row windows, output phases, query kinds and call styles against the oracle, all bit-exact."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE


def test_random_windows_queries_and_call_styles():
    import torch

    import mpc_iris_code_b200 as iris

    n = 2111                                           # 16.5 tiles: last pair tile is ragged
    shares = O.gen_share_rows(SEED, 10_000, n, threads=8)
    masks = O.gen_mask_rows(SEED, 10_000, n, threads=8)
    r = np.random.default_rng(2024)
    with iris.Database(n) as db:
        db.generate(SEED, 10_000, n)                   # device generator == oracle generator
        for case in range(24):
            rb = int(r.integers(0, n - 1))
            re = int(r.integers(rb + 1, min(n, rb + int(r.integers(1, 900))) + 1))
            kind = case % 3
            if kind == 0:                              # ternary query (signed two-product path)
                qm = r.integers(0, 2**64, 200, dtype=np.uint64)
                q = O.encode(r.integers(0, 2**64, 200, dtype=np.uint64), qm)
            elif kind == 1:                            # uniform u16 query (three-product path)
                qm = r.integers(0, 2**64, 200, dtype=np.uint64)
                q = r.integers(0, 2**16, O.BITS, dtype=np.uint16)
            else:                                      # sparse mask / small-valued query
                qm = r.integers(0, 2**64, 200, dtype=np.uint64) & r.integers(0, 2**64, 200, dtype=np.uint64)
                q = r.integers(0, 3, O.BITS, dtype=np.uint16)
            exp_d = O.distance_batch(q, shares[rb:re], threads=8)
            exp_n = O.masks_batch(qm, masks[rb:re], threads=8)
            de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
            m = (re - rb) * 31
            style = case % 4
            if style == 0:                             # fused, host outputs
                d, dn = np.zeros((re - rb, 31), np.uint16), np.zeros((re - rb, 31), np.uint16)
                iris.match(de, me, db, rb, re, d, dn)
            elif style == 1:                           # separate engines, host outputs
                d, dn = np.zeros((re - rb, 31), np.uint16), np.zeros((re - rb, 31), np.uint16)
                de.batch_process(d, db, rb, re)
                me.batch_process(dn, db, rb, re)
            elif style == 2:                           # fused, device outputs at a random element phase
                off = int(r.integers(0, 9))
                bd = torch.zeros(m + 16, dtype=torch.int16, device="cuda")
                bn = torch.zeros(m + 16, dtype=torch.int16, device="cuda")
                iris.match(de, me, db, rb, re, bd[off : off + m], bn[off : off + m])
                db.synchronize()
                d = bd.cpu().numpy().view(np.uint16)[off : off + m].reshape(-1, 31)
                dn = bn.cpu().numpy().view(np.uint16)[off : off + m].reshape(-1, 31)
            else:                                      # batched kernels with this query among others
                others = [O.encode(r.integers(0, 2**64, 200, dtype=np.uint64), r.integers(0, 2**64, 200, dtype=np.uint64)) for _ in range(2)]
                pos = int(r.integers(0, 3))
                qs = others[:pos] + [q] + others[pos:]
                out = np.zeros((3, re - rb, 31), np.uint16)
                iris.distances_batch([iris.DistanceEngine(x) for x in qs], db, rb, re, out)
                d = out[pos].copy()
                iris.denominators_batch([iris.MasksEngine(qm)] * 3, db, rb, re, out)
                dn = out[1].copy()
            assert np.array_equal(d, exp_d), (case, rb, re, kind, style)
            assert np.array_equal(dn, exp_n), (case, rb, re, kind, style)
