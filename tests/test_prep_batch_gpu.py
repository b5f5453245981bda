"""iris_engines_new_from_templates: Q engines of each kind from Q wire Templates in one batch must behave exactly like
engines built one at a time (DistanceEngine::new(&encode(&t)), MasksEngine::new(&t.mask); src/main.rs:427, 512)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE


@pytest.mark.parametrize("nq", [1, 7, 64, 70])
def test_engines_from_templates_match_individual_engines(nq):
    import mpc_iris_code_b200 as iris

    n = 600
    shares = O.gen_share_rows(SEED, 0, n, threads=8)
    masks = O.gen_mask_rows(SEED, 0, n, threads=8)
    templates = np.stack([np.concatenate([O.gen_mask_rows(8000 + i, 0, 1)[0], O.gen_mask_rows(8000 + i, 1, 1)[0]]) for i in range(nq)])
    with iris.Database(n) as db:
        db.append_shares(shares)
        db.append_masks(masks)
        des, mes = iris.engines_from_templates(templates)
        assert len(des) == len(mes) == nq
        for i in sorted({0, nq // 2, nq - 1}):
            p, m = templates[i, :200], templates[i, 200:]
            d, dn = np.zeros((n, 31), np.uint16), np.zeros((n, 31), np.uint16)
            iris.match(des[i], mes[i], db, 0, n, d, dn)
            assert np.array_equal(d, O.distance_batch(O.encode(p, m), shares, threads=8))
            assert np.array_equal(dn, O.masks_batch(m, masks, threads=8))
        out = np.zeros((nq, n, 31), np.uint16)
        iris.distances_batch(des, db, 0, n, out)          # all engines are flagged s8: signed two-product GEMM
        for i in sorted({0, nq - 1}):
            assert np.array_equal(out[i], O.distance_batch(O.encode(templates[i, :200], templates[i, 200:]), shares, threads=8))
        only_d, none = iris.engines_from_templates(templates, masks=False)
        assert len(only_d) == nq and none == []
