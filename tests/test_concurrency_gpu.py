"""The C ABI's threading contract (include/iris_b200.h): calls on ONE handle are serialised by the
caller, calls on DIFFERENT handles may run concurrently from different host threads.  The reference's engines are
`Sync` and are driven from rayon workers / spawn_blocking threads (src/lib.rs:44-50, src/main.rs:425-431)."""
import threading

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE


def test_independent_handles_from_concurrent_threads():
    import mpc_iris_code_b200 as iris

    n_threads, rounds = 4, 6
    errors, results = [], {}

    def worker(t):
        try:
            n = 300 + 257 * t                                         # different shapes per thread
            shares = O.gen_share_rows(SEED + t, 0, n, threads=1)
            masks = O.gen_mask_rows(SEED + t, 0, n, threads=1)
            with iris.Database(n) as db:
                db.append_shares(shares)
                db.append_masks(masks)
                for r in range(rounds):
                    pattern, mask = O.gen_mask_rows(500 + 10 * t + r, 0, 1)[0], O.gen_mask_rows(500 + 10 * t + r, 1, 1)[0]
                    de = iris.DistanceEngine.from_template(pattern, mask)          # per-call pool allocations
                    me = iris.MasksEngine(mask)
                    dist = np.zeros((n, 31), np.uint16)
                    den = np.zeros((n, 31), np.uint16)
                    iris.match(de, me, db, 0, n, dist, den)                       # host outputs: chunked pipeline
                    q = O.encode(pattern, mask)
                    assert np.array_equal(dist, O.distance_batch(q, shares)), (t, r, "distances")
                    assert np.array_equal(den, O.masks_batch(mask, masks)), (t, r, "denominators")
                    got = iris.combine_min([dist], den)                            # stream-ordered temporaries
                    assert got == O.combine_min(dist[None], den), (t, r, "combine")
                    results[(t, r)] = got
                    de.close()
                    me.close()
        except BaseException as ex:  # noqa: BLE001
            errors.append((t, repr(ex)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors
    assert len(results) == n_threads * rounds
