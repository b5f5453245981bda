"""Diagnostic run for a B200 box: exercises every kernel at small size against the oracle and
prints what differs and how (layout, generator, per-product raw accumulators, final outputs).

    python tests/diagnostics/gpu_probe.py [n_rows]
"""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402
import oracle as O  # noqa: E402
from mpc_iris_code_b200 import api  # noqa: E402

SEED = 0x1715C0DE
FAILS = []


def report(name, ok, extra=""):
    print(f"[{'PASS' if ok else 'FAIL'}] {name} {extra}", flush=True)
    if not ok:
        FAILS.append(name)


def step(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        traceback.print_exc()
        report(fn.__name__, False, f"exception: {e}")


def expected_raw(q, qm, shares, masks):
    """numpy model of the four s32 accumulators per (row, rotation)."""
    rot = np.stack([O.np_encoded_rotated(q, j - 15) for j in range(31)]).astype(np.int64)
    d = shares.astype(np.int64)
    qlo, qhi = rot & 0xFF, rot >> 8
    dlo, dhi = d & 0xFF, d >> 8
    s00 = dlo @ qlo.T
    s10 = dlo @ qhi.T
    s01 = dhi @ qlo.T
    dm = 128 * O.np_masks_batch(qm, masks).astype(np.int64)
    return s00, s10, s01, dm


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    print("devices:", iris.device_count(), flush=True)
    shares = O.gen_share_rows(SEED, 0, n, threads=8)
    masks = O.gen_mask_rows(SEED, 0, n, threads=8)
    q_uni = O.gen_share_rows(77, 0, 1)[0]
    qp, qm = O.gen_mask_rows(78, 0, 1)[0], O.gen_mask_rows(78, 1, 1)[0]
    q_ter = O.encode(qp, qm)

    db = iris.Database(n + 77)

    def loader_roundtrip():
        db.append_shares(shares)
        db.append_masks(masks)
        report("loader_roundtrip.shares", np.array_equal(db.read_shares(0, n), shares))
        report("loader_roundtrip.masks", np.array_equal(db.read_masks(0, n), masks))

    step(loader_roundtrip)

    def generator_parity():
        g = iris.Database(n)
        g.generate(SEED, 0, n)
        report("generator.shares", np.array_equal(g.read_shares(0, n), shares))
        report("generator.masks", np.array_equal(g.read_masks(0, n), masks))
        g.close()

    step(generator_parity)

    exp_d_uni = O.distance_batch(q_uni, shares, threads=8)
    exp_d_ter = O.distance_batch(q_ter, shares, threads=8)
    exp_den = O.masks_batch(qm, masks, threads=8)

    def simt_kernels():
        report("simt.distances.uniform", np.array_equal(db.check_distances_simt(q_uni, 0, n), exp_d_uni))
        report("simt.distances.ternary", np.array_equal(db.check_distances_simt(q_ter, 0, n), exp_d_ter))
        report("simt.denominators", np.array_equal(db.check_denominators_simt(qm, 0, n), exp_den))

    step(simt_kernels)

    def dots():
        report("dot_u16", iris.dot_u16(q_uni, shares[0]) == O.dot_u16(q_uni, shares[0]))
        report("dot_bool", iris.dot_bool(qm, masks[0]) == O.dot_bool(qm, masks[0]))

    step(dots)

    def raw_accumulators():
        for name, q in (("uniform", q_uni), ("ternary", q_ter)):
            de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
            raw = api.raw_accumulators(de, me, db, 0, n)[:n].astype(np.int64)
            exp = expected_raw(q, qm, shares, masks)
            for blk, nm in enumerate(("S00", "S10", "S01", "128*pop")):
                got = raw[:, 32 * blk : 32 * blk + 31]
                ok = np.array_equal(got, exp[blk])
                extra = ""
                if not ok:
                    bad = np.argwhere(got != exp[blk])
                    extra = f"mismatches={len(bad)}/{got.size} first={bad[:4].tolist()} got={got[tuple(bad[0])]} exp={exp[blk][tuple(bad[0])]}"
                    rows_bad = np.unique(bad[:, 0])
                    extra += f" bad_rows={len(rows_bad)} (first {rows_bad[:8].tolist()}) bad_cols={np.unique(bad[:, 1]).tolist()[:16]}"
                report(f"raw.{name}.{nm}", ok, extra)
            pad = raw[:, [31, 63, 95, 127]]
            report(f"raw.{name}.pad_columns_zero", bool((pad == 0).all()), f"max|pad|={np.abs(pad).max()}")

    step(raw_accumulators)

    def final_outputs():
        de_u, de_t, me = iris.DistanceEngine(q_uni), iris.DistanceEngine(q_ter), iris.MasksEngine(qm)
        for rb, re in ((0, n), (3, n - 5), (130, min(n, 257)), (127, 129)):
            if re <= rb:
                continue
            m = re - rb
            out = np.zeros((m, 31), np.uint16)
            de_u.batch_process(out, db, rb, re)
            report(f"scan.distances.uniform[{rb}:{re}]", np.array_equal(out, exp_d_uni[rb:re]))
            out2 = np.zeros((m, 31), np.uint16)
            me.batch_process(out2, db, rb, re)
            report(f"scan.denominators[{rb}:{re}]", np.array_equal(out2, exp_den[rb:re]))
            o3, o4 = np.zeros((m, 31), np.uint16), np.zeros((m, 31), np.uint16)
            iris.match(de_t, me, db, rb, re, o3, o4)
            report(f"scan.fused.ternary[{rb}:{re}]", np.array_equal(o3, exp_d_ter[rb:re]) and np.array_equal(o4, exp_den[rb:re]))

    step(final_outputs)

    def literal_api():
        de = iris.DistanceEngine(q_uni)
        out = np.zeros((n, 31), np.uint16)
        de.batch_process(out, shares)
        report("literal.DistanceEngine.batch_process(host slice)", np.array_equal(out, exp_d_uni))
        me = iris.MasksEngine(qm)
        out = np.zeros((n, 31), np.uint16)
        me.batch_process(out, masks)
        report("literal.MasksEngine.batch_process(host slice)", np.array_equal(out, exp_den))
        report("distances()", np.array_equal(iris.distances(q_ter, shares[1]), exp_d_ter[1]))
        report("denominators()", np.array_equal(iris.denominators(qm, masks[1]), exp_den[1]))

    step(literal_api)

    print("FAILED:" if FAILS else "ALL PASS", FAILS, flush=True)
    return 1 if FAILS else 0


if __name__ == "__main__":
    sys.exit(main())
