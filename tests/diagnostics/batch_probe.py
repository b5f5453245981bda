"""Diagnostic + timing run of the batched (dense GEMM) distances path on a B200 box.

    python tests/diagnostics/batch_probe.py [n_rows_timing] [n_queries_timing]
"""
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402
import oracle as O  # noqa: E402

SEED = 0x1715C0DE


def ternary(seed):
    return O.encode(O.gen_mask_rows(seed, 0, 1)[0], O.gen_mask_rows(seed, 1, 1)[0])


def describe(got, exp):
    bad = np.argwhere(got != exp)
    qs = np.unique(bad[:, 0]).tolist()
    rows = np.unique(bad[:, 1])
    return f"mismatch {len(bad)}/{got.size}; queries {qs[:16]}; rows {len(rows)} first {rows[:8].tolist()}; cols {np.unique(bad[:, 2]).tolist()[:8]}"


def correctness():
    n = 1000
    shares = O.gen_share_rows(SEED, 0, n, threads=8)
    db = iris.Database(n, masks=False)
    db.append_shares(shares)
    ok_all = True
    for name, queries in (
        ("ternary x8", [ternary(100 + i) for i in range(8)]),
        ("uniform x8", [O.gen_share_rows(200 + i, 0, 1)[0] for i in range(8)]),
        ("ternary x5", [ternary(300 + i) for i in range(5)]),
        ("mixed x20", [ternary(400 + i) if i % 3 else O.gen_share_rows(400 + i, 0, 1)[0] for i in range(20)]),
    ):
        try:
            engines = [iris.DistanceEngine(q) for q in queries]
            exp = np.stack([O.distance_batch(q, shares, threads=8) for q in queries])
            for rb, re in ((0, n), (3, 997), (256, 512), (300, 301)):
                out = np.zeros((len(queries), re - rb, 31), np.uint16)
                iris.distances_batch(engines, db, rb, re, out)
                ok = np.array_equal(out, exp[:, rb:re])
                ok_all &= ok
                print(f"[{'PASS' if ok else 'FAIL'}] batch {name} rows[{rb}:{re}] {'' if ok else describe(out, exp[:, rb:re])}", flush=True)
        except Exception as e:  # noqa: BLE001
            traceback.print_exc()
            ok_all = False
            print(f"[FAIL] batch {name}: {e}", flush=True)
    db.close()
    # batched denominators
    masks = O.gen_mask_rows(SEED, 0, n, threads=8)
    dbm = iris.Database(n, shares=False)
    dbm.append_masks(masks)
    for nqm in (16, 5, 37):
        try:
            qms = [O.gen_mask_rows(900 + i, 1, 1)[0] for i in range(nqm)]
            engines = [iris.MasksEngine(q) for q in qms]
            exp = np.stack([O.masks_batch(q, masks, threads=8) for q in qms])
            for rb, re in ((0, n), (3, 997), (256, 512), (300, 301)):
                out = np.zeros((nqm, re - rb, 31), np.uint16)
                iris.denominators_batch(engines, dbm, rb, re, out)
                ok = np.array_equal(out, exp[:, rb:re])
                ok_all &= ok
                print(f"[{'PASS' if ok else 'FAIL'}] batch denominators x{nqm} rows[{rb}:{re}] {'' if ok else describe(out, exp[:, rb:re])}", flush=True)
        except Exception as e:  # noqa: BLE001
            traceback.print_exc()
            ok_all = False
            print(f"[FAIL] batch denominators x{nqm}: {e}", flush=True)
    dbm.close()
    return ok_all


def timing_masks(n, nq):
    stream = torch.cuda.Stream()
    db = iris.Database(n, shares=False)
    db.generate(SEED, 0, n)
    db.set_stream(stream.cuda_stream)
    out = torch.empty((nq, n, 31), dtype=torch.int16, device="cuda")
    qms = [O.gen_mask_rows(950 + i, 1, 1)[0] for i in range(nq)]
    engines = [iris.MasksEngine(q) for q in qms]
    for _ in range(2):
        iris.denominators_batch(engines, db, 0, n, out)
    db.synchronize()
    iters = 5
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(iters):
        iris.denominators_batch(engines, db, 0, n, out)
    e.record(stream)
    db.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"denominators batch: n={n} Q={nq} {ms:.3f} ms -> {n * nq / (ms * 1e-3):.3e} cmp/s; issued "
          f"{2 * n * nq * 32 * 12800 / (ms * 1e-3) / 1e15:.3f} Pop/s", flush=True)
    got = out.cpu().numpy().view(np.uint16)
    ok = all(np.array_equal(got[qi, i], O.masks_batch(qms[qi], O.gen_mask_rows(SEED, i, 1))[0]) for qi in (0, nq - 1) for i in (0, 255, 256, n - 1))
    print("  sampled parity:", ok, flush=True)
    db.close()


def timing(n, nq):
    stream = torch.cuda.Stream()
    db = iris.Database(n, masks=False)
    db.generate(SEED, 0, n)
    db.set_stream(stream.cuda_stream)
    out = torch.empty((nq, n, 31), dtype=torch.int16, device="cuda")
    for name, queries in (("ternary(s8, 2 products)", [ternary(500 + i) for i in range(nq)]),
                          ("uniform(u8, 3 products)", [O.gen_share_rows(600 + i, 0, 1)[0] for i in range(nq)])):
        engines = [iris.DistanceEngine(q) for q in queries]
        for _ in range(2):
            iris.distances_batch(engines, db, 0, n, out)
        db.synchronize()
        iters = 5
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(iters):
            iris.distances_batch(engines, db, 0, n, out)
        e.record(stream)
        db.synchronize()
        ms = s.elapsed_time(e) / iters
        prods = 2 if name.startswith("ternary") else 3
        macs_issued = n * nq * 32 * 12800 * prods
        macs_useful = n * nq * 31 * 12800 * prods
        print(f"{name}: n={n} Q={nq} {ms:.3f} ms -> {n * nq / (ms * 1e-3):.3e} cmp/s; issued {2 * macs_issued / (ms * 1e-3) / 1e15:.3f} Pop/s,"
              f" useful {2 * macs_useful / (ms * 1e-3) / 1e15:.3f} Pop/s", flush=True)
        # sampled parity at size
        idx = [0, 255, 256, n // 2, n - 1]
        got = out.cpu().numpy().view(np.uint16)
        ok = all(np.array_equal(got[qi, i], O.distance_batch(queries[qi], O.gen_share_rows(SEED, i, 1))[0]) for qi in (0, nq // 2, nq - 1) for i in idx)
        print("  sampled parity:", ok, flush=True)
    # int8 peak of this box with the library GEMM (denominator for tensor-pipe utilisation)
    try:
        a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
        b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            torch._int_mm(a, b)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print(f"torch._int_mm 8192^3: {ms:.3f} ms -> {2 * 8192**3 / (ms * 1e-3) / 1e15:.3f} Pop/s (int8 library peak on this box)", flush=True)
    except Exception as ex:  # noqa: BLE001
        print("int8 library GEMM unavailable:", ex, flush=True)


if __name__ == "__main__":
    ok = correctness()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    if ok or "--force-timing" in sys.argv:
        timing_masks(n, nq)
        timing(n, nq)
    sys.exit(0 if ok else 1)
