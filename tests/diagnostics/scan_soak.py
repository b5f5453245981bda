"""Soak test of the single-query scans on a shard with shares and masks: random row ranges, ternary and uniform queries,
fused / distances-only launches, every launch compared on the device with the CUDA-core kernels.

    python tests/diagnostics/scan_soak.py [rows] [seconds]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 90.0
    rng = np.random.default_rng(20261019)
    db = iris.Database(rows)
    db.generate(0x1715C0DE, 0, rows)
    gd = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    gn = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    rd = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    rn = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    t0, launches, checked = time.time(), 0, 0
    while time.time() - t0 < seconds:
        tmpl = rng.integers(0, 2**64, size=400, dtype=np.uint64)
        qm = tmpl[200:].copy()
        if rng.random() < 0.5:
            q = iris.encode(tmpl[:200].copy(), qm)                       # ternary: the signed two-product path
        else:
            q = rng.integers(0, 2**16, size=12800, dtype=np.uint16)      # uniform u16: three products
        de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
        db.check_distances_simt(q, 0, rows, rd)
        db.check_denominators_simt(qm, 0, rows, rn)
        for _ in range(6):
            rb = int(rng.integers(0, rows - 1))
            re = int(rng.integers(rb + 1, rows + 1)) if rng.random() < 0.6 else min(rows, rb + int(rng.integers(1, 1500)))
            gd.fill_(0x5A5A)
            gn.fill_(0x5A5A)
            fused = rng.random() < 0.6
            iris.match(de, me if fused else None, db, rb, re, gd[: re - rb], gn[: re - rb] if fused else None)
            db.synchronize()
            ok = torch.equal(gd[: re - rb], rd[rb:re]) and (not fused or torch.equal(gn[: re - rb], rn[rb:re]))
            if not ok:
                print(f"MISMATCH rows [{rb},{re}) fused={fused}", flush=True)
                sys.exit(1)
            if re - rb < rows and not (bool((gd[re - rb :] == 0x5A5A).all()) and bool((gn[re - rb :] == 0x5A5A).all())):
                print(f"OVERRUN rows [{rb},{re})", flush=True)
                sys.exit(1)
            launches += 1
            checked += re - rb
        de.close()
        me.close()
    print(f"soak ok: {launches} launches, {checked:,} rows compared in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
