"""Soak test of the denominators-only scan: many launches over random row ranges and queries, every launch compared on
the device with the CUDA-core kernel (a rare barrier-protocol race would show up as a mismatch or a watchdog trap).

    python tests/diagnostics/maskscan_soak.py [rows] [seconds]
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
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
    rng = np.random.default_rng(20261018)
    db = iris.Database(rows, shares=False)
    db.generate(0x1715C0DE, 0, rows)
    got = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    ref = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    t0, launches, checked = time.time(), 0, 0
    while time.time() - t0 < seconds:
        qm = rng.integers(0, 2**64, size=200, dtype=np.uint64)
        if rng.random() < 0.2:
            qm &= rng.integers(0, 2**64, size=200, dtype=np.uint64)      # sparser query masks now and then
        me = iris.MasksEngine(qm)
        db.check_denominators_simt(qm, 0, rows, ref)
        for _ in range(8):
            kind = rng.random()
            if kind < 0.3:
                rb, re = 0, rows
            elif kind < 0.6:
                rb = int(rng.integers(0, rows - 1))
                re = int(rng.integers(rb + 1, rows + 1))
            else:                                                     # short ranges: one to a few tile pairs
                rb = int(rng.integers(0, rows - 1))
                re = min(rows, rb + int(rng.integers(1, 2000)))
            got.fill_(0x5A5A)
            reps = int(rng.integers(1, 4))
            for _ in range(reps):                                     # back-to-back launches on the same barriers
                me.batch_process(got[: re - rb], db, rb, re)
            db.synchronize()
            if not torch.equal(got[: re - rb], ref[rb:re]):
                bad = (got[: re - rb] != ref[rb:re]).any(dim=1).nonzero()[:5].flatten().tolist()
                print(f"MISMATCH rows [{rb},{re}) first bad {bad}", flush=True)
                sys.exit(1)
            if re - rb < rows and not bool((got[re - rb :] == 0x5A5A).all()):
                print(f"OVERRUN rows [{rb},{re})", flush=True)
                sys.exit(1)
            launches += reps
            checked += re - rb
        me.close()
        # the batched path: groups of four masks on mask_scan_fp4_multi_kernel, left-overs on the single-query kernel
        nq = int(rng.integers(2, 10))
        qms = rng.integers(0, 2**64, size=(nq, 200), dtype=np.uint64)
        mes = [iris.MasksEngine(q) for q in qms]
        rb = int(rng.integers(0, rows - 1))
        re = min(rows, rb + int(rng.integers(1, max(2, rows // nq))))
        bout = got.reshape(-1)[: nq * (re - rb) * 31].view(nq, re - rb, 31)
        bout.fill_(0x5A5A)
        iris.denominators_batch(mes, db, rb, re, bout)
        db.synchronize()
        for k in range(nq):
            db.check_denominators_simt(qms[k], rb, re, ref[: re - rb])
            db.synchronize()
            if not torch.equal(bout[k], ref[: re - rb]):
                print(f"BATCH MISMATCH query {k} of {nq} rows [{rb},{re})", flush=True)
                sys.exit(1)
        for m in mes:
            m.close()
        launches += (nq + 3) // 4
        checked += nq * (re - rb)
    print(f"soak ok: {launches} launches, {checked:,} rows compared in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
