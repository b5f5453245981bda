"""Sustained (seconds, power-limited) vs burst timing of the batched distances GEMM, 64 ternary queries x rows.

    [IRIS_B200_DIAG_LIB=1 IRIS_BATCH_CLUSTERS=72] python tests/diagnostics/batch_sustained.py [rows] [seconds]

Prints ms per launch, useful int8 Pop/s and the median SM clock for a burst (3 launches after a 2 s pause) and for the
steady state (launches back to back for `seconds`, the last ten timed).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402


def sm_clock():
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:  # noqa: BLE001
        return lambda: 0


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    nq = 64
    clk = sm_clock()
    db = iris.Database(rows, masks=False)
    db.generate(0x1715C0DE, 0, rows)
    tq = np.random.default_rng(7000).integers(0, 2**64, size=(nq, 400), dtype=np.uint64)
    des, _ = iris.engines_from_templates(tq, masks=False)
    out = torch.empty((nq, rows, 31), dtype=torch.int16, device="cuda")
    stream = torch.cuda.Stream()
    db.set_stream(stream.cuda_stream)

    def launch():
        iris.distances_batch(des, db, 0, rows, out)

    def timed(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(n):
            launch()
        e.record(stream)
        clocks = []
        while not e.query():
            clocks.append(clk())
            time.sleep(0.002)
        db.synchronize()
        return s.elapsed_time(e) / n, float(np.median(clocks)) if clocks else 0.0

    launch()
    db.synchronize()
    time.sleep(2.0)
    ms, mhz = timed(3)
    pops = 2 * rows * nq * 31 * 12800 * 2 / (ms * 1e-3) / 1e15
    print(f"burst     : {ms:7.2f} ms  {pops:.3f} useful int8 Pop/s ({pops / 4.5:.3f} of nominal 4.5)  SM {mhz:.0f} MHz", flush=True)
    t0 = time.time()
    while time.time() - t0 < seconds:
        for _ in range(5):
            launch()
        db.synchronize()
    ms, mhz = timed(10)
    pops = 2 * rows * nq * 31 * 12800 * 2 / (ms * 1e-3) / 1e15
    print(f"sustained : {ms:7.2f} ms  {pops:.3f} useful int8 Pop/s ({pops / 4.5:.3f} of nominal 4.5)  SM {mhz:.0f} MHz", flush=True)


if __name__ == "__main__":
    main()
