"""Batched denominators (64 query masks x rows) in isolation: burst and sustained timings with the SM clock sampled
while the kernels run.  IRIS_BATCHDEN=i8 selects the int8 GEMM kernel.  python tests/diagnostics/batchden_bench.py [rows]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402
import pynvml  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
stream = torch.cuda.Stream()
db = iris.Database(rows, shares=False)
db.generate(0x1715C0DE, 0, rows)
db.set_stream(stream.cuda_stream)
mes = [iris.MasksEngine(m) for m in np.random.default_rng(5).integers(0, 2**64, size=(64, 200), dtype=np.uint64)]
out = torch.empty((64, rows, 31), dtype=torch.int16, device="cuda")


def timed(iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(iters):
        iris.denominators_batch(mes, db, 0, rows, out)
    e.record(stream)
    clk = []
    while not e.query():
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        time.sleep(0.002)
    db.synchronize()
    return s.elapsed_time(e) / iters, float(np.median(clk)) if clk else 0.0


iris.denominators_batch(mes, db, 0, rows, out)
db.synchronize()
time.sleep(2.0)
ms, mhz = timed(3)
print(f"mode={os.environ.get('IRIS_BATCHDEN', 'fp4x4')} burst     : {ms:7.3f} ms per 64 x {rows} ({64 * rows / ms / 1e6:6.2f} e9 cmp/s) at {mhz:.0f} MHz", flush=True)
ms, mhz = timed(40)
print(f"mode={os.environ.get('IRIS_BATCHDEN', 'fp4x4')} sustained : {ms:7.3f} ms per 64 x {rows} ({64 * rows / ms / 1e6:6.2f} e9 cmp/s) at {mhz:.0f} MHz", flush=True)
