"""Does the fused scan's time depend on WHAT the database holds?  Times 50 back-to-back launches of the result-writing
fused scan and of the search-mode scan over (a) uniform u16 shares (one party's view, what `value` in bench.py runs on)
and (b) plaintext encodings {0, 1, 0xFFFF} (the n = 1 sharing the scored search runs on), with the SM clock and the
board power sampled.  python tests/diagnostics/data_power_bench.py [rows]"""
import os
import sys
import threading
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
db = iris.Database(rows)
t = np.random.default_rng(3).integers(0, 2**64, size=400, dtype=np.uint64)
de, me = iris.DistanceEngine.from_template(t[:200].copy(), t[200:].copy()), iris.MasksEngine(t[200:].copy())
dd = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
dn = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
pair = torch.zeros(2, dtype=torch.int64, device="cuda")
stream = torch.cuda.Stream()
db.set_stream(stream.cuda_stream)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    db.synchronize()
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            time.sleep(0.01)

    th = threading.Thread(target=sampler)
    th.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(n):
        fn()
    e.record(stream)
    db.synchronize()
    stop.set()
    th.join()
    tail = samples[len(samples) // 2:] or [(0, 0)]
    return s.elapsed_time(e) / n, sorted(x[0] for x in tail)[len(tail) // 2], sorted(x[1] for x in tail)[len(tail) // 2]


for name, parties in (("uniform u16 shares", 0), ("plaintext encodings", 1), ("uniform u16 shares", 0), ("plaintext encodings", 1)):
    db.clear()
    if parties:
        db.generate_shares(0x1715C0DE, 0, 1, 0, rows)
    else:
        db.generate(0x1715C0DE, 0, rows)
    db.synchronize()
    time.sleep(1.0)
    full = timed(lambda: iris.match(de, me, db, 0, rows, dd, dn))
    time.sleep(1.0)
    search = timed(lambda: iris.match_min_async(de, me, db, 0, rows, pair))
    print(f"{name:20s}: fused+results {full[0]:.3f} ms ({full[1]} MHz, {full[2]:.0f} W)   search mode {search[0]:.3f} ms "
          f"({search[1]} MHz, {search[2]:.0f} W)", flush=True)
