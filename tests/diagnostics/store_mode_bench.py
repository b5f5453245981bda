"""Kernel-only times of the fused scan with per-row results (device outputs) and in search mode (no per-row results),
1 M rows.  IRIS_B200_LIB=<path> selects an A/B build of the library (IRIS_STORE_MODE)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
db = iris.Database(rows)
db.generate_shares(0x1715C0DE, 0, 1, 0, rows)
t = np.random.default_rng(3).integers(0, 2**64, size=400, dtype=np.uint64)
de, me = iris.DistanceEngine.from_template(t[:200].copy(), t[200:].copy()), iris.MasksEngine(t[200:].copy())
dd = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
dn = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
stream = torch.cuda.Stream()
db.set_stream(stream.cuda_stream)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    db.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(n):
        fn()
    e.record(stream)
    db.synchronize()
    return s.elapsed_time(e) / n


full = timed(lambda: iris.match(de, me, db, 0, rows, dd, dn))
dist = timed(lambda: iris.match(de, None, db, 0, rows, dd, None))
t0 = time.perf_counter()
for _ in range(20):
    iris.match_min(de, me, db, 0, rows)
search = (time.perf_counter() - t0) / 20 * 1e3
print(f"{os.environ.get('IRIS_B200_LIB', 'product')}: fused+results {full:.3f} ms, distances+results {dist:.3f} ms, "
      f"search mode (host-timed, incl. 16 B D2H) {search:.3f} ms", flush=True)
