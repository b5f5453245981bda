"""Time the batched entry points for batch sizes around the accumulator-tile boundaries (left-over queries are sent to
the single-query scans, iris_abi.cu).  python tests/diagnostics/batch_tail_bench.py [rows]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
stream = torch.cuda.Stream()
db = iris.Database(rows)
db.generate(0x1715C0DE, 0, rows)
db.set_stream(stream.cuda_stream)
tmpl = np.random.default_rng(5).integers(0, 2**64, size=(64, 400), dtype=np.uint64)
des, mes = iris.engines_from_templates(tmpl)
out = torch.empty((64, rows, 31), dtype=torch.int16, device="cuda")


def timed(fn, iters=3):
    fn()
    db.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(iters):
        fn()
    e.record(stream)
    db.synchronize()
    return s.elapsed_time(e) / iters


for q in (1, 2, 8, 9, 10, 16, 17, 26, 27, 32, 64):
    td = timed(lambda: iris.distances_batch(des[:q], db, 0, rows, out[:q]))
    tn = timed(lambda: iris.denominators_batch(mes[:q], db, 0, rows, out[:q]))
    print(f"Q={q:2d}: distances {td:7.3f} ms ({td / q:6.3f} per query)   denominators {tn:7.3f} ms ({tn / q:6.3f} per query)", flush=True)
