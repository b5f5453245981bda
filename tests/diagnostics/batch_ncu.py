"""Two launches of the batched distances GEMM (64 ternary queries x rows) for an ncu capture.

    ncu -k regex:batch_distances --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        python tests/diagnostics/batch_ncu.py [rows]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
nq = 64
db = iris.Database(rows, masks=False)
db.generate(0x1715C0DE, 0, rows)
tq = np.random.default_rng(7000).integers(0, 2**64, size=(nq, 400), dtype=np.uint64)
des, _ = iris.engines_from_templates(tq, masks=False)
out = torch.empty((nq, rows, 31), dtype=torch.int16, device="cuda")
for _ in range(2):
    iris.distances_batch(des, db, 0, rows, out)
    db.synchronize()
print("algorithmic bytes read per launch:", rows * 25600)
