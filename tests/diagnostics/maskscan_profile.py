"""Per-role wait-time profile of the denominators-only scan (IRIS_M4_VARIANT with flag 256): one launch, CTA 0 prints
the cycles each warp spent in its waits.  python tests/diagnostics/maskscan_profile.py [rows]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
db = iris.Database(rows, shares=False)
db.generate(0x1715C0DE, 0, rows)
me = iris.MasksEngine(np.random.default_rng(5).integers(0, 2**64, size=200, dtype=np.uint64))
den = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
iris.match(None, me, db, 0, rows, None, den)
db.synchronize()
torch.cuda.synchronize()
