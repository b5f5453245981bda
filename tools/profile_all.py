"""Launches every product kernel a few times at bench-like sizes; the command ncu is pointed at for profiles/.

    python tools/profile_all.py [rows] [batch_rows]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    brows = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    db = iris.Database(rows)
    db.generate(0x1715C0DE, 0, rows)
    rng = np.random.default_rng(5)
    tmpl = rng.integers(0, 2**64, size=(1 + 64, 400), dtype=np.uint64)      # wire Templates {pattern, mask}
    qm = tmpl[0, 200:].copy()
    de, me = iris.DistanceEngine.from_template(tmpl[0, :200].copy(), qm), iris.MasksEngine(qm)
    dist = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
    den = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
    for _ in range(3):
        iris.match(de, me, db, 0, rows, dist, den)         # scan_kernel<1,1>
    for _ in range(3):
        iris.match(de, None, db, 0, rows, dist, None)      # scan_kernel<1,0>
    for _ in range(3):
        iris.match(None, me, db, 0, rows, None, den)       # mask_scan_kernel
    db.synchronize()
    nq = 64
    big = torch.empty((nq, brows, 31), dtype=torch.int16, device="cuda")
    tern, mes = iris.engines_from_templates(tmpl[1:])
    unif = [iris.DistanceEngine(x) for x in rng.integers(0, 2**16, size=(nq, 12800), dtype=np.uint16)]
    for _ in range(3):
        iris.distances_batch(tern, db, 0, brows, big)      # batch_distances_kernel<1>
    for _ in range(3):
        iris.distances_batch(unif, db, 0, brows, big)      # batch_distances_kernel<0>
    for _ in range(3):
        iris.denominators_batch(mes, db, 0, brows, big)    # mask_scan_fp4_multi_kernel x 16 (four masks per pass)
    db.synchronize()
    for _ in range(3):
        best = iris.match_min(de, me, db, 0, rows)          # scan_kernel<1,1,1,search> + final_min_kernel
    print("min/argmin:", best)
    iris.match(de, me, db, 0, rows, dist, den)
    db.synchronize()
    print("combine:", iris.combine_min([dist], den))        # combine_decode_kernel + final_min_kernel
    print("launches:", iris.launch_count())


if __name__ == "__main__":
    main()
