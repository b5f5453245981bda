"""Kernel-only timing of the three scan modes on a synthetic resident database.

    python tests/diagnostics/quick_bench.py [n_rows] [iters]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpc_iris_code_b200 as iris  # noqa: E402
import oracle as O  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    torch.cuda.init()
    stream = torch.cuda.Stream()
    db = iris.Database(n)
    db.generate(0x1715C0DE, 0, n)
    db.set_stream(stream.cuda_stream)
    q = O.encode(O.gen_mask_rows(5, 0, 1)[0], O.gen_mask_rows(5, 1, 1)[0])
    qm = O.gen_mask_rows(5, 1, 1)[0]
    de, me = iris.DistanceEngine(q), iris.MasksEngine(qm)
    dist = torch.empty((n, 31), dtype=torch.int16, device="cuda")
    den = torch.empty((n, 31), dtype=torch.int16, device="cuda")
    for name, a, b, bytes_per_row in (("fused", de, me, 27200 + 124), ("distances", de, None, 25600 + 62), ("denominators", None, me, 1600 + 62)):
        for _ in range(3):
            iris.match(a, b, db, 0, n, dist, den)
        db.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for s, e in evs:
            s.record(stream)
            iris.match(a, b, db, 0, n, dist, den)
            e.record(stream)
        db.synchronize()
        ms = np.array([s.elapsed_time(e) for s, e in evs])
        gbs = n * bytes_per_row / (ms.mean() * 1e-3) / 1e9
        print(f"{name:13s} n={n} mean {ms.mean():.3f} ms min {ms.min():.3f} ms -> {n / (ms.mean() * 1e-3):.3e} cmp/s, {gbs:.1f} GB/s algorithmic", flush=True)
    # spot parity on a few sampled rows
    idx = np.array([0, 1, 127, 128, n // 2, n - 1])
    exp_d = np.stack([O.distance_batch(q, O.gen_share_rows(0x1715C0DE, int(i), 1))[0] for i in idx])
    exp_n = np.stack([O.masks_batch(qm, O.gen_mask_rows(0x1715C0DE, int(i), 1))[0] for i in idx])
    iris.match(de, me, db, 0, n, dist, den)
    db.synchronize()
    got_d = dist.cpu().numpy().view(np.uint16)[idx]
    got_n = den.cpu().numpy().view(np.uint16)[idx]
    print("sample parity:", np.array_equal(got_d, exp_d), np.array_equal(got_n, exp_n), flush=True)


if __name__ == "__main__":
    main()
