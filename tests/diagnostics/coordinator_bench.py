"""Whole-query latency of the two front-ends on one box: P iris_participant processes with synthetic shards and one
iris_coordinator with a synthetic masks file, all on GPU 0, talking over loopback (reference deployment:
src/main.rs:384-640).  The shares are random, so the answer is meaningless; the timing is not.

    python tests/diagnostics/coordinator_bench.py [rows] [parties] [requests] [n_gpus]
"""
import os
import socket
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpc_iris_code_b200 import build  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    parties = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    requests = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    n_gpus = int(sys.argv[4]) if len(sys.argv) > 4 else 1     # participant i runs on GPU (i + 1) % n_gpus, the coordinator on 0
    procs, ports = [], []
    with tempfile.TemporaryDirectory() as tmp:
        masks = os.path.join(tmp, "mpc.masks")
        rng = np.random.default_rng(3)
        with open(masks, "wb") as f:
            left = rows
            while left:
                n = min(left, 100_000)
                f.write(rng.integers(0, 2**64, size=(n, 200), dtype=np.uint64).tobytes())
                left -= n
        try:
            for i in range(parties):
                with socket.socket() as s:
                    s.bind(("127.0.0.1", 0))
                    ports.append(s.getsockname()[1])
                procs.append(subprocess.Popen(
                    [build.PARTICIPANT_PATH, "--synthetic", str(rows), "--seed", str(100 + i), "--bind", f"127.0.0.1:{ports[-1]}",
                     "--max-requests", str(requests), "--device", str((i + 1) % n_gpus)], stderr=subprocess.PIPE, text=True))
            for p in procs:
                while "Listening on" not in p.stderr.readline():
                    assert p.poll() is None
            coord = subprocess.Popen([build.COORDINATOR_PATH, "--masks", masks, "--requests", str(requests),
                                      *[f"127.0.0.1:{p}" for p in ports]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            t_prev = None
            for r in range(requests):
                line = coord.stdout.readline()
                now = time.perf_counter()
                if t_prev is not None:
                    dt = now - t_prev
                    print(f"request {r}: {dt * 1e3:8.2f} ms for {rows} rows x {parties} parties -> {rows / dt:.3e} rows/s   [{line.strip()}]", flush=True)
                t_prev = now
            err = coord.stderr.read()
            assert coord.wait(timeout=60) == 0
            print("\n".join(line for line in err.splitlines() if line.startswith("Timing:")), flush=True)
            for p in procs:
                p.wait(timeout=60)
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()


if __name__ == "__main__":
    main()
