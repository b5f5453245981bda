"""Wire-level throughput of bin/iris_participant over loopback: one synthetic shard, a few requests, rows/s seen by
the client (the reference coordinator's view, src/main.rs:486-504).

    python tests/diagnostics/participant_bench.py [rows] [requests] [extra participant args...]
"""
import os
import socket
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpc_iris_code_b200 import build  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    requests = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    extra = sys.argv[3:]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    proc = subprocess.Popen([build.PARTICIPANT_PATH, "--synthetic", str(rows), "--bind", f"127.0.0.1:{port}",
                             "--max-requests", str(requests), *extra], stderr=subprocess.PIPE, text=True)
    try:
        while "Listening on" not in proc.stderr.readline():
            assert proc.poll() is None
        template = np.random.default_rng(1).integers(0, 2**64, size=400, dtype=np.uint64).tobytes()
        buf = bytearray(8 << 20)
        view = memoryview(buf)
        for r in range(requests):
            t0 = time.perf_counter()
            with socket.create_connection(("127.0.0.1", port)) as c:
                c.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 8 << 20)
                c.sendall(template)
                got, first = 0, None
                while True:
                    n = c.recv_into(view)
                    if n == 0:
                        break
                    if first is None:
                        first = time.perf_counter() - t0
                    got += n
            dt = time.perf_counter() - t0
            assert got == rows * 62, (got, rows * 62)
            print(f"request {r}: {dt * 1e3:8.2f} ms, first byte after {first * 1e3:6.2f} ms, {rows / dt:.3e} rows/s, "
                  f"{got / dt / 1e9:.2f} GB/s on the wire", flush=True)
        proc.wait(timeout=60)
    finally:
        if proc.poll() is None:
            proc.kill()


if __name__ == "__main__":
    main()
