"""Denominators-only scan (BASELINE configs[2]): steady-state timing of the kernel variants + full parity.

    python tests/diagnostics/maskscan_bench.py [rows] [variant ...]

Each variant (IRIS_M4_VARIANT / IRIS_MASKSCAN value, see iris_maskscan4.cu) runs in its own process because the
library reads the switch once.  Every row is compared with the CUDA-core cross-check kernel; the timing is taken
after 0.3 s of back-to-back launches so the SM clock has settled.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child(rows: int) -> None:
    import numpy as np
    import torch

    import mpc_iris_code_b200 as iris

    stream = torch.cuda.Stream()
    db = iris.Database(rows, shares=False)
    db.generate(0x1715C0DE, 0, rows)
    db.set_stream(stream.cuda_stream)
    qm = np.random.default_rng(5).integers(0, 2**64, size=200, dtype=np.uint64)
    me = iris.MasksEngine(qm)
    den = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    iris.match(None, me, db, 0, rows, None, den)
    db.synchronize()
    ref = torch.zeros((rows, 31), dtype=torch.int16, device="cuda")
    db.check_denominators_simt(qm, 0, rows, ref)
    db.synchronize()
    ok = bool(torch.equal(den, ref))
    t0 = time.time()
    while time.time() - t0 < 0.3:
        for _ in range(20):
            iris.match(None, me, db, 0, rows, None, den)
        db.synchronize()
    best = []
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(200):
            iris.match(None, me, db, 0, rows, None, den)
        e.record(stream)
        db.synchronize()
        best.append(s.elapsed_time(e) / 200)
    ms = min(best)
    # sustained: 1.5 s back to back with the SM clock and the board power sampled (the kernel follows the clock)
    import threading

    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            time.sleep(0.02)

    th = threading.Thread(target=sampler)
    th.start()
    t0 = time.time()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    s.record(stream)
    while time.time() - t0 < 1.5:
        for _ in range(100):
            iris.match(None, me, db, 0, rows, None, den)
        n += 100
        db.synchronize()
    e.record(stream)
    db.synchronize()
    stop.set()
    th.join()
    tail = samples[len(samples) // 2:]
    mhz = sorted(x[0] for x in tail)[len(tail) // 2]
    watts = sorted(x[1] for x in tail)[len(tail) // 2]
    sustained = s.elapsed_time(e) / n
    print(f"variant={os.environ.get('IRIS_M4_VARIANT', '-')}/{os.environ.get('IRIS_MASKSCAN', 'f')} rows={rows} "
          f"parity={'ok' if ok else 'FAIL'} ms={ms:.4f} (runs {' '.join(f'{b:.4f}' for b in best)}) "
          f"{rows * 1662 / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic; sustained 1.5 s: {sustained:.4f} ms, {mhz} MHz, {watts:.0f} W",
          flush=True)


def main() -> None:
    if os.environ.get("_MASKSCAN_CHILD"):
        child(int(sys.argv[1]))
        return
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    variants = sys.argv[2:] or ["20"]
    for v in variants:
        env = dict(os.environ, _MASKSCAN_CHILD="1")
        if v in ("i8", "smem"):
            env["IRIS_MASKSCAN"] = v
        else:
            env["IRIS_M4_VARIANT"] = v
        subprocess.run([sys.executable, os.path.abspath(__file__), str(rows)], env=env, check=False, timeout=300)


if __name__ == "__main__":
    main()
