#!/usr/bin/env python
"""bench.py -- headline benchmark of the matching hot path (BASELINE.json configs[1]).

One step = one fused scan (distances + denominators, all 31 rotations) of ONE query against the
HBM-resident shard of this rank.  N=1: 1 M synthetic templates on one B200.  N>1: one process per
GPU (torchrun), every rank holds its own 1 M-row shard (rows are independent: weak scaling, no
data-path collective), value = rows of all ranks / max-over-ranks time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--impl b200|reference]

Prints ONE JSON line (rank 0).  `value` is kernel-only throughput with inputs resident in HBM;
`e2e` is the same metric through the public engine API with HOST buffers (query H2D, prepare,
scan, results D2H inside the timed region); `roofline` is the scan kernel against the measured
HBM copy peak; `cpu_baseline` times the CPU oracle (a C port of the reference's generic path --
the Rust crate cannot be built in this image) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x1715C0DE
BYTES_PER_ROW_FUSED = 25600 + 1600 + 2 * 62  # algorithmic bytes per comparison
METRIC = "template_comparisons_per_sec"
UNIT = "comparisons/s"
WORKLOAD = ("1 query (31 rotations) vs {rows} synthetic EncodedBits+Bits templates per GPU, fused "
            "distances+denominators (BASELINE configs[1])")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--rows", type=int, default=1_000_000, help="database rows per GPU (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --rows-total rows split over the ranks instead of --rows per rank")
    ap.add_argument("--rows-total", type=int, default=4_000_000, help="database rows in total (strong scaling)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs (denominators only, batched GEMM)")
    return ap.parse_args()


def random_templates(seed: int, n: int) -> np.ndarray:
    """n synthetic wire Templates ([n][400] u64 = {pattern[200], mask[200]}, uniform bits like the reference's
    `rng.gen::<Template>()`, src/template.rs:67-74).  Plain numpy: the GPU arm never touches oracle/."""
    return np.random.default_rng(seed).integers(0, 2**64, size=(n, 400), dtype=np.uint64)


def make_template():
    t = random_templates(0xBEEF, 1)[0]
    return t[:200].copy(), t[200:].copy()


def make_query_cpu():
    """(encoded query, mask) for the CPU legs -- encode() by the oracle (the CPU restatement being timed)."""
    import oracle as O

    pattern, mask = make_template()
    return O.encode(pattern, mask), mask


# ----------------------------------------------------------------------------------- CPU arm
def time_oracle(rows: int, threads: int, q, qm, seconds: float = 0.0):
    """Runs fused passes (distances + denominators) of the CPU oracle over a `rows`-row sample until `seconds`
    of wall time have accumulated (at least one pass); returns (passes, total seconds)."""
    import oracle as O

    shares = O.gen_share_rows(SEED, 0, rows, threads=threads)
    masks = O.gen_mask_rows(SEED, 0, rows, threads=threads)
    rot = O.distance_rotations(q)
    mrot = O.mask_rotations(qm)
    out_d = np.empty((rows, 31), np.uint16)
    out_n = np.empty((rows, 31), np.uint16)
    passes, total = 0, 0.0
    while passes == 0 or total < seconds:
        t0 = time.perf_counter()
        O.distance_batch_prepared(rot, shares, out_d, threads)
        O.masks_batch_prepared(mrot, masks, out_n, threads)
        total += time.perf_counter() - t0
        passes += 1
    return passes, total


def cpu_baseline(target_seconds: float, q, qm):
    cores = os.cpu_count() or 1
    rows = 4096 * cores                            # bounded sample: 1.8 GB at 16 cores, far beyond the CPU caches
    time_oracle(256 * cores, cores, q, qm)         # warm-up (page faults, OpenMP pool)
    passes, t_all = time_oracle(rows, cores, q, qm, target_seconds)
    rows1 = 2048
    p1, t_one = time_oracle(rows1, 1, q, qm, min(3.0, target_seconds / 3))
    return {
        "value": rows * passes / t_all,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{passes} passes over {rows} of the workload's rows x 31 rotations, distances+denominators, C port of "
                  f"src/arch/generic.rs + engine loops (gcc -O3 -march=native, OpenMP static over rows = rayon par_iter), "
                  f"{t_all:.1f} s of wall time on {cores} threads",
        "single_thread_value": rows1 * p1 / t_one,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    q, qm = make_query_cpu()
    cores = os.cpu_count() or 1
    probe = 256 * cores
    time_oracle(probe, cores, q, qm)
    t = time_oracle(probe, cores, q, qm)[1]
    total = args.steps + args.warmup
    per_step_target = min(4.0, 150.0 / max(total, 1))
    rows = int(max(probe, min(200_000, probe * per_step_target / max(t, 1e-6))))
    import oracle as O

    shares = O.gen_share_rows(SEED, 0, rows, threads=cores)
    masks = O.gen_mask_rows(SEED, 0, rows, threads=cores)
    rot, mrot = O.distance_rotations(q), O.mask_rotations(qm)
    out_d = np.empty((rows, 31), np.uint16)
    out_n = np.empty((rows, 31), np.uint16)

    def step():
        O.distance_batch_prepared(rot, shares, out_d, cores)
        O.masks_batch_prepared(mrot, masks, out_n, cores)
        O.combine_min(out_d, out_n)        # decode_distance + running min (src/main.rs:597-621): what the GPU arm's e2e returns

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    sample = (f"each step = {rows} rows of the workload (bounded sample): distances + denominators on all {cores} host threads, "
              f"then the coordinator's decode + min on one thread like src/main.rs:611-621; C port of the reference generic "
              f"path (Rust toolchain absent, crate not buildable here)")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(rows=args.rows), "rows_per_step_sample": rows,
                   "path": "generic (x86-64, auto-vectorised); SVE path not available"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.active = threading.Event()
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    try:
                        mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:  # noqa: BLE001
                        mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.005)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def nvml_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:  # noqa: BLE001
            return local_rank
    return local_rank


# ----------------------------------------------------------------------------------- secondary configs
_NVML = {"nv": None, "h": None, "last_mhz": None}


def _time_ms(stream, fn, sync, warmup=2, iters=5, cool_s=0.0):
    """Mean device time of fn (CUDA events on the launch stream); the SM clock is sampled while the timed launches
    run (the 1 kW cap moves it between ~1.2 and 1.97 GHz depending on what ran before) and left in _NVML['last_mhz']."""
    import torch

    for _ in range(warmup):
        fn()
    sync()
    if cool_s:
        time.sleep(cool_s)      # let the power-averaging window recover: a burst figure, like the library GEMM's
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(iters):
        fn()
    e.record(stream)
    clk = []
    if _NVML["h"] is not None:
        while not e.query():
            try:
                clk.append(_NVML["nv"].nvmlDeviceGetClockInfo(_NVML["h"], _NVML["nv"].NVML_CLOCK_SM))
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.001)
    sync()
    _NVML["last_mhz"] = float(np.median(clk)) if clk else None
    return s.elapsed_time(e) / iters


def hbm_peak_gbs():
    """The roofline denominator: the driver's measured copy bandwidth, else the profiling recipe's fallback."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _steady_ms(stream, fn, sync, settle_s=0.3, iters=100):
    """Mean device time of a sub-millisecond kernel in steady state: launch back to back for settle_s, then time
    `iters` launches with CUDA events on the launch stream (SM clock sampled meanwhile, left in _NVML['last_mhz'])."""
    t0 = time.time()
    while time.time() - t0 < settle_s:
        for _ in range(20):
            fn()
        sync()
    return _time_ms(stream, fn, sync, warmup=0, iters=iters)


def secondary_configs(iris, db, stream, rows, de, me, d_dist, d_den):
    """BASELINE configs[2] (denominators only) and configs[3] (64 queries as a dense int8 GEMM) on the same shard,
    kernel-only, reported next to the headline so the driver's own run carries them."""
    import torch

    out = {}
    # BASELINE configs[2]: denominators-only sweep, 100 k - 4 M masks (a masks-only shard; uniform bits, one query mask).
    # These are 0.03 - 1.1 ms kernels: each size is timed over 100 launches after 0.3 s of back-to-back launches so the
    # SM clock has settled (the kernel is within ~10 % of its HBM floor but still follows the SM clock).
    peak = hbm_peak_gbs()[0]
    sweep_rows = [n for n in (100_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000) if n <= 4 * rows]
    mdb = iris.Database(max(sweep_rows), device=db.device, shares=False)
    mdb.generate(SEED, 0, max(sweep_rows))
    mdb.set_stream(stream.cuda_stream)
    m_den = torch.empty((max(sweep_rows), 31), dtype=torch.int16, device="cuda")
    sweep = []
    for n in sweep_rows:
        # consecutive launches scan DIFFERENT n-row windows of the shard (offsets are multiples of a 256-row pair tile),
        # cycling over all of it: 100 k masks are 160 MB, the size of the L2, and a loop over one window would be
        # served partly from cache
        windows = max(1, max(sweep_rows) // n)
        state = {"i": 0}
        out_n = m_den[:n]

        def one(n=n, windows=windows, state=state, out_n=out_n):
            b = (state["i"] % windows) * n // 256 * 256
            state["i"] += 1
            iris.match(None, me, mdb, b, b + n, None, out_n)

        ms = _steady_ms(stream, one, mdb.synchronize)
        sweep.append({"rows": n, "ms": ms, "comparisons_per_s": n / (ms * 1e-3),
                      "algorithmic_GBps": n * 1662 / (ms * 1e-3) / 1e9,
                      "frac_of_hbm_peak": n * 1662 / (ms * 1e-3) / 1e9 / peak, "sm_mhz": _NVML["last_mhz"],
                      "windows": windows})
    mdb.close()
    del m_den
    one_m = next((x for x in sweep if x["rows"] == 1_000_000), sweep[-1])
    out["denominators_only_1q"] = dict(one_m, note="4-bit tcgen05 operands expanded into tensor memory (DESIGN.md 5.3); "
                                       "algorithmic bytes = 1 600 B read + 62 B written per row")
    out["denominators_sweep"] = sweep
    # the reference's own calling pattern: batch_process on 20 000-row chunks (src/main.rs:427-430, 512-515), here as
    # back-to-back asynchronous calls on the shard's stream with device outputs
    chunk = 20_000

    def chunked():
        for c in range(0, rows, chunk):
            e_ = min(rows, c + chunk)
            iris.match(de, me, db, c, e_, d_dist[c:e_], d_den[c:e_])

    ms_serial = _time_ms(stream, chunked, db.synchronize, warmup=1, iters=5)
    # the same on the library's own stream (host-timed)
    db.set_stream(None)
    chunked()
    db.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        chunked()
    db.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    db.set_stream(stream.cuda_stream)
    out["fused_in_20000_row_calls_1q"] = {"ms": ms, "calls": (rows + chunk - 1) // chunk, "comparisons_per_s": rows / (ms * 1e-3),
                                          "ms_on_a_caller_stream": ms_serial,
                                          "note": "157 tiles of 128 rows on 148 SMs per call; the next call starts on the SMs the "
                                                  "previous call's tail leaves idle (programmatic dependent launch), on the "
                                                  "library's stream and on a caller-supplied stream alike"}
    # the coordinator's side of the same pattern: MasksEngine::batch_process on 20 000-row chunks (src/main.rs:512-515)
    def chunked_masks():
        for c in range(0, rows, chunk):
            e_ = min(rows, c + chunk)
            iris.match(None, me, db, c, e_, None, d_den[c:e_])

    ms_serial = _steady_ms(stream, chunked_masks, db.synchronize, settle_s=0.2, iters=20)
    db.set_stream(None)
    for _ in range(20):
        chunked_masks()
    db.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        chunked_masks()
    db.synchronize()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    db.set_stream(stream.cuda_stream)
    out["denominators_in_20000_row_calls_1q"] = {"ms": ms, "calls": (rows + chunk - 1) // chunk,
                                                 "comparisons_per_s": rows / (ms * 1e-3), "ms_on_a_caller_stream": ms_serial}
    ms = _time_ms(stream, lambda: iris.match(de, None, db, 0, rows, d_dist, None), db.synchronize, warmup=3, iters=10)
    out["distances_only_1q"] = {"ms": ms, "comparisons_per_s": rows / (ms * 1e-3),
                                "algorithmic_GBps": rows * 25662 / (ms * 1e-3) / 1e9, "sm_mhz": _NVML["last_mhz"]}
    # the general path: a uniform-u16 query (what the reference's criterion bench draws, src/arch/mod.rs:56-61) needs
    # all three limb products; the headline's ternary encode(Template) query takes the two-product path
    qu = np.random.default_rng(77).integers(0, 2**16, size=12800, dtype=np.uint16)
    du = iris.DistanceEngine(qu, device=db.device)
    ms = _time_ms(stream, lambda: iris.match(du, me, db, 0, rows, d_dist, d_den), db.synchronize, warmup=3, iters=10)
    out["fused_uniform_u16_query_1q"] = {"ms": ms, "comparisons_per_s": rows / (ms * 1e-3),
                                         "algorithmic_GBps": rows * BYTES_PER_ROW_FUSED / (ms * 1e-3) / 1e9,
                                         "limb_products": 3, "sm_mhz": _NVML["last_mhz"]}
    du.close()
    # BASELINE configs[0] on the GPU: the reference's criterion grids (src/arch/mod.rs:22-72) -- dot_u16 on every pair of
    # 31 INDEPENDENT uniform vectors x 100 000 rows, dot_bool on 31 x 1 000 and 1 x 100 000 -- through the arch-level
    # entry points (iris_dot_*_batch_resident: the 31 vectors take the slots of the 31 rotations); Elements = a x b
    grid = {}
    try:
        n_b = 100_000
        gdb = iris.Database(n_b, device=db.device)
        gdb.generate(SEED + 1, 0, n_b)
        gdb.set_stream(stream.cuda_stream)
        a16 = np.random.default_rng(31).integers(0, 2**16, size=(31, 12800), dtype=np.uint16)
        am = np.random.default_rng(32).integers(0, 2**64, size=(31, 200), dtype=np.uint64)
        g_out = torch.empty((n_b, 31), dtype=torch.int16, device="cuda")
        for name, fn, n_a, nb in (("dot_u16_31x100000", lambda: iris.dot_u16_batch(a16, gdb, out=g_out), 31, n_b),
                                  ("dot_bool_31x100000", lambda: iris.dot_bool_batch(am, gdb, out=g_out), 31, n_b)):
            for _ in range(3):
                fn()
            gdb.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            gdb.synchronize()
            dt = (time.perf_counter() - t0) / 10
            grid[name] = {"ms_per_grid": dt * 1e3, "elements_per_s": n_a * nb / dt,
                          "note": "31 host vectors uploaded + prepared per call, rows resident; out = [rows][31] in HBM"}
        # parity of one column against a plain numpy restatement of src/arch/generic.rs:11-16 on rows read back
        back = gdb.read_shares(0, 64).astype(np.uint64)
        want = ((back @ a16.astype(np.uint64).T) & 0xFFFF).astype(np.uint16)
        iris.dot_u16_batch(a16, gdb, out=g_out)
        gdb.synchronize()
        grid["dot_u16_parity_ok"] = bool(np.array_equal(g_out[:64].cpu().numpy().view(np.uint16), want))
        gdb.close()
        del g_out
    except Exception as ex:  # noqa: BLE001
        grid = {"error": repr(ex)}
    out["criterion_grids_on_gpu"] = grid
    # the reference's LITERAL signature: batch_process(out, db) with `db` a HOST slice (src/lib.rs:42-52) -- rows are
    # uploaded, re-tiled and scanned in a double-buffered pipeline; PCIe-bound by construction, so the yardstick is a
    # bare pinned host->device copy of the same bytes
    hs_rows = 100_000
    h_rows = torch.empty((hs_rows, 12800), dtype=torch.int16).pin_memory()
    h_rows.random_(-32768, 32767)
    h_out = torch.empty((hs_rows, 31), dtype=torch.int16).pin_memory()
    d_tmp = torch.empty((hs_rows, 12800), dtype=torch.int16, device="cuda")
    rows_np, out_np = h_rows.numpy().view(np.uint16), h_out.numpy().view(np.uint16)
    de.batch_process(out_np, rows_np)
    t0 = time.perf_counter()
    for _ in range(3):
        de.batch_process(out_np, rows_np)
    hs = (time.perf_counter() - t0) / 3
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        d_tmp.copy_(h_rows, non_blocking=True)
        torch.cuda.synchronize()
    bare = (time.perf_counter() - t0) / 3
    out["host_slice_batch_process_1q"] = {"rows": hs_rows, "ms": hs * 1e3, "comparisons_per_s": hs_rows / hs,
                                          "host_to_device_GBps": hs_rows * 25600 / hs / 1e9,
                                          "bare_h2d_GBps": hs_rows * 25600 / bare / 1e9,
                                          "frac_of_bare_h2d": bare / hs,
                                          "note": "DistanceEngine::batch_process(out, db) with db and out in (pinned) host memory"}
    del h_rows, h_out, d_tmp
    # int8 library GEMM on this box: the measured tensor-core denominator
    a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
    b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        torch._int_mm(a, b)
    e.record()
    torch.cuda.synchronize()
    lib_pops = 2 * 8192**3 / (s.elapsed_time(e) / 20 * 1e-3) / 1e15
    # the same library GEMM back to back for 2.5 s: what the 1 kW cap leaves of it (the sustained denominator)
    t0 = time.time()
    while time.time() - t0 < 2.5:
        for _ in range(50):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
    s.record()
    for _ in range(50):
        torch._int_mm(a, b)
    e.record()
    torch.cuda.synchronize()
    lib_pops_sustained = 2 * 8192**3 / (s.elapsed_time(e) / 50 * 1e-3) / 1e15
    del a, b
    time.sleep(2.0)
    nq = 64
    tt = random_templates(7000, nq)
    tern = [iris.encode(tt[i, :200].copy(), tt[i, 200:].copy(), device=db.device) for i in range(nq)]   # encode() on the device
    unif = list(np.random.default_rng(8000).integers(0, 2**16, size=(nq, 12800), dtype=np.uint16))
    qms = [tt[i, 200:].copy() for i in range(nq)]
    big = torch.empty((nq, rows, 31), dtype=torch.int16, device="cuda")
    res = {"queries": nq, "rows": rows, "int8_library_gemm_Pops": lib_pops, "int8_library_gemm_Pops_sustained": lib_pops_sustained,
           "int8_nominal_Pops": 4.5,
           "timing": "3 launches after a 2 s pause (burst, like the 20-launch library GEMM); sm_mhz = median SM clock "
                     "while they ran; back to back for seconds the 1 kW cap pulls the clock to ~1.2-1.5 GHz"}
    for name, qs, prods in (("ternary", tern, 2), ("uniform_u16", unif, 3)):
        eng = [iris.DistanceEngine(x) for x in qs]
        ms = _time_ms(stream, lambda: iris.distances_batch(eng, db, 0, rows, big), db.synchronize, warmup=1, iters=3, cool_s=2.0)
        useful = 2 * rows * nq * 31 * 12800 * prods / (ms * 1e-3) / 1e15
        res[f"distances_{name}"] = {"ms": ms, "comparisons_per_s": rows * nq / (ms * 1e-3), "limb_products": prods,
                                    "useful_int8_Pops": useful, "frac_of_nominal": useful / 4.5,
                                    "frac_of_library_gemm": useful / lib_pops, "sm_mhz": _NVML["last_mhz"]}
        if name == "ternary":
            # sustained: back to back for 2.5 s (the 1 kW cap settles the clock), then ten timed launches
            t0 = time.time()
            while time.time() - t0 < 2.5:
                for _ in range(5):
                    iris.distances_batch(eng, db, 0, rows, big)
                db.synchronize()
            ms_s = _time_ms(stream, lambda: iris.distances_batch(eng, db, 0, rows, big), db.synchronize, warmup=0, iters=10)
            useful_s = 2 * rows * nq * 31 * 12800 * prods / (ms_s * 1e-3) / 1e15
            res["distances_ternary_sustained"] = {"ms": ms_s, "comparisons_per_s": rows * nq / (ms_s * 1e-3),
                                                  "useful_int8_Pops": useful_s, "frac_of_nominal": useful_s / 4.5,
                                                  "frac_of_library_gemm_sustained": useful_s / lib_pops_sustained,
                                                  "sm_mhz": _NVML["last_mhz"],
                                                  "timing": "ten launches after 2.5 s of back-to-back launches"}
        for x in eng:
            x.close()
    eng = [iris.MasksEngine(x) for x in qms]
    ms = _time_ms(stream, lambda: iris.denominators_batch(eng, db, 0, rows, big), db.synchronize, warmup=1, iters=3, cool_s=2.0)
    useful = 2 * rows * nq * 31 * 12800 / (ms * 1e-3) / 1e15
    res["denominators"] = {"ms": ms, "comparisons_per_s": rows * nq / (ms * 1e-3), "useful_Pops": useful,
                           "kernel": "mask_scan_fp4_multi_kernel: four query masks per pass over the database operand "
                                     "expanded into tensor memory (tcgen05.mma kind::mxf4, N = 128); the int8 GEMM kernel "
                                     "(IRIS_BATCHDEN=i8) takes 13.0 ms for the same work",
                           "frac_of_nominal_fp4": useful / 9.0, "sm_mhz": _NVML["last_mhz"]}
    both = res["distances_ternary"]["ms"] + res["denominators"]["ms"]
    res["distances_plus_denominators_ternary"] = {"ms": both, "comparisons_per_s": rows * nq / (both * 1e-3)}
    out["batched_64q_int8_gemm"] = res
    return out


# ----------------------------------------------------------------------------------- GPU arm
def plain_template_of_row(iris, row_id: int, device: int):
    """(pattern & mask, mask) of synthetic row `row_id`, read back from the GPU: a one-row shard is generated with the
    plaintext encoding (n_parties = 1) and decoded (encode() maps mask-2*(pattern&mask) -> 1, 0, 0xFFFF, src/lib.rs:16-26)."""
    with iris.Database(1, device=device) as d1:
        d1.generate_shares(SEED, 0, 1, row_id, 1)
        enc = d1.read_shares(0, 1)[0]
        mask = d1.read_masks(0, 1)[0]
    bits = (enc == 0xFFFF).astype(np.uint8)
    pattern = np.packbits(bits, bitorder="little").view("<u8").astype(np.uint64)
    return pattern, mask


def _bits_matrix(limbs):
    b = np.ascontiguousarray(limbs, dtype="<u8").view(np.uint8)
    return np.unpackbits(b, bitorder="little").reshape(64, 200)      # bit k = row k/200, column k%200 (src/bits.rs:44-57)


def _matrix_bits(m):
    return np.packbits(m.reshape(-1).astype(np.uint8), bitorder="little").view("<u8").astype(np.uint64)


def noisy_copy(pattern, mask, flips: int, rotation: int, seed: int):
    """A query Template close to (pattern, mask): `flips` pattern bits toggled, both matrices rotated by `rotation`
    columns (Bits::rotate, src/bits.rs:80-92: out[row][col] = in[row][(col - amount) mod 200])."""
    rng = np.random.default_rng(seed)
    pm = _bits_matrix(pattern).copy().reshape(-1)
    pm[rng.choice(pm.size, size=flips, replace=False)] ^= 1
    return (_matrix_bits(np.roll(pm.reshape(64, 200), rotation, axis=1)),
            _matrix_bits(np.roll(_bits_matrix(mask), rotation, axis=1)))


def plain_distance(qp, qm, p, m) -> float:
    """Template::distance in the clear (src/template.rs:43-64): min over rotations -15..=15 of the QUERY of the fractional
    Hamming distance under both masks -- plain numpy, the yardstick for the planted match (not the oracle)."""
    qpm, qmm, pm, mm = _bits_matrix(qp), _bits_matrix(qm), _bits_matrix(p), _bits_matrix(m)
    best = float("inf")
    for r in range(-15, 16):
        msk = np.roll(qmm, r, axis=1) & mm
        den = int(msk.sum())
        num = int(((np.roll(qpm, r, axis=1) ^ pm) & msk).sum())
        if den:
            best = min(best, num / den)
        elif num == 0:
            continue                                          # 0/0 = NaN, dropped by f64::min
    return best


def run_b200(args):
    import torch
    import torch.distributed as dist

    import mpc_iris_code_b200 as iris

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        try:
            # run this rank (and first-touch its pinned buffers) on the CPUs / NUMA node next to its GPU;
            # only for N > 1 -- at N = 1 the cpu_baseline leg needs every host core
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(nvml_index(local_rank)))
        except Exception:  # noqa: BLE001
            pass
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout while the communicator comes up (NCCL_DEBUG=VERSION/INFO);
        # stdout must carry exactly one JSON line, so point fd 1 at stderr until the communicators are up.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            host_group = dist.new_group(backend="gloo")      # host-side barriers that leave the GPUs idle
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    strong = args.scaling == "strong"
    rows = args.rows if not strong else iris.cluster_partition(args.rows_total, world, rank)[1] - iris.cluster_partition(args.rows_total, world, rank)[0]
    row0 = rank * rows if not strong else iris.cluster_partition(args.rows_total, world, rank)[0]
    n_total = rows * world if not strong else args.rows_total
    c5_rows = min(16_000_000 // world, 4_000_000)               # BASELINE configs[4]: 16 M rows over the ranks, 109 GB cap
    ss_rows = (4_000_000 + world - 1) // world                  # strong scaling at 4 M rows in total
    capacity = max(rows, c5_rows if not args.no_extras else 0, ss_rows if not args.no_extras else 0)

    # ---- the product's multi-GPU handle: this process owns one shard of a cluster that spans all ranks; the library
    # (not this script) gathers the per-query (min, argmin) pairs over its own NCCL communicator.
    cluster = iris.Cluster([local_rank], capacity)
    if world > 1:
        uid = [iris.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)                  # 128 bytes by any channel; torch.distributed is plumbing
        cluster.join(uid[0], rank, world)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    db = cluster.shard(0)[0]
    qp, qm = make_template()
    q = iris.encode(qp, qm, device=local_rank)        # encode(&template) on the device (src/lib.rs:16-26)
    stream = torch.cuda.Stream()
    cluster.generate(SEED, rows, first_row_id=row0, n_parties=0)   # uniform u16 shares: what one party holds
    db.set_stream(stream.cuda_stream)
    d_dist = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
    d_den = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
    de, me = iris.DistanceEngine(q, device=local_rank), iris.MasksEngine(qm, device=local_rank)

    sampler = ClockSampler(nvml_index(local_rank))
    sampler.start()
    if sampler.ok:
        _NVML["nv"], _NVML["h"] = sampler.nv, sampler.h

    # ---- kernel-only: inputs resident in HBM, results stay in HBM; the database (27.2 GB at 1 M rows)
    # is far larger than L2, so every step streams from DRAM.
    for _ in range(max(args.warmup, 3)):
        iris.match(de, me, db, 0, rows, d_dist, d_den)
    db.synchronize()
    barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = iris.launch_count()
    sampler.active.set()
    evs[0].record(stream)
    for i in range(args.steps):
        iris.match(de, me, db, 0, rows, d_dist, d_den)
        evs[i + 1].record(stream)
    db.synchronize()
    torch.cuda.synchronize()
    sampler.active.clear()
    launches = iris.launch_count() - launches0
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = float(np.mean([evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]))
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # the search-mode scan (the kernel behind `e2e`) on the SAME uniform shares, for comparison: the scan's time depends a
    # little on what the database holds (uniform u16 words cost more power than the encodings' {0, 1, 0xFFFF}: DESIGN.md 5.1)
    pair_u = torch.zeros(2, dtype=torch.int64, device="cuda")
    for _ in range(3):
        iris.match_min_async(de, me, db, 0, rows, pair_u, index_base=row0)
    db.synchronize()
    uev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    uev[0].record(stream)
    for _ in range(max(3, min(args.steps, 20))):
        iris.match_min_async(de, me, db, 0, rows, pair_u, index_base=row0)
    uev[1].record(stream)
    db.synchronize()
    search_kernel_ms_uniform = max_over_ranks(uev[0].elapsed_time(uev[1]) / max(3, min(args.steps, 20)))

    # ---- the engine API with HOST result buffers (pinned): per step the query and its mask go host->device, both
    # engines are prepared, the shard is scanned, and BOTH result arrays (124 B per row) come back over PCIe.
    q_pin = torch.from_numpy(q.view(np.int16).copy()).pin_memory()
    qm_pin = torch.from_numpy(qm.view(np.int64).copy()).pin_memory()
    h_dist = torch.empty((rows, 31), dtype=torch.int16).pin_memory()
    h_den = torch.empty((rows, 31), dtype=torch.int16).pin_memory()
    q_np, qm_np = q_pin.numpy().view(np.uint16), qm_pin.numpy().view(np.uint64)
    hd_np, hn_np = h_dist.numpy().view(np.uint16), h_den.numpy().view(np.uint16)

    def full_results_step(dist_only=False):
        e1 = iris.DistanceEngine(q_np, device=local_rank)
        e2 = None if dist_only else iris.MasksEngine(qm_np, device=local_rank)
        iris.match(e1, e2, db, 0, rows, hd_np, None if dist_only else hn_np)   # returns after the last D2H copy completed
        e1.close()
        if e2:
            e2.close()

    def timed_host_loop(fn, steps, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt)

    e2e_steps = max(3, min(args.steps, 20))
    full_s = timed_host_loop(full_results_step, e2e_steps)
    part_s = timed_host_loop(lambda: full_results_step(True), e2e_steps)
    # bare device->host copies of the same two arrays, all ranks at once: the PCIe ceiling of this box for that traffic
    cp_s = torch.cuda.Stream()

    def bare_d2h():
        with torch.cuda.stream(cp_s):
            h_dist.copy_(d_dist, non_blocking=True)
            h_den.copy_(d_den, non_blocking=True)
        cp_s.synchronize()

    d2h_s = timed_host_loop(bare_d2h, e2e_steps)

    # rows of the host result kept for the oracle spot check done inside the cpu_baseline leg (rank 0, N = 1)
    full_results_step()
    sample_idx = np.array([0, 127, 128, rows // 3, rows - 1])
    sample_d, sample_n = hd_np[sample_idx].copy(), hn_np[sample_idx].copy()
    del h_den, hn_np

    # ---- END TO END = the search a user of the cluster makes (iris_cluster_search through the C ABI): one wire Template
    # in host memory in, (min distance, row) in host memory out.  Every rank's shard now holds whole encodings of
    # synthetic Templates (the n = 1 sharing, so results MEAN something); per step the library copies the Template to its
    # GPU, prepares both engines, scans + decodes + reduces its shard on the device, all-gathers 16 bytes per shard over
    # NCCL and merges.  A noisy copy of the LAST row of the LAST rank is the query: the winner must come from rank N-1.
    db.synchronize()
    db.set_stream(None)
    cluster.generate(SEED, rows, first_row_id=row0, n_parties=1)
    cluster.set_index_base(row0)
    target = n_total - 1
    t_p, t_m = plain_template_of_row(iris, target, local_rank)
    s_p, s_m = noisy_copy(t_p, t_m, flips=1500, rotation=4, seed=1)
    expected = plain_distance(s_p, s_m, t_p, t_m)
    tq_pin = torch.from_numpy(np.concatenate([s_p, s_m]).view(np.int64).copy()).pin_memory()
    tq1 = tq_pin.numpy().view(np.uint64).reshape(1, 400)
    found = {}

    def search_step():
        md, mi = cluster.search(tq1)
        found["d"], found["i"] = float(md[0]), int(mi[0])

    sampler.active.set()
    search_s = timed_host_loop(search_step, e2e_steps)
    sampler.active.clear()
    sampler.stop_flag.set()
    # the kernel behind that call, device-timed on its own: the fused scan in SEARCH MODE (decode + running minimum in
    # the epilogue, no per-row results written), plus the 148-entry final reduction
    e1, e2 = iris.DistanceEngine.from_template(s_p, s_m, device=local_rank), iris.MasksEngine(s_m, device=local_rank)
    pair = torch.zeros(2, dtype=torch.int64, device="cuda")
    db.set_stream(stream.cuda_stream)
    for _ in range(3):
        iris.match_min_async(e1, e2, db, 0, rows, pair, index_base=row0)
    db.synchronize()
    sev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    sev[0].record(stream)
    for _ in range(e2e_steps):
        iris.match_min_async(e1, e2, db, 0, rows, pair, index_base=row0)
    sev[1].record(stream)
    db.synchronize()
    search_kernel_ms = max_over_ranks(sev[0].elapsed_time(sev[1]) / e2e_steps)
    db.set_stream(None)
    e1.close()
    e2.close()
    e2e_value = n_total * e2e_steps / search_s
    search_ok = bool(found["i"] == target and found["d"] == expected)

    # ---- secondary lines on every rank (skipped with --no-extras)
    extra_lines = {}
    if not args.no_extras:
        try:
            extra_lines = multi_gpu_extras(iris, cluster, rank, world, local_rank, ss_rows, c5_rows, timed_host_loop, host_group)
        except Exception as ex:  # noqa: BLE001
            extra_lines = {"error": repr(ex)}

    if rank != 0:
        cluster.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak_gbs()
    achieved = rows * BYTES_PER_ROW_FUSED / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_fused_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
        if traffic is not None and rows != 1_000_000:
            traffic = None                                     # the capture was taken at 1 M rows
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(rows=rows), "rows_per_gpu": rows, "rows_total": n_total,
                   "query": "ternary encode(random Template)",
                   "l2": "inputs larger than L2 (27.2 GB streamed per step at 1 M rows); no flush needed",
                   "sharding": "rows, one shard per rank (iris_cluster_* in the C ABI); no data-path collective; the search "
                               "all-gathers 16 bytes per shard over the library's NCCL communicator",
                   "value_data": "uniform u16 shares (one party's view)",
                   "e2e_data": "whole encodings of synthetic Templates (n = 1 sharing), query = noisy copy of the last row"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "scan_kernel<shares,masks>", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": rows * BYTES_PER_ROW_FUSED, "launch_ms": per_launch_ms},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3200, "d2h_bytes_per_step": 16, "steps": e2e_steps,
                "ms_per_query": search_s / e2e_steps * 1e3,
                "kernel": "scan_kernel<shares,masks,search>: the fused scan with decode_distance + running min/argmin in its "
                          "epilogue; no per-row results are written (worth 1.5-2 %), and the database of this leg holds "
                          "encodings {0, 1, 0xFFFF}, which cost less power than `value`'s uniform u16 shares (another 2-2.5 %: "
                          "kernel_only_ms_on_uniform_shares is the same kernel over the `value` data) -- which is why this "
                          "path can be FASTER than `value`",
                "kernel_only_ms": search_kernel_ms,
                "kernel_only_ms_on_uniform_shares": search_kernel_ms_uniform,
                "kernel_roofline": {"achieved_GBps": rows * 27200 / (search_kernel_ms * 1e-3) / 1e9,
                                    "frac_of_hbm_peak": rows * 27200 / (search_kernel_ms * 1e-3) / 1e9 / hbm_peak_gbs()[0],
                                    "algorithmic_bytes_per_launch": rows * 27200},
                "path": "iris_cluster_search (C ABI): wire Template from pinned host memory -> encode + both engines on the "
                        "device -> fused scan -> decode_distance + min/argmin on the device -> "
                        + ("all-gather of the shards' (min, argmin) pairs over NCCL -> " if world > 1 else "")
                        + "16 bytes to the host",
                "collective": "ncclAllGather of 16 bytes per rank inside the library" if world > 1 else "none (one shard)",
                "result": [found["d"], found["i"]], "expected": [expected, target],
                "winner_rank": world - 1, "parity_ok": search_ok},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    line["e2e_full_results"] = {
        "note": "the engine API with HOST result arrays, every rank copying its own results over its own PCIe link at the "
                "same time; the search above is the scored end-to-end path because in the deployed roles "
                "(src/main.rs:510-519, 597-621) the coordinator's denominators never leave its device and the decoded "
                "minimum is what a query returns",
        "both_arrays": {"comparisons_per_s": n_total * e2e_steps / full_s, "ms_per_query": full_s / e2e_steps * 1e3,
                        "d2h_bytes_per_step_per_gpu": rows * 124},
        "distances_only_participant": {"comparisons_per_s": n_total * e2e_steps / part_s, "ms_per_query": part_s / e2e_steps * 1e3,
                                       "d2h_bytes_per_step_per_gpu": rows * 62},
        "bare_concurrent_d2h": {"ms": d2h_s / e2e_steps * 1e3, "GBps_per_gpu": rows * 124 / (d2h_s / e2e_steps) / 1e9,
                                "GBps_aggregate": world * rows * 124 / (d2h_s / e2e_steps) / 1e9,
                                "note": "two cudaMemcpyAsync (62 MB each at 1 M rows) from HBM into pinned host memory per "
                                        "rank, all ranks at once, nothing else running: the ceiling for that traffic"},
    }
    line.update(extra_lines)
    if world == 1 and not args.no_extras:
        try:
            db.set_stream(stream.cuda_stream)
            cluster.generate(SEED, rows, first_row_id=row0, n_parties=0)
            line["extras"] = secondary_configs(iris, db, stream, rows, de, me, d_dist, d_den)
        except Exception as ex:  # noqa: BLE001
            line["extras"] = {"error": repr(ex)}
        try:
            line["participant_wire_1q"] = participant_wire(rows)
        except Exception as ex:  # noqa: BLE001
            line["participant_wire_1q"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline:
        # the only leg of this arm that touches oracle/: times the CPU port and uses it as the checker for the
        # sampled rows of the GPU result above
        import oracle as O

        q_cpu = O.encode(qp, qm)
        line["cpu_baseline"] = cpu_baseline(args.cpu_seconds, q_cpu, qm)
        ok = bool(np.array_equal(q_cpu, q))
        for k, i in enumerate(sample_idx):
            ok &= np.array_equal(sample_d[k], O.distance_batch(q_cpu, O.gen_share_rows(SEED, int(row0 + i), 1))[0])
            ok &= np.array_equal(sample_n[k], O.masks_batch(qm, O.gen_mask_rows(SEED, int(row0 + i), 1))[0])
        ok &= found["d"] == O.template_distance(s_p, s_m, *[x[0] for x in (O.gen_pattern_rows(SEED, target, 1), O.gen_mask_rows(SEED, target, 1))])
        line["config"]["sample_parity_ok"] = bool(ok)
    print(json.dumps(line), flush=True)
    cluster.close()
    if world > 1:
        dist.destroy_process_group()


def participant_wire(rows: int, requests: int = 5):
    """The reference's participant role on the wire (src/main.rs:384-452): bin/iris_participant (C++ over the C ABI) holds a
    synthetic share database in HBM; a client sends one 3 200-byte Template over loopback TCP and reads rows x 62 bytes
    back until EOF, as the reference coordinator does (src/main.rs:486-504).  Times are the client's."""
    import socket
    import subprocess

    from mpc_iris_code_b200 import build

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    proc = subprocess.Popen([build.PARTICIPANT_PATH, "--synthetic", str(rows), "--bind", f"127.0.0.1:{port}",
                             "--max-requests", str(requests)], stderr=subprocess.PIPE)
    times, firsts = [], []
    try:
        import select

        deadline = time.time() + 120.0                       # never let a stuck front-end hang the bench
        err_fd, err_buf = proc.stderr.fileno(), b""
        while True:
            if proc.poll() is not None:
                raise RuntimeError("iris_participant exited early")
            if time.time() > deadline:
                raise RuntimeError("iris_participant did not start listening within 120 s")
            if select.select([err_fd], [], [], 1.0)[0]:
                err_buf += os.read(err_fd, 4096)
                if b"Listening on" in err_buf:
                    break
        template = random_templates(1, 1)[0].tobytes()
        view = memoryview(bytearray(8 << 20))
        for _ in range(requests):
            t0 = time.perf_counter()
            with socket.create_connection(("127.0.0.1", port), timeout=30.0) as c:
                c.settimeout(30.0)
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
            times.append(time.perf_counter() - t0)
            firsts.append(first)
            if got != rows * 62:
                raise RuntimeError(f"short reply: {got} of {rows * 62} bytes")
        proc.wait(timeout=60)
    finally:
        if proc.poll() is None:
            proc.kill()
    t = sorted(times[1:])[len(times[1:]) // 2]                # the first request warms the connection path up
    return {"rows": rows, "requests": requests, "ms_per_request": t * 1e3, "best_ms": min(times) * 1e3,
            "first_byte_ms": sorted(firsts[1:])[len(firsts[1:]) // 2] * 1e3, "comparisons_per_s": rows / t,
            "wire_GBps": rows * 62 / t / 1e9, "reply_bytes": rows * 62, "batch_rows": 20_000,
            "note": "request -> EOF seen by a single-connection loopback client (Python recv_into); the scan itself takes "
                    "~3.7 ms per 1 M rows, the rest is one TCP stream carrying 62 MB (DESIGN.md 6.1)"}


def multi_gpu_extras(iris, cluster, rank, world, local_rank, ss_rows, c5_rows, timed_host_loop, host_group):
    """Lines every rank takes part in: strong scaling of the search at 4 M rows in total, and BASELINE configs[4]
    (64 queries against 16 M rows row-sharded over the ranks) through iris_cluster_search, each with a planted match in
    the last rank's shard; then (N > 1) rank 0 alone drives all N GPUs through ONE in-process cluster handle."""
    import torch
    import torch.distributed as dist

    out = {}
    # -- strong scaling: 4 M rows in total
    total = 4_000_000
    b, e = iris.cluster_partition(total, world, rank)
    cluster.generate(SEED, e - b, first_row_id=b, n_parties=1)
    cluster.set_index_base(b)
    t_p, t_m = plain_template_of_row(iris, total - 1, local_rank)
    s_p, s_m = noisy_copy(t_p, t_m, flips=1400, rotation=-7, seed=2)
    tq = np.concatenate([s_p, s_m]).reshape(1, 400).copy()
    got = {}

    def step():
        md, mi = cluster.search(tq)
        got["r"] = (float(md[0]), int(mi[0]))

    steps = 10
    s = timed_host_loop(step, steps, warm=2)
    out["strong_scaling_search"] = {
        "rows_total": total, "rows_per_gpu": e - b, "ms_per_query": s / steps * 1e3, "comparisons_per_s": total * steps / s,
        "result": list(got["r"]), "expected": [plain_distance(s_p, s_m, t_p, t_m), total - 1],
        "parity_ok": bool(got["r"] == (plain_distance(s_p, s_m, t_p, t_m), total - 1)),
        "path": "iris_cluster_search, 1 query, fixed 4 M-row database split over the ranks (HBM-bound fused scan per shard)"}
    # -- BASELINE configs[0]'s size (the reference's own CPU-runnable case): 1 query vs 100 000 rows in total -- the
    # latency floor of a search (engine preparation, launches, the gather, 16 bytes back) rather than a bandwidth figure
    small = 100_000
    b0, e0 = iris.cluster_partition(small, world, rank)
    cluster.generate(SEED, e0 - b0, first_row_id=b0, n_parties=1)
    cluster.set_index_base(b0)
    t_p0, t_m0 = plain_template_of_row(iris, small - 1, local_rank)
    s_p0, s_m0 = noisy_copy(t_p0, t_m0, flips=1350, rotation=5, seed=4)
    tq0 = np.concatenate([s_p0, s_m0]).reshape(1, 400).copy()
    got0 = {}

    def step0():
        md, mi = cluster.search(tq0)
        got0["r"] = (float(md[0]), int(mi[0]))

    s0 = timed_host_loop(step0, 50, warm=5)
    exp0 = (plain_distance(s_p0, s_m0, t_p0, t_m0), small - 1)
    out["search_100k_rows"] = {
        "rows_total": small, "rows_per_gpu": e0 - b0, "ms_per_query": s0 / 50 * 1e3, "comparisons_per_s": small * 50 / s0,
        "result": list(got0["r"]), "expected": list(exp0), "parity_ok": bool(got0["r"] == exp0),
        "path": "iris_cluster_search, 1 query, BASELINE configs[0]'s 100 000 rows split over the ranks (latency floor)"}
    # -- BASELINE configs[4]: 64 queries vs 16 M rows row-sharded (4 M rows = 109 GB per GPU at most)
    nq = 64
    total5 = c5_rows * world
    cluster.generate(SEED, c5_rows, first_row_id=rank * c5_rows, n_parties=1)
    cluster.set_index_base(rank * c5_rows)
    tq64 = random_templates(9000, nq)
    planted = {5: total5 - 1, 40: total5 - c5_rows // 2}          # both in the last rank's shard
    exp = {}
    for k, r in planted.items():
        p_, m_ = plain_template_of_row(iris, r, local_rank)
        tq64[k, :200], tq64[k, 200:] = noisy_copy(p_, m_, flips=1200 + k, rotation=(k % 9) - 4, seed=k)
        exp[k] = (plain_distance(tq64[k, :200], tq64[k, 200:], p_, m_), r)
    res = {}

    def step64():
        res["md"], res["mi"] = cluster.search(tq64)

    bsteps = 3
    s5 = timed_host_loop(step64, bsteps, warm=1)
    ok = all((float(res["md"][k]), int(res["mi"][k])) == exp[k] for k in planted)
    out["batched_search_64q_sharded"] = {
        "config": "BASELINE configs[4]" if world > 1 else "BASELINE configs[3] shape at 4 M rows on one GPU",
        "queries": nq, "rows_total": total5, "rows_per_gpu": c5_rows,
        "note": (f"16 M rows / {world} = {c5_rows} rows per GPU" if 16_000_000 // world <= c5_rows else
                 f"capped at 4 M rows per GPU (109 GB of shares + masks): {total5} rows in total at N = {world} instead of 16 M"),
        "ms_per_batch": s5 / bsteps * 1e3, "comparisons_per_s": nq * total5 * bsteps / s5,
        "h2d_bytes_per_step": nq * 3200, "d2h_bytes_per_step": nq * 16,
        "useful_int8_Pops_per_gpu": 2 * c5_rows * nq * 31 * 12800 * 2 / (s5 / bsteps) / 1e15,
        "path": "iris_cluster_search (C ABI), 64 wire Templates: batched int8 tcgen05 GEMM distances + 4-bit denominators, "
                "decode + min/argmin on the device per 524 288-row slice, "
                + ("ncclAllGather of 64 x 16 bytes per rank, merge" if world > 1 else "merge"),
        "planted": {str(k): {"expected": list(exp[k]), "found": [float(res["md"][k]), int(res["mi"][k])]} for k in planted},
        "parity_ok": bool(ok)}
    if world == 1:
        return out
    # -- the full result vectors of all ranks gathered on every rank: scan into the rank's slot, then one grouped
    # ncclBroadcast per rank over NVLink ("compute, then collective" -- the in-process handle below stores the same
    # vectors straight into one GPU from the scan epilogues instead)
    try:
        rows_g = 1_000_000
        cluster.generate(SEED, rows_g, first_row_id=rank * rows_g, n_parties=1)
        cluster.set_index_base(rank * rows_g)
        p_, m_ = plain_template_of_row(iris, world * rows_g - 1, local_rank)
        qg = iris.encode(p_, m_, device=local_rank)
        gd = torch.empty((world * rows_g, 31), dtype=torch.int16, device="cuda")
        gn = torch.empty((world * rows_g, 31), dtype=torch.int16, device="cuda")
        sg = timed_host_loop(lambda: cluster.match_allgather(qg, m_, gd, gn), 5, warm=2)
        mdg, mig = iris.combine_min([gd], gn, device=local_rank)
        out["full_vectors_allgather"] = {
            "rows_total": world * rows_g, "ms_per_query": sg / 5 * 1e3, "comparisons_per_s": world * rows_g * 5 / sg,
            "nvlink_bytes_received_per_gpu": (world - 1) * rows_g * 124,
            "path": "iris_cluster_match_allgather: fused scan into the rank's slot of two [N_total][31] arrays, then "
                    "ncclBroadcast of every rank's block (grouped) -- every rank ends up with all vectors",
            "parity_ok": bool((mdg, mig) == (0.0, world * rows_g - 1))}
        del gd, gn
    except Exception as ex:  # noqa: BLE001
        out["full_vectors_allgather"] = {"error": repr(ex)}
    # -- ONE process, ONE handle, N GPUs: rank 0 drives every GPU of the box through iris_cluster_* (what a Rust host
    # does); the other ranks release their GPUs and wait at a HOST barrier (gloo) so that nothing of theirs runs.
    cluster.close()                                   # every rank frees its shard (and leaves the NCCL communicator)
    torch.cuda.synchronize()
    dist.barrier(group=host_group)
    if rank == 0:
        try:
            rows1 = 1_000_000
            with iris.Cluster(list(range(world)), rows1 * world) as c:
                n = rows1 * world
                c.generate(SEED, n, n_parties=1)
                t_p, t_m = plain_template_of_row(iris, n - 1, 0)
                s_p, s_m = noisy_copy(t_p, t_m, flips=1300, rotation=3, seed=3)
                tq = np.concatenate([s_p, s_m]).reshape(1, 400).copy()
                for _ in range(3):
                    md, mi = c.search(tq)
                t0 = time.perf_counter()
                for _ in range(10):
                    md, mi = c.search(tq)
                s1 = (time.perf_counter() - t0) / 10
                exp1 = (plain_distance(s_p, s_m, t_p, t_m), n - 1)
                # full result vectors of all shards into ONE array on GPU 0: every other GPU's scan stores over NVLink
                dd = torch.empty((n, 31), dtype=torch.int16, device="cuda:0")
                dn = torch.empty((n, 31), dtype=torch.int16, device="cuda:0")
                for _ in range(2):
                    c.match_template(s_p, s_m, dd, dn)
                t0 = time.perf_counter()
                for _ in range(5):
                    c.match_template(s_p, s_m, dd, dn)
                s2 = (time.perf_counter() - t0) / 5
                md2, mi2 = iris.combine_min([dd], dn, device=0)
                out["in_process_cluster"] = {
                    "gpus": world, "rows_total": n, "host_threads": world,
                    "search": {"ms_per_query": s1 * 1e3, "comparisons_per_s": n / s1, "result": [float(md[0]), int(mi[0])],
                               "expected": list(exp1), "parity_ok": bool((float(md[0]), int(mi[0])) == exp1),
                               "gather": "peer stores of 16 bytes per shard into GPU 0 over NVLink, merge kernel on GPU 0"},
                    "match_into_one_gpu_array": {
                        "ms_per_query": s2 * 1e3, "comparisons_per_s": n / s2,
                        "nvlink_bytes_per_query": (n - rows1) * 124,
                        "note": "each shard's scan epilogue stores its [rows][31] distances and denominators straight into "
                                "GPU 0's arrays (peer stores); reduced afterwards on GPU 0 with iris_combine_min",
                        "parity_ok": bool((md2, mi2) == exp1)}}
        except Exception as ex:  # noqa: BLE001
            out["in_process_cluster"] = {"error": repr(ex)}
    dist.barrier(group=host_group)
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
