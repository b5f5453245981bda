"""Performance guards at 1.3x the measured times of an idle B200 (boxes of the pool differ by 5-10 %), so that a later
change which falls off the fast path -- a kernel variant switch left on, a serialising sync, a lost overlap, a lost
tail overlap of the chunked calls -- fails a test instead of only moving a bench number.  Times are CUDA-event means over
back-to-back launches after a warm-up; each figure is the best of three attempts (a neighbour on a shared box must not
fail the suite).  The HBM-bound scans are held to 1.3x outright; the tensor-bound batched kernels follow the SM clock,
which the 1 kW cap moves between 1.2 and 1.97 GHz depending on what ran before, so their allowance is 1.3x scaled by
max clock / the clock sampled while measuring (and that clock is part of the failure message)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE


@pytest.fixture(scope="module")
def ctx():
    import torch

    import mpc_iris_code_b200 as iris

    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("not enough free HBM for the 1 M-row shard")
    rows = 1_000_000
    stream = torch.cuda.Stream()
    db = iris.Database(rows)
    db.generate(SEED, 0, rows)
    db.set_stream(stream.cuda_stream)
    tmpl = np.random.default_rng(11).integers(0, 2**64, size=(16, 400), dtype=np.uint64)
    yield iris, torch, stream, db, rows, tmpl
    db.close()


_CLOCK = {"mhz": None}


def _sm_mhz(kind="now"):
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        if kind == "max":
            return pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:  # noqa: BLE001
        return None


def _clock_allowance():
    """max clock / sampled clock (>= 1): how much slower a clock-bound kernel may legitimately be right now."""
    now, top = _CLOCK["mhz"], _sm_mhz("max")
    return max(1.0, top / now) if now and top else 1.0


def _mean_ms(torch, stream, db, fn, warm, iters, attempts=3):
    import time

    best = float("inf")
    for _ in range(attempts):
        time.sleep(0.5)                    # let the power-averaging window recover between attempts
        for _ in range(warm):
            fn()
        db.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(iters):
            fn()
        e.record(stream)
        mhz = _sm_mhz()                    # sampled while the timed launches are running
        db.synchronize()
        ms = s.elapsed_time(e) / iters
        if ms < best:
            best, _CLOCK["mhz"] = ms, mhz
    return best


def test_single_query_scans_stay_near_the_hbm_rate(ctx):
    iris, torch, stream, db, rows, tmpl = ctx
    de = iris.DistanceEngine.from_template(tmpl[0, :200].copy(), tmpl[0, 200:].copy())
    me = iris.MasksEngine(tmpl[0, 200:].copy())
    dist = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
    den = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
    fused = _mean_ms(torch, stream, db, lambda: iris.match(de, me, db, 0, rows, dist, den), 3, 10)
    dists = _mean_ms(torch, stream, db, lambda: iris.match(de, None, db, 0, rows, dist, None), 3, 10)
    clk = _CLOCK["mhz"]
    assert fused < 5.1, f"fused scan {fused:.2f} ms per 1 M rows (measured: 3.9), SM {clk} MHz"
    assert dists < 4.7, f"distances-only scan {dists:.2f} ms per 1 M rows (measured: 3.6), SM {clk} MHz"
    masks = _mean_ms(torch, stream, db, lambda: iris.match(None, me, db, 0, rows, None, den), 50, 100)
    lim, clk = 1.3 * 0.29 * _clock_allowance(), _CLOCK["mhz"]    # this one follows the SM clock
    assert masks < lim, f"denominators-only scan {masks:.3f} ms per 1 M rows (0.28-0.30 at full clock; limit {lim:.3f}), SM {clk} MHz"


def test_chunked_calls_keep_their_tail_overlap(ctx):
    # the reference's 20 000-row calls (src/main.rs:427-430) on a caller-supplied stream: 4.0 ms per 1 M rows with the
    # tails overlapped (= one full-range call), 7.7 ms strictly serial
    iris, torch, stream, db, rows, tmpl = ctx
    de = iris.DistanceEngine.from_template(tmpl[1, :200].copy(), tmpl[1, 200:].copy())
    me = iris.MasksEngine(tmpl[1, 200:].copy())
    dist = torch.empty((rows, 31), dtype=torch.int16, device="cuda")
    den = torch.empty((rows, 31), dtype=torch.int16, device="cuda")

    def chunked():
        for b in range(0, rows, 20_000):
            iris.match(de, me, db, b, b + 20_000, dist[b:b + 20_000], den[b:b + 20_000])

    ms = _mean_ms(torch, stream, db, chunked, 1, 3)
    assert ms < 5.2, f"fused scan in 20 000-row calls {ms:.2f} ms per 1 M rows (measured: 4.0), SM {_CLOCK['mhz']} MHz"


def test_batched_paths_stay_on_the_tensor_kernels(ctx):
    iris, torch, stream, db, rows, tmpl = ctx
    n = 200_000
    des, mes = iris.engines_from_templates(tmpl)
    out = torch.empty((16, n, 31), dtype=torch.int16, device="cuda")
    dists = _mean_ms(torch, stream, db, lambda: iris.distances_batch(des, db, 0, n, out), 2, 5)
    lim, clk = 1.3 * 1.6 * _clock_allowance(), _CLOCK["mhz"]
    assert dists < lim, f"16 queries x 200 k rows, distances: {dists:.2f} ms (1.6 at full clock; limit {lim:.2f}), SM {clk} MHz"
    masks = _mean_ms(torch, stream, db, lambda: iris.denominators_batch(mes, db, 0, n, out), 2, 5)
    lim, clk = 1.3 * 0.45 * _clock_allowance(), _CLOCK["mhz"]
    assert masks < lim, f"16 masks x 200 k rows, denominators: {masks:.2f} ms (0.45 at full clock; limit {lim:.2f}), SM {clk} MHz"
