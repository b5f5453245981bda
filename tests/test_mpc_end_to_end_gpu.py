"""Whole-protocol check on the GPU: the reference's deployment (src/main.rs:384-640) played by the two C++
front-ends over the C ABI -- three `iris_participant` processes, each with one additive share of the encoded
database resident in HBM, and one `iris_coordinator` with the masks -- must find the same closest entry, at the
same f64 distance, as the plaintext definition (src/template.rs:43-64) evaluated by the oracle."""
import os
import socket
import subprocess
import time

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

PARTIES = 3
ROWS = 1500


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _wait_listening(proc, deadline):
    while True:
        line = proc.stderr.readline()
        if "Listening on" in line:
            return
        assert proc.poll() is None and time.time() < deadline, line


def test_participants_and_coordinator_find_the_plaintext_minimum(tmp_path):
    import mpc_iris_code_b200  # noqa: F401  (fails loudly when the CUDA library is missing)
    from mpc_iris_code_b200 import build

    if not (os.path.exists(build.PARTICIPANT_PATH) and os.path.exists(build.COORDINATOR_PATH)):
        build.build_participant()

    rng = np.random.default_rng(20261018)
    patterns = rng.integers(0, 2**64, size=(ROWS, 200), dtype=np.uint64)
    masks = rng.integers(0, 2**64, size=(ROWS, 200), dtype=np.uint64) | rng.integers(0, 2**64, size=(ROWS, 200), dtype=np.uint64)
    # query 0: row 1234 with a few pattern bits flipped; the same entry also sits at row 300 -> first minimum wins
    patterns[300], masks[300] = patterns[1234], masks[1234]
    q0p, q0m = patterns[1234].copy(), masks[1234].copy()
    q0p[:3] ^= np.uint64(0x00F0_0000_0F00_0001)
    # query 1: unrelated
    q1p = rng.integers(0, 2**64, size=200, dtype=np.uint64)
    q1m = rng.integers(0, 2**64, size=200, dtype=np.uint64)
    queries = [(q0p, q0m), (q1p, q1m)]

    # `prepare` (src/main.rs:320-372): masks file + one file per additive share of encode(template)
    encoded = np.stack([O.encode(p, m) for p, m in zip(patterns, masks)])
    shares = [rng.integers(0, 2**16, size=encoded.shape, dtype=np.uint16) for _ in range(PARTIES - 1)]
    last = encoded.copy()
    for s in shares:
        last -= s                                                    # wrapping u16, src/encoded_bits.rs:23-38
    shares.append(last)
    assert np.array_equal(O.share_last(encoded[7], np.stack([s[7] for s in shares[:-1]])), last[7])
    masks.tofile(tmp_path / "mpc.masks")
    for i, s in enumerate(shares):
        s.tofile(tmp_path / f"mpc.share-{i}")
    with open(tmp_path / "queries.bin", "wb") as f:
        for p, m in queries:
            f.write(p.tobytes() + m.tobytes())                      # Template {pattern, mask}

    n_gpus = mpc_iris_code_b200.device_count()
    procs, ports = [], []
    try:
        deadline = time.time() + 180
        for i in range(PARTIES):
            port = _free_port()
            ports.append(port)
            procs.append(subprocess.Popen(
                [build.PARTICIPANT_PATH, "--input", str(tmp_path / f"mpc.share-{i}"), "--bind", f"127.0.0.1:{port}",
                 "--batch-size", "700", "--max-requests", str(len(queries)),
                 # participant 1 row-shards its file over three shards (distinct GPUs when the box has them)
                 *(["--devices", ",".join(str(d % n_gpus) for d in range(3))] if i == 1 else [])],
                stderr=subprocess.PIPE, text=True))
        for p in procs:
            _wait_listening(p, deadline)
        out = subprocess.run(
            [build.COORDINATOR_PATH, "--masks", str(tmp_path / "mpc.masks"), "--queries", str(tmp_path / "queries.bin"),
             "--requests", str(len(queries)), "--batch-rows", "400", *[f"127.0.0.1:{p}" for p in ports]],
            capture_output=True, text=True, timeout=180)
        assert out.returncode == 0, out.stderr
        for p in procs:
            assert p.wait(timeout=60) == 0
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()

    lines = out.stdout.strip().splitlines()
    assert len(lines) == len(queries), out.stdout
    for (qp, qm), line in zip(queries, lines):
        idx, rows, dist = line.split()
        idx, rows, dist = int(idx), int(rows), float(dist)
        assert rows == ROWS
        # the coordinator's arithmetic (src/main.rs:597-621) on the true numerators ...
        want_d, want_i = O.combine_min(O.distance_batch(O.encode(qp, qm), encoded, threads=8)[None], O.masks_batch(qm, masks, threads=8))
        assert (idx, dist) == (want_i, want_d)
        # ... which is the plaintext fractional Hamming distance of the winning pair (src/template.rs:43-64)
        assert dist == O.template_distance(qp, qm, patterns[idx], masks[idx])
        assert "Found closest entry at %d out of %d" % (idx, ROWS) in out.stderr
    assert int(lines[0].split()[0]) == 300                            # the first of the two planted copies
