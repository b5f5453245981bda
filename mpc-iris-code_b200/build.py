"""Builds libiris_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python -m mpc_iris_code_b200.build        # or: python mpc-iris-code_b200/build.py [--force] [-v] [--diag]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the tree.

Two libraries come out of the same sources:
  lib/libiris_b200.so       the product.  No environment switches, no timing-only kernel variants.
  lib/libiris_b200_diag.so  -DIRIS_DIAGNOSTICS: adds the superseded A/B kernels (iris_maskscan.cu, the int8 GEMM
                            denominators) and the IRIS_* environment switches that tests/diagnostics and the
                            "kernel variants agree" tests use.  Selected with IRIS_B200_DIAG_LIB=1, never by default.
"""
from __future__ import annotations

import fcntl
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libiris_b200.so")
DIAG_LIB_PATH = os.path.join(LIB_DIR, "libiris_b200_diag.so")
SOURCES = ["iris_kernels.cu", "iris_abi.cu", "iris_batch.cu", "iris_reduce.cu", "iris_maskscan4.cu", "iris_dotbatch.cu",
           "iris_cluster.cu"]
DIAG_SOURCES = SOURCES + ["iris_maskscan.cu"]
HEADERS = ["iris_layout.h", "iris_ptx.cuh", "iris_kernels.cuh", "iris_epilogue.cuh",
           "../../include/iris_b200.h"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
    "--threads",
    "0",
    "-shared",
]
LINK_FLAGS = ["-ldl", "-lpthread"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources(names=SOURCES):
    return [os.path.join(CSRC, s) for s in names if os.path.exists(os.path.join(CSRC, s))]


def is_stale(lib_path: str = LIB_PATH, names=SOURCES) -> bool:
    if not os.path.exists(lib_path):
        return True
    t = os.path.getmtime(lib_path)
    deps = _sources(names) + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _compile(lib_path: str, names, extra, verbose: bool) -> None:
    os.makedirs(LIB_DIR, exist_ok=True)
    # one builder at a time (torchrun ranks and pytest-xdist workers may all find a stale library)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not is_stale(lib_path, names) and not extra.get("force"):
            return
        tmp = lib_path + f".tmp{os.getpid()}"
        cmd = [find_nvcc(), *NVCC_FLAGS, *extra.get("flags", []), "-o", tmp, *_sources(names), *LINK_FLAGS]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        os.replace(tmp, lib_path)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    _compile(LIB_PATH, SOURCES, {"force": force}, verbose)
    build_participant()
    return LIB_PATH


def build_diagnostics(force: bool = False, verbose: bool = False) -> str:
    """The diagnostics library (see the module docstring); not loaded unless IRIS_B200_DIAG_LIB=1."""
    if not force and not is_stale(DIAG_LIB_PATH, DIAG_SOURCES):
        return DIAG_LIB_PATH
    _compile(DIAG_LIB_PATH, DIAG_SOURCES, {"force": force, "flags": ["-DIRIS_DIAGNOSTICS"]}, verbose)
    return DIAG_LIB_PATH


BIN_DIR = os.path.join(HERE, "bin")
PARTICIPANT_PATH = os.path.join(BIN_DIR, "iris_participant")
COORDINATOR_PATH = os.path.join(BIN_DIR, "iris_coordinator")


def _build_front_end(source: str, target: str) -> str:
    os.makedirs(BIN_DIR, exist_ok=True)
    tmp = target + f".tmp{os.getpid()}"
    subprocess.check_call([
        "g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", tmp, os.path.join(CSRC, source),
        "-L" + LIB_DIR, "-liris_b200", "-Wl,-rpath,$ORIGIN/../lib",
    ])
    os.replace(tmp, target)
    return target


def build_participant() -> str:
    """The wire-protocol front-ends over the C ABI, plain g++: reference `participant` (src/main.rs:384-452) and
    the query loop of `coordinator` (src/main.rs:453-640).  Returns the participant's path."""
    _build_front_end("coordinator_main.cpp", COORDINATOR_PATH)
    return _build_front_end("participant_main.cpp", PARTICIPANT_PATH)


if __name__ == "__main__":
    if "--diag" in sys.argv:
        print(build_diagnostics(force="--force" in sys.argv, verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
