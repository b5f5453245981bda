"""Builds libiris_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python -m mpc_iris_code_b200.build        # or: python mpc-iris-code_b200/build.py

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libiris_b200.so")
SOURCES = ["iris_kernels.cu", "iris_abi.cu", "iris_batch.cu", "iris_reduce.cu", "iris_maskscan.cu", "iris_maskscan4.cu"]
HEADERS = ["iris_layout.h", "iris_ptx.cuh", "iris_kernels.cuh", "iris_epilogue.cuh", "../../include/iris_b200.h"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
    "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", tmp, *_sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)
    build_participant()
    return LIB_PATH


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
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
