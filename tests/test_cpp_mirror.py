"""Builds tests/cpp/test_engine.cpp (the reference's unit tests, in C++, against include/iris_engine.hpp) and
runs it on the GPU.  The compile step alone runs on CPU so the header is checked every round."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_engine")


def build_exe():
    import oracle
    from mpc_iris_code_b200 import build

    build.build()
    olib = oracle.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call([
        "g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", EXE,
        os.path.join(ROOT, "tests", "cpp", "test_engine.cpp"),
        "-L" + build.LIB_DIR, "-liris_b200", "-Wl,-rpath," + build.LIB_DIR,
        olib, "-Wl,-rpath," + os.path.dirname(olib), "-fopenmp",
    ])
    return EXE


def test_cpp_mirror_compiles():
    assert os.path.exists(build_exe())


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_unit_tests():
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout + r.stderr
