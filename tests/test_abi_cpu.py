"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/iris_b200.h declares, and fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "iris_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(iris_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for required in ("iris_dot_u16", "iris_dot_bool", "iris_distance_engine_new", "iris_distance_engine_batch_process",
                     "iris_masks_engine_new", "iris_masks_engine_batch_process", "iris_distances", "iris_denominators"):
        assert required in names


def test_library_exports_every_declared_symbol():
    import mpc_iris_code_b200 as iris
    from mpc_iris_code_b200 import build

    build.build()
    L = ctypes.CDLL(iris.library_path())
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing
    iris.lib()  # binds argtypes for every function the wrapper uses


def test_sass_contains_tcgen05_and_bulk_copy():
    """The product kernel must be the tensor-core one: UTCIMMA (tcgen05.mma kind::i8), LDTM (tcgen05.ld),
    UBLKCP (cp.async.bulk) in the SASS of the shipped library."""
    import shutil
    import subprocess

    import mpc_iris_code_b200 as iris

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", iris.library_path()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCIMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    # the denominators-only scan: block-scaled 4-bit UMMA (tcgen05.mma kind::mxf4) fed from tensor memory (tcgen05.st)
    for mnemonic in ("UTCOMMA", "STTM"):
        assert mnemonic in sass, mnemonic


def test_diagnostic_cuda_sources_compile(tmp_path):
    """tests/diagnostics/*.cu are the micro-benchmarks DESIGN.md quotes; keep them building for sm_100a."""
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    diag = os.path.join(ROOT, "tests", "diagnostics")
    for src in ("umma_bench.cu", "sttm_bench.cu"):
        out = tmp_path / (src + ".cubin")
        subprocess.run([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-o", str(out),
                        os.path.join(diag, src)], check=True, capture_output=True)
        assert out.stat().st_size > 0


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mpc_iris_code_b200 as iris

    q = np.zeros(iris.BITS, np.uint16)
    m = np.zeros(iris.LIMBS, np.uint64)
    for call in (lambda: iris.Database(128), lambda: iris.DistanceEngine(q), lambda: iris.MasksEngine(m),
                 lambda: iris.dot_u16(q, q), lambda: iris.dot_bool(m, m), lambda: iris.distances(q, q),
                 lambda: iris.denominators(m, m)):
        with pytest.raises(iris.IrisError) as ei:
            call()
        assert ei.value.code == -2


def test_front_ends_build_and_fail_loudly_without_a_gpu(tmp_path):
    """iris_participant / iris_coordinator (reference src/main.rs:384-640) link against the C ABI only; without a
    device they must exit non-zero with the library's error, not serve anything."""
    import subprocess

    import torch

    from mpc_iris_code_b200 import build

    build.build_participant()
    assert subprocess.run([build.PARTICIPANT_PATH]).returncode == 2          # usage
    assert subprocess.run([build.COORDINATOR_PATH]).returncode == 2
    if torch.cuda.is_available():
        return
    np.zeros((3, 200), np.uint64).tofile(tmp_path / "mpc.masks")
    r = subprocess.run([build.COORDINATOR_PATH, "--masks", str(tmp_path / "mpc.masks"), "127.0.0.1:9"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "iris_db_create" in r.stderr
    np.zeros((3, 12800), np.uint16).tofile(tmp_path / "mpc.share-0")
    r = subprocess.run([build.PARTICIPANT_PATH, "--input", str(tmp_path / "mpc.share-0")], capture_output=True, text=True)
    assert r.returncode == 1 and "iris_db_create" in r.stderr


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mpc-iris-code_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "iris_oracle" not in text, f
    for f in ("include/iris_b200.h",):
        assert "oracle/" not in open(os.path.join(ROOT, f)).read().replace("oracle/iris_oracle.c restates the generator", "")


def test_wrapper_validates_buffers():
    import mpc_iris_code_b200 as iris
    from mpc_iris_code_b200 import api

    with pytest.raises(TypeError):
        api._ptr(np.zeros(iris.BITS, np.int32), np.uint16, iris.BITS, "q")
    with pytest.raises(ValueError):
        api._ptr(np.zeros(10, np.uint16), np.uint16, iris.BITS, "q")
    with pytest.raises(ValueError):
        api._ptr(np.zeros((4, iris.BITS), np.uint16)[:, ::2], np.uint16, 1, "q")


# ---------------------------------------------------------------------------------- Rust binding stays in sync
_C2RUST = {
    "int": "c_int", "uint64_t": "u64", "uint32_t": "u32", "const char*": "*const c_char", "void*": "*mut c_void",
    "void**": "*mut *mut c_void", "const void*": "*const c_void", "const uint16_t*": "*const u16", "uint16_t*": "*mut u16",
    "const uint64_t*": "*const u64", "uint64_t*": "*mut u64", "uint32_t*": "*mut u32", "double*": "*mut f64",
    "const int*": "*const c_int", "int*": "*mut c_int", "const uint16_t*const*": "*const *const u16",
    "iris_db*": "*mut IrisDb", "iris_db**": "*mut *mut IrisDb", "const iris_db*": "*const IrisDb",
    "iris_distance_engine*": "*mut IrisDistanceEngine", "iris_distance_engine**": "*mut *mut IrisDistanceEngine",
    "iris_masks_engine*": "*mut IrisMasksEngine", "iris_masks_engine**": "*mut *mut IrisMasksEngine",
    "iris_cluster*": "*mut IrisCluster", "iris_cluster**": "*mut *mut IrisCluster", "const iris_cluster*": "*const IrisCluster",
}


def _c_prototypes():
    text = open(os.path.join(ROOT, "include", "iris_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|uint64_t|const char \*)\s*\**(iris_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        params = []
        for a in [x.strip() for x in args.split(",") if x.strip() and x.strip() != "void"]:
            a = re.sub(r"\[[^\]]*\]", "*", a)                      # array parameters decay to pointers
            m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)?\s*(\**)$", a)
            m = re.match(r"^(.*?[\s\*])([A-Za-z_][A-Za-z0-9_]*)(\**)$", a)
            ctype = (m.group(1) + m.group(3)) if m else a
            params.append(re.sub(r"\s+", " ", ctype).replace(" *", "*").replace("* ", "*").strip())
        protos[name] = params
    return protos


def _rust_prototypes():
    text = open(os.path.join(ROOT, "integration", "rust", "src", "arch", "cuda.rs")).read()
    block = text[text.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    block = re.sub(r"//[^\n]*", "", block)
    protos = {}
    for name, args in re.findall(r"fn\s+(iris_[a-z0-9_]+)\s*\(([^)]*)\)", block, flags=re.S):
        params = [re.sub(r"\s+", " ", a.split(":", 1)[1]).strip() for a in args.split(",") if ":" in a]
        protos[name] = params
    return protos


def test_rust_extern_block_matches_the_header():
    """integration/rust/src/arch/cuda.rs cannot be compiled here (no Rust toolchain): keep at least its extern "C"
    declarations identical, name by name and parameter by parameter, to include/iris_b200.h."""
    c, r = _c_prototypes(), _rust_prototypes()
    assert len(r) >= 30
    for name, rparams in r.items():
        assert name in c, f"{name} is bound in cuda.rs but not declared in iris_b200.h"
        want = [_C2RUST[p] for p in c[name]]
        assert rparams == want, f"{name}: cuda.rs has {rparams}, the header means {want}"


def test_product_library_has_no_diagnostic_switches():
    """The timing-only kernel variants and every IRIS_* environment switch live in the diagnostics build only: a stray
    variable in a deployment must not be able to change what the product computes."""
    import mpc_iris_code_b200 as iris
    from mpc_iris_code_b200 import build

    build.build()
    blob = open(build.LIB_PATH, "rb").read()
    for name in (b"IRIS_M4_VARIANT", b"IRIS_MQ_VARIANT", b"IRIS_MASKSCAN", b"IRIS_BATCHDEN", b"IRIS_BATCH_CLUSTERS"):
        assert name not in blob, name
    assert not iris.library_path().endswith("_diag.so") or os.environ.get("IRIS_B200_DIAG_LIB")
    diag = open(build.build_diagnostics(), "rb").read()
    assert b"IRIS_M4_VARIANT" in diag


def test_rust_build_script_compiles_the_same_sources():
    from mpc_iris_code_b200 import build

    text = open(os.path.join(ROOT, "integration", "rust", "build.rs")).read()
    listed = re.findall(r'"(iris_[a-z0-9_]+\.cu)"', text)
    assert sorted(listed) == sorted(build.SOURCES)


def test_first_nccl_call_does_not_break_a_later_torch_import():
    """A process holds one libnccl.so.2.  The Python binding points the library at the copy bundled with torch
    (IRIS_NCCL_LIB) before its first NCCL call, so importing torch AFTERWARDS still resolves the newer symbols
    torch needs (seen on a two-GPU box: `undefined symbol: ncclDevCommCreate` with the system's older copy)."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r)\n"
            "import mpc_iris_code_b200 as iris\n"
            "assert len(iris.comm_unique_id()) == 128\n"
            "import torch\n"
            "print('ok', torch.__version__)\n") % ROOT
    env = {k: v for k, v in os.environ.items() if k != "IRIS_NCCL_LIB"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_a_wrong_nccl_path_falls_back_to_the_system_copy():
    """IRIS_NCCL_LIB only NAMES a copy; a path that cannot be loaded must not disable multi-process clusters."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r)\n"
            "import mpc_iris_code_b200 as iris\n"
            "assert len(iris.comm_unique_id()) == 128\n"
            "print('ok')\n") % ROOT
    env = dict(os.environ, IRIS_NCCL_LIB="/nonexistent/libnccl.so.2")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
