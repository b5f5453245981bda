"""ctypes binding of include/iris_b200.h, shaped like the reference's Rust API.

Reference names kept: DistanceEngine / MasksEngine with new (constructor) and batch_process,
distances(), denominators(), dot_u16(), dot_bool()  (src/lib.rs:28-94, src/arch/mod.rs:5).
Buffers are numpy arrays (host) or torch CUDA tensors (device); layouts are the reference's
(EncodedBits = u16[12800], Bits = u64[200], result rows = u16[31]).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

COLS = 200          # src/lib.rs:10
ROWS = 64           # src/lib.rs:11
BITS = ROWS * COLS  # src/lib.rs:12
LIMBS = BITS // 64  # src/bits.rs:10
ROTATIONS = 31      # src/lib.rs:34-35

IRIS_DB_SHARES = 1
IRIS_DB_MASKS = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
PROGRESS_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64)   # iris_progress_fn


class IrisError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"iris_b200 error {code}: {message}")
        self.code = code


def _diagnostics_requested() -> bool:
    return os.environ.get("IRIS_B200_DIAG_LIB", "") not in ("", "0")


def library_path() -> str:
    """The product library; IRIS_B200_DIAG_LIB=1 selects the diagnostics build (tests/diagnostics and the
    kernel-variant tests only -- it carries A/B kernels and timing-only variants the product does not)."""
    return os.path.join(_HERE, "lib", "libiris_b200_diag.so" if _diagnostics_requested() else "libiris_b200.so")


def lib():
    """Loads libiris_b200.so, (re)building it with nvcc when it is absent or older than its sources.  No fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    from . import build as _build

    path = _build.build_diagnostics() if _diagnostics_requested() else _build.build()
    if os.environ.get("IRIS_B200_LIB"):          # an explicitly built A/B library (tests/diagnostics only)
        path = os.environ["IRIS_B200_LIB"]
    L = ctypes.CDLL(path)
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    pp = ctypes.POINTER(ctypes.c_void_p)
    sigs = {
        "iris_device_count": [ctypes.POINTER(i32)],
        "iris_dot_u16": [i32, vp, vp, vp],
        "iris_dot_bool": [i32, vp, vp, vp],
        "iris_db_create": [i32, u64, u32, pp],
        "iris_db_destroy": [vp],
        "iris_db_clear": [vp],
        "iris_db_len": [vp, ctypes.POINTER(u64), ctypes.POINTER(u64)],
        "iris_db_append_shares": [vp, vp, u64],
        "iris_db_append_masks": [vp, vp, u64],
        "iris_db_generate": [vp, u64, u64, u64],
        "iris_db_generate_shares": [vp, u64, u32, u32, u64, u64],
        "iris_db_write_shares": [vp, u64, vp, u64],
        "iris_db_write_masks": [vp, u64, vp, u64],
        "iris_db_get_stream": [vp, pp],
        "iris_db_device": [vp, ctypes.POINTER(i32)],
        "iris_db_check": [vp],
        "iris_dot_u16_batch": [i32, vp, u32, vp, u64, vp],
        "iris_dot_bool_batch": [i32, vp, u32, vp, u64, vp],
        "iris_dot_u16_batch_resident": [vp, u32, vp, u64, u64, vp],
        "iris_dot_bool_batch_resident": [vp, u32, vp, u64, u64, vp],
        "iris_match_min_resident_async": [vp, vp, vp, u64, u64, u64, vp],
        "iris_search_batch_resident_async": [vp, vp, u32, vp, u64, u64, u64, vp],
        "iris_cluster_create": [ctypes.POINTER(i32), u32, u64, u32, pp],
        "iris_cluster_destroy": [vp],
        "iris_cluster_partition": [u64, u32, u32, ctypes.POINTER(u64), ctypes.POINTER(u64)],
        "iris_cluster_shard": [vp, u32, pp, ctypes.POINTER(i32), ctypes.POINTER(u64), ctypes.POINTER(u64)],
        "iris_cluster_len": [vp, ctypes.POINTER(u32), ctypes.POINTER(u64), ctypes.POINTER(u64)],
        "iris_cluster_generate": [vp, u64, u32, u32, u64, u64],
        "iris_cluster_load_files": [vp, ctypes.c_char_p, ctypes.c_char_p],
        "iris_cluster_load_rows": [vp, vp, vp, u64],
        "iris_cluster_set_index_base": [vp, u64],
        "iris_cluster_match": [vp, vp, vp, vp, vp],
        "iris_cluster_match_template": [vp, vp, vp, vp, vp],
        "iris_cluster_search": [vp, vp, u32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)],
        "iris_cluster_match_template_streamed": [vp, vp, vp, vp, vp, PROGRESS_FN, vp],
        "iris_match_resident_streamed": [vp, vp, vp, u64, u64, vp, vp, PROGRESS_FN, vp],
        "iris_db_set_overlap": [vp, i32],
        "iris_comm_unique_id": [vp],
        "iris_cluster_join": [vp, vp, i32, i32],
        "iris_cluster_match_allgather": [vp, vp, vp, vp, vp],
        "iris_db_load_shares_file": [vp, ctypes.c_char_p, u64, u64],
        "iris_db_load_masks_file": [vp, ctypes.c_char_p, u64, u64],
        "iris_db_read_shares": [vp, u64, u64, vp],
        "iris_db_read_masks": [vp, u64, u64, vp],
        "iris_db_set_stream": [vp, vp],
        "iris_db_synchronize": [vp],
        "iris_distance_engine_new": [i32, vp, pp],
        "iris_distance_engine_free": [vp],
        "iris_distance_engine_new_from_template": [i32, vp, vp, pp],
        "iris_engines_new_from_templates": [i32, vp, u32, vp, vp],
        "iris_encode": [i32, vp, vp, vp],
        "iris_distance_engine_batch_process": [vp, vp, u64, vp, u64],
        "iris_distance_engine_batch_process_resident": [vp, vp, u64, vp, u64, u64],
        "iris_masks_engine_new": [i32, vp, pp],
        "iris_masks_engine_free": [vp],
        "iris_masks_engine_batch_process": [vp, vp, u64, vp, u64],
        "iris_masks_engine_batch_process_resident": [vp, vp, u64, vp, u64, u64],
        "iris_match_resident": [vp, vp, vp, u64, u64, vp, vp],
        "iris_distances_batch_resident": [vp, u32, vp, u64, u64, vp],
        "iris_denominators_batch_resident": [vp, u32, vp, u64, u64, vp],
        "iris_combine_min": [i32, vp, u32, vp, u64, u64, vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)],
        "iris_match_min_resident": [vp, vp, vp, u64, u64, u64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)],
        "iris_combine_min_batch": [i32, vp, vp, u32, u64, u64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)],
        "iris_distances": [i32, vp, vp, vp],
        "iris_denominators": [i32, vp, vp, vp],
        "iris_check_distances_simt": [vp, vp, u64, u64, vp],
        "iris_check_denominators_simt": [vp, vp, u64, u64, vp],
        "iris_debug_raw_accumulators": [vp, vp, vp, u64, u64, vp],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    L.iris_last_error.restype = ctypes.c_char_p
    L.iris_last_error.argtypes = []
    L.iris_launch_count.restype = u64
    L.iris_launch_count.argtypes = []
    _LIB = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise IrisError(rc, lib().iris_last_error().decode(errors="replace"))


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x, dtype, min_elems: int, what: str) -> int:
    """Raw address of a numpy array (host) or torch tensor (host or CUDA) after layout checks."""
    if x is None:
        return 0
    if _is_torch(x):
        import torch

        want = {np.uint16: (torch.uint16, torch.int16), np.uint64: (torch.uint64, torch.int64), np.int32: (torch.int32,)}[dtype]
        if x.dtype not in want:
            raise TypeError(f"{what}: expected torch dtype in {want}, got {x.dtype}")
        if not x.is_contiguous():
            raise ValueError(f"{what}: tensor must be contiguous")
        if x.numel() < min_elems:
            raise ValueError(f"{what}: needs {min_elems} elements, has {x.numel()}")
        return x.data_ptr()
    if not isinstance(x, np.ndarray):
        raise TypeError(f"{what}: expected numpy array or torch tensor, got {type(x)}")
    if x.dtype != dtype:
        raise TypeError(f"{what}: expected dtype {np.dtype(dtype)}, got {x.dtype}")
    if not x.flags["C_CONTIGUOUS"]:
        raise ValueError(f"{what}: array must be C-contiguous")
    if x.size < min_elems:
        raise ValueError(f"{what}: needs {min_elems} elements, has {x.size}")
    return x.ctypes.data


def _numel(x) -> int:
    return x.numel() if _is_torch(x) else x.size


def device_count() -> int:
    n = ctypes.c_int(0)
    _check(lib().iris_device_count(ctypes.byref(n)))
    return n.value


def launch_count() -> int:
    return int(lib().iris_launch_count())


# ------------------------------------------------------------------ arch entry points
def dot_u16(a, b, device: int = 0) -> int:
    """arch::dot_u16 (src/arch/generic.rs:11-16)."""
    out = np.zeros(1, np.uint16)
    _check(lib().iris_dot_u16(device, _ptr(a, np.uint16, BITS, "a"), _ptr(b, np.uint16, BITS, "b"), out.ctypes.data))
    return int(out[0])


def dot_bool(a, b, device: int = 0) -> int:
    """arch::dot_bool (src/arch/generic.rs:4-9)."""
    out = np.zeros(1, np.uint16)
    _check(lib().iris_dot_bool(device, _ptr(a, np.uint64, LIMBS, "a"), _ptr(b, np.uint64, LIMBS, "b"), out.ctypes.data))
    return int(out[0])


def dot_u16_batch(a, b, out=None, device: int = 0):
    """The reference's criterion grid for dot_u16 (src/arch/mod.rs:46-72) in one call: a = [n_a][12800] u16 independent
    vectors, b = [n_b][12800] u16 (or a resident Database); returns out[i][j] = dot_u16(a[j], b[i]), [n_b][n_a] u16."""
    n_a = _numel(a) // BITS
    if isinstance(b, Database):
        n_b = b.len_shares
        out = np.empty((n_b, n_a), np.uint16) if out is None else out
        _check(lib().iris_dot_u16_batch_resident(_ptr(a, np.uint16, n_a * BITS, "a"), n_a, b._h, 0, n_b, _ptr(out, np.uint16, n_a * n_b, "out")))
        return out
    n_b = _numel(b) // BITS
    out = np.empty((n_b, n_a), np.uint16) if out is None else out
    _check(lib().iris_dot_u16_batch(device, _ptr(a, np.uint16, n_a * BITS, "a"), n_a, _ptr(b, np.uint16, n_b * BITS, "b"), n_b,
                                    _ptr(out, np.uint16, n_a * n_b, "out")))
    return out


def dot_bool_batch(a, b, out=None, device: int = 0):
    """The criterion grid for dot_bool (src/arch/mod.rs:22-44): a = [n_a][200] u64, b = [n_b][200] u64 or a Database."""
    n_a = _numel(a) // LIMBS
    if isinstance(b, Database):
        n_b = b.len_masks
        out = np.empty((n_b, n_a), np.uint16) if out is None else out
        _check(lib().iris_dot_bool_batch_resident(_ptr(a, np.uint64, n_a * LIMBS, "a"), n_a, b._h, 0, n_b, _ptr(out, np.uint16, n_a * n_b, "out")))
        return out
    n_b = _numel(b) // LIMBS
    out = np.empty((n_b, n_a), np.uint16) if out is None else out
    _check(lib().iris_dot_bool_batch(device, _ptr(a, np.uint64, n_a * LIMBS, "a"), n_a, _ptr(b, np.uint64, n_b * LIMBS, "b"), n_b,
                                     _ptr(out, np.uint16, n_a * n_b, "out")))
    return out


# ------------------------------------------------------------------ database shard
class Database:
    """HBM-resident shard: the reference's mmapped &[EncodedBits] / &[Bits] (src/main.rs:389-391, 458-461)."""

    def __init__(self, capacity_rows: int, device: int = 0, shares: bool = True, masks: bool = True, _borrowed=None):
        self._h = ctypes.c_void_p()
        self.device = device
        self._owned = _borrowed is None
        if _borrowed is not None:          # a shard of a Cluster: the cluster owns it
            self._h = _borrowed
            return
        flags = (IRIS_DB_SHARES if shares else 0) | (IRIS_DB_MASKS if masks else 0)
        _check(lib().iris_db_create(device, capacity_rows, flags, ctypes.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            if self._owned:
                lib().iris_db_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _lens(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        _check(lib().iris_db_len(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    @property
    def len_shares(self) -> int:
        return self._lens()[0]

    @property
    def len_masks(self) -> int:
        return self._lens()[1]

    def clear(self) -> None:
        _check(lib().iris_db_clear(self._h))

    def append_shares(self, rows) -> None:
        n = _numel(rows) // BITS
        if n * BITS != _numel(rows):
            raise ValueError("shares must be [n][12800] u16")
        _check(lib().iris_db_append_shares(self._h, _ptr(rows, np.uint16, n * BITS, "rows"), n))

    def append_masks(self, rows) -> None:
        n = _numel(rows) // LIMBS
        if n * LIMBS != _numel(rows):
            raise ValueError("masks must be [n][200] u64")
        _check(lib().iris_db_append_masks(self._h, _ptr(rows, np.uint64, n * LIMBS, "rows"), n))

    def load_shares_file(self, path: str, first_row: int = 0, n_rows: int = 0) -> None:
        """Append rows of a reference `mpc.share-i` file (raw EncodedBits rows, src/main.rs:355-371)."""
        _check(lib().iris_db_load_shares_file(self._h, os.fsencode(path), first_row, n_rows))

    def load_masks_file(self, path: str, first_row: int = 0, n_rows: int = 0) -> None:
        """Append rows of a reference `mpc.masks` file (raw Bits rows, src/main.rs:341-367)."""
        _check(lib().iris_db_load_masks_file(self._h, os.fsencode(path), first_row, n_rows))

    def generate(self, seed: int, first_row_id: int, n: int) -> None:
        _check(lib().iris_db_generate(self._h, seed, first_row_id, n))

    def generate_shares(self, seed: int, party: int, n_parties: int, first_row_id: int, n: int) -> None:
        """Party `party`'s additive shares (EncodedBits::share, src/encoded_bits.rs:23-38) of the encodings of the
        synthetic Templates with row ids first_row_id..; n_parties = 1 stores the plaintext encodings."""
        _check(lib().iris_db_generate_shares(self._h, seed, party, n_parties, first_row_id, n))

    def write_shares(self, row: int, rows) -> None:
        n = _numel(rows) // BITS
        _check(lib().iris_db_write_shares(self._h, row, _ptr(rows, np.uint16, n * BITS, "rows"), n))

    def write_masks(self, row: int, rows) -> None:
        n = _numel(rows) // LIMBS
        _check(lib().iris_db_write_masks(self._h, row, _ptr(rows, np.uint64, n * LIMBS, "rows"), n))

    def read_shares(self, row_begin: int, n: int) -> np.ndarray:
        out = np.empty((n, BITS), np.uint16)
        _check(lib().iris_db_read_shares(self._h, row_begin, n, out.ctypes.data))
        return out

    def read_masks(self, row_begin: int, n: int) -> np.ndarray:
        out = np.empty((n, LIMBS), np.uint64)
        _check(lib().iris_db_read_masks(self._h, row_begin, n, out.ctypes.data))
        return out

    def set_overlap(self, allow: bool) -> None:
        """Whether consecutive device-output scans of this shard may overlap each other's tails (default: yes)."""
        _check(lib().iris_db_set_overlap(self._h, 1 if allow else 0))

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        _check(lib().iris_db_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def synchronize(self) -> None:
        _check(lib().iris_db_synchronize(self._h))

    # verification helpers (CUDA-core kernels, not the product path)
    def check_distances_simt(self, query, row_begin: int, row_end: int, out=None):
        if out is None:
            out = np.empty((row_end - row_begin, ROTATIONS), np.uint16)
        _check(lib().iris_check_distances_simt(self._h, _ptr(query, np.uint16, BITS, "query"), row_begin, row_end,
                                               _ptr(out, np.uint16, (row_end - row_begin) * ROTATIONS, "out")))
        return out

    def check_denominators_simt(self, qmask, row_begin: int, row_end: int, out=None):
        if out is None:
            out = np.empty((row_end - row_begin, ROTATIONS), np.uint16)
        _check(lib().iris_check_denominators_simt(self._h, _ptr(qmask, np.uint64, LIMBS, "qmask"), row_begin, row_end,
                                                  _ptr(out, np.uint16, (row_end - row_begin) * ROTATIONS, "out")))
        return out


def _out_len(out) -> int:
    n = _numel(out)
    if n % ROTATIONS:
        raise ValueError("out must be [n][31] u16")
    return n // ROTATIONS


# ------------------------------------------------------------------ engines
def encode(pattern, mask, device: int = 0) -> np.ndarray:
    """encode(&Template) -> EncodedBits (src/lib.rs:16-26), on the device."""
    out = np.empty(BITS, np.uint16)
    _check(lib().iris_encode(device, _ptr(pattern, np.uint64, LIMBS, "pattern"), _ptr(mask, np.uint64, LIMBS, "mask"), out.ctypes.data))
    return out


class DistanceEngine:
    """DistanceEngine (src/lib.rs:28-52): new(query) prepares rotations -15..=15; batch_process fills out[i][j]."""

    def __init__(self, query, device: int = 0, _handle=None):
        self._h = ctypes.c_void_p()
        self.device = device
        if _handle is not None:
            self._h = _handle
            return
        _check(lib().iris_distance_engine_new(device, _ptr(query, np.uint16, BITS, "query"), ctypes.byref(self._h)))

    @classmethod
    def from_template(cls, pattern, mask, device: int = 0) -> "DistanceEngine":
        """DistanceEngine::new(&encode(&template)) (src/main.rs:427) with encode done on the device."""
        h = ctypes.c_void_p()
        _check(lib().iris_distance_engine_new_from_template(
            device, _ptr(pattern, np.uint64, LIMBS, "pattern"), _ptr(mask, np.uint64, LIMBS, "mask"), ctypes.byref(h)))
        return cls(None, device, _handle=h)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().iris_distance_engine_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def batch_process(self, out, db, row_begin: Optional[int] = None, row_end: Optional[int] = None) -> None:
        """batch_process(&self, out, db).  `db` is either a host array of EncodedBits ([n][12800] u16,
        the reference's literal signature) or a resident Database (optionally a row range of it)."""
        if isinstance(db, Database):
            rb = 0 if row_begin is None else row_begin
            re = db.len_shares if row_end is None else row_end
            _check(lib().iris_distance_engine_batch_process_resident(
                self._h, _ptr(out, np.uint16, 0, "out"), _out_len(out), db._h, rb, re))
        else:
            n = _numel(db) // BITS
            if n * BITS != _numel(db):
                raise ValueError("db must be [n][12800] u16")
            _check(lib().iris_distance_engine_batch_process(
                self._h, _ptr(out, np.uint16, 0, "out"), _out_len(out), _ptr(db, np.uint16, 0, "db"), n))


class MasksEngine:
    """MasksEngine (src/lib.rs:55-79)."""

    def __init__(self, query_mask, device: int = 0, _handle=None):
        self._h = ctypes.c_void_p()
        self.device = device
        if _handle is not None:
            self._h = _handle
            return
        _check(lib().iris_masks_engine_new(device, _ptr(query_mask, np.uint64, LIMBS, "query_mask"), ctypes.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().iris_masks_engine_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def batch_process(self, out, db, row_begin: Optional[int] = None, row_end: Optional[int] = None) -> None:
        if isinstance(db, Database):
            rb = 0 if row_begin is None else row_begin
            re = db.len_masks if row_end is None else row_end
            _check(lib().iris_masks_engine_batch_process_resident(
                self._h, _ptr(out, np.uint16, 0, "out"), _out_len(out), db._h, rb, re))
        else:
            n = _numel(db) // LIMBS
            if n * LIMBS != _numel(db):
                raise ValueError("db must be [n][200] u64")
            _check(lib().iris_masks_engine_batch_process(
                self._h, _ptr(out, np.uint16, 0, "out"), _out_len(out), _ptr(db, np.uint64, 0, "db"), n))


def engines_from_templates(templates, device: int = 0, masks: bool = True):
    """Q wire Templates ([Q][400] u64 = {pattern[200], mask[200]} each, src/template.rs:26-29) -> (list of
    DistanceEngine, list of MasksEngine) prepared in one batch (one copy, three launches, one synchronisation)."""
    q = _numel(templates) // (2 * LIMBS)
    if q * 2 * LIMBS != _numel(templates):
        raise ValueError("templates must be [Q][400] u64")
    de = (ctypes.c_void_p * q)()
    me = (ctypes.c_void_p * q)() if masks else None
    _check(lib().iris_engines_new_from_templates(device, _ptr(templates, np.uint64, q * 2 * LIMBS, "templates"), q, de, me))
    des = [DistanceEngine(None, device, _handle=ctypes.c_void_p(h)) for h in de]
    mes = [MasksEngine(None, device, _handle=ctypes.c_void_p(h)) for h in me] if masks else []
    return des, mes


def match(distance_engine: Optional[DistanceEngine], masks_engine: Optional[MasksEngine], db: Database,
          row_begin: int, row_end: int, distances_out=None, denominators_out=None) -> None:
    """Fused scan: both engines over rows [row_begin,row_end) in one pass over HBM."""
    n = (row_end - row_begin) * ROTATIONS
    _check(lib().iris_match_resident(
        distance_engine._h if distance_engine else None,
        masks_engine._h if masks_engine else None,
        db._h, row_begin, row_end,
        _ptr(distances_out, np.uint16, n, "distances_out") if distance_engine else None,
        _ptr(denominators_out, np.uint16, n, "denominators_out") if masks_engine else None))


def match_streamed(distance_engine, masks_engine, db: Database, row_begin: int, row_end: int, distances_out, denominators_out,
                   progress) -> None:
    """match() with HOST outputs; progress(b, e) is called as the rows [b, e) land in host memory while later rows
    are still being scanned (iris_match_resident_streamed)."""
    n = (row_end - row_begin) * ROTATIONS
    cb = PROGRESS_FN(lambda _u, b, e: progress(b, e))
    _check(lib().iris_match_resident_streamed(
        distance_engine._h if distance_engine else None, masks_engine._h if masks_engine else None, db._h, row_begin, row_end,
        _ptr(distances_out, np.uint16, n, "distances_out") if distance_engine else None,
        _ptr(denominators_out, np.uint16, n, "denominators_out") if masks_engine else None, cb, None))


def distances_batch(engines, db: Database, row_begin: int, row_end: int, out) -> None:
    """All `engines` (DistanceEngine list) against rows [row_begin,row_end) as one tensor-core GEMM;
    out = [len(engines)][rows][31] u16, numpy (host) or torch CUDA tensor."""
    n = len(engines) * (row_end - row_begin) * ROTATIONS
    arr = (ctypes.c_void_p * len(engines))(*[e._h.value for e in engines])
    _check(lib().iris_distances_batch_resident(arr, len(engines), db._h, row_begin, row_end, _ptr(out, np.uint16, n, "out")))


def denominators_batch(engines, db: Database, row_begin: int, row_end: int, out) -> None:
    """All `engines` (MasksEngine list) against rows [row_begin,row_end), four query masks per pass of the 4-bit
    tensor-core scan; out = [len(engines)][rows][31] u16."""
    n = len(engines) * (row_end - row_begin) * ROTATIONS
    arr = (ctypes.c_void_p * len(engines))(*[e._h.value for e in engines])
    _check(lib().iris_denominators_batch_resident(arr, len(engines), db._h, row_begin, row_end, _ptr(out, np.uint16, n, "out")))


def combine_min(distance_shares, denominators, index_base: int = 0, device: int = 0, want_distances: bool = False):
    """Coordinator reduction (src/main.rs:597-621 + decode_distance src/lib.rs:97-107) on the device.
    `distance_shares`: list of [n][31] u16 arrays (one per party), `denominators`: [n][31] u16.
    Returns (min_distance, min_index) or (min_distance, min_index, distances[n]) ; min_index is -1 when no
    distance is below +inf (the reference leaves usize::MAX)."""
    n = _numel(denominators) // ROTATIONS
    arr = (ctypes.c_void_p * len(distance_shares))(*[_ptr(s, np.uint16, n * ROTATIONS, "share") for s in distance_shares])
    md, mi = ctypes.c_double(), ctypes.c_uint64()
    dist = np.empty(n, np.float64) if want_distances else None
    _check(lib().iris_combine_min(device, arr, len(distance_shares), _ptr(denominators, np.uint16, n * ROTATIONS, "denominators"),
                                  n, index_base, dist.ctypes.data if want_distances else None,
                                  ctypes.byref(md), ctypes.byref(mi)))
    idx = -1 if mi.value == 2**64 - 1 else mi.value
    return (md.value, idx, dist) if want_distances else (md.value, idx)


def combine_min_batch(distances, denominators, num_queries: int, index_base: int = 0, device: int = 0):
    """Per-query (min distance, argmin) over [Q][n][31] device arrays produced by the batched kernels."""
    n = _numel(denominators) // (ROTATIONS * num_queries)
    md = (ctypes.c_double * num_queries)()
    mi = (ctypes.c_uint64 * num_queries)()
    _check(lib().iris_combine_min_batch(device, _ptr(distances, np.uint16, num_queries * n * ROTATIONS, "distances"),
                                        _ptr(denominators, np.uint16, num_queries * n * ROTATIONS, "denominators"),
                                        num_queries, n, index_base, md, mi))
    return np.array(md[:], np.float64), np.array([-1 if v == 2**64 - 1 else v for v in mi[:]], np.int64)


def match_min(distance_engine: DistanceEngine, masks_engine: MasksEngine, db: Database, row_begin: int, row_end: int,
              index_base: int = 0):
    """Fused scan + decode + min/argmin on the device for a shard holding whole encodings; returns
    (min_distance, global_row) with global_row = index_base + row, or -1 if nothing is below +inf."""
    md, mi = ctypes.c_double(), ctypes.c_uint64()
    _check(lib().iris_match_min_resident(distance_engine._h, masks_engine._h, db._h, row_begin, row_end, index_base,
                                         ctypes.byref(md), ctypes.byref(mi)))
    return md.value, (-1 if mi.value == 2**64 - 1 else mi.value)


def match_min_async(distance_engine: DistanceEngine, masks_engine: MasksEngine, db: Database, row_begin: int, row_end: int,
                    result, index_base: int = 0) -> None:
    """The asynchronous form: the {f64 min, u64 row} pair (16 bytes) is written to `result` (a torch CUDA tensor of
    two int64 / uint64 elements, or any device / mapped address) in stream order; no host synchronisation."""
    ptr = result.data_ptr() if _is_torch(result) else int(result)
    _check(lib().iris_match_min_resident_async(distance_engine._h, masks_engine._h, db._h, row_begin, row_end, index_base,
                                               ctypes.c_void_p(ptr)))


def raw_accumulators(distance_engine, masks_engine, db: Database, row_begin: int, row_end: int) -> np.ndarray:
    tiles = (row_end - row_begin + 127) // 128
    raw = np.empty((tiles * 128, 128), np.int32)
    _check(lib().iris_debug_raw_accumulators(
        distance_engine._h if distance_engine else None, masks_engine._h if masks_engine else None,
        db._h, row_begin, row_end, raw.ctypes.data))
    return raw


def distances(query, entry, device: int = 0) -> np.ndarray:
    """distances(query, entry) -> [u16;31]  (src/lib.rs:82-87)."""
    out = np.zeros(ROTATIONS, np.uint16)
    _check(lib().iris_distances(device, _ptr(query, np.uint16, BITS, "query"), _ptr(entry, np.uint16, BITS, "entry"), out.ctypes.data))
    return out


def denominators(query, entry, device: int = 0) -> np.ndarray:
    """denominators(query, entry) -> [u16;31]  (src/lib.rs:89-94)."""
    out = np.zeros(ROTATIONS, np.uint16)
    _check(lib().iris_denominators(device, _ptr(query, np.uint64, LIMBS, "query"), _ptr(entry, np.uint64, LIMBS, "entry"), out.ctypes.data))
    return out


# ------------------------------------------------------------------ cluster: one database over several GPUs
def cluster_partition(n_total: int, n_shards: int, shard: int):
    b, e = ctypes.c_uint64(), ctypes.c_uint64()
    _check(lib().iris_cluster_partition(n_total, n_shards, shard, ctypes.byref(b), ctypes.byref(e)))
    return b.value, e.value


def _prefer_bundled_nccl() -> None:
    """A process can hold one libnccl.so.2.  torch links against the copy bundled with it (nvidia-nccl wheel), which is
    newer than the system's; if the library loaded the system copy first, a later `import torch` would fail with an
    undefined symbol.  So a Python host points the library (IRIS_NCCL_LIB, read at its first NCCL call) at the
    bundled copy, without importing torch."""
    if os.environ.get("IRIS_NCCL_LIB"):
        return
    import importlib.util

    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for d in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
        path = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(path):
            os.environ["IRIS_NCCL_LIB"] = path
            return


def comm_unique_id() -> bytes:
    """The id rank 0 draws for a multi-process cluster (128 bytes; hand it to the other processes by any channel)."""
    _prefer_bundled_nccl()
    buf = ctypes.create_string_buffer(128)
    _check(lib().iris_comm_unique_id(buf))
    return buf.raw


class Cluster:
    """One database row-sharded over `devices` (the compiled library runs one host thread + stream per GPU).
    search() gathers 16 bytes per query and shard over NVLink; match() lets every GPU store its block of the
    [n][31] results into one array (on any GPU of the cluster, or in host memory)."""

    def __init__(self, devices, capacity_rows: int, shares: bool = True, masks: bool = True):
        self._h = ctypes.c_void_p()
        self.devices = list(devices)
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        flags = (IRIS_DB_SHARES if shares else 0) | (IRIS_DB_MASKS if masks else 0)
        _check(lib().iris_cluster_create(arr, len(self.devices), capacity_rows, flags, ctypes.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().iris_cluster_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self) -> int:
        ns, a, b = ctypes.c_uint32(), ctypes.c_uint64(), ctypes.c_uint64()
        _check(lib().iris_cluster_len(self._h, ctypes.byref(ns), ctypes.byref(a), ctypes.byref(b)))
        return max(a.value, b.value)

    def shard(self, i: int):
        """(Database view, device, row_begin, row_end) of shard i."""
        h, dev, b, e = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_uint64(), ctypes.c_uint64()
        _check(lib().iris_cluster_shard(self._h, i, ctypes.byref(h), ctypes.byref(dev), ctypes.byref(b), ctypes.byref(e)))
        return Database(0, dev.value, _borrowed=h), dev.value, b.value, e.value

    def generate(self, seed: int, n: int, first_row_id: int = 0, party: int = 0, n_parties: int = 1) -> None:
        _check(lib().iris_cluster_generate(self._h, seed, party, n_parties, first_row_id, n))

    def load_files(self, shares_path: Optional[str], masks_path: Optional[str]) -> None:
        _check(lib().iris_cluster_load_files(self._h, os.fsencode(shares_path) if shares_path else None,
                                             os.fsencode(masks_path) if masks_path else None))

    def load_rows(self, shares=None, masks=None) -> None:
        n = _numel(shares) // BITS if shares is not None else _numel(masks) // LIMBS
        _check(lib().iris_cluster_load_rows(self._h, _ptr(shares, np.uint16, n * BITS, "shares"),
                                            _ptr(masks, np.uint64, n * LIMBS, "masks"), n))

    def set_index_base(self, base: int) -> None:
        _check(lib().iris_cluster_set_index_base(self._h, base))

    def join(self, unique_id: bytes, rank: int, world_size: int) -> None:
        _prefer_bundled_nccl()
        _check(lib().iris_cluster_join(self._h, unique_id, rank, world_size))

    def match(self, query=None, query_mask=None, distances_out=None, denominators_out=None) -> None:
        n = len(self) * ROTATIONS
        _check(lib().iris_cluster_match(self._h, _ptr(query, np.uint16, BITS, "query"), _ptr(query_mask, np.uint64, LIMBS, "query_mask"),
                                        _ptr(distances_out, np.uint16, n, "distances_out"),
                                        _ptr(denominators_out, np.uint16, n, "denominators_out")))

    def match_allgather(self, query=None, query_mask=None, distances_out=None, denominators_out=None) -> None:
        """Joined (multi-process) clusters: every process scans its rows and all processes end up with the result
        vectors of ALL rows (device arrays [total rows][31]); collective."""
        _check(lib().iris_cluster_match_allgather(self._h, _ptr(query, np.uint16, BITS, "query"),
                                                  _ptr(query_mask, np.uint64, LIMBS, "query_mask"),
                                                  _ptr(distances_out, np.uint16, 0, "distances_out"),
                                                  _ptr(denominators_out, np.uint16, 0, "denominators_out")))

    def match_template(self, pattern, mask, distances_out, denominators_out=None) -> None:
        n = len(self) * ROTATIONS
        _check(lib().iris_cluster_match_template(self._h, _ptr(pattern, np.uint64, LIMBS, "pattern"), _ptr(mask, np.uint64, LIMBS, "mask"),
                                                 _ptr(distances_out, np.uint16, n, "distances_out"),
                                                 _ptr(denominators_out, np.uint16, n, "denominators_out")))

    def match_template_streamed(self, pattern, mask, distances_out, progress, denominators_out=None) -> None:
        """match_template with HOST outputs; progress(b, e) is called (from the library's per-GPU threads) as cluster
        rows [b, e) land in host memory."""
        n = len(self) * ROTATIONS
        cb = PROGRESS_FN(lambda _u, b, e: progress(b, e))
        _check(lib().iris_cluster_match_template_streamed(
            self._h, _ptr(pattern, np.uint64, LIMBS, "pattern"), _ptr(mask, np.uint64, LIMBS, "mask"),
            _ptr(distances_out, np.uint16, n, "distances_out"), _ptr(denominators_out, np.uint16, n, "denominators_out"), cb, None))

    def search(self, templates):
        """templates: [Q][400] u64 wire Templates -> (min distances [Q] f64, rows [Q] i64, -1 = none)."""
        q = _numel(templates) // (2 * LIMBS)
        md = (ctypes.c_double * q)()
        mi = (ctypes.c_uint64 * q)()
        _check(lib().iris_cluster_search(self._h, _ptr(templates, np.uint64, q * 2 * LIMBS, "templates"), q, md, mi))
        return np.array(md[:], np.float64), np.array([-1 if v == 2**64 - 1 else v for v in mi[:]], np.int64)
