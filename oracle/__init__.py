"""CPU oracle for the matching hot path of recmo/mpc-iris-code (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (mpc-iris-code_b200/) never does.

Two independent restatements live here:
  * iris_oracle.c  -- plain C, statement-by-statement after the reference (the timed CPU arm);
  * the np_* functions below -- numpy, written from the mathematical definition
    (np.roll for the rotation, integer matmul for the dots) as a second opinion.
tests/test_oracle.py checks them against each other and against every reference test that
can run without the (absent) data/ directory.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import numpy as np

COLS = 200          # src/lib.rs:10
ROWS = 64           # src/lib.rs:11
BITS = ROWS * COLS  # src/lib.rs:12
LIMBS = BITS // 64  # src/bits.rs:10
ROTATIONS = 31      # src/lib.rs:34-35  (-15..=15)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _cpu_tag() -> str:
    """-march=native output is CPU specific; key the build on the CPU's feature flags."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def lib_path() -> str:
    return os.path.join(_HERE, "_build", f"libiris_oracle_{_cpu_tag()}.so")


def build(force: bool = False) -> str:
    out = lib_path()
    src = os.path.join(_HERE, "iris_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        tmp = out + f".tmp{os.getpid()}"
        subprocess.check_call(
            ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", tmp, src, "-lm"]
        )
        os.replace(tmp, out)
    return out


def _u16p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16))


def _u64p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def load():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.oracle_dot_u16.restype = ctypes.c_uint16
        lib.oracle_dot_bool.restype = ctypes.c_uint16
        lib.oracle_bits_count_ones.restype = ctypes.c_uint16
        lib.oracle_encoded_sum.restype = ctypes.c_uint16
        lib.oracle_bits_get.restype = ctypes.c_int
        lib.oracle_decode_distance.restype = ctypes.c_double
        lib.oracle_fraction_hamming.restype = ctypes.c_double
        lib.oracle_template_distance.restype = ctypes.c_double
        _LIB = lib
    return _LIB


def _c16(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    if n is not None:
        assert a.size == n, (a.shape, n)
    return a


def _c64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if n is not None:
        assert a.size == n, (a.shape, n)
    return a


# ----------------------------------------------------------------------------- C oracle wrappers
def dot_u16(a, b) -> int:
    a, b = _c16(a, BITS), _c16(b, BITS)
    return int(load().oracle_dot_u16(_u16p(a), _u16p(b)))


def dot_bool(a, b) -> int:
    a, b = _c64(a, LIMBS), _c64(b, LIMBS)
    return int(load().oracle_dot_bool(_u64p(a), _u64p(b)))


def bits_rotated(bits, amount: int) -> np.ndarray:
    out = _c64(bits, LIMBS).copy()
    load().oracle_bits_rotate(_u64p(out), ctypes.c_int(amount))
    return out


def encoded_rotated(v, amount: int) -> np.ndarray:
    out = _c16(v, BITS).copy()
    load().oracle_encoded_rotate(_u16p(out), ctypes.c_int(amount))
    return out


def encoded_from_bits(bits) -> np.ndarray:
    bits = _c64(bits, LIMBS)
    out = np.empty(BITS, np.uint16)
    load().oracle_encoded_from_bits(_u64p(bits), _u16p(out))
    return out


def encode(pattern, mask) -> np.ndarray:
    pattern, mask = _c64(pattern, LIMBS), _c64(mask, LIMBS)
    out = np.empty(BITS, np.uint16)
    load().oracle_encode(_u64p(pattern), _u64p(mask), _u16p(out))
    return out


def share_last(self_v, rest) -> np.ndarray:
    self_v = _c16(self_v, BITS)
    rest = _c16(rest).reshape(-1, BITS)
    out = np.empty(BITS, np.uint16)
    load().oracle_share_last(_u16p(self_v), _u16p(rest), ctypes.c_size_t(rest.shape[0]), _u16p(out))
    return out


def distance_rotations(query) -> np.ndarray:
    query = _c16(query, BITS)
    rot = np.empty((ROTATIONS, BITS), np.uint16)
    load().oracle_distance_rotations(_u16p(query), _u16p(rot))
    return rot


def mask_rotations(qmask) -> np.ndarray:
    qmask = _c64(qmask, LIMBS)
    rot = np.empty((ROTATIONS, LIMBS), np.uint64)
    load().oracle_mask_rotations(_u64p(qmask), _u64p(rot))
    return rot


def distance_batch(query, db, threads: int = 1) -> np.ndarray:
    """DistanceEngine::new(query).batch_process(out, db)  (src/lib.rs:33-52)."""
    db = _c16(db).reshape(-1, BITS)
    rot = distance_rotations(query)
    out = np.empty((db.shape[0], ROTATIONS), np.uint16)
    load().oracle_distance_batch(_u16p(rot), _u16p(db), ctypes.c_size_t(db.shape[0]), _u16p(out), ctypes.c_int(threads))
    return out


def distance_batch_prepared(rot, db, out, threads: int = 1) -> None:
    load().oracle_distance_batch(_u16p(rot), _u16p(db), ctypes.c_size_t(db.shape[0]), _u16p(out), ctypes.c_int(threads))


def masks_batch(qmask, db, threads: int = 1) -> np.ndarray:
    """MasksEngine::new(qmask).batch_process(out, db)  (src/lib.rs:60-79)."""
    db = _c64(db).reshape(-1, LIMBS)
    rot = mask_rotations(qmask)
    out = np.empty((db.shape[0], ROTATIONS), np.uint16)
    load().oracle_masks_batch(_u64p(rot), _u64p(db), ctypes.c_size_t(db.shape[0]), _u16p(out), ctypes.c_int(threads))
    return out


def masks_batch_prepared(rot, db, out, threads: int = 1) -> None:
    load().oracle_masks_batch(_u64p(rot), _u64p(db), ctypes.c_size_t(db.shape[0]), _u16p(out), ctypes.c_int(threads))


def distances(query, entry) -> np.ndarray:
    query, entry = _c16(query, BITS), _c16(entry, BITS)
    out = np.empty(ROTATIONS, np.uint16)
    load().oracle_distances(_u16p(query), _u16p(entry), _u16p(out))
    return out


def denominators(qmask, emask) -> np.ndarray:
    qmask, emask = _c64(qmask, LIMBS), _c64(emask, LIMBS)
    out = np.empty(ROTATIONS, np.uint16)
    load().oracle_denominators(_u64p(qmask), _u64p(emask), _u16p(out))
    return out


def decode_distance(dist31, den31) -> float:
    d, n = _c16(dist31, ROTATIONS), _c16(den31, ROTATIONS)
    return float(load().oracle_decode_distance(_u16p(d), _u16p(n)))


def template_distance(ap, am, bp, bm) -> float:
    ap, am, bp, bm = (_c64(x, LIMBS) for x in (ap, am, bp, bm))
    return float(load().oracle_template_distance(_u64p(ap), _u64p(am), _u64p(bp), _u64p(bm)))


def combine_min(dist_shares, denoms):
    dist_shares = _c16(dist_shares)
    denoms = _c16(denoms).reshape(-1, ROTATIONS)
    n = denoms.shape[0]
    parties = dist_shares.size // (n * ROTATIONS)
    md = ctypes.c_double()
    mi = ctypes.c_int64()
    load().oracle_combine_min(
        _u16p(dist_shares), ctypes.c_size_t(parties), _u16p(denoms), ctypes.c_size_t(n), ctypes.byref(md), ctypes.byref(mi)
    )
    return md.value, mi.value


def gen_share_rows(seed: int, row0: int, n: int, threads: int = 1) -> np.ndarray:
    out = np.empty((n, BITS), np.uint16)
    load().oracle_gen_share_rows(ctypes.c_uint64(seed), ctypes.c_uint64(row0), ctypes.c_size_t(n), _u16p(out), ctypes.c_int(threads))
    return out


def gen_mask_rows(seed: int, row0: int, n: int, threads: int = 1) -> np.ndarray:
    out = np.empty((n, LIMBS), np.uint64)
    load().oracle_gen_mask_rows(ctypes.c_uint64(seed), ctypes.c_uint64(row0), ctypes.c_size_t(n), _u64p(out), ctypes.c_int(threads))
    return out


def gen_pattern_rows(seed: int, row0: int, n: int, threads: int = 1) -> np.ndarray:
    """pattern_R of the synthetic Templates (mask_R = gen_mask_rows)."""
    out = np.empty((n, LIMBS), np.uint64)
    load().oracle_gen_pattern_rows(ctypes.c_uint64(seed), ctypes.c_uint64(row0), ctypes.c_size_t(n), _u64p(out), ctypes.c_int(threads))
    return out


def gen_party_share_rows(seed: int, party: int, n_parties: int, row0: int, n: int, threads: int = 1) -> np.ndarray:
    """Party `party`'s additive share (EncodedBits::share, src/encoded_bits.rs:23-38) of encode(Template R)."""
    out = np.empty((n, BITS), np.uint16)
    load().oracle_gen_party_share_rows(ctypes.c_uint64(seed), ctypes.c_uint32(party), ctypes.c_uint32(n_parties),
                                       ctypes.c_uint64(row0), ctypes.c_size_t(n), _u16p(out), ctypes.c_int(threads))
    return out


# ----------------------------------------------------------------------------- numpy second opinion
def np_bits_to_bool(bits) -> np.ndarray:
    """[..., 200] u64 -> [..., 12800] {0,1}; bit k = byte k/8, bit k%8 (src/bits.rs:44-57, test_index)."""
    b = np.ascontiguousarray(bits, dtype="<u8")
    return np.unpackbits(b.view(np.uint8).reshape(*b.shape[:-1], LIMBS * 8), axis=-1, bitorder="little")


def np_bool_to_bits(x) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint8)
    return np.packbits(x, axis=-1, bitorder="little").view("<u8").reshape(*x.shape[:-1], LIMBS)


def np_encoded_rotated(v, amount: int) -> np.ndarray:
    """out[row][col] = in[row][(col - amount) mod 200]  (pinned by test_rotated_number)."""
    v = np.asarray(v, dtype=np.uint16).reshape(ROWS, COLS)
    return np.roll(v, amount, axis=1).reshape(BITS)


def np_bits_rotated(bits, amount: int) -> np.ndarray:
    x = np_bits_to_bool(np.asarray(bits, dtype=np.uint64).reshape(LIMBS)).reshape(ROWS, COLS)
    return np_bool_to_bits(np.roll(x, amount, axis=1).reshape(BITS))


def np_encode(pattern, mask) -> np.ndarray:
    m = np_bits_to_bool(np.asarray(mask, np.uint64).reshape(LIMBS)).astype(np.uint16)
    p = np_bits_to_bool(np.asarray(pattern, np.uint64).reshape(LIMBS)).astype(np.uint16) & m
    return (m - p - p).astype(np.uint16)


def np_distance_batch(query, db) -> np.ndarray:
    db = np.asarray(db, dtype=np.uint16).reshape(-1, BITS).astype(np.int64)
    rot = np.stack([np_encoded_rotated(query, j - 15) for j in range(ROTATIONS)]).astype(np.int64)
    return ((db @ rot.T) & 0xFFFF).astype(np.uint16)


def np_masks_batch(qmask, db) -> np.ndarray:
    db = np_bits_to_bool(np.asarray(db, dtype=np.uint64).reshape(-1, LIMBS)).astype(np.int64)
    rot = np.stack([np_bits_to_bool(np_bits_rotated(qmask, j - 15)) for j in range(ROTATIONS)]).astype(np.int64)
    return (db @ rot.T).astype(np.uint16)


def np_decode_distance(dist31, den31) -> float:
    d = np.asarray(dist31, np.uint16)
    n = np.asarray(den31, np.uint16)
    num = ((n - d).astype(np.uint16) // 2).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = num / n.astype(np.float64)
    f = f[~np.isnan(f)]
    return float(min(np.inf, f.min())) if f.size else float("inf")
