"""The denominators-only scan (BASELINE config 3; `mask_scan_fp4_kernel`, 4-bit operands expanded into tensor memory)
on its own: MasksEngine::batch_process (src/lib.rs:69-79 -> src/arch/generic.rs:4-9) over ragged row ranges, unaligned
device outputs, all-ones / all-zeros / single-bit masks, shards with several tile pairs per CTA, and the three kernel
variants (4-bit, int8 in TMEM, int8 through shared memory) against each other.  Bit-exact against the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x1715C0DE
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def iris():
    import mpc_iris_code_b200 as iris_mod

    assert iris_mod.device_count() >= 1
    return iris_mod


@pytest.fixture(scope="module")
def shard(iris):
    """1 500 masks (5.9 tile pairs, the last one partial) loaded from reference-layout host rows, masks only."""
    n = 1500
    masks = O.gen_mask_rows(SEED, 7_000, n, threads=8)
    db = iris.Database(n, shares=False)
    db.append_masks(masks[:129])
    db.append_masks(masks[129:])
    yield db, masks
    db.close()


@pytest.mark.parametrize("rb,re", [(0, 1), (0, 128), (0, 256), (0, 257), (255, 257), (1, 1500), (256, 512), (300, 1301),
                                   (1499, 1500), (1280, 1500)])
def test_ragged_row_ranges_with_guards(iris, shard, rb, re):
    # the reference calls batch_process on arbitrary chunks of its mmap (src/main.rs:510-516)
    db, masks = shard
    qm = O.gen_mask_rows(71, 0, 1)[0]
    guard = 0xBEEF
    out = np.full((re - rb + 2, 31), guard, np.uint16)
    iris.MasksEngine(qm).batch_process(out[1:-1], db, rb, re)
    assert np.array_equal(out[1:-1], O.masks_batch(qm, masks[rb:re]))
    assert (out[0] == guard).all() and (out[-1] == guard).all()


def test_unaligned_device_output(iris, shard):
    import torch

    db, masks = shard
    qm = O.gen_mask_rows(72, 0, 1)[0]
    me = iris.MasksEngine(qm)
    exp = O.masks_batch(qm, masks, threads=8)
    buf = torch.full((1500 * 31 + 64,), 0x7777, dtype=torch.int16, device="cuda")
    for off, rb, re in ((1, 5, 300), (3, 128, 1500), (7, 0, 1500), (5, 257, 258)):
        buf.fill_(0x7777)
        me.batch_process(buf[off : off + (re - rb) * 31], db, rb, re)
        db.synchronize()
        got = buf.cpu().numpy().view(np.uint16)
        assert np.array_equal(got[off : off + (re - rb) * 31].reshape(-1, 31), exp[rb:re])
        assert (got[:off] == 0x7777).all() and (got[off + (re - rb) * 31 :] == 0x7777).all()


def test_extreme_masks(iris):
    # all ones (12 800 = the largest count, every nibble 0xF incl. the shifted sign position), all zeros, one bit per
    # nibble position, and single bits at the row-wrap columns of the rotation (src/bits.rs:178-205)
    rows = [np.full(O.LIMBS, 2**64 - 1, np.uint64), np.zeros(O.LIMBS, np.uint64)]
    for t in range(4):
        rows.append(np.full(O.LIMBS, int("0x" + "%x" % (1 << t) * 16, 16), np.uint64))
    for bit in (0, 199, 200, 12_799, 3, 7, 64 * 199 + 63):
        r = np.zeros(O.LIMBS, np.uint64)
        r[bit // 64] = np.uint64(1) << np.uint64(bit % 64)
        rows.append(r)
    masks = np.stack(rows)
    with iris.Database(len(masks), shares=False) as db:
        db.append_masks(masks)
        for qm in (masks[0], masks[1], masks[3], masks[5], O.gen_mask_rows(73, 0, 1)[0]):
            out = np.zeros((len(masks), 31), np.uint16)
            iris.MasksEngine(qm).batch_process(out, db)
            assert np.array_equal(out, O.masks_batch(qm, masks))
    full = np.zeros((1, 31), np.uint16)
    with iris.Database(1, shares=False) as db:
        db.append_masks(masks[:1])
        iris.MasksEngine(masks[0]).batch_process(full, db)
    assert (full == 12_800).all()


def test_several_pairs_per_cta_every_row(iris):
    """200 000 masks = 782 tile pairs over 148 CTAs (5-6 pairs each, ring state carried across pairs): every row against
    the CUDA-core kernel, samples and a ragged sub-range against the oracle."""
    import torch

    n = 200_000
    with iris.Database(n, shares=False) as db:
        db.generate(SEED, 0, n)
        qm = O.gen_mask_rows(74, 1, 1)[0]
        me = iris.MasksEngine(qm)
        dn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        for _ in range(3):                                     # back-to-back launches reuse the barriers' initial state
            me.batch_process(dn, db)
        db.synchronize()
        cn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        db.check_denominators_simt(qm, 0, n, cn)
        assert torch.equal(dn, cn)
        host = dn.cpu().numpy().view(np.uint16)
        for i in np.concatenate([[0, 255, 256, 37_887, 37_888, n - 1], np.random.default_rng(75).integers(0, n, 60)]):
            assert np.array_equal(host[i], O.masks_batch(qm, O.gen_mask_rows(SEED, int(i), 1))[0]), i
        rb, re = 12_345, 187_654
        part = torch.zeros((re - rb, 31), dtype=torch.int16, device="cuda")
        me.batch_process(part, db, rb, re)
        db.synchronize()
        assert torch.equal(part, cn[rb:re])


_VARIANT_SCRIPT = r"""
import sys, hashlib
import numpy as np
sys.path.insert(0, %r)
import mpc_iris_code_b200 as iris
n = 20_000
with iris.Database(n, shares=False) as db:
    db.generate(0x1715C0DE, 0, n)
    qm = np.random.default_rng(76).integers(0, 2**64, size=200, dtype=np.uint64)
    out = np.zeros((n - 77, 31), np.uint16)
    iris.MasksEngine(qm).batch_process(out, db, 33, n - 44)
    print(hashlib.sha256(out.tobytes()).hexdigest())
"""


def test_kernel_variants_agree(iris):
    """The diagnostics build (libiris_b200_diag.so, IRIS_B200_DIAG_LIB=1) keeps the superseded kernels behind
    IRIS_MASKSCAN: the int8 TMEM-operand kernel ('i8') and the shared-memory-operand kernel ('smem').  The library
    reads the switch once, so each variant runs in its own process.  All of them, and the product library (which has
    no switch at all), must produce identical bytes."""
    digests = {}
    for mode in ("product", "f", "i8", "smem"):
        env = dict(os.environ)
        for k in ("IRIS_MASKSCAN", "IRIS_M4_VARIANT", "IRIS_B200_DIAG_LIB"):
            env.pop(k, None)
        if mode != "product":
            env.update(IRIS_MASKSCAN=mode, IRIS_B200_DIAG_LIB="1")
        res = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT % ROOT], env=env, capture_output=True, text=True,
                             timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        digests[mode] = res.stdout.strip().splitlines()[-1]
    assert len(set(digests.values())) == 1, digests


def test_product_library_ignores_diagnostic_switches(iris):
    """A stray IRIS_M4_VARIANT / IRIS_MQ_VARIANT (timing-only kernels with WRONG results in the diagnostics build)
    must not change what the product library computes."""
    digests = []
    for extra in ({}, {"IRIS_M4_VARIANT": "6", "IRIS_MQ_VARIANT": "6", "IRIS_MASKSCAN": "i8", "IRIS_BATCHDEN": "i8"}):
        env = dict(os.environ)
        for k in ("IRIS_MASKSCAN", "IRIS_M4_VARIANT", "IRIS_MQ_VARIANT", "IRIS_BATCHDEN", "IRIS_B200_DIAG_LIB"):
            env.pop(k, None)
        env.update(extra)
        res = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT % ROOT], env=env, capture_output=True, text=True,
                             timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        digests.append(res.stdout.strip().splitlines()[-1])
    assert digests[0] == digests[1]


# ------------------------------------------------------------------------------------------------------------------
# batched denominators: four query masks per pass of the 4-bit scan (mask_scan_fp4_multi_kernel), left-overs on the
# single-query kernel


@pytest.mark.parametrize("nq,rb,re", [(4, 0, 1500), (5, 1, 1499), (7, 255, 257), (8, 300, 1301), (3, 0, 1), (12, 1280, 1500)])
def test_four_query_kernel_ragged_ranges(iris, shard, nq, rb, re):
    db, masks = shard
    qms = [O.gen_mask_rows(80 + i, 0, 1)[0] for i in range(nq)]
    guard = 0xC0DE
    buf = np.full(nq * (re - rb) * 31 + 62, guard, np.uint16)      # a guard row before and after the [nq][rows][31] block
    view = buf[31:-31].reshape(nq, re - rb, 31)
    iris.denominators_batch([iris.MasksEngine(q) for q in qms], db, rb, re, view)
    assert np.array_equal(view, np.stack([O.masks_batch(q, masks[rb:re]) for q in qms]))
    assert (buf[:31] == guard).all() and (buf[-31:] == guard).all()


def test_four_query_kernel_unaligned_device_output(iris, shard):
    import torch

    db, masks = shard
    nq, rb, re = 6, 129, 1400
    qms = [O.gen_mask_rows(90 + i, 0, 1)[0] for i in range(nq)]
    exp = np.stack([O.masks_batch(q, masks[rb:re]) for q in qms])
    m = nq * (re - rb) * 31
    for off in (1, 3, 5):
        buf = torch.full((m + 64,), 0x7777, dtype=torch.int16, device="cuda")
        iris.denominators_batch([iris.MasksEngine(q) for q in qms], db, rb, re, buf[off : off + m])
        db.synchronize()
        got = buf.cpu().numpy().view(np.uint16)
        assert np.array_equal(got[off : off + m].reshape(nq, re - rb, 31), exp)
        assert (got[:off] == 0x7777).all() and (got[off + m :] == 0x7777).all()


def test_four_query_kernel_several_pairs_per_cta_every_row(iris):
    import torch

    n, nq = 150_000, 9                                          # two full groups of four + one left-over mask
    with iris.Database(n, shares=False) as db:
        db.generate(SEED, 0, n)
        qms = [O.gen_mask_rows(100 + i, 1, 1)[0] for i in range(nq)]
        out = torch.zeros((nq, n, 31), dtype=torch.int16, device="cuda")
        for _ in range(2):
            iris.denominators_batch([iris.MasksEngine(q) for q in qms], db, 0, n, out)
        db.synchronize()
        cn = torch.zeros((n, 31), dtype=torch.int16, device="cuda")
        for k in range(nq):
            db.check_denominators_simt(qms[k], 0, n, cn)
            db.synchronize()
            assert torch.equal(out[k], cn), k
        host = out[5].cpu().numpy().view(np.uint16)
        for i in (0, 255, 256, 75_775, 75_776, n - 1):
            assert np.array_equal(host[i], O.masks_batch(qms[5], O.gen_mask_rows(SEED, int(i), 1))[0]), i


_BATCH_VARIANT_SCRIPT = r"""
import sys, hashlib
import numpy as np
sys.path.insert(0, %r)
import mpc_iris_code_b200 as iris
n, nq = 20_000, 21
with iris.Database(n, shares=False) as db:
    db.generate(0x1715C0DE, 0, n)
    qms = np.random.default_rng(77).integers(0, 2**64, size=(nq, 200), dtype=np.uint64)
    out = np.zeros((nq, n - 77, 31), np.uint16)
    iris.denominators_batch([iris.MasksEngine(q) for q in qms], db, 33, n - 44, out)
    print(hashlib.sha256(out.tobytes()).hexdigest())
"""


def test_batched_kernel_variants_agree(iris):
    """Diagnostics build: IRIS_BATCHDEN=i8 selects the int8 GEMM kernel for the batched denominators; identical bytes
    either way, and identical to the product library."""
    digests = {}
    for mode in ("product", "", "i8"):
        env = dict(os.environ)
        for k in ("IRIS_BATCHDEN", "IRIS_MQ_VARIANT", "IRIS_B200_DIAG_LIB"):
            env.pop(k, None)
        if mode != "product":
            env["IRIS_B200_DIAG_LIB"] = "1"
        if mode == "i8":
            env["IRIS_BATCHDEN"] = mode
        res = subprocess.run([sys.executable, "-c", _BATCH_VARIANT_SCRIPT % ROOT], env=env, capture_output=True, text=True,
                             timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        digests[mode] = res.stdout.strip().splitlines()[-1]
    assert len(set(digests.values())) == 1, digests
