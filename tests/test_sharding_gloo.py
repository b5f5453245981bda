"""world_size-2 gloo test (CPU) of the multi-GPU host logic: row sharding + post-scan gather.

The per-shard compute is done by the CPU oracle here (this is a host-logic test; the GPU scan itself
is covered by tests/test_parity_gpu.py); the sharded result must equal the unsharded one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mpc_iris_code_b200.sharding import owner_of, shard_rows  # noqa: E402

N, SEED = 37, 0x1715C0DE


def test_shard_rows_partition():
    for n in (0, 1, 7, 37, 1_000_000, 16_000_001):
        for w in (1, 2, 3, 4, 8):
            blocks = [shard_rows(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    for row in (0, 18, 19, 36):
        r = owner_of(row, N, 2)
        b, e = shard_rows(N, 2, r)
        assert b <= row < e
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _problem():
    import oracle as O

    qp, qm = O.gen_mask_rows(61, 0, 1)[0], O.gen_mask_rows(61, 1, 1)[0]
    ep, em = O.gen_mask_rows(62, 0, N), O.gen_mask_rows(63, 0, N)
    ep[29], em[29] = qp, qm                      # planted exact match on rank 1's shard
    ep[3], em[3] = qp, qm                        # and an equal-distance one on rank 0's: lowest row must win
    enc = np.stack([O.encode(ep[i], em[i]) for i in range(N)])
    return O.encode(qp, qm), qm, enc, em


def _worker(rank, world, port, out):
    import torch.distributed as dist

    import oracle as O
    from mpc_iris_code_b200.sharding import gather_best, shard_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, qm, enc, em = _problem()
    b, e = shard_rows(N, world, rank)
    d = O.distance_batch(q, enc[b:e])
    den = O.masks_batch(qm, em[b:e])
    md, mi = O.combine_min(d[None], den)
    best = gather_best(md, mi, b)
    # batched form: three "queries" per shard -- the real one, one with no finite distance on rank 1, one all-inf
    from mpc_iris_code_b200.sharding import gather_best_batch

    mins = np.array([md, 0.25 if rank == 0 else np.inf, np.inf])
    idxs = np.array([b + mi if mi >= 0 else -1, b + 1 if rank == 0 else -1, -1], np.int64)
    bm, bi = gather_best_batch(mins, idxs)
    if rank == 0:
        out.put((best, (b, e), bm.tolist(), bi.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_unsharded():
    import oracle as O

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    best, _, bm, bi = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    q, qm, enc, em = _problem()
    md, mi = O.combine_min(O.distance_batch(q, enc)[None], O.masks_batch(qm, em))
    assert best == (md, mi) == (0.0, 3)
    assert bm == [0.0, 0.25, float("inf")] and bi == [3, 1, -1]
