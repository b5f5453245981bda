"""Row sharding of one party's database across the GPUs of a box (one process per GPU).

Rows are independent (src/lib.rs:44-51: one output row per database row, no cross-row state), so
the scan needs no data-path collective: rank g owns the contiguous block returned by shard_rows and
produces its slice of the [N][31] results.  The only exchange is after the scan, on small per-query
vectors: the coordinator-style reduction (running min / argmin over rows, src/main.rs:611-621)
becomes a per-shard (min, argmin) pair that is all-gathered and reduced.
"""
from __future__ import annotations

from typing import Tuple


def shard_rows(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of rank `rank`; sizes differ by at most one row."""
    if world_size <= 0 or not (0 <= rank < world_size) or n_total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_total, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def owner_of(row: int, n_total: int, world_size: int) -> int:
    base, extra = divmod(n_total, world_size)
    edge = extra * (base + 1)
    if row < edge:
        return row // (base + 1)
    return extra + (row - edge) // base if base else world_size - 1


def gather_best_batch(local_mins, global_indices, group=None):
    """Batched form of gather_best: per-query arrays (float64 mins, int64 GLOBAL row indices, -1 = none) of this
    shard -> per-query global (min, row) over all shards, lowest row on ties.  One all-gather of 16*Q bytes."""
    import numpy as np
    import torch
    import torch.distributed as dist

    mins = np.asarray(local_mins, np.float64)
    idx = np.asarray(global_indices, np.int64)
    if not (dist.is_available() and dist.is_initialized()):
        return mins, idx
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    q = mins.shape[0]
    mine = torch.from_numpy(np.concatenate([mins, idx.astype(np.float64)])).to(dev)
    allv = torch.empty(world * 2 * q, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.cpu().numpy().reshape(world, 2, q)
    m, g = allv[:, 0, :], allv[:, 1, :].astype(np.int64)
    g_cmp = np.where(g >= 0, g, np.iinfo(np.int64).max)
    m_cmp = np.where(g >= 0, m, np.inf)
    order = np.lexsort((g_cmp, m_cmp), axis=0)[0]          # per query: smallest min, then smallest row
    cols = np.arange(q)
    return m_cmp[order, cols], np.where(np.isfinite(m_cmp[order, cols]), g[order, cols], -1)


def gather_best(local_min: float, local_index: int, row_offset: int, group=None):
    """All-gathers each shard's (min distance, local argmin) and returns the global (min, global row).

    Ties resolve to the lowest global row, which is what the reference's sequential scan with
    `if distance < min_distance` (src/main.rs:617) yields.  `group` is a torch.distributed group
    (NCCL on the GPU box, gloo in CPU tests); without an initialised group this is the identity."""
    import torch
    import torch.distributed as dist

    glob = row_offset + local_index if local_index >= 0 else -1
    if not (dist.is_available() and dist.is_initialized()):
        return local_min, glob
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([local_min, float(glob)], dtype=torch.float64, device=dev)
    allv = torch.empty(world * 2, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.cpu().view(world, 2)
    best, best_row = float("inf"), -1
    for r in range(world):
        m, g = float(allv[r, 0]), int(allv[r, 1])
        if g >= 0 and (m < best or (m == best and g < best_row)):
            best, best_row = m, g
    return best, best_row
