"""Utterance sharding across the GPUs of one box.

Utterances are independent end to end (per-utterance mean, clamped STC edges, fresh decoder
state: srec.cpp:1148-1167), so the path shards with NO data-path collective: rank r of W takes a
contiguous slice of the list balanced by frame count, runs the whole pipeline on its GPU, and
the only exchange is a host gather of the label arrays in original list order (the MLF writer
needs list order, srec.cpp:1246-1291).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(frames_per_utt, world_size: int):
    """Contiguous split of the utterance list into `world_size` slices with near-equal total frames.

    Returns a list of (begin, end) index pairs, one per rank; slices are in list order and cover it.
    """
    f = np.asarray(frames_per_utt, dtype=np.int64)
    n = int(f.size)
    cum = np.concatenate([[0], np.cumsum(f)])
    total = int(cum[-1])
    bounds, begin = [], 0
    for r in range(world_size):
        if r == world_size - 1:
            end = n
        else:
            target = total * (r + 1) / world_size
            end = int(np.searchsorted(cum, target, side="left"))
            # pick the boundary (end or end-1) whose cumulative frame count is closer to the target
            if end > begin and end <= n and abs(cum[end - 1] - target) <= abs(cum[min(end, n)] - target):
                end -= 1
            end = max(begin, min(end, n))
        bounds.append((begin, end))
        begin = end
    return bounds


def gather_in_list_order(local_items, bounds, rank: int, world_size: int, group=None):
    """Host gather of per-utterance results to rank 0, restoring list order.

    `local_items` is this rank's list (one entry per utterance of its slice).  Uses
    torch.distributed.gather_object (host side, tiny payload: ~16 B per label); no GPU collective.
    Returns the full list on rank 0, None elsewhere.  With world_size == 1 no process group is needed.
    """
    b, e = bounds[rank]
    if len(local_items) != e - b:
        raise ValueError(f"rank {rank}: {len(local_items)} results for slice [{b},{e})")
    if world_size == 1:
        return list(local_items)
    import torch.distributed as dist
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(list(local_items), gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = []
    for r in range(world_size):
        rb, re_ = bounds[r]
        if len(gathered[r]) != re_ - rb:
            raise ValueError(f"rank {r} returned {len(gathered[r])} results for slice [{rb},{re_})")
        out.extend(gathered[r])
    return out
