"""Data-parallel plumbing of the hot path: one process per GPU, image pairs sharded across ranks,
weights replicated, and exactly ONE collective — the sum of the integer confusion matrix
(SURVEY.md §8e).  The reference itself only ever scatters the batch (``nn.DataParallel``,
models/networks.py:132-133) and keeps the metric on one process.

Pure ``torch.distributed``: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ``[begin, end)`` slice of ``n_total`` image pairs owned by ``rank``: the first
    ``n_total % world`` ranks take one extra pair (uneven last shards are legal; empty ones too)."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError(f"bad shard request n={n_total} rank={rank} world={world}")
    base, extra = divmod(n_total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_confusion(cm: torch.Tensor, group=None, pixels: Optional[int] = None):
    """In-place SUM of an int64 confusion matrix over the ranks of ``group``.  Integer addition is
    associative, so the result is bit-exact whatever the reduction order.  No-op without an
    initialised process group (single-process use).

    ``pixels``: the number of pixels this rank accumulated (SegmentationMetric's out-of-range check);
    it rides in the SAME all-reduce as one extra int64 word -- the path keeps exactly one collective --
    and the all-rank total is returned instead of ``cm`` (an int without a process group, else a 0-dim tensor on
    ``cm``'s device so that the call itself never synchronises)."""
    import torch.distributed as dist
    if cm.dtype != torch.int64:
        raise TypeError("the confusion matrix is accumulated as int64 (exact); got %s" % cm.dtype)
    active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if pixels is None:
        if active:
            dist.all_reduce(cm, op=dist.ReduceOp.SUM, group=group)
        return cm
    if not active:
        return int(pixels)
    flat = cm.reshape(-1)
    buf = torch.cat([flat, torch.tensor([int(pixels)], dtype=torch.int64, device=cm.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    flat.copy_(buf[:-1])
    return buf[-1]          # 0-dim device tensor: reading it is the caller's (only) synchronisation point


def rank_world(group: Optional[object] = None) -> Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1
