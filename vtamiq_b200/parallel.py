"""Data-parallel sharding of image pairs across the GPUs of one box.

Pairs are independent end to end (no BatchNorm, no cross-sample op: modules/vtamiq/vtamiq.py:18), so the
path shards with weight replicas and NO hot-path collective; the only exchange is one gather of the
per-pair scores after the forward.  One process per GPU, ``torch.distributed``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_pairs(num_pairs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [start, stop) slice of the pair dimension owned by ``rank``; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(num_pairs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_scores(q_local: torch.Tensor, num_pairs: int, group=None) -> torch.Tensor:
    """All ranks receive the full (num_pairs,) score vector in pair order.  Ragged shards are padded to the
    largest shard for the collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        return q_local
    world = dist.get_world_size(group)
    if world == 1:
        return q_local
    longest = -(-num_pairs // world)
    padded = q_local.new_zeros(longest)
    padded[: q_local.numel()] = q_local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    parts = []
    for r, t in enumerate(out):
        a, b = shard_pairs(num_pairs, r, world)
        parts.append(t[: b - a])
    return torch.cat(parts)
