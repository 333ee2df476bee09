"""Sharding of the hot path across the GPUs of one node (one process per GPU).

The path shards by independent units -- configurations, edges, query rows: every
rank takes a contiguous block of rows, the scene blob / corpus is replicated, and the only
exchange is the final gather of flag bytes or neighbour indices (`torch.distributed`; NCCL over
NVLink on GPUs, gloo in the CPU tests).  No collective runs inside the data path.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of rank `rank`; the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of per-rank row blocks produced with shard_range (ragged last blocks allowed)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = shard_sizes(n_total, world)
    width = max(sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.dtype == torch.bool:  # collectives on bytes
        dist.all_gather_into_tensor(out.view(torch.uint8), pad.view(torch.uint8), group=group)
    else:
        dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * width:r * width + s] for r, s in enumerate(sizes)])


def sharded_map(fn: Callable[[int, int], torch.Tensor], n_total: int, group=None) -> torch.Tensor:
    """Run fn(start, end) on this rank's block and gather every rank's rows (same result on all ranks)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    s, e = shard_range(n_total, rank, world)
    return gather_rows(fn(s, e), n_total, group)
