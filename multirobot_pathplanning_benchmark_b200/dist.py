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


def gather_csr(offsets: torch.Tensor, indices: torch.Tensor, n_rows_total: int, group=None):
    """All-gather of ragged per-row lists (r-disc neighbour search sharded by query rows): every rank
    holds the CSR (offsets [rows_local + 1], indices) of its shard_range block; returns the CSR of all rows.
    Two exchanges: the per-row counts, then the indices padded to the longest shard."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return offsets, indices
    counts = (offsets[1:] - offsets[:-1]).to(torch.int64)
    all_counts = gather_rows(counts, n_rows_total, group)
    sizes = shard_sizes(n_rows_total, world)
    row_starts = [0]
    for s_ in sizes:
        row_starts.append(row_starts[-1] + s_)
    per_rank = [int(all_counts[row_starts[r]:row_starts[r + 1]].sum().item()) for r in range(world)]
    width = max(max(per_rank), 1)
    pad = torch.zeros(width, dtype=indices.dtype, device=indices.device)
    pad[:indices.shape[0]] = indices
    out = torch.empty(world * width, dtype=indices.dtype, device=indices.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    all_idx = torch.cat([out[r * width:r * width + n] for r, n in enumerate(per_rank)])
    all_off = torch.zeros(n_rows_total + 1, dtype=torch.int64, device=offsets.device)
    all_off[1:] = torch.cumsum(all_counts, 0)
    return all_off, all_idx


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pins this process (one rank per GPU) to the CPUs of the NUMA node its GPU hangs on, BEFORE the rank allocates its
    pinned host buffers: first-touch then places them in that node's memory, and the H2D copies of the host-buffer API
    (backend.check_configs_host) stay off the inter-socket link.  Round 1 measured the end-to-end rate at 0.45 of linear
    on 8 GPUs with every rank's staging memory wherever the launcher happened to run.
    -> what was done ({"numa_node", "cpus", "bound"}); never raises (containers may hide sysfs or forbid affinity)."""
    import os
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpus = _parse_cpulist(open(f"{base}/local_cpulist").read())
        info["numa_node"] = node
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        info["cpus"] = len(cpus)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as e:   # noqa: BLE001 -- report, do not fail the run
        info["error"] = repr(e)[:120]
    return info
