"""Host-side sharding of the hot path across GPUs.

The path has no cross-image coupling, so N GPUs are N independent replicas working
on disjoint images: no data-path collective.  `torch.distributed` is used for the start/stop
barrier and for reducing the device time to its maximum over ranks (bench contract)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `total` images owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]


def max_over_ranks(value: float, device: torch.device) -> float:
    """max of a per-rank scalar (device time in seconds)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
