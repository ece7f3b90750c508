"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed.

  * target views are independent units -> `shard_views` partitions them round-robin, every rank renders
    its share of the replicated Gaussian set, no data-path collective ("weak" scaling in bench.py);
  * context views (backbone + cost volume) shard the same way after one all-gather of the stride-4
    matching features;
  * PTF is an order-dependent fold -> "replicas only": `all_gather_views` collects the per-view
    candidates of all ranks (the "cross-view PTF gather" of BASELINE.json) in global view order and every
    rank runs the identical deterministic fusion.

Works with the NCCL backend on GPUs and with gloo on the CPU (tests/test_parallel_gloo.py).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: view v belongs to rank v % world."""
    return [v for v in range(num_views) if v % world == rank]


def owner_of(view: int, world: int) -> int:
    return view % world


def all_gather_views(local: torch.Tensor, num_views: int, group=None) -> torch.Tensor:
    """local: [n_local, ...] holding this rank's views (in increasing global view index).
    Returns [num_views, ...] in global view order on every rank.  Ranks may own different counts."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [len(shard_views(num_views, r, world)) for r in range(world)]
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    out = local.new_empty((num_views,) + tuple(local.shape[1:]))
    for r in range(world):
        idx = shard_views(num_views, r, world)
        if idx:
            out[torch.tensor(idx, device=out.device)] = bufs[r][: len(idx)]
    return out


def max_over_ranks(values: Sequence[float], device, group=None) -> List[float]:
    """Timing reduction used by bench.py (device-timed milliseconds, max over ranks)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t]
