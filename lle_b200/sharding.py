"""Multi-GPU layout of a batch of worlds: contiguous env ranges per rank, no data-path collective.

Worlds never interact (the reference's `World` owns all of its state, src/core/world.rs:21-44), so N GPUs simply
step N independent slices.  The Philox action stream is keyed by the *global* env id (`env_id_base`), so results do
not depend on the number of GPUs.  The only collective is an optional end-of-run reduction of a few counters.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world_size: int) -> tuple[int, int]:
    """[begin, end) of the global env ids owned by `rank` (contiguous, sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reduce_stats(stats: torch.Tensor, op=None) -> torch.Tensor:
    """Sum (or `op`) a small tensor of counters over all ranks (NCCL on GPUs, gloo on CPU); identity when not distributed."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=op if op is not None else dist.ReduceOp.SUM)
    return stats
