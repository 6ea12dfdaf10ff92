"""Sample-sharded evaluation across the GPUs of one box.

Images are independent units: rank r of W takes a contiguous slice of the batch, weights are
replicated, and there is no collective inside the forward.  The only exchanges are the eval
loop's accuracy counts (one all-reduce of two int64, reference validate/test.py:120-127) and,
when a caller wants every logit on every rank, one all-gather.  Works with any
``torch.distributed`` backend (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``total`` samples for ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_counts(correct: torch.Tensor, total: torch.Tensor) -> torch.Tensor:
    """Global [correct, total] as an int64 tensor on the inputs' device."""
    counts = torch.stack([correct.to(torch.int64).reshape(()), total.to(torch.int64).reshape(())])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def max_over_ranks(value: float, device=None) -> float:
    """The slowest rank's value (pass times: the whole job is as fast as its slowest shard)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_logits(local_logits: torch.Tensor, total: int) -> torch.Tensor:
    """All ranks' logits in global sample order, [total, C] (ragged shards are padded for the collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_logits
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = local_logits.new_zeros(pad, local_logits.shape[1])
    buf[:local_logits.shape[0]] = local_logits
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
