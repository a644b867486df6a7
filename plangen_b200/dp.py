"""Prompt-level data parallelism: batch k of the evaluation set goes to rank
k mod world (what accelerate.prepare(test_dataloader) does in the reference, plangen_base.py:994);
weights are replicated; no collective inside the decode loop.  NCCL (gloo on CPU) is used only to
gather generated images and timing scalars."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_batches(n_batches: int, rank: int, world: int) -> List[int]:
    """Indices of the batches this rank decodes."""
    return list(range(rank, n_batches, world))


def gather_images(local: torch.Tensor, world: int) -> torch.Tensor:
    """all_gather of a rank's (B, 3, H, W) uint8 images -> (world*B, 3, H, W) in rank order."""
    if world == 1 or not dist.is_initialized():
        return local
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local.contiguous())
    return torch.cat(out, dim=0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def interleave_results(per_rank: Sequence[Sequence[int]], world: int) -> List[int]:
    """Inverse of shard_batches: global batch order from rank-major lists."""
    out = []
    for i in range(max(len(p) for p in per_rank)):
        for r in range(world):
            if i < len(per_rank[r]):
                out.append(per_rank[r][i])
    return out
