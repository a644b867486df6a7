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
    """all_gather of a rank's (B_r, 3, H, W) uint8 images -> (sum B_r, 3, H, W) in rank order.  Ranks may hold
    different image counts (`shard_batches` hands out unequal shares when n_batches % world != 0): the counts are
    exchanged first and every rank contributes a buffer padded to the largest, so all ranks always enter the same two
    collectives - a rank with nothing to send passes an empty (0, 3, H, W) tensor, it must not skip the call."""
    if world == 1 or not dist.is_initialized():
        return local
    local = local.contiguous()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    top = max(counts)
    if top == 0:
        return local
    if local.shape[0] < top:
        pad = torch.zeros((top - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(out, local)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


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
