"""Data-parallel plumbing around the loss path: one process per GPU (torchrun), contiguous batch shards,
and the only collective the path needs — an all-reduce of the per-term loss sums for logging.

Every reduction inside the loss path is per sample (``mean((1,2,3))``), so a rank's shard is
computed with no exchange at all and gives bit-identical per-sample numbers to the unsharded batch; the
reference's ``nn.DataParallel`` gather of the ``(B,)`` loss vectors (train.py:59-60, 211-214) becomes a
K-float all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced slice [start, stop) of the batch owned by ``rank`` (first ranks take the remainder)."""
    if not (0 <= rank < world) or global_batch < 0:
        raise ValueError("bad rank/world/batch: %d/%d/%d" % (rank, world, global_batch))
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_tensors(tensors: Sequence[torch.Tensor], rank: int, world: int):
    """Slice dim 0 of every tensor to this rank's shard."""
    out = []
    for t in tensors:
        a, b = shard_range(t.shape[0], rank, world)
        out.append(t[a:b].contiguous())
    return out


def global_loss_means(local_losses: Dict[str, torch.Tensor], global_batch: int, group: Optional[dist.ProcessGroup] = None) -> Dict[str, torch.Tensor]:
    """mean over the GLOBAL batch of every (B_local,) loss vector: local sums, one all-reduce(SUM) of a K-vector,
    divide by the global batch."""
    keys = [k for k, v in local_losses.items() if v.dim() == 1]
    vec = torch.stack([local_losses[k].detach().sum() for k in keys])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return {k: vec[i] / float(global_batch) for i, k in enumerate(keys)}


def weighted_total(loss_means: Dict[str, torch.Tensor], weights: Dict[str, float]) -> torch.Tensor:
    """train.py:211-214 on the already batch-averaged terms."""
    return sum(weights[k] * v for k, v in loss_means.items() if k in weights)
