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


# ---- the data-parallel step around the loss path (train.py:59-60, 211-216) ---------------------------------------------
def bind_to_gpu_numa(cuda_index: int) -> Dict[str, object]:
    """Pin this process (and therefore the pinned host buffers it allocates afterwards: first touch) to the CPU cores NVML
    reports as local to the GPU.  One process per GPU: without this every rank of a torchrun job may run on -- and stage its
    batches from -- the same NUMA node.  Returns what was done (for the bench line); never raises."""
    import os
    info: Dict[str, object] = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed), first_cpu=allowed[0], last_cpu=allowed[-1])
        try:
            info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
    except Exception as e:          # no NVML / no permission: run unbound
        info["error"] = repr(e)[:80]
    return info


class LossAllReduce:
    """The path's one collective: the (K,) vector of per-term loss sums, all-reduced (SUM) over the ranks and divided by the
    global batch -- what replaces ``DataParallel``'s gather of the ``(B,)`` loss vectors (train.py:59-60, 211-214).  Works on
    preallocated tensors and issues one collective on the current stream: capturable into the step's CUDA graph."""

    def __init__(self, n_terms: int, global_batch: int, device, group: Optional[dist.ProcessGroup] = None):
        self.vec = torch.zeros(n_terms, dtype=torch.float32, device=device)
        self.global_batch, self.group = float(global_batch), group

    def __call__(self, loss_matrix: torch.Tensor) -> torch.Tensor:
        """``loss_matrix`` (K, B_local) -> (K,) means over the GLOBAL batch (in ``self.vec``)."""
        torch.sum(loss_matrix, dim=1, out=self.vec)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.vec, op=dist.ReduceOp.SUM, group=self.group)
        self.vec.div_(self.global_batch)
        return self.vec


class GradBucketAllReduce:
    """Stand-in for the parameter-gradient all-reduce of the surrounding data-parallel step (21.57 M fp32 =
    86.3 MB for PWC-Net + depth net + pose net), bucketed like DDP and issued on its own stream so that it overlaps the loss
    step.  ``launch()`` makes the side stream wait for the current stream's position and enqueues the buckets; ``join()``
    makes the current stream wait for them."""

    def __init__(self, device, numel: int = 21_570_000, bucket_mb: float = 25.0, group: Optional[dist.ProcessGroup] = None):
        per = max(1, int(bucket_mb * (1 << 20) / 4))
        self.buckets = [torch.zeros(min(per, numel - o), dtype=torch.float32, device=device) for o in range(0, numel, per)]
        self.nbytes = 4 * numel
        self.stream = torch.cuda.Stream(device=device)
        self.group = group

    def launch(self) -> None:
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for b in self.buckets:
                if dist.is_available() and dist.is_initialized():
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)

    def join(self) -> None:
        torch.cuda.current_stream().wait_stream(self.stream)
