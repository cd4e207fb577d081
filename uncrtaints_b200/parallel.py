"""Data-parallel plumbing: one process per GPU, batch sharded by sample, ONE all-reduce per step.

The reference is single-process; the path shards naturally by batch (§8e): encoder GroupNorm,
L-TAE and aggregation are per sample; decoder BatchNorm statistics and the MGNLL batch-summed log-determinant
stay per rank (reference semantics has no SyncBN).  The 570 010 fp32 gradients (2.28 MB) live in one flat
buffer that ``param.grad`` views alias, so the step needs exactly one NCCL all-reduce over NVLink
(latency-bound at this size; nothing to bucket or overlap).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_batch(n: int, rank: int, world: int) -> slice:
    """Samples [rank*n/world, (rank+1)*n/world) of a global batch of n (n divisible by world)."""
    if n % world != 0:
        raise ValueError(f"global batch {n} not divisible by world size {world}")
    per = n // world
    return slice(rank * per, (rank + 1) * per)


class FlatGradAllReduce:
    """Owns a flat fp32 gradient buffer; ``p.grad`` of every parameter is a view into it."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero_(self) -> None:
        self.flat.zero_()

    def all_reduce_mean(self) -> None:
        """Sum over ranks, scale by 1/world: the gradient of the mean of the per-rank losses."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / dist.get_world_size(self.group))
