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
    """Owns a flat fp32 gradient buffer; ``p.grad`` of every parameter is a view into it.  The parameters are flagged so
    that ``uncrtaints_b200.UNCRTAINTS``'s backward accumulates straight into these views (the C ABI adds into its gradient
    slots) instead of handing autograd one temporary per parameter; call ``zero_()`` before every backward.

    The aliasing is ENFORCED, not assumed: ``optimizer.zero_grad()`` of torch >= 2.0 sets ``p.grad = None`` (the reference's
    own step does that, base_model.py:120), after which autograd would create fresh gradient tensors and an all-reduce of the
    flat buffer would silently reduce stale zeros.  ``ensure_attached()`` -- called by ``all_reduce_mean()`` and
    ``zero_()`` -- folds any such foreign gradient into the flat buffer and re-points ``p.grad`` at its view.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        self.reattached = 0                      # how many gradients had to be folded back (0 in a well-behaved loop)
        off = 0
        for p in self.params:
            n = p.numel()
            v = self.flat[off:off + n].view_as(p)
            self.views.append(v)
            p.grad = v
            p._ub200_grad_in_place = True
            off += n

    def ensure_attached(self, discard: bool = False) -> None:
        """Every ``p.grad`` must be its view of the flat buffer.  If it is not, somebody set it to None since the last
        ``zero_()`` (``zero_grad(set_to_none=True)`` semantics: None means zero), so the view's content is stale: a foreign
        tensor (autograd allocated it during backward because the view was gone) REPLACES the view's content, a gradient
        that is still None zeroes it -- unless ``discard`` (we are about to zero the whole buffer anyway)."""
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is not None and g.data_ptr() == v.data_ptr() and g.shape == v.shape:
                continue
            if not discard:
                if g is None:
                    v.zero_()
                else:
                    v.copy_(g.to(device=v.device, dtype=torch.float32))
            p.grad = v
            self.reattached += 1

    def zero_(self) -> None:
        self.ensure_attached(discard=True)
        self.flat.zero_()

    zero_grad = zero_         # bucket-aware replacement of optimizer.zero_grad(): zeroes in place, grads stay views

    def all_reduce_mean(self) -> None:
        """Mean over ranks of the flat gradient (the gradient of the mean of the per-rank losses): ONE collective; NCCL's
        AVG reduction applies the 1/world factor inside the all-reduce kernel (no separate scaling launch)."""
        self.ensure_attached()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:                                 # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / dist.get_world_size(self.group))


class HostToDevicePrefetcher:
    """Double-buffered H2D input pipeline: the pinned-host batch of step k+1 is copied on a side stream while step k
    computes (the reference moves every batch synchronously in `prepare_data_multi`, train_reconstruct.py:161-179, and
    `BaseModel.set_input`, base_model.py:87-91).  Buffers are recycled only after the step that consumed them finished."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.slots = [None] * depth          # device tensors per slot
        self.ready = [None] * depth          # event: copy into slot finished
        self.free = [None] * depth           # event: compute that consumed the slot finished
        self.head = 0                        # next slot to fill
        self.tail = 0                        # next slot to consume
        self.pending = 0

    def submit(self, *host_tensors):
        """Start the asynchronous copy of one batch (pinned host tensors)."""
        i = self.head
        if self.slots[i] is None:
            self.slots[i] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
        with torch.cuda.stream(self.copy_stream):
            if self.free[i] is not None:
                self.copy_stream.wait_event(self.free[i])
            for d, h in zip(self.slots[i], host_tensors):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.ready[i] = ev
        self.head = (i + 1) % self.depth
        self.pending += 1

    def get(self):
        """Device tensors of the oldest submitted batch; the current stream waits for its copy."""
        assert self.pending > 0, "get() without submit()"
        i = self.tail
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        self.tail = (i + 1) % self.depth
        self.pending -= 1
        self._last = i
        return self.slots[i]

    def release(self):
        """Call after the step that used the tensors of the last get() has been enqueued."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[self._last] = ev


class HostScalarReader:
    """Device -> host read of a per-step scalar (the loss) without draining the GPU: the copy into pinned host memory runs on
    a side stream behind an event recorded after the step, and the value is consumed one step later.  Every step's scalar
    is still read; the host just never waits for the step it has only just enqueued (``loss.item()`` would)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.done = [None, None]
        self.keep = [None, None]
        self.i = 0
        self.bytes_per_read = 4

    def submit(self, scalar: torch.Tensor) -> None:
        """Enqueue the read of `scalar` (0-d CUDA tensor produced on the current stream)."""
        k = self.i & 1
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self.host[k].copy_(scalar.detach(), non_blocking=True)
            scalar.record_stream(self.stream)
            d = torch.cuda.Event()
            d.record(self.stream)
        self.done[k], self.keep[k] = d, scalar
        self.i += 1

    def previous(self):
        """Value submitted one call before the latest one (None if there is none yet)."""
        if self.i < 2:
            return None
        k = self.i & 1                      # slot of submission i-2
        self.done[k].synchronize()
        return float(self.host[k])

    def latest(self):
        """Value of the latest submission (waits for that step to finish)."""
        if self.i < 1:
            return None
        k = (self.i - 1) & 1
        self.done[k].synchronize()
        return float(self.host[k])
