"""Fused Adam over flat buffers.

The reference steps ``torch.optim.Adam(netG.parameters(), lr=config.lr)`` with an ``ExponentialLR`` scheduler
(model/src/backbones/base_model.py:48-51,115-122): ``optimizer_G.zero_grad()`` + ``optimizer_G.step()`` are two multi-tensor
launch chains over 91 small tensors per step.  ``FusedAdam`` is a ``torch.optim.Optimizer`` (so ``ExponentialLR`` and
``state_dict`` work on it unchanged) whose parameters, gradients and moments live in four flat fp32 buffers; ``step()`` is ONE
kernel of libuncrtaints_b200.so (``ub200_adam_step``) that also folds the data-parallel 1/world gradient scale and re-zeroes the
gradient buffer for the next backward.  No host synchronisation, no per-tensor launches.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _lib
from .parallel import FlatGradAllReduce


class FusedAdam(torch.optim.Optimizer):
    """Adam with torch.optim.Adam's defaults and update rule (no amsgrad, L2 weight decay) over one flat buffer.

    ``params``: the model's parameters (fp32, one CUDA device).  Their ``.data`` is re-pointed into a flat parameter buffer
    and their ``.grad`` into the flat gradient buffer of a ``FlatGradAllReduce`` (created here unless ``bucket`` is given), so
    ``UNCRTAINTS.backward`` accumulates straight into it.  ``step()`` zeroes the gradients (``zero_grad()`` is then a no-op kept
    for the reference's call order, base_model.py:120).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, bucket: Optional[FlatGradAllReduce] = None, grad_scale: float = 1.0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdam: no parameters")
        if any((not p.is_cuda) or p.dtype != torch.float32 for p in params):
            raise RuntimeError("uncrtaints_b200.FusedAdam needs float32 CUDA parameters (no CPU fallback)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.bucket = bucket if bucket is not None else FlatGradAllReduce(params)
        if [id(p) for p in self.bucket.params] != [id(p) for p in params]:
            raise ValueError("FusedAdam: the bucket must own exactly these parameters, in this order")
        dev = params[0].device
        n = self.bucket.flat.numel()
        self.flat_params = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            view = self.flat_params[off:off + k].view_as(p)
            view.copy_(p.data)
            p.data = view                         # the module's parameters now alias the flat buffer
            off += k
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad_scale = float(grad_scale)
        self._step = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.bucket.ensure_attached()             # grads set to None / re-created by a foreign zero_grad are folded back in
        g = self.param_groups[0]
        self._step += 1
        b1, b2 = g["betas"]
        dev = self.flat_params.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ub200_adam_step(self.flat_params.data_ptr(), self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(),
                                                  self.exp_avg_sq.data_ptr(), self.flat_params.numel(), self._step, float(g["lr"]),
                                                  float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                                                  self.grad_scale, 1, torch.cuda.current_stream(dev).cuda_stream), "ub200_adam_step")
        return loss

    def zero_grad(self, set_to_none: bool = False):
        """The gradient buffer is re-zeroed inside ``step()``; an explicit call zeroes it in place (never sets grads to None:
        they must stay views of the flat buffer)."""
        self.bucket.zero_()

    def state_dict(self):
        sd = super().state_dict()
        sd["fused"] = {"step": self._step, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone()}
        return sd

    def load_state_dict(self, state_dict):
        fused = state_dict.get("fused")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "fused"})
        if fused is not None:
            self._step = int(fused["step"])
            self.exp_avg.copy_(fused["exp_avg"])
            self.exp_avg_sq.copy_(fused["exp_avg_sq"])
