"""MGNLL / GNLL losses on the B200 path, with the reference's call surface.

Mirrors model/src/losses.py: ``get_loss(config)`` (:14-32), ``calc_loss(criterion, config, out, y, var)`` (:35-43),
``MultiGaussianNLLLoss(*, full, eps, reduction, mode, chunk)(input, target, var) -> (loss, variance)`` (:288-354) and
``GaussianNLLLoss(*, full, eps, reduction)(input, target, var) -> (loss, clamped var)`` (:46-128,222-284; ``--loss GNLL``).

* ``loss`` is a CUDA tensor with autograd: 0-d for reduction 'mean' (what ``get_loss`` builds; one fused kernel computes the loss
  and both gradients) and 'sum'; ``[H,W,B]`` (MGNLL, the nested vmap's output order) / ``[B,1,13,H,W]`` (GNLL) for 'none'.
* ``variance`` is diag_embed(max(var, eps)) of shape [B,1,13,13,H,W] (losses.py:145,211).  The reference
  builds it on the CPU (a 44 MB/sample D2H copy inside the loss); here it is a tensor on the loss's device,
  built by one kernel, and ``covariance='none'`` skips it.  Callers only use ``.cpu()``, scalar multiply,
  ``.diagonal``, ``.mean`` and indexing on it (base_model.py:112-113,131; train_reconstruct.py:185-191,324-352),
  all of which work unchanged.
* ``var`` with a negative entry raises ``ValueError`` like the reference (:199-200).  Reading the device flag costs a host
  sync per call (the CPU cannot enqueue the backward kernels until the forward has drained); ``check_negative="deferred"``
  keeps the error but reads the flag of call k at call k+1 (or at ``.check()``), ``False`` never reads it.
  A NaN variance propagates to a NaN loss, as in the reference (``clamp_`` keeps NaN).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

S2_BANDS = 13


def _plane_view(t: torch.Tensor):
    """(tensor, batch stride in elements) with a contiguous [C,H,W] block per sample, without copying slices
    of the network output (base_model.py:83 passes fake_B[:, :, :13] and fake_B[:, :, 13:26])."""
    B, one, Cc, H, W = t.shape
    if t.stride(4) == 1 and t.stride(3) == W and t.stride(2) == H * W:
        return t, t.stride(0)
    t = t.contiguous()
    return t, t.stride(0)


def _raise_if_negative(flag, check_negative):
    if check_negative and check_negative != "deferred" and int(flag.item()) != 0:
        raise ValueError("var has negative entry/entries")


class _MGNLLFunction(torch.autograd.Function):
    """reduction='mean'; returns (loss, flag): flag is a device int32, 1 if any var < 0."""

    @staticmethod
    def forward(ctx, pred, target, var, eps, check_negative):
        L = _lib.lib()
        B, _, C, H, W = pred.shape
        P = H * W
        var_ch = var.shape[2]
        pred_v, pred_sb = _plane_view(pred)
        targ_v, targ_sb = _plane_view(target)
        var_v, var_sb = _plane_view(var)
        dev = pred.device
        need_grad = pred.requires_grad or var.requires_grad
        loss = torch.empty((), dtype=torch.float32, device=dev)
        scratch = torch.empty(4, dtype=torch.float64, device=dev)   # 16 B accumulators + the int flag
        flag = scratch[2:].view(torch.int32)[:1]
        dpred = torch.empty((B, 1, S2_BANDS, H, W), dtype=torch.float32, device=dev) if need_grad else None
        dvar = torch.empty((B, 1, var_ch, H, W), dtype=torch.float32, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(L.ub200_mgnll_forward(pred_v.data_ptr(), pred_sb, targ_v.data_ptr(), targ_sb, var_v.data_ptr(), var_sb,
                                             var_ch, B, P, float(eps), loss.data_ptr(),
                                             dpred.data_ptr() if need_grad else None, dvar.data_ptr() if need_grad else None,
                                             flag.data_ptr(), scratch.data_ptr(), stream), "ub200_mgnll_forward")
        _raise_if_negative(flag, check_negative)
        if need_grad:
            ctx.save_for_backward(dpred, dvar)
        ctx.mark_non_differentiable(flag)
        return loss, flag

    @staticmethod
    def backward(ctx, g, _gflag):
        L = _lib.lib()
        dpred, dvar = ctx.saved_tensors
        g = g.contiguous().float()
        gp, gv = torch.empty_like(dpred), torch.empty_like(dvar)
        with torch.cuda.device(dpred.device):
            stream = torch.cuda.current_stream(dpred.device).cuda_stream
            _lib.check(L.ub200_scale_by_scalar(dpred.data_ptr(), g.data_ptr(), gp.data_ptr(), dpred.numel(), stream), "scale")
            _lib.check(L.ub200_scale_by_scalar(dvar.data_ptr(), g.data_ptr(), gv.data_ptr(), dvar.numel(), stream), "scale")
        return gp, None, gv, None, None


class _MGNLLNoneFunction(torch.autograd.Function):
    """reduction='none': loss [H,W,B] (losses.py:206-218)."""

    @staticmethod
    def forward(ctx, pred, target, var, eps, check_negative):
        L = _lib.lib()
        B, _, C, H, W = pred.shape
        pred_v, pred_sb = _plane_view(pred)
        targ_v, targ_sb = _plane_view(target)
        var_v, var_sb = _plane_view(var)
        dev = pred.device
        loss = torch.empty((H, W, B), dtype=torch.float32, device=dev)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.ub200_mgnll_none(pred_v.data_ptr(), pred_sb, targ_v.data_ptr(), targ_sb, var_v.data_ptr(), var_sb,
                                          var.shape[2], B, H * W, float(eps), None, loss.data_ptr(), None, None, flag.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream), "ub200_mgnll_none")
        _raise_if_negative(flag, check_negative)
        ctx.save_for_backward(pred_v, targ_v, var_v)
        ctx.strides, ctx.eps = (pred_sb, targ_sb, var_sb), float(eps)
        ctx.mark_non_differentiable(flag)
        return loss, flag

    @staticmethod
    def backward(ctx, g, _gflag):
        L = _lib.lib()
        pred, target, var = ctx.saved_tensors
        B, _, C, H, W = pred.shape
        g = g.contiguous().float()
        dpred = torch.empty((B, 1, S2_BANDS, H, W), dtype=torch.float32, device=pred.device)
        dvar = torch.empty((B, 1, var.shape[2], H, W), dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(L.ub200_mgnll_none(pred.data_ptr(), ctx.strides[0], target.data_ptr(), ctx.strides[1], var.data_ptr(),
                                          ctx.strides[2], var.shape[2], B, H * W, ctx.eps, g.data_ptr(), None, dpred.data_ptr(),
                                          dvar.data_ptr(), None, torch.cuda.current_stream(pred.device).cuda_stream), "ub200_mgnll_none")
        return dpred, None, dvar, None, None


class _GNLLFunction(torch.autograd.Function):
    """reduction='mean'; returns (loss, clamped var, flag)."""

    @staticmethod
    def forward(ctx, pred, target, var, eps, full, check_negative):
        L = _lib.lib()
        B, _, C, H, W = pred.shape
        P = H * W
        pred_v, pred_sb = _plane_view(pred)
        targ_v, targ_sb = _plane_view(target)
        var_v, var_sb = _plane_view(var)
        dev = pred.device
        need_grad = pred.requires_grad or var.requires_grad
        loss = torch.empty((), dtype=torch.float32, device=dev)
        scratch = torch.empty(4, dtype=torch.float64, device=dev)
        flag = scratch[2:].view(torch.int32)[:1]
        dpred = torch.empty((B, 1, S2_BANDS, H, W), dtype=torch.float32, device=dev) if need_grad else None
        dvar = torch.empty((B, 1, S2_BANDS, H, W), dtype=torch.float32, device=dev) if need_grad else None
        var_out = torch.empty((B, 1, S2_BANDS, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(L.ub200_gnll_forward(pred_v.data_ptr(), pred_sb, targ_v.data_ptr(), targ_sb, var_v.data_ptr(), var_sb, B, P,
                                            float(eps), int(bool(full)), loss.data_ptr(), dpred.data_ptr() if need_grad else None,
                                            dvar.data_ptr() if need_grad else None, var_out.data_ptr(), flag.data_ptr(),
                                            scratch.data_ptr(), stream), "ub200_gnll_forward")
        _raise_if_negative(flag, check_negative)
        if need_grad:
            ctx.save_for_backward(dpred, dvar)
        ctx.mark_non_differentiable(var_out, flag)
        return loss, var_out, flag

    @staticmethod
    def backward(ctx, g, _g_var, _gflag):
        L = _lib.lib()
        dpred, dvar = ctx.saved_tensors
        g = g.contiguous().float()
        gp, gv = torch.empty_like(dpred), torch.empty_like(dvar)
        with torch.cuda.device(dpred.device):
            stream = torch.cuda.current_stream(dpred.device).cuda_stream
            _lib.check(L.ub200_scale_by_scalar(dpred.data_ptr(), g.data_ptr(), gp.data_ptr(), dpred.numel(), stream), "scale")
            _lib.check(L.ub200_scale_by_scalar(dvar.data_ptr(), g.data_ptr(), gv.data_ptr(), dvar.numel(), stream), "scale")
        return gp, None, gv, None, None, None


class _GNLLNoneFunction(torch.autograd.Function):
    """reduction='none': element-wise loss [B,1,13,H,W] (losses.py:122-128)."""

    @staticmethod
    def forward(ctx, pred, target, var, eps, full, check_negative):
        L = _lib.lib()
        B, _, C, H, W = pred.shape
        pred_v, pred_sb = _plane_view(pred)
        targ_v, targ_sb = _plane_view(target)
        var_v, var_sb = _plane_view(var)
        dev = pred.device
        loss = torch.empty((B, 1, S2_BANDS, H, W), dtype=torch.float32, device=dev)
        var_out = torch.empty((B, 1, S2_BANDS, H, W), dtype=torch.float32, device=dev)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.ub200_gnll_none(pred_v.data_ptr(), pred_sb, targ_v.data_ptr(), targ_sb, var_v.data_ptr(), var_sb, B, H * W,
                                         float(eps), int(bool(full)), None, loss.data_ptr(), var_out.data_ptr(), None, None,
                                         flag.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "ub200_gnll_none")
        _raise_if_negative(flag, check_negative)
        ctx.save_for_backward(pred_v, targ_v, var_v)
        ctx.strides, ctx.eps, ctx.full = (pred_sb, targ_sb, var_sb), float(eps), int(bool(full))
        ctx.mark_non_differentiable(var_out, flag)
        return loss, var_out, flag

    @staticmethod
    def backward(ctx, g, _g_var, _gflag):
        L = _lib.lib()
        pred, target, var = ctx.saved_tensors
        B, _, C, H, W = pred.shape
        g = g.contiguous().float()
        dpred, dvar = torch.empty_like(g), torch.empty_like(g)
        with torch.cuda.device(pred.device):
            _lib.check(L.ub200_gnll_none(pred.data_ptr(), ctx.strides[0], target.data_ptr(), ctx.strides[1], var.data_ptr(),
                                         ctx.strides[2], B, H * W, ctx.eps, ctx.full, g.data_ptr(), None, None, dpred.data_ptr(),
                                         dvar.data_ptr(), None, torch.cuda.current_stream(pred.device).cuda_stream), "ub200_gnll_none")
        return dpred, None, dvar, None, None, None


def _check_reduction(reduction):
    if reduction != "none" and reduction != "mean" and reduction != "sum":
        raise ValueError(reduction + " is not valid")                      # losses.py:100-101,195-196


def gaussian_nll_loss(input, target, var, full=False, eps=1e-8, reduction="mean", check_negative=True, _flag_out=None):
    """gaussian_nll_loss (losses.py:46-128) -> (loss, clamped variance).  Heteroscedastic [B,1,13,H,W] variances (what the
    'uni' covariance head produces); reduction 'mean' (what get_loss builds, losses.py:16), 'sum' or 'none'."""
    _check_reduction(reduction)
    if not input.is_cuda:
        raise RuntimeError("uncrtaints_b200 GNLL runs on CUDA tensors only (no CPU fallback)")
    if var.size() != input.size():
        if input.size()[:-1] == var.size() or (input.size()[:-1] == var.size()[:-1] and var.size(-1) == 1):
            raise NotImplementedError("B200 path: homoscedastic variances are not built (the network predicts one per element)")
        raise ValueError("var is of incorrect size")
    if input.dim() != 5 or input.shape[2] != S2_BANDS:
        raise NotImplementedError("B200 path: expects [B,1,13,H,W] predictions and variances")
    fn = _GNLLNoneFunction if reduction == "none" else _GNLLFunction
    loss, var_out, flag = fn.apply(input.float(), target.float(), var.float(), eps, full, check_negative)
    if reduction == "sum":
        loss = loss * float(input.numel())
    if _flag_out is not None:
        _flag_out.append(flag)
    return loss, var_out


class _DeferredCheck:
    """Mixin: the device flag of the previous call is read (host sync) at the next call or at ``check()``."""

    def check(self):
        """Deferred mode: raise the reference's ValueError if the previous call saw a negative variance."""
        flag, self._pending_flag = self._pending_flag, None
        if flag is not None and int(flag.item()) != 0:
            raise ValueError("var has negative entry/entries")


class GaussianNLLLoss(nn.Module, _DeferredCheck):
    """Same constructor / call as the reference class (losses.py:222-284): (input, target, var) -> (loss, clamped var)."""

    def __init__(self, *, full: bool = False, eps: float = 1e-8, reduction: str = "mean", check_negative=True) -> None:
        super().__init__()
        self.full, self.eps, self.reduction, self.check_negative = full, eps, reduction, check_negative
        self._pending_flag = None

    def forward(self, input, target, var):
        if self.check_negative == "deferred":
            self.check()
        flags = []
        out = gaussian_nll_loss(input, target, var, full=self.full, eps=self.eps, reduction=self.reduction,
                                check_negative=self.check_negative, _flag_out=flags)
        if self.check_negative == "deferred":
            self._pending_flag = flags[0]
        return out


def covariance_diag(var: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """diag_embed(max(var, eps)) -> [B,1,13,13,H,W] on var's device (losses.py:145,211)."""
    if not var.is_cuda:
        raise RuntimeError("uncrtaints_b200 covariance_diag runs on CUDA tensors only (no CPU fallback)")
    if var.dim() != 5 or var.shape[1] != 1 or var.shape[2] not in (1, S2_BANDS):
        raise ValueError("covariance_diag expects a [B,1,13|1,H,W] variance tensor")
    L = _lib.lib()
    var = var.detach().float()                       # the kernel reads float32
    B, _, var_ch, H, W = var.shape
    var_v, var_sb = _plane_view(var)
    cov = torch.empty((B, 1, S2_BANDS, S2_BANDS, H, W), dtype=torch.float32, device=var.device)
    with torch.cuda.device(var.device):
        stream = torch.cuda.current_stream(var.device).cuda_stream
        _lib.check(L.ub200_covariance(var_v.data_ptr(), var_sb, var_ch, B, H * W, float(eps), cov.data_ptr(), stream),
                   "ub200_covariance")
    return cov


def multi_gaussian_nll_loss(input, target, var, full=False, eps=1e-8, reduction="mean", mode="diag", chunk=None,
                            covariance="dense", check_negative=True, _flag_out=None):
    """multi_gaussian_nll_loss (losses.py:149-218).  ``full`` and ``chunk`` are accepted and ignored, as in the
    reference (the constant is always included, :143)."""
    _check_reduction(reduction)
    if mode not in ("iso", "diag"):
        raise NotImplementedError("B200 path: covmode must be 'iso' or 'diag' for MGNLL")
    if not input.is_cuda:
        raise RuntimeError("uncrtaints_b200 MGNLL runs on CUDA tensors only (no CPU fallback)")
    if input.dim() != 5 or input.shape[2] != S2_BANDS or var.shape[2] not in (1, S2_BANDS):
        raise NotImplementedError("B200 path: expects [B,1,13,H,W] predictions and [B,1,13|1,H,W] variances")
    fn = _MGNLLNoneFunction if reduction == "none" else _MGNLLFunction
    loss, flag = fn.apply(input.float(), target.float(), var.float(), eps, check_negative)
    if reduction == "sum":                           # sum over [H,W,B] = B*P times the mean
        loss = loss * float(input.shape[0] * input.shape[3] * input.shape[4])
    if _flag_out is not None:
        _flag_out.append(flag)
    variance = covariance_diag(var, eps) if covariance == "dense" else None
    return loss, variance


class MultiGaussianNLLLoss(nn.Module, _DeferredCheck):
    """Same constructor / call as the reference class (losses.py:288-354)."""

    def __init__(self, *, full: bool = False, eps: float = 1e-8, reduction: str = "mean", mode: str = "diag", chunk=None,
                 covariance: str = "dense", check_negative: bool = True) -> None:
        super().__init__()
        self.full, self.eps, self.reduction, self.mode, self.chunk = full, eps, reduction, mode, chunk
        self.covariance, self.check_negative = covariance, check_negative
        self._pending_flag = None

    def forward(self, input, target, var):
        if self.check_negative == "deferred":
            self.check()                 # flag of the previous call: that step has long finished, no pipeline bubble
        flags = []
        loss, variance = multi_gaussian_nll_loss(input, target, var, full=self.full, eps=self.eps, reduction=self.reduction,
                                                 mode=self.mode, chunk=self.chunk, covariance=self.covariance,
                                                 check_negative=self.check_negative, _flag_out=flags)
        if self.check_negative == "deferred":
            self._pending_flag = flags[0]
        return loss, variance


def get_loss(config):
    """losses.get_loss (losses.py:14-32) for the MGNLL branch; other losses are outside the hot path."""
    if config.loss == "GNLL":                      # losses.py:15-17
        criterion0 = GaussianNLLLoss(reduction="mean", eps=1e-8, full=True)
        return lambda pred, targ, var: criterion0(pred, targ, var)
    if config.loss != "MGNLL":
        raise NotImplementedError("B200 path builds the GNLL / MGNLL losses only; use the reference's losses for " + str(config.loss))
    criterion1 = MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode=config.covmode,
                                      chunk=getattr(config, "chunk_size", None))
    return lambda pred, targ, var: criterion1(pred, targ, var)


def calc_loss(criterion, config, out, y, var=None):
    """losses.calc_loss (losses.py:35-43)."""
    if config.loss in ["GNLL", "MGNLL"]:
        loss, variance = criterion(out, y, var)
    else:
        loss, variance = criterion(out, y), None
    return loss, variance
