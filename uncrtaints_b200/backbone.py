"""Drop-in ``UNCRTAINTS`` generator whose forward/backward run in libuncrtaints_b200.so.

Boundary: the reference constructs ``uncrtaints.UNCRTAINTS(...)`` with keywords
(model/src/model_utils.py:86-108), calls ``netG(real_A, batch_positions=dates)``
(model/src/backbones/base_model.py:63), reads ``mean_idx`` / ``vars_idx`` / ``variance``
(base_model.py:58,82-83) and saves / loads ``state_dict()`` (model_utils.py:117-219).  This class keeps the
constructor signature (uncrtaints.py:231-254), the forward signature (uncrtaints.py:391) and the exact
module tree / state-dict keys (real ``nn.Conv2d`` / ``nn.GroupNorm`` / ``nn.BatchNorm2d`` / ``nn.Linear`` /
``nn.Conv1d`` instances as parameter holders, so ``netG.apply(weight_init)`` and checkpoints interoperate)
-- but none of the sub-modules computes anything: ``forward`` hands raw device pointers to the C ABI.

There is no PyTorch fallback: without the CUDA library or on a CPU tensor ``forward`` raises.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib

S2_BANDS = 13
_WIDTH, _HID, _SE, _HEADS, _DK, _DMODEL = 128, 256, 32, 16, 4, 256


class _Holder(nn.Module):
    """Parameter container: the compute lives in the fused CUDA path of ``UNCRTAINTS.forward``."""

    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is a parameter holder of the fused B200 path; call UNCRTAINTS.forward")


def _norm_layer(kind: str, channels: int) -> nn.Module:
    # get_norm_layer (uncrtaints.py:16-22) / ConvLayer norm choice (utae.py:464-475)
    if kind == "batch":
        return nn.BatchNorm2d(channels)
    if kind == "group":
        return nn.GroupNorm(num_channels=channels, num_groups=4)
    raise NotImplementedError(f"norm '{kind}' is not supported by the B200 path (group | batch)")


class ConvLayer(_Holder):
    """in_conv / out_conv holder: keys ``conv.0`` (Conv2d k=1), ``conv.1`` (norm) (utae.py:453-497)."""

    def __init__(self, cin: int, cout: int, norm: str, last_relu: bool, k: int = 1, p: int = 0):
        super().__init__()
        layers: List[nn.Module] = [nn.Conv2d(cin, cout, kernel_size=k, padding=p, stride=1, padding_mode="reflect")]
        if norm in ("batch", "group"):
            layers.append(_norm_layer(norm, cout))
        if last_relu:
            layers.append(nn.ReLU())
        self.conv = nn.Sequential(*layers)


class ConvBlock(_Holder):
    """keys ``conv.conv.N`` (utae.py:500-520)."""

    def __init__(self, cin: int, cout: int, norm: str, last_relu: bool = True):
        super().__init__()
        self.conv = ConvLayer(cin, cout, norm, last_relu)


class ResidualConvBlock(_Holder):
    """ResidualConvBlock(nkernels=[128, 128], k=3, p=1) holder (uncrtaints.py:24-69): keys ``conv{1,2,3}.conv.0`` (Conv2d 3x3, reflect
    padding, bias) and ``conv{1,2,3}.conv.1`` (norm); every layer ends in a ReLU."""

    def __init__(self, width: int, norm: str = "batch"):
        super().__init__()
        self.conv1 = ConvLayer(width, width, norm, last_relu=True, k=3, p=1)
        self.conv2 = ConvLayer(width, width, norm, last_relu=True, k=3, p=1)
        self.conv3 = ConvLayer(width, width, norm, last_relu=True, k=3, p=1)


class SE(_Holder):
    """keys ``fc.0.weight`` [32,256], ``fc.2.weight`` [256,32] (uncrtaints.py:82-97)."""

    def __init__(self, inp: int, oup: int, expansion: float = 0.25):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(oup, int(inp * expansion), bias=False), nn.GELU(),
                                nn.Linear(int(inp * expansion), oup, bias=False), nn.Sigmoid())


class PreNorm(_Holder):
    """keys ``norm.*`` and ``fn.*`` (uncrtaints.py:72-79)."""

    def __init__(self, dim: int, fn: nn.Module, norm: str):
        super().__init__()
        self.norm = _norm_layer(norm, dim)
        self.fn = fn


class MBConv(_Holder):
    """MBConv(inp, oup, expansion=2, downsample=False) holder (uncrtaints.py:100-146): keys ``conv.norm``,
    ``conv.fn.{0,1,3,4,6,7,8}``."""

    def __init__(self, inp: int, oup: int, expansion: int = 2, norm: str = "batch"):
        super().__init__()
        hidden = int(inp * expansion)
        body = nn.Sequential(
            nn.Conv2d(inp, hidden, 1, stride=1, padding=0, bias=False),
            _norm_layer(norm, hidden),
            nn.GELU(),
            nn.Conv2d(hidden, hidden, 3, stride=1, padding=1, padding_mode="reflect", groups=hidden, bias=False),
            _norm_layer(norm, hidden),
            nn.GELU(),
            SE(inp, hidden),
            nn.Conv2d(hidden, oup, 1, stride=1, padding=0, bias=False),
            _norm_layer(norm, oup),
        )
        self.conv = PreNorm(inp, body, norm)


class MultiHeadAttentionSmall(_Holder):
    """keys ``Q`` [16,4], ``fc1_k.{weight,bias}`` (ltae.py:312-339)."""

    def __init__(self, n_head: int, d_k: int, d_in: int):
        super().__init__()
        self.n_head, self.d_k, self.d_in = n_head, d_k, d_in
        self.Q = nn.Parameter(torch.zeros((n_head, d_k)))
        nn.init.normal_(self.Q, mean=0, std=np.sqrt(2.0 / d_k))
        self.fc1_k = nn.Linear(d_in, n_head * d_k)
        nn.init.normal_(self.fc1_k.weight, mean=0, std=np.sqrt(2.0 / d_k))


class LTAE2dtiny(_Holder):
    """keys ``inconv``, ``attention_heads``, ``in_norm`` (ltae.py:145-194)."""

    def __init__(self, in_channels: int, n_head: int, d_k: int, d_model: int, positional_encoding: bool, T: int = 1000):
        super().__init__()
        self.in_channels, self.n_head, self.d_model, self.T = in_channels, n_head, d_model, T
        self.inconv = nn.Conv1d(in_channels, d_model, 1)
        self.use_positional_encoding = positional_encoding
        self.attention_heads = MultiHeadAttentionSmall(n_head=n_head, d_k=d_k, d_in=d_model)
        self.in_norm = nn.GroupNorm(num_groups=n_head, num_channels=in_channels)


class MultiHeadAttention(MultiHeadAttentionSmall):
    """Same parameters as the Small form (ltae.py:244-262); the dropout on the attention is 0 in UNCRTAINTS (use_dropout=False)."""


class LTAE2d(_Holder):
    """Full L-TAE of the ``use_v`` variant (ltae.py:10-94): keys ``inconv``, ``attention_heads``, ``in_norm``, ``out_norm``,
    ``mlp.0`` (Linear 256->128), ``mlp.1`` (BatchNorm1d)."""

    def __init__(self, in_channels: int, n_head: int, d_k: int, d_model: int, mlp: List[int], positional_encoding: bool,
                 dropout: float = 0.2, T: int = 1000):
        super().__init__()
        assert mlp[0] == d_model and len(mlp) == 2
        self.in_channels, self.n_head, self.d_model, self.T = in_channels, n_head, d_model, T
        self.inconv = nn.Conv1d(in_channels, d_model, 1)
        self.use_positional_encoding = positional_encoding
        self.attention_heads = MultiHeadAttention(n_head=n_head, d_k=d_k, d_in=d_model)
        self.in_norm = nn.GroupNorm(num_groups=n_head, num_channels=in_channels)
        self.out_norm = nn.GroupNorm(num_groups=n_head, num_channels=mlp[-1])
        self.mlp = nn.Sequential(nn.Linear(mlp[0], mlp[1]), nn.BatchNorm1d(mlp[1]), nn.ReLU())
        self.dropout = nn.Dropout(dropout)


class Compact_Temporal_Aggregator(_Holder):
    """Holds the dropout module of the aggregator (uncrtaints.py:149-154); its ``p`` is honoured."""

    def __init__(self, mode: str = "att_group"):
        super().__init__()
        self.mode = mode
        self.attn_dropout = nn.Dropout(0.1)


def fold_ltae(te, batch_positions: Optional[torch.Tensor], B: int, T: int, device, with_pe: bool = False) -> tuple:
    """Fold the input-independent query into the key projection (differentiable, weights only).

    Reference chain (ltae.py:210-224,347-350,432-433): score = Q_h . (W_k (W_in gn(x) + b_in + pe) + b_k)_h / 2.
    Returns Ap [16,128] = (Ak W_in) diag(gamma) and e [B,T,16] = pe Ak^T + t-independent constants, such that
    score[h,t] = (Ap[h] . x_hat[:,t] + e[b,t,h]) / 2 with x_hat the affine-free GroupNorm of the pooled features.
    The full LTAE2d of the ``use_v`` variant computes the same scores (ltae.py:103-120,276-286,404-407); ``with_pe`` also
    returns the positional table [B,T,256] its value path adds to the projected features (ltae.py:107-116).
    """
    mh = te.attention_heads
    nh, dk = mh.n_head, mh.d_k
    wk = mh.fc1_k.weight.view(nh, dk, -1)                       # [16,4,256]
    ak = torch.einsum("hd,hdk->hk", mh.Q, wk)                   # [16,256]
    a = ak @ te.inconv.weight[:, :, 0]                          # [16,128]
    ap = a * te.in_norm.weight[None, :]
    const = ak @ te.inconv.bias + (mh.Q * mh.fc1_k.bias.view(nh, dk)).sum(dim=1) + a @ te.in_norm.bias   # [16]
    pe = None
    if te.use_positional_encoding and batch_positions is not None:
        d = te.d_model // nh
        # PositionalEncoder (positional_encoding.py:11-13,20-29): float32 denominators, even->sin, odd->cos, tiled x n_head
        denom = torch.pow(torch.tensor(float(te.T)), 2 * (torch.arange(0, d).float() // 2) / d).to(device)
        tab = batch_positions.to(ak.dtype)[:, :, None] / denom.to(ak.dtype)[None, None, :]
        tab = torch.stack([torch.sin(tab[:, :, 0::2]), torch.cos(tab[:, :, 1::2])], dim=-1).reshape(B, T, d)
        pe = tab.repeat(1, 1, nh)                               # [B,T,256]
        e = pe @ ak.t() + const[None, None, :]
    else:
        e = const[None, None, :].expand(B, T, nh)
    if with_pe:
        return ap.contiguous(), e.contiguous(), (pe.detach().contiguous() if pe is not None else None)
    return ap.contiguous(), e.contiguous()


class _UncrtaintsFunction(torch.autograd.Function):
    """forward/backward through ub200_forward / ub200_backward."""

    @staticmethod
    def forward(ctx, net, x, keep_mask, need_grad, extra_buffers, v_keep_mask, *tensors):
        L = _lib.lib()
        desc, slots, buffers = net._describe(x, need_grad, extra_buffers)
        table = [0] * len(net._slot_names)
        for slot, t in zip(slots, tensors):
            table[slot] = t.data_ptr()
        for slot, t in buffers:
            table[slot] = t.data_ptr()
        ws_bytes = L.ub200_workspace_bytes(desc)
        if ws_bytes == 0:
            raise NotImplementedError(
                f"unsupported configuration for the B200 path: B={desc.B} T={desc.T} C_in={desc.C_in} H={desc.H} W={desc.W} "
                "(need H, W multiples of 32, T <= 64, C_in <= 16)")
        # The workspace lives from this forward to its backward only (at B=32, T=5 it is 102 GB: two would not fit in 180 GB).
        # ``net.keep_workspace = True`` (tests / debugging: ub200_workspace_tap) additionally keeps the last one on the module.
        net._last_workspace = None
        with torch.cuda.device(x.device):            # the C ABI launches on the CURRENT device: make it the tensors' device
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            out = torch.empty((desc.B, 1, desc.out_dim, desc.H, desc.W), dtype=torch.float32, device=x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            params = _lib.ptr_table(table)
            _lib.check(L.ub200_forward_v(desc, x.data_ptr(), params, keep_mask.data_ptr() if keep_mask is not None else None,
                                         v_keep_mask.data_ptr() if v_keep_mask is not None else None,
                                         out.data_ptr(), ws.data_ptr(), ws_bytes, stream), "ub200_forward")
        if net.keep_workspace:
            net._last_workspace = (desc, ws)
        if need_grad:
            ctx.net, ctx.desc, ctx.ws, ctx.table, ctx.slots = net, desc, ws, table, slots
            ctx.keep_mask, ctx.v_keep_mask, ctx.extra_buffers = keep_mask, v_keep_mask, extra_buffers
            ctx.save_for_backward(x, out, *tensors)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        L = _lib.lib()
        if ctx.ws is None:
            raise RuntimeError("uncrtaints_b200: backward through UNCRTAINTS.forward a second time -- the activation workspace is "
                               "released after the first backward (retain_graph is not supported); run the forward again")
        x, out, *tensors = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        # The C ABI ACCUMULATES into every gradient slot.  Parameters whose .grad is owned by a FlatGradAllReduce (flagged
        # _ub200_grad_in_place) receive their gradient directly in that buffer and autograd gets None for them: no zeroed
        # staging buffer and no per-parameter add_ kernel.  Everything else goes through a zero-initialised flat buffer.
        def in_place(t):
            if not (t.is_leaf and getattr(t, "_ub200_grad_in_place", False)):
                return False
            g = t.grad
            return (g is not None and g.is_cuda and g.dtype == torch.float32
                    and g.is_contiguous() and g.shape == t.shape)
        direct = [in_place(t) for t in tensors]
        sizes = [0 if d else t.numel() for t, d in zip(tensors, direct)]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=x.device)
        gtable = [0] * len(ctx.table)
        views, off = [], 0
        for slot, t, n, d in zip(ctx.slots, tensors, sizes, direct):
            if d:
                gtable[slot] = t.grad.data_ptr()
                views.append(None)
                continue
            v = flat[off:off + n]
            gtable[slot] = v.data_ptr()
            views.append(v.view_as(t))
            off += n
        km, vkm = ctx.keep_mask, ctx.v_keep_mask
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(L.ub200_backward_v(ctx.desc, x.data_ptr(), _lib.ptr_table(ctx.table), km.data_ptr() if km is not None else None,
                                          vkm.data_ptr() if vkm is not None else None,
                                          out.data_ptr(), grad_out.data_ptr(), _lib.ptr_table(gtable), ctx.ws.data_ptr(),
                                          ctx.ws.numel(), stream), "ub200_backward")
        ctx.ws = None
        return (None, None, None, None, None, None, *views)


class UNCRTAINTS(nn.Module):
    """Same constructor as the reference (uncrtaints.py:231-254).  Supported: the UnCRtainTS architecture with MBConv blocks
    (encoder_widths=[128], decoder_widths=[128]*k, block_type='mbconv', agg_mode='att_group', n_head=16, d_model=256, d_k=4,
    reflect padding, group/batch norms, covmode diag|iso|uni|None) and its constructor variants ``block_type='residual'``
    (ResidualConvBlock: 3x3 convolutions as implicit GEMMs, :24-69,291-294,324-327), ``use_v`` (full LTAE2d value path + include_v,
    :300-314,414-417), ``is_mono`` (single date, no temporal encoder, :296,418) and ``separate_out`` (two output convolutions,
    :376-379,424-430); other combinations (other widths / aggregation modes) raise NotImplementedError, as the reference does
    for its own unsupported ones (:320)."""

    def __init__(self, input_dim, encoder_widths=[128], decoder_widths=[128, 128, 128, 128, 128], out_conv=[S2_BANDS],
                 out_nonlin_mean=False, out_nonlin_var="relu", agg_mode="att_group", encoder_norm="group",
                 decoder_norm="batch", n_head=16, d_model=256, d_k=4, pad_value=0, padding_mode="reflect",
                 positional_encoding=True, covmode="diag", scale_by=1, separate_out=False, use_v=False,
                 block_type="mbconv", is_mono=False, gemm_backend: Optional[int] = None):
        super().__init__()
        if list(encoder_widths) != [_WIDTH] or decoder_widths is None or any(w != _WIDTH for w in decoder_widths):
            raise NotImplementedError("B200 path: encoder_widths must be [128] and decoder_widths [128]*k")
        if block_type not in ("mbconv", "residual"):
            raise NotImplementedError                             # as the reference (uncrtaints.py:294,327)
        if agg_mode != "att_group":
            raise NotImplementedError("B200 path: only agg_mode='att_group' is built")
        if block_type == "residual" and "none" in (encoder_norm, decoder_norm):
            raise NotImplementedError("B200 path: residual blocks need a group or batch norm")
        if (n_head, d_model, d_k) != (_HEADS, _DMODEL, _DK) or padding_mode != "reflect" or len(out_conv) != 1:
            raise NotImplementedError("B200 path: n_head=16, d_model=256, d_k=4, padding_mode='reflect', single out_conv layer")
        if input_dim > 16:
            raise NotImplementedError("B200 path: at most 16 input channels")
        self.n_stages = len(encoder_widths)
        self.encoder_widths, self.decoder_widths, self.out_widths = encoder_widths, decoder_widths, out_conv
        self.is_mono, self.use_v, self.block_type = is_mono, use_v, block_type
        self.enc_dim, self.stack_dim = decoder_widths[0], sum(decoder_widths)
        self.pad_value, self.padding_mode = pad_value, padding_mode
        self.scale_by, self.separate_out = scale_by, separate_out
        self.encoder_norm, self.decoder_norm = encoder_norm, decoder_norm
        self.input_dim = input_dim

        self.in_conv = ConvBlock(input_dim, _WIDTH, norm=encoder_norm)
        def make_block(norm):                                                            # uncrtaints.py:291-294,324-327
            return MBConv(_WIDTH, _WIDTH, expansion=2, norm=norm) if block_type == "mbconv" else ResidualConvBlock(_WIDTH, norm=norm)
        self.in_block = nn.ModuleList([make_block(encoder_norm)])
        if not self.is_mono:                                                              # uncrtaints.py:296-322
            if self.use_v:
                self.temporal_encoder = LTAE2d(in_channels=_WIDTH, n_head=n_head, d_k=d_k, d_model=d_model, mlp=[d_model, _WIDTH],
                                               positional_encoding=positional_encoding)
                self.include_v = nn.Conv2d(2 * _WIDTH, _WIDTH, 1)
            else:
                self.temporal_encoder = LTAE2dtiny(in_channels=_WIDTH, n_head=n_head, d_k=d_k, d_model=d_model,
                                                   positional_encoding=positional_encoding)
            self.temporal_aggregator = Compact_Temporal_Aggregator(mode=agg_mode)
        self.out_block = nn.ModuleList([make_block(decoder_norm) for _ in decoder_widths])

        self.covmode = covmode
        covar_dim = {"uni": S2_BANDS, "iso": 1, "diag": S2_BANDS}.get(covmode, 0)      # uncrtaints.py:357-365
        self.mean_idx = S2_BANDS
        self.vars_idx = self.mean_idx + covar_dim
        self.out_dims = out_conv[-1]
        if self.out_dims != self.vars_idx:
            raise NotImplementedError(f"B200 path: out_conv[-1]={self.out_dims} must equal 13 + covar_dim = {self.vars_idx}")
        if covar_dim and out_nonlin_var != "softplus":
            raise NotImplementedError("B200 path: out_nonlin_var must be 'softplus' (what the CLI forces, train_reconstruct.py:61)")
        self.var_eps = 1e-9 if self.scale_by == 1.0 else 1e-3                             # uncrtaints.py:374
        self.out_nonlin_mean = bool(out_nonlin_mean)
        if self.separate_out:                                                             # uncrtaints.py:376-381
            self.out_conv_mean_1 = ConvBlock(_WIDTH, S2_BANDS, norm="none", last_relu=False)
            if self.out_dims - self.mean_idx > 0:
                self.out_conv_var_1 = ConvBlock(_WIDTH, self.out_dims - S2_BANDS, norm="none", last_relu=False)
        else:
            self.out_conv = ConvBlock(_WIDTH, self.out_dims, norm="none", last_relu=False)
        self.variance = None
        self.gemm_backend = gemm_backend
        self._injected_keep_mask = None      # tests: explicit dropout keep mask uint8 [16,B,T,H,W]
        self._injected_v_keep_mask = None    # tests (use_v): explicit keep mask of the value dropout, uint8 [B*1024,128]
        self.keep_workspace = False          # debugging: keep (desc, workspace) of the last forward in _last_workspace
        self._last_workspace = None
        self._build_slots()

    # ---- parameter table ------------------------------------------------------------------------------
    def _build_slots(self):
        P0, S = _lib.UB200_P_BLOCK0, _lib.UB200_BLOCK_STRIDE
        n_blocks = 1 + len(self.out_block)
        names = [None] * (P0 + n_blocks * S)
        names[_lib.UB200_P_IN_W] = "in_conv.conv.conv.0.weight"
        names[_lib.UB200_P_IN_B] = "in_conv.conv.conv.0.bias"
        names[_lib.UB200_P_IN_NORM_W] = "in_conv.conv.conv.1.weight"
        names[_lib.UB200_P_IN_NORM_B] = "in_conv.conv.conv.1.bias"
        names[_lib.UB200_P_IN_NORM_RM] = "in_conv.conv.conv.1.running_mean"
        names[_lib.UB200_P_IN_NORM_RV] = "in_conv.conv.conv.1.running_var"
        if not self.is_mono:
            names[_lib.UB200_P_LTAE_AP] = "<Ap>"
            names[_lib.UB200_P_LTAE_E] = "<e>"
        if self.use_v:
            te = "temporal_encoder."
            for slot, key in ((_lib.UB200_P_LTAE_GN_W, te + "in_norm.weight"), (_lib.UB200_P_LTAE_GN_B, te + "in_norm.bias"),
                              (_lib.UB200_P_LTAE_WIN, te + "inconv.weight"), (_lib.UB200_P_LTAE_BIN, te + "inconv.bias"),
                              (_lib.UB200_P_LTAE_PE, "<pe>"),
                              (_lib.UB200_P_LTAE_MLP_W, te + "mlp.0.weight"), (_lib.UB200_P_LTAE_MLP_B, te + "mlp.0.bias"),
                              (_lib.UB200_P_LTAE_BN_W, te + "mlp.1.weight"), (_lib.UB200_P_LTAE_BN_B, te + "mlp.1.bias"),
                              (_lib.UB200_P_LTAE_BN_RM, te + "mlp.1.running_mean"), (_lib.UB200_P_LTAE_BN_RV, te + "mlp.1.running_var"),
                              (_lib.UB200_P_LTAE_ON_W, te + "out_norm.weight"), (_lib.UB200_P_LTAE_ON_B, te + "out_norm.bias"),
                              (_lib.UB200_P_INCV_W, "include_v.weight"), (_lib.UB200_P_INCV_B, "include_v.bias")):
                names[slot] = key
        # separate_out: the two output convolutions act on the same input, i.e. they are ONE convolution with stacked weights
        names[_lib.UB200_P_OUT_W] = "<out_w>" if self.separate_out else "out_conv.conv.conv.0.weight"
        names[_lib.UB200_P_OUT_B] = "<out_b>" if self.separate_out else "out_conv.conv.conv.0.bias"
        rel = {
            _lib.UB200_B_N0_W: "conv.norm.weight", _lib.UB200_B_N0_B: "conv.norm.bias",
            _lib.UB200_B_N0_RM: "conv.norm.running_mean", _lib.UB200_B_N0_RV: "conv.norm.running_var",
            _lib.UB200_B_W1: "conv.fn.0.weight",
            _lib.UB200_B_N1_W: "conv.fn.1.weight", _lib.UB200_B_N1_B: "conv.fn.1.bias",
            _lib.UB200_B_N1_RM: "conv.fn.1.running_mean", _lib.UB200_B_N1_RV: "conv.fn.1.running_var",
            _lib.UB200_B_WDW: "conv.fn.3.weight",
            _lib.UB200_B_N2_W: "conv.fn.4.weight", _lib.UB200_B_N2_B: "conv.fn.4.bias",
            _lib.UB200_B_N2_RM: "conv.fn.4.running_mean", _lib.UB200_B_N2_RV: "conv.fn.4.running_var",
            _lib.UB200_B_F1: "conv.fn.6.fc.0.weight", _lib.UB200_B_F2: "conv.fn.6.fc.2.weight",
            _lib.UB200_B_W2: "conv.fn.7.weight",
            _lib.UB200_B_N3_W: "conv.fn.8.weight", _lib.UB200_B_N3_B: "conv.fn.8.bias",
            _lib.UB200_B_N3_RM: "conv.fn.8.running_mean", _lib.UB200_B_N3_RV: "conv.fn.8.running_var",
        }
        if self.block_type == "residual":             # three ConvLayers per block share the block's slot range (UB200_R_*)
            rel = {}
            for l in range(3):
                base, pre = l * _lib.UB200_R_STRIDE, f"conv{l + 1}.conv."
                rel.update({base + _lib.UB200_R_W: pre + "0.weight", base + _lib.UB200_R_B: pre + "0.bias",
                            base + _lib.UB200_R_N_W: pre + "1.weight", base + _lib.UB200_R_N_B: pre + "1.bias",
                            base + _lib.UB200_R_N_RM: pre + "1.running_mean", base + _lib.UB200_R_N_RV: pre + "1.running_var"})
        for bi in range(n_blocks):
            prefix = "in_block.0." if bi == 0 else f"out_block.{bi - 1}."
            for k, v in rel.items():
                names[P0 + bi * S + k] = prefix + v
        self._slot_names = names

    _PSEUDO = ("<Ap>", "<e>", "<out_w>", "<out_b>")      # differentiable tensors derived from the parameters by tiny torch ops

    def _describe(self, x, need_grad, extra_buffers=None):
        """(ub200_desc, slots of the differentiable tensors in call order, [(slot, buffer tensor)])"""
        B, T, C, H, W = x.shape
        d = _lib.Desc()
        d.B, d.T, d.C_in, d.H, d.W = B, T, C, H, W
        d.n_dec_blocks = len(self.out_block)
        d.out_dim = self.out_dims
        d.enc_groups = 4 if self.encoder_norm == "group" else 0
        d.dec_groups = 4 if self.decoder_norm == "group" else 0
        d.training = int(self.training)
        d.need_grad = int(need_grad)
        d.mean_sigmoid = int(self.out_nonlin_mean)
        d.gemm_backend = int(self.gemm_backend if self.gemm_backend is not None else _default_backend())
        d.scale_by, d.var_eps, d.pad_value = float(self.scale_by), float(self.var_eps), float(self.pad_value)
        d.norm_eps, d.bn_momentum = 1e-5, 0.1
        # the reference applies the dropout only inside the upsampling branch, i.e. when H > 32 (uncrtaints.py:197-202)
        d.is_mono, d.use_v = int(self.is_mono), int(self.use_v)
        d.block_type = 1 if self.block_type == "residual" else 0
        d.dropout_p = float(self.temporal_aggregator.attn_dropout.p) if (H > 32 and not self.is_mono) else 0.0
        d.v_dropout_p = float(self.temporal_encoder.dropout.p) if self.use_v else 0.0
        if self.training and ((d.dropout_p > 0 and self._injected_keep_mask is None) or
                              (d.v_dropout_p > 0 and self._injected_v_keep_mask is None)):
            d.seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        d.offset = 0
        named = dict(self.named_parameters())
        bufs = dict(self.named_buffers())
        slots, buffers = [], []
        for slot, name in enumerate(self._slot_names):
            if name is None:
                continue
            if name in self._PSEUDO or name in named:
                slots.append(slot)
            elif name in bufs:
                buffers.append((slot, bufs[name]))
            elif extra_buffers and extra_buffers.get(name) is not None:
                buffers.append((slot, extra_buffers[name]))
        return d, slots, buffers

    def _tensors_in_slot_order(self, pseudo):
        named = dict(self.named_parameters())
        out = []
        for name in self._slot_names:
            if name in self._PSEUDO:
                out.append(pseudo[name])
            elif name is not None and name in named:
                out.append(named[name])
        return out

    # ---- forward --------------------------------------------------------------------------------------
    def forward(self, input, batch_positions=None):
        """input [B,T,C,H,W] float32 CUDA, batch_positions [B,T] or None -> [B,1,13+covdim,H,W] (uncrtaints.py:391-446)."""
        if not input.is_cuda:
            raise RuntimeError("uncrtaints_b200.UNCRTAINTS runs on CUDA tensors only (no CPU fallback)")
        if input.dim() != 5:
            raise NotImplementedError("B200 path expects a 5-D [B,T,C,H,W] input")
        x = input.contiguous().float()
        B, T = x.shape[:2]
        if self.is_mono and T != 1:
            raise NotImplementedError("B200 path: is_mono squeezes the time dimension (uncrtaints.py:418) and needs T == 1")
        pseudo, extra = {}, {}
        if not self.is_mono:
            pseudo["<Ap>"], pseudo["<e>"], pe = fold_ltae(self.temporal_encoder, batch_positions, B, T, x.device, with_pe=True)
            if self.use_v and pe is not None:
                extra["<pe>"] = pe.float()
        if self.separate_out:                                     # uncrtaints.py:424-430: cat of the two convolutions' outputs
            heads = [self.out_conv_mean_1] + ([self.out_conv_var_1] if self.out_dims - self.mean_idx > 0 else [])
            pseudo["<out_w>"] = torch.cat([h.conv.conv[0].weight for h in heads], dim=0).contiguous()
            pseudo["<out_b>"] = torch.cat([h.conv.conv[0].bias for h in heads], dim=0).contiguous()
        tensors = self._tensors_in_slot_order(pseudo)
        for t in tensors:
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise RuntimeError("B200 path: parameters must be contiguous float32 CUDA tensors")
        need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in tensors)
        km = self._injected_keep_mask
        if km is not None:
            km = km.to(device=x.device, dtype=torch.uint8).contiguous()
        vkm = self._injected_v_keep_mask
        if vkm is not None:
            vkm = vkm.to(device=x.device, dtype=torch.uint8).contiguous()
        out = _UncrtaintsFunction.apply(self, x, km, need_grad, extra, vkm, *tensors)
        if self.training:
            counters = [m.num_batches_tracked for m in self.modules()
                        if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None]
            if counters:
                torch._foreach_add_(counters, 1)          # one launch for all BatchNorm step counters
        if not self.covmode:
            return out[:, :, :self.mean_idx]
        return out


_BACKEND = 3     # 3 = tcgen05 tensor-core path (forward fp16 hi/lo, gradients bf16 hi/lo, fused expand-convolution backward); flags: +4 single-pass
                 # bf16 MMAs, +32 bf16 storage of the hidden tensors (39 = BASELINE config #3, reduced precision); 0 = fp32 CUDA-core comparator


def set_default_gemm_backend(backend: int) -> None:
    """See _BACKEND: default 3 (tcgen05 tensor-core path); 0 = fp32 CUDA cores."""
    global _BACKEND
    _BACKEND = int(backend)


def _default_backend() -> int:
    return _BACKEND
