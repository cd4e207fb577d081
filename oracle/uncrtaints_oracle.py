"""CPU restatement (oracle) of the UnCRtainTS forward / MGNLL hot path.

THIS IS TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
path (``uncrtaints_b200``) never does.

The oracle re-states, with elementary torch CPU ops (so that it runs unchanged in
fp32 and fp64), the algorithm of the reference at /root/reference (PatrickTUM/UnCRtainTS
@ 5e1f1b5).  Every function cites the reference file:line it follows.  Backward is
obtained by autograd over this restatement -- exactly how the reference obtains its own.

Parity pinning: the reference ships no tests / golden vectors, therefore
the oracle is pinned against the *live* reference imported from /root/reference in the
build container (tests/test_oracle_vs_reference.py, skipped where /root/reference is
absent) and against the committed fixtures tests/golden/*.npz that were generated from
the unmodified reference by tests/golden/make_golden.py.

Parameters are passed as a flat ``dict[str, Tensor]`` with exactly the reference's
state-dict keys, so a reference ``state_dict()`` can be fed directly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

S2_BANDS = 13
ATT_DOWN = 32  # model/src/backbones/uncrtaints.py:403 (hard-coded low-res size)


@dataclass
class OracleConfig:
    """Mirror of the UNCRTAINTS constructor arguments that change the arithmetic
    (model/src/backbones/uncrtaints.py:231-254)."""
    input_dim: int = 15
    n_enc_blocks: int = 1
    n_dec_blocks: int = 5
    encoder_norm: str = "group"
    decoder_norm: str = "batch"
    n_head: int = 16
    d_model: int = 256
    d_k: int = 4
    pad_value: float = 0.0
    covmode: str = "diag"
    scale_by: float = 10.0
    out_nonlin_mean: bool = True
    positional_encoding: bool = True
    dropout_p: float = 0.1          # Compact_Temporal_Aggregator.attn_dropout, uncrtaints.py:154
    bn_momentum: float = 0.1
    norm_eps: float = 1e-5
    use_v: bool = False             # full LTAE2d with a value output + include_v (uncrtaints.py:300-314,414-417)
    v_dropout_p: float = 0.2        # LTAE2d.dropout on the MLP-processed values (ltae.py:17,85,129)
    block_type: str = "mbconv"      # 'mbconv' | 'residual' (ResidualConvBlock: three 3x3 ConvLayers, uncrtaints.py:24-69,291-294,324-327)
    is_mono: bool = False           # single date, no temporal encoder (uncrtaints.py:296,418)
    separate_out: bool = False      # separate mean / variance output convolutions (uncrtaints.py:376-379,424-430)

    @property
    def covar_dim(self) -> int:      # uncrtaints.py:357-365
        return {"uni": S2_BANDS, "diag": S2_BANDS, "iso": 1}.get(self.covmode, 0)

    @property
    def out_dim(self) -> int:
        return S2_BANDS + self.covar_dim

    @property
    def var_eps(self) -> float:      # uncrtaints.py:374
        return 1e-9 if self.scale_by == 1.0 else 1e-3


# --------------------------------------------------------------------------------------
# elementary ops
# --------------------------------------------------------------------------------------
# FUSED = True routes the elementary ops below through the same torch.nn.functional kernels the reference's
# nn.Modules dispatch to (oneDNN conv, ATen group/batch norm, ...).  Used for the CPU-baseline timing in bench.py,
# where the point is to time what the reference executes on the host; tests/test_oracle.py checks FUSED == elementary.
FUSED = False


def set_fused(flag: bool) -> None:
    global FUSED
    FUSED = bool(flag)


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """Exact-erf GELU, nn.GELU() default (uncrtaints.py:88,128,133)."""
    if FUSED:
        return F.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def group_norm(x: torch.Tensor, groups: int, weight, bias, eps: float = 1e-5) -> torch.Tensor:
    """nn.GroupNorm over [N, C, *]: biased variance per (sample, group)
    (uncrtaints.py:22, utae.py:470-473, ltae.py:191-194)."""
    if FUSED:
        return F.group_norm(x, groups, weight, bias, eps)
    n, c = x.shape[:2]
    xg = x.reshape(n, groups, -1)
    mu = xg.mean(dim=2, keepdim=True)
    var = ((xg - mu) ** 2).mean(dim=2, keepdim=True)
    xh = ((xg - mu) / torch.sqrt(var + eps)).reshape(x.shape)
    shape = [1, c] + [1] * (x.dim() - 2)
    return xh * weight.reshape(shape) + bias.reshape(shape)


def batch_norm(x, weight, bias, running_mean, running_var, training: bool,
               momentum: float = 0.1, eps: float = 1e-5,
               new_buffers: Optional[dict] = None, key: str = ""):
    """nn.BatchNorm2d (uncrtaints.py:18).  Train: batch mean / biased var normalise, running
    stats updated with the *unbiased* var.  Eval: running stats.  Buffer updates are
    returned through ``new_buffers`` (the oracle is functional)."""
    shape = [1, -1, 1, 1]
    if FUSED and new_buffers is None:
        return F.batch_norm(x, running_mean.clone(), running_var.clone(), weight, bias, training, momentum, eps)
    if training:
        mu = x.mean(dim=(0, 2, 3))
        var = ((x - mu.reshape(shape)) ** 2).mean(dim=(0, 2, 3))
        if new_buffers is not None:
            cnt = x.numel() // x.shape[1]
            unbiased = var.detach() * (cnt / max(cnt - 1, 1))
            new_buffers[key + "running_mean"] = (1 - momentum) * running_mean + momentum * mu.detach()
            new_buffers[key + "running_var"] = (1 - momentum) * running_var + momentum * unbiased
    else:
        mu, var = running_mean, running_var
    xh = (x - mu.reshape(shape)) / torch.sqrt(var.reshape(shape) + eps)
    return xh * weight.reshape(shape) + bias.reshape(shape)


def reflect_pad1(x: torch.Tensor) -> torch.Tensor:
    """padding_mode='reflect', pad 1 (uncrtaints.py:130-131): index -1 -> 1, H -> H-2."""
    x = torch.cat([x[..., 1:2, :], x, x[..., -2:-1, :]], dim=-2)
    x = torch.cat([x[..., :, 1:2], x, x[..., :, -2:-1]], dim=-1)
    return x


def depthwise3x3_reflect(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """groups=C 3x3 cross-correlation, stride 1, reflect padding, no bias
    (uncrtaints.py:130-131).  w: [C,1,3,3]."""
    if FUSED:
        return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, groups=x.shape[1])
    xp = reflect_pad1(x)
    h, wd = x.shape[-2:]
    out = torch.zeros_like(x)
    for i in range(3):
        for j in range(3):
            out = out + xp[..., i:i + h, j:j + wd] * w[:, 0, i, j].reshape(1, -1, 1, 1)
    return out


def conv1x1(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Conv2d(k=1): per-pixel matmul.  w: [Cout, Cin, 1, 1]."""
    if FUSED:
        return F.conv2d(x, w, b)
    out = torch.einsum("nchw,oc->nohw", x, w[:, :, 0, 0])
    if b is not None:
        out = out + b.reshape(1, -1, 1, 1)
    return out


def adaptive_max_pool(x: torch.Tensor, out_hw: int = ATT_DOWN) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn.AdaptiveMaxPool2d((32,32)) for H,W multiples of 32 (uncrtaints.py:403-404).
    Returns values and flat (h*W + w) int64 argmax with the *first* maximum in row-major
    window order."""
    if FUSED:
        return F.adaptive_max_pool2d(x, (out_hw, out_hw), return_indices=True)
    n, c, h, w = x.shape
    assert h % out_hw == 0 and w % out_hw == 0
    kh, kw = h // out_hw, w // out_hw
    win = x.reshape(n, c, out_hw, kh, out_hw, kw).permute(0, 1, 2, 4, 3, 5).reshape(n, c, out_hw, out_hw, kh * kw)
    mx = win.max(dim=-1, keepdim=True).values
    # first index attaining the max
    is_max = (win == mx)
    first = torch.argmax(is_max.to(torch.uint8), dim=-1)
    val = torch.gather(win, -1, first.unsqueeze(-1)).squeeze(-1)
    wy, wx = first // kw, first % kw
    oy = torch.arange(out_hw).reshape(1, 1, out_hw, 1) * kh
    ox = torch.arange(out_hw).reshape(1, 1, 1, out_hw) * kw
    idx = (oy + wy) * w + (ox + wx)
    return val, idx


def bilinear_weights(out_size: int, in_size: int, dtype) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """1-D taps of nn.Upsample(mode='bilinear', align_corners=False) (uncrtaints.py:198-200):
    src = max((o+0.5)*in/out - 0.5, 0); i0=floor(src); i1=min(i0+1,in-1); lam = src-i0."""
    o = torch.arange(out_size, dtype=dtype)
    src = torch.clamp((o + 0.5) * (in_size / out_size) - 0.5, min=0.0)
    i0 = torch.floor(src).to(torch.int64)
    i1 = torch.clamp(i0 + 1, max=in_size - 1)
    lam = src - i0.to(dtype)
    return i0, i1, lam


def bilinear_upsample(a: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """[..., h, w] -> [..., out_h, out_w], separable 4-tap lerp."""
    if FUSED:
        lead = a.shape[:-2]
        up = F.interpolate(a.reshape(-1, 1, *a.shape[-2:]), size=(out_h, out_w), mode="bilinear", align_corners=False)
        return up.reshape(*lead, out_h, out_w)
    y0, y1, ly = bilinear_weights(out_h, a.shape[-2], a.dtype)
    x0, x1, lx = bilinear_weights(out_w, a.shape[-1], a.dtype)
    rows = a[..., y0, :] * (1 - ly).reshape(-1, 1) + a[..., y1, :] * ly.reshape(-1, 1)
    return rows[..., :, x0] * (1 - lx) + rows[..., :, x1] * lx


# --------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------
def _norm(x, p, key, kind, training, cfg: OracleConfig, new_buffers):
    """get_norm_layer (uncrtaints.py:16-22): 'group' -> GroupNorm(4), 'batch' -> BatchNorm2d."""
    if kind == "group":
        return group_norm(x, 4, p[key + "weight"], p[key + "bias"], cfg.norm_eps)
    if kind == "batch":
        return batch_norm(x, p[key + "weight"], p[key + "bias"], p[key + "running_mean"],
                          p[key + "running_var"], training, cfg.bn_momentum, cfg.norm_eps, new_buffers, key)
    raise NotImplementedError(kind)


def squeeze_excite(g: torch.Tensor, f1: torch.Tensor, f2: torch.Tensor) -> torch.Tensor:
    """SE (uncrtaints.py:82-97): g * sigmoid(F2 gelu(F1 mean_hw(g))), no biases."""
    pooled = g.mean(dim=(2, 3))
    z = gelu_erf(pooled @ f1.t())
    s = torch.sigmoid(z @ f2.t())
    return g * s.reshape(s.shape[0], s.shape[1], 1, 1)


def mbconv(x, p, pre, kind, training, cfg, new_buffers, taps: Optional[dict] = None):
    """MBConv(expansion=2, downsample=False) with PreNorm (uncrtaints.py:100-146,72-79):
    x + N3(W2 SE(gelu(N2(DW(gelu(N1(W1 N0(x))))))))."""
    k = pre + "conv."
    n0 = _norm(x, p, k + "norm.", kind, training, cfg, new_buffers)
    h1 = conv1x1(n0, p[k + "fn.0.weight"])
    g1 = gelu_erf(_norm(h1, p, k + "fn.1.", kind, training, cfg, new_buffers))
    h2 = depthwise3x3_reflect(g1, p[k + "fn.3.weight"])
    g2 = gelu_erf(_norm(h2, p, k + "fn.4.", kind, training, cfg, new_buffers))
    u = squeeze_excite(g2, p[k + "fn.6.fc.0.weight"], p[k + "fn.6.fc.2.weight"])
    y = conv1x1(u, p[k + "fn.7.weight"])
    out = x + _norm(y, p, k + "fn.8.", kind, training, cfg, new_buffers)
    if taps is not None:
        taps[pre + "h1"], taps[pre + "h2"], taps[pre + "y"], taps[pre + "out"] = h1, h2, y, out
    return out


def conv3x3_reflect(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d(k=3, padding=1, padding_mode='reflect', bias=True) (utae.py:478-487): dense cross-correlation over the
    reflect-padded input.  w: [Cout, Cin, 3, 3]."""
    if FUSED:
        return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b)
    xp = reflect_pad1(x)
    h, wd = x.shape[-2:]
    out = b.reshape(1, -1, 1, 1).expand(x.shape[0], -1, h, wd)
    for i in range(3):
        for j in range(3):
            out = out + torch.einsum("nchw,oc->nohw", xp[..., i:i + h, j:j + wd], w[:, :, i, j])
    return out


RELU_MASKS: Optional[dict] = None      # diagnostic, see residual_block


def residual_block(x, p, pre, kind, training, cfg, new_buffers, taps: Optional[dict] = None):
    """ResidualConvBlock (uncrtaints.py:24-69): x + L3(L2(L1(x))), L = ConvLayer = Conv3x3(reflect, bias) -> norm -> ReLU
    (utae.py:453-497 with last_relu=True; the final ReLU and norm are NOT omitted, see the comments at :55-60).
    Diagnostic: with the module-level RELU_MASKS = {"<pre>conv<i>": bool [N,128,H,W]} the active set of a layer's ReLU is imposed
    instead of decided (like ``pool_idx`` / ``v_relu_mask``): on the small test shapes ONE pre-activation within fp32 rounding of
    zero that an fp32 implementation decides differently from fp64 changes a weight gradient by ~1e-3 (a decoder frame pair has
    8192 pixels; measured by shifting the threshold by +-1e-6 in fp64), with no visible change of the forward output."""
    f = x
    for i in (1, 2, 3):
        k = f"{pre}conv{i}.conv."
        f = conv3x3_reflect(f, p[k + "0.weight"], p[k + "0.bias"])
        f = _norm(f, p, k + "1.", kind, training, cfg, new_buffers)
        mask = RELU_MASKS.get(f"{pre}conv{i}") if RELU_MASKS is not None else None
        f = torch.relu(f) if mask is None else f * mask.to(f.dtype)
    out = x + f
    if taps is not None:
        taps[pre + "out"] = out
    return out


def _block(x, p, pre, kind, training, cfg, new_buffers, taps):
    if cfg.block_type == "residual":
        return residual_block(x, p, pre, kind, training, cfg, new_buffers, taps)
    return mbconv(x, p, pre, kind, training, cfg, new_buffers, taps)


def positional_table(batch_positions: torch.Tensor, d: int, T: float, repeat: int) -> torch.Tensor:
    """PositionalEncoder (positional_encoding.py:5-31): denominators are built in float32
    (torch.pow on a float32 arange, :11-13) whatever the compute dtype; even -> sin, odd -> cos;
    the d-wide table is tiled ``repeat`` times."""
    denom32 = torch.pow(torch.tensor(float(T), dtype=torch.float32),
                        2 * (torch.arange(0, d).float() // 2) / d)
    tab = batch_positions[:, :, None] / denom32.to(batch_positions.dtype)[None, None, :]
    even = torch.sin(tab[:, :, 0::2])
    odd = torch.cos(tab[:, :, 1::2])
    tab = torch.stack([even, odd], dim=-1).reshape(tab.shape)
    return tab.repeat(1, 1, repeat)


def ltae_tiny(down, batch_positions, pad_mask, p, cfg: OracleConfig) -> torch.Tensor:
    """LTAE2dtiny.forward + MultiHeadAttentionSmall + ScaledDotProductAttentionSmall
    (ltae.py:197-239, 341-385, 431-458).  down: [B,T,C,h,w] -> attn [n_head,B,T,h,w]."""
    pre = "temporal_encoder."
    b, t, c, h, w = down.shape
    nh, dk = cfg.n_head, cfg.d_k
    seq = down.permute(0, 3, 4, 2, 1).reshape(b * h * w, c, t)                 # [N, C, T]  (ltae.py:210-211)
    seq = group_norm(seq, nh, p[pre + "in_norm.weight"], p[pre + "in_norm.bias"], cfg.norm_eps)
    z = torch.einsum("nct,dc->ndt", seq, p[pre + "inconv.weight"][:, :, 0]) + p[pre + "inconv.bias"].reshape(1, -1, 1)
    if cfg.positional_encoding:
        pe = positional_table(batch_positions, cfg.d_model // nh, 1000.0, nh)   # [B,T,d_model] (ltae.py:181-184)
        pe = pe.reshape(b, 1, t, cfg.d_model).expand(b, h * w, t, cfg.d_model).reshape(b * h * w, t, cfg.d_model)
        z = z + pe.permute(0, 2, 1)
    key = torch.einsum("ndt,ed->net", z, p[pre + "attention_heads.fc1_k.weight"]) \
        + p[pre + "attention_heads.fc1_k.bias"].reshape(1, -1, 1)               # [N, nh*dk, T] (ltae.py:349)
    key = key.reshape(b * h * w, nh, dk, t)
    score = torch.einsum("nhdt,hd->nht", key, p[pre + "attention_heads.Q"]) / math.sqrt(dk)   # ltae.py:432-433
    if pad_mask is not None:
        pm = pad_mask.reshape(b, 1, 1, t).expand(b, h * w, 1, t).reshape(b * h * w, 1, t)
        score = score.masked_fill(pm, -1e3)                                     # ltae.py:435
    attn = torch.softmax(score, dim=-1)                                         # [N, nh, T]
    return attn.reshape(b, h, w, nh, t).permute(3, 0, 4, 1, 2)                  # [nh,B,T,h,w] (ltae.py:233-235)


def ltae_full(down, batch_positions, pad_mask, p, cfg: OracleConfig, training: bool, new_buffers=None, v_keep_mask=None,
              v_relu_mask=None):
    """LTAE2d.forward + MultiHeadAttention + ScaledDotProductAttention (ltae.py:96-141, 265-307, 400-416) as built by UNCRTAINTS for
    ``use_v`` (use_dropout=False: no dropout on the attention; mlp=[256,128]; dropout 0.2 on the MLP output; return_att=True).
    down: [B,T,C,h,w] -> (v [B,128,h,w], attn [n_head,B,T,h,w]).  ``v_keep_mask`` [B*h*w,128]: keep mask of the value dropout.
    ``v_relu_mask`` [B*h*w,128] (diagnostic, like ``pool_idx`` of `forward`): impose the active set of the MLP's ReLU.  The ReLU
    sits between a BatchNorm and a GroupNorm over groups of 8 channels; about 2 of the 262k pre-activations of a B=2 batch lie
    within fp32 rounding (1e-5) of zero, an fp32 implementation decides those differently from fp64, and where the rest of the
    group is zero the GroupNorm multiplies that element's gradient by up to 1/sqrt(eps) = 316: ONE such flip moves every
    temporal-encoder and encoder gradient by 1e-3..7e-3 (measured by shifting the threshold by +-1e-5 in fp64), although the
    forward output changes by < 1e-7.  Imposing the implementation's own mask separates this discrete effect from arithmetic error."""
    pre = "temporal_encoder."
    b, t, c, h, w = down.shape
    nh, dk = cfg.n_head, cfg.d_k
    n = b * h * w
    seq = down.permute(0, 3, 4, 2, 1).reshape(n, c, t)                                      # [N, C, T]  (ltae.py:102-103)
    seq = group_norm(seq, nh, p[pre + "in_norm.weight"], p[pre + "in_norm.bias"], cfg.norm_eps)
    z = torch.einsum("nct,dc->ndt", seq, p[pre + "inconv.weight"][:, :, 0]) + p[pre + "inconv.bias"].reshape(1, -1, 1)
    if cfg.positional_encoding:
        pe = positional_table(batch_positions, cfg.d_model // nh, 1000.0, nh)
        pe = pe.reshape(b, 1, t, cfg.d_model).expand(b, h * w, t, cfg.d_model).reshape(n, t, cfg.d_model)
        z = z + pe.permute(0, 2, 1)
    zt = z.permute(0, 2, 1)                                                                  # [N, T, d_model]
    key = zt @ p[pre + "attention_heads.fc1_k.weight"].t() + p[pre + "attention_heads.fc1_k.bias"]      # ltae.py:276
    key = key.reshape(n, t, nh, dk)
    score = torch.einsum("nthd,hd->nht", key, p[pre + "attention_heads.Q"]) / math.sqrt(dk)               # ltae.py:404-405
    if pad_mask is not None:
        pm = pad_mask.reshape(b, 1, 1, t).expand(b, h * w, 1, t).reshape(n, 1, t)
        score = score.masked_fill(pm, -1e3)                                                  # ltae.py:407
    attn = torch.softmax(score, dim=-1)                                                      # [N, nh, T]
    val = zt.reshape(n, t, nh, cfg.d_model // nh)                                            # heads = consecutive channel slices (ltae.py:286)
    out = torch.einsum("nht,nthe->nhe", attn, val).reshape(n, cfg.d_model)                   # concatenated heads (ltae.py:124-126)
    m = out @ p[pre + "mlp.0.weight"].t() + p[pre + "mlp.0.bias"]
    # nn.BatchNorm1d over the N rows (ltae.py:79)
    k = pre + "mlp.1."
    if training:
        mu = m.mean(dim=0)
        var = ((m - mu) ** 2).mean(dim=0)
        if new_buffers is not None:
            unbiased = var.detach() * (n / max(n - 1, 1))
            new_buffers[k + "running_mean"] = (1 - cfg.bn_momentum) * p[k + "running_mean"] + cfg.bn_momentum * mu.detach()
            new_buffers[k + "running_var"] = (1 - cfg.bn_momentum) * p[k + "running_var"] + cfg.bn_momentum * unbiased
    else:
        mu, var = p[k + "running_mean"], p[k + "running_var"]
    m = (m - mu) / torch.sqrt(var + cfg.norm_eps) * p[k + "weight"] + p[k + "bias"]
    m = torch.relu(m) if v_relu_mask is None else m * v_relu_mask.to(m.dtype)
    if training and cfg.v_dropout_p > 0:
        assert v_keep_mask is not None, "train-mode use_v oracle needs an explicit value-dropout keep mask"
        m = m * v_keep_mask.to(m.dtype) / (1.0 - cfg.v_dropout_p)
    v = group_norm(m, nh, p[pre + "out_norm.weight"], p[pre + "out_norm.bias"], cfg.norm_eps)            # ltae.py:131
    v = v.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    return v, attn.reshape(b, h, w, nh, t).permute(3, 0, 4, 1, 2)


def aggregate(x, attn, pad_mask, training, cfg: OracleConfig, keep_mask=None) -> torch.Tensor:
    """Compact_Temporal_Aggregator, mode 'att_group' (uncrtaints.py:156-210).
    x: [B,T,C,H,W], attn: [nh,B,T,h,w].  ``keep_mask`` [nh,B,T,H,W] (bool / 0-1) is the
    Bernoulli keep mask of nn.Dropout(0.1) for train mode (injected for reproducibility)."""
    nh, b, t, h, w = attn.shape
    H, W = x.shape[-2:]
    a = attn
    if H > w:                                                                   # uncrtaints.py:197-202
        a = bilinear_upsample(attn, H, W)
        if training and cfg.dropout_p > 0:      # dropout lives in the upsampling branch only
            assert keep_mask is not None, "train-mode oracle needs an explicit dropout keep mask"
            a = a * keep_mask.to(a.dtype) / (1.0 - cfg.dropout_p)
    if pad_mask is not None and bool(pad_mask.any()):
        a = a * (~pad_mask).to(a.dtype)[None, :, :, None, None]                 # uncrtaints.py:172
    c = x.shape[2]
    xg = x.reshape(b, t, nh, c // nh, H, W)
    out = (a.permute(1, 2, 0, 3, 4).unsqueeze(3) * xg).sum(dim=1)               # sum over T
    return out.reshape(b, c, H, W)


def head(o: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """Output nonlinearities (uncrtaints.py:436-446, 223-228, 384-385)."""
    o = o.unsqueeze(1)
    mean = o[:, :, :S2_BANDS]
    if cfg.out_nonlin_mean:
        mean = cfg.scale_by * torch.sigmoid(mean)
    if cfg.covar_dim == 0:
        return mean
    v = o[:, :, S2_BANDS:S2_BANDS + cfg.covar_dim]
    var = torch.where(v > 20.0, v, torch.log1p(torch.exp(torch.clamp(v, max=20.0)))) + cfg.var_eps
    return torch.cat([mean, var], dim=2)


def gather_pool(x: torch.Tensor, idx: torch.Tensor, out_hw: int = ATT_DOWN) -> torch.Tensor:
    """Max-pool *values* taken at given flat (h*W + w) positions: the pooling of `adaptive_max_pool` with the index selection
    imposed from outside.  Diagnostic only (``forward(..., pool_idx=...)``): an fp32 implementation whose encoder output differs
    from the fp64 oracle's in the 6th digit picks the other element of a near-tie in a handful of the 8x8 windows; the discrete
    re-routing of that window's (large, attention-path) gradient is an O(1) change of a few elements of dEnc, which shows up as
    5-8e-4 relative L2 on every encoder gradient although all arithmetic is accurate to ~5e-5.  Imposing the implementation's
    own argmax (itself checked bit-exact against ATen on identical input) separates the two effects."""
    n, c, h, w = x.shape
    return torch.gather(x.reshape(n, c, h * w), 2, idx.reshape(n, c, -1)).reshape(n, c, out_hw, out_hw)


def forward(p: Dict[str, torch.Tensor], x: torch.Tensor, batch_positions: Optional[torch.Tensor],
            cfg: OracleConfig, training: bool = True, keep_mask: Optional[torch.Tensor] = None,
            new_buffers: Optional[dict] = None, taps: Optional[dict] = None,
            pool_idx: Optional[torch.Tensor] = None, v_keep_mask: Optional[torch.Tensor] = None,
            v_relu_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """UNCRTAINTS.forward (uncrtaints.py:391-446).  x: [B,T,C_in,H,W] -> [B,1,13+covdim,H,W].
    ``pool_idx`` (int64 [B*T,128,32,32], diagnostic): impose the max-pool index selection, see `gather_pool`."""
    b, t, cin, H, W = x.shape
    pad_mask = (x == cfg.pad_value).all(dim=-1).all(dim=-1).all(dim=-1)          # :392-394
    f = x.reshape(b * t, cin, H, W)                                              # smart_forward, utae.py:422-450
    f = conv1x1(f, p["in_conv.conv.conv.0.weight"], p["in_conv.conv.conv.0.bias"])
    f = _norm(f, p, "in_conv.conv.conv.1.", cfg.encoder_norm, training, cfg, new_buffers)
    f = torch.relu(f)                                                            # utae.py:490-491
    if taps is not None:
        taps["pad_mask"], taps["in_conv"] = pad_mask, f
    for i in range(cfg.n_enc_blocks):
        f = _block(f, p, f"in_block.{i}.", cfg.encoder_norm, training, cfg, new_buffers, taps)
    c = f.shape[1]
    if cfg.is_mono:                                                              # :418 (needs T == 1)
        assert t == 1
    else:
        down, idx = adaptive_max_pool(f, ATT_DOWN)                               # :403-404
        if pool_idx is not None:
            down, idx = gather_pool(f, pool_idx, ATT_DOWN), pool_idx
        if cfg.use_v:
            v, attn = ltae_full(down.reshape(b, t, c, ATT_DOWN, ATT_DOWN), batch_positions, pad_mask, p, cfg, training, new_buffers,
                                v_keep_mask, v_relu_mask)
        else:
            attn = ltae_tiny(down.reshape(b, t, c, ATT_DOWN, ATT_DOWN), batch_positions, pad_mask, p, cfg)
        agg = aggregate(f.reshape(b, t, c, H, W), attn, pad_mask, training, cfg, keep_mask)
        if taps is not None:
            taps["pool_idx"], taps["down"], taps["attn"], taps["agg"] = idx, down, attn, agg
        f = agg
        if cfg.use_v:                                                            # :414-417
            up_v = bilinear_upsample(v, H, W)
            f = conv1x1(torch.cat([f, up_v], dim=1), p["include_v.weight"], p["include_v.bias"])
            if taps is not None:
                taps["v"], taps["mix"] = v, f
    for i in range(cfg.n_dec_blocks):
        f = _block(f, p, f"out_block.{i}.", cfg.decoder_norm, training, cfg, new_buffers, taps)
    if cfg.separate_out:                                                         # :376-379,424-430
        o = conv1x1(f, p["out_conv_mean_1.conv.conv.0.weight"], p["out_conv_mean_1.conv.conv.0.bias"])
        if cfg.covar_dim > 0:
            o = torch.cat([o, conv1x1(f, p["out_conv_var_1.conv.conv.0.weight"], p["out_conv_var_1.conv.conv.0.bias"])], dim=1)
    else:
        o = conv1x1(f, p["out_conv.conv.conv.0.weight"], p["out_conv.conv.conv.0.bias"])   # :381,432
    if new_buffers is not None and training:
        for k in p:
            if k.endswith("num_batches_tracked"):
                new_buffers[k] = p[k] + 1
    return head(o, cfg)


# --------------------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------------------
def mgnll(pred: torch.Tensor, target: torch.Tensor, var: torch.Tensor, mode: str = "diag",
          eps: float = 1e-8, reduction: str = "mean") -> torch.Tensor:
    """multi_gaussian_nll_loss / multi_diag_gaussian_nll (losses.py:131-145,149-218) in closed
    form.  The nested vmap hands every pixel a [B,1,13] slice, and ``var.log().sum()`` (:138)
    therefore sums the log-determinant over the *batch* as well as the channels; the
    Mahalanobis term stays per sample.  The clamp to eps (:203-205) is invisible to autograd.
    The additive constant is 13/2*log(2*float32(pi)) evaluated in float32 (:143)."""
    if reduction not in ("none", "mean", "sum"):
        raise ValueError(reduction + " is not valid")
    if mode == "iso":
        var = var.expand(-1, -1, S2_BANDS, -1, -1)                               # :190-192
    if torch.any(var < 0):
        raise ValueError("var has negative entry/entries")                       # :199-200
    v = var + (torch.clamp(var, min=eps) - var).detach()                         # identity gradient
    k = pred.shape[2]
    const = ((k / 2) * torch.log(2 * torch.tensor(torch.pi))).to(pred.dtype)    # product rounded in float32
    logdet = torch.log(v).sum(dim=(0, 1, 2))                                     # [H,W]: over B and C
    maha = ((pred - target) ** 2 / v).sum(dim=(1, 2))                            # [B,H,W]
    maha = torch.clamp(torch.nan_to_num(maha), min=1e-9)                         # :141
    loss = const + 0.5 * logdet.unsqueeze(0) + 0.5 * maha                        # [B,H,W]
    loss = loss.permute(1, 2, 0)                                                 # vmap output order [H,W,B]
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def gnll(pred: torch.Tensor, target: torch.Tensor, var: torch.Tensor, full: bool = True, eps: float = 1e-8,
         reduction: str = "mean"):
    """gaussian_nll_loss (losses.py:46-128), heteroscedastic case: 1/2 (log v + e^2 / v) [+ 1/2 log(2 pi)], v = var clamped to
    eps without a gradient of the clamp (:108-111); returns (loss, clamped var) like the reference (:122-128)."""
    import math
    if var.size() != pred.size():
        raise ValueError("var is of incorrect size")
    if reduction not in ("none", "mean", "sum"):
        raise ValueError(reduction + " is not valid")
    if torch.any(var < 0):
        raise ValueError("var has negative entry/entries")
    v = var + (torch.clamp(var, min=eps) - var).detach()
    loss = 0.5 * (torch.log(v) + (pred - target) ** 2 / v)
    if full:
        loss = loss + 0.5 * math.log(2 * math.pi)
    if reduction == "mean":
        return loss.mean(), v
    if reduction == "sum":
        return loss.sum(), v
    return loss, v


def covariance(var: torch.Tensor, mode: str = "diag", eps: float = 1e-8) -> torch.Tensor:
    """Second return value of the loss (losses.py:145,211): diag_embed of the clamped
    variance, [B,1,13,13,H,W]."""
    if mode == "iso":
        var = var.expand(-1, -1, S2_BANDS, -1, -1)
    v = torch.clamp(var, min=eps)
    return torch.diag_embed(v.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 5, 2, 3)


# --------------------------------------------------------------------------------------
# calibration statistics of the validation loop (BASELINE config #5)
# --------------------------------------------------------------------------------------
def errvar_samplewise(target: torch.Tensor, pred: torch.Tensor, var: torch.Tensor) -> Dict[str, float]:
    """Image-wise error / uncertainty statistics of img_metrics (model/src/learning/metrics.py:41-53): one sample
    = one image; `target`, `pred`, `var` are that sample's [1,13,H,W] tensors after BaseModel.rescale
    (base_model.py:103-112: means / scale_by, variances / scale_by^2)."""
    error = target - pred
    return {"error": float(error.nanmean()), "mean ae": float(error.abs().nanmean()),
            "mean se": float(error.square().nanmean()), "mean var": float(var.nanmean())}


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """pytorch_ssim.ssim (util/pytorch_ssim/__init__.py:7-37,72-80): per-channel 11x11 Gaussian (sigma 1.5, normalised in float32)
    moments with zero padding; mean of the SSIM map.  img: [N,C,H,W]."""
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    c = img1.shape[1]
    win = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(c, 1, window_size, window_size).contiguous().to(img1.dtype)
    pad = window_size // 2
    mu1, mu2 = F.conv2d(img1, win, padding=pad, groups=c), F.conv2d(img2, win, padding=pad, groups=c)
    s11 = F.conv2d(img1 * img1, win, padding=pad, groups=c) - mu1 * mu1
    s22 = F.conv2d(img2 * img2, win, padding=pad, groups=c) - mu2 * mu2
    s12 = F.conv2d(img1 * img2, win, padding=pad, groups=c) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))).mean()


def img_metrics(target: torch.Tensor, pred: torch.Tensor, var: Optional[torch.Tensor] = None) -> Dict[str, float]:
    """img_metrics (model/src/learning/metrics.py:20-57) for one [1,13,H,W] sample, without the pixel-wise maps."""
    rmse = torch.sqrt(torch.mean(torch.square(target - pred)))
    mat = torch.sum(target * pred, 1)
    mat = mat / torch.sqrt(torch.sum(target * target, 1)) / torch.sqrt(torch.sum(pred * pred, 1))
    sam = torch.mean(torch.acos(torch.clamp(mat, -1, 1)) * 180 / math.pi)
    out = {"RMSE": float(rmse), "MAE": float(torch.mean(torch.abs(target - pred))), "PSNR": float(20 * torch.log10(1 / rmse)),
           "SAM": float(sam), "SSIM": float(ssim(target, pred))}
    if var is not None:
        out.update(errvar_samplewise(target, pred, var))
    return out


def variance_from_covariance(cov: torch.Tensor) -> torch.Tensor:
    """[B,1,13,13,H,W] covariance -> [B,1,13,H,W] variances (train_reconstruct.py:321-324)."""
    return cov.diagonal(dim1=2, dim2=3).moveaxis(-1, 2)


def compute_uce_auce(var, errors, n_samples: int, percent: int = 5, l2: bool = True):
    """Uncertainty calibration errors of the validation loop (model/train_reconstruct.py:489-512, plotting omitted):
    the per-sample variances are grouped into 100/percent equal-width bins between their min and max
    (np.digitize against linspace(min, max, n_bins)[1:], :487); per bin the root-mean variance and the RMSE
    (l2) or mean std / mean |error| are compared; UCE weights the bin discrepancies by the bin population
    / n_samples, AUCE is their unweighted nan-mean."""
    import numpy as np
    n_bins = 100 // percent
    var, errors = torch.Tensor(var), torch.Tensor(errors)

    def metric(arg):
        return torch.sqrt(torch.mean(arg ** 2)) if l2 else torch.mean(torch.abs(arg))
    edges = np.linspace(float(var.min()), float(var.max()), num=n_bins)[1:]
    var_idx = torch.Tensor(np.digitize(var, bins=edges))
    bk_var, bk_err = torch.empty(n_bins), torch.empty(n_bins)
    for b in range(n_bins):
        bk_var[b] = metric(var[var_idx == b].sqrt())
        bk_err[b] = metric(errors[var_idx == b])
    calib = torch.abs(bk_err - bk_var)
    weight = torch.histogram(var_idx, n_bins)[0] / n_samples
    return float(torch.nansum(weight * calib)), float(torch.nanmean(calib))


# --------------------------------------------------------------------------------------
# synthetic inputs + init
# --------------------------------------------------------------------------------------
def synthetic_batch(b: int, t: int, h: int, w: int, cin: int = 15, scale_by: float = 10.0,
                    seed: int = 1234, pad_last: bool = False, dtype=torch.float32):
    """x = scale_by*U[0,1) [B,T,cin,H,W]; y likewise [B,1,13,H,W]; dates: sorted integer days
    in [1400,1900] (data/dataLoader.py:10,377-378).  ``pad_last`` zeroes the last frame of
    every sample (pad path, uncrtaints.py:157-178)."""
    g = torch.Generator("cpu").manual_seed(seed)
    x = scale_by * torch.rand(b, t, cin, h, w, generator=g, dtype=torch.float32)
    y = scale_by * torch.rand(b, 1, S2_BANDS, h, w, generator=g, dtype=torch.float32)
    dates = torch.sort(torch.randint(1400, 1901, (b, t), generator=g), dim=1).values.float()
    if pad_last:
        x[:, -1] = 0.0
    return x.to(dtype), y.to(dtype), dates.to(dtype)


def dropout_keep_mask(n_head: int, b: int, t: int, h: int, w: int, p: float = 0.1, seed: int = 4321):
    g = torch.Generator("cpu").manual_seed(seed)
    return torch.rand(n_head, b, t, h, w, generator=g) >= p


def init_params(cfg: OracleConfig, seed: int = 1) -> Dict[str, torch.Tensor]:
    """Parameter dict with the reference's keys/shapes and the distributions of
    weight_init (learning/weight_init.py:13-47) + module defaults (ltae.py:324-336).
    Not bit-identical to the reference's RNG stream; used where only *a* valid set of weights
    is needed (bench, smoke).  Parity tests load a reference state-dict instead."""
    g = torch.Generator("cpu").manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}

    def xavier(shape):
        rf = 1
        for s in shape[2:]:
            rf *= s
        fan_in, fan_out = shape[1] * rf, shape[0] * rf
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))

    def norm(key, c, kind):
        if kind == "batch":
            p[key + "weight"] = torch.randn(c, generator=g)
            p[key + "bias"] = torch.zeros(c)
            p[key + "running_mean"] = torch.zeros(c)
            p[key + "running_var"] = torch.ones(c)
            p[key + "num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
        else:
            p[key + "weight"] = torch.ones(c)
            p[key + "bias"] = torch.zeros(c)

    w, hid, se = 128, 256, 32
    p["in_conv.conv.conv.0.weight"] = xavier((w, cfg.input_dim, 1, 1))
    p["in_conv.conv.conv.0.bias"] = torch.randn(w, generator=g)
    norm("in_conv.conv.conv.1.", w, cfg.encoder_norm)

    def block(pre, kind):
        if cfg.block_type == "residual":
            for i in (1, 2, 3):
                p[f"{pre}conv{i}.conv.0.weight"] = xavier((w, w, 3, 3))
                p[f"{pre}conv{i}.conv.0.bias"] = torch.randn(w, generator=g)
                norm(f"{pre}conv{i}.conv.1.", w, kind)
            return
        norm(pre + "conv.norm.", w, kind)
        p[pre + "conv.fn.0.weight"] = xavier((hid, w, 1, 1))
        norm(pre + "conv.fn.1.", hid, kind)
        p[pre + "conv.fn.3.weight"] = xavier((hid, 1, 3, 3))
        norm(pre + "conv.fn.4.", hid, kind)
        p[pre + "conv.fn.6.fc.0.weight"] = xavier((se, hid))
        p[pre + "conv.fn.6.fc.2.weight"] = xavier((hid, se))
        p[pre + "conv.fn.7.weight"] = xavier((w, hid, 1, 1))
        norm(pre + "conv.fn.8.", w, kind)

    for i in range(cfg.n_enc_blocks):
        block(f"in_block.{i}.", cfg.encoder_norm)
    t = "temporal_encoder."
    if not cfg.is_mono:
        p[t + "inconv.weight"] = torch.randn(cfg.d_model, w, 1, generator=g)
        p[t + "inconv.bias"] = torch.randn(cfg.d_model, generator=g)
        p[t + "attention_heads.Q"] = torch.randn(cfg.n_head, cfg.d_k, generator=g) * math.sqrt(2.0 / cfg.d_k)
        p[t + "attention_heads.fc1_k.weight"] = xavier((cfg.n_head * cfg.d_k, cfg.d_model))
        p[t + "attention_heads.fc1_k.bias"] = torch.randn(cfg.n_head * cfg.d_k, generator=g)
        # GroupNorm affine parameters are left at (1, 0) by weight_init; perturbed here so that their gradients are exercised
        p[t + "in_norm.weight"] = torch.ones(w) + (0.2 * torch.randn(w, generator=g) if cfg.use_v else 0.0)
        p[t + "in_norm.bias"] = torch.zeros(w) + (0.2 * torch.randn(w, generator=g) if cfg.use_v else 0.0)
    if cfg.use_v:                                       # LTAE2d extras + include_v (ltae.py:68-85, uncrtaints.py:338)
        p[t + "out_norm.weight"] = torch.ones(w) + 0.2 * torch.randn(w, generator=g)
        p[t + "out_norm.bias"] = 0.2 * torch.randn(w, generator=g)
        p[t + "mlp.0.weight"] = xavier((w, cfg.d_model))
        p[t + "mlp.0.bias"] = torch.randn(w, generator=g)
        norm(t + "mlp.1.", w, "batch")
        p["include_v.weight"] = xavier((w, 2 * w, 1, 1))
        p["include_v.bias"] = torch.randn(w, generator=g)
    for i in range(cfg.n_dec_blocks):
        block(f"out_block.{i}.", cfg.decoder_norm)
    if cfg.separate_out:
        p["out_conv_mean_1.conv.conv.0.weight"] = xavier((S2_BANDS, w, 1, 1))
        p["out_conv_mean_1.conv.conv.0.bias"] = torch.randn(S2_BANDS, generator=g)
        if cfg.covar_dim > 0:
            p["out_conv_var_1.conv.conv.0.weight"] = xavier((cfg.covar_dim, w, 1, 1))
            p["out_conv_var_1.conv.conv.0.bias"] = torch.randn(cfg.covar_dim, generator=g)
    else:
        p["out_conv.conv.conv.0.weight"] = xavier((cfg.out_dim, w, 1, 1))
        p["out_conv.conv.conv.0.bias"] = torch.randn(cfg.out_dim, generator=g)
    return p


def value_keep_mask(b: int, h: int = ATT_DOWN, w: int = ATT_DOWN, c: int = 128, p: float = 0.2, seed: int = 987):
    """Bernoulli keep mask of LTAE2d.dropout on the [B*h*w, 128] MLP output (ltae.py:129)."""
    g = torch.Generator("cpu").manual_seed(seed)
    return torch.rand(b * h * w, c, generator=g) >= p


def step(p: Dict[str, torch.Tensor], x, y, dates, cfg: OracleConfig, training=True, keep_mask=None, loss_name: str = "MGNLL",
         pool_idx=None, v_keep_mask=None, v_relu_mask=None):
    """One fwd + loss (MGNLL, or GNLL for `--loss GNLL`) + bwd of the oracle.  Returns (out, loss, grads dict, new BN buffers)."""
    leaf = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in p.items()}
    new_buffers: dict = {}
    out = forward(leaf, x, dates, cfg, training, keep_mask, new_buffers, None, pool_idx, v_keep_mask, v_relu_mask)
    if loss_name == "GNLL":
        loss, _ = gnll(out[:, :, :S2_BANDS], y, out[:, :, S2_BANDS:S2_BANDS + cfg.covar_dim], full=True)
    else:
        loss = mgnll(out[:, :, :S2_BANDS], y, out[:, :, S2_BANDS:S2_BANDS + cfg.covar_dim], cfg.covmode)
    names = [k for k, v in leaf.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(leaf[k])) for k, g in zip(names, grads)}
    return out.detach(), loss.detach(), gd, new_buffers
