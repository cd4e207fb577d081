"""Generate the golden fixtures from the UNMODIFIED reference at /root/reference (build container only).

    python tests/golden/make_golden.py

Writes (all float32 unless noted; the reference is evaluated in float64 on CPU, SURVEY.md §8c "gold"):
  state_dict_keys.json      key -> shape of the reference UNCRTAINTS state-dict (diag, 15 input channels)
  weights_seed1.npz         reference weights after torch.manual_seed(1); model.apply(weight_init), with the
                            BatchNorm running statistics perturbed so that eval mode is not trivially (0, 1)
  case_diag_train_pad.npz   B=2,T=3,64x64, covmode diag, train mode, injected dropout keep mask, last frame padded:
                            x, y, dates, keep (packed bits), out, loss, every parameter gradient, updated BN buffers,
                            pool argmax indices (int16) and pad mask (bit-exact items)
  case_iso_eval.npz         B=2,T=2,64x96, covmode iso, eval mode: x, y, dates, out, loss
  case_mgnll.npz            loss-only sweep (losses.py:149-218): var logits spanning the softplus threshold and the eps floor
The reference ships no tests or vectors of its own, hence this script.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import, uncrtaints_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def build(U, winit, covmode, dtype=torch.float64):
    cov = {"diag": 13, "iso": 1}[covmode]
    m = U.UNCRTAINTS(input_dim=15, encoder_widths=[128], decoder_widths=[128] * 5, out_conv=[13 + cov],
                     out_nonlin_mean=True, out_nonlin_var="softplus", covmode=covmode, scale_by=10.0)
    return m.to(dtype)


def main():
    U, Lm, winit = ref_import.load()
    torch.manual_seed(1)
    base = build(U, winit, "diag", torch.float32)
    base.apply(winit)
    g = torch.Generator("cpu").manual_seed(77)
    sd = {k: v.clone() for k, v in base.state_dict().items()}
    for k in sd:
        if k.endswith("running_mean"):
            sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g)
        if k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
    json.dump({k: list(v.shape) for k, v in sd.items()}, open(os.path.join(OUT, "state_dict_keys.json"), "w"), indent=0)
    np.savez(os.path.join(OUT, "weights_seed1.npz"), **{k: v.numpy() for k, v in sd.items()})

    def load_into(m, sd_, dtype):
        m.load_state_dict({k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd_.items()}, strict=False)

    # ---- case 1: diag, train, dropout mask injected, padded last frame ---------------------------------
    B, T, H, W = 2, 3, 64, 64
    x, y, d = O.synthetic_batch(B, T, H, W, pad_last=True)
    keep = O.dropout_keep_mask(16, B, T, H, W)
    m = build(U, winit, "diag")
    load_into(m, sd, torch.float64)
    m.train()
    m.temporal_aggregator.attn_dropout = ref_import.InjectedDropout(keep)
    taps = {}
    enc_block, enc_fwd = m.in_block[0], m.in_block[0].smart_forward

    def tapped(inp):            # instance-level wrapper; the reference calls layer.smart_forward (uncrtaints.py:400)
        o = enc_fwd(inp)
        taps["enc"] = o.detach()
        return o
    enc_block.smart_forward = tapped
    out = m(x.double(), batch_positions=d.double())
    crit = Lm.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode="diag", chunk=None)
    loss, _ = crit(out[:, :, :13], y.double(), out[:, :, 13:26])
    loss.backward()
    enc = taps["enc"]
    _, idx = torch.nn.functional.adaptive_max_pool2d(enc.reshape(B * T, 128, H, W), (32, 32), return_indices=True)
    pad_mask = (x == 0).all(dim=-1).all(dim=-1).all(dim=-1)
    case = {"x": x.numpy(), "y": y.numpy(), "dates": d.numpy(), "keep": np.packbits(keep.numpy()),
            "out": out.detach().float().numpy(), "loss": np.float64(loss.item()),
            "pool_idx": idx.to(torch.int16).numpy(), "pad_mask": pad_mask.numpy()}
    for k, p in m.named_parameters():
        case["grad." + k] = p.grad.float().numpy()
    for k, v in m.state_dict().items():
        if "running" in k or "num_batches" in k:
            case["buf." + k] = v.float().numpy() if v.is_floating_point() else v.numpy()
    np.savez(os.path.join(OUT, "case_diag_train_pad.npz"), **case)
    print("case_diag_train_pad: loss", loss.item())

    # ---- case 2: iso, eval ------------------------------------------------------------------------------
    B, T, H, W = 2, 2, 64, 96
    x, y, d = O.synthetic_batch(B, T, H, W, seed=99)
    m = build(U, winit, "iso")
    sd_iso = dict(sd)
    sd_iso["out_conv.conv.conv.0.weight"] = sd["out_conv.conv.conv.0.weight"][:14].clone()
    sd_iso["out_conv.conv.conv.0.bias"] = sd["out_conv.conv.conv.0.bias"][:14].clone()
    load_into(m, sd_iso, torch.float64)
    m.eval()
    with torch.no_grad():
        out = m(x.double(), batch_positions=d.double())
        crit = Lm.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode="iso", chunk=None)
        loss, _ = crit(out[:, :, :13], y.double(), out[:, :, 13:14])
    np.savez(os.path.join(OUT, "case_iso_eval.npz"), x=x.numpy(), y=y.numpy(), dates=d.numpy(),
             out=out.float().numpy(), loss=np.float64(loss.item()))
    print("case_iso_eval: loss", loss.item())

    # ---- case 3: loss-only sweep ------------------------------------------------------------------------
    g = torch.Generator("cpu").manual_seed(5)
    Bm, Hm, Wm = 3, 8, 16
    res = {}
    for mode, vc in (("diag", 13), ("iso", 1)):
        pred = (10 * torch.rand(Bm, 1, 13, Hm, Wm, generator=g)).double().requires_grad_(True)
        targ = (10 * torch.rand(Bm, 1, 13, Hm, Wm, generator=g)).double()
        logits = torch.linspace(-30, 30, Bm * vc * Hm * Wm).reshape(Bm, 1, vc, Hm, Wm)[..., torch.randperm(Wm, generator=g)]
        var = (torch.nn.functional.softplus(logits.double(), beta=1, threshold=20) + 1e-3)
        var[0, 0, 0, 0, :4] = 1e-12            # below the eps clamp (losses.py:203-205)
        var = var.requires_grad_(True)
        crit = Lm.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode=mode, chunk=None)
        loss, cov = crit(pred, targ, var)
        loss.backward()
        res.update({f"{mode}.pred": pred.detach().float().numpy(), f"{mode}.target": targ.float().numpy(),
                    f"{mode}.var": var.detach().float().numpy(), f"{mode}.loss": np.float64(loss.item()),
                    f"{mode}.dpred": pred.grad.float().numpy(), f"{mode}.dvar": var.grad.float().numpy(),
                    f"{mode}.cov_diag_sum": np.float64(cov.double().sum().item())})
        print("mgnll", mode, loss.item())
    np.savez(os.path.join(OUT, "case_mgnll.npz"), **res)


if __name__ == "__main__":
    main()
