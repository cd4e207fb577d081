"""Golden fixture of the GNLL loss (`--loss GNLL`, covmode 'uni') from the UNMODIFIED reference (build container only):

    python tests/golden/make_gnll.py        ->  tests/golden/case_gnll.npz

gaussian_nll_loss / GaussianNLLLoss (model/src/losses.py:46-128,222-284) in float64 on a [3,1,13,8,16] problem whose variances
span the softplus range and fall below the eps clamp in a few places."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def main():
    _, Lm, _ = ref_import.load()
    g = torch.Generator("cpu").manual_seed(11)
    B, H, W = 3, 8, 16
    pred = (10 * torch.rand(B, 1, 13, H, W, generator=g)).float()
    targ = (10 * torch.rand(B, 1, 13, H, W, generator=g)).float()
    logits = torch.linspace(-30, 30, B * 13 * H * W).reshape(B, 1, 13, H, W)[..., torch.randperm(W, generator=g)]
    var = (torch.nn.functional.softplus(logits, beta=1, threshold=20) + 1e-3).float()
    var[0, 0, 0, 0, :4] = 1e-12
    a, v = pred.double().requires_grad_(True), var.double().requires_grad_(True)
    crit = Lm.GaussianNLLLoss(reduction="mean", eps=1e-8, full=True)
    loss, vout = crit(a, targ.double(), v)
    loss.backward()
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "case_gnll.npz"), pred=pred.numpy(), target=targ.numpy(),
             var=var.numpy(), loss=np.float64(loss.item()), dpred=a.grad.float().numpy(), dvar=v.grad.float().numpy(),
             var_out=vout.detach().float().numpy())
    print("gnll loss", loss.item())


if __name__ == "__main__":
    main()
