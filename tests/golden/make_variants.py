"""Golden fixture for the constructor variants use_v / is_mono / separate_out, from the UNMODIFIED reference (build container only).

    python tests/golden/make_variants.py

Writes case_variants.npz.  Per variant V in VARIANTS (tests/variants.py): the reference UNCRTAINTS (fp64, CPU, train mode, injected
dropout masks) is run on O.synthetic_batch with the weights of O.init_params (both seeded, so neither is stored) and
"V.out" (float32), "V.loss" (float64) and "V.grad.<state-dict key>" (float32) are saved.  One decoder block keeps the file small.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_import, uncrtaints_oracle as O  # noqa: E402
from variants import VARIANTS, NO_FIXTURE_GRADS, variant_inputs, reference_model  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    U, Lm, _ = ref_import.load()
    case = {}
    for name, kw in VARIANTS.items():
        cfg, p, x, y, d, keep, vkeep = variant_inputs(kw)
        m = reference_model(U, cfg, p, keep, vkeep).double().train()
        out = m(x.double(), batch_positions=d.double())
        loss, _ = Lm.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode=cfg.covmode, chunk=None)(
            out[:, :, :13], y.double(), out[:, :, 13:13 + cfg.covar_dim])
        loss.backward()
        case[name + ".out"] = out.detach().float().numpy()
        case[name + ".loss"] = np.float64(loss.item())
        for k, prm in m.named_parameters():
            if name in NO_FIXTURE_GRADS:
                break
            case[name + ".grad." + k] = (prm.grad if prm.grad is not None else torch.zeros_like(prm)).float().numpy()
        print(name, "loss", loss.item(), "params", sum(q.numel() for q in m.parameters()))
    np.savez(os.path.join(OUT, "case_variants.npz"), **case)


if __name__ == "__main__":
    main()
