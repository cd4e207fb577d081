"""Golden fixture for BASELINE config #5 (iso vs diag MGNLL head: calibration parity sweep), generated from the UNMODIFIED
reference at /root/reference (build container only):

    python tests/golden/make_calibration.py        ->  tests/golden/case_calibration.npz

S samples of [1,13,8,16] logits whose variance logits sweep [-30, 30] (softplus threshold 20 and the eps floor are both
crossed) go through the reference's own head nonlinearities (uncrtaints.py:223-228,384-385,436-446), MultiGaussianNLLLoss
(losses.py:346-354) in float64, BaseModel.rescale's scaling (base_model.py:103-112), img_metrics (metrics.py:20-53) per
sample as in the validation loop (train_reconstruct.py:316-329) and compute_uce_auce (train_reconstruct.py:489-530; the
function is extracted from the script's source with ast because the script itself needs torchnet / rasterio / matplotlib).
"""
import ast
import os
import sys
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
S, H, W, SCALE, EPS_VAR = 40, 8, 16, 10.0, 1e-3


def reference_uce_auce():
    src = open("/root/reference/model/train_reconstruct.py").read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "compute_uce_auce")
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "binarize")]
    glb = {"torch": torch, "np": np, "plt": mock.MagicMock(), "writer": mock.MagicMock()}
    glb["plt"].subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    exec(compile(ast.Module(body=keep, type_ignores=[]), "train_reconstruct.py", "exec"), glb)
    return glb["compute_uce_auce"]


def main():
    U, Lm, _ = ref_import.load()
    sys.path.insert(0, "/root/reference")
    cwd = os.getcwd()
    os.chdir("/root/reference/model")
    from src.learning import metrics as ref_metrics
    os.chdir(cwd)
    uce_fn = reference_uce_auce()
    g = torch.Generator("cpu").manual_seed(2024)
    lm = (1.5 * torch.randn(S, 13, H, W, generator=g)).float()
    y = (SCALE * torch.rand(S, 1, 13, H, W, generator=g)).float()
    level = torch.linspace(-30, 30, S)[torch.randperm(S, generator=g)]
    res = {"lm": lm.numpy(), "y": y.numpy()}
    for mode, vc in (("diag", 13), ("iso", 1)):
        lv = (level[:, None, None, None] + 0.5 * torch.randn(S, vc, H, W, generator=g)).float()
        a = lm.double().requires_grad_(True)
        b = lv.double().requires_grad_(True)
        mean = SCALE * torch.sigmoid(a)                                         # uncrtaints.py:384-385,441
        var = U.get_nonlinearity("softplus", EPS_VAR)(b)                         # uncrtaints.py:223-228,444
        out = torch.cat([mean, var], dim=1).unsqueeze(1)                         # [S,1,13+vc,H,W]
        crit = Lm.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode=mode, chunk=None)
        loss, variance = crit(out[:, :, :13], y.double(), out[:, :, 13:13 + vc])
        loss.backward()
        fake = (out[:, :, :13] / SCALE).detach()                                 # base_model.py:103-112
        real = y.double() / SCALE
        variance = variance.detach() / SCALE ** 2
        v5 = variance.diagonal(dim1=2, dim2=3).moveaxis(-1, 2)                   # train_reconstruct.py:321-324
        mvar, err = [], []
        for s in range(S):
            m = ref_metrics.img_metrics(real[s], fake[s], var=v5[s])
            mvar.append(m["mean var"]); err.append(m["error"])
        uce, auce = uce_fn(mvar, err, S, percent=5, l2=True, mode="val", step=0)
        res.update({f"{mode}.lv": lv.numpy(), f"{mode}.loss": np.float64(loss.item()), f"{mode}.dlm": a.grad.float().numpy(),
                    f"{mode}.dlv": b.grad.float().numpy(), f"{mode}.mean_var": np.array(mvar), f"{mode}.error": np.array(err),
                    f"{mode}.uce": np.float64(float(uce)), f"{mode}.auce": np.float64(float(auce))})
        print(mode, "loss", loss.item(), "uce", float(uce), "auce", float(auce))
    np.savez(os.path.join(OUT, "case_calibration.npz"), **res)


if __name__ == "__main__":
    main()
