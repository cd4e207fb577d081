"""Patch the B200 path into the reference's own package so its scripts run unchanged.

The reference scripts do ``from src.model_utils import get_model`` (model/train_reconstruct.py:25,
model/test_reconstruct.py:21); ``model_utils`` imports ``src.backbones.uncrtaints`` (model_utils.py:6) and
``base_model`` imports ``src.losses`` (base_model.py:4).  ``install()`` must therefore run after ``src`` is
importable (cwd = model/, README.md:77) and before ``main()`` builds the model.
"""
from __future__ import annotations

import importlib

import torch


def install(verbose: bool = True, metrics: bool = True) -> None:
    """metrics=True additionally patches ``src.learning.metrics.img_metrics`` (the per-sample image metrics of the validation loop,
    train_reconstruct.py:318-353) with the two-kernel GPU version; it only takes effect for scripts imported AFTER this call
    (``from src.learning.metrics import img_metrics`` binds the name at import, train_reconstruct.py:26) -- the launcher
    ``python -m uncrtaints_b200.run`` guarantees that order."""
    from . import backbone, losses
    ref_uncrtaints = importlib.import_module("src.backbones.uncrtaints")
    ref_losses = importlib.import_module("src.losses")
    ref_uncrtaints.UNCRTAINTS = backbone.UNCRTAINTS            # model_utils.get_generator looks it up at call time (:86)
    ref_losses.MultiGaussianNLLLoss = losses.MultiGaussianNLLLoss   # get_loss looks it up at call time (losses.py:19)
    ref_losses.GaussianNLLLoss = losses.GaussianNLLLoss             # --loss GNLL (losses.py:16)
    if metrics:
        try:
            ref_metrics = importlib.import_module("src.learning.metrics")
            from . import metrics as gpu_metrics
            if not hasattr(ref_metrics, "_reference_img_metrics"):
                ref_metrics._reference_img_metrics = ref_metrics.img_metrics

            def img_metrics(target, pred, var=None, pixelwise=True):
                if torch.is_tensor(target) and target.is_cuda and target.dim() == 4 and target.shape[1] == 13:
                    return gpu_metrics.img_metrics(target, pred, var, pixelwise)
                return ref_metrics._reference_img_metrics(target, pred, var, pixelwise)      # CPU tensors / other band counts
            ref_metrics.img_metrics = img_metrics
        except ImportError:                     # util.pytorch_ssim not importable from this cwd: leave the metrics alone
            pass
    if verbose:
        print("[uncrtaints_b200] installed UNCRTAINTS, MultiGaussianNLLLoss and GaussianNLLLoss into the reference's src package")
