"""Patch the B200 path into the reference's own package so its scripts run unchanged.

The reference scripts do ``from src.model_utils import get_model`` (model/train_reconstruct.py:25,
model/test_reconstruct.py:21); ``model_utils`` imports ``src.backbones.uncrtaints`` (model_utils.py:6) and
``base_model`` imports ``src.losses`` (base_model.py:4).  ``install()`` must therefore run after ``src`` is
importable (cwd = model/, README.md:77) and before ``main()`` builds the model.
"""
from __future__ import annotations

import importlib


def install(verbose: bool = True) -> None:
    from . import backbone, losses
    ref_uncrtaints = importlib.import_module("src.backbones.uncrtaints")
    ref_losses = importlib.import_module("src.losses")
    ref_uncrtaints.UNCRTAINTS = backbone.UNCRTAINTS            # model_utils.get_generator looks it up at call time (:86)
    ref_losses.MultiGaussianNLLLoss = losses.MultiGaussianNLLLoss   # get_loss looks it up at call time (losses.py:19)
    ref_losses.GaussianNLLLoss = losses.GaussianNLLLoss             # --loss GNLL (losses.py:16)
    if verbose:
        print("[uncrtaints_b200] installed UNCRTAINTS, MultiGaussianNLLLoss and GaussianNLLLoss into the reference's src package")
