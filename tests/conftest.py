import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REPORT_DIR = os.path.join(ROOT, "gpurun_out")

# Gradients that are analytically zero in the reference (SURVEY.md §7 "Analytically-zero gradients"): softmax over T
# is shift invariant, and a per-channel constant through a bias-free 1x1 conv is removed by the following BatchNorm.
ZERO_GRAD_SUFFIXES = ("temporal_encoder.in_norm.bias", "temporal_encoder.inconv.bias",
                      "temporal_encoder.attention_heads.fc1_k.bias")


def is_zero_grad_param(name: str, decoder_norm: str = "batch", training: bool = True) -> bool:
    if name in ZERO_GRAD_SUFFIXES:
        return True
    # only a BatchNorm that normalises with *batch* statistics (training mode) removes the per-channel constant
    return training and decoder_norm == "batch" and name.startswith("out_block.") and name.endswith("conv.norm.bias")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def golden_weights():
    return {k: torch.from_numpy(v) for k, v in load_npz("weights_seed1.npz").items()}


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def report(name: str, lines):
    os.makedirs(REPORT_DIR, exist_ok=True)
    with open(os.path.join(REPORT_DIR, name), "a") as f:
        for ln in lines:
            f.write(ln + "\n")
