"""Import the *unmodified* reference modules from /root/reference (build container only).

Test infrastructure: used by tests/test_oracle_vs_reference.py and tests/golden/make_golden.py
to pin the oracle and to generate golden vectors.  /root/reference does not exist on the GPU
box; ``available()`` is False there and every caller skips."""
import os
import sys

REF_MODEL_DIR = "/root/reference/model"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_MODEL_DIR, "src", "backbones"))


def load():
    """Returns (uncrtaints module, losses module, weight_init fn) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present")
    if REF_MODEL_DIR not in sys.path:
        sys.path.insert(0, REF_MODEL_DIR)
    from src.backbones import uncrtaints as ref_uncrtaints  # noqa: E402
    from src import losses as ref_losses                     # noqa: E402
    from src.learning.weight_init import weight_init         # noqa: E402
    return ref_uncrtaints, ref_losses, weight_init


import torch


class InjectedDropout(torch.nn.Module):
    """Stand-in for Compact_Temporal_Aggregator.attn_dropout (uncrtaints.py:154) that applies a
    given keep mask (laid out [n_head*B, T, H, W] like the tensor it is applied to, :198-202)
    so that train-mode runs of the reference are reproducible."""
    def __init__(self, keep_mask, p=0.1):
        super().__init__()
        self.keep, self.p = keep_mask, p

    def forward(self, a):
        return a * self.keep.reshape(a.shape).to(a.dtype) / (1.0 - self.p)
