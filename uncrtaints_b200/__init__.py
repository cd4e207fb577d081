"""uncrtaints_b200 -- B200 (sm_100a) implementation of the UnCRtainTS forward/backward hot path.

Public surface (mirrors the reference's):
  UNCRTAINTS(...)                      <- model/src/backbones/uncrtaints.py:230
  MultiGaussianNLLLoss, get_loss, calc_loss   <- model/src/losses.py:14,35,288
  install()                            patch the above into the reference's ``src`` package
  FlatGradAllReduce                    one-collective data-parallel gradient reduction (new; reference is single-GPU)
The compute lives in libuncrtaints_b200.so (C ABI: include/uncrtaints_b200.h); there is no CPU fallback.
"""
from .backbone import UNCRTAINTS, set_default_gemm_backend  # noqa: F401
from .losses import (MultiGaussianNLLLoss, GaussianNLLLoss, get_loss, calc_loss, multi_gaussian_nll_loss,  # noqa: F401
                     gaussian_nll_loss, covariance_diag)
from .install import install  # noqa: F401
from .parallel import FlatGradAllReduce, HostToDevicePrefetcher, HostScalarReader, shard_batch  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .metrics import img_metrics, img_metrics_batch  # noqa: F401
from .data import prepare_data_multi  # noqa: F401

__all__ = ["UNCRTAINTS", "MultiGaussianNLLLoss", "GaussianNLLLoss", "gaussian_nll_loss", "get_loss", "calc_loss", "multi_gaussian_nll_loss", "covariance_diag",
           "install", "prepare_data_multi", "img_metrics", "img_metrics_batch", "FusedAdam", "FlatGradAllReduce", "HostToDevicePrefetcher", "HostScalarReader", "shard_batch", "set_default_gemm_backend"]
