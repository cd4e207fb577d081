"""ctypes binding of libuncrtaints_b200.so (the C ABI in include/uncrtaints_b200.h).

There is no fallback: if the shared library is missing or a symbol is absent, importing the product
path fails loudly.  Build it with ``python -m uncrtaints_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuncrtaints_b200.so")

(UB200_P_IN_W, UB200_P_IN_B, UB200_P_IN_NORM_W, UB200_P_IN_NORM_B, UB200_P_IN_NORM_RM, UB200_P_IN_NORM_RV,
 UB200_P_LTAE_AP, UB200_P_LTAE_E, UB200_P_OUT_W, UB200_P_OUT_B,
 UB200_P_LTAE_GN_W, UB200_P_LTAE_GN_B, UB200_P_LTAE_WIN, UB200_P_LTAE_BIN, UB200_P_LTAE_PE, UB200_P_LTAE_MLP_W, UB200_P_LTAE_MLP_B,
 UB200_P_LTAE_BN_W, UB200_P_LTAE_BN_B, UB200_P_LTAE_BN_RM, UB200_P_LTAE_BN_RV, UB200_P_LTAE_ON_W, UB200_P_LTAE_ON_B,
 UB200_P_INCV_W, UB200_P_INCV_B, UB200_P_BLOCK0) = range(26)
(UB200_B_N0_W, UB200_B_N0_B, UB200_B_N0_RM, UB200_B_N0_RV, UB200_B_W1, UB200_B_N1_W, UB200_B_N1_B, UB200_B_N1_RM,
 UB200_B_N1_RV, UB200_B_WDW, UB200_B_N2_W, UB200_B_N2_B, UB200_B_N2_RM, UB200_B_N2_RV, UB200_B_F1, UB200_B_F2, UB200_B_W2,
 UB200_B_N3_W, UB200_B_N3_B, UB200_B_N3_RM, UB200_B_N3_RV, UB200_BLOCK_STRIDE) = range(22)

UB200_R_W, UB200_R_B, UB200_R_N_W, UB200_R_N_B, UB200_R_N_RM, UB200_R_N_RV, UB200_R_STRIDE = range(7)

ERRORS = {-1: "UB200_ERR_ARG (unsupported shape / configuration)", -2: "UB200_ERR_CUDA (kernel launch failed)",
          -3: "UB200_ERR_WORKSPACE (workspace too small)"}

# every symbol include/uncrtaints_b200.h declares
SYMBOLS = [
    "ub200_version", "ub200_launch_count", "ub200_prof_enable", "ub200_prof_num_kernels", "ub200_prof_kernel_name",
    "ub200_prof_read", "ub200_gemm1_forward", "ub200_num_param_slots", "ub200_workspace_bytes", "ub200_workspace_tap", "ub200_forward",
    "ub200_backward", "ub200_forward_v", "ub200_backward_v", "ub200_mgnll_forward", "ub200_gnll_forward", "ub200_mgnll_none", "ub200_gnll_none", "ub200_scale_by_scalar", "ub200_covariance", "ub200_mbconv_workspace_bytes",
    "ub200_mbconv_forward", "ub200_mbconv_backward", "ub200_head_forward", "ub200_head_backward", "ub200_adam_step", "ub200_img_metrics", "ub200_assemble_input",
]


class Desc(C.Structure):
    """struct ub200_desc"""
    _fields_ = [
        ("B", C.c_int), ("T", C.c_int), ("C_in", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("n_dec_blocks", C.c_int), ("out_dim", C.c_int), ("enc_groups", C.c_int), ("dec_groups", C.c_int),
        ("training", C.c_int), ("need_grad", C.c_int), ("mean_sigmoid", C.c_int), ("gemm_backend", C.c_int),
        ("scale_by", C.c_float), ("var_eps", C.c_float), ("pad_value", C.c_float), ("norm_eps", C.c_float),
        ("bn_momentum", C.c_float), ("dropout_p", C.c_float),
        ("seed", C.c_ulonglong), ("offset", C.c_ulonglong),
        ("block_type", C.c_int), ("use_v", C.c_int), ("v_dropout_p", C.c_float), ("is_mono", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU / PyTorch fallback). "
            "Build it with `python -m uncrtaints_b200.build`.")
    L = C.CDLL(LIB_PATH)
    for s in SYMBOLS:
        if not hasattr(L, s):
            raise RuntimeError(f"libuncrtaints_b200.so lacks symbol {s}")
    vp, sz, i, ll, f = C.c_void_p, C.c_size_t, C.c_int, C.c_longlong, C.c_float
    dp = C.POINTER(Desc)
    L.ub200_version.restype = i
    L.ub200_launch_count.restype = C.c_ulonglong
    L.ub200_prof_enable.argtypes = [C.c_ulonglong]
    L.ub200_prof_kernel_name.argtypes = [i]
    L.ub200_prof_kernel_name.restype = C.c_char_p
    L.ub200_prof_read.argtypes = [i, C.POINTER(C.c_double), C.POINTER(i)]
    L.ub200_gemm1_forward.argtypes = [i, vp, vp, vp, vp, vp, i, i, vp, vp]
    L.ub200_num_param_slots.argtypes = [dp]
    L.ub200_workspace_bytes.argtypes = [dp]
    L.ub200_workspace_bytes.restype = sz
    L.ub200_workspace_tap.argtypes = [dp, C.c_char_p, C.POINTER(sz), C.POINTER(sz)]
    L.ub200_forward.argtypes = [dp, vp, C.POINTER(vp), vp, vp, vp, sz, vp]
    L.ub200_backward.argtypes = [dp, vp, C.POINTER(vp), vp, vp, vp, C.POINTER(vp), vp, sz, vp]
    L.ub200_forward_v.argtypes = [dp, vp, C.POINTER(vp), vp, vp, vp, vp, sz, vp]
    L.ub200_backward_v.argtypes = [dp, vp, C.POINTER(vp), vp, vp, vp, vp, C.POINTER(vp), vp, sz, vp]
    L.ub200_mgnll_forward.argtypes = [vp, ll, vp, ll, vp, ll, i, i, i, f, vp, vp, vp, vp, vp, vp]
    L.ub200_gnll_forward.argtypes = [vp, ll, vp, ll, vp, ll, i, i, f, i, vp, vp, vp, vp, vp, vp, vp]
    L.ub200_mgnll_none.argtypes = [vp, ll, vp, ll, vp, ll, i, i, i, f, vp, vp, vp, vp, vp, vp]
    L.ub200_gnll_none.argtypes = [vp, ll, vp, ll, vp, ll, i, i, f, i, vp, vp, vp, vp, vp, vp, vp]
    L.ub200_scale_by_scalar.argtypes = [vp, vp, vp, sz, vp]
    L.ub200_covariance.argtypes = [vp, ll, i, i, i, f, vp, vp]
    L.ub200_mbconv_workspace_bytes.argtypes = [i, i, i]
    L.ub200_mbconv_workspace_bytes.restype = sz
    L.ub200_mbconv_forward.argtypes = [vp, C.POINTER(vp), i, i, i, i, i, f, f, i, vp, vp, sz, vp]
    L.ub200_mbconv_backward.argtypes = [vp, C.POINTER(vp), vp, C.POINTER(vp), i, i, i, i, i, i, vp, vp, sz, vp]
    L.ub200_adam_step.argtypes = [vp, vp, vp, vp, sz, i, f, f, f, f, f, f, i, vp]
    L.ub200_img_metrics.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp]
    L.ub200_assemble_input.argtypes = [vp, vp, i, i, i, i, i, vp]
    L.ub200_head_forward.argtypes = [vp, vp, vp, vp, i, i, i, f, i, f, vp]
    L.ub200_head_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, f, i, f, vp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {ERRORS.get(rc, rc)}")


def ptr_table(ptrs):
    """list of int device addresses (0 = NULL) -> ctypes void* array"""
    arr = (C.c_void_p * len(ptrs))()
    for k, p in enumerate(ptrs):
        arr[k] = p if p else None
    return arr
