#!/usr/bin/env python
"""Per-kernel timing of ONE standalone MBConv block (ub200_mbconv_forward/backward) at full resolution, for kernel variants.

    python scripts/bench_block.py [--n 16] [--hw 256] [--groups 0] [--modes 0,3,7] [--reps 3]

For every depthwise-kernel mode (ub200_dwconv_set_mode) it runs forward + backward, reports the CUDA-event time of every
kernel class (ub200_prof_*), and the relative L2 difference of out / dx / parameter gradients against the first mode.
Development tool: not part of the product path.
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from uncrtaints_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--groups", type=int, default=0)
    ap.add_argument("--modes", default="0,3,7")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--backend", type=int, default=3)
    a = ap.parse_args()
    L = _lib.lib()
    N, H, W = a.n, a.hw, a.hw
    g = torch.Generator("cpu").manual_seed(5)
    dev = {}
    B = _lib

    def rnd(*shape, scale=1.0, shift=0.0):
        return (torch.randn(*shape, generator=g) * scale + shift).cuda().contiguous()
    dev[B.UB200_B_W1] = rnd(256, 128, scale=0.08)
    dev[B.UB200_B_WDW] = rnd(256, 9, scale=0.3)
    dev[B.UB200_B_F1] = rnd(32, 256, scale=0.08)
    dev[B.UB200_B_F2] = rnd(256, 32, scale=0.08)
    dev[B.UB200_B_W2] = rnd(128, 256, scale=0.08)
    for wk, bk, rm, rv, C in ((B.UB200_B_N0_W, B.UB200_B_N0_B, B.UB200_B_N0_RM, B.UB200_B_N0_RV, 128),
                              (B.UB200_B_N1_W, B.UB200_B_N1_B, B.UB200_B_N1_RM, B.UB200_B_N1_RV, 256),
                              (B.UB200_B_N2_W, B.UB200_B_N2_B, B.UB200_B_N2_RM, B.UB200_B_N2_RV, 256),
                              (B.UB200_B_N3_W, B.UB200_B_N3_B, B.UB200_B_N3_RM, B.UB200_B_N3_RV, 128)):
        dev[wk] = rnd(C, scale=1.0)
        dev[bk] = rnd(C, scale=0.2)
        if not a.groups:
            dev[rm] = torch.zeros(C, device="cuda")
            dev[rv] = torch.ones(C, device="cuda")
    gkeys = [k for k in dev if k not in (B.UB200_B_N0_RM, B.UB200_B_N0_RV, B.UB200_B_N1_RM, B.UB200_B_N1_RV, B.UB200_B_N2_RM,
                                         B.UB200_B_N2_RV, B.UB200_B_N3_RM, B.UB200_B_N3_RV)]
    gdev = {k: torch.zeros_like(dev[k]) for k in gkeys}
    ptab = _lib.ptr_table([dev[k].data_ptr() if k in dev else 0 for k in range(B.UB200_BLOCK_STRIDE)])
    gtab = _lib.ptr_table([gdev[k].data_ptr() if k in gdev else 0 for k in range(B.UB200_BLOCK_STRIDE)])
    x = rnd(N, H * W, 128, scale=1.5, shift=0.3)
    dout = rnd(N, H * W, 128)
    out, dx = torch.empty_like(x), torch.empty_like(x)
    nbytes = L.ub200_mbconv_workspace_bytes(N, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    nk = L.ub200_prof_num_kernels()
    names = [L.ub200_prof_kernel_name(k).decode() for k in range(nk)]

    def run():
        for t in gdev.values():
            t.zero_()
        _lib.check(L.ub200_mbconv_forward(x.data_ptr(), ptab, N, H, W, a.groups, 1, 1e-5, 0.1, a.backend, out.data_ptr(),
                                          ws.data_ptr(), nbytes, st), "mbconv_forward")
        _lib.check(L.ub200_mbconv_backward(x.data_ptr(), ptab, dout.data_ptr(), gtab, N, H, W, a.groups, 1, a.backend,
                                           dx.data_ptr(), ws.data_ptr(), nbytes, st), "mbconv_backward")

    def rel(p, r):
        return float((p.double() - r.double()).norm() / (r.double().norm() + 1e-300))

    base = None
    for mode in [0]:
        run(); run()
        torch.cuda.synchronize()
        L.ub200_prof_enable((1 << nk) - 1)
        for _ in range(a.reps):
            run()
        torch.cuda.synchronize()
        line = []
        for k in range(nk):
            ms, n = ctypes.c_double(), ctypes.c_int()
            _lib.check(L.ub200_prof_read(k, ctypes.byref(ms), ctypes.byref(n)), "prof_read")
            if n.value:
                line.append(f"{names[k]}={ms.value / a.reps:.3f}")
        L.ub200_prof_enable(0)
        snap = {"out": out.clone(), "dx": dx.clone(), **{f"g{k}": v.clone() for k, v in gdev.items()}}
        if base is None:
            base = snap
            diff = ""
        else:
            diff = " | diff vs first: " + " ".join(f"{k}={rel(v, base[k]):.1e}" for k, v in snap.items())
        print(f"[block N={N} {H}x{W} groups={a.groups}] mode={mode} ms: " + " ".join(line) + diff, flush=True)


if __name__ == "__main__":
    main()
