"""Throughput of the constructor variants (block_type='residual', use_v, is_mono, separate_out) beside the unmodified reference in
torch eager on the same GPU:  python scripts/bench_variants.py [--batch 16] [--steps 5] > gpurun_out/bench_variants.log

One JSON line per variant: samples/s of fwd + MGNLL + bwd (train mode, synthetic 15x256x256 input, inputs resident in HBM, CUDA-event
timed after 3 warm-up steps) for this library and for the reference modules (TF32 off and on; batch halves on OOM)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BREAKDOWN = False
VARIANTS = {"mbconv (default)": {}, "residual": dict(block_type="residual"), "use_v": dict(use_v=True),
            "is_mono": dict(is_mono=True), "separate_out": dict(separate_out=True), "residual+use_v": dict(block_type="residual", use_v=True)}


def ours(variant, B, T, hw, steps, warmup=3):
    import uncrtaints_b200 as ub
    from bench import init_like_reference, synthetic
    dev = torch.device("cuda:0")
    net = ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="diag", scale_by=10.0, **variant)
    init_like_reference(net, seed=1)
    net = net.to(dev).train()
    crit = ub.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode="diag", chunk=None, check_negative="deferred")
    bucket = ub.FlatGradAllReduce(net.parameters())
    x, y, d = (v.to(dev) for v in synthetic(B, T, hw))

    def step():
        bucket.zero_()
        out = net(x, batch_positions=d)
        loss, _ = crit(out[:, :, :13], y, out[:, :, 13:26])
        loss.backward()
        return loss
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    breakdown = {}
    if BREAKDOWN:                       # per-kernel-class device time of one step (ub200_prof_*)
        import ctypes
        from uncrtaints_b200 import _lib
        L = _lib.lib()
        nk = L.ub200_prof_num_kernels()
        L.ub200_prof_enable((1 << nk) - 1)
        step()
        torch.cuda.synchronize()
        for k in range(nk):
            ms_k, n_k = ctypes.c_double(), ctypes.c_int()
            _lib.check(L.ub200_prof_read(k, ctypes.byref(ms_k), ctypes.byref(n_k)), "prof_read")
            if n_k.value:
                breakdown[L.ub200_prof_kernel_name(k).decode()] = [round(ms_k.value, 2), n_k.value]
        L.ub200_prof_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = {"samples_per_s": round(B * 1e3 / ms, 2), "ms_per_step": round(ms, 2), "batch": B, "loss": float(loss),
           "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    if breakdown:
        res["kernel_ms_launches"] = breakdown
    del net, bucket, x, y, d
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--t", type=int, default=3)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--breakdown", action="store_true")
    a = ap.parse_args()
    global BREAKDOWN
    BREAKDOWN = a.breakdown
    from baseline import ref_runner as R
    for name, var in VARIANTS.items():
        if a.only and name not in a.only.split(","):
            continue
        T = 1 if var.get("is_mono") else a.t
        line = {"variant": name, "T": T, "hw": a.hw, "ours": ours(var, a.batch, T, a.hw, a.steps)}
        if not a.no_reference and R.available():
            for tf32 in (0, 1):
                r = R.time_cuda(a.batch, T, a.hw, "diag", max(2, a.steps // 2), 2, bool(tf32), **var)
                line["reference_eager_tf32" if tf32 else "reference_eager_fp32"] = r and {k: r[k] for k in ("value", "ms_per_step", "batch", "peak_mem_gb")}
            if line.get("reference_eager_fp32"):
                line["speedup_vs_eager_fp32"] = round(line["ours"]["samples_per_s"] / line["reference_eager_fp32"]["value"], 2)
                line["speedup_vs_eager_tf32"] = round(line["ours"]["samples_per_s"] / line["reference_eager_tf32"]["value"], 2)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
