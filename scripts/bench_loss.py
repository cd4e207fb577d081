#!/usr/bin/env python
"""BASELINE config #5: throughput of the MGNLL loss kernel alone, iso vs diag, at the full-size head output (B=16, 256x256).

    python scripts/bench_loss.py [--batch 16] [--reps 20]

Times ub200_mgnll_forward (loss + both gradients in one pass) and ub200_covariance (second return value) with CUDA events on
the launch stream and reports samples/s and achieved GB/s against the algorithmic bytes:
  diag: reads pred 13 + target 13 + var 13 planes, writes dpred 13 + dvar 13  -> 65 planes of B*P*4 bytes
  iso : reads 13 + 13 + 1, writes 13 + 1                                       -> 41 planes
Development / measurement tool (numbers quoted in DESIGN.md); not part of the product path.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import uncrtaints_b200 as ub  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    B, H, W = a.batch, a.hw, a.hw
    P = H * W
    g = torch.Generator("cpu").manual_seed(1)
    res = {}
    for mode, vc, planes in (("diag", 13, 65), ("iso", 1, 41)):
        out = torch.empty(B, 1, 13 + vc, H, W)
        out[:, :, :13] = 10 * torch.rand(B, 1, 13, H, W, generator=g)
        out[:, :, 13:] = torch.nn.functional.softplus(torch.randn(B, 1, vc, H, W, generator=g)) + 1e-3
        out = out.cuda().requires_grad_(True)
        y = (10 * torch.rand(B, 1, 13, H, W, generator=g)).cuda()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
        for cov, tag in (("none", "loss+grads"), ("dense", "loss+grads+covariance")):
            crit = ub.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode=mode, chunk=None, covariance=cov,
                                           check_negative=False)
            ms = []
            for r in range(a.reps + 3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                loss, _ = crit(out[:, :, :13], y, out[:, :, 13:13 + vc])
                e1.record()
                torch.cuda.synchronize()
                if r >= 3:
                    ms.append(e0.elapsed_time(e1))
            t = sorted(ms)[len(ms) // 2]
            nbytes = planes * B * P * 4 + (0 if cov == "none" else (vc + 169) * B * P * 4)
            res[f"{mode}.{tag}"] = {"ms": round(t, 4), "samples_per_s": round(B / t * 1e3, 1), "GBps": round(nbytes / t / 1e6, 1)}
    print(json.dumps({"workload": f"MGNLL loss kernel, B={B}, 13x{H}x{W}, L2 flushed between calls", **res}))


if __name__ == "__main__":
    main()
