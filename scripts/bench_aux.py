#!/usr/bin/env python
"""Measurements of the SURVEY §8f pieces around the hot path (one JSON line; quoted in DESIGN.md §6):
  eval   : forward-only validation pass (eval mode, no_grad; B=16, T=3, 256x256): this path's specialised forward vs the unmodified
           reference in torch eager on the same GPU (test_reconstruct.py:104-109)
  adam   : optimizer step + zero_grad over the 570 010 parameters: FusedAdam (one kernel) vs torch.optim.Adam (base_model.py:48-51,120-122)
  metrics: img_metrics of a 16-sample batch: two kernels + one read vs the reference's per-sample loop on GPU tensors
           (metrics.py:20-57, train_reconstruct.py:318-353)
  prepare: prepare_data_multi of a collated host batch (B=16, T=3): pinned staging + async copies vs the reference's function
Development / measurement tool; not part of the product path."""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import uncrtaints_b200 as ub  # noqa: E402
from baseline import ref_runner as R  # noqa: E402


def cuda_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wall_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def main():
    dev = torch.device("cuda:0")
    B, T, HW = 16, 3, 256
    res = {}
    x, y, d = R.synthetic(B, T, HW)
    xd, yd, dd = x.to(dev), y.to(dev), d.to(dev)
    # ---- eval forward -------------------------------------------------------------------------------------------------
    ref, _ = R.build("diag", 1, dev)
    ref.eval()
    net = ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="diag", scale_by=10.0)
    net.load_state_dict(ref.state_dict(), strict=True)
    net = net.to(dev).eval()
    with torch.no_grad():
        ours = cuda_ms(lambda: net(xd, batch_positions=dd))
        o = net(xd, batch_positions=dd)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        theirs = cuda_ms(lambda: ref(xd, batch_positions=dd), reps=3, warm=1)
        r = ref(xd, batch_positions=dd)
    res["eval_forward"] = {"ours_ms": round(ours, 3), "ours_samples_per_s": round(B / ours * 1e3, 1), "reference_eager_ms": round(theirs, 2),
                           "reference_eager_samples_per_s": round(B / theirs * 1e3, 1),
                           "out_rel_l2": float((o - r).double().norm() / r.double().norm())}
    del ref, r, o
    torch.cuda.empty_cache()
    # ---- optimizer step -----------------------------------------------------------------------------------------------
    nets = [ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="diag", scale_by=10.0).to(dev)
            for _ in range(2)]
    opt_t = torch.optim.Adam(nets[0].parameters(), lr=1e-3)
    opt_f = ub.FusedAdam(nets[1].parameters(), lr=1e-3)
    for p in nets[0].parameters():
        p.grad = torch.randn_like(p)
    opt_f.bucket.flat.normal_()

    def torch_step():
        opt_t.step()
        opt_t.zero_grad(set_to_none=False)
    res["adam_step"] = {"torch_adam_ms": round(wall_ms(torch_step, reps=20), 4), "fused_adam_ms": round(wall_ms(lambda: opt_f.step(), reps=20), 4),
                        "params": int(opt_f.flat_params.numel()), "note": "wall clock incl. launch overhead (the GPU is idle otherwise)"}
    # ---- image metrics ------------------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(0)
    t_, p_, v_ = (torch.rand(B, 1, 13, HW, HW, generator=g).to(dev) for _ in range(3))
    ours_m = wall_ms(lambda: ub.img_metrics_batch(t_, p_, v_))
    R.load()
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    cwd = os.getcwd()
    os.chdir(R.REF_MODEL_DIR)
    from src.learning.metrics import img_metrics as ref_metrics
    os.chdir(cwd)
    theirs_m = wall_ms(lambda: [ref_metrics(t_[b], p_[b], var=v_[b]) for b in range(B)], reps=2, warm=1)
    res["img_metrics_batch16"] = {"ours_ms": round(ours_m, 3), "reference_loop_ms": round(theirs_m, 2)}
    # ---- prepare_data_multi -------------------------------------------------------------------------------------------
    batch = {"input": {"S1": [torch.rand(B, 2, HW, HW, generator=g) for _ in range(T)], "S2": [torch.rand(B, 13, HW, HW, generator=g) for _ in range(T)],
                       "masks": [torch.rand(B, HW, HW, generator=g) for _ in range(T)],
                       "S1 TD": [torch.randint(1400, 1900, (B,), generator=g) for _ in range(T)],
                       "S2 TD": [torch.randint(1400, 1900, (B,), generator=g) for _ in range(T)]},
             "target": {"S2": [torch.rand(B, 13, HW, HW, generator=g)]}}
    cfg = types.SimpleNamespace(use_sar=True, batch_size=B)

    def ref_prepare():      # train_reconstruct.py:161-179, restated inline (importing the script parses sys.argv)
        dev_ = "cuda:0"
        in_S2 = [t.to(dev_) for t in batch["input"]["S2"]]
        in_S1 = [t.to(dev_) for t in batch["input"]["S1"]]
        in_m = torch.stack([t.to(dev_) for t in batch["input"]["masks"]]).swapaxes(0, 1)
        yy = torch.cat([t.to(dev_) for t in batch["target"]["S2"]], dim=0).unsqueeze(1)
        xx = torch.cat((torch.stack(in_S1, dim=1), torch.stack(in_S2, dim=1)), dim=2)
        return xx, yy, in_m
    res["prepare_data_multi_b16"] = {"ours_ms": round(wall_ms(lambda: ub.prepare_data_multi(batch, dev, cfg)), 2),
                                     "reference_ms": round(wall_ms(ref_prepare), 2)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
