"""Drive the UNMODIFIED reference (baseline/_ref, see install_ref.py) through its own public API for bench.py's reference arm.

Nothing of this repo's product path is imported here: the model is the reference's ``src.backbones.uncrtaints.UNCRTAINTS``
built with the keyword set ``model_utils.get_generator`` passes (model/src/model_utils.py:86-108) and the CLI-effective values
of the README command (README.md:78; train_reconstruct.py:57-61), initialised by the reference's ``weight_init``
(model/src/learning/weight_init.py:13-47), and the loss is the reference's ``losses.MultiGaussianNLLLoss`` as ``get_loss``
builds it (model/src/losses.py:18-20).  One step = ``netG(x, batch_positions=dates)`` -> loss -> ``loss.backward()``, i.e.
BaseModel.forward / get_loss_G / backward_G (model/src/backbones/base_model.py:60-84) without the optimizer, which the metric
excludes.
"""
from __future__ import annotations

import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_MODEL_DIR = os.path.join(HERE, "_ref", "model")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_MODEL_DIR, "src", "backbones", "uncrtaints.py"))


def load():
    """(uncrtaints module, losses module, weight_init function) of the unmodified reference."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_ref.py` in the build container")
    if REF_MODEL_DIR not in sys.path:
        sys.path.insert(0, REF_MODEL_DIR)
    from src.backbones import uncrtaints          # noqa: E402  (the reference's package name is `src`)
    from src import losses                        # noqa: E402
    from src.learning.weight_init import weight_init   # noqa: E402
    return uncrtaints, losses, weight_init


def build(covmode: str = "diag", seed: int = 1, device: str = "cpu", **variant):
    """``variant``: constructor overrides (block_type='residual', use_v=True, is_mono=True, separate_out=True) for scripts/bench_variants.py."""
    uncrtaints, losses, weight_init = load()
    cov = {"diag": 13, "iso": 1, "uni": 13}[covmode]
    torch.manual_seed(seed)
    kw = dict(input_dim=15, encoder_widths=[128], decoder_widths=[128, 128, 128, 128, 128], out_conv=[13 + cov],
              out_nonlin_mean=True, out_nonlin_var="softplus", agg_mode="att_group", encoder_norm="group",
              decoder_norm="batch", n_head=16, d_model=256, d_k=4, pad_value=0, padding_mode="reflect",
              positional_encoding=True, covmode=covmode, scale_by=10.0, separate_out=False, use_v=False,
              block_type="mbconv", is_mono=False)
    kw.update(variant)
    net = uncrtaints.UNCRTAINTS(**kw)
    net.apply(weight_init)
    net = net.to(device).train()
    crit = losses.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode=covmode, chunk=None)
    return net, crit


def synthetic(b, t, hw, seed=1234, scale_by=10.0):
    g = torch.Generator("cpu").manual_seed(seed)
    x = scale_by * torch.rand(b, t, 15, hw, hw, generator=g)
    y = scale_by * torch.rand(b, 1, 13, hw, hw, generator=g)
    d = torch.sort(torch.randint(1400, 1901, (b, t), generator=g), dim=1).values.float()
    return x, y, d


def step(net, crit, x, y, d):
    for p in net.parameters():
        p.grad = None
    out = net(x, batch_positions=d)
    loss, _variance = crit(out[:, :, :net.mean_idx], y, out[:, :, net.mean_idx:net.vars_idx])
    loss.backward()
    return loss


def cpu_model_name() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_cpu(batch: int, t: int, hw: int, covmode: str, steps: int, warmup: int, budget_s: float = 240.0):
    """Reference on the host cores.  The step is B=`batch` samples (BASELINE config #1: 2); if `steps` of them would not fit
    the time budget the per-step sample shrinks to B=1 (said in `sample`)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net, crit = build(covmode, 1, "cpu")
    x, y, d = synthetic(batch, t, hw)
    t0 = time.perf_counter()
    step(net, crit, x, y, d)                    # first step: allocator / oneDNN primitive warm-up, never timed
    first = time.perf_counter() - t0
    done_warm = 1
    if batch > 1 and first * (steps + max(warmup - 1, 0)) > budget_s:
        batch = 1
        x, y, d = synthetic(batch, t, hw)
    for _ in range(max(warmup - done_warm, 0)):
        step(net, crit, x, y, d)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step(net, crit, x, y, d)
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": "samples/s", "cores": cores, "kind": "reference", "ms_per_step": 1e3 * dt / steps,
            "batch": batch, "last_loss": float(loss.detach()),
            "sample": f"{steps} step(s) of the unmodified reference (baseline/_ref: src.backbones.uncrtaints.UNCRTAINTS + "
                      f"src.losses.MultiGaussianNLLLoss), B={batch}, T={t}, 15x{hw}x{hw}, {covmode}, fp32, train mode, "
                      f"fwd+MGNLL+bwd, {cores} host threads ({cpu_model_name()}), torch {torch.__version__} CPU / oneDNN"}


def time_cuda(batch: int, t: int, hw: int, covmode: str, steps: int, warmup: int, tf32: bool, device="cuda:0", **variant):
    """Reference in PyTorch eager on the GPU (cuDNN / cuBLAS / ATen kernels): the "existing Blackwell kernel" the CUDA path has
    to beat.  Falls back to smaller batches on OOM.  Returns None if even B=1 does not fit."""
    dev = torch.device(device)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    net, crit = build(covmode, 1, dev, **variant)
    b = batch
    while b >= 1:
        oom, out = False, None
        try:
            x, y, d = (v.to(dev) for v in synthetic(b, t, hw))
            for _ in range(max(warmup, 1)):
                step(net, crit, x, y, d)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step(net, crit, x, y, d)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            peak = torch.cuda.max_memory_allocated(dev)
            out = {"value": b * steps / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms / steps, "batch": b, "tf32": bool(tf32),
                   "steps": steps, "last_loss": float(loss.detach()), "peak_mem_gb": round(peak / 2 ** 30, 1),
                   "what": "unmodified reference modules, torch eager on this GPU (cuDNN/cuBLAS/ATen), fwd + reference MGNLL "
                           "(nested vmap, covariance built and copied to the host inside the loss, losses.py:145) + bwd"}
        except torch.OutOfMemoryError:
            oom = True                          # clean up OUTSIDE the handler: the traceback pins the failed step's tensors
        x = y = d = loss = None
        for p in net.parameters():
            p.grad = None
        torch.cuda.empty_cache()
        if not oom:
            return out
        b //= 2
    return None
