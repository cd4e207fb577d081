"""Multi-GPU parity on hardware (BASELINE config #4's mechanism): needs >= 2 GPUs (`gpurun --gpus 2`; skipped on
the single-GPU box).  Per rank: loss and gradients equal the fp64 oracle run on that rank's shard (decoder BatchNorm statistics
and the MGNLL batch-summed log-determinant are per rank by definition); the NCCL-reduced flat gradient equals the mean of the
per-rank gradients on every rank."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import rel_l2, is_zero_grad_param
from oracle import uncrtaints_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_data_parallel_parity(tmp_path, golden_weights):
    world, B, T, HW = 2, 2, 3, 64
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_check.py"), str(tmp_path), str(B), str(T), str(HW)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    ranks = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world)]
    x, y, d = O.synthetic_batch(B * world, T, HW, HW, seed=500)
    keep = O.dropout_keep_mask(16, B * world, T, HW, HW, seed=501)
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in golden_weights.items()}
    cfg = O.OracleConfig()
    lines = []
    for r, rec in enumerate(ranks):
        sl = slice(r * B, (r + 1) * B)
        _, o_loss, o_grads, _ = O.step(p64, x[sl].double(), y[sl].double(), d[sl].double(), cfg, True, keep[:, sl])
        assert abs(float(rec["loss"]) - float(o_loss)) <= 1e-3 * abs(float(o_loss))
        off = 0
        for name, n in zip(rec["names"], rec["sizes"]):
            g = torch.from_numpy(rec["local"][off:off + n])
            off += int(n)
            if is_zero_grad_param(str(name)):
                continue
            e = rel_l2(g, o_grads[str(name)].reshape(-1))
            assert e <= 1e-3, (r, str(name), e)
        lines.append(f"rank {r}: loss {float(rec['loss']):.6f} oracle {float(o_loss):.6f}")
    mean = sum(torch.from_numpy(rec["local"]).double() for rec in ranks) / world
    for rec in ranks:
        assert rel_l2(torch.from_numpy(rec["reduced"]), mean) <= 1e-6
    assert np.array_equal(ranks[0]["reduced"], ranks[1]["reduced"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "nccl_parity.txt"), "w") as f:
        f.write("\n".join(lines + [f"reduced == mean of per-rank gradients on {world} ranks (rel_l2 <= 1e-6), identical across ranks"]) + "\n")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_model_on_a_non_current_device(golden_weights):
    """the C ABI launches on the current device; the shim must switch to the tensors' device, and the per-device
    'dynamic shared memory attribute set' caches must not leak from device 0 to device 1 (one process driving two GPUs)."""
    import uncrtaints_b200 as ub
    x, y, d = O.synthetic_batch(1, 2, 64, 64, seed=71)
    keep = O.dropout_keep_mask(16, 1, 2, 64, 64, seed=72).to(torch.uint8)
    outs, grads = [], []
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)                      # the current device stays 0 throughout
        net = ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="diag", scale_by=10.0)
        net.load_state_dict({k: v.clone() for k, v in golden_weights.items()}, strict=True)
        net = net.to(dev).train()
        net._injected_keep_mask = keep
        out = net(x.to(dev), batch_positions=d.to(dev))
        loss, cov = ub.MultiGaussianNLLLoss(mode="diag", chunk=None)(out[:, :, :13], y.to(dev), out[:, :, 13:26])
        loss.backward()
        torch.cuda.synchronize(dev)
        assert out.device == torch.device(dev) and cov.device == torch.device(dev)
        outs.append(out.detach().cpu())
        grads.append(net.in_block[0].conv.fn[0].weight.grad.cpu())
    assert rel_l2(outs[1], outs[0]) < 1e-5 and rel_l2(grads[1], grads[0]) < 1e-4
