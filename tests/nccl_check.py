"""Worker of tests/test_gpu_multi.py (one process per GPU under torchrun): data-parallel parity of SURVEY.md §8e.
Rank r runs forward + MGNLL + backward on samples [r*B, (r+1)*B) of a global batch through the CUDA path, the flat gradient is
reduced with ONE NCCL all-reduce (uncrtaints_b200.FlatGradAllReduce); every rank writes its loss, its local flat gradient and the
reduced one to `out_dir` for the checker."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_dir, B, T, HW = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import uncrtaints_b200 as ub
    from oracle import uncrtaints_oracle as O
    w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(ROOT, "tests", "golden", "weights_seed1.npz")).items()}
    net = ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="diag", scale_by=10.0)
    net.load_state_dict(w, strict=True)
    net = net.to(dev).train()
    x, y, d = O.synthetic_batch(B * world, T, HW, HW, seed=500)
    keep = O.dropout_keep_mask(16, B * world, T, HW, HW, seed=501)
    sl = ub.shard_batch(B * world, rank, world)
    net._injected_keep_mask = keep[:, sl].contiguous().to(torch.uint8)
    crit = ub.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode="diag", chunk=None, covariance="none")
    bucket = ub.FlatGradAllReduce(net.parameters())
    bucket.zero_()
    out = net(x[sl].to(dev), batch_positions=d[sl].to(dev))
    loss, _ = crit(out[:, :, :13], y[sl].to(dev), out[:, :, 13:26])
    loss.backward()
    local_flat = bucket.flat.clone()
    bucket.all_reduce_mean()
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=float(loss), local=local_flat.cpu().numpy(), reduced=bucket.flat.cpu().numpy(),
             names=np.array([k for k, p in net.named_parameters() if p.requires_grad]),
             sizes=np.array([p.numel() for k, p in net.named_parameters() if p.requires_grad]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
