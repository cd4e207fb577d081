"""world_size-2 gloo test of the data-parallel plumbing (runs on CPU): the flat-buffer all-reduce equals the mean of
the per-rank gradients, and the batch sharding is a partition."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uncrtaints_b200.parallel import FlatGradAllReduce, shard_batch
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.Linear(4, 2))
    data = torch.arange(4 * 8, dtype=torch.float32).reshape(4, 8) / 10
    sl = shard_batch(4, rank, world)
    bucket = FlatGradAllReduce(lin.parameters())
    bucket.zero_()
    lin(data[sl]).pow(2).mean().backward()
    local = bucket.flat.clone()
    bucket.all_reduce_mean()
    first = bucket.flat.clone()
    # the reference's own step calls optimizer.zero_grad() (base_model.py:120), which sets the grads to None in torch >= 2:
    # the next backward then creates fresh gradient tensors.  The bucket must fold them back in, not reduce stale data.
    torch.optim.SGD(lin.parameters(), lr=0.1).zero_grad()
    assert all(p.grad is None for p in lin.parameters())
    (3.0 * lin(data[sl]).pow(2).mean()).backward()
    bucket.all_reduce_mean()
    assert bucket.reattached == 4 and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    assert torch.allclose(bucket.flat, 3.0 * first, rtol=1e-5, atol=1e-7), "gradients created after zero_grad(set_to_none) were lost"
    q.put((rank, sl.start, sl.stop, local.tolist(), first.tolist(), [p.grad.data_ptr() for p in lin.parameters()],
           bucket.flat.data_ptr()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, l0, f0, ptrs0, base0), (r1, a1, b1, l1, f1, _, _) = res
    l0, f0, l1, f1 = (torch.tensor(t) for t in (l0, f0, l1, f1))
    assert (a0, b0, a1, b1) == (0, 2, 2, 4)
    assert torch.allclose(f0, f1) and torch.allclose(f0, (l0 + l1) / 2, atol=1e-7)
    assert ptrs0[0] == base0                      # param.grad aliases the flat buffer: one collective per step


def test_shard_batch_rejects_ragged():
    from uncrtaints_b200.parallel import shard_batch
    assert shard_batch(256, 3, 8) == slice(96, 128)
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
