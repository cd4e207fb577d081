"""Size-independent properties checked at BASELINE's full size (config #2: B=16, T=3, 15x256x256) on the GPU, where the
CPU oracle would take minutes: determinism, sample independence (batch permutation equivariance in eval mode), attention
normalisation / padding, MGNLL closed-form identities, and gradient linearity in the upstream gradient."""
import pytest
import torch

from conftest import rel_l2
from test_gpu_parity import make_net, tap

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_batch():
    g = torch.Generator("cpu").manual_seed(1234)
    B, T, H, W = 16, 3, 256, 256
    x = 10.0 * torch.rand(B, T, 15, H, W, generator=g)
    y = 10.0 * torch.rand(B, 1, 13, H, W, generator=g)
    d = torch.sort(torch.randint(1400, 1901, (B, T), generator=g), dim=1).values.float()
    x[3, 2] = 0.0            # one padded frame (pad path, uncrtaints.py:157-178)
    return x.cuda(), y.cuda(), d.cuda()


def test_full_size_eval_determinism_and_permutation(golden_weights, full_batch):
    x, y, d = full_batch
    net = make_net(golden_weights, "diag").eval()
    with torch.no_grad():
        a = net(x, batch_positions=d).clone()
        b = net(x, batch_positions=d).clone()
        assert a.shape == (16, 1, 26, 256, 256)
        assert rel_l2(a, b) < 1e-6                               # fp64 atomics may reorder; values must agree to rounding
        attn = tap(net, "attn", (16, 16, 3, 32, 32))
        assert float((attn.sum(dim=2) - 1).abs().max()) < 1e-5   # softmax over T
        notpad = tap(net, "notpad", (16, 3), torch.int32)
        assert int((notpad == 0).sum()) == 1 and int(notpad[3, 2]) == 0
        perm = torch.randperm(16, generator=torch.Generator().manual_seed(0)).cuda()
        c = net(x[perm].contiguous(), batch_positions=d[perm].contiguous())
        assert rel_l2(c, a[perm]) < 1e-5                         # eval: GroupNorm per sample, BatchNorm running stats
    assert torch.isfinite(a).all()
    assert float(a[:, :, :13].min()) >= 0 and float(a[:, :, :13].max()) <= 10.0     # scale_by * sigmoid
    assert float(a[:, :, 13:].min()) >= 1e-3                                          # softplus + eps


def test_full_size_padded_frame_gets_zero_attention(golden_weights, full_batch):
    x, y, d = full_batch
    net = make_net(golden_weights, "diag").eval()
    with torch.no_grad():
        net(x, batch_positions=d)
        attn = tap(net, "attn", (16, 16, 3, 32, 32))
        assert float(attn[:, 3, 2].abs().max()) == 0.0           # masked_fill(-1e3) underflows to exactly 0 (ltae.py:435)
        agg = tap(net, "agg", (16, 256 * 256, 128)).clone()
        x2 = x.clone()
        x2[3, 1] = x2[3, 1] * 0.5 + 1.0                          # changing an unpadded frame changes the aggregate ...
        net(x2, batch_positions=d)
        agg2 = tap(net, "agg", (16, 256 * 256, 128))
        assert rel_l2(agg2[3], agg[3]) > 1e-3
        assert rel_l2(agg2[4], agg[4]) < 1e-6                    # ... of that sample only


def test_full_size_train_step_and_gradient_linearity(golden_weights, full_batch):
    import uncrtaints_b200 as ub
    x, y, d = full_batch
    net = make_net(golden_weights, "diag").train()
    net.temporal_aggregator.attn_dropout.p = 0.0                 # deterministic
    crit = ub.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode="diag", chunk=None, covariance="none")
    grads = []
    for scale in (1.0, 3.0):
        for p in net.parameters():
            p.grad = None
        for m in net.modules():                                  # same BatchNorm buffers for both runs
            if isinstance(m, torch.nn.BatchNorm2d):
                m.reset_running_stats()
        out = net(x, batch_positions=d)
        loss, _ = crit(out[:, :, :13], y, out[:, :, 13:26])
        (scale * loss).backward()
        grads.append({k: p.grad.clone() for k, p in net.named_parameters()})
    assert torch.isfinite(loss)
    for k in ("out_conv.conv.conv.0.weight", "out_block.4.conv.fn.7.weight", "out_block.0.conv.fn.3.weight",
              "in_block.0.conv.fn.0.weight", "temporal_encoder.attention_heads.Q", "in_conv.conv.conv.0.weight"):
        assert rel_l2(grads[1][k], 3.0 * grads[0][k]) < 1e-4, k
    # MGNLL closed-form identities (losses.py:131-145): pred == target and var == 1  =>  loss = 6.5*log(2*pi_f32) + 0.5*1e-9
    ones = torch.ones(16, 1, 13, 256, 256, device="cuda")
    l0, _ = crit(y, y, ones)
    assert abs(l0.item() - 11.946200370788574) < 1e-5
    # the log-determinant is summed over the batch (losses.py:138): var == e  =>  loss = const + 0.5*B*13 + 0.5*1e-9
    l1, _ = crit(y, y, ones * 2.718281828459045)
    assert abs(l1.item() - (11.946200370788574 + 0.5 * 16 * 13)) < 1e-3
