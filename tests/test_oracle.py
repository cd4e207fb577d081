"""CPU tests of the oracle: against the committed golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py) and, where /root/reference exists, against the live reference."""
import numpy as np
import pytest
import torch

from conftest import load_npz, rel_l2, is_zero_grad_param
from oracle import uncrtaints_oracle as O, ref_import


def _sd64(w):
    return {k: (v.double() if v.is_floating_point() else v) for k, v in w.items()}


def test_oracle_matches_golden_train_pad(golden_weights):
    c = load_npz("case_diag_train_pad.npz")
    x, y, d = (torch.from_numpy(c[k]).double() for k in ("x", "y", "dates"))
    B, T, _, H, W = x.shape
    keep = torch.from_numpy(np.unpackbits(c["keep"])[:16 * B * T * H * W].reshape(16, B, T, H, W)).bool()
    cfg = O.OracleConfig()
    out, loss, grads, newbuf = O.step(_sd64(golden_weights), x, y, d, cfg, True, keep)
    assert rel_l2(out, torch.from_numpy(c["out"])) < 1e-6
    assert abs(loss.item() - float(c["loss"])) / abs(float(c["loss"])) < 1e-9
    for k, g in grads.items():
        ref = torch.from_numpy(c["grad." + k])
        if is_zero_grad_param(k):
            assert g.norm() < 1e-6 * max(1.0, float(ref.norm()) * 1e6)
        else:
            assert rel_l2(g, ref) < 1e-5, k
    for k, v in newbuf.items():
        assert torch.allclose(v.double(), torch.from_numpy(c["buf." + k]).double(), rtol=1e-5, atol=1e-6), k
    # bit-exact items
    taps = {}
    O.forward(_sd64(golden_weights), x, d, cfg, True, keep, None, taps)
    assert torch.equal(taps["pad_mask"], torch.from_numpy(c["pad_mask"]))
    assert torch.equal(taps["pool_idx"].to(torch.int16), torch.from_numpy(c["pool_idx"]))


def test_oracle_matches_golden_iso_eval(golden_weights):
    c = load_npz("case_iso_eval.npz")
    x, y, d = (torch.from_numpy(c[k]).double() for k in ("x", "y", "dates"))
    cfg = O.OracleConfig(covmode="iso")
    sd = _sd64(golden_weights)
    sd["out_conv.conv.conv.0.weight"] = sd["out_conv.conv.conv.0.weight"][:14]
    sd["out_conv.conv.conv.0.bias"] = sd["out_conv.conv.conv.0.bias"][:14]
    out = O.forward(sd, x, d, cfg, training=False)
    loss = O.mgnll(out[:, :, :13], y, out[:, :, 13:14], "iso")
    assert rel_l2(out, torch.from_numpy(c["out"])) < 1e-6
    assert abs(loss.item() - float(c["loss"])) / abs(float(c["loss"])) < 1e-9


@pytest.mark.parametrize("mode", ["diag", "iso"])
def test_oracle_mgnll_matches_golden(mode):
    c = load_npz("case_mgnll.npz")
    pred = torch.from_numpy(c[f"{mode}.pred"]).double().requires_grad_(True)
    var = torch.from_numpy(c[f"{mode}.var"]).double().requires_grad_(True)
    targ = torch.from_numpy(c[f"{mode}.target"]).double()
    loss = O.mgnll(pred, targ, var, mode)
    loss.backward()
    assert abs(loss.item() - float(c[f"{mode}.loss"])) / abs(float(c[f"{mode}.loss"])) < 1e-6
    assert rel_l2(pred.grad, torch.from_numpy(c[f"{mode}.dpred"])) < 1e-5
    assert rel_l2(var.grad, torch.from_numpy(c[f"{mode}.dvar"])) < 1e-5
    cov = O.covariance(var.detach(), mode)
    assert cov.shape == (pred.shape[0], 1, 13, 13, pred.shape[3], pred.shape[4])
    assert abs(cov.sum().item() - float(c[f"{mode}.cov_diag_sum"])) / float(c[f"{mode}.cov_diag_sum"]) < 1e-5


def test_mgnll_error_semantics():
    p = torch.rand(2, 1, 13, 4, 4)
    with pytest.raises(ValueError, match="var has negative entry/entries"):
        O.mgnll(p, p, -torch.ones_like(p))
    with pytest.raises(ValueError, match="is not valid"):
        O.mgnll(p, p, torch.ones_like(p), reduction="bogus")


def test_maxpool_first_max_tie_rule():
    x = torch.zeros(1, 1, 64, 64)
    val, idx = O.adaptive_max_pool(x)
    ref_v, ref_i = torch.nn.functional.adaptive_max_pool2d(x, (32, 32), return_indices=True)
    assert torch.equal(idx, ref_i) and torch.equal(val, ref_v)
    x = torch.rand(2, 3, 96, 64)
    val, idx = O.adaptive_max_pool(x)
    ref_v, ref_i = torch.nn.functional.adaptive_max_pool2d(x, (32, 32), return_indices=True)
    assert torch.equal(idx, ref_i) and torch.equal(val, ref_v)


def test_bilinear_matches_torch():
    a = torch.rand(3, 2, 32, 32, dtype=torch.float64)
    up = O.bilinear_upsample(a, 256, 64)
    ref = torch.nn.functional.interpolate(a, size=(256, 64), mode="bilinear", align_corners=False)
    assert torch.allclose(up, ref, atol=1e-12)


def test_depthwise_reflect_matches_torch():
    x = torch.rand(2, 8, 9, 7, dtype=torch.float64)
    w = torch.rand(8, 1, 3, 3, dtype=torch.float64)
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect"), w, groups=8)
    assert torch.allclose(O.depthwise3x3_reflect(x, w), ref, atol=1e-12)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference_fp64():
    U, Lm, winit = ref_import.load()
    torch.manual_seed(3)
    m = U.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="diag",
                     scale_by=10.0)
    m.apply(winit)
    m = m.double().train()
    B, T, H, W = 1, 2, 64, 64
    x, y, d = O.synthetic_batch(B, T, H, W, seed=11, dtype=torch.float64)
    keep = O.dropout_keep_mask(16, B, T, H, W, seed=12)
    m.temporal_aggregator.attn_dropout = ref_import.InjectedDropout(keep)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    out = m(x, batch_positions=d)
    loss, var = Lm.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode="diag", chunk=None)(
        out[:, :, :13], y, out[:, :, 13:26])
    loss.backward()
    o_out, o_loss, o_g, _ = O.step(sd, x, y, d, O.OracleConfig(), True, keep)
    assert rel_l2(o_out, out) < 1e-12
    assert abs(o_loss.item() - loss.item()) / abs(loss.item()) < 1e-9
    for k, p in m.named_parameters():
        if not is_zero_grad_param(k):
            assert rel_l2(o_g[k], p.grad) < 1e-9, k
    assert torch.equal(O.covariance(out[:, :, 13:26].detach(), "diag"), var)


def test_fused_mode_equals_elementary(golden_weights):
    """bench.py times the oracle with FUSED=True (the ATen/oneDNN kernels the reference's modules dispatch to)."""
    cfg = O.OracleConfig()
    sd = _sd64(golden_weights)
    x, y, d = O.synthetic_batch(1, 2, 64, 64, seed=8, dtype=torch.float64)
    keep = O.dropout_keep_mask(16, 1, 2, 64, 64)
    a_out, a_loss, a_g, _ = O.step(sd, x, y, d, cfg, True, keep)
    O.set_fused(True)
    try:
        b_out, b_loss, b_g, _ = O.step(sd, x, y, d, cfg, True, keep)
    finally:
        O.set_fused(False)
    assert rel_l2(b_out, a_out) < 1e-12 and abs(a_loss.item() - b_loss.item()) < 1e-9 * abs(a_loss.item())
    for k in a_g:
        if not is_zero_grad_param(k):
            assert rel_l2(b_g[k], a_g[k]) < 1e-9, k


@pytest.mark.parametrize("mode", ["diag", "iso"])
def test_oracle_calibration_sweep_matches_golden(mode):
    """BASELINE config #5: head nonlinearities + MGNLL + rescale + per-sample statistics + UCE/AUCE of the oracle against the
    fixture generated from the unmodified reference (tests/golden/make_calibration.py)."""
    c = load_npz("case_calibration.npz")
    S = c["lm"].shape[0]
    vc = 13 if mode == "diag" else 1
    cfg = O.OracleConfig(covmode=mode)
    lm = torch.from_numpy(c["lm"]).double().requires_grad_(True)
    lv = torch.from_numpy(c[f"{mode}.lv"]).double().requires_grad_(True)
    y = torch.from_numpy(c["y"]).double()
    out = O.head(torch.cat([lm, lv], dim=1), cfg)
    loss = O.mgnll(out[:, :, :13], y, out[:, :, 13:13 + vc], mode)
    loss.backward()
    assert abs(loss.item() - float(c[f"{mode}.loss"])) / abs(float(c[f"{mode}.loss"])) < 1e-9
    assert rel_l2(lm.grad, torch.from_numpy(c[f"{mode}.dlm"])) < 1e-6
    assert rel_l2(lv.grad, torch.from_numpy(c[f"{mode}.dlv"])) < 1e-6
    var5 = O.variance_from_covariance(O.covariance(out[:, :, 13:13 + vc].detach(), mode) / cfg.scale_by ** 2)
    stats = [O.errvar_samplewise(y[s] / cfg.scale_by, out[s, :, :13].detach() / cfg.scale_by, var5[s]) for s in range(S)]
    mvar, err = [s["mean var"] for s in stats], [s["error"] for s in stats]
    assert np.allclose(mvar, c[f"{mode}.mean_var"], rtol=1e-9) and np.allclose(err, c[f"{mode}.error"], rtol=1e-9, atol=1e-12)
    uce, auce = O.compute_uce_auce(mvar, err, S)
    assert abs(uce - float(c[f"{mode}.uce"])) < 1e-6 and abs(auce - float(c[f"{mode}.auce"])) < 1e-6
    # the restated metric also agrees with the reference's on its own golden inputs
    uce2, auce2 = O.compute_uce_auce(c[f"{mode}.mean_var"], c[f"{mode}.error"], S)
    assert abs(uce2 - float(c[f"{mode}.uce"])) < 1e-7 and abs(auce2 - float(c[f"{mode}.auce"])) < 1e-7


def test_oracle_gnll_matches_golden():
    """GNLL (`--loss GNLL`, covmode 'uni'): oracle restatement of gaussian_nll_loss against the fixture generated from the
    unmodified reference (tests/golden/make_gnll.py)."""
    c = load_npz("case_gnll.npz")
    pred = torch.from_numpy(c["pred"]).double().requires_grad_(True)
    var = torch.from_numpy(c["var"]).double().requires_grad_(True)
    loss, vout = O.gnll(pred, torch.from_numpy(c["target"]).double(), var, full=True, eps=1e-8)
    loss.backward()
    assert abs(loss.item() - float(c["loss"])) / abs(float(c["loss"])) < 1e-9
    assert rel_l2(pred.grad, torch.from_numpy(c["dpred"])) < 1e-6 and rel_l2(var.grad, torch.from_numpy(c["dvar"])) < 1e-6
    assert rel_l2(vout, torch.from_numpy(c["var_out"])) < 1e-6
    with pytest.raises(ValueError, match="var has negative entry/entries"):
        O.gnll(pred.detach(), torch.from_numpy(c["target"]).double(), -var.detach())


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_oracle_gnll_matches_live_reference_fp64():
    _, Lm, _ = ref_import.load()
    g = torch.Generator("cpu").manual_seed(3)
    pred = torch.rand(2, 1, 13, 8, 8, generator=g, dtype=torch.float64).requires_grad_(True)
    targ = torch.rand(2, 1, 13, 8, 8, generator=g, dtype=torch.float64)
    var = (torch.rand(2, 1, 13, 8, 8, generator=g, dtype=torch.float64) * 2).requires_grad_(True)
    with torch.no_grad():
        var[0, 0, 0, 0, :3] = 1e-12
    l_ref, v_ref = Lm.GaussianNLLLoss(reduction="mean", eps=1e-8, full=True)(pred, targ, var)
    g_ref = torch.autograd.grad(l_ref, [pred, var])
    l_own, v_own = O.gnll(pred, targ, var, full=True, eps=1e-8)
    g_own = torch.autograd.grad(l_own, [pred, var])
    assert abs(l_ref.item() - l_own.item()) < 1e-12 * abs(l_ref.item())
    assert torch.allclose(v_ref, v_own) and all(torch.allclose(a, b, rtol=1e-12, atol=1e-15) for a, b in zip(g_ref, g_own))


def test_imposed_pool_index_is_a_noop_for_the_oracles_own_argmax(golden_weights):
    """`forward(..., pool_idx=...)` (diagnostic used by the 256x256 GPU gradient test): imposing the oracle's own argmax leaves
    output and gradients unchanged; imposing a different element of one window changes the encoder gradients only."""
    x, y, d = O.synthetic_batch(1, 2, 32, 32, seed=77)
    keep = O.dropout_keep_mask(16, 1, 2, 32, 32, seed=78)
    cfg = O.OracleConfig()
    sd = _sd64(golden_weights)
    taps = {}
    O.forward(sd, x.double(), d.double(), cfg, True, keep, None, taps)
    idx = taps["pool_idx"]
    out0, loss0, g0, _ = O.step(sd, x.double(), y.double(), d.double(), cfg, True, keep)
    out1, loss1, g1, _ = O.step(sd, x.double(), y.double(), d.double(), cfg, True, keep, pool_idx=idx)
    assert torch.equal(out0, out1) and float(loss0) == float(loss1)
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k


def test_oracle_img_metrics_match_live_reference():
    """O.img_metrics / O.ssim (checker of the GPU metrics kernels) against the reference's own metrics.py + pytorch_ssim."""
    import os, sys
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "model", "src")):
        pytest.skip("/root/reference absent")
    for p in (os.path.join(ref, "model"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    cwd = os.getcwd()
    os.chdir(os.path.join(ref, "model"))
    try:
        from src.learning.metrics import img_metrics as ref_img_metrics
    finally:
        os.chdir(cwd)
    g = torch.Generator().manual_seed(3)
    t, p = torch.rand(1, 13, 48, 40, generator=g), torch.rand(1, 13, 48, 40, generator=g)
    v = torch.rand(1, 13, 48, 40, generator=g) * 0.01
    want = ref_img_metrics(t, p, var=v)
    got = O.img_metrics(t, p, v)
    for k in ("RMSE", "MAE", "PSNR", "SAM", "SSIM", "error", "mean ae", "mean se", "mean var"):
        assert abs(got[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), (k, got[k], want[k])


# ---- constructor variants use_v / is_mono / separate_out (uncrtaints.py:296-338,376-379,414-430) ----
from variants import VARIANTS, variant_inputs, reference_model  # noqa: E402


def _variant_oracle_step(kw):
    cfg, p, x, y, d, keep, vkeep = variant_inputs(kw)
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
    out, loss, grads, newbuf = O.step(p64, x.double(), y.double(), d.double() if cfg.positional_encoding else None, cfg, True, keep,
                                      v_keep_mask=vkeep)
    return cfg, p, out, loss, grads, newbuf


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_variants_match_fixture(name):
    """The oracle against case_variants.npz (unmodified reference in fp64, tests/golden/make_variants.py)."""
    c = load_npz("case_variants.npz")
    cfg, p, out, loss, grads, _ = _variant_oracle_step(VARIANTS[name])
    assert rel_l2(out, torch.from_numpy(c[name + ".out"])) < 1e-6
    assert abs(loss.item() - float(c[name + ".loss"])) / abs(float(c[name + ".loss"])) < 1e-9
    if not any(k.startswith(name + ".grad.") for k in c):
        return                                            # outputs / loss only (variants.NO_FIXTURE_GRADS)
    scale = max(float(np.linalg.norm(c[k])) for k in c if k.startswith(name + ".grad."))
    for k, g in grads.items():
        ref = torch.from_numpy(c[name + ".grad." + k])
        if float(ref.norm()) <= 1e-7 * scale:
            assert float(g.norm()) <= 1e-6 * scale, k
        else:
            assert rel_l2(g, ref) < 1e-5, k


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_variants_match_live_reference_fp64(name):
    U, Lm, _ = ref_import.load()
    cfg, p, x, y, d, keep, vkeep = variant_inputs(VARIANTS[name])
    m = reference_model(U, cfg, p, keep, vkeep).double().train()
    out = m(x.double(), batch_positions=d.double())
    loss, _ = Lm.MultiGaussianNLLLoss(reduction="mean", eps=1e-8, full=True, mode=cfg.covmode, chunk=None)(
        out[:, :, :13], y.double(), out[:, :, 13:13 + cfg.covar_dim])
    loss.backward()
    _, _, o_out, o_loss, o_g, o_buf = _variant_oracle_step(VARIANTS[name])
    assert rel_l2(o_out, out) < 1e-12
    assert abs(o_loss.item() - loss.item()) / abs(loss.item()) < 1e-10
    scale = max(float(q.grad.norm()) for q in m.parameters() if q.grad is not None)
    for k, q in m.named_parameters():
        ref = q.grad if q.grad is not None else torch.zeros_like(q)
        if float(ref.norm()) <= 1e-9 * scale:
            assert float(o_g[k].norm()) <= 1e-8 * scale, k
        else:
            assert rel_l2(o_g[k], ref) < 1e-9, k
    sd = m.state_dict()
    for k, v in o_buf.items():                        # BatchNorm running statistics incl. the BatchNorm1d of the use_v MLP
        assert torch.allclose(v.double(), sd[k].double(), rtol=1e-10, atol=1e-12), k
    # eval mode (running statistics, no dropout: put real nn.Dropout modules back, the injected stand-ins ignore .training)
    if not cfg.is_mono:
        m.temporal_aggregator.attn_dropout = torch.nn.Dropout(cfg.dropout_p)
    if cfg.use_v:
        m.temporal_encoder.dropout = torch.nn.Dropout(cfg.v_dropout_p)
    m.eval()
    with torch.no_grad():
        e_ref = m(x.double(), batch_positions=d.double())
    p_eval = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    e_or = O.forward(p_eval, x.double(), d.double() if cfg.positional_encoding else None, cfg, training=False)
    assert rel_l2(e_or, e_ref) < 1e-12
