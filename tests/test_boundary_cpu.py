"""CPU-side tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the Python shim reproduces the reference's constructor / state-dict surface, and nothing falls back to the CPU."""
import ctypes
import json
import os
import re

import pytest
import torch

from conftest import ROOT, GOLDEN


def test_library_exports_every_declared_symbol():
    from uncrtaints_b200 import _lib
    header = open(os.path.join(ROOT, "include", "uncrtaints_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ub200_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    assert sorted(_lib.SYMBOLS) == declared
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    assert _lib.lib().ub200_version() >= 100


def test_desc_struct_and_workspace_query_without_gpu():
    from uncrtaints_b200 import _lib
    L = _lib.lib()
    d = _lib.Desc()
    d.B, d.T, d.C_in, d.H, d.W, d.n_dec_blocks, d.out_dim, d.need_grad = 16, 3, 15, 256, 256, 5, 26, 1
    n = L.ub200_workspace_bytes(d)
    assert 30e9 < n < 60e9                      # ~40 GB of saved activations + scratch at BASELINE config #2
    assert L.ub200_num_param_slots(d) == _lib.UB200_P_BLOCK0 + 6 * _lib.UB200_BLOCK_STRIDE
    d.H = 250                                   # not a multiple of 32 -> unsupported
    assert L.ub200_workspace_bytes(d) == 0
    d.H, d.T = 256, 9                           # T > 8: run-time-T temporal kernels (up to 64)
    assert L.ub200_workspace_bytes(d) > n
    d.T = 65
    assert L.ub200_workspace_bytes(d) == 0
    d.T, d.need_grad = 3, 0                     # forward only: shared hidden buffers, ping-pong decoder outputs
    assert L.ub200_workspace_bytes(d) < n / 2
    d.need_grad, d.gemm_backend = 1, 1          # mixed backends do not exist: 0 (CUDA cores) or 3 (tcgen05) plus flags
    assert L.ub200_workspace_bytes(d) == 0
    d.gemm_backend = 3 | 32                     # bf16 hidden storage
    assert L.ub200_workspace_bytes(d) < 0.8 * n
    off, nb = ctypes.c_size_t(), ctypes.c_size_t()
    d.T = 3
    assert L.ub200_workspace_tap(d, b"blk3.h2", ctypes.byref(off), ctypes.byref(nb)) == 0
    assert nb.value == 16 * 65536 * 256 * 2     # bf16 hidden storage: 2 bytes per element
    d.gemm_backend = 3
    assert L.ub200_workspace_tap(d, b"blk3.h2", ctypes.byref(off), ctypes.byref(nb)) == 0 and nb.value == 16 * 65536 * 256 * 4
    assert L.ub200_workspace_tap(d, b"nonsense", ctypes.byref(off), ctypes.byref(nb)) == -1


def test_state_dict_surface_matches_reference():
    import uncrtaints_b200 as ub
    want = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    net = ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="diag", scale_by=10.0)
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(got.keys()) == list(want.keys())
    assert got == want
    assert sum(p.numel() for p in net.parameters()) == 570010
    assert (net.mean_idx, net.vars_idx, net.variance) == (13, 26, None)
    iso = ub.UNCRTAINTS(input_dim=15, out_conv=[14], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="iso", scale_by=10.0)
    assert (iso.mean_idx, iso.vars_idx) == (13, 14)
    # weight_init dispatches on isinstance (learning/weight_init.py:13-47): the holders must be the real nn types
    kinds = {type(m).__name__ for m in net.modules()}
    assert {"Conv2d", "Conv1d", "Linear", "GroupNorm", "BatchNorm2d"} <= kinds


def test_unsupported_configurations_raise():
    import uncrtaints_b200 as ub
    kw = dict(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", scale_by=10.0)
    for bad in (dict(block_type="bogus"), dict(n_head=4), dict(encoder_widths=[64]),
                dict(agg_mode="mean"), dict(out_nonlin_var="relu"), dict(out_conv=[13])):
        with pytest.raises(NotImplementedError):
            ub.UNCRTAINTS(**{**kw, **bad})


def test_variant_module_trees_match_the_reference_state_dict():
    """use_v / is_mono / separate_out build the reference's module tree: same keys, same order, same shapes (checkpoints and
    weight_init interoperate).  Needs /root/reference for the comparison; the key lists of the variants are also pinned by
    oracle.init_params, which the fixtures of tests/golden/make_variants.py were generated with."""
    import uncrtaints_b200 as ub
    from oracle import ref_import, uncrtaints_oracle as O
    kw = dict(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", scale_by=10.0)
    for var in (dict(use_v=True), dict(is_mono=True), dict(separate_out=True), dict(separate_out=True, covmode=None, out_conv=[13]),
                dict(block_type="residual"), dict(block_type="residual", use_v=True)):
        net = ub.UNCRTAINTS(**{**kw, **var})
        cfg = O.OracleConfig(use_v=var.get("use_v", False), is_mono=var.get("is_mono", False), separate_out=var.get("separate_out", False),
                             covmode=var.get("covmode", "diag"), block_type=var.get("block_type", "mbconv"))
        mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert mine == {k: tuple(v.shape) for k, v in O.init_params(cfg).items()}
        if ref_import.available():
            U, _, winit = ref_import.load()
            ref = U.UNCRTAINTS(**{**kw, **var})
            assert list(ref.state_dict().keys()) == list(net.state_dict().keys())
            assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == mine
            ref.apply(winit)
            net.apply(winit)                              # isinstance dispatch of weight_init works on the holders
            net.load_state_dict(ref.state_dict(), strict=True)


def test_cpu_tensors_are_rejected_not_emulated():
    import uncrtaints_b200 as ub
    net = ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", scale_by=10.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 2, 15, 32, 32), batch_positions=torch.zeros(1, 2))
    p = torch.rand(1, 1, 13, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ub.MultiGaussianNLLLoss(mode="diag", chunk=None)(p, p, p + 1)
    with pytest.raises(ValueError, match="is not valid"):
        ub.MultiGaussianNLLLoss(mode="diag", chunk=None, reduction="bogus")(p, p, p + 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "uncrtaints_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn


def test_ltae_fold_is_exact_in_fp64(golden_weights):
    """Ap / e folding (uncrtaints_b200/backbone.py:fold_ltae) reproduces the reference score chain."""
    import uncrtaints_b200 as ub
    from oracle import uncrtaints_oracle as O
    net = ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", scale_by=10.0)
    net.load_state_dict(golden_weights)
    net = net.double()
    with torch.no_grad():
        net.temporal_encoder.in_norm.weight.mul_(1.3).add_(0.1)
        net.temporal_encoder.in_norm.bias.add_(0.2)
    B, T = 2, 3
    g = torch.Generator().manual_seed(0)
    down = torch.randn(B, T, 128, 32, 32, generator=g, dtype=torch.float64)
    dates = torch.tensor([[1400., 1500., 1890.], [1411., 1412., 1700.]], dtype=torch.float64)
    pad = torch.tensor([[False, False, True], [False, False, False]])
    p = {k: v for k, v in net.state_dict().items()}
    ref = O.ltae_tiny(down, dates, pad, p, O.OracleConfig())                     # [16,B,T,32,32]
    ap, e = ub.backbone.fold_ltae(net.temporal_encoder, dates, B, T, "cpu")
    seq = down.permute(0, 3, 4, 2, 1).reshape(B * 1024, 128, T)
    xg = seq.reshape(B * 1024, 16, -1)
    xh = ((xg - xg.mean(2, keepdim=True)) / torch.sqrt(xg.var(2, unbiased=False, keepdim=True) + 1e-5)).reshape(seq.shape)
    score = 0.5 * (torch.einsum("hc,nct->nht", ap, xh) + e.double().permute(0, 2, 1).repeat_interleave(1024, 0))
    score = score.masked_fill(pad.reshape(B, 1, T).repeat_interleave(1024, 0), -1e3)
    mine = torch.softmax(score, -1).reshape(B, 32, 32, 16, T).permute(3, 0, 4, 1, 2)
    assert float((mine - ref).abs().max()) < 1e-9


@pytest.mark.skipif(not os.path.isdir("/root/reference/model/src"), reason="/root/reference not present (GPU box)")
def test_install_patches_reference_factory_and_loss():
    """uncrtaints_b200.install() makes the reference's own factory (model_utils.get_generator, model_utils.py:85-108) and
    loss factory (losses.get_loss, losses.py:14-32) build the B200 classes, with the reference's config namespace."""
    import sys
    import types
    from unittest import mock
    import argparse
    sys.path.insert(0, "/root/reference/model")
    cwd = os.getcwd()
    stubs = {"fvcore": types.ModuleType("fvcore"), "fvcore.nn": types.ModuleType("fvcore.nn")}
    stubs["fvcore.nn"].FlopCountAnalysis = mock.MagicMock()
    stubs["fvcore.nn"].flop_count_table = mock.MagicMock()
    saved = None
    try:
        with mock.patch.dict(sys.modules, stubs):
            import uncrtaints_b200 as ub
            from src import losses as ref_losses
            from src.backbones import uncrtaints as ref_uncrtaints
            saved = (ref_uncrtaints.UNCRTAINTS, ref_losses.MultiGaussianNLLLoss, ref_losses.GaussianNLLLoss)     # restored below: other tests use the reference
            ub.install(verbose=False)
            from src import model_utils                                # model_utils chdirs into ./model if it exists
            assert ref_uncrtaints.UNCRTAINTS is ub.UNCRTAINTS
            assert ref_losses.MultiGaussianNLLLoss is ub.MultiGaussianNLLLoss
            cfg = argparse.Namespace(model="uncrtaints", use_sar=True, encoder_widths=[128], decoder_widths=[128] * 5,
                                     out_conv=[26], mean_nonLinearity=True, var_nonLinearity="softplus", agg_mode="att_group",
                                     encoder_norm="group", decoder_norm="batch", n_head=16, d_model=256, d_k=4, pad_value=0,
                                     padding_mode="reflect", positional_encoding=True, covmode="diag", scale_by=10.0,
                                     separate_out=False, use_v=False, block_type="mbconv", pretrain=False, loss="MGNLL",
                                     chunk_size=None)
            net = model_utils.get_generator(cfg)
            assert isinstance(net, ub.UNCRTAINTS) and (net.mean_idx, net.vars_idx) == (13, 26)
            # the factory's variant flags (model_utils.py:86-108: --use_v, --block_type residual, --pretrain -> is_mono, separate_out)
            for flags in (dict(use_v=True), dict(block_type="residual"), dict(pretrain=True), dict(separate_out=True)):
                vnet = model_utils.get_generator(argparse.Namespace(**{**vars(cfg), **flags}))
                assert isinstance(vnet, ub.UNCRTAINTS)
                assert (vnet.use_v, vnet.block_type, vnet.is_mono, vnet.separate_out) == (
                    flags.get("use_v", False), flags.get("block_type", "mbconv"), flags.get("pretrain", False), flags.get("separate_out", False))
            crit = ref_losses.get_loss(cfg)
            p = torch.rand(1, 1, 13, 4, 4)
            with pytest.raises(RuntimeError, match="CUDA"):           # our loss, reached through the reference's wrapper
                ref_losses.calc_loss(crit, cfg, p, p, p + 1)
            # weight_init of the reference dispatches on module types (weight_init.py:13-47) and must touch every tensor it would
            from src.learning.weight_init import weight_init
            before = {k: v.clone() for k, v in net.state_dict().items()}
            torch.manual_seed(0)
            net.apply(weight_init)
            changed = [k for k, v in net.state_dict().items() if not torch.equal(v, before[k])]
            assert "in_conv.conv.conv.0.weight" in changed and "out_block.4.conv.fn.8.weight" in changed
            assert "temporal_encoder.inconv.bias" in changed and "temporal_encoder.attention_heads.Q" not in changed
    finally:
        os.chdir(cwd)
        if saved is not None:
            ref_uncrtaints.UNCRTAINTS, ref_losses.MultiGaussianNLLLoss, ref_losses.GaussianNLLLoss = saved


def test_prepare_data_multi_matches_the_reference(tmp_path):
    """uncrtaints_b200.prepare_data_multi (pinned staging in the final layout; here on the CPU) returns what the reference's
    prepare_data_multi (model/train_reconstruct.py:161-179) returns for the same collated batch, with and without SAR."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_loop
    if not ref_loop.available():
        pytest.skip("baseline/_ref not installed")
    import uncrtaints_b200 as ub
    tr = ref_loop.load_train_module(ref_loop.README_FLAGS + ["--device", "cpu"], str(tmp_path))
    data = ref_loop.SyntheticSEN12MSCRTS(4, 3, 32, seed=5)
    batch = next(iter(torch.utils.data.DataLoader(data, batch_size=2, shuffle=False)))
    for use_sar in (True, False):
        tr.config.use_sar = use_sar
        want = tr.prepare_data_multi(batch, "cpu", tr.config)
        got = ub.prepare_data_multi(batch, "cpu", tr.config)
        for name, a, b in zip(("x", "y", "in_m", "dates"), got, want):
            assert a.shape == b.shape, (name, a.shape, b.shape)
            assert torch.equal(a.float(), b.float()), name


def test_bench_byte_model_of_the_block_kernel_classes():
    """bench.py's roofline numerators: per-class bytes of one step add up to the pass structure the library runs --
    the Norm3-backward statistics pass on its own for the encoder block and the last decoder block only, one more read stream (A)
    in the residual pass of the four decoder blocks above a fused one."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    B, T, n_dec, P = 16, 3, 5, 256 * 256
    A, Hh = 128 * P * 4, 256 * P * 4
    frames = B * T + n_dec * B
    assert bench.class_bytes("dwconv_bwd", B, T, n_dec, P) == 4 * Hh * frames
    assert bench.class_bytes("norm_bwd_stats", B, T, n_dec, P) == 2 * A * (B * T + B)
    assert bench.class_bytes("residual_bwd", B, T, n_dec, P) == 4 * A * frames + A * (n_dec - 1) * B
    assert bench.class_bytes("dwconv_bwd", B, T, n_dec, P, hid_bytes=2) == 2 * Hh * frames
    # the whole-step model of SURVEY.md §8(d): 12.06 GB per sample at T=3, diag, fp32
    assert abs(bench.survey_bytes_per_sample(3, 13, P) / 1e9 - 12.06) < 0.01
