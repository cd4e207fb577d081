"""Test infrastructure: run the reference's UNMODIFIED training / validation loop (baseline/_ref/model/train_reconstruct.py:
``prepare_data_multi`` :161-179, ``iterate`` :279-447; ``BaseModel.set_input / optimize_parameters / rescale``,
model/src/backbones/base_model.py:87-131; ``save_model`` / ``load_checkpoint``, model/src/model_utils.py:117-219) on a synthetic
dataset, either with the reference's own modules or with ``uncrtaints_b200.install()`` patched in.

The reference imports six packages that are neither installed nor needed for the loop itself: ``fvcore``
(FLOP counting, base_model.py:5-6), ``torchnet`` (a running-average meter, train_reconstruct.py:29), ``matplotlib`` (plots),
``rasterio`` / ``s2cloudless`` / ``natsort`` (GeoTIFF loader, data/dataLoader.py:6-15).  They are stubbed in ``sys.modules``;
the loop, the model wrapper and the checkpoint code are the reference's own files.
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import types
from unittest import mock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "model", "train_reconstruct.py"))


class _AverageValueMeter:                       # torchnet.meter.AverageValueMeter: add(value), value() -> (mean, std)
    def __init__(self):
        self.vals = []

    def add(self, v, n=1):
        self.vals.append(float(v))

    def value(self):
        m = sum(self.vals) / max(len(self.vals), 1)
        return m, 0.0

    def reset(self):
        self.vals = []


def _install_stubs():
    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m
    fv = stub("fvcore")
    fv.nn = stub("fvcore.nn", FlopCountAnalysis=mock.MagicMock(), flop_count_table=mock.MagicMock())
    tnt = stub("torchnet")
    tnt.meter = stub("torchnet.meter", AverageValueMeter=_AverageValueMeter)
    mpl = stub("matplotlib")
    plt = mock.MagicMock()                       # every plotting call is a no-op; subplots() must unpack into (fig, ax)
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    plt.Figure = type("Figure", (), {})
    mpl.pyplot = plt
    sys.modules["matplotlib.pyplot"] = plt
    ras = stub("rasterio", open=mock.MagicMock())
    ras.merge = stub("rasterio.merge", merge=mock.MagicMock())
    stub("s2cloudless", S2PixelCloudDetector=mock.MagicMock())
    stub("natsort", natsorted=sorted)
    try:
        import PIL  # noqa: F401  (util/utils.py:5)
    except ImportError:
        pil = stub("PIL")
        pil.Image = stub("PIL.Image")


def load_train_module(argv, res_dir):
    """Import baseline/_ref/model/train_reconstruct.py as a module (its top level parses sys.argv and builds `config`)."""
    _install_stubs()
    model_dir = os.path.join(REF, "model")
    for p in (model_dir, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = ["train_reconstruct.py"] + list(argv) + ["--res_dir", res_dir]
    os.chdir(model_dir)                        # README.md:77: the scripts run from model/
    try:
        for name in ("train_reconstruct",):
            sys.modules.pop(name, None)
        with mock.patch("torch.utils.tensorboard.SummaryWriter", mock.MagicMock()):
            tr = importlib.import_module("train_reconstruct")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    return tr


class SyntheticSEN12MSCRTS(torch.utils.data.Dataset):
    """Samples in the format of data/dataLoader.py:364-380 (`SEN12MSCRTS.__getitem__`, sample_type 'cloudy_cloudfree'): per time
    point lists of S1 [2,H,W] / S2 [13,H,W] / masks [H,W] in [0,1], acquisition-day offsets, one cloud-free S2 target."""

    def __init__(self, n, t, hw, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.items = []
        for _ in range(n):
            days = torch.sort(torch.randint(1400, 1901, (t,), generator=g)).values
            self.items.append({
                "input": {"S1": [torch.rand(2, hw, hw, generator=g) for _ in range(t)],
                          "S2": [torch.rand(13, hw, hw, generator=g) for _ in range(t)],
                          "masks": [(torch.rand(hw, hw, generator=g) > 0.5).float() for _ in range(t)],
                          "coverage": [0.5] * t,
                          "S1 TD": [int(d) for d in days], "S2 TD": [int(d) + 1 for d in days],
                          "S1 path": [""] * t, "S2 path": [""] * t, "idx": 0},
                "target": {"S2": [torch.rand(13, hw, hw, generator=g)], "S2 TD": [int(days[-1]) + 5], "S2 path": [""]},
                "coverage bin": 0})

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


README_FLAGS = ["--experiment_name", "loop_test", "--model", "uncrtaints", "--input_t", "3", "--region", "all", "--epochs", "1",
                "--lr", "0.001", "--batch_size", "2", "--gamma", "1.0", "--scale_by", "10.0", "--trained_checkp", "", "--loss", "MGNLL",
                "--covmode", "diag", "--var_nonLinearity", "softplus", "--display_step", "1000000", "--use_sar", "--block_type", "mbconv",
                "--n_head", "16", "--plot_every", "-1", "--export_every", "-1", "--num_workers", "0"]     # README.md:78


def run_loop(device: str, use_b200: bool, res_dir: str, hw: int = 64, steps: int = 2, init_state=None, dropout_p=None):
    """Build BaseModel through the reference's get_model, run `iterate(mode='train')` for `steps` batches and one validation
    pass, save and reload a checkpoint.  Returns dict(losses, state_dict (cpu), val_metrics, model class name)."""
    tr = load_train_module(README_FLAGS + ["--device", device], res_dir)
    config = tr.config
    config.device = device
    if use_b200:
        import uncrtaints_b200
        uncrtaints_b200.install(verbose=False)
    else:                                         # make sure a previous install() in this process is undone
        import importlib as _il
        ref_unc = _il.import_module("src.backbones.uncrtaints")
        ref_losses = _il.import_module("src.losses")
        _il.reload(ref_unc)
        _il.reload(ref_losses)
    os.makedirs(os.path.join(res_dir, config.experiment_name), exist_ok=True)
    torch.manual_seed(config.rdm_seed)
    model = tr.get_model(config)
    model = model.to(device)
    model.netG.apply(tr.weight_init)              # train_reconstruct.py:627
    if init_state is not None:
        model.load_state_dict(init_state, strict=True)
    if dropout_p is not None:
        model.netG.temporal_aggregator.attn_dropout.p = dropout_p
    init = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    data = SyntheticSEN12MSCRTS(steps * config.batch_size, config.input_t, hw)
    loader = torch.utils.data.DataLoader(data, batch_size=config.batch_size, shuffle=False, num_workers=0)
    losses = []
    orig_opt = model.optimize_parameters

    def spy():
        orig_opt()
        losses.append(float(model.loss_G.detach()))
    model.optimize_parameters = spy
    model.train()
    writer = mock.MagicMock()
    tr.iterate(model, loader, config, writer, mode="train", epoch=1, device=device)
    model.optimize_parameters = orig_opt
    model.eval()
    val = tr.iterate(model, loader, config, writer, mode="val", epoch=1, device=device)
    tr.save_model(config, 1, model, "model")
    after = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    # reload through the reference's own checkpoint code into a freshly built model: strict=True must succeed
    model2 = tr.get_model(config).to(device)
    tr.load_checkpoint(config, res_dir, model2, "model")
    reloaded = {k: v.detach().cpu().clone() for k, v in model2.state_dict().items()}
    return {"losses": losses, "init": init, "after": after, "reloaded": reloaded, "val": val,
            "netG_class": type(model.netG).__module__ + "." + type(model.netG).__name__,
            "ckpt": os.path.join(res_dir, config.experiment_name, "model.pth.tar")}
