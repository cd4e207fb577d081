"""Shared definition of the constructor-variant cases (use_v / is_mono / separate_out / block_type='residual';
uncrtaints.py:24-69,291-338,376-379,414-430):
used by tests/golden/make_variants.py (fixture from the unmodified reference), tests/test_oracle.py (oracle vs fixture / live
reference) and tests/test_gpu_extras.py (CUDA path vs oracle and fixture)."""
import torch

from oracle import uncrtaints_oracle as O, ref_import

VARIANTS = {
    "use_v": dict(use_v=True, n_dec_blocks=1),
    "use_v_pad_nope": dict(use_v=True, n_dec_blocks=1, positional_encoding=False, pad_last=True, covmode="iso"),
    "is_mono": dict(is_mono=True, n_dec_blocks=1),
    "separate_out": dict(separate_out=True, n_dec_blocks=1),
    "residual": dict(block_type="residual", n_dec_blocks=1, pad_last=True),
    "residual_gn_dec": dict(block_type="residual", n_dec_blocks=1, encoder_norm="batch", decoder_norm="group", covmode="iso"),
}


# fixtures of these cases hold outputs and loss only (their gradients are checked against the oracle, which the other cases pin)
NO_FIXTURE_GRADS = {"residual_gn_dec"}


def variant_inputs(kw):
    """(cfg, params, x, y, dates, attention keep mask, value keep mask) of a variant; everything seeded."""
    kw = dict(kw)
    pad_last = kw.pop("pad_last", False)
    cfg = O.OracleConfig(**kw)
    p = O.init_params(cfg, seed=21)
    B, T, H, W = 2, (1 if cfg.is_mono else 3), 64, 64
    x, y, d = O.synthetic_batch(B, T, H, W, cin=cfg.input_dim, scale_by=cfg.scale_by, seed=31, pad_last=pad_last)
    keep = O.dropout_keep_mask(16, B, T, H, W, seed=32)
    vkeep = O.value_keep_mask(B, seed=33) if cfg.use_v else None
    return cfg, p, x, y, d, keep, vkeep


def reference_model(U, cfg, p, keep, vkeep):
    """The unmodified reference class with the variant's weights and injected dropout masks (float32; caller casts)."""
    m = U.UNCRTAINTS(input_dim=cfg.input_dim, decoder_widths=[128] * cfg.n_dec_blocks, out_conv=[13 + cfg.covar_dim],
                     out_nonlin_mean=cfg.out_nonlin_mean, out_nonlin_var="softplus", encoder_norm=cfg.encoder_norm,
                     decoder_norm=cfg.decoder_norm, positional_encoding=cfg.positional_encoding, covmode=cfg.covmode,
                     scale_by=cfg.scale_by, use_v=cfg.use_v, is_mono=cfg.is_mono, separate_out=cfg.separate_out,
                     block_type=cfg.block_type)
    m.load_state_dict(p, strict=True)
    if not cfg.is_mono:
        m.temporal_aggregator.attn_dropout = ref_import.InjectedDropout(keep, p=cfg.dropout_p)
    if cfg.use_v:
        m.temporal_encoder.dropout = ref_import.InjectedDropout(vkeep, p=cfg.v_dropout_p)
    return m
