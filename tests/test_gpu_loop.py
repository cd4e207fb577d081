"""The reference's UNMODIFIED training loop on the B200 path (SURVEY.md §4 "boundary" level, §8 row a14): `iterate`
(train + val), `BaseModel.set_input / optimize_parameters / rescale`, `save_model -> load_checkpoint` from baseline/_ref, with
`uncrtaints_b200.install()` patched in, against the same loop run with the reference's own modules on the CPU."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_loop  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not ref_loop.available(), reason="baseline/_ref not installed (python baseline/install_ref.py)")
def test_unmodified_training_loop_on_the_b200_path(tmp_path):
    ref = ref_loop.run_loop("cpu", False, str(tmp_path / "ref"), hw=64, steps=2, dropout_p=0.0)
    assert ref["netG_class"] == "src.backbones.uncrtaints.UNCRTAINTS"
    new = ref_loop.run_loop("cuda", True, str(tmp_path / "b200"), hw=64, steps=2, init_state=ref["init"], dropout_p=0.0)
    assert new["netG_class"] == "uncrtaints_b200.backbone.UNCRTAINTS"
    # same state-dict surface (151 netG entries under the `netG.` prefix), a reference checkpoint loaded strict=True above
    assert list(new["init"].keys()) == list(ref["init"].keys())
    # step 1: forward + MGNLL + backward; step 2 additionally sees the Adam update of step 1 (base_model.py:115-122)
    for a, b in zip(new["losses"], ref["losses"]):
        assert abs(a - b) <= 2e-3 * abs(b), (new["losses"], ref["losses"])
    # parameters after two optimizer steps (Adam's normalised step amplifies tiny gradient differences: loose tolerance)
    num = sum(float((new["after"][k].double() - ref["after"][k].double()).pow(2).sum()) for k in ref["after"] if ref["after"][k].is_floating_point())
    den = sum(float(ref["after"][k].double().pow(2).sum()) for k in ref["after"] if ref["after"][k].is_floating_point())
    assert (num / den) ** 0.5 <= 5e-3
    # the reference's own checkpoint code round-trips the B200 model strictly
    for k, v in new["after"].items():
        assert torch.equal(v, new["reloaded"][k]), k
    # validation pass of the unmodified loop (eval mode, no_grad, img_metrics incl. SSIM, UCE/AUCE) on the B200 path
    (m_new, im_new), (m_ref, im_ref) = new["val"], ref["val"]
    assert abs(m_new["val_loss"] - m_ref["val_loss"]) <= 1e-2 * abs(m_ref["val_loss"])
    for key in ("RMSE", "MAE", "PSNR", "SSIM", "mean var"):
        assert abs(im_new[key] - im_ref[key]) <= 1e-2 * abs(im_ref[key]) + 1e-6, (key, im_new[key], im_ref[key])
    # and a checkpoint written by the reference loads into the B200 model through the reference's loader
    ck = torch.load(ref["ckpt"], map_location="cpu")["state_dict_G"]
    import uncrtaints_b200 as ub
    net = ub.UNCRTAINTS(input_dim=15, out_conv=[26], out_nonlin_mean=True, out_nonlin_var="softplus", covmode="diag", scale_by=10.0)
    net.load_state_dict(ck, strict=True)
