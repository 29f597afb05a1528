"""CPU: the oracle restatement (oracle/unet3d_oracle.py) against golden vectors produced by the
reference's own classes (oracle/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

from oracle import unet3d_oracle as O
from tests.helpers import assert_close_l2, load_golden, rel_l2

TOL = 2e-5  # same ATen CPU kernels on both sides; slack only for thread-count dependent reductions

MODEL_CASES = {
    "unet3d_f16_l3_s16": (dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3), "bce_dice_loss"),
    "unet3d_f16_l3_dice_b2": (dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3), "dice_loss"),
    "unet3d_f16_l3_odd": (dict(name="UNet3D", in_channels=2, out_channels=3, f_maps=16, num_levels=3, final_sigmoid=False), "bce_dice_loss"),
    "unet3d_f16_l2_cgr": (dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, layer_order="cgr"), "bce_dice_loss"),
    "resunet3d_f16_l3_s16": (dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3), "bce_dice_loss"),
    "resunetse3d_f16_l3_s16": (dict(name="ResidualUNetSE3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3), "bce_dice_loss"),
    # further layer orders / activations / heads (pin the oracle; CPU only)
    "unet3d_f16_l2_cl": (dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, layer_order="cl"), "bce_dice_loss"),
    "unet3d_f16_l2_crg": (dict(name="UNet3D", in_channels=1, out_channels=2, f_maps=16, num_levels=2, layer_order="crg",
                               final_sigmoid=False), "bce_dice_loss"),
    "resunet3d_f16_l2_gcl": (dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, layer_order="gcl"),
                             "bce_dice_loss"),
    "resunetse3d_f16_l2_gce": (dict(name="ResidualUNetSE3D", in_channels=2, out_channels=1, f_maps=16, num_levels=2, layer_order="gce"),
                               "dice_loss"),
    # upsampling modes reachable through the model config + the residual block's default order
    "unet3d_f16_l2_trilinear": (dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, upsample="trilinear"),
                                "bce_dice_loss"),
    "unet3d_f16_l2_deconv": (dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, upsample="deconv"), "bce_dice_loss"),
    "resunet3d_f16_l2_cge": (dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, layer_order="cge"),
                             "bce_dice_loss"),
    "resunet3d_f16_l2_deconvcat": (dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, upsample="deconv"),
                                   "bce_dice_loss"),
}


@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_model_matches_reference_golden(name):
    cfg, loss_name = MODEL_CASES[name]
    rec, sd, grads = load_golden(name)
    # the state_dict contract: exactly the reference's keys and shapes
    shapes = O.param_shapes(cfg)
    assert set(shapes) == set(sd), set(shapes) ^ set(sd)
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == tuple(s), k
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = rec["x"].clone().requires_grad_(True)
    out, logits = O.forward(sd, cfg, x)
    loss = getattr(O, loss_name)(logits, rec["target"])
    loss.backward()
    assert rel_l2(out, rec["out"]) < TOL
    assert rel_l2(logits, rec["logits"]) < TOL
    assert abs(loss.item() - rec["loss"].item()) < 1e-5
    assert_close_l2(x.grad, rec["grad_x"], 1e-4, 1e-8, "grad_x")
    for k, g in grads.items():
        assert_close_l2(sd[k].grad, g, 1e-4, 1e-7, k)


def _block_fn(name):
    if name.startswith("block_singleconv_gcr"):
        return lambda sd, x, enc: O.single_conv(x, sd, "", "gcr", 8)
    if name.startswith("block_singleconv_cr"):
        return lambda sd, x, enc: O.single_conv(x, sd, "", "cr", 8)
    if name.startswith("block_doubleconv"):
        return lambda sd, x, enc: O.double_conv(x, sd, "", "gcr", 8)
    if name.startswith("block_encoder_pool"):
        return lambda sd, x, enc: O.double_conv(torch.nn.functional.max_pool3d(x, 2), sd, "basic_module.", "gcr", 8)
    if name.startswith("block_decoder_cat"):
        def f(sd, x, enc):
            u = torch.nn.functional.interpolate(x, size=enc.shape[2:], mode="nearest")
            return O.double_conv(torch.cat((enc, u), 1), sd, "basic_module.", "gcr", 8)
        return f
    if name.startswith("block_resnetse"):
        return lambda sd, x, enc: O.res_block(x, sd, "", "gcr", 8, se=True)
    if name.startswith("block_resnet"):
        return lambda sd, x, enc: O.res_block(x, sd, "", "gcr", 8)
    if name.startswith("block_decoder_deconv"):
        def f(sd, x, enc):
            u = torch.nn.functional.conv_transpose3d(x, sd["upsampling.upsample.conv_transposed.weight"], None, stride=2, padding=1)
            u = torch.nn.functional.interpolate(u, size=enc.shape[2:])
            return O.res_block(enc + u, sd, "basic_module.", "gcr", 8)
        return f
    raise KeyError(name)


BLOCKS = ["block_singleconv_gcr_16_32", "block_singleconv_cr_16_16", "block_doubleconv_enc_32_64",
          "block_doubleconv_dec_96_32", "block_encoder_pool_32_64", "block_decoder_cat_64_32",
          "block_decoder_cat_odd", "block_resnet_16_32", "block_resnetse_32_32", "block_decoder_deconv_32_16"]


@pytest.mark.parametrize("name", BLOCKS)
def test_block_matches_reference_golden(name):
    rec, sd, grads = load_golden(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = rec["x"].clone().requires_grad_(True)
    enc = rec["enc"].clone().requires_grad_(True) if "enc" in rec else None
    y = _block_fn(name)(sd, x, enc)
    (y * rec["r"]).sum().backward()
    assert rel_l2(y, rec["y"]) < TOL
    assert_close_l2(x.grad, rec["grad_x"], 1e-4, 1e-8, "grad_x")
    if enc is not None:
        assert_close_l2(enc.grad, rec["grad_enc"], 1e-4, 1e-8, "grad_enc")
    for k, g in grads.items():
        assert_close_l2(sd[k].grad, g, 1e-4, 1e-7, k)


def test_random_state_dict_runs():
    cfg = dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=8, num_levels=3)
    sd = O.random_state_dict(cfg, seed=3)
    out, logits = O.forward(sd, cfg, torch.rand(1, 1, 8, 8, 8))
    assert out.shape == (1, 1, 8, 8, 8) and 0 <= out.min() and out.max() <= 1
