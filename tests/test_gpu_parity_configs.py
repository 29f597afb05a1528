"""GPU: parity AT THE BENCHED CONFIGURATIONS (BASELINE.json configs[1..3]): the engine against the oracle run on the same GPU in
fp32 (TF32 off: the reference's own GPU path is cuDNN fp32, buildingblocks.py:56,75), with the tolerance MEASURED in the test:

  (i)   probabilities within north_star's 1e-2 relative L2;
  (ii)  logits no further from the fp32 reference than 1.25x the drift torch's OWN bf16 path (autocast over the same oracle) shows on
        the same inputs, or within 1e-2 outright (SURVEY.md section 7 hard part 5: random-weight logits sit near zero, so their
        relative error is dominated by bf16 storage of O(1) activations -- for any bf16 implementation);
  (iii) the activation-pattern flip rate between the engine and the fp32 reference stays below 0.5 % of the elements of every
        ReLU, or -- in the deep decoder layers, where the inputs of a ReLU have themselves drifted by a few per cent -- below 1.25x
        the flip rate torch's own bf16 path shows at the same layer (so that pinning those masks for the gradient comparison
        cannot hide a forward bug: a wrong layer would flip a large fraction of its outputs);
  (iv)  every parameter gradient against the oracle evaluated at the engine's activation pattern / pool argmax.
"""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu

DRIFT_FACTOR = 1.25
PROB_TOL = 1e-2
FLIP_TOL = 5e-3
GRAD_TOL = 8e-2


def _run_case(cfg, shape, loss_name, seed=0):
    import pytorch3dunet_b200 as P
    from pytorch3dunet_b200 import engine as E
    from oracle import unet3d_oracle as O  # checker only
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    torch.manual_seed(seed)
    model = P.get_model(cfg)                      # the reference's default init under the seed (tests/test_host.py)
    with torch.no_grad():                         # perturb the GroupNorm affine so that path is exercised
        g = torch.Generator().manual_seed(seed + 1)
        for k, p in model.named_parameters():
            if "groupnorm" in k:
                p.add_(0.2 * torch.randn(p.shape, generator=g))
    model = model.to(dev)
    x = torch.rand(shape, device=dev)
    t = (torch.rand((shape[0], cfg["out_channels"]) + tuple(shape[2:]), device=dev) > 0.5).float()

    # ---- engine -------------------------------------------------------------------------------------------------------
    E.DEBUG = {}
    try:
        out, logits = model(x, return_logits=True)
        loss = getattr(P.losses, loss_name)(logits, t)
        loss.backward()
        torch.cuda.synchronize()
        masks = {k: (v > 0).permute(0, 4, 1, 2, 3) for k, v in E.DEBUG["fwd"].items()}
        for k, (y, gg, q) in E.DEBUG.get("se", {}).items():
            n, d, h, w, c = y.shape
            masks[k + "#select"] = ((y.float() * gg.view(n, 1, 1, 1, c)) >= (y.float() * q.view(n, d, h, w, 1))).permute(0, 4, 1, 2, 3)
        pool_idx = [F.max_pool3d(v.float().permute(0, 4, 1, 2, 3), 2, return_indices=True)[1] for v in E.DEBUG.get("pool", [])]
    finally:
        E.DEBUG = None
    eg = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    e_logits, e_out, e_loss = logits.detach(), out.detach(), loss.item()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    del model, out, logits, loss
    torch.cuda.empty_cache()

    # ---- the fp32 reference on the same GPU: forward, its own activation pattern, torch's own bf16 drift ----------------
    with torch.no_grad():
        rec = {}
        r_out, r_logits = O.forward(sd0, cfg, x, masks={"__record__": rec})
        rec_b = {}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            b_out, b_logits = O.forward(sd0, cfg, x, masks={"__record__": rec_b})
        drift = rel_l2(b_logits.float(), r_logits)
        drift_p = rel_l2(b_out.float(), r_out)
        flips = {k: (masks[k] != rec[k]).float().mean().item() for k in rec if k in masks}
        flips_torch = {k: (rec_b[k] != rec[k]).float().mean().item() for k in rec if k in rec_b}
        del b_out, b_logits, rec, rec_b
    rep = {"logits": rel_l2(e_logits, r_logits), "torch_bf16_logits": drift, "probs": rel_l2(e_out, r_out), "torch_bf16_probs": drift_p,
           "max_flip": max(flips.values()) if flips else 0.0}
    torch.cuda.empty_cache()

    # ---- gradients: the reference arithmetic at the engine's activation pattern ---------------------------------------
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    o_out, o_logits = O.forward(sd, cfg, x, masks=masks, pool_idx=pool_idx)
    o_loss = getattr(O, loss_name)(o_logits, t)
    o_loss.backward()
    rep["loss_abs"] = abs(e_loss - o_loss.item())
    worst = ("", 0.0)
    for k, g_e in eg.items():
        g_o = sd[k].grad
        r = rel_l2(g_e, g_o)
        # the GroupNorm affine in front of a conv feeds another GroupNorm: its true gradient is ~0 by scale/shift invariance
        # (the stem's is EXACTLY the invariant direction), so judge it against the sibling conv's gradient scale, not its own norm
        floor = 0.0
        if "groupnorm" in k:
            floor = 5e-3 * sd[k.rsplit("groupnorm", 1)[0] + "conv.weight"].grad.norm().item()
        err = (g_e.double() - g_o.double()).norm().item()
        if err > floor and r > worst[1] and g_o.norm() > 1e-4:
            worst = (k, r)
    rep["worst_grad"] = worst
    print(cfg["name"], shape, {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in rep.items()})
    print("  flip rates (engine / torch bf16):", {k: f"{v:.2e} / {flips_torch.get(k, float('nan')):.2e}"
                                                 for k, v in sorted(flips.items(), key=lambda kv: -kv[1])[:4]})
    assert rep["probs"] <= PROB_TOL, rep
    assert rep["logits"] <= max(DRIFT_FACTOR * drift, 1e-2), rep
    bad_flips = {k: (v, flips_torch.get(k)) for k, v in flips.items() if v >= max(FLIP_TOL, DRIFT_FACTOR * flips_torch.get(k, 0.0))}
    assert not bad_flips, bad_flips
    assert rep["loss_abs"] < 5e-3, rep
    assert worst[1] < GRAD_TOL, worst
    return rep


def test_cfg2_unet3d_f32_d4_2x128_bce_dice():
    """BASELINE configs[1], exactly: UNet3D f_maps=32, 4 levels, batch 2x1x128^3, BCEDiceLoss (the 384->128 and 192->64 virtual-concat
    layers, 221 tiles per persistent CTA, split-K wgrad over the whole chip)."""
    _run_case(dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=32, num_levels=4), (2, 1, 128, 128, 128), "bce_dice_loss")


def test_cfg3_residual_unet3d_f32_l5_widths():
    """BASELINE configs[2] widths: ResidualUNet3D f_maps=32, 5 levels (512 channels at the bottom, C_out > 256 slicing, transposed-conv
    joins), 2x1x96^3."""
    _run_case(dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=32, num_levels=5), (2, 1, 96, 96, 96), "bce_dice_loss", seed=1)


def test_cfg4_residual_unet_se3d_f64_l5_widths():
    """BASELINE configs[3] widths: ResidualUNetSE3D f_maps=64, 5 levels (1024 channels, scSE at every level), 1x1x64^3."""
    _run_case(dict(name="ResidualUNetSE3D", in_channels=1, out_channels=1, f_maps=64, num_levels=5), (1, 1, 64, 64, 64), "bce_dice_loss", seed=2)


def test_cfg4_residual_unet_se3d_f64_l5_widths_fp16_operands():
    """the same with fp16 activations / operands (what BASELINE configs[3] names): the drift gates are still torch's bf16 path's (looser)"""
    _run_case(dict(name="ResidualUNetSE3D", in_channels=1, out_channels=1, f_maps=64, num_levels=5, operand_dtype="fp16"), (1, 1, 64, 64, 64),
              "bce_dice_loss", seed=2)
