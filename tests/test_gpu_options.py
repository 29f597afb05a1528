"""GPU: options behind the boundary beyond the defaults -- block-level AvgPool3d encoder, trilinear / nearest joins at odd sizes,
gradient w.r.t. both decoder inputs, the 2-replica nn.DataParallel path the reference trainer / predictor use, the flat gradient
buffer, and the fused Adam step against torch.optim.Adam."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def test_encoder_avgpool_block_matches_torch(monkeypatch):
    import pytorch3dunet_b200 as P
    from pytorch3dunet_b200 import engine as E
    from oracle import unet3d_oracle as O
    from tests.test_gpu_model import _engine_masks
    torch.manual_seed(0)
    mod = P.Encoder(16, 32, pool_type="avg").cuda()
    x = (torch.rand(1, 16, 9, 10, 12) * 2 - 0.5)
    xe = x.cuda().requires_grad_(True)
    r = torch.randn(1, 32, 4, 5, 6)
    monkeypatch.setattr(E, "DEBUG", {})
    y = mod(xe)
    (y * r.cuda()).sum().backward()
    torch.cuda.synchronize()
    masks = _engine_masks(E)   # gradients at the engine's ReLU pattern (see tests/test_gpu_model.py)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    xo = x.bfloat16().float().requires_grad_(True)
    yo = O.double_conv(F.avg_pool3d(xo, 2), sd, "basic_module.", "gcr", 8, masks=masks)
    (yo * r).sum().backward()
    rep = {"y": rel_l2(y, yo), "dx": rel_l2(xe.grad, xo.grad)}
    for k, p in mod.named_parameters():
        rep[k] = rel_l2(p.grad, sd[k].grad)
    print("avgpool encoder:", {k: f"{v:.2e}" for k, v in rep.items()})
    assert rep["y"] < 1e-2 and all(v < 3e-2 for k, v in rep.items()), rep


@pytest.mark.parametrize("mode,enc_shape,low_shape", [("trilinear", (8, 8, 8), (4, 4, 4)), ("trilinear", (9, 11, 7), (4, 5, 3)),
                                                      ("nearest", (9, 11, 7), (4, 5, 3))])
def test_upcat_join_kernels_match_torch(mode, enc_shape, low_shape):
    """the materialised join (b200_upcat*_fwd / _bwd) alone, through the engine, against F.interpolate + cat and its autograd"""
    from pytorch3dunet_b200 import engine as E
    torch.manual_seed(1)
    dev = torch.device("cuda")
    enc = torch.randn((2, 16) + enc_shape)
    low = torch.randn((2, 24) + low_shape)
    eng = E.Engine(dev, record=True)
    a_enc = eng.input_bf16(enc.to(dev), True)
    a_low = eng.input_bf16(low.to(dev), True)
    cat = eng.upcat(a_enc, a_low, want_stats=True, mode=mode)
    out = eng.to_ncdhw_f32(cat.t)
    g = torch.randn(out.shape)
    eng.grad_from_ncdhw(cat, g.to(dev))
    eng.run_backward()
    g_enc, g_low = eng.to_ncdhw_f32(a_enc.grad).cpu(), eng.to_ncdhw_f32(a_low.grad).cpu()
    sums = eng.sums_of(cat).cpu()
    eo = enc.bfloat16().float().requires_grad_(True)
    lo = low.bfloat16().float().requires_grad_(True)
    ref = torch.cat((eo, F.interpolate(lo, size=enc_shape, mode=mode)), 1)
    (ref * g.bfloat16().float()).sum().backward()
    assert rel_l2(out, ref) < 4e-3
    assert rel_l2(g_enc, eo.grad) < 4e-3 and rel_l2(g_low, lo.grad) < 6e-3
    refb = out.cpu().double()
    assert torch.allclose(sums[..., 0], refb.sum(dim=(2, 3, 4)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(sums[..., 1], (refb * refb).sum(dim=(2, 3, 4)), rtol=1e-4, atol=1e-2)


def test_flat_gradient_buffer_receives_the_engine_gradients():
    import pytorch3dunet_b200 as P
    cfg = dict(name="ResidualUNetSE3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2)
    torch.manual_seed(0)
    m1 = P.get_model(cfg).cuda()
    m2 = P.get_model(cfg).cuda()
    m2.load_state_dict(m1.state_dict())
    flat = P.optim.FlatParameters(m2)
    x = torch.rand(1, 1, 16, 16, 16, device="cuda")
    t = (torch.rand(1, 1, 16, 16, 16, device="cuda") > 0.5).float()
    for m in (m1, m2):
        _, logits = m(x, return_logits=True)
        P.losses.bce_dice_loss(logits, t).backward()
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert p2.grad.data_ptr() == flat.grad_views[k].data_ptr(), k     # still the view of the flat buffer
        # and equal to the autograd-accumulated path (not bit-for-bit: two small backward reductions -- the border-class sums and the
        # fp32-input stem wgrad -- use floating-point atomics, so their summation order varies from run to run at the 1e-7 level)
        assert torch.allclose(p1.grad, p2.grad, rtol=1e-4, atol=1e-7 * float(p1.grad.abs().max()) + 1e-12), k
    # a second step OVERWRITES (no accumulation): same values again
    before = flat.grad.clone()
    _, logits = m2(x, return_logits=True)
    P.losses.bce_dice_loss(logits, t).backward()
    assert torch.allclose(before, flat.grad, rtol=1e-4, atol=1e-7 * float(before.abs().max()))


def test_fused_adam_matches_torch_adam():
    import pytorch3dunet_b200 as P
    cfg = dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2)
    torch.manual_seed(0)
    m1 = P.get_model(cfg).cuda()
    m2 = P.get_model(cfg).cuda()
    m2.load_state_dict(m1.state_dict())
    flat = P.optim.FlatParameters(m2)
    opt1 = torch.optim.Adam(m1.parameters(), lr=2e-3, betas=(0.9, 0.999), weight_decay=1e-5)   # create_optimizer defaults, utils.py:246-316
    opt2 = P.optim.FusedAdam(flat, lr=2e-3, betas=(0.9, 0.999), weight_decay=1e-5)
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(5):
        for p1, p2 in zip(m1.parameters(), m2.parameters()):
            gr = torch.randn(p1.shape, device="cuda", generator=g) * 0.1
            p1.grad = gr.clone()
            p2.grad.copy_(gr)
        opt1.step()
        opt2.step()
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(p1, p2, rtol=2e-5, atol=2e-7), (k, (p1 - p2).abs().max().item())


def test_adam_training_steps_reduce_the_loss():
    """forward + fused loss + backward into the flat buffer + fused Adam: a few steps on one patch must fit it better"""
    import pytorch3dunet_b200 as P
    torch.manual_seed(0)
    m = P.get_model(dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3)).cuda()
    flat = P.optim.FlatParameters(m)
    opt = P.optim.FusedAdam(flat, lr=1e-3)
    x = torch.rand(1, 1, 32, 32, 32, device="cuda")
    t = (x > 0.5).float()
    losses = []
    for _ in range(12):
        opt.zero_grad()
        _, logits = m(x, return_logits=True)
        loss = P.losses.bce_dice_loss(logits, t, fused=True)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses", [f"{v:.4f}" for v in losses])
    assert losses[-1] < 0.8 * losses[0]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_two_replicas_forward_backward():
    """what trainer.py:203-204 / predict.py:63-65 do when several GPUs are visible"""
    import pytorch3dunet_b200 as P
    cfg = dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2)
    torch.manual_seed(0)
    single = P.get_model(cfg).cuda(0)
    dp_model = P.get_model(cfg).cuda(0)
    dp_model.load_state_dict(single.state_dict())
    dp = torch.nn.DataParallel(dp_model, device_ids=[0, 1])
    x = torch.rand(4, 1, 16, 16, 16, device="cuda:0")
    t = (torch.rand(4, 1, 16, 16, 16, device="cuda:0") > 0.5).float()
    out_s, logits_s = single(x, return_logits=True)
    out_d, logits_d = dp(x, return_logits=True)
    assert logits_d.shape == logits_s.shape and logits_d.device == x.device
    assert torch.allclose(logits_d, logits_s, atol=1e-5, rtol=1e-5)   # GroupNorm is per sample: replicas are exact
    F.binary_cross_entropy_with_logits(logits_s, t).backward()
    F.binary_cross_entropy_with_logits(logits_d, t).backward()
    for (k, a), (_, b) in zip(single.named_parameters(), dp_model.named_parameters()):
        assert b.grad is not None, k
        assert rel_l2(b.grad, a.grad) < 2e-2, (k, rel_l2(b.grad, a.grad))  # replica split changes the bf16 summation order only
    with torch.no_grad():
        assert torch.allclose(dp(x), out_s, atol=1e-5)


@pytest.mark.parametrize("phases", [True, False])
@pytest.mark.parametrize("shape", [((2, 16, 12, 10, 8), (2, 32, 6, 5, 4)), ((1, 32, 16, 16, 16), (1, 64, 8, 8, 8)),
                                   ((1, 16, 9, 11, 7), (1, 32, 5, 6, 4))])   # the last one: odd encoder size -> zero-insert path either way
def test_deconv_join_matches_torch(shape, phases, monkeypatch):
    """ConvTranspose3d(k3,s2,p1) + nearest resize + sum join (buildingblocks.py:617-664, :493), through the engine: the phase-decomposed
    path (exact 2x sizes) and the zero-insert path against conv_transpose3d + interpolate and their autograd"""
    from pytorch3dunet_b200 import engine as E
    monkeypatch.setattr(E, "DECONV_PHASES", phases)
    (n, c_out, D, H, W), (_, c_in, d, h, w) = shape
    torch.manual_seed(3)
    dev = torch.device("cuda")
    enc = torch.randn(n, c_out, D, H, W)
    x = torch.randn(n, c_in, d, h, w)
    Wt = torch.randn(c_in, c_out, 3, 3, 3) * (1.0 / (27 * c_in)) ** 0.5
    eng = E.Engine(dev, record=True)
    a_enc = eng.input_bf16(enc.to(dev), True)
    a_x = eng.input_bf16(x.to(dev), True)
    out = eng.deconv_up_add(a_enc, a_x, Wt.to(dev), "up.weight", want_stats=True)
    y = eng.to_ncdhw_f32(out.t)
    g = torch.randn(y.shape)
    eng.grad_from_ncdhw(out, g.to(dev))
    eng.run_backward()
    g_enc, g_x = eng.to_ncdhw_f32(a_enc.grad).cpu(), eng.to_ncdhw_f32(a_x.grad).cpu()
    dWt = eng.param_grads["up.weight"].cpu()
    sums = eng.sums_of(out).cpu()
    eo, xo = enc.bfloat16().float().requires_grad_(True), x.bfloat16().float().requires_grad_(True)
    Wo = Wt.bfloat16().float().requires_grad_(True)
    ref = eo + F.interpolate(F.conv_transpose3d(xo, Wo, None, stride=2, padding=1), size=(D, H, W))
    (ref * g.bfloat16().float()).sum().backward()
    rep = {"y": rel_l2(y, ref), "g_enc": rel_l2(g_enc, eo.grad), "g_x": rel_l2(g_x, xo.grad), "dWt": rel_l2(dWt, Wo.grad)}
    print("deconv join", shape, "phases" if phases else "zero-insert", {k: f"{v:.2e}" for k, v in rep.items()})
    assert rep["y"] < 6e-3 and rep["g_enc"] < 4e-3 and rep["g_x"] < 1e-2 and rep["dWt"] < 1e-2, rep
    refb = y.cpu().double()
    assert torch.allclose(sums[..., 0], refb.sum(dim=(2, 3, 4)), rtol=1e-4, atol=1e-2)
