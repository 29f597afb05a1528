"""GPU: kernels of the residual model family (1x1x1 conv, transposed-conv upsampling + sum-join, scSE) through the C-ABI
against plain PyTorch restatements."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _ctx():
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    from pytorch3dunet_b200._lib import lib
    return U, E, lib()


def _rand(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale + shift).bfloat16()


@pytest.mark.parametrize("cin,cout,f32", [(16, 32, False), (1, 32, True), (64, 128, False), (96, 24, False)])
def test_pointwise_fwd_dgrad_wgrad(cin, cout, f32):
    U, E, L = _ctx()
    N, D, H, W = 2, 3, 5, 7
    vox = D * H * W
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((N, D, H, W, cin), device="cuda", generator=g) if f32 else _rand((N, D, H, W, cin), 2)
    Wt = torch.randn((cout, cin), device="cuda", generator=g) * 0.2
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    y = torch.empty((N, D, H, W, cout), dtype=torch.bfloat16, device="cuda")
    P = L.query("b200_pointwise_partials_count", N, vox, cout)
    part = torch.full((N, P, cout, 2), float("nan"), device="cuda")
    L.call("b200_pointwise_fwd", U.p(x), int(f32), U.p(Wt), 0, U.p(b), N, vox, cin, cout, U.p(y), U.p(part), U.stream())
    ref = x.float() @ Wt.t() + b
    assert U.rel_l2(y, ref) < 5e-3
    yd = y.double()
    assert U.rel_l2(part.double().sum(1)[..., 0], yd.sum((1, 2, 3))) < 1e-4
    assert U.rel_l2(part.double().sum(1)[..., 1], (yd * yd).sum((1, 2, 3))) < 1e-4
    dy = _rand((N, D, H, W, cout), 3)
    Pw = L.query("b200_pointwise_wgrad_partials_count", N, vox)
    K = cout * cin + cout
    wp = torch.full((N * Pw, K), float("nan"), device="cuda")
    L.call("b200_pointwise_wgrad", U.p(x), int(f32), U.p(dy), N, vox, cin, cout, U.p(wp), U.stream())
    red = wp.double().sum(0)
    dW = torch.einsum("ndhwo,ndhwi->oi", dy.double(), x.double())
    assert U.rel_l2(red[:cout * cin].view(cout, cin), dW) < 1e-4
    assert U.rel_l2(red[cout * cin:], dy.double().sum((0, 1, 2, 3))) < 1e-4
    if not f32 and cin % 8 == 0:
        dx = torch.empty((N, D, H, W, cin), dtype=torch.bfloat16, device="cuda")
        L.call("b200_pointwise_fwd", U.p(dy), 0, U.p(Wt), 1, None, N, vox, cout, cin, U.p(dx), None, U.stream())
        assert U.rel_l2(dx, dy.float() @ Wt) < 5e-3


@pytest.mark.parametrize("cin,cout,vox", [(32, 64, 1000), (64, 128, 216), (256, 512, 130), (512, 1024, 27), (16, 16, 4096), (96, 48, 333)])
def test_pointwise_tc_fwd_dgrad_wgrad(cin, cout, vox):
    """1x1x1 conv on the tcgen05 conv / wgrad kernels (flat voxel list) against fp32 matmuls on the same bf16 operands."""
    U, E, L = _ctx()
    N = 2
    assert L.query("b200_pointwise_tc_supported", N, vox, cin, cout)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = _rand((N, vox, cin), 2)
    Wt = torch.randn((cout, cin), device="cuda", generator=g) / cin ** 0.5
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    wq = torch.empty((cout, cin), dtype=torch.bfloat16, device="cuda")
    L.call("b200_pointwise_prep_weights", U.p(Wt), cin, cout, 0, U.p(wq), U.stream())
    assert torch.equal(wq, Wt.bfloat16())
    y = torch.empty((N, vox, cout), dtype=torch.bfloat16, device="cuda")
    P = L.query("b200_pointwise_tc_partials_count", N, vox)
    part = torch.full((N, P, cout, 2), float("nan"), device="cuda")
    L.call("b200_pointwise_tc_fwd", U.p(x), U.p(wq), U.p(b), N, vox, cin, cout, U.p(y), U.p(part), U.stream())
    ref = x.float() @ wq.float().t() + b
    assert U.rel_l2(y, ref) < 4e-3
    yd = y.double()
    assert U.rel_l2(part.double().sum(1)[..., 0], yd.sum(1)) < 1e-4
    assert U.rel_l2(part.double().sum(1)[..., 1], (yd * yd).sum(1)) < 1e-4
    dy = _rand((N, vox, cout), 3)
    wqt = torch.empty((cin, cout), dtype=torch.bfloat16, device="cuda")
    L.call("b200_pointwise_prep_weights", U.p(Wt), cin, cout, 1, U.p(wqt), U.stream())
    assert torch.equal(wqt, Wt.t().bfloat16())
    dx = torch.full((N, vox, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.call("b200_pointwise_tc_fwd", U.p(dy), U.p(wqt), None, N, vox, cout, cin, U.p(dx), None, U.stream())
    assert U.rel_l2(dx, dy.float() @ wq.float()) < 4e-3
    S = L.query("b200_pointwise_tc_wgrad_splits", N, vox, cin, cout)
    G = torch.full((N, S, cin, cout), float("nan"), device="cuda")
    L.call("b200_pointwise_tc_wgrad", U.p(x), U.p(dy), N, vox, cin, cout, U.p(G), U.stream())
    dW = torch.einsum("nvi,nvo->io", x.double(), dy.double())
    assert U.rel_l2(G.double().sum((0, 1)), dW) < 1e-4


@pytest.mark.parametrize("small,big,cin,cout", [((4, 4, 4), (8, 8, 8), 32, 16), ((3, 5, 4), (5, 9, 7), 16, 8), ((2, 2, 2), (4, 4, 4), 64, 32)])
def test_deconv_pieces(small, big, cin, cout):
    """zero-insert / weight permute / resize+add / gather / subsample: each against its torch restatement (bit-exact data movement)."""
    U, E, L = _ctx()
    N = 2
    (d, h, w), (D, H, W) = small, big
    sd, sh, sw = 2 * d - 1, 2 * h - 1, 2 * w - 1
    x = F.relu(_rand((N, d, h, w, cin), 4).float()).bfloat16()
    xz = torch.full((N, sd, sh, sw, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.call("b200_zero_insert", U.p(x), N, d, h, w, cin, U.p(xz), U.stream())
    ref = torch.zeros_like(xz)
    ref[:, ::2, ::2, ::2] = x
    assert torch.equal(xz, ref)
    g = torch.Generator(device="cuda").manual_seed(6)
    Wt = torch.randn((cin, cout, 3, 3, 3), device="cuda", generator=g) * 0.1
    Wc = torch.empty((cout, cin, 3, 3, 3), device="cuda")
    L.call("b200_deconv_weight_permute", U.p(Wt), cin, cout, 1, U.p(Wc), U.stream())
    assert torch.equal(Wc, Wt.flip(2, 3, 4).transpose(0, 1).contiguous())
    back = torch.empty_like(Wt)
    L.call("b200_deconv_weight_permute", U.p(Wc), cin, cout, 0, U.p(back), U.stream())
    assert torch.equal(back, Wt)
    # the identity the engine relies on
    xr = x.float().permute(0, 4, 1, 2, 3)
    t_ref = F.conv_transpose3d(xr, Wt, None, stride=2, padding=1)
    t_conv = F.conv3d(xz.float().permute(0, 4, 1, 2, 3), Wc, None, padding=1)
    assert U.rel_l2(t_conv, t_ref) < 1e-5
    # resize (nearest) + add + statistics
    T = _rand((N, sd, sh, sw, cout), 11)
    enc = _rand((N, D, H, W, cout), 5)
    out = torch.empty((N, D, H, W, cout), dtype=torch.bfloat16, device="cuda")
    P = L.query("b200_resize_add_partials_count", N, D, H, W, cout)
    part = torch.full((N, P, cout, 2), float("nan"), device="cuda")
    L.call("b200_resize_add_fwd", U.p(T), U.p(enc), N, sd, sh, sw, D, H, W, cout, U.p(out), U.p(part), U.stream())
    Tr = T.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    ref = enc.float().permute(0, 4, 1, 2, 3) + F.interpolate(Tr, size=(D, H, W))
    assert torch.equal(out, ref.permute(0, 2, 3, 4, 1).bfloat16())
    od = out.double()
    assert U.rel_l2(part.double().sum(1)[..., 0], od.sum((1, 2, 3))) < 1e-4
    assert U.rel_l2(part.double().sum(1)[..., 1], (od * od).sum((1, 2, 3))) < 1e-4
    dout = _rand((N, D, H, W, cout), 7)
    ref.backward(dout.float().permute(0, 4, 1, 2, 3))
    dT = torch.empty((N, sd, sh, sw, cout), dtype=torch.bfloat16, device="cuda")
    L.call("b200_deconv_gather", U.p(dout), N, d, h, w, D, H, W, cout, U.p(dT), U.stream())
    assert U.rel_l2(dT, Tr.grad.permute(0, 2, 3, 4, 1)) < 4e-3
    # subsample of the zero-inserted gradient, masked by the producer's ReLU, plus an existing gradient
    dxz = _rand((N, sd, sh, sw, cin), 12)
    gadd = _rand((N, d, h, w, cin), 13)
    gx = torch.empty_like(x)
    L.call("b200_subsample2_bwd", U.p(dxz), U.p(x), N, d, h, w, cin, E.ACT_RELU, 0.0, U.p(gadd), U.p(gx), U.stream())
    want = dxz[:, ::2, ::2, ::2].float() * (x.float() > 0) + gadd.float()
    assert U.rel_l2(gx, want) < 4e-3
    L.call("b200_subsample2_bwd", U.p(dxz), U.p(x), N, d, h, w, cin, E.ACT_NONE, 0.0, None, U.p(gx), U.stream())
    assert torch.equal(gx, dxz[:, ::2, ::2, ::2])


@pytest.mark.parametrize("small,big,cin,cout", [((4, 4, 4), (8, 8, 8), 32, 16), ((3, 5, 4), (5, 9, 7), 16, 16), ((6, 6, 6), (12, 12, 12), 128, 64)])
def test_deconv_up_add_engine(small, big, cin, cout):
    """Engine.deconv_up_add (zero-insert -> tcgen05 conv -> resize+add) and its backward against
    conv_transpose3d + interpolate + add in fp32 torch (buildingblocks.py:617-664, :493)."""
    U, E, L = _ctx()
    N = 2
    (d, h, w), (D, H, W) = small, big
    x = F.relu(_rand((N, d, h, w, cin), 4).float()).bfloat16()
    enc = F.relu(_rand((N, D, H, W, cout), 5).float()).bfloat16()
    g = torch.Generator(device="cuda").manual_seed(6)
    Wt = torch.randn((cin, cout, 3, 3, 3), device="cuda", generator=g) * 0.1
    eng = E.Engine(torch.device("cuda"), record=True)
    xa = E.Act(x, E.ACT_RELU, 0.0, requires_grad=True)
    ea = E.Act(enc, E.ACT_RELU, 0.0, requires_grad=True)
    out = eng.deconv_up_add(ea, xa, Wt, "up.weight", want_stats=True)
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    er = enc.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    Wr = Wt.bfloat16().float().requires_grad_(True)
    ref = er + F.interpolate(F.conv_transpose3d(xr, Wr, None, stride=2, padding=1), size=(D, H, W))
    assert U.rel_l2(out.t, ref.permute(0, 2, 3, 4, 1)) < 5e-3
    od = out.t.double()
    sums = eng.sums_of(out)
    assert U.rel_l2(sums[..., 0], od.sum((1, 2, 3))) < 1e-4
    dout = _rand((N, D, H, W, cout), 7)
    ref.backward(dout.float().permute(0, 4, 1, 2, 3))
    out.grad = dout
    eng.run_backward()
    assert U.rel_l2(xa.grad, xr.grad.permute(0, 2, 3, 4, 1) * (x.float() > 0)) < 1e-2
    assert U.rel_l2(ea.grad, er.grad.permute(0, 2, 3, 4, 1) * (enc.float() > 0)) < 1e-2
    assert U.rel_l2(eng.param_grads["up.weight"], Wr.grad) < 1e-2


@pytest.mark.parametrize("C", [32, 64, 256, 512])
def test_scse_fwd_bwd(C):
    U, E, L = _ctx()
    N, D, H, W = 2, 3, 4, 5
    vox = D * H * W
    y = F.relu(_rand((N, D, H, W, C), 8).float()).bfloat16()
    g = torch.Generator(device="cuda").manual_seed(9)
    W1 = torch.randn((C, C), device="cuda", generator=g) / C ** 0.5
    b1 = torch.randn(C, device="cuda", generator=g) * 0.1
    W2 = torch.randn((C, C), device="cuda", generator=g) / C ** 0.5
    b2 = torch.randn(C, device="cuda", generator=g) * 0.1
    ws = torch.randn(C, device="cuda", generator=g) / C ** 0.5
    bs = 0.05
    yd = y.double()
    sums = torch.stack([yd.sum((1, 2, 3)), (yd * yd).sum((1, 2, 3))], -1).contiguous()
    sm = torch.empty((N, C), device="cuda")
    hh, gg = torch.empty_like(sm), torch.empty_like(sm)
    L.call("b200_se_gates_fwd", U.p(sums), float(vox), U.p(W1), U.p(b1), U.p(W2), U.p(b2), N, C, U.p(sm), U.p(hh), U.p(gg), U.stream())
    out = torch.empty_like(y)
    q = torch.empty((N, vox), device="cuda")
    bs_t = torch.tensor([bs], device="cuda")
    L.call("b200_scse_apply_fwd", U.p(y), U.p(gg), U.p(ws), U.p(bs_t), N, vox, C, U.p(out), U.p(q), U.stream())
    # torch reference (fp64) with autograd
    yr = y.double().permute(0, 4, 1, 2, 3).requires_grad_(True)
    P_ = [t.double().requires_grad_(True) for t in (W1, b1, W2, b2, ws)]
    bsr = torch.tensor([bs], dtype=torch.float64, device="cuda", requires_grad=True)
    s = yr.mean((2, 3, 4))
    hr = F.relu(F.linear(s, P_[0], P_[1]))
    gr = torch.sigmoid(F.linear(hr, P_[2], P_[3]))
    qr = torch.sigmoid((yr * P_[4].view(1, C, 1, 1, 1)).sum(1, keepdim=True) + bsr)
    ref = torch.max(yr * gr.view(N, C, 1, 1, 1), yr * qr)
    assert U.rel_l2(gg, gr) < 1e-4
    assert U.rel_l2(q, qr.reshape(N, vox)) < 1e-4
    assert U.rel_l2(out, ref.permute(0, 2, 3, 4, 1)) < 5e-3
    dout = _rand((N, D, H, W, C), 10)
    ref.backward(dout.double().permute(0, 4, 1, 2, 3))
    P = L.query("b200_scse_partials_count", N, vox, C)
    tmp = torch.empty_like(y)
    part = torch.full((N, P, C, 2), float("nan"), device="cuda")
    dbp = torch.full((N * P, 1), float("nan"), device="cuda")
    L.call("b200_scse_bwd1", U.p(dout), U.p(y), U.p(gg), U.p(q), U.p(ws), N, vox, C, U.p(tmp), U.p(part), U.p(dbp), U.stream())
    sums2 = torch.empty((N, C, 2), dtype=torch.float64, device="cuda")
    L.call("b200_partials_finalize", U.p(part), N, P, C, U.p(sums2), U.stream())
    coef = torch.empty((N, C, 3), device="cuda")
    dW1, db1, dW2, db2 = torch.empty_like(W1), torch.empty_like(b1), torch.empty_like(W2), torch.empty_like(b2)
    dws = torch.empty(C, device="cuda")
    scratch = torch.empty((N, 2, C), device="cuda")
    L.call("b200_se_gates_bwd", U.p(sums2), U.p(sm), U.p(hh), U.p(gg), U.p(W1), U.p(W2), N, C, float(vox), U.p(coef), U.p(dW1), U.p(db1),
           U.p(dW2), U.p(db2), U.p(dws), U.p(scratch), U.stream())
    gy = torch.empty_like(y)
    L.call("b200_gn_bwd_apply", U.p(tmp), U.p(y), U.p(coef), N, C, vox, E.ACT_NONE, 0.0, None, U.p(gy), U.stream())
    assert U.rel_l2(gy, yr.grad.permute(0, 2, 3, 4, 1)) < 1e-2
    assert U.rel_l2(dW2, P_[2].grad) < 1e-2 and U.rel_l2(db2, P_[3].grad) < 1e-2
    assert U.rel_l2(dW1, P_[0].grad) < 1e-2 and U.rel_l2(db1, P_[1].grad) < 1e-2
    assert U.rel_l2(dws, P_[4].grad) < 1e-2
    assert U.rel_l2(dbp.double().sum(), bsr.grad) < 1e-2
