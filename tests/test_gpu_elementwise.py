"""GPU: the HBM-bound kernels of libb200unet.so (through the C-ABI) against plain PyTorch fp32/fp64 restatements."""
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


def test_stats_and_finalize():
    U, E, L = _ctx()
    x = _rand((2, 5, 6, 7, 24), 1, 0.7, 0.4)
    N, D, H, W, C = x.shape
    vox = D * H * W
    P = L.query("b200_stats_partials_count", N, C, vox)
    part = torch.full((N, P, C, 2), float("nan"), device="cuda")
    L.call("b200_stats_ndhwc_bf16", U.p(x), N, C, vox, U.p(part), U.stream())
    sums = torch.empty((N, C, 2), dtype=torch.float64, device="cuda")
    L.call("b200_partials_finalize", U.p(part), N, P, C, U.p(sums), U.stream())
    xd = x.double()
    assert U.rel_l2(sums[..., 0], xd.sum((1, 2, 3))) < 1e-5
    assert U.rel_l2(sums[..., 1], (xd * xd).sum((1, 2, 3))) < 1e-5
    # NCDHW fp32 input statistics
    xin = torch.rand(2, 3, 9, 8, 7, device="cuda")
    P2 = L.query("b200_stats_ncdhw_f32_partials_count", 9 * 8 * 7)
    part2 = torch.full((2, P2, 3, 2), float("nan"), device="cuda")
    L.call("b200_stats_ncdhw_f32", U.p(xin), 2, 3, 9 * 8 * 7, U.p(part2), U.stream())
    s2 = part2.double().sum(1)
    assert U.rel_l2(s2[..., 0], xin.double().sum((2, 3, 4))) < 1e-5
    assert U.rel_l2(s2[..., 1], (xin.double() ** 2).sum((2, 3, 4))) < 1e-5


def test_gn_fold_matches_group_norm_then_conv():
    """fold(a,b into weights + border-class bias) == conv(zero_pad(group_norm(x)))"""
    U, E, L = _ctx()
    N, D, H, W, Cin, Cout, G = 2, 6, 5, 7, 16, 24, 8
    x = _rand((N, D, H, W, Cin), 2, 0.6, 1.5)   # mean >> std: exercises the cancellation correction
    g = torch.Generator(device="cuda").manual_seed(3)
    Wt = torch.randn((Cout, Cin, 3, 3, 3), device="cuda", generator=g) * 0.05
    gamma = 1 + 0.2 * torch.randn(Cin, device="cuda", generator=g)
    beta = 0.2 * torch.randn(Cin, device="cuda", generator=g)
    xd = x.double()
    sums = torch.stack([xd.sum((1, 2, 3)), (xd * xd).sum((1, 2, 3))], dim=-1).contiguous()
    wf = torch.empty((N, 27, Cout, Cin), dtype=torch.bfloat16, device="cuda")
    bc = torch.empty((N, 64, Cout), device="cuda")
    mr = torch.empty((N, G, 2), device="cuda")
    ab = torch.empty((N, Cin, 2), device="cuda")
    L.call("b200_gn_fold", U.p(sums), U.p(gamma), U.p(beta), G, float(D * H * W), U.p(Wt), None, N, Cin, Cout,
           U.p(wf), U.p(bc), U.p(mr), U.p(ab), U.stream())
    y, _ = U.run_conv3(E.IMPL_DIRECT, x, wf, bc)
    xn = F.group_norm(x.float().permute(0, 4, 1, 2, 3), G, gamma, beta, 1e-5)
    ref = F.conv3d(xn, Wt, padding=1).permute(0, 2, 3, 4, 1)
    assert U.rel_l2(y, ref) < 1e-2
    xg = xd.permute(0, 4, 1, 2, 3).reshape(N, G, -1)
    assert U.rel_l2(mr[..., 0], xg.mean(-1)) < 1e-5
    assert U.rel_l2(mr[..., 1], 1 / torch.sqrt(xg.var(-1, unbiased=False) + 1e-5)) < 1e-4


@pytest.mark.parametrize("P,C,G", [(37, 32, 8), (1024, 64, 8), (5, 256, 8), (3, 16, 1)])
def test_fused_stats_coeffs_and_fold_equal_the_four_kernel_chain(P, C, G):
    """b200_gn_stats_coeffs == partials_finalize + gn_coeffs and b200_fold_weights_bias == gn_fold's two fold kernels, bit for bit"""
    U, E, L = _ctx()
    N, Cin, Cout, vox = 2, C, 32, 4096.0
    g = torch.Generator(device="cuda").manual_seed(31)
    partials = torch.rand((N, P, C, 2), device="cuda", generator=g) * 50 + 1.0
    partials[..., 1] = partials[..., 0] ** 2 / 40 + partials[..., 1]   # sum of squares large enough for a positive variance
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.2 * torch.randn(C, device="cuda", generator=g)
    Wt = torch.randn((Cout, Cin, 3, 3, 3), device="cuda", generator=g) * 0.05
    cb = torch.randn(Cout, device="cuda", generator=g)
    sums = torch.empty((N, C, 2), dtype=torch.float64, device="cuda")
    L.call("b200_partials_finalize", U.p(partials), N, P, C, U.p(sums), U.stream())
    wf = torch.empty((N, 27, Cout, Cin), dtype=torch.bfloat16, device="cuda")
    bc = torch.empty((N, 64, Cout), device="cuda")
    mr = torch.empty((N, G, 2), device="cuda")
    ab = torch.empty((N, Cin, 2), device="cuda")
    L.call("b200_gn_fold", U.p(sums), U.p(gamma), U.p(beta), G, vox, U.p(Wt), U.p(cb), N, Cin, Cout, U.p(wf), U.p(bc), U.p(mr), U.p(ab),
           U.stream())
    sums2, mr2, ab2 = torch.full_like(sums, float("nan")), torch.full_like(mr, float("nan")), torch.full_like(ab, float("nan"))
    L.call("b200_gn_stats_coeffs", U.p(partials), N, P, C, U.p(gamma), U.p(beta), G, vox, U.p(sums2), U.p(mr2), U.p(ab2), U.stream())
    assert torch.equal(sums, sums2) and torch.equal(mr, mr2) and torch.equal(ab, ab2)
    wf2 = torch.zeros_like(wf)
    bc2 = torch.full_like(bc, float("nan"))
    L.call("b200_fold_weights_bias", U.p(Wt), U.p(ab2), U.p(cb), U.p(sums2), vox, N, Cin, Cout, U.p(wf2), U.p(bc2), U.stream())
    assert torch.equal(wf, wf2) and torch.equal(bc, bc2)
    # no GroupNorm: one weight copy, conv bias only
    wf1 = torch.empty((1, 27, Cout, Cin), dtype=torch.bfloat16, device="cuda")
    bc1 = torch.empty((1, 64, Cout), device="cuda")
    L.call("b200_gn_fold", None, None, None, 1, vox, U.p(Wt), U.p(cb), N, Cin, Cout, U.p(wf1), U.p(bc1), None, None, U.stream())
    wf3, bc3 = torch.zeros_like(wf1), torch.full_like(bc1, float("nan"))
    L.call("b200_fold_weights_bias", U.p(Wt), None, U.p(cb), None, vox, N, Cin, Cout, U.p(wf3), U.p(bc3), U.stream())
    assert torch.equal(wf1, wf3) and torch.equal(bc1, bc3)


def test_border_tap_sums_and_gn_bwd_sums():
    U, E, L = _ctx()
    N, D, H, W, Cin, Cout = 2, 5, 6, 4, 16, 24
    x = _rand((N, D, H, W, Cin), 4)
    dz = _rand((N, D, H, W, Cout), 5)
    g = torch.Generator(device="cuda").manual_seed(6)
    Wt = torch.randn((Cout, Cin, 3, 3, 3), device="cuda", generator=g) * 0.05
    T = torch.full((N, 27, Cout), float("nan"), device="cuda")
    scratch = torch.empty((L.query("b200_border_tap_sums_workspace", N, D, H, W, Cout),), device="cuda")
    L.call("b200_border_tap_sums", U.p(dz), N, D, H, W, Cout, U.p(T), U.p(scratch), U.stream())
    ones = F.pad(torch.ones((1, 1, D, H, W), device="cuda", dtype=torch.float64), (1, 1, 1, 1, 1, 1))
    dzd = dz.double()
    Tref = torch.zeros((N, 27, Cout), dtype=torch.float64, device="cuda")
    for td in range(3):
        for th in range(3):
            for tw in range(3):
                m = ones[0, 0, td:td + D, th:th + H, tw:tw + W]
                Tref[:, (td * 3 + th) * 3 + tw] = torch.einsum("dhw,ndhwo->no", m, dzd)
    assert U.rel_l2(T, Tref) < 1e-5
    Gd = U.run_wgrad(E.IMPL_DIRECT, x, dz)            # [N,27,Cin,Cout] (fp64 sum over splits)
    Gf = Gd.float().reshape(N, 1, 27, Cin, Cout).contiguous()
    sums2 = torch.empty((N, Cin, 2), dtype=torch.float64, device="cuda")
    L.call("b200_gn_bwd_sums_from_wgrad", U.p(Gf), 1, U.p(T), U.p(Wt), N, Cin, Cout, U.p(sums2), U.stream())
    dxhat = F.conv_transpose3d(dzd.permute(0, 4, 1, 2, 3), Wt.double(), padding=1)
    xd = x.double().permute(0, 4, 1, 2, 3)
    assert U.rel_l2(sums2[..., 0], dxhat.sum((2, 3, 4))) < 1e-4
    assert U.rel_l2(sums2[..., 1], (dxhat * xd).sum((2, 3, 4))) < 1e-4
    # wgrad finalize with a,b
    ab = torch.rand((N, Cin, 2), device="cuda") + 0.5
    dW = torch.empty_like(Wt)
    Gsum = torch.full((N, 27, Cin, Cout), float("nan"), device="cuda")
    L.call("b200_wgrad_finalize", U.p(Gf), N, 1, Cin, Cout, U.p(ab), U.p(T), U.p(dW), U.p(Gsum), U.stream())
    assert U.rel_l2(Gsum, Gd) < 1e-6
    ref = torch.einsum("nc,ntco->oct", ab[..., 0].double(), Gd) + torch.einsum("nc,nto->oct", ab[..., 1].double(), Tref)
    assert U.rel_l2(dW.reshape(Cout, Cin, 27), ref) < 1e-4


def test_gn_backward_coeffs_and_apply_match_autograd():
    U, E, L = _ctx()
    N, D, H, W, C, G = 2, 4, 5, 6, 32, 8
    vox = D * H * W
    x = _rand((N, D, H, W, C), 7, 0.8, 0.5)
    dxhat = _rand((N, D, H, W, C), 8)
    g = torch.Generator(device="cuda").manual_seed(9)
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.2 * torch.randn(C, device="cuda", generator=g)
    xr = x.double().permute(0, 4, 1, 2, 3).requires_grad_(True)
    gm, bt = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    yn = F.group_norm(xr, G, gm, bt, 1e-5)
    dxh = dxhat.double().permute(0, 4, 1, 2, 3)
    (yn * dxh).sum().backward()
    xd = x.double()
    sums = torch.stack([xd.sum((1, 2, 3)), (xd * xd).sum((1, 2, 3))], dim=-1).contiguous()
    mr = torch.empty((N, G, 2), device="cuda")
    ab = torch.empty((N, C, 2), device="cuda")
    L.call("b200_gn_coeffs", U.p(sums), U.p(gamma), U.p(beta), G, float(vox), N, C, U.p(mr), U.p(ab), U.stream())
    dd = dxhat.double()
    sums2 = torch.stack([dd.sum((1, 2, 3)), (dd * xd).sum((1, 2, 3))], dim=-1).contiguous()
    coef = torch.empty((N, C, 3), device="cuda")
    dgamma, dbeta = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    L.call("b200_gn_bwd_coeffs", U.p(sums2), U.p(gamma), U.p(mr), G, float(vox), N, C, U.p(coef), U.p(dgamma), U.p(dbeta), U.stream())
    out = torch.empty_like(x)
    L.call("b200_gn_bwd_apply", U.p(dxhat), U.p(x), U.p(coef), N, C, vox, E.ACT_NONE, 0.0, None, U.p(out), U.stream())
    assert U.rel_l2(dgamma, gm.grad) < 1e-4
    assert U.rel_l2(dbeta, bt.grad) < 1e-4
    assert U.rel_l2(out, xr.grad.permute(0, 2, 3, 4, 1)) < 5e-3
    # with relu mask and an already accumulated gradient
    gadd = _rand((N, D, H, W, C), 10)
    out2 = torch.empty_like(x)
    L.call("b200_gn_bwd_apply", U.p(dxhat), U.p(x), U.p(coef), N, C, vox, E.ACT_RELU, 0.0, U.p(gadd), U.p(out2), U.stream())
    ref2 = xr.grad.permute(0, 2, 3, 4, 1) * (xd > 0) + gadd.double()
    assert U.rel_l2(out2, ref2) < 5e-3


@pytest.mark.parametrize("dims", [(8, 8, 8), (5, 9, 7), (2, 3, 2)])
def test_maxpool_fwd_bwd(dims):
    U, E, L = _ctx()
    N, C = 2, 16
    D, H, W = dims
    x = F.relu(_rand((N, D, H, W, C), 11).float()).bfloat16()   # post-ReLU values incl. ties at 0
    od, oh, ow = D // 2, H // 2, W // 2
    y = torch.empty((N, od, oh, ow, C), dtype=torch.bfloat16, device="cuda")
    P = L.query("b200_maxpool_partials_count", N, D, H, W, C)
    part = torch.full((N, P, C, 2), float("nan"), device="cuda")
    L.call("b200_maxpool_fwd", U.p(x), N, D, H, W, C, U.p(y), U.p(part), U.stream())
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    yr = F.max_pool3d(xr, 2)
    assert torch.equal(y.float(), yr.detach().permute(0, 2, 3, 4, 1))
    assert U.rel_l2(part.double().sum(1)[..., 0], yr.detach().double().sum((2, 3, 4))) < 1e-5
    dp = _rand((N, od, oh, ow, C), 12)
    gadd = _rand((N, D, H, W, C), 13)
    out = torch.empty_like(x)
    L.call("b200_maxpool_bwd", U.p(dp), U.p(x), N, D, H, W, C, E.ACT_RELU, 0.0, U.p(gadd), U.p(out), U.stream())
    yr.backward(dp.float().permute(0, 4, 1, 2, 3))
    ref = (xr.grad.permute(0, 2, 3, 4, 1) * (x.float() > 0) + gadd.float())
    assert U.rel_l2(out, ref) < 5e-3


def test_act_bwd_stats_variant_matches_plain_and_emits_totals():
    U, E, L = _ctx()
    N, D, H, W, C, CS = 2, 5, 6, 7, 24, 40
    g = _rand((N, D, H, W, CS), 41)
    y = F.elu(_rand((N, D, H, W, C), 42).float()).bfloat16()
    gadd = _rand((N, D, H, W, C), 43)
    vox = D * H * W
    ref = torch.empty_like(y)
    L.call("b200_act_bwd", U.p(g), CS, 8, U.p(y), N, C, vox, E.ACT_ELU, 0.0, U.p(gadd), U.p(ref), U.stream())
    P = L.query("b200_stats_partials_count", N, C, vox)
    parts = torch.full((N, P, C, 2), float("nan"), device="cuda")
    out = gadd.clone()   # in place over the accumulated gradient, as Engine.accumulate_grad uses it
    L.call("b200_act_bwd_stats", U.p(g), CS, 8, U.p(y), N, C, vox, E.ACT_ELU, 0.0, U.p(out), U.p(out), U.p(parts), U.stream())
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    od = out.double()
    assert U.rel_l2(parts.double().sum(1)[..., 0], od.sum((1, 2, 3))) < 1e-5
    assert U.rel_l2(parts.double().sum(1)[..., 1], (od * od).sum((1, 2, 3))) < 1e-5


@pytest.mark.parametrize("C", [32, 48, 128])   # C/8 a power of two: lane-pair kernel; 48: one thread per cell
@pytest.mark.parametrize("dims", [(8, 8, 8), (5, 9, 7), (16, 12, 20)])
def test_maxpool_bwd_fused_with_deferred_groupnorm_backward(dims, C, monkeypatch):
    """b200_maxpool_bwd_gn == b200_gn_bwd_apply followed by b200_maxpool_bwd (+ the channel totals of the result), in place over dxhat"""
    U, E, L = _ctx()
    N = 2
    D, H, W = dims
    x = F.relu(_rand((N, D, H, W, C), 21).float()).bfloat16()
    dxhat = _rand((N, D, H, W, C), 22)
    coef = (torch.randn((N, C, 3), device="cuda", generator=torch.Generator(device="cuda").manual_seed(23)) * 0.5).contiguous()
    dp = _rand((N, D // 2, H // 2, W // 2, C), 24)
    vox = D * H * W
    tmp = torch.empty_like(x)
    L.call("b200_gn_bwd_apply", U.p(dxhat), U.p(x), U.p(coef), N, C, vox, E.ACT_RELU, 0.0, None, U.p(tmp), U.stream())
    ref = torch.empty_like(x)
    monkeypatch.setenv("B200UNET_MAXPOOL_PAIR", "0")   # reference route: the one-thread-per-cell kernel
    L.call("b200_maxpool_bwd", U.p(dp), U.p(x), N, D, H, W, C, E.ACT_RELU, 0.0, U.p(tmp), U.p(ref), U.stream())
    monkeypatch.delenv("B200UNET_MAXPOOL_PAIR")
    ref_pair = torch.empty_like(x)
    L.call("b200_maxpool_bwd", U.p(dp), U.p(x), N, D, H, W, C, E.ACT_RELU, 0.0, U.p(tmp), U.p(ref_pair), U.stream())
    assert torch.equal(ref, ref_pair)   # same argmax (ties at the ReLU zeros included), same sums
    P = L.query("b200_maxpool_bwd_partials_count", N, D, H, W, C)
    parts = torch.full((N, P, C, 2), float("nan"), device="cuda")
    out = dxhat.clone()
    L.call("b200_maxpool_bwd_gn", U.p(dp), U.p(x), N, D, H, W, C, E.ACT_RELU, 0.0, U.p(out), U.p(coef), U.p(out), U.p(parts), U.stream())
    torch.cuda.synchronize()
    # the two-pass route rounds the GroupNorm term to 16 bits before the scatter is added; the fused one rounds once
    assert U.rel_l2(out, ref) < 4e-3
    od = out.double()
    assert U.rel_l2(parts.double().sum(1)[..., 0], od.sum((1, 2, 3))) < 1e-5
    assert U.rel_l2(parts.double().sum(1)[..., 1], (od * od).sum((1, 2, 3))) < 1e-5


@pytest.mark.parametrize("dims", [((8, 8, 8), (4, 4, 4)), ((5, 9, 7), (2, 4, 3)), ((3, 3, 3), (1, 1, 1))])
def test_upcat_fwd_bwd(dims):
    U, E, L = _ctx()
    (D, H, W), (d, h, w) = dims
    N, C0, C1 = 2, 16, 24
    enc = _rand((N, D, H, W, C0), 14)
    xs = F.relu(_rand((N, d, h, w, C1), 15).float()).bfloat16()
    cat = torch.empty((N, D, H, W, C0 + C1), dtype=torch.bfloat16, device="cuda")
    P = L.query("b200_upcat_partials_count", N, D, H, W, C0 + C1)
    part = torch.full((N, P, C0 + C1, 2), float("nan"), device="cuda")
    L.call("b200_upcat_fwd", U.p(enc), C0, U.p(xs), C1, N, D, H, W, d, h, w, U.p(cat), U.p(part), U.stream())
    xr = xs.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    up = F.interpolate(xr, size=(D, H, W), mode="nearest")
    ref = torch.cat((enc.float().permute(0, 4, 1, 2, 3), up), 1)
    assert torch.equal(cat.float(), ref.detach().permute(0, 2, 3, 4, 1))
    cd = cat.double()
    assert U.rel_l2(part.double().sum(1)[..., 1], (cd * cd).sum((1, 2, 3))) < 1e-5
    dcat = _rand((N, D, H, W, C0 + C1), 16)
    out = torch.empty_like(xs)
    L.call("b200_upcat_bwd", U.p(dcat), C0, C1, U.p(xs), N, D, H, W, d, h, w, E.ACT_RELU, 0.0, U.p(out), U.stream())
    up.backward(dcat.float().permute(0, 4, 1, 2, 3)[:, C0:])
    refg = xr.grad.permute(0, 2, 3, 4, 1) * (xs.float() > 0)
    assert U.rel_l2(out, refg) < 5e-3


@pytest.mark.parametrize("cout,final", [(1, 1), (3, 2), (2, 0)])
def test_final_conv_fwd_bwd(cout, final):
    U, E, L = _ctx()
    N, D, H, W, C = 2, 4, 5, 6, 16
    vox = D * H * W
    x = F.relu(_rand((N, D, H, W, C), 17).float()).bfloat16()
    g = torch.Generator(device="cuda").manual_seed(18)
    Wt = torch.randn((cout, C), device="cuda", generator=g) * 0.3
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    logits = torch.empty((N, cout, D, H, W), device="cuda")
    probs = torch.empty_like(logits)
    L.call("b200_final_conv_fwd", U.p(x), N, vox, C, U.p(Wt), U.p(b), cout, final, U.p(logits), U.p(probs) if final else None, U.stream())
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    Wr, br = Wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    lr = F.conv3d(xr, Wr.view(cout, C, 1, 1, 1), br)
    assert U.rel_l2(logits, lr) < 1e-5
    if final == 1:
        assert U.rel_l2(probs, torch.sigmoid(lr)) < 1e-5
    elif final == 2:
        assert U.rel_l2(probs, torch.softmax(lr, 1)) < 1e-5
    dl = torch.randn(logits.shape, device="cuda", generator=g)
    lr.backward(dl)
    P = L.query("b200_final_conv_bwd_partials_count", N, vox, C, cout)
    K = cout * C + cout
    part = torch.full((N * P, K), float("nan"), device="cuda")
    dz = torch.empty_like(x)
    L.call("b200_final_conv_bwd", U.p(dl), U.p(x), N, vox, C, U.p(Wt), cout, E.ACT_RELU, 0.0, U.p(dz), U.p(part), U.stream())
    red = torch.empty(K, device="cuda")
    L.call("b200_reduce_rows", U.p(part), N * P, K, U.p(red), U.stream())
    assert U.rel_l2(red[:cout * C].view(cout, C), Wr.grad) < 1e-4
    assert U.rel_l2(red[cout * C:], br.grad) < 1e-4
    assert U.rel_l2(dz, xr.grad.permute(0, 2, 3, 4, 1) * (x.float() > 0)) < 5e-3


def test_layout_roundtrip_and_act_bwd():
    U, E, L = _ctx()
    x = torch.randn(2, 16, 3, 4, 5, device="cuda")
    t = torch.empty((2, 3, 4, 5, 16), dtype=torch.bfloat16, device="cuda")
    L.call("b200_ncdhw_f32_to_ndhwc_bf16", U.p(x), U.p(t), 2, 16, 3, 4, 5, U.stream())
    assert torch.equal(t, x.permute(0, 2, 3, 4, 1).bfloat16())
    back = torch.empty_like(x)
    L.call("b200_ndhwc_bf16_to_ncdhw_f32", U.p(t), U.p(back), 2, 16, 3, 4, 5, U.stream())
    assert torch.equal(back, x.bfloat16().float())
    # act_bwd on a channel slice with an accumulated gradient, ELU
    g = _rand((2, 3, 4, 5, 40), 19)
    y = F.elu(_rand((2, 3, 4, 5, 16), 20).float()).bfloat16()
    gadd = _rand((2, 3, 4, 5, 16), 21)
    out = torch.empty_like(y)
    L.call("b200_act_bwd", U.p(g), 40, 8, U.p(y), 2, 16, 60, E.ACT_ELU, 1.0, U.p(gadd), U.p(out), U.stream())
    yf = y.float()
    ref = g[..., 8:24].float() * torch.where(yf > 0, torch.ones_like(yf), yf + 1) + gadd.float()
    assert U.rel_l2(out, ref) < 5e-3
