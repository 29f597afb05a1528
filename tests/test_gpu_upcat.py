"""GPU: the "virtual concat" decoder convolution -- conv3(GN(cat(enc, nearest_up2x(b)))) computed as conv3_enc(enc) + per-phase
2x2x2 convolutions of the low-res b (csrc/upcat_conv.cu, b200_conv3_up_*) -- against plain PyTorch on the same operands and
against the engine's own materialised-concat path."""
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


def _ncdhw(t):
    return t.float().permute(0, 4, 1, 2, 3)


SHAPES = [((4, 4, 4), 16, 32, 32), ((3, 5, 6), 16, 16, 16), ((8, 8, 8), 32, 64, 32), ((2, 2, 2), 64, 128, 64), ((16, 16, 16), 32, 64, 32),
          # low-res planes large enough (h >= 18, w >= 10) for the z-stacked phase kernel (csrc/upzs_sm100.cu): N = 4*C_out per instruction
          ((12, 20, 12), 32, 64, 32), ((5, 33, 17), 16, 32, 16), ((9, 18, 10), 16, 128, 64), ((1, 18, 10), 16, 16, 32),
          # c1 % 64 == 0, h >= 17, w >= 10, d >= 2: the stacked-offset weight-gradient kernel (csrc/wgrad_up_sm100.cu); C_out 16 / 128 slices
          ((6, 34, 16), 32, 64, 16), ((4, 17, 10), 16, 64, 128), ((3, 40, 24), 32, 64, 32)]


@pytest.mark.parametrize("small,c0,c1,cout", SHAPES)
def test_up_phase_fwd_dgrad_wgrad(small, c0, c1, cout):
    U, E, L = _ctx()
    N = 2
    d, h, w = small
    D, H, W = 2 * d, 2 * h, 2 * w
    C = c0 + c1
    g = torch.Generator(device="cuda").manual_seed(3)
    Wt = torch.randn((cout, C, 3, 3, 3), device="cuda", generator=g) / (27 * C) ** 0.5
    b = _rand((N, d, h, w, c1), 4)
    assert L.query("b200_conv3_up_supported", N, d, h, w, c1, cout)
    # forward: no GroupNorm, no bias -> wp = phase sums of W (bf16)
    wf_enc = torch.empty((1, 27, cout, c0), dtype=torch.bfloat16, device="cuda")
    wp = torch.empty((1, 64, cout, c1), dtype=torch.bfloat16, device="cuda")
    L.call("b200_gn_fold_upcat", None, None, None, 1, float(D * H * W), U.p(Wt), None, N, c0, c1, cout, U.p(wf_enc), U.p(wp), None, None,
           None, U.stream())
    assert torch.equal(wf_enc[0], Wt[:, :c0].reshape(cout, c0, 27).permute(2, 0, 1).bfloat16())
    R = torch.full((N, D, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.call("b200_conv3_up_phase_fwd", U.p(b), U.p(wp), 1, N, d, h, w, c1, cout, U.p(R), U.stream())
    br = _ncdhw(b).requires_grad_(True)
    Wu = Wt[:, c0:].clone().requires_grad_(True)
    ref = F.conv3d(F.interpolate(br, scale_factor=2, mode="nearest"), Wu, padding=1)
    assert U.rel_l2(R, ref.permute(0, 2, 3, 4, 1)) < 6e-3
    # transpose: gradient w.r.t. the low-res tensor and the weights
    dz = _rand((N, D, H, W, cout), 5)
    ref.backward(_ncdhw(dz))
    wd_enc = torch.empty((27, c0, cout), dtype=torch.bfloat16, device="cuda")
    wd_up = torch.empty((64, c1, cout), dtype=torch.bfloat16, device="cuda")
    L.call("b200_upcat_prep_dgrad_weights", U.p(Wt), c0, c1, cout, U.p(wd_enc), U.p(wd_up), U.stream())
    assert torch.equal(wd_enc, Wt[:, :c0].reshape(cout, c0, 27).flip(2).permute(2, 1, 0).bfloat16())
    dxb = torch.full((N, d, h, w, c1), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.call("b200_conv3_up_dgrad", U.p(dz), U.p(wd_up), N, d, h, w, cout, c1, U.p(dxb), U.stream())
    assert U.rel_l2(dxb, br.grad.permute(0, 2, 3, 4, 1)) < 6e-3
    S2 = L.query("b200_conv3_up_wgrad_splits", N, d, h, w, cout, c1)
    assert S2 > 0
    Q = torch.full((N, S2, 64, cout, c1), float("nan"), device="cuda")
    L.call("b200_conv3_up_wgrad", U.p(dz), U.p(b), N, d, h, w, cout, c1, U.p(Q), U.stream())
    G_enc = torch.zeros((N, 1, 27, c0, cout), device="cuda")
    G = torch.full((N, 27, C, cout), float("nan"), device="cuda")
    L.call("b200_upcat_assemble_wgrad", U.p(G_enc), 1, U.p(Q), S2, N, c0, c1, cout, U.p(G), U.stream())
    dWu = G[:, :, c0:].sum(0).permute(2, 1, 0).reshape(cout, c1, 3, 3, 3)  # [27][c1][co] -> [co][c1][27]
    assert U.rel_l2(dWu, Wu.grad) < 1e-3
    assert float(G[:, :, :c0].abs().max()) == 0.0


@pytest.mark.parametrize("small,c1,cout", [((12, 20, 12), 64, 32), ((9, 18, 10), 128, 64), ((3, 40, 24), 64, 32), ((4, 17, 10), 64, 128)])
@pytest.mark.parametrize("splits", [1, 3])
def test_up_wgrad_stacked_matches_tap_loop(small, c1, cout, splits, monkeypatch):
    """wgrad_up_kernel (offsets as views, M = 2 w-views x 64 channels, N = 4 h-offsets x C_out) against the tap-loop kernel and fp64"""
    U, E, L = _ctx()
    N = 2
    d, h, w = small
    b = _rand((N, d, h, w, c1), 21)
    dz = _rand((N, 2 * d, 2 * h, 2 * w, cout), 22)

    def run():
        S = L.query("b200_conv3_up_wgrad_splits", N, d, h, w, cout, c1)
        assert S > 0
        Q = torch.full((N, S, 64, cout, c1), float("nan"), device="cuda")
        L.call("b200_conv3_up_wgrad", U.p(dz), U.p(b), N, d, h, w, cout, c1, U.p(Q), U.stream())
        torch.cuda.synchronize()
        return Q.double().sum(1), S

    monkeypatch.setenv("B200UNET_WGRAD_SPLITS", str(splits))
    Q1, S1 = run()
    assert S1 == splits
    monkeypatch.delenv("B200UNET_WGRAD_SPLITS")
    monkeypatch.setenv("B200UNET_UP_WGRAD_HS", "0")
    Q0, _ = run()
    # fp64 reference: Q[t] = sum_u dz[2u + t - 1] * b[u]
    dzp = F.pad(dz.double().permute(0, 4, 1, 2, 3), (1, 2, 1, 2, 1, 2))  # index o + 1
    bb = b.double()
    ref = torch.zeros((N, 64, cout, c1), dtype=torch.float64, device="cuda")
    for td in range(4):
        for th in range(4):
            for tw in range(4):
                sub = dzp[:, :, td:td + 2 * d:2, th:th + 2 * h:2, tw:tw + 2 * w:2]  # dz[2u + t - 1]
                ref[:, (td * 4 + th) * 4 + tw] = torch.einsum("ncdhw,ndhwk->nck", sub, bb)
    print("up wgrad", small, c1, cout, splits, "vs fp64", U.rel_l2(Q1, ref), "tap loop vs fp64", U.rel_l2(Q0, ref))
    assert U.rel_l2(Q1, ref) < 1e-4
    assert U.rel_l2(Q1, Q0) < 1e-4
    worst = max(U.rel_l2(Q1[:, t], ref[:, t]) for t in range(64))
    assert worst < 1e-3, worst


@pytest.mark.parametrize("small,c0,c1,cout", SHAPES)
@pytest.mark.parametrize("mode", ["gn", "bias", "plain"])
def test_virtual_concat_conv_matches_materialised(small, c0, c1, cout, mode):
    """Engine.conv3 on a VirtualCat == Engine.conv3 on the materialised concat (forward, statistics, every gradient)."""
    U, E, L = _ctx()
    N = 2
    d, h, w = small
    D, H, W = 2 * d, 2 * h, 2 * w
    C = c0 + c1
    g = torch.Generator(device="cuda").manual_seed(7)
    Wt = torch.randn((cout, C, 3, 3, 3), device="cuda", generator=g) / (27 * C) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1 if mode == "bias" else None
    groups = 8 if C % 8 == 0 else 1
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.2
    # ReLU outputs with a large mean: the case the weight-rounding correction of the bias table exists for
    enc_t = F.relu(_rand((N, D, H, W, c0), 8, 1.0, 1.5).float()).bfloat16()
    low_t = F.relu(_rand((N, d, h, w, c1), 9, 1.0, 1.5).float()).bfloat16()
    # gradient in "dz form" (already masked by the output activation); the same mask for both runs: the two forward results differ
    # by bf16 rounding, and a mask taken from each run's own output would flip a few ReLUs
    dz = (_rand((N, D, H, W, cout), 10).float() * (_rand((N, D, H, W, cout), 11).float() > 0)).bfloat16()

    def run(virtual):
        eng = E.Engine(torch.device("cuda"), record=True)
        enc = E.Act(enc_t, E.ACT_RELU, 0.0, requires_grad=True)
        low = E.Act(low_t, E.ACT_RELU, 0.0, requires_grad=True)
        if virtual:
            x = E.VirtualCat(enc, low)
            assert eng._vcat_conv_ok(x, cout)
        else:
            x = eng._upcat_materialize(enc, low, want_stats=True)
        gn = (gamma, beta, groups, "gn.weight", "gn.bias") if mode == "gn" else None
        y = eng.conv3(x, Wt, bias, gn, "c.", act=(E.ACT_RELU, 0.0), want_stats=True)
        sums = eng.sums_of(y).clone()
        y.grad = dz.clone()
        eng.run_backward()
        return y.t, sums, enc.grad, low.grad, eng.param_grads

    yv, sv, gev, glv, pv = run(True)
    ym, sm, gem, glm, pm = run(False)
    assert U.rel_l2(yv, ym) < 6e-3
    assert U.rel_l2(sv, sm) < 4e-3   # the statistics of two bf16 roundings of the same tensor (R is rounded once more on the virtual path)
    assert U.rel_l2(gev, gem) < 1e-2
    assert U.rel_l2(glv, glm) < 1e-2
    assert set(pv) == set(pm)
    for k in pm:
        scale = float(pm[k].float().norm()) + 1e-6
        assert float((pv[k].float() - pm[k].float()).norm()) / scale < 2e-2, k


@pytest.mark.parametrize("case", [((40, 20, 12), 32, 32, 1), ((40, 20, 12), 32, 32, 3), ((21, 36, 20), 64, 32, 2), ((30, 18, 10), 128, 64, 1)])
def test_zstacked_phase_conv_matches_tap_loop_and_torch(case, monkeypatch):
    """csrc/upzs_sm100.cu with few CTAs per sample (long depth walks: TMEM ring wrap, segment cuts inside a column, channel slices)
    against the tap-loop phase kernel on the same operands and against conv3(nearest_up2x(.)) in fp32"""
    U, E, L = _ctx()
    (d, h, w), c1, cout, cps = case
    N = 2
    g = torch.Generator(device="cuda").manual_seed(9)
    Wu = torch.randn((cout, c1, 3, 3, 3), device="cuda", generator=g) / (27 * c1) ** 0.5
    b = _rand((N, d, h, w, c1), 6)
    # phase weights straight from the definition (per axis: p=0: j=0 <- tap -1, j=1 <- taps 0,+1;  p=1: j=0 <- taps -1,0, j=1 <- tap +1)
    sets = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}
    wp = torch.zeros((1, 64, cout, c1), device="cuda")
    for phase in range(8):
        pp = ((phase >> 2) & 1, (phase >> 1) & 1, phase & 1)
        for j in range(8):
            jj = ((j >> 2) & 1, (j >> 1) & 1, j & 1)
            acc = 0
            for td in sets[(pp[0], jj[0])]:
                for th in sets[(pp[1], jj[1])]:
                    for tw in sets[(pp[2], jj[2])]:
                        acc = acc + Wu[:, :, td, th, tw]
            wp[0, phase * 8 + j] = acc
    wp = wp.bfloat16()
    monkeypatch.setenv("B200UNET_ZS_CTAS", str(cps))
    R = torch.full((N, 2 * d, 2 * h, 2 * w, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.call("b200_conv3_up_phase_fwd", U.p(b), U.p(wp), 1, N, d, h, w, c1, cout, U.p(R), U.stream())
    monkeypatch.setenv("B200UNET_UPZS", "0")
    R0 = torch.full_like(R, float("nan"))
    L.call("b200_conv3_up_phase_fwd", U.p(b), U.p(wp), 1, N, d, h, w, c1, cout, U.p(R0), U.stream())
    torch.cuda.synchronize()
    # reference with the SAME bf16-rounded phase weights is what both kernels compute; against torch: the unrounded 27-tap weights
    ref = F.conv3d(F.interpolate(_ncdhw(b), scale_factor=2, mode="nearest"), Wu, padding=1).permute(0, 2, 3, 4, 1)
    print("upzs", case, "vs tap-loop", U.rel_l2(R, R0.float()), "vs torch", U.rel_l2(R, ref))
    assert not torch.isnan(R.float()).any()
    assert U.rel_l2(R, R0.float()) < 4e-3
    assert U.rel_l2(R, ref) < 8e-3
    # the transpose (gradient w.r.t. the low-res tensor): z-stacked kernel vs the tap-loop one vs autograd
    monkeypatch.delenv("B200UNET_UPZS")
    dz = _rand((N, 2 * d, 2 * h, 2 * w, cout), 7)
    Wfull = torch.cat([torch.zeros((cout, 16, 3, 3, 3), device="cuda"), Wu], 1)   # b200_upcat_prep_dgrad_weights wants (C_out, C0 + C1, 27)
    wd_enc = torch.empty((27, 16, cout), dtype=torch.bfloat16, device="cuda")
    wd_up = torch.empty((64, c1, cout), dtype=torch.bfloat16, device="cuda")
    L.call("b200_upcat_prep_dgrad_weights", U.p(Wfull), 16, c1, cout, U.p(wd_enc), U.p(wd_up), U.stream())
    assert L.query("b200_conv3_up_dgrad_zs_supported", N, d, h, w, cout, c1)
    parts = torch.full((4, N, d, h, w, c1), float("nan"), dtype=torch.bfloat16, device="cuda")
    g1 = torch.full((N, d, h, w, c1), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.call("b200_conv3_up_dgrad_zs", U.p(dz), U.p(wd_up), N, d, h, w, cout, c1, U.p(parts), U.p(g1), U.stream())
    g0 = torch.full_like(g1, float("nan"))
    L.call("b200_conv3_up_dgrad", U.p(dz), U.p(wd_up), N, d, h, w, cout, c1, U.p(g0), U.stream())
    br = _ncdhw(b).requires_grad_(True)
    F.conv3d(F.interpolate(br, scale_factor=2, mode="nearest"), Wu, padding=1).backward(_ncdhw(dz))
    gref = br.grad.permute(0, 2, 3, 4, 1)
    print("updzs", case, "vs tap-loop", U.rel_l2(g1, g0.float()), "vs torch", U.rel_l2(g1, gref))
    assert not torch.isnan(g1.float()).any()
    assert U.rel_l2(g1, g0.float()) < 6e-3      # four bf16 partials are rounded before their sum
    assert U.rel_l2(g1, gref) < 1e-2
