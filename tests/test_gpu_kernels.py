"""GPU: each conv kernel of libb200unet.so (through the C-ABI) against a plain PyTorch fp32 restatement of its
contract, and the tcgen05 kernels against the direct CUDA-core kernels on identical operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

TOL = 1e-2  # north_star: within 1e-2 relative (fp32 reference), bf16 operands / fp32 accumulation


def _mk(N, D, H, W, Cin, Cout, n_w, seed, bias=True):
    from tests import gpu_util as U  # noqa
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn((N, D, H, W, Cin), device="cuda", generator=g) * 0.7 + 0.3).bfloat16()
    wf = (torch.randn((n_w, 27, Cout, Cin), device="cuda", generator=g) * (2.0 / (27 * Cin)) ** 0.5).bfloat16()
    b = torch.randn((n_w, 64, Cout), device="cuda", generator=g) * 0.2 if bias else None
    return x, wf, b


SHAPES = [
    # N, D, H, W, Cin, Cout
    (1, 8, 8, 8, 16, 32),
    (2, 8, 8, 16, 32, 32),
    (1, 8, 16, 8, 64, 64),
    (2, 4, 8, 8, 96, 32),
    (1, 8, 8, 8, 128, 128),
    (1, 4, 8, 8, 128, 256),
    (1, 4, 4, 8, 384, 128),
    (1, 5, 9, 7, 32, 16),      # ragged: exercises TMA out-of-bounds fill and masked stores
    (1, 16, 16, 16, 32, 64),
    (2, 4, 4, 4, 64, 64),      # volume smaller than any 128-voxel box: the TMA box overhangs the tensor
    (1, 6, 6, 6, 128, 128),
    # large enough for the halo kernel (D>=3, H>=18, W>=10): resident weights + shifted halo views
    (2, 3, 18, 10, 16, 32),
    (1, 4, 20, 12, 32, 32),
    (1, 5, 33, 17, 96, 32),    # ragged, three weight blocks, 16-channel halo chunks
    (1, 4, 32, 16, 32, 96),
    (1, 4, 20, 12, 64, 32),
    (1, 4, 20, 12, 32, 64),
    (2, 6, 24, 24, 32, 16),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_direct_conv_matches_torch(shape):
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    N, D, H, W, Cin, Cout = shape
    x, wf, b = _mk(N, D, H, W, Cin, Cout, N, 1)
    y, sums = U.run_conv3(E.IMPL_DIRECT, x, wf, b, act=E.ACT_RELU, want_stats=True)
    ref = U.conv3_contract_ref(x, wf, b, act=E.ACT_RELU)
    torch.cuda.synchronize()
    assert U.rel_l2(y, ref) < TOL
    yf = y.double()
    assert U.rel_l2(sums[..., 0], yf.sum(dim=(1, 2, 3))) < 1e-4
    assert U.rel_l2(sums[..., 1], (yf * yf).sum(dim=(1, 2, 3))) < 1e-4


@pytest.mark.parametrize("shape", SHAPES)
def test_tcgen05_conv_matches_torch_and_direct(shape):
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    N, D, H, W, Cin, Cout = shape
    x, wf, b = _mk(N, D, H, W, Cin, Cout, N, 2)
    y, sums = U.run_conv3(E.IMPL_TCGEN05, x, wf, b, act=E.ACT_RELU, want_stats=True)
    torch.cuda.synchronize()
    ref = U.conv3_contract_ref(x, wf, b, act=E.ACT_RELU)
    yd, _ = U.run_conv3(E.IMPL_DIRECT, x, wf, b, act=E.ACT_RELU)
    assert U.rel_l2(y, ref) < TOL
    assert U.rel_l2(y, yd.float()) < 5e-3
    yf = y.double()
    assert U.rel_l2(sums[..., 0], yf.sum(dim=(1, 2, 3))) < 1e-4
    assert U.rel_l2(sums[..., 1], (yf * yf).sum(dim=(1, 2, 3))) < 1e-4


# the z-stacked kernel (csrc/conv_zs_sm100.cu): depth taps stacked along the MMA's N dimension, output planes in a TMEM ring.
# Few CTAs per sample force long walks along the depth axis: ring wrap-around (R = 512 / C_out blocks), segments that start /
# end in the middle of a column (halo planes with a single target block), several columns per CTA.
ZS_CASES = [
    # N, D, H, W, Cin, Cout, ctas per sample
    (1, 40, 36, 20, 32, 32, 1),
    (1, 40, 36, 20, 32, 32, 3),
    (1, 40, 36, 20, 32, 32, 7),
    (2, 21, 18, 10, 16, 32, 2),
    (1, 19, 33, 17, 64, 48, 2),     # ragged tiles, N = 144, ring of 10 blocks
    (1, 12, 20, 12, 32, 64, 1),     # N = 192, ring of 8 blocks
    (1, 9, 18, 10, 32, 80, 1),      # C_out = 80: the TMEM ring would be too short for two lanes -> halo / tap-loop kernel
    (1, 1, 18, 10, 16, 16, 1),      # a single plane
    (1, 2, 20, 12, 96, 32, 2),      # three 32-channel chunks per plane
    (1, 70, 18, 10, 32, 16, 1),     # N = 48, ring of 16 blocks wraps four times
    (2, 20, 36, 20, 64, 64, 2),     # resident weights of 64 output channels do not fit: two 32-channel slices (grid.z = 2)
    (1, 10, 18, 26, 32, 128, 1),    # two 64-channel slices, N = 192
]


@pytest.mark.parametrize("case", ZS_CASES)
def test_zstacked_conv_matches_torch_and_direct(case, monkeypatch):
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    from pytorch3dunet_b200._lib import lib
    N, D, H, W, Cin, Cout, cps = case
    monkeypatch.setenv("B200UNET_ZS_CTAS", str(cps))
    tiles = D * ((H + 15) // 16) * ((W + 7) // 8)
    P = lib().query("b200_conv3_igemm_partials_count", N, D, H, W, Cin, Cout)
    # one partial row per persistent CTA when the z-stacked kernel takes the layer (resident weights + >= 3 halo stages fit shared
    # memory); otherwise the halo / tap-loop kernels' one row per tile
    assert P == min(cps, (tiles + 1) // 2) or P >= tiles // 2, (P, tiles)   # one row per persistent CTA (z-stacked) or per tile
    x, wf, b = _mk(N, D, H, W, Cin, Cout, N, 11)
    res = (torch.randn((N, D, H, W, Cout), device="cuda") * 0.3).bfloat16()
    y, sums = U.run_conv3(E.IMPL_TCGEN05, x, wf, b, act=E.ACT_LEAKY, slope=0.1, residual=res, want_stats=True)
    torch.cuda.synchronize()
    ref = U.conv3_contract_ref(x, wf, b, act=E.ACT_LEAKY, slope=0.1, residual=res)
    yd, _ = U.run_conv3(E.IMPL_DIRECT, x, wf, b, act=E.ACT_LEAKY, slope=0.1, residual=res)
    print("zs", case, "vs torch", U.rel_l2(y, ref), "vs direct", U.rel_l2(y, yd.float()))
    assert U.rel_l2(y, ref) < TOL
    assert U.rel_l2(y, yd.float()) < 5e-3
    yf = y.double()
    assert U.rel_l2(sums[..., 0], yf.sum(dim=(1, 2, 3))) < 1e-4
    assert U.rel_l2(sums[..., 1], (yf * yf).sum(dim=(1, 2, 3))) < 1e-4
    # same result whatever the work split (fp32 accumulation order inside a voxel does not depend on it)
    monkeypatch.setenv("B200UNET_ZS_CTAS", str(cps + 1))
    y2, _ = U.run_conv3(E.IMPL_TCGEN05, x, wf, b, act=E.ACT_LEAKY, slope=0.1, residual=res)
    assert torch.equal(y, y2)
    # and the 27-separate-taps kernels (halo / tap-loop) agree with it
    monkeypatch.setenv("B200UNET_ZS", "0")
    y3, _ = U.run_conv3(E.IMPL_TCGEN05, x, wf, b, act=E.ACT_LEAKY, slope=0.1, residual=res)
    assert U.rel_l2(y, y3.float()) < 5e-3


def test_tcgen05_conv_residual_no_bias_shared_weights():
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    x, wf, _ = _mk(2, 8, 8, 8, 32, 32, 1, 3, bias=False)
    res = torch.randn((2, 8, 8, 8, 32), device="cuda").bfloat16()
    y, _ = U.run_conv3(E.IMPL_TCGEN05, x, wf, None, act=E.ACT_LEAKY, slope=0.1, residual=res)
    ref = U.conv3_contract_ref(x, wf, None, act=E.ACT_LEAKY, slope=0.1, residual=res)
    assert U.rel_l2(y, ref) < TOL


WG_SHAPES = [(1, 8, 8, 8, 16, 32), (2, 8, 8, 8, 32, 32), (1, 8, 8, 8, 64, 64), (1, 4, 8, 8, 96, 32), (1, 8, 8, 8, 128, 128),
             (1, 4, 8, 8, 192, 64), (1, 4, 4, 8, 384, 128), (1, 4, 4, 8, 128, 256), (1, 5, 9, 7, 32, 16), (1, 16, 16, 16, 32, 32),
             (2, 4, 4, 4, 64, 64), (1, 6, 6, 6, 128, 128),
             # large enough for the halo / stacked-tap wgrad kernel (D>=3, H>=18, W>=10)
             (2, 3, 18, 10, 16, 32), (1, 4, 20, 12, 32, 32), (1, 5, 33, 17, 96, 32), (1, 4, 32, 16, 32, 96), (1, 4, 20, 12, 64, 64),
             (1, 3, 18, 10, 128, 128), (1, 3, 20, 10, 32, 256), (2, 6, 24, 24, 32, 16),
             # C_out > 256: processed in output-channel slices (plain and stacked-tap kernels)
             (1, 6, 6, 6, 256, 512), (1, 4, 4, 8, 64, 320), (1, 3, 18, 10, 32, 512), (2, 6, 6, 6, 512, 512)]


# the h-stacked wgrad kernel (wgrad_hs_kernel): three dh taps stacked along N through line-shifted views of the dz tile
HS_CASES = [
    # N, D, H, W, Cin, Cout, splits per sample
    (1, 5, 36, 20, 32, 32, 1),      # every tile accumulated by one CTA
    (1, 5, 36, 20, 32, 32, 4),
    (2, 4, 33, 17, 16, 32, 2),      # ragged tiles (zero-filled x rows / dz lines)
    (1, 3, 18, 10, 64, 64, 1),      # C_out = 64: one depth tap per CTA (grid.z = 3), two C_in slices
    (1, 6, 20, 12, 32, 16, 3),
    (1, 4, 40, 24, 96, 32, 2),
]


@pytest.mark.parametrize("case", HS_CASES)
def test_hstacked_wgrad_matches_torch(case, monkeypatch):
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    N, D, H, W, Cin, Cout, splits = case
    monkeypatch.setenv("B200UNET_WGRAD_SPLITS", str(splits))
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn((N, D, H, W, Cin), device="cuda", generator=g).bfloat16()
    dz = torch.randn((N, D, H, W, Cout), device="cuda", generator=g).bfloat16()
    G = U.run_wgrad(E.IMPL_TCGEN05, x, dz)
    torch.cuda.synchronize()
    ref = U.wgrad_contract_ref(x, dz)
    monkeypatch.setenv("B200UNET_WGRAD_HS", "0")
    G0 = U.run_wgrad(E.IMPL_TCGEN05, x, dz)   # the 9-accumulator halo kernel on the same operands
    print("hs", case, "vs torch", U.rel_l2(G, ref), "vs halo kernel", U.rel_l2(G, G0))
    assert U.rel_l2(G, ref) < 1e-3
    assert U.rel_l2(G, G0) < 1e-4
    worst_tap = max(U.rel_l2(G[:, t], ref[:, t]) for t in range(27))
    assert worst_tap < 2e-3, worst_tap


@pytest.mark.parametrize("shape", WG_SHAPES)
def test_direct_wgrad_matches_torch(shape):
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    N, D, H, W, Cin, Cout = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((N, D, H, W, Cin), device="cuda", generator=g).bfloat16()
    dz = torch.randn((N, D, H, W, Cout), device="cuda", generator=g).bfloat16()
    G = U.run_wgrad(E.IMPL_DIRECT, x, dz)
    assert U.rel_l2(G, U.wgrad_contract_ref(x, dz)) < 1e-3


@pytest.mark.parametrize("shape", WG_SHAPES)
def test_tcgen05_wgrad_matches_torch(shape):
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    N, D, H, W, Cin, Cout = shape
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn((N, D, H, W, Cin), device="cuda", generator=g).bfloat16()
    dz = torch.randn((N, D, H, W, Cout), device="cuda", generator=g).bfloat16()
    G = U.run_wgrad(E.IMPL_TCGEN05, x, dz)
    torch.cuda.synchronize()
    assert U.rel_l2(G, U.wgrad_contract_ref(x, dz)) < 1e-3


@pytest.mark.parametrize("shift,group_rows", [(0, 8), (1, 8), (3, 8), (8, 8), (0, 10), (1, 10), (2, 10), (11, 10)])
def test_probe_umma_row_shifted_swizzled_view(shift, group_rows):
    """hardware fact the halo-reuse conv design depends on (probe kernels: tools/probes/umma_probe.cu, their own small library)"""
    from tests import gpu_util as U
    from tools.probes import probe_lib
    if not probe_lib.available():
        pytest.skip("tools/probes/libb200probe.so not built")
    lib = probe_lib.ProbeLib
    rows = 200
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.randn((rows, 64), device="cuda", generator=g).bfloat16()
    B = torch.randn((16, 64), device="cuda", generator=g).bfloat16()
    D = torch.zeros((128, 16), device="cuda")
    lib().call("b200_probe_umma_rowshift", U.p(A), rows, U.p(B), shift, group_rows, U.p(D), U.stream())
    torch.cuda.synchronize()
    r = torch.arange(128, device="cuda")
    idx = shift + (r // 8) * group_rows + (r % 8)
    ref = A[idx].float() @ B.float().t()
    print("probe shift", shift, "group_rows", group_rows, "rel", U.rel_l2(D, ref))
    assert U.rel_l2(D, ref) < 1e-3


@pytest.mark.parametrize("shape", [(2, 6, 9, 35, 1, 16), (1, 5, 8, 40, 2, 8), (1, 4, 17, 33, 1, 32), (1, 16, 16, 16, 3, 16),
                                   (1, 3, 5, 70, 1, 8), (2, 7, 20, 64, 1, 16), (1, 2, 8, 8, 1, 32), (1, 1, 1, 1, 1, 16)])
def test_stem_kernels_fp32_input(shape):
    """network stem: fp32 NDHWC input with C_in <= 4 (C_in == 1: warp-level MMA with x split into hi + lo halves, stem_mma.cu;
    C_in 2..4: CUDA-core kernels); both HBM-bound"""
    from tests import gpu_util as U
    from pytorch3dunet_b200 import engine as E
    N, D, H, W, Cin, Cout = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand((N, D, H, W, Cin), device="cuda", generator=g)
    wf = (torch.randn((N, 27, Cout, Cin), device="cuda", generator=g) * 0.2).bfloat16()
    b = torch.randn((N, 64, Cout), device="cuda", generator=g) * 0.2
    y, sums = U.run_conv3(E.IMPL_AUTO, x, wf, b, act=E.ACT_RELU, want_stats=True)
    ref = U.conv3_contract_ref(x, wf, b, act=E.ACT_RELU)
    assert U.rel_l2(y, ref) < 5e-3
    yf = y.double()
    assert U.rel_l2(sums[..., 0], yf.sum(dim=(1, 2, 3))) < 1e-4
    assert U.rel_l2(sums[..., 1], (yf * yf).sum(dim=(1, 2, 3))) < 1e-4
    dz = torch.randn((N, D, H, W, Cout), device="cuda", generator=g).bfloat16()
    G = U.run_wgrad(E.IMPL_AUTO, x, dz)
    assert U.rel_l2(G, U.wgrad_contract_ref(x, dz)) < 1e-4
    # no atomics anywhere on the stem path: bit-reproducible
    y2, sums2 = U.run_conv3(E.IMPL_AUTO, x, wf, b, act=E.ACT_RELU, want_stats=True)
    assert torch.equal(y, y2) and torch.equal(sums, sums2)
    assert torch.equal(G, U.run_wgrad(E.IMPL_AUTO, x, dz))
