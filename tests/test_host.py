"""CPU: host-side logic -- the C-ABI library loads and exports every symbol the header declares, the module tree
carries the reference's parameter names/shapes, unsupported configurations fail loudly (no silent fallback)."""
import ctypes

import pytest
import torch

from tests.helpers import load_golden


def test_library_exports_every_header_symbol():
    from pytorch3dunet_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 40
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, missing
    assert _lib.lib().query("b200_version") >= 1


@pytest.mark.parametrize("golden,cfg", [
    ("unet3d_f16_l3_s16", dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3)),
    ("unet3d_f16_l3_odd", dict(name="UNet3D", in_channels=2, out_channels=3, f_maps=16, num_levels=3, final_sigmoid=False)),
    ("unet3d_f16_l2_cgr", dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, layer_order="cgr")),
    ("resunet3d_f16_l3_s16", dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3)),
    ("resunetse3d_f16_l3_s16", dict(name="ResidualUNetSE3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3)),
    ("unet3d_f16_l2_cl", dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, layer_order="cl")),
    ("unet3d_f16_l2_crg", dict(name="UNet3D", in_channels=1, out_channels=2, f_maps=16, num_levels=2, layer_order="crg", final_sigmoid=False)),
    ("resunet3d_f16_l2_gcl", dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, layer_order="gcl")),
    ("resunetse3d_f16_l2_gce", dict(name="ResidualUNetSE3D", in_channels=2, out_channels=1, f_maps=16, num_levels=2, layer_order="gce")),
    ("unet3d_f16_l2_trilinear", dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, upsample="trilinear")),
    ("unet3d_f16_l2_deconv", dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, upsample="deconv")),
    ("resunet3d_f16_l2_cge", dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, layer_order="cge")),
    ("resunet3d_f16_l2_deconvcat", dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=2, upsample="deconv")),
])
def test_state_dict_contract_matches_reference(golden, cfg):
    import pytorch3dunet_b200 as P
    _, sd, _ = load_golden(golden)
    model = P.get_model(cfg)
    msd = model.state_dict()
    assert list(msd.keys()) == list(sd.keys())  # same names in the same order as the reference's state_dict
    for k in sd:
        assert tuple(msd[k].shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd)


def test_default_init_equals_reference_under_seed():
    """same construction order => same RNG consumption => identical default init as the reference (golden sd minus
    the GroupNorm perturbation make_golden.py adds)"""
    import pytorch3dunet_b200 as P
    _, sd, _ = load_golden("unet3d_f16_l3_s16")
    torch.manual_seed(0)
    model = P.get_model(dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3))
    for k, v in model.state_dict().items():
        if "groupnorm" not in k:
            assert torch.equal(v, sd[k]), k


def test_headline_model_parameter_count():
    import pytorch3dunet_b200 as P
    m = P.get_model(dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=32, num_levels=4))
    assert sum(p.numel() for p in m.parameters()) == 4081267


def test_cpu_tensor_is_rejected_not_silently_computed():
    import pytorch3dunet_b200 as P
    m = P.get_model(dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=8, num_levels=2))
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 1, 8, 8, 8))


def test_unsupported_orders_fail_loudly():
    import pytorch3dunet_b200 as P
    with pytest.raises(NotImplementedError):   # UnsupportedConfig is a NotImplementedError
        P.get_model(dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=8, num_levels=2, layer_order="bcr"))
    with pytest.raises(NotImplementedError):
        P.get_model(dict(name="UNet2D", in_channels=1, out_channels=1))


def test_query_functions_without_gpu():
    from pytorch3dunet_b200._lib import lib
    L = lib()
    assert L.query("b200_conv3_igemm_supported", 2, 128, 128, 128, 96, 32) == 1
    assert L.query("b200_conv3_igemm_supported", 2, 128, 128, 128, 1, 16) == 0
    # z-stacked kernel (resident weights fit): one partial row per persistent CTA, 148 SMs / 2 samples
    assert L.query("b200_conv3_igemm_partials_count", 2, 128, 128, 128, 96, 32) == 148 // 2
    # tap-loop kernel (weights too large to stay resident): one partial row per 128-voxel tile
    assert L.query("b200_conv3_igemm_partials_count", 2, 32, 32, 32, 128, 128) == 32 ** 3 // 128
    assert L.query("b200_maxpool_partials_count", 1, 64, 64, 64, 32) >= 1


def test_planning_entry_points_over_the_model_shapes():
    """Host-only planning code (tile / split / buffer sizing) for every layer shape of the BASELINE configurations, plus ragged
    and tiny volumes: supported shapes must yield positive counts, the tensor-core path must be chosen where it exists."""
    from pytorch3dunet_b200._lib import lib
    from pytorch3dunet_b200 import engine as E
    L = lib()
    layers = []
    for (N, S, f, levels) in [(2, 128, 32, 4), (4, 96, 32, 5), (1, 96, 64, 5), (1, 17, 16, 3)]:
        for lv in range(levels):
            s = max(S >> lv, 1)
            c = f << lv
            layers += [(N, s, s, s, c, c), (N, s, s, s, max(c // 2, 16), c), (N, s, s, s, c + 2 * c, c)]
    layers += [(1, 5, 9, 7, 32, 16), (2, 3, 18, 10, 16, 32), (1, 6, 6, 6, 512, 512), (1, 4, 4, 8, 64, 320)]
    for (N, D, H, W, Cin, Cout) in layers:
        if Cin % 16 or Cout % 16:
            continue
        assert L.query("b200_conv3_igemm_supported", N, D, H, W, Cin, Cout) == 1, (N, D, H, W, Cin, Cout)
        assert L.query("b200_conv3_igemm_partials_count", N, D, H, W, Cin, Cout) >= 1
        assert L.query("b200_conv3_wgrad_igemm_supported", N, D, H, W, Cin, Cout) == 1
        assert L.query("b200_conv3_wgrad_igemm_splits", N, D, H, W, Cin, Cout) >= 1
        assert L.query("b200_border_tap_sums_workspace", N, D, H, W, Cout) > 0
        vox = D * H * W
        assert L.query("b200_pointwise_tc_supported", N, vox, Cin, Cout) == 1
        assert L.query("b200_pointwise_tc_partials_count", N, vox) == (vox + 127) // 128
        assert L.query("b200_pointwise_tc_wgrad_splits", N, vox, Cin, Cout) >= 1
        if D % 2 == 0 and H % 2 == 0 and W % 2 == 0:  # virtual concat: low-res (D/2,H/2,W/2)
            assert L.query("b200_conv3_up_supported", N, D // 2, H // 2, W // 2, Cin, Cout) == 1
            assert L.query("b200_conv3_up_wgrad_splits", N, D // 2, H // 2, W // 2, Cout, Cin) >= 1
    # shapes the tensor-core kernels do not take: "auto" resolves to the direct CUDA-core kernels, "tcgen05" is refused
    # (b200_conv3_resolve_impl also requires an sm_100 device, so without a GPU it never answers tcgen05)
    assert L.query("b200_conv3_igemm_supported", 1, 8, 8, 8, 24, 8) == 0
    assert L.query("b200_conv3_resolve_impl", E.IMPL_AUTO, 1, 8, 8, 8, 24, 8, 0) == E.IMPL_DIRECT
    assert L.query("b200_conv3_resolve_impl", E.IMPL_TCGEN05, 1, 8, 8, 8, 24, 8, 0) < 0
    assert L.query("b200_pointwise_tc_supported", 1, 512, 24, 8) == 0
    assert L.query("b200_conv3_up_supported", 1, 4, 4, 4, 24, 16) == 0


def test_engine_only_calls_declared_entry_points():
    """every `b200_*` name the Python host passes to the C-ABI is declared in include/b200unet.h (a prototype dropped from the header
    would otherwise only surface on the GPU box)"""
    import glob
    import os
    import re
    from pytorch3dunet_b200 import _lib
    protos = set(_lib.parse_header())
    used = set()
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for f in glob.glob(os.path.join(root, "*.py")):
        used |= set(re.findall(r'"(b200_\w+)"', open(f).read()))
    missing = sorted(used - protos - {"b200_stream_t"})
    assert not missing, missing


def test_deferred_groupnorm_backward_is_applied_by_any_reader_of_grad():
    """engine.Act: a skip connection's GroupNorm backward left pending for the max-pool backward (Act.deferred) must not get lost when
    something else reads `.grad` first -- the read applies it (host logic, stub engine)."""
    import torch
    from pytorch3dunet_b200 import engine as E

    class StubEngine:
        def __init__(self):
            self.calls = []

        def gn_bwd_apply(self, dxhat, x, coef, n, c, vox):
            self.calls.append((n, c, vox, x._grad))
            x.grad = dxhat + 1   # stands for the kernel's result

    a = E.Act(torch.zeros(2, 3, 4, 5, 8))
    assert a.grad is None and a.deferred is None and not a.pool_pending
    eng = StubEngine()
    dxhat = torch.full((2, 3, 4, 5, 8), 2.0)
    a.deferred = (dxhat, torch.zeros(2, 8, 3), eng)
    g = a.grad                                   # first reader: applies the pending term exactly once
    assert torch.equal(g, dxhat + 1) and a.deferred is None
    assert eng.calls == [(2, 8, 60, None)]
    assert a.grad is g and len(eng.calls) == 1   # later reads do not re-apply
    a.grad = None
    assert a.grad is None


def test_partials_count_of_new_entry_points_without_gpu():
    """query-style entry points added in round 2 answer on the host (no device work)"""
    from pytorch3dunet_b200._lib import lib
    L = lib()
    assert L.query("b200_maxpool_bwd_partials_count", 2, 16, 16, 16, 32) >= 1
    assert L.query("b200_conv3_direct_wgrad_splits", 2, 16, 16, 16, 1, 16, 1) >= 1      # first conv: one split slot per block
    assert L.query("b200_conv3_direct_wgrad_splits", 2, 16, 16, 16, 24, 16, 0) == 1     # generic CUDA-core fallback
