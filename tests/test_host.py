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
    with pytest.raises(NotImplementedError):
        P.get_model(dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=8, num_levels=2, layer_order="bcr"))
    with pytest.raises(NotImplementedError):
        P.get_model(dict(name="UNet2D", in_channels=1, out_channels=1))


def test_query_functions_without_gpu():
    from pytorch3dunet_b200._lib import lib
    L = lib()
    assert L.query("b200_conv3_igemm_supported", 2, 128, 128, 128, 96, 32) == 1
    assert L.query("b200_conv3_igemm_supported", 2, 128, 128, 128, 1, 16) == 0
    assert L.query("b200_conv3_igemm_partials_count", 2, 128, 128, 128, 96, 32) == 128 ** 3 // 128
    assert L.query("b200_maxpool_partials_count", 1, 64, 64, 64, 32) >= 1
