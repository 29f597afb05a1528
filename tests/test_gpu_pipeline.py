"""GPU: the device-resident sliding-window pipeline (pytorch3dunet_b200.pipeline) against the host
restatement of the reference's predictor loop (mirror_pad -> halo-extended slices -> model -> remove_padding -> sequential
`prediction_array[index] = pred`, predictor.py:148-193): the assembled volume must be BIT-identical when both sides run the same
model, and within 1e-2 of the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_loop(model_fn, vol, patch, stride, halo, c_out):
    from pytorch3dunet_b200 import patches as PT
    padded = PT.mirror_pad(vol, halo)
    idx = PT.build_slices(vol.shape[-3:], patch, stride)
    preds = []
    for i in idx:
        pi = PT.padded_index(i, halo)
        p = padded[(slice(None),) + tuple(pi)] if vol.ndim == 4 else padded[tuple(pi)][None]
        preds.append(model_fn(np.ascontiguousarray(p)[None])[0])
    return PT.assemble_last_writer_wins(preds, idx, (c_out,) + tuple(vol.shape[-3:]), halo)


@pytest.mark.parametrize("shape,patch,stride,halo,cin,cout", [
    ((40, 72, 56), (32, 32, 32), (16, 24, 24), (0, 0, 0), 1, 1),
    ((40, 72, 56), (24, 32, 24), (16, 24, 24), (4, 4, 4), 1, 2),
    ((33, 48, 40), (16, 32, 32), (9, 16, 8), (2, 4, 4), 2, 1),      # ragged grid, 2 input channels
])
def test_volume_predictor_matches_reference_loop(shape, patch, stride, halo, cin, cout):
    import pytorch3dunet_b200 as P
    from oracle import unet3d_oracle as O
    torch.manual_seed(0)
    cfg = dict(name="UNet3D", in_channels=cin, out_channels=cout, f_maps=16, num_levels=2, final_sigmoid=cout == 1)
    model = P.get_model(cfg).cuda().eval()
    rng = np.random.default_rng(0)
    vol = rng.random((cin,) + shape, dtype=np.float32) if cin > 1 else rng.random(shape, dtype=np.float32)

    def engine_fn(p):
        with torch.no_grad():
            return model(torch.from_numpy(p).cuda()).cpu().numpy()

    want = _reference_loop(engine_fn, vol, patch, stride, halo, cout)
    vp = P.pipeline.VolumePredictor(model, patch, stride, halo, slab_planes=8)
    got = np.array(vp.predict(vol))
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert vp.stats["patches_this_rank"] == vp.stats["patches_total"] == len(P.patches.build_slices(shape, patch, stride))
    # a second call reuses the pinned ring / output buffer and gives the same volume
    assert np.array_equal(np.array(vp.predict(vol)), want)

    # patch sharding (patch i -> rank i mod world): the ranks' volumes are disjoint and sum to the single-GPU result
    dvol = torch.from_numpy(vol if vol.ndim == 4 else vol[None]).cuda()
    parts = []
    for r in range(3):
        vpr = P.pipeline.VolumePredictor(model, patch, stride, halo, world=3, rank=r)
        parts.append(vpr.predict_device(dvol).cpu().numpy())
    assert np.array_equal(sum(parts), want)
    assert np.all(sum((p != 0).astype(np.int32) for p in parts) <= 1)

    # and against the CPU oracle (fp32 torch ops = what the reference executes)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}

    def oracle_fn(p):
        with torch.no_grad():
            return O.forward(sd, cfg, torch.from_numpy(p))[0].numpy()

    ref = _reference_loop(oracle_fn, vol, patch, stride, halo, cout)
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print("pipeline vs oracle rel-L2", rel)
    assert rel < 1e-2


def test_patch_gather_is_reflect_padding():
    import pytorch3dunet_b200 as P
    from pytorch3dunet_b200._lib import lib
    rng = np.random.default_rng(1)
    vol = rng.random((2, 10, 12, 9), dtype=np.float32)
    halo = (3, 4, 2)
    padded = P.patches.mirror_pad(vol, halo)
    d = torch.from_numpy(vol).cuda()
    out = torch.empty((2, 10 + 6, 12 + 8, 9 + 4), device="cuda")
    lib().call("b200_patch_gather_f32", d.data_ptr(), 2, 10, 12, 9, -3, -4, -2, 16, 20, 13, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(out.cpu().numpy(), padded)
