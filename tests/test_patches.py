"""CPU: host-side patch geometry (pytorch3dunet_b200.patches) against index lists produced by
the reference's own SliceBuilder / mirror_pad / remove_padding (oracle/make_patch_golden.py -> tests/golden/patch_indices.json)."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "patch_indices.json")))


@pytest.mark.parametrize("case", GOLD["slices"], ids=lambda c: "x".join(map(str, c["shape"])))
def test_build_slices_matches_reference_slicebuilder(case):
    from pytorch3dunet_b200 import patches as PT
    sl = PT.build_slices(case["shape"], case["patch"], case["stride"])
    assert [[s.start for s in idx] for idx in sl] == case["starts"]
    assert [[s.stop for s in idx] for idx in sl] == case["stops"]


def test_cfg5_patch_count_and_sharding():
    from pytorch3dunet_b200 import patches as PT
    sl = PT.build_slices((256, 512, 512), (128, 128, 128), (64, 64, 64))
    assert len(sl) == 147  # 3 x 7 x 7
    parts = [PT.shard_patches(len(sl), 8, r) for r in range(8)]
    assert sorted(sum(parts, [])) == list(range(147)) and max(map(len, parts)) == 19 and min(map(len, parts)) == 18


@pytest.mark.parametrize("case", GOLD["mirror"], ids=lambda c: "x".join(map(str, c["shape"])))
def test_mirror_pad_and_halo_crop_match_reference(case):
    from pytorch3dunet_b200 import patches as PT
    a = np.asarray(case["input"])
    p = PT.mirror_pad(a, case["pad"])
    assert np.array_equal(p, np.asarray(case["padded"]))
    crop = PT.halo_crop(case["pad"]) if sum(case["pad"]) else (...,)
    assert np.array_equal(p[crop], a) and case["unpadded_equal"]


def test_halo_round_trip_reassembles_the_volume():
    """cut a mirror-padded volume into halo-extended patches, 'predict' the identity, crop the halo, write back last-writer-wins:
    the volume must come back exactly (the predictor's contract, predictor.py:148-193 with hdf5.py:16-20)."""
    from pytorch3dunet_b200 import patches as PT
    rng = np.random.default_rng(1)
    vol = rng.standard_normal((70, 90, 80)).astype(np.float32)
    halo = (4, 6, 8)
    padded = PT.mirror_pad(vol, halo)
    idx = PT.build_slices(vol.shape, (32, 40, 32), (24, 25, 32))
    preds = [padded[PT.padded_index(i, halo)][None] for i in idx]      # (C=1, z+2h, y+2h, x+2h)
    out = PT.assemble_last_writer_wins(preds, idx, (1,) + vol.shape, halo)
    assert np.array_equal(out[0], vol)


def test_patch_origins_errors():
    from pytorch3dunet_b200 import patches as PT
    assert PT.patch_origins(10, 4, 3) == [0, 3, 6]
    assert PT.patch_origins(11, 4, 3) == [0, 3, 6, 7]
    with pytest.raises(ValueError):
        PT.patch_origins(3, 4, 1)
