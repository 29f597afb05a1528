"""Golden patch-index fixtures from the reference's own SliceBuilder / halo helpers (run in the build container; the reference's
`datasets` package imports h5py / skimage, which are stubbed -- the index logic itself is pure Python).  Writes
tests/golden/patch_indices.json.  Test infrastructure only."""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def import_reference():
    for name in ("h5py", "skimage", "skimage.color", "skimage.filters", "skimage.segmentation", "imageio"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["h5py"].Dataset = object
    sys.modules["h5py"].File = object
    sys.modules["skimage.color"].label2rgb = lambda *a, **k: None
    sys.modules["skimage.filters"].gaussian = lambda *a, **k: None
    sys.modules["skimage.segmentation"].find_boundaries = lambda *a, **k: None
    sys.modules["skimage"].color = sys.modules["skimage.color"]
    sys.path.insert(0, "/root/reference")
    from pytorch3dunet.datasets import utils as du
    return du


class Vol:
    def __init__(self, shape):
        self.shape, self.ndim = tuple(shape), len(shape)


def main():
    du = import_reference()
    cases = [((256, 512, 512), (128, 128, 128), (64, 64, 64)),      # BASELINE cfg 5: 147 patches
             ((100, 200, 170), (64, 64, 64), (32, 40, 48)),          # ragged tails on every axis
             ((2, 64, 96, 130), (64, 64, 64), (64, 32, 64)),         # 4-D volume (channel slice prepended)
             ((64, 64, 64), (64, 64, 64), (16, 16, 16))]             # a single patch
    out = {"slices": [], "mirror": []}
    for shape, patch, stride in cases:
        sb = du.SliceBuilder(Vol(shape), None, patch, stride)
        out["slices"].append(dict(shape=shape, patch=patch, stride=stride,
                                  starts=[[s.start for s in idx] for idx in sb.raw_slices],
                                  stops=[[s.stop for s in idx] for idx in sb.raw_slices]))
    rng = np.random.default_rng(0)
    for shape, pad in [((5, 6, 7), (2, 1, 3)), ((2, 4, 5, 6), (1, 2, 2)), ((4, 4, 4), (0, 0, 0))]:
        a = rng.integers(0, 1000, size=shape).astype(np.int64)
        p = du.mirror_pad(a, pad)
        out["mirror"].append(dict(shape=shape, pad=pad, input=a.tolist(), padded=p.tolist(),
                                  unpadded_equal=bool(np.array_equal(du.remove_padding(p, pad), a))))
    path = os.path.join(ROOT, "tests", "golden", "patch_indices.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, [len(c["starts"]) for c in out["slices"]])


if __name__ == "__main__":
    main()
