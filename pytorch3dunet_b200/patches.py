"""Host-side patch geometry of the reference's sliding-window pipeline -- index logic
only, no device code yet: which patches a volume is cut into, how a halo-padded patch maps back into the output volume and in
which order overlapping patches overwrite each other.  The device-side staging (pinned ring buffers, copy stream) and the
device-resident write-back are the next rows; they will consume exactly these index lists.

Reference semantics mirrored (file:line in the reference checkout):
  SliceBuilder._gen_indices / _build_slices      pytorch3dunet/datasets/utils.py:263-282, :235-261
  _create_padded_indexes (halo)                   pytorch3dunet/datasets/hdf5.py:16-20
  mirror_pad / remove_padding                     pytorch3dunet/datasets/utils.py:518-565
  StandardPredictor write-back (last writer wins) pytorch3dunet/unet3d/predictor.py:148-193
"""
from __future__ import annotations

import numpy as np


def patch_origins(size: int, patch: int, stride: int) -> list[int]:
    """Start offsets along one axis: 0, s, 2s, ... while the patch fits, plus a final patch flush with the end when the regular
    grid does not reach it (so the last two patches may overlap by more than patch - stride)."""
    if size < patch:
        raise ValueError("Sample size has to be bigger than the patch size")
    if stride < 1:
        raise ValueError("stride must be positive")
    out = list(range(0, size - patch + 1, stride))
    if out[-1] + patch < size:
        out.append(size - patch)
    return out


def build_slices(shape, patch_shape, stride_shape) -> list[tuple[slice, ...]]:
    """All patch positions of a (Z,Y,X) or (C,Z,Y,X) volume in the reference's order (z outermost, x innermost); a 4-D volume gets
    the full channel slice prepended."""
    shape = tuple(int(v) for v in shape)
    if len(shape) not in (3, 4):
        raise ValueError("volume must be (Z,Y,X) or (C,Z,Y,X)")
    spatial = shape[-3:]
    zs, ys, xs = (patch_origins(spatial[a], int(patch_shape[a]), int(stride_shape[a])) for a in range(3))
    kz, ky, kx = (int(v) for v in patch_shape)
    out = []
    for z in zs:
        for y in ys:
            for x in xs:
                idx = (slice(z, z + kz), slice(y, y + ky), slice(x, x + kx))
                if len(shape) == 4:
                    idx = (slice(0, shape[0]),) + idx
                out.append(idx)
    return out


def padded_index(index, halo_shape):
    """Index of the halo-extended patch inside the mirror-padded volume: the padded volume is shifted by `halo`, so the patch
    [start, stop) with its halo is [start, stop + 2*halo) there."""
    if sum(halo_shape) == 0:
        return tuple(index)
    return tuple(slice(i.start, i.stop + 2 * h) for i, h in zip(index, halo_shape, strict=True))


def mirror_pad(image: np.ndarray, padding_shape) -> np.ndarray:
    """Reflect-pad the three spatial axes (not the channel axis) by `padding_shape` on both sides."""
    if len(padding_shape) != 3:
        raise ValueError("Padding shape must be specified for each dimension: ZYX")
    if any(p < 0 for p in padding_shape):
        raise ValueError("padding_shape must be non-negative")
    if all(p == 0 for p in padding_shape):
        return image
    pad_width = [(int(p), int(p)) for p in padding_shape]
    if image.ndim == 4:
        pad_width = [(0, 0)] + pad_width
    return np.pad(image, pad_width, mode="reflect")


def halo_crop(halo_shape):
    """Index that removes the halo from the trailing len(halo_shape) axes of a prediction."""
    return (..., *(slice(int(p), -int(p) or None) for p in halo_shape))


def assemble_last_writer_wins(predictions, indices, out_shape, halo_shape=(0, 0, 0), dtype=np.float32) -> np.ndarray:
    """Host restatement of the predictor's write-back: every (C,z,y,x) prediction has its halo removed and is written at its
    (unpadded) index; later patches overwrite earlier ones where they overlap."""
    out = np.zeros(out_shape, dtype=dtype)
    crop = halo_crop(halo_shape) if sum(halo_shape) > 0 else (...,)
    for pred, idx in zip(predictions, indices, strict=True):
        out[(slice(0, out_shape[0]),) + tuple(idx)[-3:]] = np.asarray(pred)[crop]
    return out


def shard_patches(num_patches: int, world: int, rank: int) -> list[int]:
    """Inference partitioning: patch i goes to GPU i mod world."""
    return list(range(rank, num_patches, world))
