"""Rebind the reference's model factory to the b200 engine, in place.

The reference resolves the model through `pytorch3dunet.unet3d.model.get_model` (model.py:361-363), which
`trainer.py:17` and `predict.py:15` import BY NAME (`from ... import get_model`), so both the defining module and the
already-imported names in its callers have to be rebound.  2-D models keep the reference implementation.
"""
from __future__ import annotations

import importlib
import sys


def install(verbose: bool = False) -> bool:
    """Returns True if a reference package was found and patched."""
    try:
        ref_model = importlib.import_module("pytorch3dunet.unet3d.model")
    except Exception:  # reference not importable here: nothing to patch
        return False
    from . import model as m

    ref_get_model = ref_model.get_model

    def get_model(model_config):
        if model_config.get("name") in ("UNet3D", "ResidualUNet3D", "ResidualUNetSE3D"):
            return m.get_model(model_config)
        return ref_get_model(model_config)  # UNet2D / ResidualUNet2D: out of scope, stay on the reference

    ref_model.get_model = get_model
    for cls in ("UNet3D", "ResidualUNet3D", "ResidualUNetSE3D"):
        setattr(ref_model, cls, getattr(m, cls))
    for caller in ("pytorch3dunet.unet3d.trainer", "pytorch3dunet.predict", "pytorch3dunet.train"):
        mod = sys.modules.get(caller)
        if mod is not None and hasattr(mod, "get_model"):
            mod.get_model = get_model
    if verbose:
        print("pytorch3dunet_b200: get_model / UNet3D / ResidualUNet3D / ResidualUNetSE3D now run on the b200 engine")
    return True
