"""Rebind the reference's model factory to the b200 engine, in place.

The reference resolves the model through `pytorch3dunet.unet3d.model.get_model` (model.py:361-363), which
`trainer.py:17` and `predict.py:15` import BY NAME (`from ... import get_model`), so both the defining module and the
already-imported names in its callers have to be rebound.

Graph-level fallback: configurations the engine does not build -- 2-D models, BatchNorm / Dropout
layer orders, `upsample` in {None, 'area', ...}, channel counts that are not multiples of 8 -- construct the REFERENCE's own
class instead (never a Python re-implementation of its arithmetic).  The engine validates a configuration in the model
constructor (`model.UnsupportedConfig`), so the decision is made inside get_model(), not on the first batch.
"""
from __future__ import annotations

import importlib
import sys

ENGINE_MODELS = ("UNet3D", "ResidualUNet3D", "ResidualUNetSE3D")
_STATE = {}


def install(verbose: bool = False) -> bool:
    """Returns True if a reference package was found and patched (idempotent)."""
    try:
        ref_model = importlib.import_module("pytorch3dunet.unet3d.model")
    except Exception:  # reference not importable here: nothing to patch
        return False
    from . import model as m

    if _STATE.get("module") is ref_model:
        return True
    ref_get_model = ref_model.get_model
    ref_classes = {name: getattr(ref_model, name) for name in ENGINE_MODELS}

    def get_model(model_config):
        name = model_config.get("name")
        if name in ENGINE_MODELS:
            try:
                return m.get_model(model_config)
            except m.UnsupportedConfig as e:
                if verbose:
                    print(f"pytorch3dunet_b200: {name} config not built in the engine ({e}); using the reference class")
                cfg = dict(model_config)
                return ref_classes[name](**cfg)   # what the reference's get_model does (model.py:362-363)
        return ref_get_model(model_config)  # UNet2D / ResidualUNet2D: out of scope, stay on the reference

    get_model.b200_reference_get_model = ref_get_model
    get_model.b200_reference_classes = ref_classes
    ref_model.get_model = get_model
    for cls in ENGINE_MODELS:
        setattr(ref_model, cls, getattr(m, cls))
    for caller in ("pytorch3dunet.unet3d.trainer", "pytorch3dunet.predict", "pytorch3dunet.train"):
        mod = sys.modules.get(caller)
        if mod is not None and hasattr(mod, "get_model"):
            mod.get_model = get_model
    _STATE.update(module=ref_model, get_model=ref_get_model, classes=ref_classes)
    if verbose:
        print("pytorch3dunet_b200: get_model / UNet3D / ResidualUNet3D / ResidualUNetSE3D now run on the b200 engine")
    return True


def uninstall() -> bool:
    """Undo install() (tests)."""
    ref_model = _STATE.get("module")
    if ref_model is None:
        return False
    current = ref_model.get_model
    ref_model.get_model = _STATE["get_model"]
    for name, cls in _STATE["classes"].items():
        setattr(ref_model, name, cls)
    for caller in ("pytorch3dunet.unet3d.trainer", "pytorch3dunet.predict", "pytorch3dunet.train"):
        mod = sys.modules.get(caller)
        if mod is not None and getattr(mod, "get_model", None) is current:
            mod.get_model = _STATE["get_model"]
    _STATE.clear()
    return True
