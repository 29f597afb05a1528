"""b200 3D U-Net engine: a B200-native (sm_100a) drop-in for the model path of wolny/pytorch-3dunet.

    from pytorch3dunet_b200 import get_model            # same contract as pytorch3dunet.unet3d.model.get_model
    model = get_model({"name": "UNet3D", "in_channels": 1, "out_channels": 1, "f_maps": 32}).cuda()

    pytorch3dunet_b200.install()                        # rebinds get_model & the model classes inside an installed
                                                        # reference package so train3dunet / predict3dunet use the engine
"""
from .model import (AbstractUNet, Decoder, DoubleConv, Encoder, ResidualUNet3D, ResidualUNetSE3D, ResNetBlock, ResNetBlockSE,  # noqa: F401
                    SingleConv, UNet3D, UnsupportedConfig,
                    get_model, is_model_2d, last_launch_counts, last_tape_length, number_of_features_per_level)
from .install import install, uninstall  # noqa: F401
from . import losses  # noqa: F401
from . import patches  # noqa: F401
from . import optim  # noqa: F401
from . import pipeline  # noqa: F401

__all__ = ["get_model", "UNet3D", "ResidualUNet3D", "ResidualUNetSE3D", "SingleConv", "DoubleConv", "Encoder", "Decoder",
           "install", "uninstall", "is_model_2d", "losses", "patches", "optim", "pipeline", "last_launch_counts", "UnsupportedConfig"]
