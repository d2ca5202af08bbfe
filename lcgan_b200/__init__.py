"""lcgan_b200: B200-native (sm_100a) implementation of the LC-GAN training hot path.

Drop-in modules `cnn`, `custom_layers`, `loss`, `ema` keep the reference's surface
(rakutentech/lcgan cnn.py, custom_layers.py, loss.py, ema.py); `lcgan_b200/dropin/` holds same-named
shims so the reference's own main.py / worker.py import them unchanged.
"""
from . import _lib, plans  # noqa: F401
from .ops import get_precision, no_weight_gradients, set_precision, set_tensor_cores  # noqa: F401

__all__ = ["set_precision", "get_precision", "set_tensor_cores", "no_weight_gradients"]
