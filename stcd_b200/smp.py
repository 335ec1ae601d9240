"""The reference's ``segmentation_models_pytorch`` construction surface for the models this library
accelerates (segmentation_models_pytorch/__init__.py: ``SegCD`` ... ``create_model``).

``import stcd_b200.smp as smp; smp.SegCD("resnet34", encoder_weights=None, classes=1)`` is what
train_stcd.py:637-638 does with the vendored package.
"""
from __future__ import annotations

from .segcd import FFCTLCD, SegCD

__all__ = ["SegCD", "FFCTLCD", "create_model"]

_ARCHS = {"segcd": SegCD, "ffctlcd": FFCTLCD}


def create_model(arch: str, encoder_name: str = "resnet34", encoder_weights=None, in_channels: int = 3, classes: int = 1,
                 **kwargs):
    """segmentation_models_pytorch/__init__.py ``create_model``: same KeyError for an unknown arch."""
    try:
        cls = _ARCHS[arch.lower()]
    except KeyError:
        raise KeyError("Wrong architecture type `{}`. Available options are: {}".format(arch, list(_ARCHS.keys())))
    return cls(encoder_name=encoder_name, encoder_weights=encoder_weights, in_channels=in_channels, classes=classes, **kwargs)
