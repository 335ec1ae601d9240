"""stcd_b200 — B200-native (sm_100a) bi-temporal change-detection inference + evaluation.

A drop-in for ONE hot path of VCISwang/STCD: ``net_G(x1, x2)`` for the Siamese change-detection
nets selected by ``define_G`` (models/networks.py:138-215) and the evaluator's confusion matrix
(train_stcd.py:515-593), executed by hand-written CUDA kernels in ``libstcd_b200.so`` behind the
C-ABI declared in ``include/stcd_b200.h``.  No CPU fallback.
"""
from ._lib import StcdError  # noqa: F401

__all__ = ["StcdError"]
