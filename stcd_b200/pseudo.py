"""The step right after the hot path in the STCD recipe: pseudo-label masks and the
reliability score between checkpoints, on the GPU.

* ``change_mask(logits, thr)`` -- ``(sigmoid(diffseg) > thr).int()``, ``[== 1] = 255`` (train_stcd.py:176,185;
  thr 0.7 in train_pse_cd.py:145): the uint8 image the script saves as PNG, written by one bandwidth kernel.
* ``ReliabilityScorer`` -- train_stcd.py:104-123: IoU of the change class between each earlier checkpoint's
  prediction and the last one's.  The reference creates ONE ``SegmentationMetric`` before the image loop and never
  resets it (:105), so the "per-image" score is the IoU accumulated over all images seen so far; ``cumulative=True``
  (default) reproduces that, ``cumulative=False`` scores each image on its own.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib
from .metric import SegmentationMetric, _LOGIT_KINDS


@torch.no_grad()
def change_mask(logits: torch.Tensor, thr: float = 0.5, kind: str = "sigmoid", on_value: int = 255) -> torch.Tensor:
    """fp32 logits [B, 1, H, W] ('sigmoid', 'raw_ge') or [B, 2, H, W] ('argmax') -> uint8 [B, H, W] in {0, on_value}."""
    if not logits.is_cuda:
        raise RuntimeError("stcd_b200 has no CPU path: move the logits to a B200 (cuda) device")
    if logits.dtype != torch.float32 or logits.dim() != 4:
        raise TypeError("logits must be float32 [B, C, H, W]")
    if kind not in _LOGIT_KINDS:
        raise ValueError(f"kind must be one of {sorted(_LOGIT_KINDS)}")
    b, c, h, w = logits.shape
    if c != (2 if kind == "argmax" else 1):
        raise ValueError(f"kind={kind} needs {2 if kind == 'argmax' else 1} logit channel(s), got {c}")
    logits = logits.contiguous()
    mask = torch.empty(b, h, w, dtype=torch.uint8, device=logits.device)
    stream = C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)
    _lib.check(_lib.lib().stcd_binarise_mask(C.c_void_p(logits.data_ptr()), _LOGIT_KINDS[kind], float(thr), b, h * w,
                                             int(on_value), C.c_void_p(mask.data_ptr()), stream), "stcd_binarise_mask")
    return mask


class ReliabilityScorer:
    def __init__(self, device="cuda", cumulative: bool = True):
        self.metric = SegmentationMetric(2, device)
        self.cumulative = cumulative

    @torch.no_grad()
    def score(self, masks: Sequence[torch.Tensor]) -> float:
        """masks: the uint8 change masks ({0,1} or {0,255}) of one image from K checkpoints, last = the final model.
        Returns sum_i IoU_change(masks[i], masks[-1]) / (K - 1)   (train_stcd.py:115-123)."""
        if len(masks) < 2:
            raise ValueError("need the predictions of at least two checkpoints")
        ious = []
        for m in masks[:-1]:
            if not self.cumulative:
                self.metric.reset()
            self.metric.addBatch(m, masks[-1], raw_masks=True)       # addBatch(preds[i], preds[-1]): rows = preds[-1]
            ious.append(float(self.metric.IntersectionOverUnion()[1]))
        return sum(ious) / len(ious)
