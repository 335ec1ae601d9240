"""The evaluator's confusion matrix on the GPU, behind the reference's ``SegmentationMetric`` API.

Mirror of ``SegmentationMetric`` (train_stcd.py:515-593; byte-identical copies in train_sup.py
and train_pse_cd.py): same constructor, ``addBatch(imgPredict, imgLabel)`` (asserts equal
shapes, :586-588), ``confusionMatrix`` (float64 ``[numClass, numClass]``, rows = ground truth,
columns = prediction, :576-578) and the score getters (:523-570, no eps: NaN when a class is
absent).  The counts themselves are accumulated as int64 on the device by the warp-aggregated
histogram kernel ``stcd_confusion_add_batch`` (csrc/aux_kernels.cuh) — the reference moves the
prediction to the CPU and runs a single-threaded ``torch.bincount`` per batch (:484).

``addLogits`` additionally fuses the binarisation that precedes ``addBatch`` in the reference's
loops (``sigmoid(x) > thr``: train_stcd.py:477,483; ``argmax``: models/evaluator.py:108-109;
``x >= thr``: models/evaluator.py:110-113) so the logits are read once and no prediction
tensor is materialised.  ``allreduce()`` sums the matrix over data-parallel ranks (the path's
only collective).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

PRED_ARGMAX2, PRED_SIGMOID_GT, PRED_RAW_GE, PRED_U8, PRED_I32, PRED_I64, PRED_U8_GE1 = range(7)
LABEL_I64, LABEL_U8, LABEL_I32, LABEL_U8_GE1 = range(4)

_PRED_KINDS = {torch.uint8: PRED_U8, torch.bool: PRED_U8, torch.int32: PRED_I32, torch.int64: PRED_I64}
_LABEL_KINDS = {torch.int64: LABEL_I64, torch.uint8: LABEL_U8, torch.bool: LABEL_U8, torch.int32: LABEL_I32}
_LOGIT_KINDS = {"argmax": PRED_ARGMAX2, "sigmoid": PRED_SIGMOID_GT, "raw_ge": PRED_RAW_GE}


def _stream_ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class SegmentationMetric(nn.Module):
    def __init__(self, numClass: int, device="cuda"):
        super().__init__()
        self.numClass = int(numClass)
        self.device = device
        self.count = 0
        self._cm: Optional[torch.Tensor] = None
        self.reset(device)

    # ------------------------------------------------------------------ state
    def reset(self, device=None):
        """train_stcd.py:590-593.  The accumulator lives on the device as int64 (exact; the
        reference's float64 is exact only below 2**53)."""
        dev = torch.device(device if device is not None else self.device)
        if not torch.cuda.is_available():
            raise _lib.StcdError("stcd_b200.SegmentationMetric needs a CUDA (B200) device: there is no CPU path")
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())   # reference callers pass 'cpu'/'cuda:0' freely
        self._cm = torch.zeros(self.numClass * self.numClass, dtype=torch.int64, device=dev)

    @property
    def confusionMatrix(self) -> torch.Tensor:
        """float64 [numClass, numClass] on the CPU, as the reference keeps it (Appendix G-7)."""
        return self._cm.reshape(self.numClass, self.numClass).to("cpu", torch.float64)

    def confusion_counts(self) -> torch.Tensor:
        """int64 [numClass, numClass] device tensor (no sync)."""
        return self._cm.reshape(self.numClass, self.numClass)

    def getConfusionMatrix(self):
        return self.confusionMatrix

    # ------------------------------------------------------------------ accumulation
    def _to_dev(self, t: torch.Tensor) -> torch.Tensor:
        if not t.is_cuda:
            t = t.to(self._cm.device, non_blocking=True)
        return t.contiguous()

    def addBatch(self, imgPredict: torch.Tensor, imgLabel: torch.Tensor, raw_masks: bool = False):
        """train_stcd.py:586-588.  raw_masks=True: both arguments are uint8 mask images ({0, 255}, the pseudo-label
        PNGs of train_stcd.py:185-196), binarised on the fly (>= 1 -> class 1) -- numClass must be 2."""
        assert imgPredict.shape == imgLabel.shape
        pred = self._to_dev(imgPredict)
        label = self._to_dev(imgLabel)
        if raw_masks:
            if pred.dtype != torch.uint8 or label.dtype != torch.uint8 or self.numClass != 2:
                raise TypeError("raw_masks needs two uint8 mask images and numClass == 2")
            _lib.check(_lib.lib().stcd_confusion_add_batch(
                C.c_void_p(pred.data_ptr()), PRED_U8_GE1, 0.0, C.c_void_p(label.data_ptr()), LABEL_U8_GE1, 1, pred.numel(), 2,
                C.c_void_p(self._cm.data_ptr()), None, _stream_ptr(pred)), "stcd_confusion_add_batch")
            self.count += 1
            return
        if pred.dtype not in _PRED_KINDS or label.dtype not in _LABEL_KINDS:
            raise TypeError(f"unsupported dtypes pred={pred.dtype} label={label.dtype}")
        if pred.dtype == torch.bool:
            pred = pred.view(torch.uint8)
        if label.dtype == torch.bool:
            label = label.view(torch.uint8)
        n = pred.numel()
        _lib.check(_lib.lib().stcd_confusion_add_batch(
            C.c_void_p(pred.data_ptr()), _PRED_KINDS[pred.dtype], 0.0, C.c_void_p(label.data_ptr()),
            _LABEL_KINDS[label.dtype], 1, n, self.numClass, C.c_void_p(self._cm.data_ptr()), None,
            _stream_ptr(pred)), "stcd_confusion_add_batch")
        self.count += 1

    def addLogits(self, logits: torch.Tensor, imgLabel: torch.Tensor, kind: str = "sigmoid", thr: float = 0.5,
                  pred_out: Optional[torch.Tensor] = None, raw_mask_label: bool = False):
        """Fused binarise + histogram.  logits: fp32 [B,2,H,W] ('argmax') or [B,1,H,W] ('sigmoid',
        'raw_ge'); imgLabel: [B,H,W] or [B,1,H,W] integer; pred_out: optional uint8 [B,H,W].
        raw_mask_label=True: imgLabel is the raw uint8 mask image ({0, 255}), binarised on the fly like the
        reference's loader does (``label[label >= 1] = 1``, data/dataset.py:206-210)."""
        if self.numClass != 2:
            raise ValueError("addLogits binarises: numClass must be 2")
        if kind not in _LOGIT_KINDS:
            raise ValueError(f"kind must be one of {sorted(_LOGIT_KINDS)}")
        logits = self._to_dev(logits)
        label = self._to_dev(imgLabel)
        if logits.dtype != torch.float32 or logits.dim() != 4:
            raise TypeError("logits must be float32 [B,C,H,W]")
        b, c, h, w = logits.shape
        if c != (2 if kind == "argmax" else 1):
            raise ValueError(f"kind={kind} needs {2 if kind == 'argmax' else 1} logit channel(s), got {c}")
        assert label.numel() == b * h * w
        if label.dtype == torch.bool:
            label = label.view(torch.uint8)
        if label.dtype not in _LABEL_KINDS:
            raise TypeError(f"unsupported label dtype {label.dtype}")
        label_kind = _LABEL_KINDS[label.dtype]
        if raw_mask_label:
            if label.dtype != torch.uint8:
                raise TypeError("raw_mask_label needs the uint8 mask image")
            label_kind = LABEL_U8_GE1
        po = None
        if pred_out is not None:
            assert pred_out.is_cuda and pred_out.dtype == torch.uint8 and pred_out.numel() == b * h * w
            po = C.c_void_p(pred_out.data_ptr())
        _lib.check(_lib.lib().stcd_confusion_add_batch(
            C.c_void_p(logits.data_ptr()), _LOGIT_KINDS[kind], float(thr), C.c_void_p(label.data_ptr()),
            label_kind, b, h * w, 2, C.c_void_p(self._cm.data_ptr()), po, _stream_ptr(logits)),
            "stcd_confusion_add_batch")
        self.count += 1

    def allreduce(self, group=None):
        """Sum the integer matrix over the data-parallel ranks (NCCL on GPUs): the only collective
        of the path.  A no-op without an initialised process group."""
        from .parallel import allreduce_confusion
        allreduce_confusion(self._cm, group)
        return self

    # ------------------------------------------------------------------ scores (train_stcd.py:523-570)
    def OverallAccuracy(self):
        cm = self.confusionMatrix
        return torch.diag(cm).sum() / cm.sum()

    def Precision(self):
        cm = self.confusionMatrix
        return torch.diag(cm) / cm.sum(0)

    def Recall(self):
        cm = self.confusionMatrix
        return torch.diag(cm) / cm.sum(1)

    def F1score(self):
        p, r = self.Precision(), self.Recall()
        return 2 * p * r / (p + r)

    def IntersectionOverUnion(self):
        cm = self.confusionMatrix
        inter = torch.diag(cm)
        return inter / (cm.sum(1) + cm.sum(0) - inter)

    def meanIntersectionOverUnion(self):
        return torch.mean(self.IntersectionOverUnion())

    def Frequency_Weighted_Intersection_over_Union(self):
        cm = self.confusionMatrix
        freq = cm.sum(1) / (cm.sum() + 1e-8)
        iu = torch.diag(cm) / (cm.sum(1) + cm.sum(0) - torch.diag(cm) + 1e-8)
        return (freq[freq > 0] * iu[freq > 0]).sum()
