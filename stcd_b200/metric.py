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

Out-of-range classes.  The reference's ``bincount(numClass * label + pred, minlength=numClass**2).reshape(numClass,
numClass)`` (train_stcd.py:572-579) RAISES when a label or prediction lies outside ``[0, numClass)`` (a raw {0, 255}
mask, a 255 ignore label: the histogram grows and the reshape fails; a negative value fails in bincount).  The kernels
skip such pixels, so this class keeps the number of pixels it was given and every read of the matrix
(``confusionMatrix`` and the score getters) checks it against the matrix total and raises ``ValueError`` when they
differ: the scores are never computed on a silently reduced pixel set.  ``ConfuseMatrixMeter`` below mirrors the other
evaluator world (models/evaluator.py), whose upstream helper masks out-of-range LABELS on purpose.
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
        self._pixels = 0            # pixels handed to addBatch / addLogits (all ranks after allreduce)
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
        self._pixels = 0

    def genConfusionMatrix(self, imgPredict: torch.Tensor, imgLabel: torch.Tensor) -> torch.Tensor:
        """train_stcd.py:572-579: the confusion matrix of ONE batch (int64 [numClass, numClass] on the CPU), without
        touching the running matrix.  Raises like the reference when a class index is out of range."""
        one = SegmentationMetric(self.numClass, self._cm.device)
        one.addBatch(imgPredict, imgLabel)
        return one._checked_counts().to("cpu")

    def _checked_counts(self) -> torch.Tensor:
        cm = self._cm.reshape(self.numClass, self.numClass)
        total = int(cm.sum().item())
        seen = int(self._pixels)
        if total != seen:
            raise ValueError(f"{seen - total} of {seen} pixels carry a label or prediction outside [0, {self.numClass}) "
                             "(e.g. a raw {0, 255} mask passed without raw_masks / raw_mask_label, or an ignore label): the "
                             "reference's bincount(...).reshape(numClass, numClass) raises on these (train_stcd.py:572-579)")
        return cm

    @property
    def confusionMatrix(self) -> torch.Tensor:
        """float64 [numClass, numClass] on the CPU, as the reference keeps it (Appendix G-7)."""
        return self._checked_counts().to("cpu", torch.float64)

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
            self._pixels += pred.numel()
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
        self._pixels += n

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
        self._pixels += b * h * w

    def allreduce(self, group=None):
        """Sum the integer matrix over the data-parallel ranks (NCCL on GPUs): the only collective
        of the path.  A no-op without an initialised process group."""
        from .parallel import allreduce_confusion
        self._pixels = allreduce_confusion(self._cm, group, pixels=self._pixels)
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


# ------------------------------------------------------------------------------------------ the other evaluator world
_EPS = float(torch.finfo(torch.float32).eps)      # np.finfo(np.float32).eps


def cm2score(hist) -> dict:
    """Score dict of ``ConfuseMatrixMeter.get_scores`` (consumed at models/evaluator.py:150-167): ``acc, miou, mf1`` and
    per class ``iou_k, F1_k, precision_k, recall_k``; rows = ground truth, columns = prediction, every quotient carries
    ``+ eps`` (eps = float32 machine epsilon).  PARITY UNPINNED: ``misc/metric_tool.py`` is imported by the reference
    (models/evaluator.py:7) but absent from its tree; this restates upstream BIT_CD's helper of the same name."""
    import numpy as np
    hist = np.asarray(hist, dtype=np.float64)
    n_class = hist.shape[0]
    tp = np.diag(hist)
    sum_a1 = hist.sum(axis=1)
    sum_a0 = hist.sum(axis=0)
    acc = tp.sum() / (hist.sum() + _EPS)
    recall = tp / (sum_a1 + _EPS)
    precision = tp / (sum_a0 + _EPS)
    f1 = 2 * recall * precision / (recall + precision + _EPS)
    iu = tp / (sum_a1 + sum_a0 - tp + _EPS)
    out = {"acc": acc, "miou": np.nanmean(iu), "mf1": np.nanmean(f1)}
    out.update({f"iou_{i}": iu[i] for i in range(n_class)})
    out.update({f"F1_{i}": f1[i] for i in range(n_class)})
    out.update({f"precision_{i}": precision[i] for i in range(n_class)})
    out.update({f"recall_{i}": recall[i] for i in range(n_class)})
    return out


class ConfuseMatrixMeter:
    """Drop-in for ``misc.metric_tool.ConfuseMatrixMeter`` as ``CDEvaluator`` / ``CDTrainer`` use it
    (models/evaluator.py:34,115,152,170; models/trainer.py:205): ``update_cm(pr, gt, weight=1)`` adds one batch and returns
    the batch's mean F1, ``get_scores()`` returns the score dict of the accumulated matrix, ``clear()`` resets.

    PARITY UNPINNED (see ``cm2score``).  The batch matrix comes from the same histogram kernel as SegmentationMetric; like
    upstream's ``get_confuse_matrix`` pixels whose LABEL is outside ``[0, n_class)`` are ignored (an ignore label is
    legal here), while an out-of-range PREDICTION raises (upstream's reshape fails on it).  ``pr`` / ``gt`` may be the
    numpy arrays the reference passes (``.cpu().numpy()``) or device tensors (no host round trip)."""

    def __init__(self, n_class: int, device="cuda"):
        self.n_class = int(n_class)
        self._metric = SegmentationMetric(self.n_class, device)
        self._batch = torch.zeros_like(self._metric._cm)
        self.clear()

    def clear(self):
        self.initialized = False
        self.sum = None          # float64 [n_class, n_class] numpy: weighted sum of the batch matrices (upstream's AverageMeter)
        self.count = 0
        self.val = None

    @staticmethod
    def _as_tensor(x, device):
        if isinstance(x, torch.Tensor):
            return x
        import numpy as np
        return torch.from_numpy(np.ascontiguousarray(x)).to(device, non_blocking=True)

    def update_cm(self, pr, gt, weight=1):
        import numpy as np
        m = self._metric
        dev = m._cm.device
        pr, gt = self._as_tensor(pr, dev), self._as_tensor(gt, dev)
        if pr.dtype not in _PRED_KINDS:
            pr = pr.to(torch.int64)
        if gt.dtype not in _LABEL_KINDS:
            gt = gt.to(torch.int64)
        before = m._cm.clone()
        m.addBatch(pr, gt)
        val = (m._cm - before).reshape(self.n_class, self.n_class).cpu().numpy().astype(np.float64)
        in_range = int(((gt >= 0) & (gt < self.n_class)).sum().item())
        if int(val.sum()) != in_range:
            raise ValueError(f"{in_range - int(val.sum())} predictions lie outside [0, {self.n_class})")
        m._pixels = int(m._cm.sum().item())           # ignored labels are legal in this world
        self.val = val
        self.sum = val * weight if self.sum is None else self.sum + val * weight
        self.count += weight
        self.initialized = True
        return float(cm2score(val)["mf1"])

    @property
    def avg(self):
        return None if self.sum is None else self.sum / self.count

    def get_scores(self) -> dict:
        return cm2score(self.sum)
