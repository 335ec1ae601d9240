"""numpy restatement of the evaluator's confusion matrix and scores.  TEST INFRASTRUCTURE.

Follows SegmentationMetric (train_stcd.py:515-593): ``cm[g, p] = #{label == g and pred == p}``
via ``bincount(numClass * label + pred, minlength=numClass**2)`` (:572-579), rows = ground
truth, columns = prediction; scores without eps (NaN when a class is absent) (:523-561).
"""
from __future__ import annotations

import numpy as np


def confusion_matrix(pred: np.ndarray, label: np.ndarray, num_class: int = 2) -> np.ndarray:
    """train_stcd.py:572-579.  int64 [num_class, num_class]."""
    assert pred.shape == label.shape                      # addBatch asserts this (:587)
    idx = num_class * label.astype(np.int64).ravel() + pred.astype(np.int64).ravel()
    return np.bincount(idx, minlength=num_class ** 2).reshape(num_class, num_class).astype(np.int64)


def binarise(logits: np.ndarray, kind: str, thr: float = 0.5) -> np.ndarray:
    """The three binarisations that feed the matrix:
    'argmax'   torch.argmax(G_pred, dim=1)  (models/evaluator.py:108-109; ties -> class 0)
    'sigmoid'  (sigmoid(x) > thr)           (train_stcd.py:477,483), evaluated in fp32 like torch
    'raw_ge'   (x >= thr)                   (models/evaluator.py:110-113)
    logits: [B, C, H, W]; returns uint8 [B, H, W]."""
    x = np.asarray(logits, dtype=np.float32)
    if kind == "argmax":
        return (x[:, 1] > x[:, 0]).astype(np.uint8)
    if kind == "sigmoid":
        import torch
        return (torch.sigmoid(torch.from_numpy(x[:, 0].copy())) > thr).numpy().astype(np.uint8)
    if kind == "raw_ge":
        return (x[:, 0] >= np.float32(thr)).astype(np.uint8)
    raise ValueError(kind)


def scores(cm: np.ndarray) -> dict:
    """train_stcd.py:523-561 (float64, no eps)."""
    cm = cm.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        diag = np.diag(cm)
        precision = diag / cm.sum(0)
        recall = diag / cm.sum(1)
        f1 = 2 * precision * recall / (precision + recall)
        iou = diag / (cm.sum(1) + cm.sum(0) - diag)
        oa = diag.sum() / cm.sum()
    return {"OA": oa, "Precision": precision, "Recall": recall, "F1": f1, "IoU": iou, "mIoU": iou.mean()}


def confuse_matrix_meter_scores(pred: np.ndarray, label: np.ndarray, n_class: int = 2) -> dict:
    """Independent numpy statement of what ``ConfuseMatrixMeter.update_cm`` + ``get_scores`` produce for one batch
    (models/evaluator.py:115,150-167).  PARITY UNPINNED: ``misc/metric_tool.py`` is absent from the reference tree; this
    follows upstream BIT_CD's ``get_confuse_matrix`` (labels outside [0, n_class) masked out) and ``cm2score`` (eps =
    float32 machine epsilon in every quotient, nanmean over classes)."""
    eps = np.finfo(np.float32).eps
    gt, pr = label.ravel().astype(np.int64), pred.ravel().astype(np.int64)
    keep = (gt >= 0) & (gt < n_class)
    hist = np.zeros((n_class, n_class), np.float64)
    for g in range(n_class):
        for p in range(n_class):
            hist[g, p] = np.count_nonzero(keep & (gt == g) & (pr == p))
    tp = np.diag(hist)
    rec = tp / (hist.sum(1) + eps)
    pre = tp / (hist.sum(0) + eps)
    f1 = 2 * rec * pre / (rec + pre + eps)
    iou = tp / (hist.sum(1) + hist.sum(0) - tp + eps)
    out = {"acc": tp.sum() / (hist.sum() + eps), "miou": np.nanmean(iou), "mf1": np.nanmean(f1)}
    for i in range(n_class):
        out.update({f"iou_{i}": iou[i], f"F1_{i}": f1[i], f"precision_{i}": pre[i], f"recall_{i}": rec[i]})
    return out
