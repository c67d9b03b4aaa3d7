"""Evaluation branch of ``predict()`` on the device: mirrors ``segmantic.seg.evaluation.confusion_matrix``
(``/root/reference/src/segmantic/seg/evaluation.py:96-125``) and the metrics ``predict()`` reports when labels are
supplied (``seg/monai_unet.py:640-725``: MONAI ``DiceMetric(include_background=False)`` and
``ConfusionMatrixMetric([sensitivity, specificity, precision, accuracy])``).

One kernel (``sgm_confusion_matrix``) reads the two uint8 label maps once; Dice and the confusion-matrix metrics are
closed forms of the ``num_classes x num_classes`` histogram.  No CPU fallback: host arrays are uploaded.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

from .. import _lib


def confusion_matrix(num_classes: int, y_pred, y, device=None) -> np.ndarray:
    """Compute confusion matrix similar to sklearn.metrics.confusion_matrix

    Args:
        num_classes (int): Number of labels including '0', i.e. max(y)+1
        y_pred: Predicted labels (numpy array or CUDA tensor, integer valued)
        y: True labels

    Returns:
        np.ndarray: Dimension num_classes x num_classes (rows: true labels, columns: predictions), int64.
        Pairs with a label outside ``[0, num_classes)`` are not counted.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device available: segmantic_b200 has no CPU fallback")

    def to_dev(a, dev):
        if isinstance(a, torch.Tensor):
            t = a
        else:
            arr = np.asarray(a)
            if arr.dtype.kind == "f":
                arr = arr.astype(np.int64)
            t = torch.from_numpy(np.ascontiguousarray(arr))
        if t.dtype != torch.uint8:
            if t.numel() and (int(t.min()) < 0 or int(t.max()) > 255):
                raise ValueError("labels must lie in [0, 255]")
            t = t.to(torch.uint8)
        return t.reshape(-1).to(dev).contiguous()

    dev = torch.device(device) if device is not None else (
        y_pred.device if isinstance(y_pred, torch.Tensor) and y_pred.is_cuda else
        (y.device if isinstance(y, torch.Tensor) and y.is_cuda else torch.device("cuda:0")))
    p, t = to_dev(y_pred, dev), to_dev(y, dev)
    if p.numel() != t.numel():
        raise ValueError(f"y_pred and y differ in size: {p.numel()} vs {t.numel()}")
    lib = _lib.load()
    cm = torch.empty((num_classes, num_classes), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sgm_confusion_matrix(p.data_ptr(), t.data_ptr(), p.numel(), int(num_classes), cm.data_ptr(),
                                            None, int(torch.cuda.current_stream(dev).cuda_stream)),
                   "sgm_confusion_matrix")
    return cm.cpu().numpy()


def class_dice(cm: np.ndarray, include_background: bool = False) -> np.ndarray:
    """Per-class Dice ``2|A n B| / (|A| + |B|)`` (NaN where the ground truth of the class is empty, as MONAI's
    ``DiceMetric(ignore_empty=True)``); background (class 0) dropped unless asked for."""
    cm = np.asarray(cm, dtype=np.float64)
    tp = np.diag(cm)
    y_o, p_o = cm.sum(1), cm.sum(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        d = np.where(y_o > 0, 2.0 * tp / (y_o + p_o), np.nan)
    return d if include_background else d[1:]


def confusion_counts(cm: np.ndarray) -> np.ndarray:
    """``[C, 4]`` = (tp, fp, tn, fn) per class."""
    cm = np.asarray(cm, dtype=np.float64)
    tp = np.diag(cm)
    fn = cm.sum(1) - tp
    fp = cm.sum(0) - tp
    tn = cm.sum() - tp - fn - fp
    return np.stack([tp, fp, tn, fn], axis=1)


CONFUSION_METRICS = ("sensitivity", "specificity", "precision", "accuracy")


def confusion_metrics(per_image_counts: Sequence[np.ndarray]) -> Dict[str, float]:
    """``ConfusionMatrixMetric(metric_name=CONFUSION_METRICS, reduction="mean").aggregate()``: the counts are averaged
    over classes and images first, the ratios are formed afterwards."""
    f = np.mean(np.stack([np.mean(c, axis=0) for c in per_image_counts]), axis=0)
    tp, fp, tn, fn = (float(v) for v in f)

    def ratio(n, d):
        return float("nan") if d == 0 else n / d

    return {"sensitivity": ratio(tp, tp + fn), "specificity": ratio(tn, tn + fp), "precision": ratio(tp, tp + fp),
            "accuracy": ratio(tp + tn, tp + fp + tn + fn)}
