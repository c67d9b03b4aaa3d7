"""Oracle: the evaluation branch of ``predict()`` restated in numpy (test infrastructure).

Call sites in the reference: ``/root/reference/src/segmantic/seg/monai_unet.py:640-725`` (``DiceMetric(
include_background=False, reduction="mean")``, ``ConfusionMatrixMetric(metric_name=[sensitivity, specificity,
precision, accuracy])``) and ``seg/evaluation.py:96-125`` (``confusion_matrix``, "similar to
sklearn.metrics.confusion_matrix").  MONAI is not installed here; the metric definitions are restated from MONAI >= 1.0
(``monai/metrics/meandice.py``: per class ``2|A n B| / (|A| + |B|)``, NaN when the ground truth of the class is empty,
NaNs ignored by the mean; ``monai/metrics/confusion_matrix.py``: (tp, fp, tn, fn) per image and class, averaged over
classes and images BEFORE the ratio is formed).

The reference's own ``confusion_matrix`` is inconsistent with its docstring (its numba branch loops over
``range(num_classes)`` elements only, its fallback indexes ``cm[y_pred, y]``); the documented convention is restated:
rows = true labels, columns = predictions, every voxel counted.
"""
from __future__ import annotations

import numpy as np


def confusion_matrix(num_classes: int, y_pred: np.ndarray, y: np.ndarray) -> np.ndarray:
    y_pred = np.asarray(y_pred).reshape(-1).astype(np.int64)
    y = np.asarray(y).reshape(-1).astype(np.int64)
    ok = (y >= 0) & (y < num_classes) & (y_pred >= 0) & (y_pred < num_classes)
    cm = np.zeros((num_classes, num_classes), dtype=np.int64)
    np.add.at(cm, (y[ok], y_pred[ok]), 1)
    return cm


def class_dice(cm: np.ndarray, include_background: bool = False) -> np.ndarray:
    """Per-class Dice of one image from its confusion matrix (NaN where the ground truth of the class is empty)."""
    cm = np.asarray(cm, dtype=np.float64)
    tp = np.diag(cm)
    y_o, p_o = cm.sum(1), cm.sum(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        d = np.where(y_o > 0, 2.0 * tp / (y_o + p_o), np.nan)
    return d if include_background else d[1:]


def confusion_counts(cm: np.ndarray) -> np.ndarray:
    """``[C, 4]`` = (tp, fp, tn, fn) per class, as MONAI's ``get_confusion_matrix`` on one-hot label maps."""
    cm = np.asarray(cm, dtype=np.float64)
    tp = np.diag(cm)
    fn = cm.sum(1) - tp
    fp = cm.sum(0) - tp
    tn = cm.sum() - tp - fn - fp
    return np.stack([tp, fp, tn, fn], axis=1)


def confusion_metrics(per_image_counts) -> dict:
    """sensitivity / specificity / precision / accuracy from the counts of all images: mean over images and classes
    of (tp, fp, tn, fn), then the ratios (``ConfusionMatrixMetric(reduction="mean").aggregate()``)."""
    f = np.mean(np.stack([np.mean(c, axis=0) for c in per_image_counts]), axis=0)
    tp, fp, tn, fn = (float(v) for v in f)

    def ratio(n, d):
        return float("nan") if d == 0 else n / d

    return {"sensitivity": ratio(tp, tp + fn), "specificity": ratio(tn, tn + fp), "precision": ratio(tp, tp + fp),
            "accuracy": ratio(tp + tn, tp + fp + tn + fn)}
