"""Oracle: MONAI ``MeanEnsemble`` / ``VoteEnsemble`` and the reference's ``SelectBestEnsemble`` restated with torch CPU ops
(test infrastructure).  Call sites: ``/root/reference/src/segmantic/seg/monai_unet.py:917-1003``; the select-best rule
is the reference's own code, ``seg/transforms.py:40-52`` (unclaimed voxels are 0 here, ``torch.empty`` there).
MONAI is not installed; ``MeanEnsemble.__call__`` (weights: ``img * w / w.mean()``, then ``torch.mean(dim=0)``) and
``VoteEnsemble.__call__`` (``one_hot`` -> ``mean`` over models -> ``argmax``) are restated from MONAI >= 1.0."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def mean_ensemble(logits: torch.Tensor, weights=None) -> torch.Tensor:
    """``[E, C, ...]`` -> ``[C, ...]``."""
    x = logits.float()
    if weights is not None:
        w = torch.as_tensor(weights, dtype=torch.float32).reshape((-1,) + (1,) * (x.dim() - 1))
        x = x * w / w.mean(dim=0, keepdim=True)
    return torch.mean(x, dim=0)


def vote_ensemble(labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """``[E, ...]`` integer labels -> majority label (ties -> lowest class, as argmax of the mean one-hot)."""
    oh = F.one_hot(labels.long().clamp(max=num_classes), num_classes + 1)[..., :num_classes].float()
    return torch.argmax(oh.mean(dim=0), dim=-1)


def select_best_ensemble(labels: torch.Tensor, label_model_dict: dict) -> torch.Tensor:
    out = torch.zeros(labels.shape[1:], dtype=torch.long)
    for tissue_id, model_id in label_model_dict.items():
        out[labels[model_id].long() == tissue_id] = tissue_id
    return out
