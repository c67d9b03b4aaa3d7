"""Ensemble combination on the device (``segmantic.seg.monai_unet.ensemble_creator`` ``:848-1004`` and
``segmantic.seg.transforms.SelectBestEnsemble`` ``seg/transforms.py:15-61``): MONAI ``MeanEnsemble`` /
``VoteEnsemble`` semantics and the reference's select-best rule as voxel-wise CUDA kernels over stacked model outputs
(``sgm_ensemble_mean_argmax`` / ``sgm_ensemble_vote`` / ``sgm_ensemble_select_best``)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from .. import _lib
from .engine import sliding_window_inference


def _st(dev) -> int:
    return int(torch.cuda.current_stream(dev).cuda_stream)


def mean_ensemble_argmax(logits: torch.Tensor, weights: Optional[Sequence[float]] = None, return_mean: bool = False):
    """``logits``: ``[E, C, *spatial]`` float32 CUDA.  Returns uint8 labels ``[*spatial]`` = argmax of
    ``mean_m(x_m * w_m / mean(w))`` (ties -> lowest class), and the mean itself when asked for."""
    if not logits.is_cuda or logits.dtype != torch.float32 or logits.dim() < 3:
        raise ValueError("logits must be a float32 CUDA tensor [E, C, *spatial]")
    logits = logits.contiguous()
    e, c = int(logits.shape[0]), int(logits.shape[1])
    spatial = tuple(logits.shape[2:])
    v = 1
    for s_ in spatial:
        v *= int(s_)
    labels = torch.empty(spatial, dtype=torch.uint8, device=logits.device)
    mean = torch.empty((c,) + spatial, dtype=torch.float32, device=logits.device) if return_mean else None
    w = None
    if weights is not None:
        if len(weights) != e:
            raise ValueError(f"{len(weights)} weights for {e} models")
        w = (C.c_float * e)(*[float(x) for x in weights])
    lib = _lib.load()
    with torch.cuda.device(logits.device):
        _lib.check(lib.sgm_ensemble_mean_argmax(logits.data_ptr(), e, c, v, w, labels.data_ptr(),
                                                mean.data_ptr() if mean is not None else None, _st(logits.device)),
                   "sgm_ensemble_mean_argmax")
    return (labels, mean) if return_mean else labels


def vote_ensemble(labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """``labels``: ``[E, *spatial]`` uint8 CUDA -> majority label (ties -> lowest class)."""
    if not labels.is_cuda or labels.dtype != torch.uint8 or labels.dim() < 2:
        raise ValueError("labels must be a uint8 CUDA tensor [E, *spatial]")
    labels = labels.contiguous()
    out = torch.empty(tuple(labels.shape[1:]), dtype=torch.uint8, device=labels.device)
    lib = _lib.load()
    with torch.cuda.device(labels.device):
        _lib.check(lib.sgm_ensemble_vote(labels.data_ptr(), int(labels.shape[0]), int(num_classes), out.numel(),
                                         out.data_ptr(), _st(labels.device)), "sgm_ensemble_vote")
    return out


def select_best_ensemble(labels: torch.Tensor, pairs: Sequence[Tuple[int, int]]) -> torch.Tensor:
    """``labels``: ``[E, *spatial]`` uint8 CUDA; ``pairs``: ``(tissue id, model index)`` in dictionary order."""
    if not labels.is_cuda or labels.dtype != torch.uint8 or labels.dim() < 2:
        raise ValueError("labels must be a uint8 CUDA tensor [E, *spatial]")
    labels = labels.contiguous()
    out = torch.empty(tuple(labels.shape[1:]), dtype=torch.uint8, device=labels.device)
    n = len(pairs)
    t = (C.c_int32 * max(n, 1))(*[int(p[0]) for p in pairs])
    m = (C.c_int32 * max(n, 1))(*[int(p[1]) for p in pairs])
    lib = _lib.load()
    with torch.cuda.device(labels.device):
        _lib.check(lib.sgm_ensemble_select_best(labels.data_ptr(), int(labels.shape[0]), out.numel(), t, m, n,
                                                out.data_ptr(), _st(labels.device)), "sgm_ensemble_select_best")
    return out


def combine(spec: dict, net_in: torch.Tensor, sw_batch_size: int, precision: str, overlap: float, mode: str):
    """Run every model of ``spec["nets"]`` over ``net_in`` (``[1, Cin, *spatial]``) and combine: returns uint8 labels
    ``[1, 1, *spatial]`` on the network grid."""
    nets = spec["nets"]
    kind = spec["mode"]
    outs = []
    for n in nets:
        eng = n.engine(precision)
        if kind == "mean":
            r = sliding_window_inference(net_in, n.spatial_size, sw_batch_size, eng, overlap=overlap, mode=mode)
            outs.append(r[0])
        else:
            r = sliding_window_inference(net_in, n.spatial_size, sw_batch_size, eng, overlap=overlap, mode=mode,
                                         return_labels=True, return_logits=False)
            outs.append(r["labels"][0, 0])
        eng.check()
    stack = torch.stack(outs)
    if kind == "mean":
        lab = mean_ensemble_argmax(stack, spec.get("weights"))
    elif kind == "vote":
        lab = vote_ensemble(stack, int(spec["num_classes"]))
    elif kind == "select_best":
        lab = select_best_ensemble(stack, spec["pairs"])
    else:
        raise ValueError(f"unknown combination mode {kind!r}")
    return lab[None, None]
