"""Oracle: MONAI ``sliding_window_inference`` restated (test infrastructure).

Call site in the reference: ``SlidingWindowInferer(roi_size=net.spatial_size, sw_batch_size=4,
device=device)`` at ``/root/reference/src/segmantic/seg/monai_unet.py:637-639`` invoked at ``:665``
(defaults overlap=0.25, mode="constant", sigma_scale=0.125, padding_mode="constant", cval=0);
``overlap=0.5`` variant at ``:840-842``.  Algorithm restated from MONAI >= 1.2
``monai/inferers/utils.py``; MONAI is not installed here.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Sequence

import torch
import torch.nn.functional as F


def get_scan_interval(image_size, roi_size, overlap: float):
    out = []
    for i, r in zip(image_size, roi_size):
        if r == i:
            out.append(int(r))
        else:
            out.append(max(int(r * (1 - overlap)), 1))
    return tuple(out)


def dense_patch_starts(image_size, roi_size, scan_interval):
    """Per-dimension window starts (MONAI ``dense_patch_slices``): last window shifted back inside."""
    starts = []
    for size, roi, interval in zip(image_size, roi_size, scan_interval):
        if interval == 0:
            num = 1
        else:
            num = int(math.ceil(float(size) / interval))
            for d in range(num):
                if d * interval + roi >= size:
                    num = d + 1
                    break
        dim_starts = []
        for i in range(num):
            s = i * interval
            s -= max(s + roi - size, 0)
            dim_starts.append(s)
        starts.append(dim_starts)
    return starts


def window_starts(image_size, roi_size, overlap: float):
    """All windows, first spatial dim slowest (``meshgrid(indexing='ij')`` flattened)."""
    interval = get_scan_interval(image_size, roi_size, overlap)
    per_dim = dense_patch_starts(image_size, roi_size, interval)
    return list(itertools.product(*per_dim))


def gaussian_1d(n: int, sigma_scale: float = 0.125) -> torch.Tensor:
    sigma = sigma_scale * n
    x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float32)
    return torch.exp(x ** 2 / (-2 * sigma ** 2))


def importance_map(patch_size: Sequence[int], mode: str = "constant", sigma_scale: float = 0.125):
    """MONAI >= 1.2 ``compute_importance_map`` (separable Gaussian, floor at max(min, 1e-3))."""
    if mode == "constant":
        return torch.ones(tuple(patch_size), dtype=torch.float32)
    if mode != "gaussian":
        raise ValueError(mode)
    imap = None
    for i, n in enumerate(patch_size):
        g = gaussian_1d(n, sigma_scale)
        imap = g if imap is None else imap.unsqueeze(-1) * g[(None,) * i]
    min_non_zero = max(float(imap.min()), 1e-3)
    return torch.clamp(imap.to(torch.float32), min=min_non_zero)


def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: Callable[[torch.Tensor], torch.Tensor],
                             overlap: float = 0.25, mode: str = "constant",
                             sigma_scale: float = 0.125, return_count: bool = False):
    """inputs ``[N=1, Cin, *spatial]`` -> ``[1, C, *spatial]`` blended logits (fp32)."""
    nd = inputs.dim() - 2
    roi_size = tuple(int(r) for r in roi_size)[:nd]
    image_size_ = tuple(inputs.shape[2:])
    batch = inputs.shape[0]
    image_size = tuple(max(image_size_[i], roi_size[i]) for i in range(nd))
    pad_size = []
    for k in range(nd - 1, -1, -1):  # F.pad wants last dim first
        diff = max(roi_size[k] - image_size_[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    if any(pad_size):
        inputs = F.pad(inputs, pad=pad_size, mode="constant", value=0.0)
    starts = window_starts(image_size, roi_size, overlap)
    num_win = len(starts)
    total = num_win * batch
    imap = importance_map(roi_size, mode, sigma_scale).to(inputs.dtype)

    out = None
    count = torch.zeros((1, 1) + image_size, dtype=inputs.dtype)
    for g in range(0, total, sw_batch_size):
        idxs = range(g, min(g + sw_batch_size, total))
        slices = []
        for idx in idxs:
            b, w = idx // num_win, idx % num_win
            sl = (slice(b, b + 1), slice(None)) + tuple(
                slice(s, s + r) for s, r in zip(starts[w], roi_size))
            slices.append(sl)
        win = torch.cat([inputs[s] for s in slices], dim=0)
        seg = predictor(win)
        if out is None:
            out = torch.zeros((batch, seg.shape[1]) + image_size, dtype=inputs.dtype)
        seg = seg * imap  # fp32 multiply, then add (two roundings)
        for k, sl in enumerate(slices):
            out[sl] += seg[k:k + 1]
            count[(slice(0, 1), slice(None)) + sl[2:]] += imap
    out = out / count
    if any(pad_size):
        crop = [slice(None), slice(None)]
        for k in range(nd):
            lo = pad_size[2 * (nd - 1 - k)]
            crop.append(slice(lo, lo + image_size_[k]))
        out = out[tuple(crop)]
        count = count[tuple(crop)]
    if return_count:
        return out, count
    return out
