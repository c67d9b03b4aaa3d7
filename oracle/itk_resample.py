"""Oracle: ITK ``ResampleImageFilter`` (identity transform) restated in numpy float64 (test infra).

Reference call sites: ``/root/reference/src/segmantic/image/processing.py:49-71`` (``resample``:
``size = ceil(n*s/t)``, same origin/direction, linear|nearest, default pixel 0, output pixel type =
input's), ``:74-98`` (``apply_transform``) and ``:101-120`` (``resample_to_ref``: the fixed image's
grid, the moving image's pixel type).  SimpleITK/ITK are not installed; the filter's published
algorithm (ITK 5.x ``ResampleImageFilter::LinearThreadedGenerateData`` +
``NearestNeighborInterpolateImageFunction`` / ``LinearInterpolateImageFunction``) is restated per
SURVEY.md appendix A.4.  Arrays are indexed ``[x, y, z]`` (ITK index order).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np


@dataclass
class Image:
    """Minimal stand-in for ``sitk.Image``: array in ITK index order + geometry (LPS)."""

    array: np.ndarray
    spacing: tuple = ()
    origin: tuple = ()
    direction: tuple = field(default=())

    def __post_init__(self):
        d = self.array.ndim
        if not self.spacing:
            self.spacing = (1.0,) * d
        if not self.origin:
            self.origin = (0.0,) * d
        if not len(self.direction):
            self.direction = tuple(np.eye(d).flatten())
        self.spacing = tuple(float(s) for s in self.spacing)
        self.origin = tuple(float(s) for s in self.origin)
        self.direction = tuple(float(s) for s in self.direction)

    def GetSize(self):
        return tuple(self.array.shape)

    def GetSpacing(self):
        return self.spacing

    def GetOrigin(self):
        return self.origin

    def GetDirection(self):
        return self.direction

    def GetDimension(self):
        return self.array.ndim


def _index_to_physical(spacing, direction, d):
    return np.asarray(direction, dtype=np.float64).reshape(d, d) @ np.diag(np.asarray(spacing, np.float64))


def continuous_index_lines(out_size, out_spacing, out_origin, out_direction, moving: Image):
    """Scanline start/end continuous indices (ITK linear-transform fast path).

    For each output scanline (all indices with the same o[1:]): ``start`` = continuous input index
    of o[0]=0, ``end`` = that of o[0]=size[0]; pixel i uses ``start + (i/size0)*(end-start)``.
    Returns arrays of shape ``out_size[1:] + (d,)``.
    """
    d = len(out_size)
    i2p_out = _index_to_physical(out_spacing, out_direction, d)
    p2i_in = np.linalg.inv(_index_to_physical(moving.spacing, moving.direction, d))
    o_out = np.asarray(out_origin, np.float64)
    o_in = np.asarray(moving.origin, np.float64)
    rest = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in out_size[1:]], indexing="ij")

    def cidx(i0):
        # ITK order of operations (TransformIndexToPhysicalPoint then
        # TransformPhysicalPointToContinuousIndex): row sums left to right, origin added last,
        # separate multiply and add roundings -- mirrored op for op by the CUDA kernel.
        idx = [np.full(tuple(out_size[1:]), float(i0))] + list(rest)
        idx = idx + [np.zeros_like(idx[0])] * (3 - d)
        ph = []
        for r in range(d):
            s = i2p_out[r, 0] * idx[0]
            for c in range(1, d):
                s = s + i2p_out[r, c] * idx[c]
            ph.append((s + o_out[r]) - o_in[r])
        out = []
        for r in range(d):
            s = p2i_in[r, 0] * ph[0]
            for c in range(1, d):
                s = s + p2i_in[r, c] * ph[c]
            out.append(s)
        return np.stack(out, axis=-1)

    return cidx(0), cidx(out_size[0])


def _cast_like_itk(val: np.ndarray, dtype) -> np.ndarray:
    """``CastPixelWithBoundsChecking``: clamp to the type's range, then static_cast (truncation)."""
    dtype = np.dtype(dtype)
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        val = np.clip(val, float(info.min), float(info.max))
        return np.trunc(val).astype(dtype)
    if dtype == np.float32:
        fi = np.finfo(np.float32)
        return np.clip(val, float(fi.min), float(fi.max)).astype(np.float32)
    return val.astype(dtype)


def resample_onto_grid(moving: Image, out_size: Sequence[int], out_spacing, out_origin,
                       out_direction, nearest: bool, default_value=0) -> Image:
    d = moving.array.ndim
    out_size = tuple(int(s) for s in out_size)
    n_in = moving.array.shape
    start, end = continuous_index_lines(out_size, out_spacing, out_origin, out_direction, moving)
    n0 = out_size[0]
    alpha = (np.arange(n0, dtype=np.float64) / float(n0)).reshape((n0,) + (1,) * (d - 1) + (1,))
    c = start[None] + alpha * (end - start)[None]  # [n0, rest..., d]
    inside = np.ones(out_size, dtype=bool)
    for ax in range(d):
        inside &= (c[..., ax] >= -0.5) & (c[..., ax] < n_in[ax] - 0.5)
    src = moving.array
    if nearest:
        idx = [np.clip(np.floor(c[..., ax] + 0.5).astype(np.int64), 0, n_in[ax] - 1) for ax in range(d)]
        vals = src[tuple(idx)]
        out = np.where(inside, vals, np.asarray(default_value, dtype=src.dtype)).astype(src.dtype)
    else:
        lo, hi, fr = [], [], []
        for ax in range(d):
            b = np.maximum(np.floor(c[..., ax]).astype(np.int64), 0)
            b = np.minimum(b, n_in[ax] - 1)
            f = np.maximum(c[..., ax] - b, 0.0)
            h = np.minimum(b + 1, n_in[ax] - 1)
            lo.append(b), hi.append(h), fr.append(f)
        srcd = src.astype(np.float64)

        def gather(sel):
            return srcd[tuple(hi[a] if sel[a] else lo[a] for a in range(d))]

        def lerp(ax, sel):
            if ax < 0:
                return gather(sel)
            a = lerp(ax - 1, sel[:ax] + (0,) + sel[ax + 1:])
            b = lerp(ax - 1, sel[:ax] + (1,) + sel[ax + 1:])
            return a + fr[ax] * (b - a)

        val = lerp(d - 1, (0,) * d)
        out = _cast_like_itk(np.where(inside, val, float(default_value)), src.dtype)
    return Image(out, tuple(out_spacing), tuple(out_origin), tuple(out_direction))


def resample(image: Image, target_spacing: Sequence[float], nearest: bool = False) -> Image:
    """``processing.resample`` (processing.py:49-71)."""
    size = list(image.GetSize())
    spacing = list(image.GetSpacing())
    for d in range(image.GetDimension()):
        size[d] = math.ceil(size[d] * spacing[d] / target_spacing[d])
        spacing[d] = target_spacing[d]
    return resample_onto_grid(image, size, spacing, image.GetOrigin(), image.GetDirection(), nearest, 0)


def resample_to_ref(moving_image: Image, fixed_image: Image, nearest: bool) -> Image:
    """``processing.resample_to_ref`` (processing.py:101-120): fixed's grid, moving's pixel type."""
    return resample_onto_grid(moving_image, fixed_image.GetSize(), fixed_image.GetSpacing(),
                              fixed_image.GetOrigin(), fixed_image.GetDirection(), nearest, 0)
