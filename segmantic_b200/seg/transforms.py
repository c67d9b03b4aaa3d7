"""Device versions of the MONAI transforms on the prediction path (geometry on the host, voxels on
the GPU).  Configured in the reference at ``/root/reference/src/segmantic/seg/monai_unet.py:151-176``
(``Orientationd("RAS")``, ``NormalizeIntensityd``, ``CropForegroundd``, ``Spacingd``) and inverted at
``:612-625`` (``Invertd(nearest_interp=False)`` + ``AsDiscreted(argmax=True)``).  Tensors are
``[C, X, Y, Z]`` float32 on a CUDA device; affines are 4x4 numpy RAS matrices.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib

AFFINE_TOL = 1e-3


def _st(dev) -> int:
    return int(torch.cuda.current_stream(dev).cuda_stream)


def _check_dev(t: torch.Tensor):
    if t.device.type != "cuda":
        raise RuntimeError("segmantic_b200 transforms run on CUDA tensors only (no CPU fallback)")


def _dbl12(xform: np.ndarray):
    m = np.asarray(xform, dtype=np.float64)[:3, :4]
    return (C.c_double * 12)(*np.ascontiguousarray(m).ravel().tolist())


# ------------------------------------------------------------------ geometry (host, tiny)
def itk_geometry_to_ras_affine(spacing, origin, direction) -> np.ndarray:
    d = len(spacing)
    aff = np.eye(d + 1)
    aff[:d, :d] = np.asarray(direction, np.float64).reshape(d, d) @ np.diag(np.asarray(spacing, np.float64))
    aff[:d, d] = np.asarray(origin, np.float64)
    flip = np.eye(d + 1)
    flip[0, 0] = flip[1, 1] = -1.0
    return flip @ aff


def ras_affine_to_itk_geometry(affine: np.ndarray):
    d = affine.shape[0] - 1
    flip = np.eye(d + 1)
    flip[0, 0] = flip[1, 1] = -1.0
    lps = flip @ np.asarray(affine, np.float64)
    spacing = np.sqrt(np.sum(np.square(lps[:d, :d]), 0))
    spacing[spacing == 0] = 1.0
    direction = lps[:d, :d] / spacing
    return tuple(spacing.tolist()), tuple(lps[:d, d].tolist()), tuple(direction.flatten().tolist())


def zoom_affine(affine: np.ndarray, pixdim: Sequence[float]) -> np.ndarray:
    affine = np.asarray(affine, dtype=np.float64)
    d = affine.shape[0] - 1
    pix = list(pixdim)[:d]
    norm = np.sqrt(np.sum(np.square(affine[:d, :d]), 0))
    pix = np.asarray(pix + list(norm[len(pix):]), dtype=np.float64)
    return affine @ np.diag(np.append(pix / norm, 1.0))


def compute_shape_offset(spatial_shape, in_affine, out_affine):
    shape = np.array(spatial_shape, dtype=float)
    d = len(shape)
    corners = np.asarray(np.meshgrid(*[(0.0, dim - 1.0) for dim in shape], indexing="ij")).reshape((d, -1))
    corners = np.concatenate((corners, np.ones_like(corners[:1])))
    corners = in_affine @ corners
    inv_mat = np.linalg.inv(out_affine)
    corners_out = inv_mat @ corners
    corners_out = corners_out[:-1] / corners_out[-1]
    out_shape = np.round(np.ptp(corners_out, axis=1) + 1.0)
    all_dist = inv_mat[:-1, :-1] @ corners[:-1, :]
    offset = None
    for i in range(corners.shape[1]):
        min_corner = np.min(all_dist - all_dist[:, i:i + 1], 1)
        if np.allclose(min_corner, 0.0, rtol=AFFINE_TOL):
            offset = corners[:-1, i]
            break
    return out_shape.astype(int), offset


# ------------------------------------------------------------------ device ops
def resample_index_affine(img: torch.Tensor, xform: np.ndarray, out_shape) -> torch.Tensor:
    """``out[c][o] = trilinear(img[c], xform @ o)``, border padding (grid_sample semantics)."""
    _check_dev(img)
    lib = _lib.load()
    img = img.contiguous().to(torch.float32)
    out_shape = tuple(int(s) for s in out_shape)
    out = torch.empty((img.shape[0],) + out_shape, dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(lib.sgm_resample_trilinear(img.data_ptr(), _lib.i3(img.shape[1:]), img.shape[0],
                                              out.data_ptr(), _lib.i3(out_shape), _dbl12(xform),
                                              _st(img.device)), "sgm_resample_trilinear")
    return out


def resample_index_affine_argmax(img: torch.Tensor, xform: np.ndarray, out_shape) -> torch.Tensor:
    """argmax over channels of the trilinear resample, fused (uint8 ``[*out_shape]``)."""
    _check_dev(img)
    lib = _lib.load()
    img = img.contiguous().to(torch.float32)
    out_shape = tuple(int(s) for s in out_shape)
    out = torch.empty(out_shape, dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(lib.sgm_resample_trilinear_argmax(img.data_ptr(), _lib.i3(img.shape[1:]), img.shape[0],
                                                     out.data_ptr(), _lib.i3(out_shape), _dbl12(xform),
                                                     _st(img.device)), "sgm_resample_trilinear_argmax")
    return out


def spacing_forward(img: torch.Tensor, affine: np.ndarray, pixdim: Sequence[float]):
    """``Spacingd(pixdim)``; returns (img, new_affine, record-for-inverse or None)."""
    affine = np.asarray(affine, dtype=np.float64)
    new_affine = zoom_affine(affine, pixdim)
    out_shape, offset = compute_shape_offset(img.shape[1:], affine, new_affine)
    new_affine[:3, -1] = offset
    if np.allclose(affine, new_affine, atol=AFFINE_TOL) and tuple(out_shape) == tuple(img.shape[1:]):
        return img, affine, None
    xform = np.linalg.solve(affine, new_affine)
    out = resample_index_affine(img, xform, out_shape)
    return out, new_affine, dict(src_affine=affine, src_shape=tuple(int(s) for s in img.shape[1:]),
                                 dst_affine=new_affine)


def spacing_inverse_xform(record) -> np.ndarray:
    return np.linalg.solve(record["dst_affine"], record["src_affine"])


def orientation_ras(img: torch.Tensor, affine: np.ndarray):
    """``Orientationd("RAS")`` for axis-aligned affines (permute + flip)."""
    d = img.dim() - 1
    a = np.asarray(affine, dtype=np.float64).copy()
    perm = [int(np.argmax(np.abs(a[i, :d]))) for i in range(d)]
    if sorted(perm) != list(range(d)):
        raise ValueError("cannot determine a unique axis permutation to RAS for this affine")
    img = img.permute([0] + [p + 1 for p in perm])
    pm = np.zeros((d + 1, d + 1))
    pm[d, d] = 1.0
    for i, p in enumerate(perm):
        pm[p, i] = 1.0
    a = a @ pm
    flips = []
    for i in range(d):
        if a[i, i] < 0:
            flips.append(i)
            fm = np.eye(d + 1)
            fm[i, i] = -1.0
            fm[i, d] = img.shape[i + 1] - 1
            a = a @ fm
    if flips:
        img = torch.flip(img, dims=[f + 1 for f in flips])
    return img.contiguous(), a, (perm, flips)


def orientation_inverse(img: torch.Tensor, record, lead: int = 1):
    perm, flips = record
    if flips:
        img = torch.flip(img, dims=[f + lead for f in flips])
    inv = [0] * len(perm)
    for i, p in enumerate(perm):
        inv[p] = i
    return img.permute(list(range(lead)) + [p + lead for p in inv]).contiguous()


def normalize_intensity(img: torch.Tensor) -> torch.Tensor:
    """``NormalizeIntensityd(nonzero=False, channel_wise=True)`` on the device."""
    _check_dev(img)
    lib = _lib.load()
    img = img.contiguous().to(torch.float32)
    out = torch.empty_like(img)
    scratch = torch.empty(4096, dtype=torch.float64, device=img.device)
    vox = int(np.prod(img.shape[1:]))
    with torch.cuda.device(img.device):
        _lib.check(lib.sgm_normalize_intensity(img.data_ptr(), out.data_ptr(), img.shape[0], vox,
                                               scratch.data_ptr(), _st(img.device)), "sgm_normalize_intensity")
    return out


def foreground_bbox(img: torch.Tensor):
    """``CropForegroundd(select_fn = x > 0)`` bounding box: (lo[3], hi[3]) with hi exclusive."""
    _check_dev(img)
    lib = _lib.load()
    img = img.contiguous().to(torch.float32)
    bbox = torch.empty(6, dtype=torch.int32, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(lib.sgm_foreground_bbox(img.data_ptr(), img.shape[0], _lib.i3(img.shape[1:]),
                                           bbox.data_ptr(), _st(img.device)), "sgm_foreground_bbox")
    b = bbox.cpu().tolist()
    return b[:3], b[3:]
