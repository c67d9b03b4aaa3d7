"""Oracle: MONAI pre/post transforms on the prediction path, restated (test infrastructure).

Configured in the reference at ``/root/reference/src/segmantic/seg/monai_unet.py:151-176``
(``Orientationd("RAS")``, ``NormalizeIntensityd(nonzero=False, channel_wise=True)``,
``CropForegroundd(source_key, allow_smaller=False)``, ``Spacingd(pixdim=spacing)``) and inverted at
``:612-625`` (``Invertd(nearest_interp=False)`` then ``AsDiscreted(argmax=True)``).  Restated from
MONAI ~1.3; the resampling itself runs on the real ``F.affine_grid`` /
``F.grid_sample`` in float64 exactly as MONAI's ``AffineTransform(normalized=False,
reverse_indexing=True)`` does.  Arrays are ``[C, X, Y, Z]`` (ITK index order), affines 4x4 RAS.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

AFFINE_TOL = 1e-3


# ---------------------------------------------------------------------------- affine helpers
def itk_geometry_to_ras_affine(spacing, origin, direction) -> np.ndarray:
    """ITK (LPS) geometry -> 4x4 RAS affine, as MONAI's ITKReader does (negate rows 0 and 1)."""
    d = len(spacing)
    direction = np.asarray(direction, dtype=np.float64).reshape(d, d)
    aff = np.eye(d + 1)
    aff[:d, :d] = direction @ np.diag(np.asarray(spacing, dtype=np.float64))
    aff[:d, d] = np.asarray(origin, dtype=np.float64)
    flip = np.eye(d + 1)
    flip[0, 0] = flip[1, 1] = -1.0
    return flip @ aff


def zoom_affine(affine: np.ndarray, pixdim, diagonal: bool = False) -> np.ndarray:
    """MONAI ``zoom_affine`` with ``diagonal=False``: rescale columns to the new pixdim."""
    affine = np.asarray(affine, dtype=np.float64)
    d = affine.shape[0] - 1
    pix = np.asarray(list(pixdim)[:d] + [1.0] * max(0, d - len(pixdim)), dtype=np.float64)
    norm = np.sqrt(np.sum(np.square(affine[:d, :d]), 0))
    scale = np.diag(np.append(pix / norm, 1.0))
    new = affine @ scale
    assert not diagonal
    return new


def compute_shape_offset(spatial_shape, in_affine, out_affine):
    """MONAI ``compute_shape_offset(scale_extent=False)``: round(ptp+1) (half-even), min-corner offset."""
    shape = np.array(spatial_shape, dtype=float)
    d = len(shape)
    in_coords = [(0.0, dim - 1.0) for dim in shape]
    corners = np.asarray(np.meshgrid(*in_coords, indexing="ij")).reshape((d, -1))
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


def _normalize_transform(shape, align_corners=False):
    shape = torch.as_tensor(shape, dtype=torch.float64)
    norm = shape.clone()
    if align_corners:
        norm[norm <= 1.0] = 2.0
        norm = 2.0 / (norm - 1.0)
        m = torch.diag(torch.cat((norm, torch.ones(1, dtype=torch.float64))))
        m[:-1, -1] = -1.0
    else:
        norm[norm <= 0.0] = 2.0
        norm = 2.0 / norm
        m = torch.diag(torch.cat((norm, torch.ones(1, dtype=torch.float64))))
        m[:-1, -1] = 1.0 / shape - 1.0
    return m


def resample_index_affine(img: torch.Tensor, xform: np.ndarray, out_shape, mode="bilinear",
                          padding_mode="border", align_corners=False) -> torch.Tensor:
    """``out[idx] = interp(img, xform @ idx)`` through affine_grid/grid_sample in float64.

    ``img`` is ``[C, *spatial]``; ``xform`` maps OUTPUT voxel indices to INPUT voxel indices
    (MONAI ``AffineTransform(normalized=False, reverse_indexing=True)``).
    """
    d = img.dim() - 1
    theta = torch.as_tensor(np.asarray(xform, dtype=np.float64)).clone()
    rev = list(range(d - 1, -1, -1))
    theta[:d] = theta[rev].clone()
    theta[:, :d] = theta[:, rev].clone()
    src_size = list(img.shape[1:])[::-1]
    dst_size = [int(s) for s in out_shape][::-1]
    src_n = _normalize_transform(src_size, align_corners)
    dst_n = _normalize_transform(dst_size, align_corners)
    theta = src_n @ theta @ torch.linalg.inv(dst_n)
    grid = F.affine_grid(theta[:d].unsqueeze(0), [1, img.shape[0]] + [int(s) for s in out_shape],
                         align_corners=align_corners)
    out = F.grid_sample(img.unsqueeze(0).to(torch.float64), grid, mode=mode,
                        padding_mode=padding_mode, align_corners=align_corners)
    return out.squeeze(0)


# ---------------------------------------------------------------------------- transforms
def orientation_ras(img: torch.Tensor, affine: np.ndarray):
    """``Orientationd("RAS")`` for axis-aligned affines: permute + flip so the affine is +diag-dominant.

    Returns (img, affine, (perm, flips)) where the last item lets ``orientation_inverse`` undo it.
    """
    d = img.dim() - 1
    a = np.asarray(affine, dtype=np.float64)
    rzs = a[:d, :d]
    perm = [int(np.argmax(np.abs(rzs[i, :]))) for i in range(d)]  # output axis i <- input axis perm[i]
    if sorted(perm) != list(range(d)):
        raise ValueError("oracle orientation supports axis-aligned (non-oblique-ambiguous) affines only")
    img = img.permute([0] + [p + 1 for p in perm])
    pm = np.eye(d + 1)
    pm[:d, :d] = 0
    for i, p in enumerate(perm):
        pm[p, i] = 1.0
    a = a @ pm
    flips = []
    for i in range(d):
        if a[i, i] < 0:
            flips.append(i)
            n = img.shape[i + 1]
            fm = np.eye(d + 1)
            fm[i, i] = -1.0
            fm[i, d] = n - 1
            a = a @ fm
    if flips:
        img = torch.flip(img, dims=[f + 1 for f in flips])
    return img.contiguous(), a, (perm, flips)


def orientation_inverse(img: torch.Tensor, record):
    perm, flips = record
    if flips:
        img = torch.flip(img, dims=[f + 1 for f in flips])
    inv = [0] * len(perm)
    for i, p in enumerate(perm):
        inv[p] = i
    return img.permute([0] + [p + 1 for p in inv]).contiguous()


def normalize_intensity(img: torch.Tensor) -> torch.Tensor:
    """``NormalizeIntensityd(nonzero=False, channel_wise=True)``: per channel (x-mean)/std (biased)."""
    out = img.clone().to(torch.float32)
    for c in range(out.shape[0]):
        x = out[c]
        mean = x.mean()
        std = x.std(unbiased=False)
        std = std if float(std) != 0.0 else torch.tensor(1.0)
        out[c] = (x - mean) / std
    return out


def foreground_bbox(src: torch.Tensor):
    """``CropForegroundd(select_fn=x>0, margin=0)``: bbox over any channel; empty -> zeros."""
    mask = (src > 0).any(dim=0)
    d = mask.dim()
    lo, hi = [], []
    for ax in range(d):
        other = [a for a in range(d) if a != ax]
        proj = mask.any(dim=other) if other else mask
        nz = torch.nonzero(proj).flatten()
        if nz.numel() == 0:
            return [0] * d, [0] * d
        lo.append(int(nz[0]))
        hi.append(int(nz[-1]) + 1)
    return lo, hi


def crop(img: torch.Tensor, lo, hi) -> torch.Tensor:
    sl = (slice(None),) + tuple(slice(a, b) for a, b in zip(lo, hi))
    return img[sl].contiguous()


def crop_inverse(img: torch.Tensor, lo, full_shape) -> torch.Tensor:
    out = torch.zeros((img.shape[0],) + tuple(full_shape), dtype=img.dtype)
    sl = (slice(None),) + tuple(slice(a, a + s) for a, s in zip(lo, img.shape[1:]))
    out[sl] = img
    return out


def spacing_forward(img: torch.Tensor, affine: np.ndarray, pixdim):
    """``Spacingd(pixdim)`` defaults: bilinear, border, align_corners=False, float64 -> float32."""
    d = img.dim() - 1
    affine = np.asarray(affine, dtype=np.float64)
    new_affine = zoom_affine(affine, pixdim)
    out_shape, offset = compute_shape_offset(img.shape[1:], affine, new_affine)
    new_affine[:d, -1] = offset
    if np.allclose(affine, new_affine, atol=AFFINE_TOL) and tuple(out_shape) == tuple(img.shape[1:]):
        return img.to(torch.float32), affine, None
    xform = np.linalg.solve(affine, new_affine)
    out = resample_index_affine(img, xform, out_shape).to(torch.float32)
    record = dict(src_affine=affine, src_shape=tuple(img.shape[1:]), dst_affine=new_affine)
    return out, new_affine, record


def spacing_inverse(img: torch.Tensor, record):
    """Inverse of ``spacing_forward`` (same mode/padding), back onto the recorded grid."""
    if record is None:
        return img
    xform = np.linalg.solve(record["dst_affine"], record["src_affine"])
    return resample_index_affine(img, xform, record["src_shape"]).to(torch.float32)
