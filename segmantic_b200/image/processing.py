"""Image helpers with the signatures of ``segmantic.image.processing`` on the B200 resampler.

Mirrors ``/root/reference/src/segmantic/image/processing.py``: ``make_image`` (:10-24),
``extract_slices`` (:27-46), ``resample`` (:49-71), ``apply_transform`` / ``resample_to_ref``
(:74-120), ``pad`` / ``crop_center`` / ``crop`` (:123-156).  SimpleITK is not available in this image,
so images are the light ``Image`` struct below (array in ITK index order ``[x, y, z]`` + spacing,
origin, direction -- the four things the reference reads from ``sitk.Image``); it offers the same
``GetSize/GetSpacing/GetOrigin/GetDirection/GetDimension`` getters.  Resampling runs in
``sgm_resample_itk`` (no CPU fallback); the pad/crop helpers are plain index bookkeeping.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence, Union

import numpy as np
import torch

from .. import _lib

_DTYPE_CODES = {np.dtype(np.uint8): 0, np.dtype(np.int16): 1, np.dtype(np.uint16): 2,
                np.dtype(np.float32): 3, np.dtype(np.int32): 4}


@dataclass
class Image:
    """Stand-in for ``sitk.Image``: ``array`` indexed ``[x, y(, z)]`` (numpy) or a CUDA tensor."""

    array: Any
    spacing: tuple = ()
    origin: tuple = ()
    direction: tuple = field(default=())

    def __post_init__(self):
        d = self.array.ndim
        self.spacing = tuple(float(s) for s in (self.spacing or (1.0,) * d))
        self.origin = tuple(float(s) for s in (self.origin or (0.0,) * d))
        self.direction = tuple(float(s) for s in (self.direction if len(self.direction) else np.eye(d).flatten()))
        if len(self.spacing) != d or len(self.origin) != d or len(self.direction) != d * d:
            raise ValueError("shape and spacing must have same dimension")

    def GetSize(self):
        return tuple(int(s) for s in self.array.shape)

    def GetSpacing(self):
        return self.spacing

    def GetOrigin(self):
        return self.origin

    def GetDirection(self):
        return self.direction

    def GetDimension(self):
        return self.array.ndim

    def SetSpacing(self, spacing):
        if len(spacing) != self.array.ndim:
            raise ValueError("shape and spacing must have same dimension")
        self.spacing = tuple(float(s) for s in spacing)

    def numpy(self) -> np.ndarray:
        a = self.array
        return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def make_image(shape: Sequence[int], spacing: Optional[Sequence[float]] = None,
               value: Union[int, float] = 0, pixel_type: Any = np.uint8) -> Image:
    """Create (2D/3D) image with specified shape and spacing"""
    if spacing and len(shape) != len(spacing):
        raise ValueError("shape and spacing must have same dimension")
    arr = np.full(tuple(int(s) for s in shape), value, dtype=pixel_type)
    return Image(arr, tuple(spacing) if spacing else ())


def extract_slices(image: Image, axis: int = 2) -> List[Image]:
    """Get 2D image slices from 3D image (axis perpendicular to the slices; default XY slices)."""
    arr = image.numpy()
    keep = [a for a in range(3) if a != axis]
    dirm = np.asarray(image.direction).reshape(3, 3)
    out = []
    for k in range(arr.shape[axis]):
        sl = np.take(arr, k, axis=axis)
        idx = np.zeros(3)
        idx[axis] = k
        org = np.asarray(image.origin) + dirm @ (np.asarray(image.spacing) * idx)
        out.append(Image(sl.copy(), tuple(image.spacing[a] for a in keep), tuple(org[a] for a in keep),
                         tuple(dirm[np.ix_(keep, keep)].flatten())))
    return out


def _geometry(spacing, direction, d):
    m = np.eye(3)
    m[:d, :d] = np.asarray(direction, dtype=np.float64).reshape(d, d) @ np.diag(np.asarray(spacing, np.float64))
    return m


def _resample_onto_grid(moving: Image, size, spacing, origin, direction, nearest: bool, default=0,
                        device=None) -> Image:
    lib = _lib.load()
    d = moving.GetDimension()
    if d not in (2, 3):
        raise ValueError("only 2D/3D images are supported")
    src = moving.array
    if isinstance(src, torch.Tensor):
        if src.device.type != "cuda":
            raise RuntimeError("tensor images must live on a CUDA device (no CPU fallback)")
        dev = src.device
        src_t = src
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available: segmantic_b200 has no CPU fallback")
        dev = torch.device(device or "cuda:0")
        src_np = np.ascontiguousarray(np.asarray(src))
        src_t = torch.from_numpy(src_np.view(np.int16) if src_np.dtype == np.uint16 else src_np).to(dev)
    if isinstance(src, torch.Tensor):
        np_dtype = {torch.uint8: np.uint8, torch.int16: np.int16, torch.int32: np.int32,
                    torch.float32: np.float32}.get(src.dtype)
        if np_dtype is None:
            raise TypeError(f"unsupported tensor pixel type {src.dtype}")
    else:
        np_dtype = np.asarray(src).dtype
    code = _DTYPE_CODES.get(np.dtype(np_dtype))
    if code is None:
        raise TypeError(f"unsupported pixel type {np_dtype} (uint8, int16, uint16, int32, float32)")
    # ITK index order [x,y,z] with x fastest == C-contiguous array of the transposed [z,y,x] view
    src_zyx = src_t.permute(*reversed(range(d))).contiguous()
    in_dims = list(moving.GetSize()) + [1] * (3 - d)
    out_dims = [int(s) for s in size] + [1] * (3 - d)
    out_zyx = torch.empty(tuple(reversed([int(s) for s in size])), dtype=src_t.dtype, device=dev)
    i2p = _geometry(spacing, direction, d)
    p2i = np.eye(3)
    p2i[:d, :d] = np.linalg.inv(_geometry(moving.spacing, moving.direction, d)[:d, :d])
    oorg = np.zeros(3)
    oorg[:d] = origin
    iorg = np.zeros(3)
    iorg[:d] = moving.origin

    def dbl(a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        return (C.c_double * a.size)(*a.tolist())

    with torch.cuda.device(dev):
        st = int(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.sgm_resample_itk(src_zyx.data_ptr(), code, _lib.i3(in_dims), out_zyx.data_ptr(),
                                        _lib.i3(out_dims), dbl(i2p), dbl(oorg), dbl(p2i), dbl(iorg),
                                        int(bool(nearest)), float(default), st), "sgm_resample_itk")
    out = out_zyx.permute(*reversed(range(d)))
    if isinstance(src, torch.Tensor):
        res: Any = out.contiguous()
    else:
        res = out.contiguous().cpu().numpy()
        if np_dtype == np.uint16:
            res = res.view(np.uint16)
    img = Image(res, tuple(spacing), tuple(origin), tuple(direction))
    return img


def resample(image: Image, target_spacing: Sequence[float], nearest: bool = False) -> Image:
    """resample (2D/3D) image to a target spacing"""
    size = list(image.GetSize())
    spacing = list(image.GetSpacing())
    for d in range(image.GetDimension()):
        size[d] = math.ceil(size[d] * spacing[d] / target_spacing[d])
        spacing[d] = target_spacing[d]
    return _resample_onto_grid(image, size, spacing, image.GetOrigin(), image.GetDirection(), nearest, 0)


def apply_transform(moving_image: Image, fixed_image: Image, transform: Any, nearest: bool) -> Image:
    """Resample the moving image onto the fixed image's grid.  Only the identity transform
    (``None``, as ``sitk.Transform()`` in the reference's only caller) is supported."""
    if transform is not None:
        raise NotImplementedError("only the identity transform is supported (resample_to_ref)")
    return _resample_onto_grid(moving_image, fixed_image.GetSize(), fixed_image.GetSpacing(),
                               fixed_image.GetOrigin(), fixed_image.GetDirection(), nearest, 0)


def resample_to_ref(moving_image: Image, fixed_image: Image, nearest: bool) -> Image:
    """resample (2D/3D) image to a reference grid"""
    return apply_transform(moving_image=moving_image, fixed_image=fixed_image, transform=None, nearest=nearest)


def pad(image: Image, target_size: Sequence[int], value: float = 0) -> Image:
    """Pad (2D/3D) image to the target size

    (The delta expression is the reference's own, processing.py:126: it is non-zero only where the
    image is LARGER than the target, so padding up to a larger target is a no-op -- kept as is.)"""
    size = image.GetSize()
    delta = [max(s, t) - t for s, t in zip(size, target_size)]
    if any(delta):
        pad_low = [(d + 1) // 2 for d in delta]
        pad_hi = [delta[i] - p for i, p in enumerate(pad_low)]
        arr = np.pad(image.numpy(), list(zip(pad_low, pad_hi)), mode="constant", constant_values=value)
        d = image.GetDimension()
        dirm = np.asarray(image.direction).reshape(d, d)
        org = np.asarray(image.origin) - dirm @ (np.asarray(image.spacing) * np.asarray(pad_low))
        image = Image(arr, image.spacing, tuple(org), image.direction)
    return image


def crop_center(image: Image, target_size: Sequence[int]) -> Image:
    """Crop (2D/3D) image to the target size (centered)"""
    size = image.GetSize()
    delta = [max(s, t) - t for s, t in zip(size, target_size)]
    if any(delta):
        crop_low = [(d + 1) // 2 for d in delta]
        image = crop(image, crop_low, [s - d for s, d in zip(size, delta)])
    return image


def crop(img: Image, target_offset: Sequence[int], target_size: Sequence[int]) -> Image:
    """Crop (2D/3D) image to the target size/offset"""
    sl = tuple(slice(int(o), int(o) + int(s)) for o, s in zip(target_offset, target_size))
    d = img.GetDimension()
    dirm = np.asarray(img.direction).reshape(d, d)
    org = np.asarray(img.origin) + dirm @ (np.asarray(img.spacing) * np.asarray(target_offset, dtype=float))
    return Image(img.numpy()[sl].copy(), img.spacing, tuple(org), img.direction)
