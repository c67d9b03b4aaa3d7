"""Minimal NIfTI-1 (.nii / .nii.gz) reader/writer.

Stands in for ``LoadImaged(reader="ITKReader", ensure_channel_first=True)`` and
``SaveImaged(writer="ITKWriter")`` on the prediction path
(``/root/reference/src/segmantic/seg/monai_unet.py:157-162,599-609``): neither ITK nor nibabel is
installed here.  Arrays are returned channel-first in ITK index order ``[C, X, Y, Z]`` (x is the
fastest axis of the file) with a 4x4 RAS affine (NIfTI's native world frame, which is also what
MONAI's ITKReader produces after its LPS->RAS flip).  single-file, little/big endian, scalar types.
"""
from __future__ import annotations

import gzip
import struct
from pathlib import Path

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8,
           512: np.uint16, 768: np.uint32}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def _open(path: Path, mode: str):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def _quat_to_affine(hdr, pixdim, qfac) -> np.ndarray:
    b, c, d = hdr["quatern_b"], hdr["quatern_c"], hdr["quatern_d"]
    a2 = 1.0 - (b * b + c * c + d * d)
    a = np.sqrt(a2) if a2 > 1e-12 else 0.0
    if a2 <= 1e-12:
        n = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d = b * n, c * n, d * n
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    S = np.diag([pixdim[1], pixdim[2], pixdim[3] * qfac])
    aff = np.eye(4)
    aff[:3, :3] = R @ S
    aff[:3, 3] = [hdr["qoffset_x"], hdr["qoffset_y"], hdr["qoffset_z"]]
    return aff


def read(path):
    """-> (array float32 ``[C, X, Y, Z]``, 4x4 RAS affine, header dict)."""
    path = Path(path)
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError(f"{path}: not a NIfTI-1 file")
    end = "<"
    if struct.unpack("<i", raw[:4])[0] != 348:
        end = ">"
        if struct.unpack(">i", raw[:4])[0] != 348:
            raise ValueError(f"{path}: bad NIfTI-1 header size")
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(end + "hh", raw[70:74])
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset = int(struct.unpack(end + "f", raw[108:112])[0])
    slope, inter = struct.unpack(end + "ff", raw[112:120])
    qform_code, sform_code = struct.unpack(end + "hh", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(end + "6f", raw[256:280])
    srow = np.array(struct.unpack(end + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    if raw[344:348] not in (b"n+1\0", b"ni1\0"):
        raise ValueError(f"{path}: bad NIfTI-1 magic")
    if datatype not in _DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype {datatype}")
    nd = dim[0]
    shape = [max(1, d) for d in dim[1:1 + max(nd, 3)]]
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=np.dtype(_DTYPES[datatype]).newbyteorder(end), count=n, offset=max(vox_offset, 352))
    arr = data.reshape(shape[::-1])  # file order: x fastest
    arr = np.transpose(arr, list(range(arr.ndim))[::-1])  # -> [x, y, z, (t, c..)]
    if slope not in (0.0,) and not (slope == 1.0 and inter == 0.0):
        arr = arr.astype(np.float64) * slope + inter
    arr = arr.astype(np.float32)
    if arr.ndim == 3:
        arr = arr[None]
    else:  # extra dims -> channels first
        arr = np.moveaxis(arr.reshape(arr.shape[:3] + (-1,)), -1, 0)
    hdr = dict(quatern_b=qb, quatern_c=qc, quatern_d=qd, qoffset_x=qx, qoffset_y=qy, qoffset_z=qz,
               pixdim=pixdim, datatype=datatype, qform_code=qform_code, sform_code=sform_code)
    if sform_code > 0:
        aff = np.eye(4)
        aff[:3, :] = srow
    elif qform_code > 0:
        aff = _quat_to_affine(hdr, pixdim, -1.0 if pixdim[0] < 0 else 1.0)
    else:
        aff = np.diag([pixdim[1] or 1.0, pixdim[2] or 1.0, pixdim[3] or 1.0, 1.0])
    return np.ascontiguousarray(arr), aff, hdr


def write(path, array: np.ndarray, affine: np.ndarray) -> None:
    """Write a 3-D array indexed ``[x, y, z]`` with a 4x4 RAS affine (sform + qform-less header)."""
    path = Path(path)
    array = np.asarray(array)
    if array.ndim == 2:
        array = array[..., None]
    if array.ndim != 3:
        raise ValueError("write() expects a 2-D or 3-D array")
    if array.dtype not in _CODES:
        array = array.astype(np.float32)
    aff = np.asarray(affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, array.shape[0], array.shape[1], array.shape[2], 1, 1, 1, 1)
    struct.pack_into("<hh", hdr, 70, _CODES[array.dtype], array.dtype.itemsize * 8)
    sp = np.sqrt((aff[:3, :3] ** 2).sum(0))
    struct.pack_into("<8f", hdr, 76, 1.0, sp[0], sp[1], sp[2], 0.0, 0.0, 0.0, 0.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<ff", hdr, 112, 1.0, 0.0)
    hdr[123] = 2  # xyzt_units: mm
    struct.pack_into("<hh", hdr, 252, 0, 1)  # qform_code 0, sform_code 1 (scanner)
    struct.pack_into("<12f", hdr, 280, *aff[:3, :].ravel().tolist())
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + b"\0\0\0\0" + np.ascontiguousarray(np.transpose(array, (2, 1, 0))).astype(
        array.dtype.newbyteorder("<")).tobytes()
    with _open(path, "wb") as f:
        f.write(payload)
