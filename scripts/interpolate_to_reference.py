"""``interpolate_to_reference``: resample a (label) image onto the grid of a reference image.

Mirrors ``/root/reference/scripts/interpolate_to_reference.py`` (``sitk_cli.make_cli(resample_to_ref)``: the function's
arguments become ``--moving-image``, ``--fixed-image``, ``--nearest`` and the result goes to ``--output``) on the B200
resampler (``segmantic_b200.image.processing.resample_to_ref`` -> ``sgm_resample_itk``, ITK ``ResampleImageFilter``
semantics, bit-exact for nearest neighbour).  NIfTI in, NIfTI out (SimpleITK is not available here).

    python scripts/interpolate_to_reference.py --moving-image labels_1mm.nii.gz --fixed-image ct.nii.gz --nearest \\
        --output labels_on_ct.nii.gz
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import typer

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmantic_b200.image import nifti, processing  # noqa: E402
from segmantic_b200.seg import transforms as T  # noqa: E402


def _load(path: Path) -> processing.Image:
    arr, affine, _ = nifti.read(path)          # [C, X, Y, Z] float32, RAS affine
    spacing, origin, direction = T.ras_affine_to_itk_geometry(affine)
    return processing.Image(np.ascontiguousarray(arr[0]), spacing, origin, direction)


def main(moving_image: Path = typer.Option(..., help="image to resample"),
         fixed_image: Path = typer.Option(..., help="reference image (defines the output grid)"),
         nearest: bool = typer.Option(False, help="nearest-neighbour (label maps) instead of linear interpolation"),
         output: Path = typer.Option(..., help="output file (.nii / .nii.gz)")) -> None:
    moving, fixed = _load(moving_image), _load(fixed_image)
    if nearest:  # label maps: keep them integer valued through the resampler (uint8 when they fit)
        a = moving.array
        if a.size and float(a.min()) >= 0 and float(a.max()) <= 255 and np.all(a == np.round(a)):
            moving = processing.Image(a.astype(np.uint8), moving.spacing, moving.origin, moving.direction)
    res = processing.resample_to_ref(moving, fixed, nearest)
    _, fixed_affine, _ = nifti.read(fixed_image)
    nifti.write(output, np.asarray(res.array).astype(np.float32), fixed_affine)


if __name__ == "__main__":
    typer.run(main)
