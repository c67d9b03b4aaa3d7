"""Oracle: the composition of ``predict()`` for one image, restated on the CPU (test infrastructure).

Follows ``/root/reference/src/segmantic/seg/monai_unet.py:589-670``: pre-transforms (``:151-176``:
Orientation RAS -> NormalizeIntensity -> CropForeground(image > 0) -> [Spacing]) -> sliding-window
inference of the eval-mode UNet (``:637-639,665``; defaults overlap 0.25, constant) -> post-transforms
(``:612-625``: Invertd with trilinear interpolation of the C-channel logits, then argmax).
``invert="labels"`` is BASELINE.json's north-star variant: argmax on the network grid, then ITK
nearest-neighbour resampling of the label map back onto the pre-Spacing grid.
"""
from __future__ import annotations

import numpy as np
import torch

from . import itk_resample as oitk
from . import sliding_window as osw
from . import spacing as osp


def _ras_to_itk(affine):
    d = affine.shape[0] - 1
    flip = np.eye(d + 1)
    flip[0, 0] = flip[1, 1] = -1.0
    lps = flip @ affine
    sp = np.sqrt((lps[:d, :d] ** 2).sum(0))
    return tuple(sp), tuple(lps[:d, d]), tuple((lps[:d, :d] / sp).flatten())


def predict_volume(net, image: torch.Tensor, affine=None, spacing=(), roi=(96, 96, 96), overlap=0.25,
                   mode="constant", sw_batch_size=4, invert="logits", predictor=None, return_gap=False, label=None):
    """``image``: ``[C, X, Y, Z]`` float32 (ITK index order); returns the uint8 label map ``[X, Y, Z]``
    and the blended logits on the network grid.  ``return_gap``: additionally the top-2 logit gap behind every
    output label, carried through the same post-transforms as the labels (+inf outside the foreground crop) -- the
    tests use it to tell argmax near-ties from real disagreements.  ``label`` (``[X, Y, Z]``): the evaluation branch
    (``default_preprocessing(keys=["image", "label"])``, ``monai_unet.py:151-176``): the label is reoriented with the
    image, the foreground crop comes from ``label > 0`` (``source_key="label"``), the label goes through the same
    Spacing (bilinear) and is truncated by ``.long()`` (``:674``); returns ``(label map, logits, network-grid
    prediction, network-grid label)`` -- the last two are what the reference scores (``:672-680``)."""
    if affine is None:
        affine = np.diag([-1.0, -1.0, 1.0, 1.0])
    lab_img = None
    if label is not None:
        lab_img = osp.orientation_ras(label.to(torch.float32)[None], affine)[0]
    img, aff, orient = osp.orientation_ras(image.to(torch.float32), affine)
    img = osp.normalize_intensity(img)
    oriented_shape = tuple(img.shape[1:])
    lo, hi = osp.foreground_bbox(lab_img if lab_img is not None else img)
    if all(h > l for l, h in zip(lo, hi)):
        img = osp.crop(img, lo, hi)
        if lab_img is not None:
            lab_img = osp.crop(lab_img, lo, hi)
        shift = np.eye(4)
        shift[:3, 3] = lo
        aff = aff @ shift
    else:
        lo, hi = [0, 0, 0], list(oriented_shape)
    record = None
    if len(spacing):
        if lab_img is not None:
            lab_img = osp.spacing_forward(lab_img, aff, spacing)[0]
        img, aff, record = osp.spacing_forward(img, aff, spacing)
    fn = predictor if predictor is not None else net
    with torch.no_grad():
        logits = osw.sliding_window_inference(img[None], roi, sw_batch_size, fn, overlap=overlap, mode=mode)[0]
    if record is None:
        lab = logits.argmax(0).to(torch.uint8)
    elif invert == "logits":
        lab = osp.spacing_inverse(logits, record).argmax(0).to(torch.uint8)
    else:
        sp_d, org_d, dir_d = _ras_to_itk(record["dst_affine"])
        sp_s, org_s, dir_s = _ras_to_itk(record["src_affine"])
        moving = oitk.Image(logits.argmax(0).to(torch.uint8).numpy(), sp_d, org_d, dir_d)
        fixed = oitk.Image(np.zeros(record["src_shape"], np.uint8), sp_s, org_s, dir_s)
        lab = torch.from_numpy(oitk.resample_to_ref(moving, fixed, True).array)
    gap = None
    if return_gap:
        top2 = logits.topk(2, dim=0).values
        gap = top2[0] - top2[1]
        if record is not None and invert == "logits":
            top2 = osp.spacing_inverse(logits, record).topk(2, dim=0).values
            gap = top2[0] - top2[1]
        elif record is not None:
            gap = torch.from_numpy(oitk.resample_to_ref(oitk.Image(gap.numpy().astype(np.float32), sp_d, org_d, dir_d),
                                                        fixed, True).array.astype(np.float32))
        if tuple(gap.shape) != oriented_shape:
            inside = osp.crop_inverse(torch.ones_like(gap)[None], lo, oriented_shape)[0] > 0
            gap = torch.where(inside, osp.crop_inverse(gap[None], lo, oriented_shape)[0], torch.tensor(float("inf")))
        gap = osp.orientation_inverse(gap[None], orient)[0]
    if tuple(lab.shape) != oriented_shape:
        lab = osp.crop_inverse(lab[None], lo, oriented_shape)[0]
    lab = osp.orientation_inverse(lab[None], orient)[0]
    if label is not None:
        return lab, logits, logits.argmax(0).to(torch.uint8), lab_img[0].long().clamp(0, 255).to(torch.uint8)
    if return_gap:
        return lab, logits, gap
    return lab, logits
