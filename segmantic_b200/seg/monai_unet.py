"""``segmantic.seg.monai_unet`` prediction API on the B200-native engine.

Mirrors the prediction half of ``/root/reference/src/segmantic/seg/monai_unet.py``:

* ``Net`` (``:99-149``): same constructor arguments / ``hparams`` names (``num_classes``,
  ``num_channels``, ``spatial_dims``, ``spatial_size``, ``channels``, ``strides``, ``dropout``,
  ``act``), ``num_classes`` / ``spatial_dims`` properties, ``spatial_size`` default ``[96]*3``,
  ``load_from_checkpoint(file, **overrides)`` accepting a Lightning ``.ckpt`` (keys ``_model.model.*``
  + ``hyper_parameters``) or a plain MONAI ``.pth`` (``scripts/extract_unet.py:17-18``);
  keyword overrides win over saved hyper-parameters as in the reference (``:571-573``).
* ``predict`` (``:551-725``): same signature; new keyword-only options default to the reference's
  behaviour (``overlap=0.25``, ``mode="constant"``, ``sw_batch_size=4``, ``precision="fp32"``,
  ``invert="logits"``).
* ``predict_volume``: the array-level core of ``predict`` (pre-transforms -> sliding window -> inverse
  -> argmax) for callers that already hold the voxels (and what ``bench.py`` times end to end).

Training, cross-validation and ensembles are out of scope.  There is no CPU
fallback: ``gpu_ids=[-1]`` raises.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import transforms as T
from .engine import UNetB200, sliding_window_inference
from .utils import make_device


class Net:
    """Inference-only stand-in for the reference's LightningModule ``Net``."""

    def __init__(self, num_classes: int, num_channels: int = 1, spatial_dims: int = 3,
                 spatial_size: Sequence[int] = None, channels: tuple = (16, 32, 64, 128, 256),
                 strides: tuple = (2, 2, 2, 2), dropout: float = 0.0, act: str = "PRELU"):
        if str(act).upper() != "PRELU":
            raise ValueError("only act='PRELU' (the reference default) is supported")
        self.hparams = SimpleNamespace(num_classes=num_classes, num_channels=num_channels,
                                       spatial_dims=spatial_dims, spatial_size=spatial_size,
                                       channels=tuple(channels), strides=tuple(strides), dropout=dropout, act=act)
        self.spatial_size = list(spatial_size) if spatial_size else [96] * 3
        self._state_dict: Optional[Dict[str, torch.Tensor]] = None
        self._engines: Dict[tuple, UNetB200] = {}
        self.device = torch.device("cpu")

    # -- reference properties (monai_unet.py:143-149)
    @property
    def num_classes(self) -> int:
        return int(self.hparams.num_classes)

    @property
    def spatial_dims(self) -> int:
        return int(self.hparams.spatial_dims)

    # -- checkpoint handling
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self._state_dict = {k: v.detach().cpu() for k, v in state_dict.items() if isinstance(v, torch.Tensor)}
        self._engines.clear()

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs) -> "Net":
        """Lightning ``.ckpt`` or plain ``.pth``; ``kwargs`` override saved hyper-parameters."""
        import pickle
        import warnings
        try:
            ckpt = torch.load(str(checkpoint_path), map_location="cpu", weights_only=True)
        except pickle.UnpicklingError as e:
            # Lightning checkpoints may pickle objects the safe loader refuses (callbacks, omegaconf hyper-parameters).
            # Full unpickling executes code from the file: only with the caller's explicit consent.  Missing / corrupt
            # files (OSError, RuntimeError, EOFError) propagate unchanged.
            if os.environ.get("SEGMANTIC_TRUST_CHECKPOINT", "0") != "1":
                raise RuntimeError(
                    f"{checkpoint_path} needs full unpickling ({e}); set SEGMANTIC_TRUST_CHECKPOINT=1 if you trust the "
                    "file (torch.load(weights_only=False) can execute arbitrary code)") from e
            warnings.warn(f"loading {checkpoint_path} with weights_only=False (SEGMANTIC_TRUST_CHECKPOINT=1)")
            ckpt = torch.load(str(checkpoint_path), map_location="cpu", weights_only=False)
        if isinstance(ckpt, dict) and "state_dict" in ckpt:
            hp = dict(ckpt.get("hyper_parameters", {}) or {})
            sd = ckpt["state_dict"]
        else:
            hp, sd = {}, ckpt
        hp.update(kwargs)
        if "num_classes" not in hp or "num_channels" not in hp:
            hp.update(_infer_io_channels(sd, hp))
        allowed = ("num_classes", "num_channels", "spatial_dims", "spatial_size", "channels", "strides",
                   "dropout", "act")
        net = cls(**{k: v for k, v in hp.items() if k in allowed})
        net.load_state_dict(sd)
        return net

    def freeze(self) -> None:  # the engine is always frozen / eval
        return None

    def eval(self) -> "Net":
        return self

    def to(self, device) -> "Net":
        self.device = torch.device(device)
        return self

    def engine(self, precision: str = "fp32") -> UNetB200:
        if self._state_dict is None:
            raise RuntimeError("Net has no weights: use Net.load_from_checkpoint or load_state_dict")
        key = (str(self.device), precision)
        if key not in self._engines:
            self._engines[key] = UNetB200(self._state_dict, spatial_dims=self.spatial_dims,
                                          in_channels=int(self.hparams.num_channels),
                                          out_channels=self.num_classes, channels=self.hparams.channels,
                                          strides=self.hparams.strides, device=self.device, precision=precision)
        return self._engines[key]

    def forward(self, x: torch.Tensor, precision: str = "fp32") -> torch.Tensor:
        return self.engine(precision)(x)

    __call__ = forward


def _infer_io_channels(sd, hp) -> dict:
    """num_channels / num_classes / spatial_dims from the first and last conv of a MONAI UNet state_dict."""
    keys = {k[len("_model."):] if k.startswith("_model.") else k: v for k, v in sd.items()}
    first = keys["model.0.conv.unit0.conv.weight"]
    last = keys["model.2.1.conv.unit0.conv.weight"]
    out = {"num_channels": int(first.shape[1]), "num_classes": int(last.shape[0])}
    if "spatial_dims" not in hp:
        out["spatial_dims"] = first.dim() - 2
    return out


_PIPE_DEPTH = 4   # pinned result buffers in rotation (predict_volumes)


def _pipe_state(eng) -> dict:
    """Streams and buffer rings of the pipelined loop, created once per engine (re-created streams and per-volume
    allocations of pinned / cross-stream device memory cost milliseconds each and stall the pipeline)."""
    st = eng.__dict__.get("_pipe")
    if st is None:
        dev = eng.device
        st = eng.__dict__["_pipe"] = dict(copy_in=torch.cuda.Stream(dev), copy_out=torch.cuda.Stream(dev),
                                          dev_in={}, host_out={})
    return st


def predict_volumes(net: Net, images, affines=None, spacing: Sequence[float] = (), *, precision: str = "fp32",
                    **kwargs):
    """Pipelined ``predict_volume`` over a sequence of HOST images (the loop of ``predict()`` over ``test_images``,
    ``monai_unet.py:663-670``): yields one host uint8 label map per image, in order.

    The z-score normalisation needs the whole volume before the first window can run, so the upload of ONE volume
    cannot overlap its own prediction; across volumes it can: the host->device copy of image ``i + 1`` (copy engine, its
    own stream, second device buffer) and the device->host copy of label map ``i - 1`` run while the SMs predict image
    ``i``.  Every image still crosses PCIe exactly once in each direction; results are those of ``predict_volume``.

    Buffers are rings owned by the engine (two device input buffers and ``_PIPE_DEPTH`` pinned result buffers per
    shape): a yielded label map is overwritten once ``_PIPE_DEPTH`` further results have been produced -- consume or
    copy it before (``predict()`` writes every map to disk at once)."""
    eng = net.engine(precision)
    dev = eng.device
    images = list(images)
    affines = list(affines) if affines is not None else [None] * len(images)
    if not images:
        return
    st = _pipe_state(eng)
    compute = torch.cuda.current_stream(dev)
    copy_in, copy_out = st["copy_in"], st["copy_out"]
    done_events = {}   # image index -> event: its prediction has finished on the compute stream

    def upload(i):
        src = images[i]
        if src.device.type != "cpu":
            return src, None
        if not src.is_pinned():
            src = src.pin_memory()
        key = (tuple(src.shape), src.dtype)
        ring = st["dev_in"].get(key)
        if ring is None:
            with torch.cuda.device(dev):
                ring = st["dev_in"][key] = [torch.empty(src.shape, dtype=src.dtype, device=dev) for _ in range(2)]
            torch.cuda.current_stream(dev).synchronize()   # (first use only) the buffers exist before another stream writes
        d = ring[i % 2]
        if i - 2 in done_events:        # the slot's previous user (image i - 2) has been predicted
            copy_in.wait_event(done_events[i - 2])
        else:                           # first uses of the slot in this call: nothing of an earlier call reads it any more
            copy_in.wait_stream(compute)
        with torch.cuda.stream(copy_in):
            d.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_in)
        return d, ev

    pending = None          # (host label map, its copy event, device label map kept alive) of the previous image
    nxt = upload(0)
    for i in range(len(images)):
        cur, ev = nxt
        nxt = upload(i + 1) if i + 1 < len(images) else None   # queued BEFORE this image's kernels are launched
        if ev is not None:
            compute.wait_event(ev)
        lab = predict_volume(net, cur, affines[i], spacing, precision=precision, return_device=True, check=False,
                             **kwargs)
        done = torch.cuda.Event()
        done.record(compute)
        done_events[i] = done
        done_events.pop(i - 3, None)
        okey = (tuple(lab.shape), lab.dtype)
        oring = st["host_out"].get(okey)
        if oring is None:
            oring = st["host_out"][okey] = [torch.empty(lab.shape, dtype=lab.dtype, pin_memory=True)
                                            for _ in range(_PIPE_DEPTH)]
        host = oring[i % _PIPE_DEPTH]
        copy_out.wait_event(done)
        with torch.cuda.stream(copy_out):
            host.copy_(lab, non_blocking=True)
            out_ev = torch.cuda.Event()
            out_ev.record(copy_out)
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]
        pending = (host, out_ev, lab)
    pending[1].synchronize()
    eng.check()   # a tcgen05 pipeline timeout anywhere in the sequence raises here
    yield pending[0]


def predict_volume(net: Net, image: torch.Tensor, affine: Optional[np.ndarray] = None,
                   spacing: Sequence[float] = (), *, overlap: float = 0.25, mode: str = "constant",
                   sw_batch_size: int = 4, precision: str = "fp32", invert: str = "logits",
                   normalize: bool = True, crop_foreground: bool = True, return_device: bool = False,
                   ensemble: Optional[dict] = None, check: bool = True, label: Optional[torch.Tensor] = None):
    """The array-level core of ``predict()`` (``monai_unet.py:589-670``) for ONE image.

    ``label`` (``[X, Y, Z]``, the image's ground truth on the same grid): the evaluation branch of the reference --
    ``default_preprocessing(keys=["image", "label"])`` (``monai_unet.py:151-176``) reorients the label with the image,
    takes the foreground crop from ``label > 0`` instead of ``image > 0`` (``source_key="label"``, ``:167``) and sends
    the label through the same ``Spacingd`` (bilinear, then ``.long()`` at ``:674``).  The call then returns
    ``(label map on the original grid, prediction on the NETWORK grid, transformed label on the network grid)``, the
    last two being what the reference scores (``:672-680``).

    ``image``: ``[C, X, Y, Z]`` (``[C, X, Y]`` for 2-D networks) in ITK index order, any device (host
    tensors are uploaded).  ``affine``: 4x4 RAS affine of that array (identity direction / unit spacing
    if omitted).  Applies Orientation(RAS) -> NormalizeIntensity -> CropForeground(image > 0) ->
    [Spacing(spacing)] -> sliding-window UNet -> inverse -> argmax, and returns the label map as uint8
    ``[X, Y, Z]`` on the original grid.

    ``invert="logits"`` reproduces the reference (``Invertd`` resamples the C-channel logits
    trilinearly, then argmax); ``invert="labels"`` is the north-star variant (argmax on the network
    grid, then nearest-neighbour resampling of the label map).  Without ``spacing`` both are identical.
    """
    eng = net.engine(precision)
    dev = eng.device
    nd = net.spatial_dims
    if image.dim() != nd + 1:
        raise ValueError(f"image must be [C, *spatial({nd})], got {tuple(image.shape)}")
    img = image.to(dev, dtype=torch.float32, non_blocking=True)
    if nd == 2:
        img = img.unsqueeze(1)  # [C, 1, X, Y]: the flat axis leads
    if affine is None:
        affine = np.eye(4)
        if nd == 3:
            affine[0, 0] = affine[1, 1] = -1.0  # identity-direction ITK image seen in RAS
    full_shape_itk = tuple(img.shape[1:])
    lab_img = None
    if label is not None:
        if nd != 3:
            raise ValueError("the labels-supplied evaluation branch is implemented for 3-D networks")
        lab_img = label.to(dev, dtype=torch.float32, non_blocking=True).reshape((1,) + full_shape_itk)
    if nd == 3:
        if lab_img is not None:
            lab_img = T.orientation_ras(lab_img, affine)[0]
        img, aff, orient = T.orientation_ras(img, affine)
    else:
        aff, orient = np.asarray(affine, dtype=np.float64), None
    if normalize:
        img = T.normalize_intensity(img)
    oriented_shape = tuple(img.shape[1:])
    lo, hi = [0, 0, 0], list(oriented_shape)
    if crop_foreground:
        lo, hi = T.foreground_bbox(lab_img if lab_img is not None else img)   # source_key = "label" when given
        if all(h > l for l, h in zip(lo, hi)):
            img = img[:, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].contiguous()
            if lab_img is not None:
                lab_img = lab_img[:, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].contiguous()
            shift = np.eye(4)
            shift[:3, 3] = lo
            aff = aff @ shift
        else:
            lo, hi = [0, 0, 0], list(oriented_shape)
    record = None
    if len(spacing) and nd == 3:
        if lab_img is not None:
            lab_img = T.spacing_forward(lab_img, aff, spacing)[0]
        img, aff, record = T.spacing_forward(img, aff, spacing)
    net_in = img if nd == 3 else img[:, 0]
    want_labels = record is None or invert == "labels" or ensemble is not None
    want_net_labels = want_labels or lab_img is not None
    if ensemble is not None:
        # several models over the same pre-processed volume, combined on the network grid (ensemble_creator)
        from . import ensemble as ENS
        res = {"labels": ENS.combine(ensemble, net_in.unsqueeze(0), sw_batch_size, precision, overlap, mode)}
    else:
        res = sliding_window_inference(net_in.unsqueeze(0), net.spatial_size, sw_batch_size, eng, overlap=overlap,
                                       mode=mode, return_labels=want_net_labels, return_logits=not want_labels)
    pred_net = res["labels"][0, 0] if (lab_img is not None and isinstance(res, dict)) else None
    cropped_shape = tuple(hi[a] - lo[a] for a in range(3))
    if want_labels:
        lab = res["labels"][0, 0]
        if nd == 2:
            lab = lab.unsqueeze(0)
        if record is not None:  # nearest-neighbour back-resample of the label map (north-star stage 4)
            lab = _nearest_back(lab, record)
    else:
        logits = res if isinstance(res, torch.Tensor) else res["logits"]
        lab = T.resample_index_affine_argmax(logits[0], T.spacing_inverse_xform(record), record["src_shape"])
    if tuple(lab.shape) != oriented_shape:  # inverse CropForeground: zero padding -> label 0
        full = torch.zeros(oriented_shape, dtype=torch.uint8, device=dev)
        full[lo[0]:lo[0] + cropped_shape[0], lo[1]:lo[1] + cropped_shape[1], lo[2]:lo[2] + cropped_shape[2]] = lab
        lab = full
    if orient is not None:
        lab = T.orientation_inverse(lab, orient, lead=0)
    assert tuple(lab.shape) == full_shape_itk
    if nd == 2:
        lab = lab[0]
    if check:  # synchronises: the pipelined caller (predict_volumes) checks once at the end instead
        eng.check()
    if lab_img is not None:
        # val_labels = test_data["label"].long() (monai_unet.py:674): truncation of the (bilinearly resampled) label
        label_net = lab_img[0].to(torch.int64).clamp_(0, 255).to(torch.uint8)
        return (lab if return_device else lab.cpu()), pred_net, label_net
    if return_device:
        return lab
    # device -> host into pinned memory (torch's caching host allocator reuses the block across calls; a pageable
    # destination costs an allocation plus a staged copy per call)
    host = torch.empty(lab.shape, dtype=lab.dtype, pin_memory=True)
    host.copy_(lab, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return host


def predict_stack_2d(net: Net, image: torch.Tensor, *, overlap: float = 0.25, mode: str = "constant",
                     sw_batch_size: int = 4, precision: str = "fp32", normalize: bool = True,
                     return_device: bool = False):
    """Slice-wise prediction of a stack with a 2-D network (BASELINE configs[4]; the reference has no code for it --
    SURVEY.md 8d defines the Z slices as independent 2-D images).  ``image``: ``[C, X, Y, Z]`` (ITK index order, any
    device); z-score normalisation per channel over the whole stack (``NormalizeIntensityd`` on the loaded image), then
    every slice ``[C, X, Y]`` goes through the 2-D sliding window; all slices' windows share the network launches.
    Returns uint8 labels ``[X, Y, Z]``."""
    if net.spatial_dims != 2:
        raise ValueError("predict_stack_2d needs a network with spatial_dims=2")
    if image.dim() != 4:
        raise ValueError(f"image must be [C, X, Y, Z], got {tuple(image.shape)}")
    eng = net.engine(precision)
    img = image.to(eng.device, dtype=torch.float32, non_blocking=True)
    if normalize:
        img = T.normalize_intensity(img)
    stack = img.permute(0, 3, 1, 2).contiguous()  # [C, Z, X, Y]
    res = sliding_window_inference(stack.unsqueeze(0), net.spatial_size, sw_batch_size, eng, overlap=overlap, mode=mode,
                                   return_labels=True, return_logits=False)
    eng.check()
    lab = res["labels"][0, 0].permute(1, 2, 0).contiguous()  # [X, Y, Z]
    if return_device:
        return lab
    host = torch.empty(lab.shape, dtype=lab.dtype, pin_memory=True)
    host.copy_(lab, non_blocking=True)
    torch.cuda.current_stream(eng.device).synchronize()
    return host


def _nearest_back(lab: torch.Tensor, record) -> torch.Tensor:
    """Nearest-neighbour resample of a label map from the network grid back onto the pre-Spacing grid
    (ITK semantics, ``image/processing.resample_to_ref``)."""
    from ..image import processing as P

    sp_d, org_d, dir_d = T.ras_affine_to_itk_geometry(record["dst_affine"])
    sp_s, org_s, dir_s = T.ras_affine_to_itk_geometry(record["src_affine"])
    moving = P.Image(lab, sp_d, org_d, dir_d)
    fixed = P.Image(torch.empty(record["src_shape"], dtype=torch.uint8, device="meta"), sp_s, org_s, dir_s)
    return P.resample_to_ref(moving, fixed, nearest=True).array


def predict(model_file: Path, test_images: List[Path], test_labels: Optional[List[Path]] = None,
            output_dir: Path = None, tissue_dict: Dict[str, int] = None,
            channels: tuple = (16, 32, 64, 128, 256), strides: tuple = (2, 2, 2, 2), dropout: float = 0.0,
            spacing: Sequence[float] = [], gpu_ids: List[int] = [], *, overlap: float = 0.25,
            mode: str = "constant", sw_batch_size: int = 4, precision: str = "fp32",
            invert: str = "logits") -> None:
    """Same signature and side effects as ``segmantic.seg.monai_unet.predict`` (``:551-562``):
    writes ``<output_dir>/<image basename>.nii.gz`` label maps (float32, as ``SaveImaged`` does)."""
    from ..image import nifti

    model_file = Path(model_file)
    model_settings_json = model_file.with_suffix(".json")
    if model_settings_json.exists():
        print(f"WARNING: Loading legacy model settings from {model_settings_json}")
        with model_settings_json.open() as json_file:
            settings = json.load(json_file)
        net = Net.load_from_checkpoint(f"{model_file}", **settings)
    else:
        net = Net.load_from_checkpoint(f"{model_file}", channels=channels, strides=strides, dropout=dropout)
    num_classes = net.num_classes
    net.freeze()
    net.eval()
    device = make_device(gpu_ids)
    if device.type != "cuda":
        raise RuntimeError("segmantic_b200 has no CPU path: pass gpu_ids=[] or a non-negative GPU id")
    net.to(device)

    tissue_names = [f"{id}" for id in range(num_classes)]
    if tissue_dict:
        for name in tissue_dict.keys():
            idx = tissue_dict[name]
            if 0 <= idx < num_classes:
                tissue_names[idx] = name
    if output_dir:
        os.makedirs(output_dir, exist_ok=True)
    have_labels = test_labels is not None and len(test_labels) == len(test_images) and len(test_labels) > 0
    from . import evaluation as E

    def print_table(header, vals, indent="\t"):
        print(indent + "\t".join(header).expandtabs(30))
        print(indent + "\t".join(f"{x}" for x in vals).expandtabs(30))

    # evaluation state (monai_unet.py:640-646): DiceMetric(include_background=False, reduction="mean"),
    # ConfusionMatrixMetric([sensitivity, specificity, precision, accuracy]), CumulativeAverage of the class Dice
    all_mean_dice: List[float] = []       # dice_metric.aggregate() after every image (running mean over images)
    image_mean_dice: List[float] = []
    class_dice_sum = np.zeros(max(num_classes - 1, 0))
    class_dice_cnt = np.zeros(max(num_classes - 1, 0))
    all_counts = []
    # Several gpu_ids: the reference keeps gpu_ids[0] only (seg/utils.py:4-12).  Here the IMAGES of the call are spread
    # round-robin over the listed devices (one worker thread and one engine per device; the C library releases the GIL),
    # results are consumed in order.  One volume across several GPUs is the torchrun driver (seg/multi_gpu.py).
    devices = [device]
    if gpu_ids and len(gpu_ids) > 1:
        devices = [torch.device(f"cuda:{int(g)}") for g in gpu_ids if int(g) >= 0] or [device]
    import copy
    nets = []
    for d in devices:
        nd_ = net if d == device else copy.copy(net)
        if nd_ is not net:
            nd_._engines = {}
            nd_.to(d)
        nets.append(nd_)

    def run_one(i: int):
        net_d = nets[i % len(nets)]
        img_path = Path(test_images[i])
        image, affine, _header = nifti.read(img_path)  # [C, X, Y, Z] float32, RAS affine
        cm = None
        with torch.cuda.device(net_d.device):
            if have_labels:
                ref_lab, _, _ = nifti.read(Path(test_labels[i]))
                if tuple(ref_lab.shape[1:]) != tuple(image.shape[1:]):
                    raise ValueError(f"label {test_labels[i]} has shape {tuple(ref_lab.shape[1:])}, image {tuple(image.shape[1:])}")
                lab_dev, pred_net, label_net = predict_volume(
                    net_d, torch.from_numpy(image), affine, spacing, overlap=overlap, mode=mode,
                    sw_batch_size=sw_batch_size, precision=precision, invert=invert, return_device=True,
                    label=torch.from_numpy(np.ascontiguousarray(ref_lab[0])))
                # scored on the pre-processed (cropped, re-spaced) grid, before Invertd -- as the reference does (:672-680)
                cm = E.confusion_matrix(num_classes, pred_net.contiguous(), label_net.contiguous())   # one device pass
            else:
                lab_dev = predict_volume(net_d, torch.from_numpy(image), affine, spacing, overlap=overlap, mode=mode,
                                         sw_batch_size=sw_batch_size, precision=precision, invert=invert,
                                         return_device=True)
            lab_host = lab_dev.cpu().numpy().astype(np.float32)
        return img_path, affine, lab_host, cm

    if len(nets) > 1 and len(test_images) > 1:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=len(nets))
        futures = [pool.submit(run_one, i) for i in range(len(test_images))]
        results = (f.result() for f in futures)
    else:
        pool = None
        results = (run_one(i) for i in range(len(test_images)))
    for img_path, affine, lab_host, cm in results:
        name = img_path.name
        for ext in (".nii.gz", ".nii", ".nrrd", ".mha", ".mhd"):
            if name.endswith(ext):
                name = name[: -len(ext)]
                break
        if output_dir:
            nifti.write(Path(output_dir) / f"{name}.nii.gz", lab_host, affine)
        if have_labels:
            dice = E.class_dice(cm, include_background=False)
            valid = ~np.isnan(dice)
            class_dice_sum[valid] += dice[valid]
            class_dice_cnt[valid] += 1
            all_counts.append(E.confusion_counts(cm))
            image_mean_dice.append(float(np.nanmean(dice)) if valid.any() else float("nan"))
            print("Mean Dice: ", np.mean(dice))   # np.mean as the reference (:684): NaN when a tissue is absent from the label
            print("Class Dice:")
            print_table(tissue_names[1:], dice)
            all_mean_dice.append(float(np.nanmean(image_mean_dice)))
            if output_dir:
                # the reference plots <base>_confusion.png (matplotlib, out of scope); the matrix itself is kept
                np.savetxt(Path(output_dir) / f"{name}_confusion.csv", cm, fmt="%d", delimiter=",")
    if pool is not None:
        pool.shutdown()
    if output_dir is None:
        print("No output path specified, dice scores won't be saved.")
    else:
        np.savetxt(Path(output_dir) / f"mean_dice_{model_file.stem}_generalized_score.txt", all_mean_dice, delimiter=",")
    if have_labels:
        print("*" * 80)
        print("Total Mean Dice: ", all_mean_dice[-1] if all_mean_dice else float("nan"))
        print("Total Class Dice:")
        with np.errstate(invalid="ignore", divide="ignore"):
            print_table(tissue_names[1:], class_dice_sum / class_dice_cnt)
        print("Total Conf. Matrix Metrics:")
        metrics = E.confusion_metrics(all_counts)
        print_table(list(E.CONFUSION_METRICS), [metrics[k] for k in E.CONFUSION_METRICS])



def ensemble_creator(model_files: List[Path], test_images: List[Path], test_labels: Optional[List[Path]] = None,
                     output_dir: Path = None, tissue_dict: Dict[str, int] = None, spacing: Sequence[float] = [],
                     combination_mode: str = "select_best", candidate_per_tissue_path: Optional[Path] = None,
                     gpu_ids: List[int] = [], *, precision: str = "fp32") -> None:
    """Same signature as ``segmantic.seg.monai_unet.ensemble_creator`` (``:848-1004``): every model predicts every image
    with ``SlidingWindowInferer(roi_size=(96, 96, 96), sw_batch_size=4, overlap=0.5)`` (``:834-846``; here the roi is the
    models' own ``spatial_size``), the predictions are combined voxel-wise on the device and written as
    ``<output_dir>/<image basename>_seg.nii.gz``.

    * ``mean``: ``MeanEnsembled(weights=...)`` then argmax; the weights are parsed from the checkpoint names as the
      reference does (``float(stem.split("-")[-1].split("=")[1])``, e.g. ``epoch=12-val_dice=0.91.ckpt``).
    * ``vote``: argmax per model, majority vote (ties -> lowest class).
    * ``select_best``: argmax per model; ``candidate_per_tissue_path`` (yaml / json ``{tissue name: model index}``)
      picks, per tissue, the model whose voxels of that tissue are kept (``seg/transforms.py:15-61``).

    Deviation: with ``spacing`` the combined LABEL map is resampled back nearest-neighbour; the reference runs
    ``Invertd(nearest_interp=False)`` -- trilinear interpolation of label values -- on it."""
    from ..image import nifti
    from . import ensemble as ENS

    mode_name = getattr(combination_mode, "value", combination_mode)
    if mode_name not in ("mean", "vote", "select_best"):
        raise ValueError(f"unknown combination mode {combination_mode!r}")
    if mode_name == "select_best" and candidate_per_tissue_path is None:
        raise ValueError("When using the 'select_best'-mode, candidate_per_tissue_path needs to be specified.")
    device = make_device(gpu_ids)
    if device.type != "cuda":
        raise RuntimeError("segmantic_b200 has no CPU path: pass gpu_ids=[] or a non-negative GPU id")
    model_files = [Path(f) for f in model_files]
    if not model_files:
        raise ValueError("ensemble_creator needs at least one model file")
    models = []
    for f in model_files:
        settings = {}
        if f.with_suffix(".json").exists():
            settings = json.loads(f.with_suffix(".json").read_text())
        m = Net.load_from_checkpoint(str(f), **settings)
        m.freeze()
        m.eval()
        m.to(device)
        models.append(m)
    num_classes = models[0].num_classes
    spec = {"nets": models, "mode": mode_name, "num_classes": num_classes}
    if mode_name == "mean":
        spec["weights"] = [float(f.stem.split("-")[-1].split("=")[1]) for f in model_files]
    elif mode_name == "select_best":
        if tissue_dict is None:
            raise RuntimeError("'select_best' mode requires a tissue list")
        text = Path(candidate_per_tissue_path).read_text()
        if Path(candidate_per_tissue_path).suffix.lower() == ".json":
            name_model = json.loads(text)
        else:
            import yaml
            name_model = yaml.safe_load(text)
        spec["pairs"] = [(int(tissue_dict[name]), int(model_id)) for name, model_id in name_model.items()]
    if test_labels:
        assert len(test_images) == len(test_labels)
    if output_dir:
        os.makedirs(output_dir, exist_ok=True)
    for img_path in test_images:
        img_path = Path(img_path)
        image, affine, header = nifti.read(img_path)
        lab = predict_volume(models[0], torch.from_numpy(image), affine, spacing, overlap=0.5, mode="constant",
                             sw_batch_size=4, precision=precision, invert="labels", ensemble=spec)
        if output_dir:
            name = img_path.name
            for ext in (".nii.gz", ".nii", ".nrrd", ".mha", ".mhd"):
                if name.endswith(ext):
                    name = name[: -len(ext)]
                    break
            nifti.write(Path(output_dir) / f"{name}_seg.nii.gz", lab.numpy().astype(np.float32), affine)
