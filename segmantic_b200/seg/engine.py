"""Device engine: the MONAI-UNet + sliding-window operators, executed by the sm_100a library.

``UNetB200`` stands where ``Net._model`` (``monai.networks.nets.UNet``, eval mode) stands in the
reference (``/root/reference/src/segmantic/seg/monai_unet.py:114-124,221-222``) and
``sliding_window_inference`` where ``SlidingWindowInferer(...)(image, net)`` does (``:637-639,665``).
PyTorch supplies device buffers and the stream only; all arithmetic is in ``libsegmantic_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence, Tuple

import torch

from .. import _lib
from .sliding_window import Schedule, make_schedule
from .unet_spec import (KIND_IDENTITY, fold_batchnorm, unet_conv_specs)


DEVICE_SW_BATCH = 128  # windows per network launch on the device (memory: ~55 MB / window in bf16); fewer, fuller waves of the persistent kernels


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("segmantic_b200 runs on CUDA devices only (no CPU fallback); "
                           f"got device {device}")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device available: segmantic_b200 has no CPU fallback")
    return device


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _precision_code(precision: str) -> int:
    p = str(precision).lower()
    if p in ("fp32", "float32", "32"):
        return _lib.PRECISION_FP32
    if p in ("bf16", "bfloat16"):
        return _lib.PRECISION_BF16
    raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")


class UNetB200:
    """Frozen, eval-mode MONAI UNet (num_res_units=2, BATCH norm folded, PReLU) on one B200."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], *, spatial_dims: int = 3, in_channels: int = 1,
                 out_channels: int, channels=(16, 32, 64, 128, 256), strides=(2, 2, 2, 2),
                 device="cuda:0", precision: str = "fp32"):
        self.device = _require_cuda(device)
        self.spatial_dims, self.in_channels, self.out_channels = int(spatial_dims), int(in_channels), int(out_channels)
        self.channels, self.strides = tuple(int(c) for c in channels), tuple(int(s) for s in strides)
        self.precision = str(precision).lower()
        self._lib = _lib.load()
        specs = unet_conv_specs(self.in_channels, self.out_channels, self.channels, self.strides)
        folded = fold_batchnorm(state_dict, specs)
        n = len(self.channels) - 1
        convs = (_lib.ConvDesc * len(folded))()
        keep = []  # keep host tensors alive until create returns
        for i, f in enumerate(folded):
            d = convs[i]
            d.kind, d.cin, d.cout = f.spec.kind, f.spec.cin, f.spec.cout
            d.kernel, d.stride = f.spec.kernel, f.spec.stride
            d.has_act, d.alpha = int(f.spec.has_adn), float(f.alpha)
            if f.spec.kind != KIND_IDENTITY:
                w = f.weight.contiguous().cpu()
                b = f.bias.contiguous().cpu()
                keep += [w, b]
                d.weight = C.cast(w.data_ptr(), C.POINTER(C.c_float))
                d.bias = C.cast(b.data_ptr(), C.POINTER(C.c_float))
        desc = _lib.UnetDesc()
        desc.spatial_dims, desc.in_channels, desc.out_channels = self.spatial_dims, self.in_channels, self.out_channels
        desc.n_levels = n
        for i, c in enumerate(self.channels):
            desc.channels[i] = c
        for i, s in enumerate(self.strides[:n]):
            desc.strides[i] = s
        desc.precision = _precision_code(precision)
        desc.n_convs, desc.convs = len(folded), convs
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sgm_unet_create(C.byref(desc), C.byref(handle)), "sgm_unet_create")
        self._handle = handle
        self._ws: Optional[torch.Tensor] = None
        del keep

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                self._lib.sgm_unet_destroy(h)
            except Exception:
                pass
            self._handle = None

    # -- helpers
    def roi3(self, roi: Sequence[int]) -> Tuple[int, int, int]:
        roi = tuple(int(r) for r in roi)
        if self.spatial_dims == 2:
            roi = roi[-2:]
            return (1,) + roi
        if len(roi) != 3:
            raise ValueError(f"roi must have 3 entries for a 3-D network, got {roi}")
        return roi

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def _buffer(self, name: str, nelem: int) -> torch.Tensor:
        """Cached float32 device buffer (grown on demand) for the multi-GPU window exchange."""
        bufs = self.__dict__.setdefault("_bufs", {})
        b = bufs.get(name)
        if b is None or b.numel() < nelem:
            bufs[name] = None
            b = bufs[name] = torch.empty(int(nelem), dtype=torch.float32, device=self.device)
        return b

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.sgm_unet_last_launch_count(self._handle))

    # -- Net.forward
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``[B, Cin, *roi]`` float32 CUDA tensor -> ``[B, C, *roi]`` float32 logits."""
        if x.device != self.device or x.dtype != torch.float32:
            raise ValueError("input must be a float32 tensor on the network's CUDA device")
        if x.dim() != self.spatial_dims + 2 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected [B, {self.in_channels}, *spatial({self.spatial_dims})], got {tuple(x.shape)}")
        x = x.contiguous()
        B = x.shape[0]
        roi = self.roi3(x.shape[2:])
        out = torch.empty((B, self.out_channels) + tuple(x.shape[2:]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            need = self._lib.sgm_unet_workspace_bytes(self._handle, _lib.i3(roi), B)
            _lib.check(need, "sgm_unet_workspace_bytes")
            ws = self._workspace(need)
            _lib.check(self._lib.sgm_unet_forward(self._handle, x.data_ptr(), out.data_ptr(), B, _lib.i3(roi),
                                                  ws.data_ptr(), ws.numel(), _stream_ptr(self.device)),
                       "sgm_unet_forward")
        return out

    def check(self) -> None:
        """Synchronise and raise if a tcgen05 pipeline reported a timeout (bf16 path)."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sgm_unet_check(self._handle, _stream_ptr(self.device)), "sgm_unet_check")

    __call__ = forward

    def set_profiling(self, on: bool) -> None:
        _lib.check(self._lib.sgm_unet_set_profiling(self._handle, int(bool(on))), "sgm_unet_set_profiling")

    def get_profile(self):
        """[(role, ms, launches)] per convolution since the last call (device time, CUDA events)."""
        specs = unet_conv_specs(self.in_channels, self.out_channels, self.channels, self.strides)
        n = len(specs) + 1  # + the gather-blend kernel
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sgm_unet_get_profile(self._handle, ms, cnt, n, _stream_ptr(self.device)),
                       "sgm_unet_get_profile")
        roles = [sp.role for sp in specs] + ["gather_blend"]
        return [(roles[i], float(ms[i]), int(cnt[i])) for i in range(n)]


def _make_cfg(sched: Schedule, sw_batch: int, a0=None, vol=None, acc=None):
    cfg = _lib.SwCfg()
    keep = []
    for a in range(3):
        cfg.dims[a] = sched.padded_size[a]
        cfg.roi[a] = sched.roi[a]
        n = len(sched.starts[a])
        if n > _lib.SGM_MAX_STARTS:
            raise ValueError(f"too many window starts along axis {a}: {n} > {_lib.SGM_MAX_STARTS}")
        cfg.n_starts[a] = n
        for j, s in enumerate(sched.starts[a]):
            cfg.starts[a][j] = s
        t = sched.tables[a].contiguous().to(torch.float32).cpu()
        keep.append(t)
        cfg.imap[a] = C.cast(t.data_ptr(), C.POINTER(C.c_float))
    cfg.imap_floor = float(sched.floor)
    cfg.sw_batch = int(sw_batch)
    cfg.a0_begin, cfg.a0_end = a0 if a0 is not None else (0, len(sched.starts[0]))
    cfg.vol_x0, cfg.vol_nx = vol if vol is not None else (0, sched.padded_size[0])
    cfg.acc_x0, cfg.acc_nx = acc if acc is not None else (0, sched.padded_size[0])
    return cfg, keep


def _sw_run(net: UNetB200, vol: torch.Tensor, sched: Schedule, sw_batch_size: int, a0, vol_rng, acc_rng,
            return_logits: bool, return_labels: bool, return_probs: bool):
    """Accumulate + finalise planes ``acc_rng`` from ``vol`` = planes ``vol_rng`` of the padded volume.
    Returns ``(logits, labels, probs)`` device tensors (or None) of shape ``[C|-, nx, Y, Z]``."""
    lib = net._lib
    C_out = net.out_channels
    x0, nx = acc_rng
    plane = (sched.padded_size[1], sched.padded_size[2])
    # The reference's sw_batch_size (4) only bounds MONAI's activation memory; results do not depend on
    # it (eval-mode network, windows are independent).  The device path batches more windows per launch
    # so that the small deep layers fill the 148 SMs (SGM_SW_BATCH overrides).
    device_batch = max(int(sw_batch_size), int(os.environ.get("SGM_SW_BATCH", DEVICE_SW_BATCH)))
    cfg, keep = _make_cfg(sched, device_batch, a0, vol_rng, (x0, nx))
    blend = os.environ.get("SGM_BLEND", "auto")
    with torch.cuda.device(net.device):
        st = _stream_ptr(net.device)
        need = lib.sgm_sw_predict_workspace_bytes(net._handle, C.byref(cfg))
        _lib.check(need, "sgm_sw_predict_workspace_bytes")
        free_b, _total = torch.cuda.mem_get_info(net.device)
        have = net._ws.numel() if net._ws is not None else 0
        deferred = blend == "gather" or (blend == "auto" and need <= have + int(free_b * 0.8))
    if deferred:  # deferred (gather) blend: no read-modify-write, bit-identical result
        with torch.cuda.device(net.device):
            ws = net._workspace(need)
            logits = torch.empty((C_out, nx) + plane, dtype=torch.float32, device=net.device) if return_logits else None
            labels = torch.empty((nx,) + plane, dtype=torch.uint8, device=net.device) if return_labels else None
            probs = torch.empty((C_out, nx) + plane, dtype=torch.float32, device=net.device) if return_probs else None
            _lib.check(lib.sgm_sw_predict(net._handle, vol.data_ptr(), C.byref(cfg),
                                          logits.data_ptr() if logits is not None else None,
                                          labels.data_ptr() if labels is not None else None,
                                          probs.data_ptr() if probs is not None else None,
                                          ws.data_ptr(), ws.numel(), st), "sgm_sw_predict")
        del keep
        return logits, labels, probs
    full_call = (tuple(a0) == (0, len(sched.starts[0])) and tuple(vol_rng) == (0, sched.padded_size[0])
                 and (x0, nx) == (0, sched.padded_size[0]))
    if full_call and blend in ("auto", "chunked") and not return_probs:
        # The deferred buffer of the whole volume does not fit (e.g. BASELINE configs[3] on ONE GPU: 2100 windows x 20
        # classes x 96^3 x 4 B = 148 GB): run the window list in CHUNKS, each with its own deferred blend -- the
        # window-ownership partition of the multi-GPU driver executed sequentially on one device (the "exchange" is a
        # device copy of the seam windows).  Bit-identical to the one-shot form; no read-modify-write.
        # memory this call may use: what is free now plus what the network already holds from earlier calls (its
        # workspace and the two chunk exchange buffers are cached on the engine and reused)
        cached = sum(b.numel() * 4 for k_, b in net.__dict__.get("_bufs", {}).items() if k_.startswith("wlchunk") and b is not None)
        out = _sw_run_chunked(net, vol, sched, sw_batch_size, return_logits, return_labels,
                              int(free_b * 0.8) + have + cached)
        if out is not None:
            del keep
            return out
    acc = torch.zeros((C_out, nx) + plane, dtype=torch.float32, device=net.device)
    with torch.cuda.device(net.device):
        need = lib.sgm_sw_workspace_bytes(net._handle, C.byref(cfg))
        _lib.check(need, "sgm_sw_workspace_bytes")
        ws = net._workspace(need)
        _lib.check(lib.sgm_sw_accumulate(net._handle, vol.data_ptr(), C.byref(cfg), acc.data_ptr(),
                                         ws.data_ptr(), ws.numel(), st), "sgm_sw_accumulate")
        logits = torch.empty_like(acc) if return_logits else None
        labels = torch.empty((nx,) + plane, dtype=torch.uint8, device=net.device) if return_labels else None
        probs = torch.empty_like(acc) if return_probs else None
        _lib.check(lib.sgm_sw_finalize(acc.data_ptr(), C_out, C.byref(cfg),
                                       logits.data_ptr() if logits is not None else None,
                                       labels.data_ptr() if labels is not None else None,
                                       probs.data_ptr() if probs is not None else None, st),
                   "sgm_sw_finalize")
    del keep
    return logits, labels, probs


def _sw_run_chunked(net: UNetB200, vol: torch.Tensor, sched: Schedule, sw_batch_size: int, return_logits: bool,
                    return_labels: bool, budget_bytes: int):
    """Deferred blend in chunks of the window list (see ``_sw_run``); ``None`` when no chunking fits ``budget_bytes``."""
    from .sliding_window import plan_chunks

    roivox = int(sched.roi[0]) * int(sched.roi[1]) * int(sched.roi[2])
    stride = net.out_channels * roivox * 4
    device_batch = max(int(sw_batch_size), int(os.environ.get("SGM_SW_BATCH", DEVICE_SW_BATCH)))
    with torch.cuda.device(net.device):
        net_ws = int(net._lib.sgm_unet_workspace_bytes(net._handle, _lib.i3(sched.roi), min(device_batch, sched.n_windows)))
    # two exchange buffers alternate (chunk r blends from buffer r % 2 while chunk r + 1's seam is copied in)
    parts = plan_chunks(sched, stride, net_ws + (256 << 20), budget_bytes)
    if parts is None:
        return None
    size3 = sched.padded_size
    logits, labels, launches = [], [], 0
    prev = None
    for r, part in enumerate(parts):
        if part["w_hi"] <= part["w_lo"]:
            continue
        run = OwnedWindows(vol[:, part["vol_x0"]:part["vol_x1"]], size3, part, sched.roi, sw_batch_size, net,
                           0.0, "constant", tag=f"chunk{r % 2}", sched=sched)  # (overlap / mode live in `sched`)
        if prev is not None and run.recv_view.numel():
            run.recv_view.copy_(prev.send_view)
        run.compute_tail()
        run.compute_rest()
        out = run.blend(return_logits)
        launches += run.launches
        if return_logits:
            logits.append(out["logits"])
        labels.append(out["labels"])
        prev = run
    net.last_launch_count_chunked = launches
    return (torch.cat(logits, dim=1) if return_logits else None, torch.cat(labels, dim=0) if return_labels else None, None)


def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: UNetB200, overlap: float = 0.25, mode: str = "constant",
                             sigma_scale: float = 0.125, *, return_labels: bool = False,
                             return_probs: bool = False, return_logits: bool = True, slab: Optional[dict] = None):
    """MONAI ``sliding_window_inference`` semantics on the device.

    ``inputs`` is ``[1, Cin, *spatial]`` float32 on the predictor's device.  Returns the blended
    logits ``[1, C, *spatial]`` (and/or ``labels`` uint8 ``[1, 1, *spatial]`` = argmax, ties -> lowest
    class; ``probs`` = softmax of the blended logits) as a dict when more than the logits are asked for.
    With ``slab`` (from ``sliding_window.slab_partition``) only the rank's planes are computed and
    returned (axis 0 of the result covers ``[slab.x0, slab.x1)``).
    """
    net = predictor
    # A 2-D network also takes a STACK of slices [1, Cin, Z, X, Y] (BASELINE configs[4]: slice-wise prediction): every
    # slice is an independent 2-D image -- one schedule with roi (1, h, w) over (Z, X, Y), all slices' windows batched
    # into the same network launches; results per slice are those of the 2-D call.
    stack = net.spatial_dims == 2 and inputs.dim() == 5
    if (inputs.dim() != net.spatial_dims + 2 and not stack) or inputs.shape[0] != 1:
        raise ValueError(f"inputs must be [1, Cin, *spatial], got {tuple(inputs.shape)}")
    if inputs.device != net.device or inputs.dtype != torch.float32:
        raise ValueError("inputs must be float32 on the predictor's CUDA device")
    if stack and inputs.shape[2] > _lib.SGM_MAX_STARTS:  # the schedule holds at most SGM_MAX_STARTS starts per axis
        if slab is not None:
            raise ValueError("slab partitions of a slice stack are not supported: partition the stack itself")
        parts = [sliding_window_inference(inputs[:, :, z:z + _lib.SGM_MAX_STARTS], roi_size, sw_batch_size, predictor,
                                          overlap, mode, sigma_scale, return_labels=return_labels,
                                          return_probs=return_probs, return_logits=return_logits)
                 for z in range(0, inputs.shape[2], _lib.SGM_MAX_STARTS)]
        if isinstance(parts[0], dict):
            return {k: torch.cat([p_[k] for p_ in parts], dim=2) for k in parts[0]}
        return torch.cat(parts, dim=2)
    spatial = tuple(inputs.shape[2:])
    size3 = (1,) + spatial if (net.spatial_dims == 2 and not stack) else spatial
    roi3 = net.roi3(roi_size)
    sched = make_schedule(size3, roi3, overlap, mode, sigma_scale)
    vol = inputs[0].reshape((net.in_channels,) + size3)
    if sched.padded_size != size3:  # symmetric zero padding up to the roi (cropped again below)
        pad = []
        for a in (2, 1, 0):
            lo = sched.pad_lo[a]
            pad += [lo, sched.padded_size[a] - size3[a] - lo]
        vol = torch.nn.functional.pad(vol, pad, mode="constant", value=0.0)
    vol = vol.contiguous()
    C_out = net.out_channels
    if slab is None:
        x0, x1 = 0, sched.padded_size[0]
        a0 = (0, len(sched.starts[0]))
        vol_rng = (0, sched.padded_size[0])
    else:
        x0, x1 = int(slab["x0"]), int(slab["x1"])
        a0 = (int(slab["a0_begin"]), int(slab["a0_end"]))
        vol_rng = (int(slab["vol_x0"]), int(slab["vol_x1"]) - int(slab["vol_x0"]))
        vol = vol[:, slab["vol_x0"]:slab["vol_x1"]].contiguous()
    nx = x1 - x0
    plane = (sched.padded_size[1], sched.padded_size[2])
    out: dict = {}
    if nx <= 0:
        empty = (0,) + plane
        if return_logits:
            out["logits"] = torch.empty((1, C_out) + empty, device=net.device)
        if return_labels:
            out["labels"] = torch.empty((1, 1) + empty, dtype=torch.uint8, device=net.device)
        if return_probs:
            out["probs"] = torch.empty((1, C_out) + empty, device=net.device)
        return out if (return_labels or return_probs) else out["logits"]
    logits, labels, probs = _sw_run(net, vol, sched, sw_batch_size, a0, vol_rng, (x0, nx), return_logits,
                                    return_labels, return_probs)

    def crop(t, lead):
        # undo the roi padding (axes 1, 2 always; axis 0 only without a slab)
        sl = [slice(None)] * lead
        for a in range(3):
            lo = sched.pad_lo[a]
            if a == 0 and slab is not None:
                sl.append(slice(None))
            else:
                sl.append(slice(lo, lo + size3[a]))
        t = t[tuple(sl)]
        if net.spatial_dims == 2 and not stack:
            t = t.squeeze(lead)
        return t.unsqueeze(0)

    if logits is not None:
        out["logits"] = crop(logits, 1)
    if labels is not None:
        out["labels"] = crop(labels.unsqueeze(0), 1)
    if probs is not None:
        out["probs"] = crop(probs, 1)
    if return_labels or return_probs:
        return out
    return out["logits"]


def sliding_window_inference_slab(vol_slab: torch.Tensor, global_size: Sequence[int], slab: dict,
                                  roi_size: Sequence[int], sw_batch_size: int, predictor: UNetB200,
                                  overlap: float = 0.25, mode: str = "constant", sigma_scale: float = 0.125,
                                  *, return_logits: bool = False):
    """Multi-GPU form: this rank holds only planes ``[slab.vol_x0, slab.vol_x1)`` of axis 0 of a
    ``[Cin, *global_size]`` volume (``vol_slab``: ``[Cin, vol_nx, Y, Z]``) and computes the label map
    (and optionally the logits) of its output planes ``[slab.x0, slab.x1)``; the volume must already be
    at least ROI-sized along every axis.  Results equal the single-GPU run bit for bit."""
    net = predictor
    size3 = tuple(int(s) for s in global_size)
    sched = make_schedule(size3, net.roi3(roi_size), overlap, mode, sigma_scale)
    if sched.padded_size != size3:
        raise ValueError("slab execution needs a volume at least as large as the roi along every axis")
    vol_nx = int(slab["vol_x1"]) - int(slab["vol_x0"])
    if tuple(vol_slab.shape) != (net.in_channels, vol_nx, size3[1], size3[2]):
        raise ValueError(f"vol_slab must be {(net.in_channels, vol_nx, size3[1], size3[2])}, got {tuple(vol_slab.shape)}")
    nx = int(slab["x1"]) - int(slab["x0"])
    logits, labels, _ = _sw_run(net, vol_slab.contiguous(), sched, sw_batch_size,
                                (int(slab["a0_begin"]), int(slab["a0_end"])), (int(slab["vol_x0"]), vol_nx),
                                (int(slab["x0"]), nx), return_logits, True, False)
    out = {"labels": labels}
    if logits is not None:
        out["logits"] = logits
    return out


class OwnedWindows:
    """One rank's share of a window-ownership run (``sliding_window.window_partition``), in three steps so that the
    exchange can be NCCL point-to-point (``sliding_window_inference_owned``) or, in the single-GPU test, a plain copy
    between two simulated ranks: ``compute_tail()`` (the windows the next rank also needs) -> send ``send_view`` /
    receive into ``recv_view`` -> ``compute_rest()`` -> ``blend()``."""

    def __init__(self, vol_slab: torch.Tensor, global_size: Sequence[int], part: dict, roi_size: Sequence[int],
                 sw_batch_size: int, net: UNetB200, overlap: float, mode: str, sigma_scale: float = 0.125,
                 tag: str = "", sched: Optional[Schedule] = None):
        size3 = tuple(int(s) for s in global_size)
        if sched is None:
            sched = make_schedule(size3, net.roi3(roi_size), overlap, mode, sigma_scale)
        if sched.padded_size != size3:
            raise ValueError("owned-window execution needs a volume at least as large as the roi along every axis")
        self.net, self.sched, self.part = net, sched, part
        self.w_lo, self.w_hi, self.wb, self.send_lo = (int(part[k]) for k in ("w_lo", "w_hi", "wb", "send_lo"))
        vol_nx = int(part["vol_x1"]) - int(part["vol_x0"])
        if tuple(vol_slab.shape) != (net.in_channels, vol_nx, size3[1], size3[2]):
            raise ValueError(f"vol_slab must be {(net.in_channels, vol_nx, size3[1], size3[2])}, got {tuple(vol_slab.shape)}")
        self.vol = vol_slab.contiguous()
        roivox = int(sched.roi[0]) * int(sched.roi[1]) * int(sched.roi[2])
        self.stride = net.out_channels * roivox
        self.nx = int(part["x1"]) - int(part["x0"])
        self.plane = (size3[1], size3[2])
        device_batch = max(int(sw_batch_size), int(os.environ.get("SGM_SW_BATCH", DEVICE_SW_BATCH)))
        self.cfg, self._keep = _make_cfg(
            sched, device_batch, (int(part["b_begin"]), max(int(part["b_end"]), int(part["b_begin"]))),
            (int(part["vol_x0"]), max(vol_nx, 1)), (int(part["x0"]) if self.nx > 0 else 0, max(self.nx, 1)))
        # weighted logits of windows [wb, w_hi): the received part first, the own windows after it
        self.wl = net._buffer("wl" + tag, max(self.w_hi - self.wb, 1) * self.stride)
        self.launches = 0

    @property
    def recv_view(self) -> torch.Tensor:
        return self.wl[: (self.w_lo - self.wb) * self.stride]

    @property
    def send_view(self) -> torch.Tensor:
        return self.wl[(self.send_lo - self.wb) * self.stride: (self.w_hi - self.wb) * self.stride]

    def _compute(self, first: int, count: int):
        if count <= 0:
            return
        net, lib = self.net, self.net._lib
        with torch.cuda.device(net.device):
            need = lib.sgm_sw_windows_workspace_bytes(net._handle, C.byref(self.cfg))
            _lib.check(need, "sgm_sw_windows_workspace_bytes")
            ws = net._workspace(need)
            off = (first - self.wb) * self.stride
            _lib.check(lib.sgm_sw_windows(net._handle, self.vol.data_ptr(), C.byref(self.cfg), first, count,
                                          self.wl[off:].data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(net.device)),
                       "sgm_sw_windows")
            self.launches += int(lib.sgm_unet_last_launch_count(net._handle))

    def compute_tail(self):
        self._compute(self.send_lo, self.w_hi - self.send_lo)

    def compute_rest(self):
        self._compute(self.w_lo, self.send_lo - self.w_lo)

    def blend(self, return_logits: bool = False):
        net, lib, sched = self.net, self.net._lib, self.sched
        C_out = net.out_channels
        logits = None
        if self.nx <= 0:
            return {"labels": torch.empty((0,) + self.plane, dtype=torch.uint8, device=net.device)}
        with torch.cuda.device(net.device):
            if return_logits:
                logits = torch.empty((C_out, self.nx) + self.plane, dtype=torch.float32, device=net.device)
            labels = torch.empty((self.nx,) + self.plane, dtype=torch.uint8, device=net.device)
            scratch = net._buffer("blend_scratch", 4096)
            per_row = len(sched.starts[1]) * len(sched.starts[2])
            wl_rows = self.wl[(int(self.part["b_begin"]) * per_row - self.wb) * self.stride:]
            _lib.check(lib.sgm_sw_blend(C.byref(self.cfg), C_out, wl_rows.data_ptr(),
                                        logits.data_ptr() if logits is not None else None, labels.data_ptr(), None,
                                        scratch.data_ptr(), _stream_ptr(net.device)), "sgm_sw_blend")
        self.launches += 1
        out = {"labels": labels}
        if logits is not None:
            out["logits"] = logits
        return out


def _seam_link(net: UNetB200, run: "OwnedWindows", key: tuple, rank: int, world_size: int, group):
    """The rank's ``p2p.SeamLink`` for this geometry (built collectively on first use: every rank sees the same
    sequence of geometries, so the handle exchange stays matched across ranks)."""
    from .p2p import SeamLink

    link = net.__dict__.get("_seam")
    if link is None or link.key != key or link.recv_buf.data_ptr() != run.wl.data_ptr():
        if link is not None:
            link.close()
        link = SeamLink(run.wl, rank, world_size, group=group)
        link.key = key
        net.__dict__["_seam"] = link
    return link


def sliding_window_inference_owned(vol_slab: torch.Tensor, global_size: Sequence[int], part: dict,
                                   roi_size: Sequence[int], sw_batch_size: int, predictor: UNetB200,
                                   overlap: float = 0.25, mode: str = "constant", sigma_scale: float = 0.125,
                                   *, rank: int = 0, world_size: int = 1, group=None, return_logits: bool = False,
                                   exchange: Optional[str] = None):
    """Multi-GPU form with window OWNERSHIP (``sliding_window.window_partition``): this rank computes the
    importance-weighted logits of its own windows only, pushes the tail that also covers the next rank's planes
    over NVLink, and blends its output planes ``[part.x0, part.x1)`` in MONAI's window order -- bit-identical to the
    single-device result, and no window is computed twice.  ``vol_slab``: ``[Cin, part.vol_x1 - part.vol_x0, Y, Z]``
    float32 on the device.

    ``exchange="p2p"`` (default; ``SGM_SEAM`` overrides): peer-memory push with the copy engines and device-side
    counters (``p2p.SeamLink``), no collective library on the data path.  ``exchange="nccl"``: one NCCL point-to-point
    send / receive per seam."""
    import torch.distributed as dist

    exchange = exchange or os.environ.get("SGM_SEAM", "p2p")
    run = OwnedWindows(vol_slab, global_size, part, roi_size, sw_batch_size, predictor, overlap, mode, sigma_scale)
    if exchange == "p2p" and world_size > 1:
        key = (tuple(int(s) for s in global_size), tuple(int(r) for r in roi_size), float(overlap), str(mode),
               int(world_size), int(rank), predictor.out_channels)
        link = _seam_link(predictor, run, key, rank, world_size, group if not isinstance(group, (tuple, list)) else None)
        link.begin()
        run.compute_tail()       # the tail the next rank waits for goes first; it travels while the rest is computed
        link.push(run.send_view)
        run.compute_rest()
        if run.recv_view.numel() > 0:
            link.wait_data()
        out = run.blend(return_logits)
        if run.recv_view.numel() > 0:
            link.ack()
        predictor.last_launch_count_owned = run.launches
        return out
    # `group` may be a pair of process groups: the seam between ranks s and s+1 then uses group[s % 2], so that a
    # rank's receive (seam rank-1) and send (seam rank) live on different NCCL communicators.  On ONE communicator
    # unbatched point-to-point operations are serialised in posting order, which chains the seams across the ranks.
    seam_group = (lambda s: group[s % 2]) if isinstance(group, (tuple, list)) else (lambda s: group)
    reqs = []
    if rank > 0 and run.recv_view.numel() > 0:
        reqs.append(dist.irecv(run.recv_view, src=rank - 1, group=seam_group(rank - 1)))
    run.compute_tail()
    if rank + 1 < world_size and run.send_view.numel() > 0:
        reqs.append(dist.isend(run.send_view, dst=rank + 1, group=seam_group(rank)))
    run.compute_rest()
    for r in reqs:
        r.wait()
    out = run.blend(return_logits)
    predictor.last_launch_count_owned = run.launches
    return out


def debug_conv(net: UNetB200, conv_index: int, in0: torch.Tensor, in1: Optional[torch.Tensor] = None,
               res: Optional[torch.Tensor] = None, *, use_tc: bool, fused: bool = False, cg_out2: int = 0):
    """Run one convolution of the network on CG8 tensors ``[n, cg, d0, d1, d2, 8]`` (diagnostic).

    Returns ``out`` (and ``out2`` for a fused strided down block).  Used by the tests to pin the
    tcgen05 kernels against the CUDA-core kernels layer by layer.
    """
    lib = net._lib
    n, cg0 = in0.shape[0], in0.shape[1]
    dims = tuple(in0.shape[2:5])
    od = (C.c_int32 * 3)()
    # dry query of the output extent: run on the CUDA-core path needs the buffers, so compute here
    specs = unet_conv_specs(net.in_channels, net.out_channels, net.channels, net.strides)
    sp = specs[conv_index]
    flat = net.spatial_dims == 2
    out_dims = []
    for a_, d_ in enumerate(dims):
        if flat and a_ == 0:
            out_dims.append(d_)
        elif sp.kind == _lib.KIND_CONV_TRANSPOSE and sp.stride == 2:
            out_dims.append(d_ * 2)
        else:
            out_dims.append((d_ + 2 * (sp.kernel // 2) - sp.kernel) // sp.stride + 1)
    pad = 16 if net.precision == "bf16" else 8
    cg_out = -(-sp.cout // pad) * (pad // 8)
    out = torch.zeros((n, cg_out) + tuple(out_dims) + (8,), dtype=in0.dtype, device=in0.device)
    out2 = torch.zeros((n, cg_out2) + tuple(out_dims) + (8,), dtype=in0.dtype, device=in0.device) if fused else None
    with torch.cuda.device(net.device):
        _lib.check(lib.sgm_debug_conv(net._handle, conv_index, int(use_tc), int(fused), in0.data_ptr(), cg0,
                                      in1.data_ptr() if in1 is not None else None,
                                      in1.shape[1] if in1 is not None else 0,
                                      res.data_ptr() if res is not None else None, out.data_ptr(),
                                      out2.data_ptr() if out2 is not None else None, n, _lib.i3(dims), od,
                                      _stream_ptr(net.device)), "sgm_debug_conv")
        _lib.check(lib.sgm_unet_check(net._handle, _stream_ptr(net.device)), "sgm_unet_check")
    assert tuple(od) == tuple(out_dims), (tuple(od), out_dims)
    return (out, out2) if fused else out
