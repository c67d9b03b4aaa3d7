"""Multi-GPU driver of the prediction path: one process per GPU, axis-0 output slabs with ROI halos.

The reference predicts on a single device (``/root/reference/src/segmantic/seg/utils.py:4-12`` picks
``gpu_ids[0]``); this is the new data-parallel driver BASELINE.json's north_star asks for.  The
network-grid volume is cut along its slowest axis into ``world_size`` output slabs
(``sliding_window.slab_partition``); a rank runs every window row that intersects its slab, in MONAI's
window order, and blends / normalises / argmaxes only its own planes -- so each voxel sees exactly the
single-GPU sequence of fp32 additions and the labels are bit-identical.  Slabs are independent (no
data-path collective); NCCL (gloo in the CPU tests) is used only to gather the uint8 label slabs.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .sliding_window import Schedule, make_schedule, slab_partition, window_partition


def rank_slab(global_size: Sequence[int], roi: Sequence[int], overlap: float, mode: str, rank: int,
              world_size: int, sigma_scale: float = 0.125):
    """(schedule, all slabs, this rank's slab) for a volume that is at least ROI-sized everywhere."""
    sched = make_schedule(tuple(global_size), tuple(roi), overlap, mode, sigma_scale)
    parts = slab_partition(sched, world_size)
    return sched, parts, parts[rank]


def rank_windows(global_size: Sequence[int], roi: Sequence[int], overlap: float, mode: str, rank: int,
                 world_size: int, sigma_scale: float = 0.125):
    """(schedule, all parts, this rank's part) of the window-ownership partition (``window_partition``): every
    window is computed once; a rank receives the windows of rank-1 that cover its planes (NCCL point-to-point over
    NVLink) -- the near-linear form of the driver.  Falls back to ``slab_partition`` semantics via ``rank_slab``
    when a rank would need windows of two earlier ranks."""
    sched = make_schedule(tuple(global_size), tuple(roi), overlap, mode, sigma_scale)
    parts = window_partition(sched, world_size)
    return sched, parts, parts[rank]


def gather_label_slabs(local: torch.Tensor, parts: List[dict], dst: int = 0,
                       group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Gather uint8 label slabs ``[nx_r, Y, Z]`` (unequal heights) on ``dst`` -> ``[X, Y, Z]``.

    ``dist.gather`` wants equal shapes, so slabs are padded to the tallest one; 1 byte per voxel
    crosses NVLink once (268 MB for a 512x512x1024 volume)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    heights = [p["x1"] - p["x0"] for p in parts]
    maxh = max(heights)
    buf = torch.zeros((maxh,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([outs[r][: heights[r]] for r in range(world)], dim=0)


def predict_labels_distributed(volume_planes: Callable[[int, int], torch.Tensor], global_size: Sequence[int],
                               roi: Sequence[int], net, *, overlap: float = 0.25, mode: str = "constant",
                               sw_batch_size: int = 4, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Slab-parallel sliding-window prediction.  ``volume_planes(x0, x1)`` returns this rank's input planes
    ``[Cin, x1-x0, Y, Z]`` (float32, on the rank's GPU); returns the full uint8 label map on ``dst``."""
    from .engine import sliding_window_inference_slab

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sched, parts, part = rank_slab(global_size, roi, overlap, mode, rank, world)
    if part["x1"] > part["x0"]:
        vol = volume_planes(part["vol_x0"], part["vol_x1"])
        local = sliding_window_inference_slab(vol, global_size, part, roi, sw_batch_size, net, overlap=overlap,
                                              mode=mode)["labels"]
    else:
        local = torch.empty((0, global_size[1], global_size[2]), dtype=torch.uint8, device=net.device)
    return gather_label_slabs(local, parts, dst, group)
