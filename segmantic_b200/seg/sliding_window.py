"""Host-side schedule of the sliding-window inferer (no device work here).

Mirrors what ``SlidingWindowInferer(roi_size=net.spatial_size, sw_batch_size=4, device=device)``
(``/root/reference/src/segmantic/seg/monai_unet.py:637-639``) computes before it calls the network:
symmetric zero padding up to the ROI, the scan interval, the per-axis window starts (last window
shifted back inside) and the importance map (constant, or MONAI >= 1.2's separable Gaussian with its
1e-3 floor).  Defaults are the reference's: overlap 0.25, mode "constant", sigma_scale 0.125.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, List, Sequence, Tuple

import torch


def scan_interval(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> Tuple[int, ...]:
    if not 0.0 <= overlap < 1.0:
        raise ValueError(f"overlap must be >= 0 and < 1, got {overlap}.")
    out = []
    for size, roi in zip(image_size, roi_size):
        out.append(int(roi) if roi == size else max(int(roi * (1 - overlap)), 1))
    return tuple(out)


def axis_starts(size: int, roi: int, interval: int) -> List[int]:
    num = int(math.ceil(float(size) / interval))
    for d in range(num):
        if d * interval + roi >= size:
            num = d + 1
            break
    starts = []
    for i in range(num):
        s = i * interval
        starts.append(s - max(s + roi - size, 0))
    return starts


def importance_tables(roi: Sequence[int], mode: str = "constant", sigma_scale: float = 0.125):
    """Per-axis 1-D tables and the floor; window weight = max((t0[i]*t1[j])*t2[k], floor)."""
    mode = str(mode).lower()
    if mode == "constant":
        return [torch.ones(int(r), dtype=torch.float32) for r in roi], 0.0
    if mode != "gaussian":
        raise ValueError(f"mode must be 'constant' or 'gaussian', got {mode!r}")
    tables = []
    for n in roi:
        n = int(n)
        sigma = sigma_scale * n
        x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float32)
        tables.append(torch.exp(x ** 2 / (-2 * sigma ** 2)))
    # floor = max(min of the separable product, 1e-3); the minimum sits at a corner
    prod = None
    for t in tables:
        m = torch.minimum(t[0], t[-1])
        prod = m if prod is None else prod * m
    # the product of per-axis minima equals the map's minimum only if multiplication order matches
    # MONAI's ((t0*t1)*t2) -- it does, all factors are the corner values.
    floor = max(float(prod), 1e-3)
    return tables, floor


@dataclass
class Schedule:
    image_size: Tuple[int, ...]      # original spatial size (3 axes; axis 0 == 1 for 2-D)
    padded_size: Tuple[int, ...]     # max(image, roi)
    pad_lo: Tuple[int, ...]
    roi: Tuple[int, ...]
    starts: List[List[int]]          # per axis
    tables: List[torch.Tensor] = field(repr=False, default_factory=list)
    floor: float = 0.0

    @property
    def n_windows(self) -> int:
        return len(self.starts[0]) * len(self.starts[1]) * len(self.starts[2])

    def windows(self):
        """All window starts in MONAI order (axis 0 slowest)."""
        return [(a, b, c) for a in self.starts[0] for b in self.starts[1] for c in self.starts[2]]


def make_schedule(image_size: Sequence[int], roi_size: Sequence[int], overlap: float = 0.25,
                  mode: str = "constant", sigma_scale: float = 0.125) -> Schedule:
    image_size = tuple(int(s) for s in image_size)
    roi = tuple(int(r) for r in roi_size)
    if len(image_size) != 3 or len(roi) != 3:
        raise ValueError("schedule works on 3 axes (use a leading axis of size 1 for 2-D)")
    padded = tuple(max(s, r) for s, r in zip(image_size, roi))
    pad_lo = tuple((p - s) // 2 for p, s in zip(padded, image_size))
    interval = scan_interval(padded, roi, overlap)
    starts = [axis_starts(padded[a], roi[a], interval[a]) for a in range(3)]
    tables, floor = importance_tables(roi, mode, sigma_scale)
    return Schedule(image_size, padded, pad_lo, roi, starts, tables, floor)


def slab_partition(schedule: Schedule, world_size: int) -> List[dict]:
    """Partition axis 0 into `world_size` output slabs with ROI halos (multi-GPU driver).

    Rank r owns output planes [x0, x1) and must run every window row (axis-0 start index) that
    intersects them, in order, which makes its planes bit-identical to the single-GPU result.  Cuts are
    placed at window starts so that window rows are balanced across ranks.  Returns per rank:
    ``dict(x0, x1, a0_begin, a0_end, vol_x0, vol_x1)``.
    """
    s0 = schedule.starts[0]
    n0, roi0, size0 = len(s0), schedule.roi[0], schedule.padded_size[0]
    world_size = max(1, int(world_size))

    def rows_of(x0, x1):
        return [j for j in range(n0) if s0[j] < x1 and s0[j] + roi0 > x0] if x1 > x0 else []

    # Cuts sit exactly on window starts: a rank owning [s0[a], s0[b]) executes rows a-1 .. b-1, i.e. one
    # redundant row per cut.  Greedy balance of the executed rows: total = n0 + (world - 1).
    cuts = [0]
    remaining_rows = n0 + world_size - 1
    j = 0
    for r in range(world_size - 1):
        target = -(-remaining_rows // (world_size - r))  # ceil
        # rank r executes rows [max(j-1,0), jn-1]  ->  jn - max(j-1, 0) rows
        jn = min(n0, max(j - 1, 0) + target)
        jn = max(jn, min(j + 1, n0))
        executed = jn - max(j - 1, 0)
        remaining_rows -= executed
        cuts.append(s0[jn] if jn < n0 else size0)
        j = jn
    cuts.append(size0)
    parts = []
    for r in range(world_size):
        x0, x1 = cuts[r], cuts[r + 1]
        rows = rows_of(x0, x1)
        if rows:
            a0b, a0e = rows[0], rows[-1] + 1
            vx0, vx1 = s0[a0b], s0[a0e - 1] + roi0
        else:
            a0b = a0e = 0
            vx0 = vx1 = x0
        parts.append(dict(x0=x0, x1=x1, a0_begin=a0b, a0_end=a0e, vol_x0=vx0, vol_x1=vx1))
    return parts


def window_partition(schedule: Schedule, world_size: int) -> List[dict]:
    """Partition the window LIST (MONAI order, axis 0 slowest) evenly over `world_size` ranks (multi-GPU
    driver, near-linear form: no window is computed twice).

    Rank r owns the contiguous windows ``[w_lo, w_hi)`` and the output planes ``[x0, x1)`` that start at the
    window row holding ``w_lo``.  Its planes are also covered by windows of earlier rows, which rank r-1 owns:
    they form the contiguous range ``[wb, w_lo)`` (``wb`` = first window of the first row that intersects the
    planes) and are RECEIVED from rank r-1; symmetrically rank r SENDS the tail ``[send_lo, w_hi)`` of its own
    range to rank r+1.  Blending rows ``[b_begin, b_end)`` in window order then reproduces the single-device
    sequence of fp32 additions for every voxel.  Per rank: ``dict(w_lo, w_hi, x0, x1, b_begin, b_end, wb,
    send_lo, vol_x0, vol_x1)``; ranks beyond the number of windows get empty ranges.
    """
    s0 = schedule.starts[0]
    n0, roi0, size0 = len(s0), schedule.roi[0], schedule.padded_size[0]
    per_row = len(schedule.starts[1]) * len(schedule.starts[2])
    total = n0 * per_row
    world_size = max(1, int(world_size))
    bounds = [(total * r) // world_size for r in range(world_size + 1)]
    parts = []
    for r in range(world_size):
        w_lo, w_hi = bounds[r], bounds[r + 1]
        if w_hi <= w_lo:
            parts.append(dict(w_lo=w_lo, w_hi=w_lo, x0=size0, x1=size0, b_begin=0, b_end=0, wb=w_lo, send_lo=w_lo,
                              vol_x0=0, vol_x1=0))
            continue
        row_lo, row_hi = w_lo // per_row, (w_hi - 1) // per_row
        x0 = 0 if r == 0 else s0[row_lo]
        parts.append(dict(w_lo=w_lo, w_hi=w_hi, x0=x0, x1=size0, row_lo=row_lo, row_hi=row_hi,
                          vol_x0=s0[row_lo], vol_x1=s0[row_hi] + roi0))
    # a rank's planes end where the next non-empty rank's begin
    nxt = size0
    for r in range(world_size - 1, -1, -1):
        p = parts[r]
        if p["w_hi"] <= p["w_lo"]:
            continue
        p["x1"] = nxt
        nxt = p["x0"]
    for r, p in enumerate(parts):
        if p["w_hi"] <= p["w_lo"]:
            continue
        if p["x1"] <= p["x0"]:  # all of its windows sit in a row shared with the previous rank: nothing to blend
            p.update(b_begin=0, b_end=0, wb=p["w_lo"])
        else:
            rows = [j for j in range(n0) if s0[j] < p["x1"] and s0[j] + roi0 > p["x0"]]
            p.update(b_begin=rows[0], b_end=rows[-1] + 1, wb=min(rows[0] * per_row, p["w_lo"]))
        p.pop("row_lo", None), p.pop("row_hi", None)
    for r, p in enumerate(parts):
        nxt_wb = None
        for q in parts[r + 1:]:
            if q["w_hi"] > q["w_lo"]:
                nxt_wb = q["wb"]
                break
        p["send_lo"] = p["w_hi"] if nxt_wb is None else min(max(nxt_wb, p["w_lo"]), p["w_hi"])
        if nxt_wb is not None and nxt_wb < p["w_lo"]:
            raise ValueError("window_partition: a rank would need windows of two earlier ranks "
                             f"(world_size {world_size} too large for {n0} window rows); use slab_partition")
    return parts


def plan_chunks(schedule: Schedule, bytes_per_window: int, fixed_bytes: int, budget_bytes: int) -> Optional[List[dict]]:
    """Single-GPU chunking of a volume whose deferred-blend buffer exceeds the memory budget: the smallest number of
    chunks of the window list (``window_partition``) such that the two alternating exchange buffers -- the largest and
    the second largest ``(w_hi - wb) * bytes_per_window`` -- plus ``fixed_bytes`` (network workspace) fit
    ``budget_bytes``.  ``None`` when no chunking fits (the caller falls back to the read-modify-write blend)."""
    for nchunks in range(2, len(schedule.starts[0]) + 1):
        try:
            cand = window_partition(schedule, nchunks)
        except ValueError:
            return None
        held = sorted((p["w_hi"] - p["wb"]) * int(bytes_per_window) for p in cand)
        if held[-1] + (held[-2] if len(held) > 1 else 0) + int(fixed_bytes) <= int(budget_bytes):
            return cand
    return None
