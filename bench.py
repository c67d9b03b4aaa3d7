#!/usr/bin/env python
"""Benchmark of segmantic's volumetric prediction hot path on B200 (driver contract: one JSON line).

Workload (BASELINE.json configs[1], the configuration `metric` is quoted on): MONAI UNet 3D
(channels 16-32-64-128-256, strides 2, 1 input channel, 10 tissues) over a synthetic 256^3 CT-like
volume, roi 96^3, overlap 0.5, Gaussian blend, bf16, argmax label map.  A "step" = one full
sliding-window prediction of the volume: 125 windows through the network, importance-weighted overlap
accumulation, normalise + argmax.  (No resampling stage in this config: spacing is already 1 mm.)

  value   Mvoxel/s with the (normalised) volume resident in HBM when the timed region starts.
  e2e     the same through the public API `segmantic_b200.seg.monai_unet.predict_volume` with HOST
          buffers: pinned H2D of the raw volume, z-score normalisation, foreground crop, prediction,
          D2H of the uint8 label map -- all inside the timed region.
  N > 1   weak scaling: the volume grows to (256*N) x 256 x 256; the window list is split evenly over the ranks
          (window ownership: no window is computed twice), a rank pushes the weighted logits of its last windows
          into the next rank's memory over NVLink (copy engines, peer mapping), every rank blends its own planes
          (bit-identical to the single-GPU result) and the uint8 label slabs are gathered over NCCL to rank 0;
          time = max over ranks.

`--impl reference` times the CPU restatement of the reference path (oracle/; the reference's own
MONAI/SimpleITK stack is not installable here) on the host cores for the same workload, on a bounded
sample of windows, extrapolated linearly and stated in `cpu_baseline.sample`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

VOL = (256, 256, 256)
ROI = (96, 96, 96)
CLASSES = 10
OVERLAP = 0.5
MODE = "gaussian"
METRIC = "3D UNet sliding-window predict throughput"
# ncu --set full captures of this round (profiles/r01c_ncu_top_kernels.md): dram__bytes_read.sum + dram__bytes_write.sum
NCU_HEAD_TRAFFIC = 3.543512e9 + 4.367288e9   # rs_conv_kernel<10,10>, one 125-window launch
NCU_BLEND_TRAFFIC = 4.545447e9 + 8.923e6     # gather_blend_vt_kernel<10>, the 256^3 blend
UNIT = "Mvoxel/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.

    NVML through `pynvml` in a polling thread (started well before the timed region; every nvidia-smi query spawned
    alongside the timed loop cost an occasional 10 ms stall of the launch stream), falling back to an `nvidia-smi
    -lms` child process.  Only the samples taken between begin() and end() are reported."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    PERIOD = 0.02

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.t0 = self.t1 = None
        self._stop = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception as e:  # noqa: BLE001
            log("pynvml unavailable, using nvidia-smi:", e)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            deadline = time.time() + 10.0
            while not self.rows and time.time() < deadline:  # wait for the first sample: start-up is over
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            log("clock sampler unavailable:", e)

    def _poll(self):
        n = self.nvml
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        masks = [(nm, getattr(n, a, None) or getattr(n, b, 0)) for nm, a, b in names]
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                bits = int(get_reasons(self.h)) if get_reasons else 0
                flags = ["Active" if (bits & m) else "Not Active" for _, m in masks]
                self.rows.append((time.time(), f"{sm}, {self.max_sm}, 0, " + ", ".join(flags)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.PERIOD)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        self._stop = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        if not self.proc and not self.nvml:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        sm, mx, reasons = [], [], set()
        rows = [r for (t, r) in list(self.rows) if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.03)]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), source="nvml" if self.nvml else "nvidia-smi")


def conv_flops(spec, roi, level_dims):
    """Algorithmic FLOPs (2*MAC, no channel padding) of one conv for one window."""
    k = spec.kernel ** 3
    return 2.0 * spec.cin * spec.cout * k * level_dims


def per_window_layer_flops(specs, roi):
    """{conv index: algorithmic FLOP per window}.  conv: 2*Cin*Cout*k^3*V_out; transposed: ...*V_in."""
    from segmantic_b200.seg.unet_spec import KIND_CONV_TRANSPOSE, KIND_IDENTITY
    n = (len(specs) - 3) // 5
    vox = [float(np.prod(roi))]
    for i in range(n):
        vox.append(vox[-1] / 8.0)
    out = {}
    idx = 0
    for i in range(n):
        for j in range(3):
            sp = specs[idx]
            if sp.kind != KIND_IDENTITY:
                out[idx] = 2.0 * sp.cin * sp.cout * sp.kernel ** 3 * vox[i + 1]
            idx += 1
    for j in range(3):
        sp = specs[idx]
        if sp.kind != KIND_IDENTITY:
            out[idx] = 2.0 * sp.cin * sp.cout * sp.kernel ** 3 * vox[n]
        idx += 1
    for i in range(n - 1, -1, -1):
        sp = specs[idx]
        out[idx] = 2.0 * sp.cin * sp.cout * 27 * vox[i + 1] if sp.kind == KIND_CONV_TRANSPOSE else 0.0
        idx += 1
        sp = specs[idx]
        out[idx] = 2.0 * sp.cin * sp.cout * 27 * vox[i]
        idx += 1
    return out


def make_block(seed=1):
    """The 256^3 normalised CT-like block (float32 [1, X, Y, Z]) + its raw version."""
    from segmantic_b200.synthetic import synthetic_volume
    raw = synthetic_volume(VOL, seed=seed)
    # float64 statistics: a float32 parallel reduction depends on the number of host threads (torchrun sets
    # OMP_NUM_THREADS=1, plain python uses every core), which moved the normalised volume by an ulp between the N = 1
    # and the N > 1 runs and with it a few thousand argmax near-ties (label checksums of configs[3], round 2)
    r64 = raw.double()
    m, sd = float(f"{float(r64.mean()):.12g}"), float(f"{float(r64.std(unbiased=False)):.12g}")  # thread-count independent
    norm = ((r64 - m) / sd).float()
    return raw, norm


# ---------------------------------------------------------------------------------------- CPU arm
def cpu_sample(n_windows: int, threads: int, seed=1):
    """Times the oracle on `n_windows` of the workload's 125 windows (+ their blend) and extrapolates."""
    from oracle import sliding_window as osw
    from oracle.unet import UNet, load_checkpoint_into
    from segmantic_b200.synthetic import synthetic_state_dict
    torch.set_num_threads(threads)
    sd = synthetic_state_dict(3, 1, CLASSES, seed=0)
    net = UNet(3, 1, CLASSES)
    load_checkpoint_into(net, sd)
    net.eval()
    starts = osw.window_starts(VOL, ROI, OVERLAP)
    g = torch.Generator().manual_seed(seed)
    if n_windows >= len(starts):
        # the WHOLE workload: the 256^3 block, all 125 windows, blend, argmax -- nothing extrapolated
        vol = make_block(seed)[1][None]
    else:
        # a strip of the volume holding n_windows consecutive (50 % overlapping) windows
        vol = torch.randn((1, 1, ROI[0], ROI[1], ROI[2] + 48 * (n_windows - 1)), generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        out = osw.sliding_window_inference(vol, ROI, 4, net, overlap=OVERLAP, mode=MODE)
        out.argmax(1)
    dt = time.perf_counter() - t0
    nwin_done = len(osw.window_starts(vol.shape[2:], ROI, OVERLAP))
    total_windows = len(starts)
    est = dt * total_windows / nwin_done
    return dict(seconds_sample=dt, windows_sample=nwin_done, windows_total=total_windows,
                seconds_extrapolated=est, mvox_s=float(np.prod(VOL)) / est / 1e6)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    vals = []
    n_win = args.cpu_windows
    n_steps = max(1, args.warmup) + args.steps
    if n_win >= 125:
        # the whole workload per step when the run then still ends within ~4 minutes; otherwise the largest strip of
        # windows that does (a probe of 8 windows measures this box's per-window time)
        probe = cpu_sample(8, threads)
        per_window = probe["seconds_sample"] / probe["windows_sample"]
        fit = int(240.0 / n_steps / per_window)
        n_win = 125 if fit >= 125 else max(8, fit)
        log(f"[reference] {per_window * 1e3:.0f} ms per window on {threads} threads, {n_steps} steps -> {n_win} windows per step")
    for i in range(n_steps):
        r = cpu_sample(n_win, threads)
        if i >= max(1, args.warmup):
            vals.append(r)
        log(f"[reference] step {i}: {r['seconds_sample']:.2f} s for {r['windows_sample']} windows "
            f"-> {r['mvox_s']:.4f} Mvoxel/s" + ("" if r['windows_sample'] >= r['windows_total'] else " extrapolated"))
    v = float(np.mean([r["mvox_s"] for r in vals]))
    ms = float(np.mean([r["seconds_extrapolated"] for r in vals])) * 1e3
    full = vals[0]['windows_sample'] >= vals[0]['windows_total']
    sample = (f"{vals[0]['windows_sample']} of {vals[0]['windows_total']} ROI windows (96^3, 10 classes) through the "
              f"torch-CPU oracle UNet + Gaussian blend + argmax" +
              (": the whole 256^3 volume, nothing extrapolated" if full else ", extrapolated linearly to the 256^3 volume"))
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload="configs[1]: UNet3D 10 tissues, 256^3, roi 96^3, overlap 0.5, gaussian",
                            impl="oracle port of the reference path (torch CPU fp32); MONAI/SimpleITK not installable"),
                cpu_baseline=dict(value=v, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch.distributed as dist
    from segmantic_b200.seg import engine
    from segmantic_b200.seg.monai_unet import Net, predict_volume, predict_volumes
    from segmantic_b200.seg.multi_gpu import gather_label_slabs
    from segmantic_b200.seg.sliding_window import make_schedule, slab_partition, window_partition
    from segmantic_b200.seg.unet_spec import unet_conv_specs
    from segmantic_b200.synthetic import synthetic_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback in segmantic_b200)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    sd = synthetic_state_dict(3, 1, CLASSES, seed=0)
    net = engine.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=CLASSES, device=dev,
                          precision=args.precision)
    raw, norm = make_block(seed=1)
    gshape = (VOL[0] * world, VOL[1], VOL[2])
    sched = make_schedule(gshape, ROI, OVERLAP, MODE)
    # N > 1: window OWNERSHIP (default): the window list is split evenly, no window is computed twice, and the
    # importance-weighted logits of a rank's last windows are pushed into the next rank's memory over NVLink by the copy
    # engines (seg/p2p.py: CUDA IPC mapping + device-side counters, no collective library on the data path).
    # SGM_MGPU=owned-nccl: the same partition with NCCL point-to-point; SGM_MGPU=slab: output slabs with ROI halos (no
    # exchange at all, one redundant window row per cut: efficiency bound 5/6 .. 5/7).
    mgpu = os.environ.get("SGM_MGPU", "owned")
    owned = world > 1 and mgpu.startswith("owned")
    exchange = "nccl" if mgpu == "owned-nccl" else "p2p"
    seam_groups = None
    if owned and exchange == "nccl":
        # two extra communicators: a rank's receive (from rank-1) and send (to rank+1) must not share one; their
        # pairwise NCCL communicators are created here, outside the timed steps
        seam_groups = (dist.new_group(), dist.new_group())
        token = torch.zeros(1, device=dev)
        for s_ in range(world - 1):
            if rank == s_:
                dist.send(token, dst=s_ + 1, group=seam_groups[s_ % 2])
            elif rank == s_ + 1:
                dist.recv(token, src=s_, group=seam_groups[s_ % 2])
        dist.barrier()
    all_parts = window_partition(sched, world) if owned else slab_partition(sched, world)
    if world > 1:
        part = all_parts[rank]
        # the global volume is the 256^3 block tiled along axis 0; a rank uploads only its halo'ed slab
        idx = torch.arange(part["vol_x0"], part["vol_x1"]) % VOL[0]
        vol_dev = norm[:, idx].contiguous().to(dev)
    else:
        part = None
        vol_dev = norm.to(dev)
    if owned:
        n_win = part["w_hi"] - part["w_lo"]
    else:
        n_win = (len(sched.starts[0]) if part is None else part["a0_end"] - part["a0_begin"]) * \
            len(sched.starts[1]) * len(sched.starts[2])
    log(f"[rank {rank}] volume {gshape}, {n_win} windows on this rank, precision {args.precision}, "
        f"sw_batch {args.sw_batch}")

    def step():
        if part is None:
            res = engine.sliding_window_inference(vol_dev[None], ROI, args.sw_batch, net, overlap=OVERLAP, mode=MODE,
                                                  return_labels=True, return_logits=False)
            return res["labels"]
        if owned:  # every window once; the tail that covers the next rank's planes travels over NVLink
            res = engine.sliding_window_inference_owned(vol_dev, gshape, part, ROI, args.sw_batch, net, overlap=OVERLAP,
                                                        mode=MODE, rank=rank, world_size=world, group=seam_groups,
                                                        exchange=exchange)
        else:
            res = engine.sliding_window_inference_slab(vol_dev, gshape, part, ROI, args.sw_batch, net, overlap=OVERLAP,
                                                       mode=MODE)
        # gather the uint8 label slabs on rank 0 (NCCL over NVLink): the only collective of the path
        return gather_label_slabs(res["labels"], all_parts, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    # settle: extra untimed steps until two consecutive ones agree within 5 % (allocator growth, clock ramp, the
    # sampler's NVML start-up); at most 10
    # (N > 1: a FIXED number of settle steps -- every step is a collective exchange between the ranks, so all ranks must
    # run the same number of them; a per-rank, timing-dependent exit left the ranks with different step counts.)
    prev = None
    for it in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize(dev)
        cur = e0.elapsed_time(e1)
        if world > 1:
            if it >= 3:
                break
        elif prev is not None and abs(cur - prev) <= 0.05 * prev:
            break
        prev = cur
    net.check()
    barrier()
    # timed region: K steps with NO per-launch instrumentation (an event pair around every launch costs ~30 us of
    # front-end serialisation per launch, ~20-40 % of a step here)
    net.set_profiling(False)
    # A K-step region is repeated (at most 3 regions) when one of its steps took > 1.3x the median step -- on a shared box
    # a host stall of tens of ms now and then starves the launch stream; the fastest region is reported, the number of
    # regions and every region's time are in the JSON line ("regions_ms").
    regions = []
    sampler.begin()
    for _region in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]  # one event per step (diagnostic spread)
        barrier()
        ev0.record()
        for i in range(args.steps):
            out = step()
            marks[i].record()
        ev1.record()
        barrier()
        ms_r = ev0.elapsed_time(ev1) / args.steps
        per_r = [(ev0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(args.steps)]
        t = torch.tensor([ms_r, max(per_r) / float(np.median(per_r))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)   # all ranks take the same decision
        regions.append((float(t[0].item()), ms_r, per_r))
        if float(t[1].item()) <= 1.3:
            break
    sampler.end()
    _, ms, per_step = min(regions, key=lambda r: r[0])
    launches_step = getattr(net, "last_launch_count_owned", 0) if owned else net.last_launch_count  # network + blend
    # per-kernel durations for the rooflines: the same K steps again, with a CUDA-event pair recorded by the
    # library around every launch on the launching stream
    net.set_profiling(not args.no_profile)
    net.get_profile()
    if not args.no_profile:
        for _ in range(args.steps):
            out = step()
    barrier()
    prof = net.get_profile()
    net.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    net.check()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_vox = float(np.prod(gshape))
    value = total_vox / (ms * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers (rank-local slab semantics for N > 1 are
    # the same compute; the e2e number is reported for the single-volume API call on every rank)
    e2e = None
    if True:
        pnet = Net(num_classes=CLASSES, num_channels=1, spatial_dims=3)
        pnet.load_state_dict(sd)
        pnet.to(dev)
        host = raw.clone().pin_memory()
        # crop_foreground=False: CropForegroundd would shrink this synthetic volume to its body ellipsoid
        # (data dependent); it is disabled so that e2e runs the same 125 windows as `value`.
        kw = dict(overlap=OVERLAP, mode=MODE, sw_batch_size=args.sw_batch, precision=args.precision,
                  crop_foreground=False)
        for _ in range(3):
            predict_volume(pnet, host, None, (), **kw)
        pipelined = True
        try:
            for _ in predict_volumes(pnet, [host] * 6, None, (), **kw):  # warm the pipelined path (streams, buffer rings)
                pass
        except Exception as e:  # noqa: BLE001  (never lose the bench line to the e2e section: fall back to plain calls)
            log(f"[rank {rank}] pipelined e2e unavailable ({e!r}); timing sequential predict_volume calls instead")
            pipelined = False
        barrier()

        def e2e_results(k):
            if pipelined:
                yield from predict_volumes(pnet, [host] * k, None, (), **kw)
            else:
                for _ in range(k):
                    yield predict_volume(pnet, host, None, (), **kw)

        e2e_steps = max(1, args.steps)
        e2e_regions = []
        for _region in range(3):  # same policy as the device-timed region: repeat a region that saw a stalled call
            barrier()
            t0 = time.perf_counter()
            calls = []
            # predict_volumes = the loop of predict() over its images: every volume is uploaded from the pinned host
            # buffer and its label map downloaded inside this region; the copies of neighbouring volumes overlap the
            # prediction of the current one (copy engines on their own streams)
            tc0 = time.perf_counter()
            for lab in e2e_results(e2e_steps):  # HOST uint8 label maps, in order
                tc1 = time.perf_counter()
                calls.append((tc1 - tc0) * 1e3)
                tc0 = tc1
            torch.cuda.synchronize(dev)
            e2e_regions.append(((time.perf_counter() - t0) * 1e3 / e2e_steps, calls))
            tt = torch.tensor([max(calls) / float(np.median(calls))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if float(tt.item()) <= 1.3:
                break
        e2e_ms, e2e_calls = min(e2e_regions, key=lambda r: r[0])
        log(f"[rank {rank}] e2e ms per call, region by region: {[[round(c, 2) for c in r[1]] for r in e2e_regions]}; "
            f"launches per call {pnet.engine(args.precision).last_launch_count}")
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = dict(value=float(np.prod(VOL)) * world / (e2e_ms * 1e-3) / 1e6, unit=UNIT,
                   h2d_bytes_per_step=int(host.numel() * 4), d2h_bytes_per_step=int(lab.numel()),
                   ms_per_step=e2e_ms, ms_per_call=dict(min=min(e2e_calls), median=float(np.median(e2e_calls)),
                                                        max=max(e2e_calls)),
                   regions_ms=[round(r[0], 3) for r in e2e_regions],
                   pipelined=pipelined,
                   note="predict_volumes(K pinned host fp32 volumes) -> K host uint8 label maps (the loop of predict() over its "
                        "images): every volume crosses PCIe once each way inside the timed region, the copies of "
                        "neighbouring volumes overlap the prediction of the current one; ms_per_call = interval between "
                        "yielded label maps; N>1: one 256^3 volume per rank and step")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (the head: conv C->C + residual + weighted blend accumulate)
    specs = unet_conv_specs(1, CLASSES)
    flops = per_window_layer_flops(specs, ROI)
    layers = []
    tot_ms = sum(p[1] for p in prof) or 1.0
    blend_prof = prof[-1]
    prof = prof[:-1]
    for i, (role, pms, cnt) in enumerate(prof):
        if cnt == 0:
            continue
        fl = flops.get(i, 0.0)
        wins = n_win * args.steps
        tf = fl * wins / (pms * 1e-3) / 1e12 if pms > 0 else 0.0
        layers.append(dict(conv=role, ms_per_step=pms / args.steps, launches_per_step=cnt // args.steps,
                           share=pms / tot_ms, gflop_per_window=fl / 1e9, tflops=tf,
                           frac_bf16_peak=tf / pk["bf16_sustained"]))
    head = prof[-1]
    roi_vox = float(np.prod(ROI))
    esz = 2 if args.precision == "bf16" else 4
    # Dominant kernel = the head conv (C->C 3x3x3 + identity residual, importance-weighted logits out).
    # Algorithmic work per window: 2*C*C*27*roi^3 FLOP; bytes: read u (C ch, bf16) + write C fp32 logits.
    roofline = None
    if head[2] > 0 and head[1] > 0:
        launches = head[2]                      # head launches in the profiled steps (one per window batch)
        t_launch = head[1] * 1e-3 / launches    # average launch duration, CUDA events on the launching stream
        wins_per_launch = n_win * args.steps / launches
        per_window_s = t_launch / wins_per_launch
        head_flop = flops[len(specs) - 1]
        head_bytes = (esz * CLASSES + 4 * CLASSES) * roi_vox   # read C bf16 channels, write C fp32 weighted logits
        gbs = head_bytes * wins_per_launch / t_launch / 1e9
        tfs = head_flop * wins_per_launch / t_launch / 1e12
        # ncu --set full (profiles/r01c_ncu_top_kernels.md): dram__bytes_read + write of one 125-window launch
        row_sweep = not os.environ.get("SGM_NO_RS")
        traffic = NCU_HEAD_TRAFFIC if (CLASSES == 10 and args.precision == "bf16" and wins_per_launch == 125
                                       and row_sweep) else None
        roofline = dict(kernel=("rs_conv_kernel<10,10>[head: conv CxC k3, d0 and d1 taps folded into MMA N through "
                                "overlapping accumulator columns + identity residual + importance-weighted fp32 logits]"
                                if row_sweep else
                                "ps_conv_kernel<10,1,PLANAR>[head: conv CxC k3 (d0 taps folded into MMA N) + identity "
                                "residual + importance-weighted fp32 logits]"),
                        bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                        traffic=traffic, traffic_note="ncu dram bytes of a 125-window launch (algorithmic: 6.64 GB + "
                                                      "16-channel padding of the bf16 input = 7.96 GB)",
                        peak_source=pk["source"], us_per_window=per_window_s * 1e6, us_per_launch=t_launch * 1e6,
                        windows_per_launch=wins_per_launch, algorithmic_bytes_per_window=head_bytes,
                        algorithmic_bytes_per_launch=head_bytes * wins_per_launch,
                        algorithmic_flop_per_window=head_flop, tensor_tflops=tfs,
                        tensor_frac=tfs / pk["bf16_sustained"], tensor_peak=pk["bf16_sustained"],
                        note="the head moves 60 B per voxel for 5.4 kFLOP: its nearer roof is HBM (fraction above), "
                             "not the tensor pipe (tensor_frac); SS-mode MMAs with N <= 48 are bound by the shared-memory "
                             "read of their A tile (tests/ubench_mma.cu), see DESIGN.md 4.1")
    conv_ms = sum(l["ms_per_step"] for l in layers)
    conv_tf = sum(flops.get(i, 0.0) for i in range(len(specs))) * n_win / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0

    blend_roof = None
    if world == 1:
        # HBM-bound stage: the deferred gather blend (+count, normalise, argmax) reads every window's
        # weighted logits once and writes one label byte per voxel.  Timed alone with CUDA events on a
        # (measured by the library's CUDA-event pair around the launch, inside the timed region).
        blend_bytes = 4.0 * CLASSES * n_win * roi_vox + float(np.prod(VOL))
        if blend_prof[2] > 0 and blend_prof[1] > 0:
            t = blend_prof[1] * 1e-3 / blend_prof[2]
            blend_roof = dict(kernel="gather_blend_vt_kernel<10>[sum covering windows in MONAI order + count + "
                                     "division-free exact argmax]",
                              bound="hbm", achieved=blend_bytes / t / 1e9, peak=pk["hbm"], unit="GB/s",
                              frac=blend_bytes / t / 1e9 / pk["hbm"],
                              traffic=NCU_BLEND_TRAFFIC if CLASSES == 10 else None, peak_source=pk["source"],
                              ms_per_launch=t * 1e3, algorithmic_bytes_per_launch=blend_bytes,
                              formula="4*C*n_windows*roi^3 (weighted logits read once) + 1*V (labels written)")

    resample_roof = None
    if world == 1 and not args.no_profile:
        try:
            resample_roof = resample_rooflines(dev, pk)
        except Exception as e:  # noqa: BLE001  (diagnostic section: never lose the headline line)
            log("resample rooflines failed:", repr(e))
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        r = cpu_sample(args.cpu_windows, threads)
        cpu = dict(value=r["mvox_s"], unit=UNIT, cores=threads, kind="port",
                   sample=f"{r['windows_sample']} of {r['windows_total']} windows ({r['seconds_sample']:.1f} s) through "
                          f"the torch-CPU oracle UNet + blend + argmax" +
                          (": the whole 256^3 volume, nothing extrapolated" if r['windows_sample'] >= r['windows_total']
                           else ", extrapolated linearly"))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=args.precision, data="synthetic",
                config=dict(workload="configs[1]: MONAI UNet3D (16-32-64-128-256, strides 2, 1 ch, 10 tissues), "
                                     f"synthetic {gshape[0]}x{gshape[1]}x{gshape[2]} volume, roi 96^3, overlap 0.5, "
                                     "gaussian blend, argmax labels",
                            windows=int(n_win if world == 1 else sched.n_windows),
                            sw_batch=max(args.sw_batch, int(os.environ.get("SGM_SW_BATCH", engine.DEVICE_SW_BATCH))),
                            parallelism=(f"owned-windows{world}+{'nvlink-peer-push' if exchange == 'p2p' else 'nccl-p2p'}" if owned
                                         else f"slab{world}") if world > 1 else "single",
                            l2="no flush: per-step working set (67 MB volume + 671 MB accumulator + activations) "
                               "exceeds the 126 MB L2"),
                step_ms=dict(min=min(per_step), median=float(np.median(per_step)), max=max(per_step)),
                regions_ms=[round(r[0], 3) for r in regions],
                e2e=e2e, gpu_launches=int(launches_step * args.steps), clocks=clocks, roofline=roofline,
                roofline_blend=blend_roof, roofline_resample=resample_roof,
                cpu_baseline=cpu,
                conv_stack=dict(ms_per_step=conv_ms, tflops=conv_tf, frac_bf16_peak=conv_tf / pk["bf16_sustained"],
                                peak_tflops=pk["bf16_sustained"], layers=layers))
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------- other BASELINE configs (not the headline)
def _tiled_block(norm: torch.Tensor, x0: int, x1: int, shape):
    """Planes [x0, x1) of the 256^3 block tiled periodically to `shape` ([1, x1 - x0, Y, Z] float32, host)."""
    i0 = torch.arange(x0, x1) % norm.shape[1]
    i1 = torch.arange(shape[1]) % norm.shape[2]
    i2 = torch.arange(shape[2]) % norm.shape[3]
    return norm[:, i0][:, :, i1][:, :, :, i2].contiguous()


def run_config3(args):
    """BASELINE configs[3]: whole-body 512 x 512 x 1024 volume, 20 tissues, STRONG scaling over N GPUs.  The 1024-voxel
    axis is the slowest memory axis (axis 0 of the [X, Y, Z] tensor), so z-slabs are contiguous; N = 1 runs the window
    list in chunks (the deferred-blend buffer of all 2100 windows is 148 GB), N > 1 splits it over the ranks (window
    ownership + NVLink peer push) and gathers the uint8 label slabs on rank 0."""
    import torch.distributed as dist
    from segmantic_b200.seg import engine
    from segmantic_b200.seg.multi_gpu import gather_label_slabs
    from segmantic_b200.seg.sliding_window import make_schedule, window_partition
    from segmantic_b200.synthetic import synthetic_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    classes, gshape = 20, (1024, 512, 512)
    sd = synthetic_state_dict(3, 1, classes, seed=0)
    net = engine.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=classes, device=dev, precision=args.precision)
    _, norm = make_block(seed=4)
    sched = make_schedule(gshape, ROI, OVERLAP, MODE)
    parts = window_partition(sched, world) if world > 1 else None
    part = parts[rank] if parts else None
    vol_dev = _tiled_block(norm, part["vol_x0"] if part else 0, part["vol_x1"] if part else gshape[0], gshape).to(dev)
    n_win = (part["w_hi"] - part["w_lo"]) if part else sched.n_windows
    log(f"[rank {rank}] configs[3]: volume {gshape}, {classes} tissues, {n_win} of {sched.n_windows} windows on this rank")

    def step():
        if part is None:
            return engine.sliding_window_inference(vol_dev[None], ROI, args.sw_batch, net, overlap=OVERLAP, mode=MODE,
                                                   return_labels=True, return_logits=False)["labels"]
        res = engine.sliding_window_inference_owned(vol_dev, gshape, part, ROI, args.sw_batch, net, overlap=OVERLAP,
                                                    mode=MODE, rank=rank, world_size=world)
        return gather_label_slabs(res["labels"], parts, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    for _ in range(warm):
        out = step()
    net.check()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    barrier()
    net.check()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    checksum = int(out.to(torch.int64).sum().item()) if rank == 0 else 0
    layers = None
    if not args.no_profile:   # one more step with a CUDA-event pair around every launch: where the time goes
        net.set_profiling(True)
        net.get_profile()
        step()
        barrier()
        prof = net.get_profile()
        net.set_profiling(False)
        layers = [dict(conv=r, ms_per_step=round(m, 3), launches=c) for r, m, c in prof if c]
        log(f"[rank {rank}] per-layer ms: " + ", ".join(f"{l['conv']} {l['ms_per_step']}" for l in layers))
    if rank == 0:
        launches = getattr(net, "last_launch_count_owned", 0) if part else getattr(net, "last_launch_count_chunked", 0)
        line = dict(metric=METRIC, value=float(np.prod(gshape)) / (ms * 1e-3) / 1e6, unit=UNIT, n_gpus=world, steps=steps,
                    warmup=warm, ms_per_step=ms, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype=args.precision, data="synthetic",
                    config=dict(workload="configs[3]: MONAI UNet3D, 20 tissues, synthetic whole-body 512x512x1024 volume "
                                         "(the 1024 axis slowest), roi 96^3, overlap 0.5, gaussian blend, argmax labels",
                                windows=int(sched.n_windows),
                                parallelism=f"owned-windows{world}+nvlink-peer-push" if world > 1 else
                                "single GPU, window list in chunks (deferred-blend buffer of 2100 windows = 148 GB)",
                                l2="no flush: the working set is far larger than the 126 MB L2"),
                    e2e=None, gpu_launches=int(launches * steps), label_checksum=checksum, layers=layers)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_config2(args):
    """BASELINE configs[2]: the four-stage pipeline on an anisotropic 512 x 512 x 120 image at 0.5 x 0.5 x 3 mm --
    trilinear Spacing to 1 mm isotropic (256 x 256 x 358), sliding-window prediction (10 tissues), argmax, nearest-
    neighbour resample of the label map back onto the input grid -- through `predict_volume(..., spacing=(1, 1, 1),
    invert="labels")`.  `value`: image resident on the device, labels left on the device; `e2e`: pinned host image in,
    host label map out.  Mvoxel = voxels of the NETWORK-grid volume."""
    from segmantic_b200.seg.monai_unet import Net, predict_volume, predict_volumes
    from segmantic_b200.synthetic import synthetic_state_dict, synthetic_volume

    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    sd = synthetic_state_dict(3, 1, CLASSES, seed=0)
    pnet = Net(num_classes=CLASSES, num_channels=1, spatial_dims=3)
    pnet.load_state_dict(sd)
    pnet.to(dev)
    src_shape, net_shape = (512, 512, 120), (256, 256, 358)
    raw = synthetic_volume(src_shape, seed=2)
    affine = np.diag([-0.5, -0.5, 3.0, 1.0])   # identity-direction ITK image (LPS) seen in RAS, 0.5 x 0.5 x 3 mm
    host = raw.clone().pin_memory()
    on_dev = raw.to(dev)
    kw = dict(spacing=(1.0, 1.0, 1.0), overlap=OVERLAP, mode=MODE, sw_batch_size=args.sw_batch, precision=args.precision,
              invert="labels", crop_foreground=False)
    steps, warm = max(1, args.steps), max(3, args.warmup)
    for _ in range(warm):
        lab = predict_volume(pnet, on_dev, affine, return_device=True, **kw)
    assert tuple(lab.shape) == src_shape
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lab = predict_volume(pnet, on_dev, affine, return_device=True, **kw)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    for _ in range(2):
        predict_volume(pnet, host, affine, **kw)
    for _ in predict_volumes(pnet, [host] * 4, [affine] * 4, **kw):   # warm the pipelined path (streams, pinned blocks)
        pass
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for hl in predict_volumes(pnet, [host] * steps, [affine] * steps, **kw):   # pipelined over the images, as predict()
        pass
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    nvox = float(np.prod(net_shape))
    eng = pnet.engine(args.precision)
    line = dict(metric=METRIC, value=nvox / (ms * 1e-3) / 1e6, unit=UNIT, n_gpus=1, steps=steps, warmup=warm, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.precision, data="synthetic",
                config=dict(workload="configs[2]: 512x512x120 @ 0.5x0.5x3 mm -> Spacing 1 mm (256x256x358) -> UNet3D 10 "
                                     "tissues, roi 96^3, overlap 0.5, gaussian -> argmax -> nearest resample back to 512x512x120",
                            stages="orientation, z-score, trilinear Spacing, sliding window + blend + argmax, ITK nearest back",
                            l2="no flush: volumes and activations exceed the 126 MB L2"),
                e2e=dict(value=nvox / (e2e_ms * 1e-3) / 1e6, unit=UNIT, ms_per_step=e2e_ms,
                         h2d_bytes_per_step=int(host.numel() * 4), d2h_bytes_per_step=int(hl.numel())),
                gpu_launches=int(eng.last_launch_count * steps), label_checksum=int(hl.to(torch.int64).sum().item()))
    print(json.dumps(line), flush=True)
    return 0


def run_config4(args):
    """BASELINE configs[4]: multi-channel (T1 + T2) 2-D UNet, slice-wise over a 512 x 512 x 400 stack
    (`predict_stack_2d`: every Z slice is an independent 2-D image, roi 96 x 96, all slices' windows share the network
    launches).  `value`: stack resident on the device; `e2e`: pinned host stack in, host label map out."""
    from segmantic_b200.seg.monai_unet import Net, predict_stack_2d
    from segmantic_b200.synthetic import synthetic_state_dict, synthetic_volume

    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    sd = synthetic_state_dict(2, 2, CLASSES, seed=0)
    pnet = Net(num_classes=CLASSES, num_channels=2, spatial_dims=2, spatial_size=[96, 96])
    pnet.load_state_dict(sd)
    pnet.to(dev)
    shape = (512, 512, 400)
    block = synthetic_volume((256, 256, 200), seed=5, channels=2)
    raw = block.repeat(1, 2, 2, 2).contiguous()          # [2, 512, 512, 400]
    host, on_dev = raw.pin_memory(), raw.to(dev)
    kw = dict(overlap=0.25, mode=MODE, sw_batch_size=args.sw_batch, precision=args.precision)
    steps, warm = max(1, min(args.steps, 5)), max(2, min(args.warmup, 3))
    for _ in range(warm):
        lab = predict_stack_2d(pnet, on_dev, return_device=True, **kw)
    assert tuple(lab.shape) == shape
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lab = predict_stack_2d(pnet, on_dev, return_device=True, **kw)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    predict_stack_2d(pnet, host, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        hl = predict_stack_2d(pnet, host, **kw)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    nvox = float(np.prod(shape))
    line = dict(metric=METRIC, value=nvox / (ms * 1e-3) / 1e6, unit=UNIT, n_gpus=1, steps=steps, warmup=warm, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.precision, data="synthetic",
                config=dict(workload="configs[4]: 2-D MONAI UNet (2 input channels T1 + T2, 10 tissues), slice-wise over a "
                                     "synthetic 512x512x400 stack, roi 96x96, overlap 0.25, gaussian blend, argmax labels",
                            l2="no flush: the stack (839 MB) and the activations exceed the 126 MB L2"),
                e2e=dict(value=nvox / (e2e_ms * 1e-3) / 1e6, unit=UNIT, ms_per_step=e2e_ms,
                         h2d_bytes_per_step=int(host.numel() * 4), d2h_bytes_per_step=int(hl.numel())),
                gpu_launches=int(pnet.engine(args.precision).last_launch_count * steps),
                label_checksum=int(hl.to(torch.int64).sum().item()))
    print(json.dumps(line), flush=True)
    return 0


def resample_rooflines(dev, pk, iters=12):
    """HBM rooflines of the resample kernels on the shapes of BASELINE configs[2] (512x512x120 @ 0.5x0.5x3 mm <-> 1 mm
    isotropic 256x256x358): Spacing forward (trilinear), the fused inverse Spacing + argmax of 10-class logits, and the
    ITK nearest-neighbour back-resample of a uint8 label map.  Outside the timed region of the headline metric; CUDA
    events around every launch; inputs rotate over distinct buffers so that no launch finds its input in the 126 MB L2.
    Algorithmic bytes: every input voxel read once + every output voxel written once."""
    from segmantic_b200.seg import transforms as T

    src_shape, dst_shape = (512, 512, 120), (256, 256, 358)
    aff = np.diag([0.5, 0.5, 3.0, 1.0])
    g = torch.Generator(device="cpu").manual_seed(3)
    out = []

    def timed(fn, nbuf):
        fn(0)
        torch.cuda.synchronize(dev)
        ts = []
        for i in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(i % nbuf)
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    def entry(kernel, ms, nbytes, formula):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return dict(kernel=kernel, bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                    ms_per_launch=ms, algorithmic_bytes_per_launch=nbytes, formula=formula, peak_source=pk["source"])

    # 1. Spacingd forward: [1, 512, 512, 120] fp32 -> [1, 256, 256, 358]
    nbuf = 6
    imgs = [torch.randn((1,) + src_shape, generator=g).to(dev) for _ in range(nbuf)]
    new_aff = T.zoom_affine(aff, (1.0, 1.0, 1.0))
    shp, offset = T.compute_shape_offset(src_shape, aff, new_aff)
    new_aff[:3, -1] = offset
    assert tuple(shp) == dst_shape, shp
    xf = np.linalg.solve(aff, new_aff)
    ms = timed(lambda i: T.resample_index_affine(imgs[i], xf, dst_shape), nbuf)
    out.append(entry("trilinear_kernel<false>[Spacingd 512x512x120 @0.5x0.5x3 -> 256x256x358 @1 mm, 1 channel; float64 arithmetic]", ms,
                     4.0 * np.prod(src_shape) + 4.0 * np.prod(dst_shape), "4*V_in + 4*V_out"))
    del imgs
    # 2. inverse Spacing of 10-class logits fused with argmax: [10, 256, 256, 358] -> uint8 [512, 512, 120]
    logits = [torch.randn((CLASSES,) + dst_shape, generator=g).to(dev) for _ in range(2)]
    xinv = np.linalg.solve(new_aff, aff)
    ms = timed(lambda i: T.resample_index_affine_argmax(logits[i], xinv, src_shape), 2)
    out.append(entry("trilinear_brick_kernel<true>[inverse Spacing of 10-class logits + argmax -> 512x512x120 labels; float64 arithmetic]", ms,
                     4.0 * CLASSES * np.prod(dst_shape) + 1.0 * np.prod(src_shape), "4*C*V_in + 1*V_out"))
    del logits
    # 3. ITK nearest-neighbour resample_to_ref of a uint8 label map: [256, 256, 358] @1 mm -> [512, 512, 120]
    # (the C ABI directly: processing.resample_to_ref adds the [x,y,z] <-> C-order permute copies around the kernel)
    import ctypes as C

    from segmantic_b200 import _lib
    lib = _lib.load()
    labs = [torch.randint(0, CLASSES, tuple(reversed(dst_shape)), generator=g, dtype=torch.uint8).to(dev) for _ in range(nbuf)]
    lout = torch.empty(tuple(reversed(src_shape)), dtype=torch.uint8, device=dev)

    def dbl(a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        return (C.c_double * a.size)(*a.tolist())

    i2p, p2i, zero = dbl(np.diag([0.5, 0.5, 3.0])), dbl(np.eye(3)), dbl(np.zeros(3))

    def itk(i):
        with torch.cuda.device(dev):
            _lib.check(lib.sgm_resample_itk(labs[i].data_ptr(), 0, _lib.i3(dst_shape), lout.data_ptr(), _lib.i3(src_shape),
                                            i2p, zero, p2i, zero, 1, 0.0, int(torch.cuda.current_stream(dev).cuda_stream)),
                       "sgm_resample_itk")

    ms = timed(itk, nbuf)
    out.append(entry("itk_resample_vec_kernel<uint8,4>[nearest resample_to_ref of labels 256x256x358 -> 512x512x120; float64 index arithmetic]", ms,
                     1.0 * np.prod(dst_shape) + 1.0 * np.prod(src_shape), "1*V_in + 1*V_out"))
    return out

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--sw-batch", type=int, default=32,
                    help="MONAI's sw_batch_size argument; the device batches max(this, engine.DEVICE_SW_BATCH) windows")
    ap.add_argument("--cpu-windows", type=int, default=125,
                    help="windows in the CPU-baseline sample: 125 = the whole 256^3 workload (about 10 s on 16 cores, nothing "
                         "extrapolated); fewer = a strip of that many windows, extrapolated linearly")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs index: 1 = the headline (256^3, 10 tissues), 2 = the four-stage "
                         "anisotropic pipeline, 3 = whole-body 512x512x1024 with 20 tissues (strong scaling), 4 = 2-D two-channel slice-wise stack")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 3:
        return run_config3(args)
    if args.config == 2:
        return run_config2(args)
    if args.config == 4:
        return run_config4(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
