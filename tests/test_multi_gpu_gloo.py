"""World-size-2 CPU (gloo) test of the multi-GPU driver's host logic: slab partition + label gather.

The per-rank slab compute needs a GPU, so here each rank fills its slab with the CPU oracle's labels;
what is under test is that the partition covers the volume, that every rank runs exactly the window
rows its planes need (so its planes equal the single-process result) and that gather_label_slabs
reassembles the volume over a real process group."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sliding_window as osw
from tests.helpers import make_oracle_net, normalized_volume

ROI, OVERLAP, MODE = (16, 16, 16), 0.5, "gaussian"
SHAPE = (56, 24, 20)


def _slab_oracle(vol, net, sched, part):
    """Oracle restricted to a rank's window rows, accumulating only its planes (what the GPU rank does)."""
    s0 = sched.starts[0]
    C = net.out_channels
    x0, x1 = part["x0"], part["x1"]
    out = torch.zeros((C, x1 - x0) + SHAPE[1:])
    cnt = torch.zeros((1, x1 - x0) + SHAPE[1:])
    imap = osw.importance_map(ROI, MODE)
    with torch.no_grad():
        for a0 in range(part["a0_begin"], part["a0_end"]):
            for s1 in sched.starts[1]:
                for s2 in sched.starts[2]:
                    s = s0[a0]
                    seg = net(vol[:, :, s:s + 16, s1:s1 + 16, s2:s2 + 16])[0] * imap
                    lo, hi = max(s, x0), min(s + 16, x1)
                    if hi <= lo:
                        continue
                    out[:, lo - x0:hi - x0, s1:s1 + 16, s2:s2 + 16] += seg[:, lo - s:hi - s]
                    cnt[:, lo - x0:hi - x0, s1:s1 + 16, s2:s2 + 16] += imap[lo - s:hi - s]
    return out / cnt


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from segmantic_b200.seg.multi_gpu import gather_label_slabs, rank_slab
        torch.set_num_threads(2)
        net, _ = make_oracle_net(3, 1, 3, seed=12, channels=(16, 32, 48), strides=(2, 2))
        vol = normalized_volume(SHAPE, seed=5)[None]
        sched, parts, part = rank_slab(SHAPE, ROI, OVERLAP, MODE, rank, world)
        logits = _slab_oracle(vol, net, sched, part)
        full = gather_label_slabs(logits.argmax(0).to(torch.uint8), parts, dst=0)
        if rank == 0:
            with torch.no_grad():
                ref = osw.sliding_window_inference(vol, ROI, 1, net, overlap=OVERLAP, mode=MODE)  # batch 1: same conv calls
            torch.save(dict(full=full, ref=ref[0].argmax(0).to(torch.uint8),
                            slab_equal=bool(torch.equal(logits, ref[0][:, part["x0"]:part["x1"]]))), tmp)
    finally:
        dist.destroy_process_group()


def test_slab_partition_and_gather_world2(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["full"].shape == SHAPE
    assert res["slab_equal"], "a slab's blended logits differ from the single-process result"
    assert torch.equal(res["full"], res["ref"])


# ---------------------------------------------------------------------------------------------------------------
# Window-ownership partition (the near-linear driver): every window is computed once, the tail of a rank's windows
# that covers the next rank's planes is sent forward, and every rank blends its planes in MONAI's window order.
def _owned_worker(rank, world, port, tmp, shape):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from segmantic_b200.seg.multi_gpu import gather_label_slabs, rank_windows
        torch.set_num_threads(2)
        net, _ = make_oracle_net(3, 1, 3, seed=12, channels=(16, 32, 48), strides=(2, 2))
        vol = normalized_volume(shape, seed=5)[None]
        sched, parts, part = rank_windows(shape, ROI, OVERLAP, MODE, rank, world)
        wins = sched.windows()
        imap = osw.importance_map(ROI, MODE)
        C = net.out_channels
        w_lo, w_hi, wb, send_lo = part["w_lo"], part["w_hi"], part["wb"], part["send_lo"]
        wl = torch.zeros((max(w_hi - wb, 1), C) + ROI)
        with torch.no_grad():
            for w in range(w_lo, w_hi):  # own windows only: nothing is computed twice
                s0, s1, s2 = wins[w]
                assert part["vol_x0"] <= s0 and s0 + ROI[0] <= part["vol_x1"]
                wl[w - wb] = net(vol[:, :, s0:s0 + 16, s1:s1 + 16, s2:s2 + 16])[0] * imap
        reqs = []
        if rank > 0 and w_lo > wb:
            reqs.append(dist.irecv(wl[: w_lo - wb], src=rank - 1))
        if rank + 1 < world and w_hi > send_lo:
            reqs.append(dist.isend(wl[send_lo - wb: w_hi - wb].contiguous(), dst=rank + 1))
        for r in reqs:
            r.wait()
        x0, x1 = part["x0"], part["x1"]
        out = torch.zeros((C, max(x1 - x0, 0)) + shape[1:])
        cnt = torch.zeros((1, max(x1 - x0, 0)) + shape[1:])
        per_row = len(sched.starts[1]) * len(sched.starts[2])
        for w in range(part["b_begin"] * per_row, part["b_end"] * per_row):  # MONAI order over the rows held
            s0, s1, s2 = wins[w]
            lo, hi = max(s0, x0), min(s0 + 16, x1)
            if hi <= lo:
                continue
            out[:, lo - x0:hi - x0, s1:s1 + 16, s2:s2 + 16] += wl[w - wb][:, lo - s0:hi - s0]
            cnt[:, lo - x0:hi - x0, s1:s1 + 16, s2:s2 + 16] += imap[lo - s0:hi - s0]
        logits = out / cnt
        slabs = [dict(x0=p["x0"], x1=max(p["x1"], p["x0"])) for p in parts]
        full = gather_label_slabs(logits.argmax(0).to(torch.uint8), slabs, dst=0)
        with torch.no_grad():
            ref = osw.sliding_window_inference(vol, ROI, 1, net, overlap=OVERLAP, mode=MODE)
        equal = bool(torch.equal(logits, ref[0][:, x0:x1])) if x1 > x0 else True
        flags = [None] * world
        dist.all_gather_object(flags, equal)
        if rank == 0:
            torch.save(dict(full=full, ref=ref[0].argmax(0).to(torch.uint8), slab_equal=all(flags),
                            owned=[p["w_hi"] - p["w_lo"] for p in parts]), tmp)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,port", [(2, (56, 24, 20), 29519), (3, (72, 24, 20), 29521)])
def test_window_ownership_exchange(tmp_path, world, shape, port):
    out = str(tmp_path / "res.pt")
    mp.spawn(_owned_worker, args=(world, port, out, shape), nprocs=world, join=True)
    res = torch.load(out)
    assert res["full"].shape == shape
    assert sum(res["owned"]) == len(set(range(sum(res["owned"]))))  # every window owned exactly once
    assert max(res["owned"]) - min(res["owned"]) <= 1               # even split
    assert res["slab_equal"], "a rank's blended logits differ from the single-process result"
    assert torch.equal(res["full"], res["ref"])


def test_window_partition_covers_configs():
    from segmantic_b200.seg.sliding_window import make_schedule, window_partition
    for shape, world in (((512, 256, 256), 2), ((1024, 256, 256), 4), ((2048, 256, 256), 8), ((1024, 512, 512), 8)):
        sched = make_schedule(shape, (96, 96, 96), 0.5, "gaussian")
        parts = window_partition(sched, world)
        per_row = len(sched.starts[1]) * len(sched.starts[2])
        assert parts[0]["w_lo"] == 0 and parts[-1]["w_hi"] == sched.n_windows
        assert parts[0]["x0"] == 0 and parts[-1]["x1"] == shape[0]
        for r, p in enumerate(parts):
            if r:
                assert p["w_lo"] == parts[r - 1]["w_hi"] and p["x0"] == parts[r - 1]["x1"]
                assert parts[r - 1]["send_lo"] == max(p["wb"], parts[r - 1]["w_lo"])  # what r-1 sends is what r lacks
            assert p["wb"] <= p["w_lo"] and p["b_end"] * per_row <= p["w_hi"]        # rows blended are all present
            assert p["wb"] == min(p["b_begin"] * per_row, p["w_lo"])
