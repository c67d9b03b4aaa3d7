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
