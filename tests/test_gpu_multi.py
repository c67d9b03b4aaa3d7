"""Multi-GPU driver on REAL devices (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py`):
window ownership with the NVLink peer-memory seam exchange (`seg/p2p.py`, `csrc/p2p.cu`) and with NCCL point-to-point,
one process per GPU.  Every rank's blended logits and the gathered label volume must equal the single-GPU result BIT
FOR BIT, over several volumes in a row (the exchange buffers and counters are reused from volume to volume)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import normalized_volume

pytestmark = pytest.mark.gpu

ROI, OVERLAP, MODE = (32, 32, 32), 0.5, "gaussian"
CLASSES = 10


def _worker(rank, world, port, tmp, shape, exchange, precision, steps):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from segmantic_b200.seg import engine
        from segmantic_b200.seg.multi_gpu import gather_label_slabs, rank_windows
        from segmantic_b200.synthetic import synthetic_state_dict

        sd = synthetic_state_dict(3, 1, CLASSES, seed=3)
        net = engine.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=CLASSES, device=dev, precision=precision)
        sched, parts, part = rank_windows(shape, ROI, OVERLAP, MODE, rank, world)
        ok_logits, ok_labels = [], []
        for step in range(steps):
            vol = normalized_volume(shape, seed=20 + step)  # a different volume every step
            slab = vol[:, part["vol_x0"]:part["vol_x1"]].contiguous().to(dev)
            res = engine.sliding_window_inference_owned(slab, shape, part, ROI, 4, net, overlap=OVERLAP, mode=MODE,
                                                        rank=rank, world_size=world, return_logits=True,
                                                        exchange=exchange)
            full = gather_label_slabs(res["labels"], parts, dst=0)
            net.check()
            link = net.__dict__.get("_seam")
            if link is not None:
                link.check()
            # single-GPU reference on this rank's device (same kernels, whole volume)
            ref = engine.sliding_window_inference(vol[None].to(dev), ROI, 4, net, overlap=OVERLAP, mode=MODE,
                                                  return_labels=True)
            mine = ref["logits"][0][:, part["x0"]:part["x1"]]
            ok_logits.append(bool(torch.equal(res["logits"], mine)))
            if rank == 0:
                ok_labels.append(bool(torch.equal(full, ref["labels"][0, 0])))
        flags = torch.tensor([int(all(ok_logits))], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            torch.save(dict(logits=bool(flags.item()), labels=all(ok_labels), parts=parts), tmp)
    finally:
        dist.destroy_process_group()


def _run(tmp_path, world, shape, exchange, precision, port, steps=3):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, port, out, shape, exchange, precision, steps), nprocs=world, join=True)
    res = torch.load(out)
    assert res["logits"], "a rank's blended logits differ from the single-GPU result"
    assert res["labels"], "the gathered label volume differs from the single-GPU result"


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_owned_windows_p2p_world2(tmp_path, precision):
    _run(tmp_path, 2, (112, 48, 64), "p2p", precision, 29531)


def test_owned_windows_nccl_world2(tmp_path):
    _run(tmp_path, 2, (112, 48, 64), "nccl", "bf16", 29533)


def test_owned_windows_p2p_world4(tmp_path):
    _run(tmp_path, 4, (208, 48, 48), "p2p", "bf16", 29535, steps=4)


def test_owned_windows_p2p_world8(tmp_path):
    _run(tmp_path, 8, (400, 32, 48), "p2p", "bf16", 29537, steps=3)
