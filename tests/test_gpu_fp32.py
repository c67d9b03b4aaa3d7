"""GPU parity: fp32 CUDA path (through the C ABI) vs the CPU oracle.  Tolerance for floating point:
1e-4 relative (BASELINE.json north_star, fp32); labels bit-exact except documented argmax near-ties."""
import numpy as np
import pytest
import torch

from oracle import sliding_window as osw
from tests.helpers import (label_mismatch_outside_ties, make_oracle_net, normalized_volume, rel_err)

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4


def _engine():
    from segmantic_b200.seg import engine
    return engine


@pytest.mark.parametrize("cout,roi,batch", [(3, (32, 32, 32), 2), (10, (48, 32, 64), 1)])
def test_forward_fp32_matches_oracle(cuda_device, cout, roi, batch):
    eng = _engine()
    onet, sd = make_oracle_net(3, 1, cout, seed=1)
    x = torch.stack([normalized_volume(roi, seed=10 + b) for b in range(batch)])
    with torch.no_grad():
        ref = onet(x)
    net = eng.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=cout, device=cuda_device, precision="fp32")
    out = net(x.to(cuda_device)).cpu()
    assert out.shape == ref.shape
    assert rel_err(out, ref) < FP32_TOL


def test_forward_fp32_roi96(cuda_device):
    eng = _engine()
    onet, sd = make_oracle_net(3, 1, 3, seed=0)
    x = normalized_volume((96, 96, 96), seed=0)[None]
    with torch.no_grad():
        ref = onet(x)
    net = eng.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=3, device=cuda_device, precision="fp32")
    out = net(x.to(cuda_device)).cpu()
    assert rel_err(out, ref) < FP32_TOL
    bad, total = label_mismatch_outside_ties(ref[0], ref[0].argmax(0), out[0].argmax(0), 1e-3 * float(ref.abs().max()))
    assert bad == 0, f"{bad} label mismatches outside near-ties ({total} total)"


def test_forward_2d_two_channels(cuda_device):
    eng = _engine()
    onet, sd = make_oracle_net(2, 2, 10, seed=3)
    x = torch.stack([normalized_volume((64, 96), seed=5 + b, channels=2) for b in range(3)])
    with torch.no_grad():
        ref = onet(x)
    net = eng.UNetB200(sd, spatial_dims=2, in_channels=2, out_channels=10, device=cuda_device, precision="fp32")
    out = net(x.to(cuda_device)).cpu()
    assert rel_err(out, ref) < FP32_TOL


def test_forward_small_unet_mixed_strides(cuda_device):
    eng = _engine()
    ch, st = (8, 16, 24), (2, 1)
    onet, sd = make_oracle_net(3, 2, 4, ch, st, seed=4)
    x = normalized_volume((16, 24, 32), seed=2, channels=2)[None]
    with torch.no_grad():
        ref = onet(x)
    net = eng.UNetB200(sd, spatial_dims=3, in_channels=2, out_channels=4, channels=ch, strides=st,
                       device=cuda_device, precision="fp32")
    out = net(x.to(cuda_device)).cpu()
    assert rel_err(out, ref) < FP32_TOL


@pytest.mark.parametrize("shape,roi,overlap,mode", [
    ((70, 64, 90), (32, 32, 48), 0.25, "constant"),
    ((70, 64, 90), (32, 32, 48), 0.5, "gaussian"),
    ((20, 40, 50), (32, 32, 32), 0.25, "gaussian"),   # axis 0 smaller than the roi -> symmetric padding
])
def test_sliding_window_fp32_matches_oracle(cuda_device, shape, roi, overlap, mode):
    eng = _engine()
    onet, sd = make_oracle_net(3, 1, 5, seed=2)
    vol = normalized_volume(shape, seed=7)[None]
    with torch.no_grad():
        ref = osw.sliding_window_inference(vol, roi, 4, onet, overlap=overlap, mode=mode)
    net = eng.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=5, device=cuda_device, precision="fp32")
    res = eng.sliding_window_inference(vol.to(cuda_device), roi, 3, net, overlap=overlap, mode=mode,
                                       return_labels=True, return_probs=True)
    out = res["logits"].cpu()
    assert out.shape == ref.shape
    assert rel_err(out, ref) < FP32_TOL
    labels = res["labels"].cpu()[0, 0].long()
    # labels are the argmax of OUR logits exactly (ties -> lowest index) ...
    assert torch.equal(labels, out[0].argmax(0))
    # ... and equal the oracle's except at near-ties
    bad, total = label_mismatch_outside_ties(ref[0], ref[0].argmax(0), labels, 1e-3 * float(ref.abs().max()))
    assert bad == 0, f"{bad} label mismatches outside near-ties ({total} total)"
    probs = res["probs"].cpu()
    assert float((probs - torch.softmax(ref, 1)).abs().max()) < 1e-3
    assert float((probs.sum(1) - 1).abs().max()) < 1e-5


def test_sliding_window_count_is_exact(cuda_device):
    """With a predictor-independent check: acc/count equals the oracle's out when logits == 1."""
    eng = _engine()
    # A network whose output is constant: zero weights, bias 1 in the head -> blended logits must be
    # exactly bias (acc = sum w, count = sum w, same order -> ratio exactly 1*bias).
    onet, sd = make_oracle_net(3, 1, 2, seed=0)
    for k in sd:
        if k.endswith(".weight") and sd[k].dim() > 1:
            sd[k] = torch.zeros_like(sd[k])
    sd["model.2.0.adn.N.weight"] = torch.zeros(2)   # u = PReLU(BN(..)) == 0 exactly, folded or not
    sd["model.2.0.adn.N.bias"] = torch.zeros(2)
    sd["model.2.1.conv.unit0.conv.bias"] = torch.tensor([1.37, -2.11])
    from oracle.unet import load_checkpoint_into
    load_checkpoint_into(onet, sd)
    vol = normalized_volume((50, 40, 70), seed=1)[None]
    with torch.no_grad():
        ref = osw.sliding_window_inference(vol, (32, 32, 32), 4, onet, overlap=0.5, mode="gaussian")
    net = eng.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=2, device=cuda_device, precision="fp32")
    out = eng.sliding_window_inference(vol.to(cuda_device), (32, 32, 32), 4, net, overlap=0.5, mode="gaussian").cpu()
    assert torch.equal(out, ref), "blend accumulation / count order differs from MONAI's"


def test_slab_partition_is_bit_identical(cuda_device):
    eng = _engine()
    from segmantic_b200.seg.sliding_window import make_schedule, slab_partition
    onet, sd = make_oracle_net(3, 1, 4, seed=5)
    vol = normalized_volume((150, 48, 64), seed=3)[None].to(cuda_device)
    roi = (32, 32, 32)
    net = eng.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=4, device=cuda_device, precision="fp32")
    full = eng.sliding_window_inference(vol, roi, 4, net, overlap=0.5, mode="gaussian", return_labels=True)
    sched = make_schedule(vol.shape[2:], roi, 0.5, "gaussian")
    for world in (2, 3, 4):
        parts = slab_partition(sched, world)
        logits = torch.cat([eng.sliding_window_inference(vol, roi, 4, net, overlap=0.5, mode="gaussian", slab=p)
                            for p in parts], dim=2)
        assert torch.equal(logits, full["logits"]), f"world={world}"


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_window_ownership_is_bit_identical(cuda_device, precision):
    """The near-linear multi-GPU form on ONE device: ranks are simulated one after the other and the NVLink
    exchange is a plain copy.  Every window is computed once; each rank's planes equal the single-device run."""
    eng = _engine()
    from segmantic_b200.seg.sliding_window import make_schedule, window_partition
    onet, sd = make_oracle_net(3, 1, 4, seed=5)
    vol = normalized_volume((150, 48, 64), seed=3)[None].to(cuda_device)
    roi = (32, 32, 32)
    net = eng.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=4, device=cuda_device, precision=precision)
    full = eng.sliding_window_inference(vol, roi, 4, net, overlap=0.5, mode="gaussian", return_labels=True)
    sched = make_schedule(vol.shape[2:], roi, 0.5, "gaussian")
    for world in (2, 3, 4):
        parts = window_partition(sched, world)
        runs = [eng.OwnedWindows(vol[0, :, p["vol_x0"]:p["vol_x1"]], vol.shape[2:], p, roi, 4, net, 0.5, "gaussian",
                                 tag=str(r)) for r, p in enumerate(parts)]
        assert sum(p["w_hi"] - p["w_lo"] for p in parts) == sched.n_windows
        for r, run in enumerate(runs):
            run.compute_tail()
            run.compute_rest()
            if r + 1 < world and run.send_view.numel():
                assert runs[r + 1].recv_view.numel() == run.send_view.numel()
                runs[r + 1].recv_view.copy_(run.send_view)  # stands in for isend / irecv
        outs = [run.blend(return_logits=True) for run in runs]
        logits = torch.cat([o["logits"] for o in outs if "logits" in o], dim=1)
        labels = torch.cat([o["labels"] for o in outs], dim=0)
        assert torch.equal(logits, full["logits"][0]), f"world={world}"
        assert torch.equal(labels, full["labels"][0, 0]), f"world={world}"


def test_slice_stack_with_2d_network_matches_per_slice_oracle(cuda_device):
    """BASELINE configs[4] in small: a 2-channel 2-D UNet over a stack of slices.  The stack call (one schedule with
    roi (1, h, w), all slices' windows batched) must equal the oracle's 2-D sliding window slice by slice, and the
    device's own single-slice 2-D call bit for bit."""
    eng = _engine()
    onet, sd = make_oracle_net(2, 2, 4, seed=5)
    stack = normalized_volume((6, 80, 72), seed=13, channels=2)       # [C, Z, X, Y]
    roi = (48, 48)
    net = eng.UNetB200(sd, spatial_dims=2, in_channels=2, out_channels=4, device=cuda_device, precision="fp32")
    res = eng.sliding_window_inference(stack[None].to(cuda_device), roi, 4, net, overlap=0.25, mode="gaussian",
                                       return_labels=True)
    out = res["logits"].cpu()
    assert out.shape == (1, 4, 6, 80, 72) and res["labels"].shape == (1, 1, 6, 80, 72)
    for z in range(6):
        sl = stack[:, z][None]
        with torch.no_grad():
            ref = osw.sliding_window_inference(sl, roi, 4, onet, overlap=0.25, mode="gaussian")
        assert rel_err(out[:, :, z], ref) < FP32_TOL
        one = eng.sliding_window_inference(sl.to(cuda_device), roi, 4, net, overlap=0.25, mode="gaussian")
        assert torch.equal(one.cpu(), out[:, :, z])
    assert torch.equal(res["labels"].cpu()[0, 0].long(), out[0].argmax(0))


def test_predict_stack_2d(cuda_device):
    from segmantic_b200.seg.monai_unet import Net, predict_stack_2d
    onet, sd = make_oracle_net(2, 2, 4, seed=5)
    net = Net(num_classes=4, num_channels=2, spatial_dims=2, spatial_size=[48, 48])
    net.load_state_dict(sd)
    net.to(cuda_device)
    raw = normalized_volume((72, 64, 5), seed=17, channels=2) * 20.0 + 3.0   # [C, X, Y, Z]
    lab = predict_stack_2d(net, raw)
    assert lab.shape == (72, 64, 5) and lab.dtype == torch.uint8
    norm = torch.stack([(raw[c] - raw[c].mean()) / raw[c].std(unbiased=False) for c in range(2)])
    bad = 0
    for z in range(5):
        with torch.no_grad():
            ref = osw.sliding_window_inference(norm[:, :, :, z][None], (48, 48), 4, onet)
        b, _ = label_mismatch_outside_ties(ref[0], ref[0].argmax(0), lab[:, :, z].long(), 1e-3 * float(ref.abs().max()))
        bad += b
    assert bad == 0
