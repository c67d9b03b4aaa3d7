"""Parity of the remaining BASELINE configs / SURVEY.md section 8(f) rows at (near) full size against the oracle:

* configs[2]: the four-stage pipeline (Spacing to 1 mm -> sliding window -> argmax -> nearest resample back) on a
  256 x 256 x 120 crop of the anisotropic 0.5 x 0.5 x 3 mm image (28 windows on the 128 x 128 x 358 network grid);
* the `validation_step` sliding window of the reference (`monai_unet.py:350-363`: roi 160^d, sw_batch_size 1) -- at an
  extent of 160 voxels the head runs on the plane-sweep kernel (the row sweep covers 33..126 voxels);
* configs[3] in small: a 20-tissue network through the chunked single-GPU path (the two-launch transposed plane sweep
  and the plane-sweep head of the 11..32-class networks) against the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import sliding_window as osw
from oracle import spacing as osp
from oracle.bf16_emulation import bf16_forward
from oracle.predict import predict_volume as oracle_predict
from tests.helpers import make_oracle_net, normalized_volume, rel_err

pytestmark = pytest.mark.gpu


def _engine():
    from segmantic_b200.seg import engine
    return engine


@pytest.mark.parametrize("invert", ["labels", "logits"])
def test_config2_pipeline_256x256x120_crop_vs_oracle(cuda_device, invert):
    from segmantic_b200.seg.monai_unet import Net, predict_volume
    from segmantic_b200.synthetic import synthetic_volume
    onet, sd = make_oracle_net(3, 1, 10, seed=0)
    raw = synthetic_volume((256, 256, 120), seed=2)
    aff = osp.itk_geometry_to_ras_affine((0.5, 0.5, 3.0), (-64.0, -64.0, 0.0), np.eye(3).flatten())
    kw = dict(overlap=0.5, mode="gaussian", invert=invert)
    ref, ref_logits, gap = oracle_predict(onet, raw, aff, (1.0, 1.0, 1.0), roi=(96, 96, 96), return_gap=True, **kw)
    net = Net(num_classes=10, num_channels=1, spatial_dims=3)
    net.load_state_dict(sd)
    net.to(cuda_device)
    lab = predict_volume(net, raw, aff, (1.0, 1.0, 1.0), precision="fp32", **kw)
    assert lab.shape == ref.shape == (256, 256, 120) and lab.dtype == torch.uint8
    bad = lab != ref
    tol = 1e-3 * float(ref_logits.abs().max())   # fp32 path: summation order only
    outside = int((bad & (gap > tol)).sum())
    print(f"configs[2] crop, invert={invert}: {int(bad.sum())} of {bad.numel()} labels differ, {outside} outside near-ties")
    assert outside == 0
    assert float(bad.float().mean()) < 2e-3


def test_validation_step_window_roi160(cuda_device):
    """roi 160^3, sw_batch_size 1, overlap 0.25, constant blend (MONAI defaults: what `sliding_window_inference(inputs,
    (160,) * 3, 1, self.forward)` at monai_unet.py:354-356 runs) on a 160 x 160 x 208 volume: 2 windows."""
    import time
    eng = _engine()
    onet, sd = make_oracle_net(3, 1, 3, seed=4)
    vol = normalized_volume((160, 160, 208), seed=17)[None]
    roi = (160, 160, 160)
    with torch.no_grad():
        ref32 = osw.sliding_window_inference(vol, roi, 1, onet)
        ref16 = osw.sliding_window_inference(vol, roi, 1, lambda w: bf16_forward(onet, sd, w))
    for precision, ref, tol in (("fp32", ref32, 1e-4), ("bf16", ref16, 1.5e-2)):
        net = eng.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=3, device=cuda_device, precision=precision)
        x = vol.to(cuda_device)
        out = eng.sliding_window_inference(x, roi, 1, net)
        net.check()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            eng.sliding_window_inference(x, roi, 1, net)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        e = rel_err(out.cpu(), ref)
        print(f"roi 160^3 {precision}: {e:.3e} of the logit range vs the oracle; {ms:.2f} ms per 160x160x208 volume "
              f"({160 * 160 * 208 / ms / 1e3:.0f} Mvoxel/s)")
        assert e < tol


def test_config3_small_20_tissues_chunked_vs_oracle(cuda_device, monkeypatch):
    eng = _engine()
    onet, sd = make_oracle_net(3, 1, 20, seed=9)
    vol = normalized_volume((192, 96, 112), seed=19)[None]
    roi = (96, 96, 96)
    with torch.no_grad():
        ref16 = osw.sliding_window_inference(vol, roi, 4, lambda w: bf16_forward(onet, sd, w), overlap=0.5, mode="gaussian")
    net = eng.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=20, device=cuda_device, precision="bf16")
    monkeypatch.setenv("SGM_BLEND", "chunked")
    res = eng.sliding_window_inference(vol.to(cuda_device), roi, 4, net, overlap=0.5, mode="gaussian", return_labels=True)
    net.check()
    assert getattr(net, "last_launch_count_chunked", 0) > 0
    e = rel_err(res["logits"].cpu(), ref16)
    print(f"20 tissues, chunked: {e:.3e} of the logit range vs the bf16-emulating oracle")
    assert e < 1.5e-2
    assert torch.equal(res["labels"].cpu()[0, 0].long(), res["logits"].cpu()[0].argmax(0))
