"""Phase timing of the end-to-end call predict_volume(host volume) -> host labels (GPU box only)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmantic_b200.seg import transforms as T  # noqa: E402
from segmantic_b200.seg.engine import sliding_window_inference  # noqa: E402
from segmantic_b200.seg.monai_unet import Net, predict_volume  # noqa: E402
from segmantic_b200.synthetic import synthetic_state_dict, synthetic_volume  # noqa: E402

dev = torch.device("cuda:0")
sd = synthetic_state_dict(3, 1, 10, seed=0)
net = Net(num_classes=10, num_channels=1, spatial_dims=3)
net.load_state_dict(sd)
net.to(dev)
host = synthetic_volume((256, 256, 256), seed=1).contiguous().pin_memory()
kw = dict(overlap=0.5, mode="gaussian", sw_batch_size=32, precision="bf16", crop_foreground=False)
for _ in range(3):
    predict_volume(net, host, None, (), **kw)
torch.cuda.synchronize()


def timed(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{name:28s} {(time.perf_counter() - t0) * 1e3:8.3f} ms", flush=True)
    return out


for rep in range(2):
    print("--- rep", rep)
    t_all = time.perf_counter()
    img = timed("H2D 67 MB (pinned)", lambda: host.to(dev, non_blocking=True))
    import numpy as np
    aff = np.eye(4)
    aff[0, 0] = aff[1, 1] = -1.0
    o = timed("orientation_ras", lambda: T.orientation_ras(img, aff))
    img2 = timed("normalize_intensity", lambda: T.normalize_intensity(o[0]))
    eng = net.engine("bf16")
    res = timed("sliding_window_inference", lambda: sliding_window_inference(
        img2.unsqueeze(0), net.spatial_size, 32, eng, overlap=0.5, mode="gaussian", return_labels=True, return_logits=False))
    lab = res["labels"][0, 0]
    lab2 = timed("orientation_inverse", lambda: T.orientation_inverse(lab, o[2], lead=0))
    timed("eng.check", lambda: eng.check())
    timed("lab.cpu() (pageable)", lambda: lab2.cpu())
    pin = torch.empty(lab2.shape, dtype=torch.uint8).pin_memory()
    timed("D2H into pinned", lambda: pin.copy_(lab2, non_blocking=True))
    print(f"sum of phases (with syncs)   {(time.perf_counter() - t_all) * 1e3:8.3f} ms")
    timed("predict_volume (whole)", lambda: predict_volume(net, host, None, (), **kw))
