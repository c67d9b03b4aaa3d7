"""Profiling driver for ncu: `fwd B reps` = forwards of B ROI windows (96^3, 10 classes, bf16);
`sw nwin reps` = sliding-window prediction of a strip holding nwin windows (blend-mode head)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmantic_b200.seg import engine  # noqa: E402
from segmantic_b200.synthetic import synthetic_state_dict  # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sd = synthetic_state_dict(3, 1, 10, seed=0)
net = engine.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=10, device=dev, precision="bf16")
if mode == "fwd":
    x = torch.randn((B, 1, 96, 96, 96), device=dev)
    for _ in range(reps):
        y = net(x)
else:
    x = torch.randn((1, 1, 96, 96, 96 + 48 * (B - 1)), device=dev)
    for _ in range(reps):
        y = engine.sliding_window_inference(x, (96, 96, 96), 4, net, overlap=0.5, mode="gaussian",
                                            return_labels=True, return_logits=False)["labels"]
net.check()
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
