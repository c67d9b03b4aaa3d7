"""Exploration (GPU box): which (steps, label smoothing, seed) of tests/confident.py gives a checkpoint on which the
DEVICE bf16 sliding-window prediction meets north_star's tolerances against the fp32 oracle at the full configs[1]
size -- and whether the recipe is reproducible run to run.  Prints one line per variant."""
import hashlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sliding_window as osw  # noqa: E402
from oracle.unet import UNet, load_checkpoint_into  # noqa: E402
from segmantic_b200.seg import engine  # noqa: E402
from tests.confident import tissue_phantom, train_confident_state_dict  # noqa: E402
from tests.helpers import dice_per_class  # noqa: E402


def digest(sd):
    h = hashlib.sha1()
    for k in sorted(sd):
        h.update(sd[k].cpu().numpy().tobytes())
    return h.hexdigest()[:10]


dev = torch.device("cuda:0")
vol, truth = tissue_phantom((256, 256, 256), 10, seed=11)
x = vol[None]
variants = [(2500, 0.1, 2), (2500, 0.1, 3), (2500, 0.1, 4), (2500, 0.1, 5), (2500, 0.1, 6), (3500, 0.1, 2)]
if len(sys.argv) > 1:
    variants = [tuple(float(v) if "." in v else int(v) for v in a.split(",")) for a in sys.argv[1:]]
for steps, ls, seed in variants:
    t0 = time.time()
    sd = train_confident_state_dict(10, steps=int(steps), seed=int(seed), device=dev, label_smoothing=float(ls))
    t1 = time.time()
    onet = UNet(3, 1, 10)
    load_checkpoint_into(onet, sd)
    onet.eval()
    with torch.no_grad():
        ref = osw.sliding_window_inference(x, (96, 96, 96), 4, onet, overlap=0.5, mode="gaussian")[0]
    net = engine.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=10, device=dev, precision="bf16")
    res = engine.sliding_window_inference(x.to(dev), (96, 96, 96), 4, net, overlap=0.5, mode="gaussian",
                                          return_labels=True, return_probs=True)
    net.check()
    p32 = torch.softmax(ref, 0)
    perr = (res["probs"].cpu()[0] - p32).abs().amax(0)
    lab, lab32 = res["labels"].cpu()[0, 0].long(), ref.argmax(0)
    d = dice_per_class(lab, lab32, 10)
    q = torch.quantile(perr.flatten()[::7].float(), torch.tensor([0.999, 0.9999]))
    print(f"steps {steps} smoothing {ls} seed {seed} sha {digest(sd)}: train {t1 - t0:.0f}s acc "
          f"{float((lab32 == truth).float().mean()):.5f} max prob err {float(perr.max()):.4f} (q99.9 {float(q[0]):.4f}, "
          f"q99.99 {float(q[1]):.4f}) mismatches {int((lab != lab32).sum())} min dice {min(d):.5f} "
          f"logits [{float(ref.min()):.1f}, {float(ref.max()):.1f}]", flush=True)
    del net
