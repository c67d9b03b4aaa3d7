"""Exploration (GPU box): how many training steps the confident checkpoint of tests/confident.py needs before the
bf16-emulating oracle meets north_star's tolerances against the fp32 oracle (prints one line per variant)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bf16_emulation import bf16_forward  # noqa: E402
from oracle.unet import UNet, load_checkpoint_into  # noqa: E402
from tests.confident import tissue_phantom, train_confident_state_dict  # noqa: E402
from tests.helpers import dice_per_class  # noqa: E402

dev = "cuda" if torch.cuda.is_available() else "cpu"
for n_classes, steps, ls in ((10, 1000, 0.0), (10, 1000, 0.1), (10, 2000, 0.1), (10, 1000, 0.2), (10, 600, 0.1)):
    t0 = time.time()
    sd = train_confident_state_dict(n_classes, steps=steps, device=dev, label_smoothing=ls)
    t1 = time.time()
    net = UNet(3, 1, n_classes)
    load_checkpoint_into(net, sd)
    net.eval()
    v, l = tissue_phantom((192, 192, 192), n_classes, seed=7)
    with torch.no_grad():
        ref = net(v[None])[0]
        b16 = bf16_forward(net, sd, v[None])[0]
    p, q = torch.softmax(ref, 0), torch.softmax(b16, 0)
    lr, lb = ref.argmax(0), b16.argmax(0)
    d = dice_per_class(lb, lr, n_classes)
    print(f"classes {n_classes} steps {steps} smoothing {ls}: train {t1 - t0:.1f}s acc {float((lr == l).float().mean()):.5f} "
          f"max prob err {float((p - q).abs().max()):.4f} mismatches {int((lr != lb).sum())} min dice {min(d):.5f} "
          f"logits [{float(ref.min()):.1f}, {float(ref.max()):.1f}]", flush=True)
