"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference's own stack (MONAI / Lightning / SimpleITK) cannot be imported in this image, so the
fixtures pin the *restatement* (oracle/) -- built on the real torch CPU primitives -- rather than the
reference itself ("parity unpinned", see oracle/__init__.py).  Run from the repo root:
    python tests/golden/make_golden.py
Inputs are regenerated from seeds by the tests; only outputs are stored (float16 / uint8, < 1 MB).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import itk_resample as oitk  # noqa: E402
from oracle import sliding_window as osw  # noqa: E402
from oracle import spacing as osp  # noqa: E402
from oracle.predict import predict_volume  # noqa: E402
from tests.helpers import make_oracle_net, normalized_volume  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SMALL = dict(channels=(16, 32, 48), strides=(2, 2))


def main():
    torch.set_num_threads(4)
    # 1. UNet forward (small 3-level net, 2 input channels, 4 classes)
    net, sd = make_oracle_net(3, 2, 4, seed=11, **SMALL)
    x = normalized_volume((16, 24, 32), seed=21, channels=2)[None]
    with torch.no_grad():
        y = net(x)
    # 2. sliding window, gaussian 0.5, volume not a multiple of the roi
    net1, sd1 = make_oracle_net(3, 1, 3, seed=12, **SMALL)
    vol = normalized_volume((32, 24, 28), seed=22)[None]
    with torch.no_grad():
        sw = osw.sliding_window_inference(vol, (16, 16, 16), 4, net1, overlap=0.5, mode="gaussian")
        swc = osw.sliding_window_inference(vol, (16, 16, 16), 4, net1, overlap=0.25, mode="constant")
    # 3. full predict() composition with anisotropic spacing (config-3 style, scaled down)
    raw = normalized_volume((48, 40, 12), seed=23) * 100.0 + 50.0
    aff = osp.itk_geometry_to_ras_affine((0.5, 0.5, 3.0), (-12.0, -10.0, 0.0), np.eye(3).flatten())
    lab_logits, _ = predict_volume(net1, raw, aff, (1.0, 1.0, 1.0), roi=(16, 16, 16), invert="logits")
    lab_labels, _ = predict_volume(net1, raw, aff, (1.0, 1.0, 1.0), roi=(16, 16, 16), invert="labels")
    # 4. ITK resample of the reference's labelfield fixture (tests/conftest.py:7-13)
    lab = np.zeros((5, 5, 5), np.uint8)
    for k in range(5):
        lab[:, :, k] = k
    img = oitk.Image(lab, (0.5, 0.6, 0.7))
    r_near = oitk.resample(img, (0.25, 0.3, 0.35), nearest=True).array
    r_lin = oitk.resample(img, (0.25, 0.3, 0.35), nearest=False).array
    ref = oitk.Image(np.zeros((12, 10, 7), np.uint16), (0.25, 0.3, 0.35), (1.3, -2.1, 0.75))
    r_ref = oitk.resample_to_ref(img, ref, True).array
    np.savez_compressed(os.path.join(OUT, "oracle_golden.npz"),
                        unet_forward=y.numpy().astype(np.float32),
                        sw_gauss=sw.numpy().astype(np.float32), sw_const=swc.numpy().astype(np.float32),
                        predict_logits_mode=lab_logits.numpy(), predict_labels_mode=lab_labels.numpy(),
                        itk_near=r_near, itk_lin=r_lin, itk_ref=r_ref)
    print("wrote", os.path.join(OUT, "oracle_golden.npz"), os.path.getsize(os.path.join(OUT, "oracle_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
