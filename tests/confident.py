"""A *confident* synthetic checkpoint: seed + recipe, no file (test infrastructure).

north_star's bf16 tolerances (probabilities within 2e-2 of the fp32 reference, Dice >= 0.999 per tissue) presume a
trained network: decisive logits away from tissue borders, thin transition zones.  There is no trained model offline,
so this module TRAINS one -- the oracle's MONAI-UNet restatement (``oracle/unet.py``, the topology of
``/root/reference/src/segmantic/seg/monai_unet.py:114-124``) fitted for a few hundred Adam steps to a phantom whose
tissues are ellipsoids at distinct HU-like levels, the way ``Net.training_step`` (``monai_unet.py:327-340``) fits real
data.  The recipe is seeded, uses deterministic cuDNN algorithms and runs on whatever torch device is handed in (about 20 s
on a B200 with torch's own CUDA kernels -- test infrastructure, not the product path; hours on CPU cores, so the tests
that need it are GPU tests); the resulting ``state_dict`` has MONAI's keys, so the device path and the oracle load the very same
weights -- parity does not depend on the training being bit-reproducible.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

from oracle.unet import UNet


def tissue_phantom(shape: Sequence[int], n_classes: int, seed: int = 0, noise: float = 4.0,
                   field_sigma: float = 6.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(volume [1, *shape] float32 z-scored, labels [*shape] int64)``.

    Class 0 = air, 1 = body ellipsoid, 2.. = organs: non-overlapping ellipsoids on a jittered 2 x 2 x 2 (x more) grid
    inside the body, each at its own intensity level (levels 90 apart against voxel noise of sigma ``noise`` plus a
    smooth field of sigma ``field_sigma``: separable by intensity with a little local context, hard edges)."""
    g = torch.Generator().manual_seed(seed)
    shape = tuple(int(s) for s in shape)
    coarse = tuple(max(4, s // 16) for s in shape)
    low = torch.randn((1, 1) + coarse, generator=g)
    field = F.interpolate(low, size=shape, mode="trilinear", align_corners=True)[0, 0] * field_sigma
    vnoise = torch.randn(shape, generator=g) * noise
    axes = [torch.linspace(-1.0, 1.0, s) for s in shape]
    grid = torch.meshgrid(*axes, indexing="ij")
    labels = torch.zeros(shape, dtype=torch.long)
    level = torch.full(shape, -1000.0)
    body = sum((gr / ra) ** 2 for gr, ra in zip(grid, (0.93, 0.9, 0.95))) < 1.0
    labels[body], level[body] = 1, 0.0
    n_org = n_classes - 2
    cells = [(i, j, k) for i in range(2) for j in range(2) for k in range(-(-n_org // 4))]
    nk = -(-n_org // 4)
    for c in range(2, n_classes):
        i, j, k = cells[c - 2]
        jit = (torch.rand(3, generator=g) - 0.5) * 0.08
        centre = (-0.36 + 0.72 * i + float(jit[0]), -0.34 + 0.68 * j + float(jit[1]),
                  (-0.72 + 1.44 * (k + 0.5) / nk) * 0.62 / 0.72 + float(jit[2]))
        radii = (torch.rand(3, generator=g) * 0.05 + 0.27).tolist()
        radii[2] = min(radii[2], 0.6 / nk)
        inside = sum(((gr - ce) / ra) ** 2 for gr, ce, ra in zip(grid, centre, radii)) < 1.0
        inside &= body
        labels[inside] = c
        level[inside] = 90.0 * (c // 2) * (1.0 if c % 2 == 0 else -1.0)
    vol = level + field + vnoise
    vol = (vol - vol.mean()) / vol.std(unbiased=False)
    return vol[None].to(torch.float32), labels


def train_confident_state_dict(n_classes: int = 10, steps: int = 2500, patch: int = 48, seed: int = 0,
                               device: str | torch.device = "cpu", verbose: bool = False,
                               label_smoothing: float = 0.0) -> Dict[str, torch.Tensor]:
    """Fit the oracle UNet to ``tissue_phantom`` and return its ``state_dict`` (MONAI keys, CPU tensors)."""
    device = torch.device(device)
    # reproducible on a given GPU / library stack: deterministic cuDNN algorithms, fixed seeds (the parity tests do not
    # depend on it -- both sides load the same weights -- but a pinned recipe keeps the tolerances' margins stable)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.manual_seed(seed)
    net = UNet(3, 1, n_classes).to(device)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=2e-3)
    g = torch.Generator().manual_seed(seed + 1)
    vols = [tissue_phantom((112, 112, 112), n_classes, seed=100 + i) for i in range(2)]
    vols = [(v.to(device), l.to(device)) for v, l in vols]
    for step in range(steps):
        xs, ys = [], []
        for b in range(2):
            v, l = vols[(step + b) % len(vols)]
            o = [int(torch.randint(0, s - patch + 1, (1,), generator=g)) for s in l.shape]
            sl = tuple(slice(oo, oo + patch) for oo in o)
            xs.append(v[(slice(None),) + sl])
            ys.append(l[sl])
        x, y = torch.stack(xs), torch.stack(ys)
        for pg in opt.param_groups:
            pg["lr"] = 2e-3 * 0.5 * (1.0 + math.cos(math.pi * step / steps))
        loss = F.cross_entropy(net(x), y, label_smoothing=label_smoothing)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if verbose and step % 25 == 0:
            print(f"step {step}: loss {float(loss):.4f}", flush=True)
    net.eval()
    return {k: v.detach().to("cpu").clone() for k, v in net.state_dict().items()}


_CACHE: Dict[tuple, Dict[str, torch.Tensor]] = {}


# The pinned recipe (tests/explore_confident.py swept steps / label smoothing / seeds on a B200 at the full configs[1]
# size): label smoothing keeps the logits within +-8 (the bf16 error of a logit scales with the logit range), 3500
# steps reach 99.9 % accuracy on the phantom (min Dice of the bf16 device path against the fp32 oracle 0.99937).
RECIPE = dict(steps=3500, label_smoothing=0.1, seed=2)
# sha1 (first 10 hex digits, `state_digest`) of the checkpoint the recipe produced on the B200 boxes of this pool (three
# different boxes, identical).  The margins of tests/test_gpu_north_star.py were measured on THIS checkpoint; if another
# GPU / library stack trains a different one, those tests say so and fall back to looser bounds.
RECIPE_DIGEST = "efc78ccdfc"


def state_digest(sd: Dict[str, torch.Tensor]) -> str:
    import hashlib
    h = hashlib.sha1()
    for k in sorted(sd):
        h.update(sd[k].cpu().numpy().tobytes())
    return h.hexdigest()[:10]


def confident_state_dict(n_classes: int = 10, steps: int = RECIPE["steps"], seed: int = RECIPE["seed"],
                         label_smoothing: float = RECIPE["label_smoothing"]) -> Dict[str, torch.Tensor]:
    """Cached (per process and under ``$TMPDIR``) result of ``train_confident_state_dict``."""
    key = (n_classes, steps, seed, label_smoothing)
    if key in _CACHE:
        return _CACHE[key]
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"),
                        f"sgm_confident_c{n_classes}_s{steps}_{seed}_ls{label_smoothing}.pt")
    if os.path.exists(path):
        sd = torch.load(path, weights_only=True)
    else:
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        sd = train_confident_state_dict(n_classes, steps, seed=seed, device=dev, label_smoothing=label_smoothing)
        torch.save(sd, path)
    _CACHE[key] = sd
    return sd
