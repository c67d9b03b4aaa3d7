"""Seeded synthetic inputs and checkpoints (there is no dataset / trained model offline).

The volume is a smooth CT-like field (air background, soft-tissue and bone ellipsoids, low-passed
noise); the checkpoint is a MONAI-UNet ``state_dict`` (schema: ``unet_spec.state_dict_schema``) with
variance-preserving random conv weights and *randomised* BatchNorm statistics / PReLU slopes so that
BN folding and the activation are really exercised.  CPU generators only, so
the same seed gives the same bits here and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

from .seg.unet_spec import state_dict_schema


def synthetic_volume(shape: Sequence[int], seed: int = 0, channels: int = 1) -> torch.Tensor:
    """``[channels, *shape]`` float32 HU-like field."""
    g = torch.Generator().manual_seed(seed)
    shape = tuple(int(s) for s in shape)
    nd = len(shape)
    out = []
    for c in range(channels):
        coarse = tuple(max(4, s // 16) for s in shape)
        low = torch.randn((1, 1) + coarse, generator=g)
        mode = "trilinear" if nd == 3 else "bilinear"
        field = F.interpolate(low, size=shape, mode=mode, align_corners=True)[0, 0] * 60.0
        field += torch.randn(shape, generator=g) * 8.0
        vol = torch.full(shape, -1000.0) + field
        axes = [torch.linspace(-1.0, 1.0, s) for s in shape]
        grid = torch.meshgrid(*axes, indexing="ij")
        # body, organ, bone ellipsoids
        for level, centre, radii in (
            (40.0, (0.0,) * nd, (0.85, 0.7, 0.9)[:nd]),
            (120.0, (0.2, -0.1, 0.1)[:nd], (0.35, 0.3, 0.45)[:nd]),
            (700.0, (-0.3, 0.25, -0.2)[:nd], (0.18, 0.22, 0.5)[:nd]),
            (700.0, (0.45, 0.3, 0.3)[:nd], (0.1, 0.12, 0.3)[:nd]),
        ):
            r2 = sum(((gr - ce) / ra) ** 2 for gr, ce, ra in zip(grid, centre, radii))
            vol = torch.where(r2 < 1.0, level + field, vol)
        if c > 0:  # correlated second modality
            vol = 0.6 * vol + 0.4 * out[0] + torch.randn(shape, generator=g) * 20.0
        out.append(vol.to(torch.float32))
    return torch.stack(out, 0)


def synthetic_state_dict(spatial_dims: int, in_channels: int, out_channels: int,
                         channels=(16, 32, 64, 128, 256), strides=(2, 2, 2, 2),
                         seed: int = 0, lightning_prefix: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    schema = state_dict_schema(spatial_dims, in_channels, out_channels, channels, strides)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in schema.items():
        if key.endswith("num_batches_tracked"):
            t = torch.tensor(100, dtype=torch.long)
        elif key.endswith("adn.A.weight"):
            t = torch.rand(shape, generator=g) * 0.3 + 0.1
        elif key.endswith("adn.N.weight"):
            t = torch.rand(shape, generator=g) + 0.5
            if key.startswith("model.2.0."):  # top-level up-sampling: keep the logits O(1)
                t = t * 0.12
        elif key.endswith("adn.N.bias"):
            t = torch.randn(shape, generator=g) * 0.1
        elif key.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        elif key.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif key.endswith(".weight"):
            is_t = ".2.0.conv.weight" in key
            fan_in = (shape[0] if is_t else shape[1]) * math.prod(shape[2:])
            if is_t:  # stride-2 transposed conv: each output sees ~1/8 of the taps
                fan_in = max(1.0, fan_in / (2 ** spatial_dims))
            t = torch.randn(shape, generator=g) * (1.3 / math.sqrt(fan_in))
        else:  # conv bias
            t = torch.randn(shape, generator=g) * 0.05
        sd[("_model." if lightning_prefix else "") + key] = t
    return sd


def synthetic_lightning_checkpoint(num_classes: int, num_channels: int = 1, spatial_dims: int = 3,
                                   spatial_size=None, channels=(16, 32, 64, 128, 256),
                                   strides=(2, 2, 2, 2), seed: int = 0) -> dict:
    """A dict with the Lightning ``.ckpt`` layout the reference loads (monai_unet.py:99-112,569-573)."""
    return {
        "state_dict": synthetic_state_dict(spatial_dims, num_channels, num_classes, channels, strides,
                                           seed, lightning_prefix=True),
        "hyper_parameters": {
            "num_classes": num_classes, "num_channels": num_channels, "spatial_dims": spatial_dims,
            "spatial_size": list(spatial_size) if spatial_size else None,
            "channels": tuple(channels), "strides": tuple(strides), "dropout": 0.0, "act": "PRELU",
        },
    }
