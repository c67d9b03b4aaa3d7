"""Layer schema of the MONAI UNet that segmantic builds, and BatchNorm folding.

The reference constructs ``UNet(spatial_dims, in_channels=num_channels, out_channels=num_classes,
channels, strides, dropout, num_res_units=2, norm=BATCH, act="PRELU")``
(``/root/reference/src/segmantic/seg/monai_unet.py:114-124``) and runs it frozen in eval mode for
prediction (``:576-577``), so every BatchNorm is a per-channel affine of its running statistics and
folds into the preceding convolution.  This module enumerates the convolutions in the canonical
order the C-ABI expects (``include/segmantic_b200.h``: ``sgm_unet_create``), names the
``state_dict`` keys of each (plain MONAI ``model.*`` as written by ``scripts/extract_unet.py:17-18``;
Lightning checkpoints prefix ``_model.``), and folds BN in float64.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

BN_EPS = 1e-5  # torch.nn.BatchNorm default, used by MONAI Norm.BATCH

# layer kinds (mirrored in include/segmantic_b200.h)
KIND_CONV = 0
KIND_CONV_TRANSPOSE = 1
KIND_IDENTITY = 2


@dataclass
class ConvSpec:
    role: str            # e.g. "down0.unit0", "bottom.residual", "up2.convT", "up0.ru"
    kind: int            # KIND_*
    cin: int
    cout: int
    kernel: int          # 3 or 1 (per spatial axis actually used)
    stride: int          # 1 or 2
    has_adn: bool        # BatchNorm + PReLU follow the conv
    key: Optional[str]   # state_dict prefix of the Convolution / Conv module (None for identity)
    bare_conv: bool = False  # key names an nn.Conv directly (residual branch), not a Convolution


def _prefix(level: int) -> str:
    return "model." + "1.submodule." * level


def unet_conv_specs(in_channels: int, out_channels: int,
                    channels: Sequence[int] = (16, 32, 64, 128, 256),
                    strides: Sequence[int] = (2, 2, 2, 2)) -> List[ConvSpec]:
    """Canonical flat list: down levels (unit0, unit1, residual), bottom (same), up levels deepest first
    (transposed conv, residual-unit conv)."""
    channels = tuple(int(c) for c in channels)
    n = len(channels) - 1
    strides = tuple(int(s) for s in strides)[:n]
    if n < 1 or len(strides) < n:
        raise ValueError("the length of `strides` should equal `len(channels) - 1`")
    for s in strides:
        if s not in (1, 2):
            raise ValueError(f"strides must be 1 or 2, got {s}")
    specs: List[ConvSpec] = []

    def residual(role, cin, cout, stride, key):
        if stride != 1:
            return ConvSpec(role, KIND_CONV, cin, cout, 3, stride, False, key, True)
        if cin != cout:
            return ConvSpec(role, KIND_CONV, cin, cout, 1, 1, False, key, True)
        return ConvSpec(role, KIND_IDENTITY, cin, cout, 1, 1, False, None, True)

    inc = in_channels
    for i in range(n):
        c, s, p = channels[i], strides[i], _prefix(i)
        specs.append(ConvSpec(f"down{i}.unit0", KIND_CONV, inc, c, 3, s, True, p + "0.conv.unit0"))
        specs.append(ConvSpec(f"down{i}.unit1", KIND_CONV, c, c, 3, 1, True, p + "0.conv.unit1"))
        specs.append(residual(f"down{i}.residual", inc, c, s, p + "0.residual"))
        inc = c
    pb = _prefix(n - 1) + "1.submodule."
    cb = channels[n]
    specs.append(ConvSpec("bottom.unit0", KIND_CONV, inc, cb, 3, 1, True, pb + "conv.unit0"))
    specs.append(ConvSpec("bottom.unit1", KIND_CONV, cb, cb, 3, 1, True, pb + "conv.unit1"))
    specs.append(residual("bottom.residual", inc, cb, 1, pb + "residual"))
    for i in range(n - 1, -1, -1):
        c, s, p = channels[i], strides[i], _prefix(i)
        upc = c + channels[i + 1] if i == n - 1 else 2 * c
        outc = out_channels if i == 0 else channels[i - 1]
        specs.append(ConvSpec(f"up{i}.convT", KIND_CONV_TRANSPOSE, upc, outc, 3, s, True, p + "2.0"))
        specs.append(ConvSpec(f"up{i}.ru", KIND_CONV, outc, outc, 3, 1, i != 0, p + "2.1.conv.unit0"))
    return specs


def state_dict_schema(spatial_dims: int, in_channels: int, out_channels: int,
                      channels=(16, 32, 64, 128, 256), strides=(2, 2, 2, 2)) -> Dict[str, tuple]:
    """Ordered ``{key: shape}`` of the plain MONAI state_dict."""
    out: Dict[str, tuple] = {}
    for sp in unet_conv_specs(in_channels, out_channels, channels, strides):
        if sp.kind == KIND_IDENTITY:
            continue
        kk = (sp.kernel,) * spatial_dims
        wshape = ((sp.cin, sp.cout) if sp.kind == KIND_CONV_TRANSPOSE else (sp.cout, sp.cin)) + kk
        ck = sp.key if sp.bare_conv else sp.key + ".conv"
        out[ck + ".weight"] = wshape
        out[ck + ".bias"] = (sp.cout,)
        if sp.has_adn:
            nk = sp.key + ".adn.N"
            out[nk + ".weight"] = (sp.cout,)
            out[nk + ".bias"] = (sp.cout,)
            out[nk + ".running_mean"] = (sp.cout,)
            out[nk + ".running_var"] = (sp.cout,)
            out[nk + ".num_batches_tracked"] = ()
            out[sp.key + ".adn.A.weight"] = (1,)
    return out


@dataclass
class FoldedConv:
    spec: ConvSpec
    weight: Optional[torch.Tensor]  # float32, torch layout, BN folded ([O,I,k..] / convT [I,O,k..])
    bias: Optional[torch.Tensor]    # float32 [O]
    alpha: float                    # PReLU slope (0 if no activation)


def strip_lightning_prefix(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in state_dict.items():
        out[k[len("_model."):] if k.startswith("_model.") else k] = v
    return out


def fold_batchnorm(state_dict: Dict[str, torch.Tensor], specs: List[ConvSpec]) -> List[FoldedConv]:
    """``y = (conv(x)+b-mu)*g/sqrt(var+eps)+beta``  =>  ``W' = W*s``, ``b' = (b-mu)*s+beta`` in float64."""
    sd = strip_lightning_prefix(state_dict)
    folded: List[FoldedConv] = []
    for sp in specs:
        if sp.kind == KIND_IDENTITY:
            folded.append(FoldedConv(sp, None, None, 0.0))
            continue
        ck = sp.key if sp.bare_conv else sp.key + ".conv"
        try:
            w = sd[ck + ".weight"].detach().to(torch.float64)
            b = sd[ck + ".bias"].detach().to(torch.float64)
        except KeyError as e:
            raise KeyError(f"checkpoint is missing {e} (layer {sp.role}); not a MONAI UNet "
                           f"state_dict with num_res_units=2?") from None
        exp_io = (sp.cin, sp.cout) if sp.kind == KIND_CONV_TRANSPOSE else (sp.cout, sp.cin)
        if tuple(w.shape[:2]) != exp_io or any(k != sp.kernel for k in w.shape[2:]):
            raise ValueError(f"{ck}.weight has shape {tuple(w.shape)}, expected {exp_io}+({sp.kernel},)*d")
        alpha = 0.0
        if sp.has_adn:
            nk = sp.key + ".adn.N"
            g = sd[nk + ".weight"].detach().to(torch.float64)
            beta = sd[nk + ".bias"].detach().to(torch.float64)
            mu = sd[nk + ".running_mean"].detach().to(torch.float64)
            var = sd[nk + ".running_var"].detach().to(torch.float64)
            s = g / torch.sqrt(var + BN_EPS)
            ch_dim = 1 if sp.kind == KIND_CONV_TRANSPOSE else 0
            shape = [1] * w.dim()
            shape[ch_dim] = -1
            w = w * s.reshape(shape)
            b = (b - mu) * s + beta
            a = sd[sp.key + ".adn.A.weight"].detach().flatten()
            if a.numel() != 1:
                raise ValueError("PReLU with per-channel slopes is not produced by the reference (act='PRELU')")
            alpha = float(a[0])
        folded.append(FoldedConv(sp, w.to(torch.float32).contiguous(), b.to(torch.float32).contiguous(), alpha))
    return folded
