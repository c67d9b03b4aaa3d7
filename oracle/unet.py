"""Oracle: MONAI ``UNet`` restated with ``torch.nn`` (test infrastructure, see ``oracle/__init__``).

Follows the construction at ``/root/reference/src/segmantic/seg/monai_unet.py:114-124``::

    UNet(spatial_dims, in_channels=num_channels, out_channels=num_classes, channels, strides,
         dropout=dropout, num_res_units=2, norm=Norm.BATCH, act=act)

and ``Net.forward`` (``seg/monai_unet.py:221-222``).  Module *names* reproduce MONAI's so a MONAI /
Lightning ``state_dict`` loads with ``strict=True`` (Lightning prefixes keys with ``_model.``).
MONAI itself is not installed; the topology is restated from MONAI's published ``UNet`` /
``ResidualUnit`` / ``Convolution`` / ``ADN`` / ``SkipConnection``.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn


def _conv_cls(dims: int, transposed: bool):
    if dims == 3:
        return nn.ConvTranspose3d if transposed else nn.Conv3d
    if dims == 2:
        return nn.ConvTranspose2d if transposed else nn.Conv2d
    raise ValueError(f"spatial_dims must be 2 or 3, got {dims}")


def _bn_cls(dims: int):
    return nn.BatchNorm3d if dims == 3 else nn.BatchNorm2d


def _drop_cls(dims: int):
    # MONAI ADN uses the plain element-wise Dropout ("dropout_dim=1")
    return nn.Dropout


class ADN(nn.Sequential):
    """MONAI ADN with ordering "NDA": BatchNorm -> Dropout -> PReLU."""

    def __init__(self, dims: int, channels: int, dropout: float | None, act: str = "PRELU"):
        super().__init__()
        self.add_module("N", _bn_cls(dims)(channels))
        if dropout is not None:
            self.add_module("D", _drop_cls(dims)(dropout))
        if act.upper() != "PRELU":
            raise ValueError("oracle restates act='PRELU' only (reference default, monai_unet.py:108)")
        self.add_module("A", nn.PReLU())


class Convolution(nn.Sequential):
    """MONAI ``Convolution``: conv (+ ADN unless ``conv_only``); kernel 3, padding 1."""

    def __init__(self, dims, cin, cout, stride, kernel=3, dropout=0.0, conv_only=False,
                 transposed=False, act="PRELU"):
        super().__init__()
        pad = (kernel - 1) // 2
        if transposed:
            conv = _conv_cls(dims, True)(cin, cout, kernel, stride, pad, output_padding=stride - 1,
                                         bias=True)
        else:
            conv = _conv_cls(dims, False)(cin, cout, kernel, stride, pad, bias=True)
        self.add_module("conv", conv)
        if not conv_only:
            self.add_module("adn", ADN(dims, cout, dropout, act))


class ResidualUnit(nn.Module):
    """MONAI ``ResidualUnit``: ``conv(x) + residual(x)``; no activation after the add."""

    def __init__(self, dims, cin, cout, stride, subunits, dropout=0.0, last_conv_only=False,
                 act="PRELU"):
        super().__init__()
        self.conv = nn.Sequential()
        sc, ss = cin, stride
        for su in range(max(1, subunits)):
            only = last_conv_only and su == subunits - 1
            self.conv.add_module(f"unit{su:d}",
                                 Convolution(dims, sc, cout, ss, 3, dropout, only, False, act))
            sc, ss = cout, 1
        if stride != 1 or cin != cout:
            rk, rp = (3, 1) if stride != 1 else (1, 0)
            self.residual = _conv_cls(dims, False)(cin, cout, rk, stride, rp, bias=True)
        else:
            self.residual = nn.Identity()

    def forward(self, x):
        return self.conv(x) + self.residual(x)


class SkipConnection(nn.Module):
    def __init__(self, submodule: nn.Module):
        super().__init__()
        self.submodule = submodule

    def forward(self, x):
        return torch.cat([x, self.submodule(x)], dim=1)


class UNet(nn.Module):
    """MONAI UNet with ``num_res_units=2``, ``norm=BATCH``, ``act=PRELU`` (monai_unet.py:114-124)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int,
                 channels: Sequence[int] = (16, 32, 64, 128, 256),
                 strides: Sequence[int] = (2, 2, 2, 2), dropout: float = 0.0, act: str = "PRELU"):
        super().__init__()
        if len(channels) < 2 or len(strides) < len(channels) - 1:
            raise ValueError("need len(channels) >= 2 and len(strides) >= len(channels)-1")
        self.dimensions = spatial_dims
        self.in_channels, self.out_channels = in_channels, out_channels
        self.channels, self.strides = tuple(channels), tuple(strides)
        d, drop = spatial_dims, dropout

        def block(inc, outc, ch, st, top):
            c, s = ch[0], st[0]
            if len(ch) > 2:
                sub = block(c, c, ch[1:], st[1:], False)
                upc = c * 2
            else:
                sub = ResidualUnit(d, c, ch[1], 1, 2, drop, False, act)  # bottom layer
                upc = c + ch[1]
            down = ResidualUnit(d, inc, c, s, 2, drop, False, act)
            up = nn.Sequential(
                Convolution(d, upc, outc, s, 3, drop, False, True, act),
                ResidualUnit(d, outc, outc, 1, 1, drop, top, act),
            )
            return nn.Sequential(down, SkipConnection(sub), up)

        self.model = block(in_channels, out_channels, self.channels, self.strides, True)

    def forward(self, x):
        return self.model(x)


def load_checkpoint_into(net: UNet, state_dict: dict) -> None:
    """Accept Lightning (``_model.model.*``) or plain MONAI (``model.*``) keys, strict."""
    sd = {}
    for k, v in state_dict.items():
        sd[k[len("_model."):] if k.startswith("_model.") else k] = v
    net.load_state_dict(sd, strict=True)
