"""Oracle variant that rounds to bf16 where the bf16 CUDA path does (test infrastructure).

The bf16 device path (``SGM_PRECISION_BF16``) stores every inter-layer activation in bf16, uses
BN-folded weights rounded to bf16, keeps biases / PReLU slopes in fp32 and accumulates in fp32; the
network input (fp32 volume) is rounded to bf16 when the tensor-core stem reads it (3-D networks whose first
block is a stride-2 k3 conv with 27 * num_channels <= 64 and at most 16 output channels,
``segmantic_b200/csrc/conv_stem_tc.cu``); every other first block runs on CUDA cores and reads the fp32
volume directly with bf16-rounded weights.  This module replays the reference topology (``oracle/unet.py``; reference construction at
``/root/reference/src/segmantic/seg/monai_unet.py:114-124``) with those rounding points so that the
tensor-core kernels can be pinned tightly (summation order is then the only difference).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from segmantic_b200.seg.unet_spec import (KIND_CONV_TRANSPOSE, KIND_IDENTITY, fold_batchnorm,
                                          unet_conv_specs)


def _bf(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _conv(f, x, dims):
    sp = f.spec
    w = _bf(f.weight)
    if sp.kind == KIND_CONV_TRANSPOSE:
        fn = F.conv_transpose3d if dims == 3 else F.conv_transpose2d
        y = fn(x, w, f.bias, stride=sp.stride, padding=1, output_padding=sp.stride - 1)
    else:
        fn = F.conv3d if dims == 3 else F.conv2d
        y = fn(x, w, f.bias, stride=sp.stride, padding=sp.kernel // 2)
    if sp.has_adn:
        y = torch.where(y > 0, y, y * f.alpha)
    return y


def stem_rounds_input(onet) -> bool:
    """Whether the device path's tensor-core stem (which rounds the network input to bf16) handles this network."""
    return (onet.dimensions == 3 and onet.strides[0] == 2 and 27 * onet.in_channels <= 64
            and onet.channels[0] <= 16)


def bf16_forward(onet, state_dict, x: torch.Tensor) -> torch.Tensor:
    """Forward of the folded network with bf16 rounding of weights and stored activations."""
    dims = onet.dimensions
    specs = unet_conv_specs(onet.in_channels, onet.out_channels, onet.channels, onet.strides)
    folded = fold_batchnorm(state_dict, specs)
    n = len(onet.channels) - 1
    it = iter(folded)
    cur = x.to(torch.float32)
    if stem_rounds_input(onet):
        cur = _bf(cur)
    skips = []
    for i in range(n + 1):  # n down levels + bottom
        u0, u1, rs = next(it), next(it), next(it)
        t = _bf(_conv(u0, cur, dims))
        r = cur if rs.spec.kind == KIND_IDENTITY else _bf(_conv(rs, cur, dims))
        cur = _bf(_conv(u1, t, dims) + r)
        if i < n:
            skips.append(cur)
    sub = cur
    for i in range(n - 1, -1, -1):
        ct, ru = next(it), next(it)
        u = _bf(_conv(ct, torch.cat([skips[i], sub], 1), dims))
        o = _conv(ru, u, dims) + u
        sub = o if i == 0 else _bf(o)
    return sub
