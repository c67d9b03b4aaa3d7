"""Shared helpers for the parity tests (oracle vs CUDA path)."""
import torch

from oracle.unet import UNet, load_checkpoint_into
from segmantic_b200.synthetic import synthetic_state_dict, synthetic_volume


def make_oracle_net(spatial_dims, cin, cout, channels=(16, 32, 64, 128, 256), strides=(2, 2, 2, 2), seed=0):
    sd = synthetic_state_dict(spatial_dims, cin, cout, channels, strides, seed)
    net = UNet(spatial_dims, cin, cout, channels, strides)
    load_checkpoint_into(net, sd)
    net.eval()
    return net, sd


def normalized_volume(shape, seed=0, channels=1):
    vol = synthetic_volume(shape, seed, channels)
    out = torch.empty_like(vol)
    for c in range(channels):
        out[c] = (vol[c] - vol[c].mean()) / vol[c].std(unbiased=False)
    return out


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| (scale-relative max error)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def label_mismatch_outside_ties(logits_ref: torch.Tensor, labels_ref: torch.Tensor, labels: torch.Tensor,
                                tol: float):
    """Count label mismatches at voxels whose reference top-2 logit gap exceeds `tol` (non-ties)."""
    top2 = logits_ref.topk(2, dim=0).values
    gap = top2[0] - top2[1]
    bad = (labels_ref != labels)
    return int((bad & (gap > tol)).sum()), int(bad.sum())


def dice_per_class(a: torch.Tensor, b: torch.Tensor, n: int):
    out = []
    for c in range(n):
        pa, pb = (a == c), (b == c)
        den = int(pa.sum()) + int(pb.sum())
        out.append(1.0 if den == 0 else 2.0 * int((pa & pb).sum()) / den)
    return out
