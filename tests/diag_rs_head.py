"""Diagnostic: the row-sweep head kernel (conv_rs.cu) vs the plane-sweep kernel (conv_ps.cu, SGM_NO_RS=1) on
identical inputs -- whole-network forwards (planar logits) and sliding-window predictions (importance-weighted logits
for the deferred blend).  Only the head differs between the two networks, so the outputs agree to fp32 rounding of
the accumulation order.  `python tests/diag_rs_head.py [time]`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmantic_b200.seg import engine  # noqa: E402
from segmantic_b200.synthetic import synthetic_state_dict, synthetic_volume  # noqa: E402


def make_nets(cout, dev):
    sd = synthetic_state_dict(3, 1, cout, seed=0)
    os.environ.pop("SGM_NO_RS", None)
    rs = engine.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=cout, device=dev, precision="bf16")
    os.environ["SGM_NO_RS"] = "1"
    ps = engine.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=cout, device=dev, precision="bf16")
    os.environ.pop("SGM_NO_RS", None)
    return rs, ps


def compare(name, a, b, tol=2e-5):
    scale = float(b.abs().max())
    err = float((a - b).abs().max())
    nbad = int(((a - b).abs() > tol * scale).sum())
    ok = err <= tol * scale and bool(torch.isfinite(a).all())
    msg = f"{'PASS' if ok else 'FAIL'} {name:46s} max|ref|={scale:.4g} maxerr={err:.3g} bad={nbad}/{a.numel()}"
    if not ok and nbad:
        bad = ((a - b).abs() > tol * scale).nonzero()
        msg += f" first={bad[0].tolist()} ref={float(b[tuple(bad[0])]):.5g} got={float(a[tuple(bad[0])]):.5g} last={bad[-1].tolist()}"
    print(msg, flush=True)
    return ok


def run(timing=False):
    dev = torch.device("cuda:0")
    results = []
    for cout in (10, 3):
        rs, ps = make_nets(cout, dev)
        for shape, n in (((96, 96, 96), 2), ((96, 96, 96), 40), ((48, 64, 96), 3), ((32, 48, 64), 1), ((64, 32, 112), 2)):
            g = torch.Generator().manual_seed(sum(shape) + n)
            x = torch.randn((n, 1) + shape, generator=g).to(dev)
            try:
                a = rs(x)
                rs.check()
                b = ps(x)
                ps.check()
                results.append(compare(f"forward C={cout} {shape} n={n}", a, b))
            except Exception as e:  # noqa: BLE001
                print(f"ERROR forward C={cout} {shape} n={n}: {e}", flush=True)
                results.append(False)
            del x
        vol = synthetic_volume((130, 100, 144), seed=3)
        vol = ((vol - vol.mean()) / vol.std(unbiased=False))[None].to(dev)
        for roi in ((96, 96, 96), (64, 48, 96)):
            try:
                ra = engine.sliding_window_inference(vol, roi, 4, rs, overlap=0.5, mode="gaussian", return_labels=True)
                rs.check()
                rb = engine.sliding_window_inference(vol, roi, 4, ps, overlap=0.5, mode="gaussian", return_labels=True)
                ps.check()
                results.append(compare(f"sliding window C={cout} roi={roi} logits", ra["logits"], rb["logits"]))
                agree = float((ra["labels"] == rb["labels"]).float().mean())
                print(f"     label agreement {agree:.6f}", flush=True)
                results.append(agree > 0.999)
            except Exception as e:  # noqa: BLE001
                print(f"ERROR sliding window C={cout} roi={roi}: {e}", flush=True)
                results.append(False)
    print(f"SUMMARY {sum(results)}/{len(results)} passed", flush=True)
    if timing:
        rs, ps = make_nets(10, dev)
        x = torch.randn((32, 1, 96, 96, 96), device=dev)
        for name, net in (("row-sweep", rs), ("plane-sweep", ps)):
            for _ in range(2):
                net(x)
            net.set_profiling(True)
            net.get_profile()
            for _ in range(3):
                net(x)
            prof = net.get_profile()
            net.set_profiling(False)
            head = [p for p in prof if p[2]][-1]
            print(f"TIMING {name}: head {head[0]} {head[1] / head[2] * 1e3 / 32:.2f} us/window ({head[2]} launches)", flush=True)
    return results


if __name__ == "__main__":
    run(len(sys.argv) > 1 and sys.argv[1] == "time")
