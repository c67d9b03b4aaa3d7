"""Layer-by-layer diagnostic: tcgen05 conv kernels vs the CUDA-core kernels on identical bf16 CG8
inputs (run on the GPU box: `python tests/diag_tc_layers.py`).  Prints one line per case."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmantic_b200.seg import engine  # noqa: E402
from segmantic_b200.synthetic import synthetic_state_dict  # noqa: E402


def cg8(n, cg, dims, seed, dev, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn((n, cg) + tuple(dims) + (8,), generator=g) * scale
    return t.to(torch.bfloat16).to(dev)


def from_cg8(t, channels):
    """CG8 [n, cg, d0, d1, d2, 8] -> NCDHW float32 [n, channels, d0, d1, d2]."""
    n, cg = t.shape[:2]
    return t.float().permute(0, 1, 5, 2, 3, 4).reshape(n, cg * 8, *t.shape[2:5])[:, :channels].contiguous()


_FOLDED = {}


def torch_conv_reference(idx, in0, in1=None, res=None, seed=0, classes=10):
    """The same layer through torch's own conv3d / conv_transpose3d (CPU, fp32) on the bf16 inputs with the
    BN-folded, bf16-rounded weights: an independent statement of what the layer computes (not a repo kernel)."""
    import torch.nn.functional as F
    from segmantic_b200.seg.unet_spec import KIND_CONV_TRANSPOSE, fold_batchnorm, unet_conv_specs
    if (seed, classes) not in _FOLDED:
        specs = unet_conv_specs(1, classes, (16, 32, 64, 128, 256), (2, 2, 2, 2))
        _FOLDED[(seed, classes)] = fold_batchnorm(synthetic_state_dict(3, 1, classes, seed=seed), specs)
    f = _FOLDED[(seed, classes)][idx]
    sp = f.spec
    x = from_cg8(in0.cpu(), in0.shape[1] * 8)
    if in1 is not None:
        x = torch.cat([x, from_cg8(in1.cpu(), in1.shape[1] * 8)], 1)
    x = x[:, :sp.cin]
    w = f.weight.to(torch.bfloat16).float()
    if sp.kind == KIND_CONV_TRANSPOSE:
        y = F.conv_transpose3d(x, w, f.bias, stride=sp.stride, padding=1, output_padding=sp.stride - 1)
    else:
        y = F.conv3d(x, w, f.bias, stride=sp.stride, padding=sp.kernel // 2)
    if sp.has_adn:
        y = torch.where(y > 0, y, y * f.alpha)
    if res is not None:
        y = y + from_cg8(res.cpu(), sp.cout)
    return y


def run_case(net, name, idx, in0, in1=None, res=None, fused=False, cg_out2=0, torch_check=False, classes=10,
             real_channels=0):
    try:
        print(f"START {name}", flush=True)
        ref = engine.debug_conv(net, idx, in0, in1, res, use_tc=False)
        print("  ref done", flush=True)
        if fused:
            ref2 = engine.debug_conv(net, idx + 2, in0, None, None, use_tc=False)
            out, out2 = engine.debug_conv(net, idx, in0, in1, res, use_tc=True, fused=True, cg_out2=cg_out2)
            pairs = [("main", ref, out), ("resid", ref2, out2)]
        else:
            out = engine.debug_conv(net, idx, in0, in1, res, use_tc=True)
            pairs = [("out", ref, out)]
        ok = True
        for tag, r, o in pairs:
            r32, o32 = r.float(), o.float()
            if real_channels:  # padded channels (beyond the real class count) are don't-cares: zero weights downstream
                r32, o32 = from_cg8(r32, real_channels), from_cg8(o32, real_channels)
            scale = float(r32.abs().max())
            err = float((r32 - o32).abs().max())
            nbad = int(((r32 - o32).abs() > 2e-2 * max(scale, 1e-6)).sum())
            good = err <= 2e-2 * max(scale, 1e-6)
            ok &= good
            msg = f"{name:34s} {tag:5s} shape={tuple(o.shape)} max|ref|={scale:.4g} maxerr={err:.4g} bad={nbad}/{o.numel()}"
            if not good:
                bad = ((r32 - o32).abs() > 2e-2 * max(scale, 1e-6)).nonzero()
                msg += f" firstbad={bad[0].tolist()} ref={float(r32[tuple(bad[0])]):.4g} got={float(o32[tuple(bad[0])]):.4g}"
                msg += f" lastbad={bad[-1].tolist()} nz_out={int((o32 != 0).sum())}"
            print(("PASS " if good else "FAIL ") + msg, flush=True)
        if torch_check:  # against torch's conv on the same bf16 inputs: <= one bf16 ulp of the output range
            refs = [torch_conv_reference(idx, in0, in1, res, classes=classes)]
            outs = [out]
            if fused:
                refs.append(torch_conv_reference(idx + 2, in0))
                outs.append(out2)
            for r, o in zip(refs, outs):
                o32 = from_cg8(o.cpu(), r.shape[1])
                scale = float(r.abs().max())
                err = float((r - o32).abs().max())
                good = err <= 1e-2 * max(scale, 1e-6)
                ok &= good
                print(("PASS " if good else "FAIL ") + f"{name:34s} torch max|ref|={scale:.4g} maxerr={err:.4g}", flush=True)
        return ok
    except Exception as e:  # noqa: BLE001
        print(f"ERROR {name}: {type(e).__name__}: {e}", flush=True)
        return False


def run(which="all"):
    """Runs one group of cases ("s1", "ps", "s2", "t2", "k1", "head" or "all"); returns the list of pass flags."""
    dev = torch.device("cuda:0")
    sd = synthetic_state_dict(3, 1, 10, seed=0)
    net = engine.UNetB200(sd, spatial_dims=3, in_channels=1, out_channels=10, device=dev, precision="bf16")
    results = []
    if which in ("k1",):
        results.append(run_case(net, "k1 128->256 bottom.residual 6^3", 14, cg8(1, 16, (6, 6, 6), 9, dev)))
    if which in ("head",):
        results.append(run_case(net, "s1 10->10 head conv 32^3", 22, cg8(1, 2, (32, 32, 32), 10, dev)))
    if which in ("all", "s1"):
        # stride-1 3x3x3 convs (indices: d0.unit1=1, d1.unit1=4, d2.unit1=7, d3.unit1=10, bottom 12/13, bottom k1 res=14,
        # up ru: 16, 18, 20, head 22)
        results.append(run_case(net, "s1 16->16 d0.unit1 (tiny 8^3)", 1, cg8(1, 2, (8, 8, 8), 1, dev)))
        results.append(run_case(net, "s1 16->16 d0.unit1 24x20x40 n2 +res", 1, cg8(2, 2, (24, 20, 40), 2, dev),
                                res=cg8(2, 2, (24, 20, 40), 3, dev)))
        results.append(run_case(net, "s1 32->32 d1.unit1 24^3", 4, cg8(1, 4, (24, 24, 24), 4, dev)))
        results.append(run_case(net, "s1 64->64 d2.unit1 12^3", 7, cg8(2, 8, (12, 12, 12), 5, dev)))
        results.append(run_case(net, "s1 128->128 d3.unit1 6^3", 10, cg8(2, 16, (6, 6, 6), 6, dev)))
        results.append(run_case(net, "s1 128->256 bottom.unit0 6^3", 12, cg8(1, 16, (6, 6, 6), 7, dev)))
        results.append(run_case(net, "s1 256->256 bottom.unit1 6^3", 13, cg8(1, 32, (6, 6, 6), 8, dev)))
        results.append(run_case(net, "k1 128->256 bottom.residual 6^3", 14, cg8(1, 16, (6, 6, 6), 9, dev)))
        results.append(run_case(net, "s1 10->10 head conv 32^3", 22, cg8(1, 2, (32, 32, 32), 10, dev)))
    if which in ("all", "ps"):
        # plane-sweep family (conv_ps.cu): identity residual from the brick centre, global residual, ragged extents
        x = cg8(1, 2, (48, 48, 48), 30, dev)
        results.append(run_case(net, "ps 16->16 up1.ru 48^3 identity res", 20, x, res=x))
        x = cg8(2, 4, (24, 24, 24), 31, dev)
        results.append(run_case(net, "ps 32->32 up2.ru 24^3 n2 identity res", 18, x, res=x))
        results.append(run_case(net, "ps 32->32 d1.unit1 24^3 global res", 4, cg8(1, 4, (24, 24, 24), 32, dev),
                                res=cg8(1, 4, (24, 24, 24), 33, dev)))
        x = cg8(1, 2, (13, 17, 29), 34, dev)
        results.append(run_case(net, "ps 16->16 ragged 13x17x29 identity", 20, x, res=x))
        x = cg8(3, 2, (31, 33, 50), 35, dev)
        results.append(run_case(net, "ps 10->10 head ragged 31x33x50 n3", 22, x, res=x))
        x = cg8(2, 2, (96, 96, 96), 36, dev)
        results.append(run_case(net, "ps 10->10 head 96^3 n2 identity", 22, x, res=x))
        # small extents (the 32^3-roi tests): few planes per column, few units
        results.append(run_case(net, "ps 16->16 d0.unit1 16^3 n2 global res", 1, cg8(2, 2, (16, 16, 16), 37, dev),
                                res=cg8(2, 2, (16, 16, 16), 38, dev)))
        x = cg8(2, 2, (16, 16, 16), 39, dev)
        results.append(run_case(net, "ps 16->16 up1.ru 16^3 n2 identity", 20, x, res=x))
        results.append(run_case(net, "ps 32->32 d1.unit1 8^3 n2 global res", 4, cg8(2, 4, (8, 8, 8), 40, dev),
                                res=cg8(2, 4, (8, 8, 8), 41, dev)))
        x = cg8(2, 4, (8, 8, 8), 42, dev)
        results.append(run_case(net, "ps 32->32 up2.ru 8^3 n2 identity", 18, x, res=x))
        x = cg8(2, 2, (32, 32, 32), 43, dev)
        results.append(run_case(net, "ps 10->10 head 32^3 n2 identity", 22, x, res=x))
        x = cg8(1, 2, (5, 7, 3), 44, dev)
        results.append(run_case(net, "ps 16->16 tiny 5x7x3 identity", 20, x, res=x))
        x = cg8(1, 2, (1, 9, 9), 45, dev)
        results.append(run_case(net, "ps 16->16 one plane 1x9x9 identity", 20, x, res=x))
    if which in ("all", "s2"):
        results.append(run_case(net, "s2 16->32(+32) d1.unit0 fused 16^3", 3, cg8(1, 2, (16, 16, 16), 11, dev),
                                fused=True, cg_out2=4))
        results.append(run_case(net, "s2 16->32(+32) d1 fused 48x32x24 n2", 3, cg8(2, 2, (48, 32, 24), 12, dev),
                                fused=True, cg_out2=4))
        results.append(run_case(net, "s2 32->64(+64) d2.unit0 fused 24^3", 6, cg8(1, 4, (24, 24, 24), 13, dev),
                                fused=True, cg_out2=8))
        results.append(run_case(net, "s2 64->128(+128) d3 fused 12^3", 9, cg8(1, 8, (12, 12, 12), 14, dev),
                                fused=True, cg_out2=16))
        results.append(run_case(net, "s2 16->32 d1.unit0 alone 16^3", 3, cg8(1, 2, (16, 16, 16), 15, dev)))
    if which in ("all", "cs"):
        # channel-streamed persistent kernel (conv_cs.cu): several windows per unit (full and partial groups), units
        # spread over the CTAs, residuals, ragged extents (sub-bricks), the two inputs of the transposed convs; every
        # case also against torch's own conv3d / conv_transpose3d
        results.append(run_case(net, "cs s1 64->64 d2.unit1 12^3 n3 +res", 7, cg8(3, 8, (12, 12, 12), 50, dev),
                                res=cg8(3, 8, (12, 12, 12), 51, dev), torch_check=True))
        results.append(run_case(net, "cs s1 128->128 d3.unit1 6^3 n5 +res", 10, cg8(5, 16, (6, 6, 6), 52, dev),
                                res=cg8(5, 16, (6, 6, 6), 53, dev), torch_check=True))
        results.append(run_case(net, "cs s1 128->256 bottom.unit0 6^3 n4", 12, cg8(4, 16, (6, 6, 6), 54, dev), torch_check=True))
        results.append(run_case(net, "cs s1 256->256 bottom.unit1 6^3 n7", 13, cg8(7, 32, (6, 6, 6), 55, dev),
                                res=cg8(7, 32, (6, 6, 6), 56, dev), torch_check=True))
        results.append(run_case(net, "cs s1 256->256 bottom.unit1 6^3 n150", 13, cg8(150, 32, (6, 6, 6), 57, dev, 0.5),
                                res=cg8(150, 32, (6, 6, 6), 58, dev)))
        x = cg8(2, 8, (12, 12, 12), 59, dev)
        results.append(run_case(net, "cs s1 64->64 up3.ru 12^3 n2 identity", 16, x, res=x, torch_check=True))
        results.append(run_case(net, "cs s1 64->64 ragged 7x9x11 n2", 7, cg8(2, 8, (7, 9, 11), 60, dev), torch_check=True))
        results.append(run_case(net, "cs s1 128->128 ragged 3x4x5 n9", 10, cg8(9, 16, (3, 4, 5), 61, dev), torch_check=True))
        results.append(run_case(net, "cs s2 32->64(+64) d2 fused 24^3 n3", 6, cg8(3, 4, (24, 24, 24), 62, dev),
                                fused=True, cg_out2=8, torch_check=True))
        results.append(run_case(net, "cs s2 64->128(+128) d3 fused 12^3 n5", 9, cg8(5, 8, (12, 12, 12), 63, dev),
                                fused=True, cg_out2=16, torch_check=True))
        results.append(run_case(net, "cs s2 64->128(+128) ragged 10x6x14 n2", 9, cg8(2, 8, (10, 6, 14), 64, dev),
                                fused=True, cg_out2=16, torch_check=True))
        results.append(run_case(net, "cs t2 384->64 up3 6^3 n5", 15, cg8(5, 16, (6, 6, 6), 65, dev),
                                cg8(5, 32, (6, 6, 6), 66, dev), torch_check=True))
        results.append(run_case(net, "cs t2 128->32 up2 12^3 n3", 17, cg8(3, 8, (12, 12, 12), 67, dev),
                                cg8(3, 8, (12, 12, 12), 68, dev), torch_check=True))
        results.append(run_case(net, "cs t2 128->32 ragged 5x7x9 n2", 17, cg8(2, 8, (5, 7, 9), 69, dev),
                                cg8(2, 8, (5, 7, 9), 70, dev), torch_check=True))
    if which in ("all", "torch"):
        # one direct torch comparison per remaining tcgen05 family (plane sweep, transposed plane sweep, brick kernel)
        x = cg8(2, 2, (48, 48, 48), 80, dev)
        results.append(run_case(net, "ps 16->16 up1.ru 48^3 n2 identity", 20, x, res=x, torch_check=True))
        results.append(run_case(net, "ps 32->32 d1.unit1 24^3 global res", 4, cg8(1, 4, (24, 24, 24), 81, dev),
                                res=cg8(1, 4, (24, 24, 24), 82, dev), torch_check=True))
        results.append(run_case(net, "pst 64->16 up1 24^3 n2", 19, cg8(2, 4, (24, 24, 24), 83, dev),
                                cg8(2, 4, (24, 24, 24), 84, dev), torch_check=True))
        results.append(run_case(net, "pst 32->10 up0 48x32x40", 21, cg8(1, 2, (48, 32, 40), 85, dev),
                                cg8(1, 2, (48, 32, 40), 86, dev), torch_check=True))
        results.append(run_case(net, "brick s2 16->32(+32) d1 fused 48^3", 3, cg8(1, 2, (48, 48, 48), 87, dev),
                                fused=True, cg_out2=4, torch_check=True))
        results.append(run_case(net, "brick k1 128->256 bottom.residual 6^3", 14, cg8(2, 16, (6, 6, 6), 88, dev), torch_check=True))
    if which in ("all", "c20"):
        # 20 tissues (BASELINE configs[3]): the transposed plane sweep in two launches of 16 output channels, the head
        # on the plane-sweep kernel (11..32 classes)
        sd20 = synthetic_state_dict(3, 1, 20, seed=0)
        net20 = engine.UNetB200(sd20, spatial_dims=3, in_channels=1, out_channels=20, device=dev, precision="bf16")
        results.append(run_case(net20, "pst 32->20 up0 24x32x40 n2 (two passes)", 21, cg8(2, 2, (24, 32, 40), 90, dev),
                                cg8(2, 2, (24, 32, 40), 91, dev), torch_check=True, classes=20))
        x = cg8(2, 4, (32, 48, 40), 92, dev)
        results.append(run_case(net20, "ps 20->20 head 32x48x40 n2 identity", 22, x, res=x, torch_check=True, classes=20,
                                real_channels=20))
    if which in ("all", "t2"):
        results.append(run_case(net, "t2 384->64 up3 6^3", 15, cg8(1, 16, (6, 6, 6), 16, dev), cg8(1, 32, (6, 6, 6), 17, dev)))
        results.append(run_case(net, "t2 128->32 up2 12^3", 17, cg8(1, 8, (12, 12, 12), 18, dev), cg8(1, 8, (12, 12, 12), 19, dev)))
        results.append(run_case(net, "t2 64->16 up1 24x16x20 n2", 19, cg8(2, 4, (24, 16, 20), 20, dev), cg8(2, 4, (24, 16, 20), 21, dev)))
        results.append(run_case(net, "t2 32->10 up0 16^3", 21, cg8(1, 2, (16, 16, 16), 22, dev), cg8(1, 2, (16, 16, 16), 23, dev)))
    print(f"SUMMARY {sum(results)}/{len(results)} passed", flush=True)
    return results


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "all")
