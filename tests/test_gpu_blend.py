"""GPU parity of the deferred blend (sgm_sw_blend through the C ABI) on CRAFTED weighted logits.

The streaming blend kernel decides argmax(acc / count) without dividing every class (blend.cu); these tests pin it
bit-exactly against the oracle's order of operations (`out[slice] += seg`, `out / count`, `argmax`,
oracle/sliding_window.py) on inputs full of exact ties and 1-ulp near-ties, for class counts that take each
instantiation (guarded / unguarded class loops, one / several passes)."""
import ctypes as C
import itertools

import pytest
import torch

from oracle import sliding_window as osw

pytestmark = pytest.mark.gpu


def _blend(cuda_device, wl, sched, channels, want_logits):
    from segmantic_b200 import _lib
    from segmantic_b200.seg import engine

    lib = _lib.load()
    cfg, keep = engine._make_cfg(sched, 4)
    size = tuple(sched.padded_size)
    wl_d = wl.to(cuda_device).contiguous()
    labels = torch.full(size, 255, dtype=torch.uint8, device=cuda_device)
    logits = torch.empty((channels,) + size, dtype=torch.float32, device=cuda_device) if want_logits else None
    scratch = torch.empty(4096, dtype=torch.float32, device=cuda_device)
    with torch.cuda.device(cuda_device):
        _lib.check(lib.sgm_sw_blend(C.byref(cfg), channels, wl_d.data_ptr(),
                                    logits.data_ptr() if want_logits else None, labels.data_ptr(), None,
                                    scratch.data_ptr(), engine._stream_ptr(cuda_device)), "sgm_sw_blend")
        torch.cuda.synchronize()
    del keep
    return labels.cpu(), (logits.cpu() if want_logits else None)


def _oracle_blend(wl, sched, channels, mode):
    """MONAI's accumulation order on the CPU: windows in schedule order, `out += seg`, `out / count`."""
    size, roi = tuple(sched.padded_size), tuple(sched.roi)
    imap = osw.importance_map(roi, mode)
    out = torch.zeros((channels,) + size)
    count = torch.zeros(size)
    for w, st in enumerate(itertools.product(*sched.starts)):
        sl = tuple(slice(s, s + r) for s, r in zip(st, roi))
        out[(slice(None),) + sl] += wl[w]
        count[sl] += imap
    out = out / count
    return out, out.argmax(0)


@pytest.mark.parametrize("channels", [3, 8, 10, 13, 20])
@pytest.mark.parametrize("mode", ["gaussian", "constant"])
def test_blend_labels_exact_under_ties(cuda_device, channels, mode):
    from segmantic_b200.seg.sliding_window import make_schedule

    size, roi = (40, 36, 64), (16, 16, 32)
    sched = make_schedule(size, roi, 0.5, mode, 0.125)
    n_win = sched.n_windows
    g = torch.Generator().manual_seed(channels)
    # few distinct values -> many exact ties between classes; a sprinkle of 1-ulp perturbations -> near-ties whose
    # quotients may or may not coincide after the division
    base = torch.randint(-3, 4, (n_win, channels) + roi, generator=g).float() * 0.37
    ulp = torch.randint(-1, 2, base.shape, generator=g).float() * (2.0 ** -23)
    wl = base * (1.0 + ulp)
    wl[:, :, :, :, ::7] = 0.0  # all-zero columns: every class ties, label must be 0
    ref_logits, ref_labels = _oracle_blend(wl, sched, channels, mode)
    labels, _ = _blend(cuda_device, wl, sched, channels, want_logits=False)
    lab2, logits = _blend(cuda_device, wl, sched, channels, want_logits=True)
    assert torch.equal(logits, ref_logits)                 # same fp32 ops, same order: bit-exact
    assert torch.equal(lab2.long(), ref_labels)
    assert torch.equal(labels.long(), ref_labels)          # division-free decision == argmax of the quotients
    assert int((ref_labels == 0).sum()) < ref_labels.numel()  # not a degenerate case


def test_blend_labels_exact_random_scale(cuda_device):
    """Random magnitudes over 60 orders of magnitude (incl. tiny values near the underflow guard)."""
    from segmantic_b200.seg.sliding_window import make_schedule

    size, roi, channels = (24, 24, 48), (16, 16, 16), 10
    sched = make_schedule(size, roi, 0.25, "gaussian", 0.125)
    g = torch.Generator().manual_seed(5)
    shape = (sched.n_windows, channels) + roi
    mant = torch.randn(shape, generator=g)
    expo = torch.randint(-36, 25, (sched.n_windows, 1) + roi, generator=g).float()
    wl = mant * torch.pow(torch.tensor(10.0), expo)
    wl[:, 3] = wl[:, 7]  # an exact duplicate class: ties -> lowest index
    ref_logits, ref_labels = _oracle_blend(wl, sched, channels, "gaussian")
    labels, _ = _blend(cuda_device, wl, sched, channels, want_logits=False)
    assert torch.equal(labels.long(), ref_labels)
    assert int((labels == 7).sum()) == 0
