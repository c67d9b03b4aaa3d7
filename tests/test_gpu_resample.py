"""GPU parity for the bandwidth-bound stages: MONAI-Spacing trilinear resample (float64 arithmetic,
tolerance 1e-6 relative after the cast to fp32), ITK resample (bit-exact), normalise, bbox."""
import numpy as np
import pytest
import torch

from oracle import itk_resample as oitk
from oracle import spacing as osp

pytestmark = pytest.mark.gpu


def test_spacing_forward_and_inverse(cuda_device):
    from segmantic_b200.seg import transforms as T
    g = torch.Generator().manual_seed(0)
    img = torch.randn((2, 40, 36, 17), generator=g)
    aff = osp.itk_geometry_to_ras_affine((0.5, 0.6, 3.0), (-10.0, 5.0, 2.0), np.eye(3).flatten())
    img_o, aff_o, rec_o = osp.orientation_ras(img, aff)
    ref, ref_aff, ref_rec = osp.spacing_forward(img_o, aff_o, (1.0, 1.0, 1.0))
    img_d, aff_d, rec_d = T.orientation_ras(img.to(cuda_device), aff)
    assert np.allclose(aff_o, aff_d)
    out, out_aff, rec = T.spacing_forward(img_d, aff_d, (1.0, 1.0, 1.0))
    assert tuple(out.shape) == tuple(ref.shape)
    assert np.allclose(out_aff, ref_aff)
    assert float((out.cpu() - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
    # inverse (Invertd): back to the pre-spacing grid
    back_ref = osp.spacing_inverse(ref, ref_rec)
    back = T.resample_index_affine(out, T.spacing_inverse_xform(rec), rec["src_shape"])
    assert float((back.cpu() - back_ref).abs().max()) <= 1e-6 * float(back_ref.abs().max())
    # fused inverse + argmax == argmax of the inverse
    lab = T.resample_index_affine_argmax(out, T.spacing_inverse_xform(rec), rec["src_shape"])
    assert torch.equal(lab.cpu().long(), back.cpu().argmax(0))
    # un-orient
    assert torch.equal(T.orientation_inverse(img_d, rec_d).cpu(), img)


def test_config3_shapes(cuda_device):
    from segmantic_b200.seg import transforms as T
    aff = osp.itk_geometry_to_ras_affine((0.5, 0.5, 3.0), (-128.0, -128.0, 0.0), np.eye(3).flatten())
    img = torch.zeros((1, 64, 64, 120), device=cuda_device)
    img_d, aff_d, _ = T.orientation_ras(img, aff)
    out, _, _ = T.spacing_forward(img_d, aff_d, (1.0, 1.0, 1.0))
    assert tuple(out.shape[1:]) == (32, 32, 358)  # round((n-1)*s/t + 1), half-to-even


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16, np.float32])
@pytest.mark.parametrize("nearest", [True, False])
def test_itk_resample_bit_exact(cuda_device, dtype, nearest):
    from segmantic_b200.image import processing as P
    rng = np.random.default_rng(1)
    arr = (rng.random((23, 17, 11)) * 200).astype(dtype)
    sp, org = (0.5, 0.6, 0.7), (1.0, -2.0, 0.5)
    th = 0.2
    direction = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]]).flatten()
    mov_o = oitk.Image(arr, sp, org, tuple(direction))
    mov = P.Image(arr, sp, org, tuple(direction))
    ref_o = oitk.Image(np.zeros((31, 20, 9), np.uint16), (0.4, 0.55, 0.9), (0.7, -2.5, 0.2))
    ref = P.Image(np.zeros((31, 20, 9), np.uint16), (0.4, 0.55, 0.9), (0.7, -2.5, 0.2))
    a = oitk.resample_to_ref(mov_o, ref_o, nearest)
    b = P.resample_to_ref(mov, ref, nearest)
    assert b.GetSize() == a.GetSize() and b.GetSpacing() == a.GetSpacing()
    assert b.array.dtype == a.array.dtype == np.dtype(dtype)
    assert np.array_equal(a.array, b.array)
    a2 = oitk.resample(mov_o, (0.25, 0.3, 0.35), nearest)
    b2 = P.resample(mov, (0.25, 0.3, 0.35), nearest)
    assert b2.GetSize() == a2.GetSize() == (46, 34, 22)
    assert np.array_equal(a2.array, b2.array)


def test_reference_labelfield_fixture(cuda_device):
    """The reference's own fixtures (tests/conftest.py:7-13, tests/image/test_image.py:33-52)."""
    from segmantic_b200.image import processing as P
    lab = P.make_image(shape=(5, 5, 5), spacing=(0.5, 0.6, 0.7))
    for i in range(5):
        lab.array[..., i] = i
    spacing = [s / 2.0 for s in lab.GetSpacing()]
    res = P.resample(lab, target_spacing=spacing)
    assert list(res.GetSize()) == [2 * s for s in lab.GetSize()]
    ref = P.make_image((12, 10, 7), spacing, pixel_type=np.uint16)
    ref.origin = (1.3, -2.1, 0.75)
    res = P.resample_to_ref(lab, ref, nearest=True)
    assert list(res.GetSize()) == list(ref.GetSize())
    assert list(res.GetSpacing()) == list(ref.GetSpacing())
    assert res.array.dtype == np.uint8
    o = oitk.resample_to_ref(oitk.Image(lab.array, lab.spacing), oitk.Image(ref.array, ref.spacing, ref.origin), True)
    assert np.array_equal(o.array, res.array)


def test_2d_itk_resample(cuda_device):
    from segmantic_b200.image import processing as P
    rng = np.random.default_rng(2)
    arr = (rng.random((19, 13)) * 100).astype(np.float32)
    a = oitk.resample(oitk.Image(arr, (1.0, 2.0)), (0.5, 0.7), False)
    b = P.resample(P.Image(arr, (1.0, 2.0)), (0.5, 0.7), False)
    assert np.array_equal(a.array, b.array)


def test_normalize_and_bbox(cuda_device):
    from segmantic_b200.seg import transforms as T
    g = torch.Generator().manual_seed(3)
    img = torch.randn((2, 30, 28, 26), generator=g) * 50 + 10
    img[:, :4] = -5.0
    img[:, :, 20:] = -5.0
    ref = osp.normalize_intensity(img)
    out = T.normalize_intensity(img.to(cuda_device)).cpu()
    assert float((out - ref).abs().max()) < 1e-5
    lo_o, hi_o = osp.foreground_bbox(ref)
    lo, hi = T.foreground_bbox(ref.to(cuda_device))
    assert (lo, hi) == (lo_o, hi_o)
    lo, hi = T.foreground_bbox(torch.full((1, 4, 4, 4), -1.0, device=cuda_device))
    assert (lo, hi) == ([0, 0, 0], [0, 0, 0])


@pytest.mark.parametrize("channels", [1, 3, 10, 37])
@pytest.mark.parametrize("case", ["up_xy_down_z", "down_xy_up_z", "shift_only", "flip_scale", "coarse"])
def test_trilinear_brick_kernel_matches_gather_kernel(cuda_device, monkeypatch, channels, case):
    """The shared-memory-staged trilinear kernel and the separable-table row kernel (axis-aligned transforms) against
    the one-thread-per-voxel gather kernel they replace: same float64 statements per voxel, so outputs and fused-argmax labels are EQUAL -- for
    up / down scaling, pure shifts, negative scales, clipped borders, several channel passes, ragged tiles."""
    from segmantic_b200.seg import transforms as T
    g = torch.Generator().manual_seed(channels)
    src, dst, diag, off = {
        "up_xy_down_z": ((33, 29, 71), (67, 58, 24), (0.5, 0.5, 3.0), (-0.25, -0.25, 1.0)),
        "down_xy_up_z": ((64, 50, 21), (31, 26, 61), (2.0, 2.0, 1.0 / 3.0), (0.5, 0.5, -1.0 / 3.0)),
        "shift_only": ((20, 22, 40), (20, 22, 40), (1.0, 1.0, 1.0), (0.3, -2.6, 5.5)),
        "flip_scale": ((24, 18, 50), (30, 20, 33), (-0.8, 0.9, -1.5), (23.5, 0.1, 49.2)),
        "coarse": ((9, 200, 130), (3, 17, 11), (4.0, 12.5, 12.0), (0.0, 3.0, -4.0)),
    }[case]
    img = torch.randn((channels,) + src, generator=g).to(cuda_device)
    xf = np.zeros((4, 4))
    xf[3, 3] = 1.0
    for a in range(3):
        xf[a, a], xf[a, 3] = diag[a], off[a]
    monkeypatch.setenv("SGM_RESAMPLE_BRICK_ALWAYS", "1")  # few channels take the table kernel by default
    monkeypatch.setenv("SGM_NO_RESAMPLE_BRICK", "1")
    ref = T.resample_index_affine(img, xf, dst)            # the one-thread-per-voxel gather kernel: the reference form
    ref_lab = T.resample_index_affine_argmax(img, xf, dst)
    monkeypatch.setenv("SGM_RESAMPLE_SEP", "1")            # separable tables + one warp per output row (opt-in)
    sep = T.resample_index_affine(img, xf, dst)
    sep_lab = T.resample_index_affine_argmax(img, xf, dst)
    monkeypatch.delenv("SGM_RESAMPLE_SEP")
    monkeypatch.delenv("SGM_NO_RESAMPLE_BRICK")            # shared-memory-staged tiles
    out = T.resample_index_affine(img, xf, dst)
    lab = T.resample_index_affine_argmax(img, xf, dst)
    assert torch.equal(sep, ref) and torch.equal(sep_lab, ref_lab)
    assert torch.equal(out, ref)
    assert torch.equal(lab, ref_lab)
    assert torch.equal(lab.long(), out.argmax(0))


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32])
@pytest.mark.parametrize("nearest", [True, False])
def test_itk_resample_vector_kernel_bit_exact(cuda_device, monkeypatch, dtype, nearest):
    """Scan lines whose length is a multiple of 4 take the 4-voxels-per-thread kernel: equal to the scalar kernel and
    to the oracle (ITK's scan-line index formula), including voxels outside the moving image."""
    from segmantic_b200.image import processing as P
    rng = np.random.default_rng(7)
    arr = (rng.random((26, 19, 13)) * 200).astype(dtype)
    sp, org = (0.5, 0.6, 0.7), (1.0, -2.0, 0.5)
    mov_o, mov = oitk.Image(arr, sp, org), P.Image(arr, sp, org)
    for size, rsp, rorg in (((32, 21, 10), (0.4, 0.55, 0.9), (0.7, -2.5, 0.2)), ((52, 8, 12), (0.25, 1.5, 0.8), (0.9, -2.2, 0.4))):
        ref_o = oitk.Image(np.zeros(size, np.uint16), rsp, rorg)
        ref = P.Image(np.zeros(size, np.uint16), rsp, rorg)
        a = oitk.resample_to_ref(mov_o, ref_o, nearest)
        b = P.resample_to_ref(mov, ref, nearest)
        monkeypatch.setenv("SGM_NO_RESAMPLE_VEC", "1")
        c = P.resample_to_ref(mov, ref, nearest)
        monkeypatch.delenv("SGM_NO_RESAMPLE_VEC")
        monkeypatch.setenv("SGM_RESAMPLE_SEP", "1")   # opt-in: separable index tables + pure gather (nearest only)
        d = P.resample_to_ref(mov, ref, nearest)
        monkeypatch.delenv("SGM_RESAMPLE_SEP")
        assert np.array_equal(b.array, c.array) and np.array_equal(b.array, d.array)
        assert np.array_equal(a.array, b.array)
