"""CPU tests of the oracle: known-answer checks derived from the published algorithms / the
reference's own fixtures, and regression against the committed golden outputs (tests/golden)."""
import os

import numpy as np
import pytest
import torch

from oracle import itk_resample as oitk
from oracle import sliding_window as osw
from oracle import spacing as osp
from oracle.predict import predict_volume
from oracle.unet import UNet
from tests.helpers import make_oracle_net, normalized_volume

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz"))
SMALL = dict(channels=(16, 32, 48), strides=(2, 2))


def test_unet_topology_matches_monai_layout():
    """148 state_dict tensors, 4 809 920 parameters at Cin=1, C=3."""
    net = UNet(3, 1, 3)
    sd = net.state_dict()
    assert len(sd) == 148
    assert sum(p.numel() for p in net.parameters()) == 4809920
    assert sd["model.1.submodule.1.submodule.1.submodule.2.0.conv.weight"].shape == (384, 64, 3, 3, 3)
    assert sd["model.2.0.conv.weight"].shape == (32, 3, 3, 3, 3)          # ConvTranspose layout [I, O, ...]
    assert sd["model.2.1.conv.unit0.conv.weight"].shape == (3, 3, 3, 3, 3)  # top residual unit is conv only
    assert "model.2.1.conv.unit0.adn.N.weight" not in sd
    assert sd["model.1.submodule.1.submodule.1.submodule.1.submodule.residual.weight"].shape == (256, 128, 1, 1, 1)
    x = torch.zeros(1, 1, 32, 32, 32)
    assert net.eval()(x).shape == (1, 3, 32, 32, 32)


def test_reference_hparams_fixture():
    """tests/seg/test_unet.py:15-20 of the reference: Net(num_classes=3, num_channels=4, spatial_dims=2,
    spatial_size=[64]*2) keeps these hparams."""
    from segmantic_b200.seg.monai_unet import Net
    net = Net(num_classes=3, num_channels=4, spatial_dims=2, spatial_size=[64] * 2)
    assert net.hparams.num_classes == 3 and net.hparams.num_channels == 4
    assert net.hparams.spatial_dims == 2 and net.hparams.spatial_size == [64] * 2
    assert net.num_classes == 3 and net.spatial_dims == 2
    assert Net(num_classes=2).spatial_size == [96, 96, 96]


@pytest.mark.parametrize("size,overlap,expected", [
    (256, 0.5, [0, 48, 96, 144, 160]), (256, 0.25, [0, 72, 144, 160]),
    (358, 0.5, [0, 48, 96, 144, 192, 240, 262]), (96, 0.5, [0]),
])
def test_window_starts_known_answers(size, overlap, expected):
    """SURVEY.md appendix A.2 step 3."""
    starts = osw.dense_patch_starts((size,), (96,), osw.get_scan_interval((size,), (96,), overlap))[0]
    assert starts == expected


def test_window_counts_of_baseline_configs():
    assert len(osw.window_starts((256,) * 3, (96,) * 3, 0.5)) == 125
    assert len(osw.window_starts((256,) * 3, (96,) * 3, 0.25)) == 64
    assert len(osw.window_starts((256, 256, 358), (96,) * 3, 0.5)) == 175
    assert len(osw.window_starts((512, 512, 1024), (96,) * 3, 0.5)) == 2100
    w = osw.window_starts((100, 100, 200), (96,) * 3, 0.25)
    assert w[0] == (0, 0, 0) and w[1] == (0, 0, 72)  # last axis fastest


def test_gaussian_importance_map_known_values():
    """MONAI >= 1.2 separable Gaussian: 1-D edge 3.96e-4, 3-D max 0.9974, 57.97 % at the 1e-3 floor."""
    g = osw.gaussian_1d(96)
    assert abs(float(g[0]) - 3.96e-4) < 2e-6
    m = osw.importance_map((96, 96, 96), "gaussian")
    assert abs(float(m.max()) - 0.99740) < 1e-4
    assert float(m.min()) == pytest.approx(1e-3)
    assert abs(float((m == m.min()).float().mean()) - 0.5797) < 2e-3
    assert torch.equal(osw.importance_map((4, 5, 6), "constant"), torch.ones(4, 5, 6))


def test_sliding_window_identity_predictor_and_padding():
    """With predictor = identity the blended output is the input (out/count), incl. roi padding."""
    vol = torch.arange(20 * 30 * 40, dtype=torch.float32).reshape(1, 1, 20, 30, 40) / 1000.0
    for mode in ("constant", "gaussian"):
        out = osw.sliding_window_inference(vol, (32, 16, 16), 4, lambda w: w, overlap=0.5, mode=mode)
        assert out.shape == vol.shape
        assert torch.allclose(out, vol, rtol=1e-5, atol=1e-6)


def test_spacing_shape_rule_and_identity():
    """Spacingd output shape round((n-1)*s/t + 1), half to even: 512@0.5 -> 256, 120@3 -> 358."""
    aff = osp.itk_geometry_to_ras_affine((0.5, 0.5, 3.0), (-128.0, -128.0, 0.0), np.eye(3).flatten())
    img = torch.zeros(1, 512, 8, 120)
    img_o, aff_o, _ = osp.orientation_ras(img, aff)
    new_aff = osp.zoom_affine(aff_o, (1.0, 1.0, 1.0))
    shape, offset = osp.compute_shape_offset(img_o.shape[1:], aff_o, new_aff)
    assert tuple(shape) == (256, 4, 358)
    # same spacing -> untouched
    out, aff2, rec = osp.spacing_forward(torch.rand(1, 6, 7, 8), np.diag([1.0, 1.0, 1.0, 1.0]), (1.0, 1.0, 1.0))
    assert rec is None and out.shape == (1, 6, 7, 8)


def test_spacing_samples_voxel_centres():
    """2:1 down-sampling with align_corners=False index mapping: out[j] = in[2j] (centres on centres)."""
    img = torch.arange(9, dtype=torch.float32).reshape(1, 9, 1, 1).repeat(1, 1, 3, 3)
    out, _, rec = osp.spacing_forward(img, np.diag([1.0, 1.0, 1.0, 1.0]), (2.0, 1.0, 1.0))
    assert out.shape == (1, 5, 3, 3)
    assert torch.allclose(out[0, :, 1, 1], torch.tensor([0.0, 2.0, 4.0, 6.0, 8.0]))
    back = osp.spacing_inverse(out, rec)
    assert torch.allclose(back[0, :, 1, 1], torch.arange(9, dtype=torch.float32))


def test_orientation_identity_direction_flips_xy():
    img = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(1, 2, 3, 4)
    aff = osp.itk_geometry_to_ras_affine((1.0, 1.0, 1.0), (0.0, 0.0, 0.0), np.eye(3).flatten())
    out, aff2, rec = osp.orientation_ras(img, aff)
    assert torch.equal(out, torch.flip(img, dims=[1, 2]))
    assert np.all(np.diag(aff2)[:3] > 0)
    assert torch.equal(osp.orientation_inverse(out, rec), img)


def test_itk_reference_fixture_sizes_and_types():
    """The reference's own assertions: tests/image/test_image.py:33-52 (sizes, spacing, pixel type)."""
    lab = np.zeros((5, 5, 5), np.uint8)
    for k in range(5):
        lab[:, :, k] = k
    img = oitk.Image(lab, (0.5, 0.6, 0.7))
    res = oitk.resample(img, [s / 2.0 for s in img.GetSpacing()])
    assert list(res.GetSize()) == [10, 10, 10]
    ref = oitk.Image(np.zeros((12, 10, 7), np.uint16), (0.25, 0.3, 0.35), (1.3, -2.1, 0.75))
    out = oitk.resample_to_ref(img, ref, nearest=True)
    assert list(out.GetSize()) == [12, 10, 7] and out.GetSpacing() == ref.GetSpacing()
    assert out.array.dtype == np.uint8  # the MOVING image's pixel type (processing.py:92)


def test_itk_nearest_semantics():
    """Index = floor(c + 0.5) (round half up), inside = [-0.5, n - 0.5), outside -> 0; exact 2:1 grid."""
    lab = np.arange(1, 5, dtype=np.uint8).reshape(4, 1, 1).repeat(2, 1).repeat(2, 2)
    img = oitk.Image(lab, (1.0, 1.0, 1.0))
    out = oitk.resample_onto_grid(img, (10, 2, 2), (0.5, 1.0, 1.0), (-0.5, 0.0, 0.0), np.eye(3).flatten(), True)
    # c = -0.5, 0, 0.5, 1, 1.5, ...  -> ties round up; c = 3.5 is outside [-0.5, 3.5)
    assert out.array[:, 0, 0].tolist() == [1, 1, 2, 2, 3, 3, 4, 4, 0, 0]


def test_itk_linear_truncates_into_integer_types():
    lab = np.array([0, 10], dtype=np.uint8).reshape(2, 1, 1)
    img = oitk.Image(lab, (1.0, 1.0, 1.0))
    out = oitk.resample_onto_grid(img, (5, 1, 1), (0.25, 1.0, 1.0), (0.0, 0.0, 0.0), np.eye(3).flatten(), False)
    assert out.array[:, 0, 0].tolist() == [0, 2, 5, 7, 10]   # 2.5 -> 2, 7.5 -> 7 (static_cast truncation)


def test_golden_unet_and_sliding_window():
    net, _ = make_oracle_net(3, 2, 4, seed=11, **SMALL)
    x = normalized_volume((16, 24, 32), seed=21, channels=2)[None]
    with torch.no_grad():
        y = net(x)
    assert np.allclose(y.numpy(), GOLD["unet_forward"], rtol=1e-4, atol=1e-5)
    net1, _ = make_oracle_net(3, 1, 3, seed=12, **SMALL)
    vol = normalized_volume((32, 24, 28), seed=22)[None]
    with torch.no_grad():
        sw = osw.sliding_window_inference(vol, (16, 16, 16), 4, net1, overlap=0.5, mode="gaussian")
    assert np.allclose(sw.numpy(), GOLD["sw_gauss"], rtol=1e-4, atol=1e-5)


def test_golden_predict_and_itk():
    net1, _ = make_oracle_net(3, 1, 3, seed=12, **SMALL)
    raw = normalized_volume((48, 40, 12), seed=23) * 100.0 + 50.0
    aff = osp.itk_geometry_to_ras_affine((0.5, 0.5, 3.0), (-12.0, -10.0, 0.0), np.eye(3).flatten())
    for mode in ("logits", "labels"):
        lab, _ = predict_volume(net1, raw, aff, (1.0, 1.0, 1.0), roi=(16, 16, 16), invert=mode)
        gold = GOLD[f"predict_{mode}_mode"]
        assert lab.shape == (48, 40, 12)
        assert float((lab.numpy() != gold).mean()) < 1e-3  # fp32 conv summation order may flip near-ties
    lab = np.zeros((5, 5, 5), np.uint8)
    for k in range(5):
        lab[:, :, k] = k
    img = oitk.Image(lab, (0.5, 0.6, 0.7))
    assert np.array_equal(oitk.resample(img, (0.25, 0.3, 0.35), True).array, GOLD["itk_near"])
    assert np.array_equal(oitk.resample(img, (0.25, 0.3, 0.35), False).array, GOLD["itk_lin"])
    ref = oitk.Image(np.zeros((12, 10, 7), np.uint16), (0.25, 0.3, 0.35), (1.3, -2.1, 0.75))
    assert np.array_equal(oitk.resample_to_ref(img, ref, True).array, GOLD["itk_ref"])


def test_evaluation_oracle_known_answers():
    """confusion matrix / Dice / MONAI confusion-matrix metrics on a hand-checked example (8 voxels, 3 classes)."""
    from oracle import evaluation as oe
    y = np.array([0, 0, 1, 1, 2, 2, 2, 1])
    p = np.array([0, 1, 1, 1, 2, 0, 2, 2])
    cm = oe.confusion_matrix(3, p, y)
    assert np.array_equal(cm, [[1, 1, 0], [0, 2, 1], [1, 0, 2]])           # rows: truth, columns: prediction
    assert np.allclose(oe.class_dice(cm), [2 * 2 / (3 + 3), 2 * 2 / (3 + 3)])
    counts = oe.confusion_counts(cm)                                         # (tp, fp, tn, fn) per class
    assert np.array_equal(counts, [[1, 1, 5, 1], [2, 1, 4, 1], [2, 1, 4, 1]])
    m = oe.confusion_metrics([counts])
    assert m == {"sensitivity": 5 / 8, "specificity": 13 / 16, "precision": 5 / 8, "accuracy": 18 / 24}
    assert np.array_equal(oe.confusion_matrix(2, np.array([0, 1, 7]), np.array([1, 1, 0])), [[0, 0], [1, 1]])


def test_ensemble_oracle_known_answers():
    """MeanEnsemble(weights) / VoteEnsemble / SelectBestEnsemble on hand-checked inputs."""
    from oracle import ensemble as oens
    labs = torch.tensor([[0, 1, 2, 2, 1], [0, 2, 2, 1, 0], [1, 2, 0, 1, 3]])
    assert oens.vote_ensemble(labs, 3).tolist() == [0, 2, 2, 1, 0]          # 1-1-1 tie / out-of-range label -> lowest class
    assert oens.select_best_ensemble(labs, {1: 0, 2: 2}).tolist() == [0, 2, 0, 0, 1]  # later pairs overwrite, unclaimed -> 0
    x = torch.tensor([[[1.0], [2.0]], [[3.0], [0.0]]])                        # [E=2, C=2, V=1]
    assert torch.allclose(oens.mean_ensemble(x, [1.0, 3.0]), torch.tensor([[2.5], [0.5]]))  # w / mean(w) = 0.5, 1.5
    assert torch.allclose(oens.mean_ensemble(x), torch.tensor([[2.0], [1.0]]))


def test_reference_fixtures_for_evaluation_and_ensemble():
    """The reference's OWN golden vectors for these rows, replayed on the oracle:
    tests/seg/test_evaluation.py:8-22 (confusion matrix of a 10x10 label field against itself: diagonal == bincount, zero
    off-diagonals) and tests/seg/test_transforms.py:8-27 (SelectBestEnsembled on three label predictions)."""
    from oracle import ensemble as oens
    from oracle import evaluation as oe
    field = np.zeros((10, 10), np.uint8)       # make_image(shape=(10, 10), value=0); sitk index [x, y]
    field[2:3, 2:4] = 1
    field[3:5, 3:4] = 2
    view = field.T.flatten()                   # GetArrayViewFromImage: numpy order [y, x]
    assert int(view.max()) + 1 == 3
    cm = oe.confusion_matrix(3, view, view)
    assert np.all(np.diagonal(cm) == np.bincount(view)) and list(np.bincount(view)) == [96, 2, 2]
    assert np.all(np.diagonal(cm, offset=1) == 0) and np.all(np.diagonal(cm, offset=-1) == 0)
    preds = torch.stack([torch.ones(3, dtype=torch.long), torch.tensor([2, 0, 2]), torch.tensor([2, 1, 0])])
    out = oens.select_best_ensemble(preds, {1: 0, 2: 1, 0: 2})
    assert out.tolist() == [2, 1, 0]


def test_oracle_evaluation_branch_crops_by_label_and_scores_on_the_network_grid():
    """default_preprocessing(keys=["image", "label"]) (monai_unet.py:151-176): CropForegroundd(source_key="label"), the
    label through the same transforms, scoring on the pre-processed grid (:672-680)."""
    onet, _ = make_oracle_net(3, 1, 3, seed=12, **SMALL)
    raw = (normalized_volume((36, 32, 24), seed=50) * 40.0 + 10.0)[0]
    aff = osp.itk_geometry_to_ras_affine((1.0, 1.0, 1.0), (0.0, 0.0, 0.0), np.eye(3).flatten())
    gt = torch.zeros(36, 32, 24)
    gt[4:30, 3:27, 2:21] = 1
    gt[10:20, 10:20, 5:15] = 2
    lab, logits, pred_net, label_net = predict_volume(onet, raw[None], aff, (), roi=(16, 16, 16), label=gt)
    assert tuple(lab.shape) == (36, 32, 24) and tuple(pred_net.shape) == tuple(label_net.shape) == (26, 24, 19)
    assert int((label_net == 2).sum()) == 1000 and int((label_net > 0).sum()) == 26 * 24 * 19
    outside = torch.ones_like(lab, dtype=torch.bool)
    outside[4:30, 3:27, 2:21] = False
    assert not lab[outside].any()                       # inverse crop: zero outside the label's bounding box
    # with Spacing the label is resampled bilinearly and truncated (.long()), as the reference does
    lab2, _, pred2, label2 = predict_volume(onet, raw[None], aff, (2.0, 2.0, 2.0), roi=(16, 16, 16), label=gt)
    assert tuple(pred2.shape) == tuple(label2.shape) == (14, 12, 10) and tuple(lab2.shape) == (36, 32, 24)
    assert set(label2.unique().tolist()) <= {0, 1, 2}


def test_oracle_tie_gap_follows_the_labels():
    """return_gap: the top-2 logit gap behind every output label, through the same inverse transforms (+inf outside the
    crop) -- what the GPU tests use to tell argmax near-ties from real disagreements."""
    onet, _ = make_oracle_net(3, 1, 3, seed=12, **SMALL)
    raw = normalized_volume((40, 36, 28), seed=31) * 30.0 + 5.0
    lab, logits, gap = predict_volume(onet, raw, None, (), roi=(16, 16, 16), overlap=0.5, mode="gaussian", return_gap=True)
    assert gap.shape == lab.shape and bool((gap >= 0).all())
    inside = torch.isfinite(gap)
    assert int(inside.sum()) == logits.shape[1] * logits.shape[2] * logits.shape[3]   # the cropped grid, nothing else
    assert not lab[~inside].any()


# ---------------------------------------------------------------------------------------------------------------
# Independent cross-checks (scipy: a third implementation that shares no code with torch or with this repo).  They do
# not replace MONAI / ITK golden vectors -- the oracle stays "parity unpinned" against those libraries -- but they pin
# the interior arithmetic of the restatements to an outside party.
def test_itk_linear_interior_matches_scipy_map_coordinates():
    """Linear interpolation at ITK's continuous indices == scipy.ndimage.map_coordinates(order=1) wherever all eight
    neighbours exist (the borders are ITK-specific: clamped neighbours, [-0.5, n - 0.5) inside test)."""
    from scipy import ndimage
    rng = np.random.default_rng(5)
    arr = rng.random((14, 11, 9)).astype(np.float32)
    mov = oitk.Image(arr, (0.5, 0.6, 0.7), (1.0, -2.0, 0.5))
    size, sp, org = (20, 13, 8), (0.31, 0.47, 0.73), (1.2, -1.8, 0.7)
    out = oitk.resample_onto_grid(mov, size, sp, org, tuple(np.eye(3).flatten()), nearest=False).array
    grids = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in size], indexing="ij")
    coords = [(g * sp[a] + org[a] - mov.origin[a]) / mov.spacing[a] for a, g in enumerate(grids)]
    ref = ndimage.map_coordinates(arr.astype(np.float64), coords, order=1, mode="nearest")
    interior = np.ones(size, dtype=bool)
    for a in range(3):
        interior &= (coords[a] >= 0.0) & (coords[a] <= arr.shape[a] - 1.0)
    assert interior.mean() > 0.5
    assert np.allclose(out[interior], ref[interior], rtol=0, atol=2e-6)


def test_itk_nearest_matches_scipy_away_from_half_way_points():
    """Nearest neighbour: floor(c + 0.5) (ITK rounds half UP) == scipy's order-0 lookup except exactly half-way between
    two voxels, where the rounding rules may differ."""
    from scipy import ndimage
    rng = np.random.default_rng(6)
    arr = (rng.random((12, 10, 7)) * 200).astype(np.uint8)
    mov = oitk.Image(arr, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0))
    size, sp, org = (17, 14, 9), (0.7, 0.65, 0.8), (0.05, 0.1, 0.15)
    out = oitk.resample_onto_grid(mov, size, sp, org, tuple(np.eye(3).flatten()), nearest=True).array
    grids = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in size], indexing="ij")
    coords = [g * sp[a] + org[a] for a, g in enumerate(grids)]
    ref = ndimage.map_coordinates(arr, coords, order=0, mode="nearest")
    ok = np.ones(size, dtype=bool)
    for a in range(3):
        frac = np.abs(coords[a] - np.floor(coords[a]) - 0.5)
        ok &= (frac > 1e-9) & (coords[a] >= -0.5 + 1e-9) & (coords[a] < arr.shape[a] - 0.5 - 1e-9)
    assert ok.mean() > 0.5
    assert np.array_equal(out[ok], ref[ok])


def test_spacing_resample_interior_matches_scipy():
    """Spacingd's trilinear resample (F.grid_sample, align_corners=False, border padding, float64) against
    scipy.ndimage.map_coordinates(order=1) at the voxel-centre coordinates the affine prescribes."""
    from scipy import ndimage
    g = torch.Generator().manual_seed(4)
    img = torch.randn((1, 18, 16, 9), generator=g)
    aff = np.diag([0.5, 0.5, 3.0, 1.0])
    out, new_aff, rec = osp.spacing_forward(img, aff, (1.0, 1.0, 1.0))
    xform = np.linalg.solve(aff, new_aff)          # output voxel index -> input voxel index
    grids = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in out.shape[1:]], indexing="ij")
    coords = [xform[a, a] * grids[a] + xform[a, 3] for a in range(3)]
    ref = ndimage.map_coordinates(img[0].double().numpy(), coords, order=1, mode="nearest")
    assert np.allclose(out[0].numpy(), ref, rtol=0, atol=1e-6)


def test_gaussian_importance_map_matches_scipy_window():
    """compute_importance_map(mode="gaussian", sigma_scale=0.125): the separable window exp(-x^2 / (2 sigma^2)) with
    sigma = 0.125 * n, x centred -- scipy.signal.windows.gaussian is the same window."""
    from scipy.signal import windows
    for roi in ((96, 96, 96), (32, 48, 64)):
        imap = osw.importance_map(roi, "gaussian", 0.125).numpy()
        w = [windows.gaussian(n, 0.125 * n).astype(np.float32) for n in roi]
        ref = w[0][:, None, None] * w[1][None, :, None] * w[2][None, None, :]
        ref = np.maximum(ref, max(float(ref.min()), 1e-3))
        assert np.allclose(imap, ref, rtol=2e-6, atol=0)
