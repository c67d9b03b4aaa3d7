"""GPU parity of the evaluation branch of predict() (labels supplied): the device confusion matrix against the numpy
oracle (bit-exact: integer counts), the metrics derived from it, and the files / prints of ``predict()``
(reference: seg/monai_unet.py:640-725, seg/evaluation.py:96-125)."""
import numpy as np
import pytest
import torch

from oracle import evaluation as oe
from oracle import spacing as osp
from tests.helpers import normalized_volume

pytestmark = pytest.mark.gpu
SMALL = dict(channels=(16, 32, 48), strides=(2, 2))


@pytest.mark.parametrize("n,classes", [(0, 3), (1, 2), (15, 3), (16, 3), (1000003, 10), (96 * 96 * 96, 20), (4099, 64)])
def test_confusion_matrix_matches_oracle(cuda_device, n, classes):
    from segmantic_b200.seg import evaluation as E
    rng = np.random.default_rng(n + classes)
    y = rng.integers(0, classes, n).astype(np.uint8)
    p = np.where(rng.random(n) < 0.7, y, rng.integers(0, classes, n)).astype(np.uint8)
    if n > 100:
        y[5], p[11] = 200, 255   # labels outside [0, num_classes): the pair is not counted
    ref = oe.confusion_matrix(classes, p, y)
    cm = E.confusion_matrix(classes, p, y)
    assert cm.dtype == np.int64 and np.array_equal(cm, ref)
    assert int(cm.sum()) == int(((y < classes) & (p < classes)).sum())
    # device tensors, unaligned views (the 16-byte path needs aligned pointers: falls back to the scalar loop), int64
    if n > 100:
        yd, pd = torch.from_numpy(y).to(cuda_device), torch.from_numpy(p).to(cuda_device)
        assert np.array_equal(E.confusion_matrix(classes, pd[3:], yd[3:]), oe.confusion_matrix(classes, p[3:], y[3:]))
        assert np.array_equal(E.confusion_matrix(classes, pd.long(), yd.long()), ref)
    assert np.allclose(E.class_dice(cm), oe.class_dice(ref), equal_nan=True)
    if n:
        a, b = E.confusion_metrics([E.confusion_counts(cm)]), oe.confusion_metrics([oe.confusion_counts(ref)])
        assert a == b


def test_metrics_known_answers():
    from segmantic_b200.seg import evaluation as E
    cm = np.array([[1, 1, 0], [0, 2, 1], [1, 0, 2]])
    assert np.allclose(E.class_dice(cm), [2 / 3, 2 / 3])
    assert np.allclose(E.class_dice(cm, include_background=True), [0.5, 2 / 3, 2 / 3])
    m = E.confusion_metrics([E.confusion_counts(cm)])
    assert m == {"sensitivity": 0.625, "specificity": 0.8125, "precision": 0.625, "accuracy": 0.75}
    empty_gt = np.array([[3, 1], [0, 0]])   # class 1 absent from the ground truth -> NaN (ignored by the means)
    assert np.isnan(E.class_dice(empty_gt)[0])


def test_predict_with_labels_reports_dice_and_confusion(cuda_device, tmp_path, capsys):
    """predict(test_labels=...) follows the reference's evaluation branch: the foreground crop comes from label > 0
    (monai_unet.py:167), Dice / confusion counts are taken on the pre-processed (cropped) grid before Invertd
    (:672-680), the tables are printed and mean_dice_<model>_generalized_score.txt holds one running mean per image
    (np.savetxt format).  Expected values: the oracle composition with the same label."""
    import json

    from oracle.predict import predict_volume as oracle_predict
    from segmantic_b200.image import nifti
    from segmantic_b200.seg.monai_unet import predict
    from segmantic_b200.synthetic import synthetic_lightning_checkpoint
    from tests.helpers import make_oracle_net
    ck = synthetic_lightning_checkpoint(num_classes=3, num_channels=1, spatial_dims=3, spatial_size=[16, 16, 16],
                                        seed=12, **SMALL)
    torch.save(ck, tmp_path / "model.ckpt")
    (tmp_path / "model.json").write_text(json.dumps({"channels": [16, 32, 48], "strides": [2, 2]}))
    onet, _ = make_oracle_net(3, 1, 3, seed=12, **SMALL)
    aff = osp.itk_geometry_to_ras_affine((1.0, 1.0, 1.0), (0.0, 0.0, 0.0), np.eye(3).flatten())
    imgs, labs, raws, gts = [], [], [], []
    for i in range(2):
        raw = (normalized_volume((36, 32, 24), seed=50 + i) * 40.0 + 10.0)[0]
        nifti.write(tmp_path / f"img{i}.nii.gz", raw.numpy(), aff)
        imgs.append(tmp_path / f"img{i}.nii.gz")
        raws.append(raw)
    out = tmp_path / "pass1"
    predict(tmp_path / "model.ckpt", imgs, None, out, {"Bone": 1, "Fat": 2})
    # ground truth = the first pass's prediction (image 1: shifted), zero outside a box that is smaller than the image
    for i in range(2):
        lab, _, _ = nifti.read(out / f"img{i}.nii.gz")
        gt = lab[0] if i == 0 else np.roll(lab[0], 2, axis=0)
        box = np.zeros_like(gt)
        box[4:30, 3:27, 2:21] = 1
        gt = (gt * box).astype(np.float32)
        nifti.write(tmp_path / f"lab{i}.nii.gz", gt, aff)
        labs.append(tmp_path / f"lab{i}.nii.gz")
        gts.append(torch.from_numpy(gt))
    capsys.readouterr()
    out2 = tmp_path / "pass2"
    predict(tmp_path / "model.ckpt", imgs, labs, out2, {"Bone": 1, "Fat": 2})
    text = capsys.readouterr().out
    assert "Class Dice:" in text and "Total Conf. Matrix Metrics:" in text and "Bone" in text and "sensitivity" in text
    scores = np.atleast_1d(np.loadtxt(out2 / "mean_dice_model_generalized_score.txt", delimiter=","))
    assert scores.shape == (2,)
    per_image = []
    for i in range(2):
        ref_lab, _, pred_net, label_net = oracle_predict(onet, raws[i][None], aff, (), roi=(16, 16, 16), label=gts[i])
        cm_ref = oe.confusion_matrix(3, pred_net.numpy(), label_net.numpy())
        cm = np.loadtxt(out2 / f"img{i}_confusion.csv", delimiter=",", dtype=np.int64)
        assert cm.sum() == cm_ref.sum() == label_net.numel()           # scored on the CROPPED grid
        assert label_net.numel() < raws[i].numel()
        assert np.abs(cm - cm_ref).sum() <= 0.004 * cm.sum()            # argmax near-ties only
        per_image.append(np.nanmean(oe.class_dice(cm)))
        pred, _, _ = nifti.read(out2 / f"img{i}.nii.gz")
        assert float((pred[0] != ref_lab.numpy()).mean()) < 2e-3
        outside = np.ones_like(pred[0], dtype=bool)
        nz = np.argwhere(gts[i].numpy() > 0)
        lo_, hi_ = nz.min(0), nz.max(0) + 1
        outside[lo_[0]:hi_[0], lo_[1]:hi_[1], lo_[2]:hi_[2]] = False
        assert not pred[0][outside].any()                               # inverse crop: label 0 outside the label's box
    assert np.isclose(scores[0], per_image[0]) and np.isclose(scores[1], np.mean(per_image))
