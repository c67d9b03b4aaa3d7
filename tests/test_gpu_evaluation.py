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
    """predict(test_labels=...) prints the reference's tables and writes mean_dice_<model>_generalized_score.txt (one
    running mean per image, np.savetxt format) next to the label maps."""
    import json

    from segmantic_b200.image import nifti
    from segmantic_b200.seg.monai_unet import predict
    from segmantic_b200.synthetic import synthetic_lightning_checkpoint
    ck = synthetic_lightning_checkpoint(num_classes=3, num_channels=1, spatial_dims=3, spatial_size=[16, 16, 16],
                                        seed=12, **SMALL)
    torch.save(ck, tmp_path / "model.ckpt")
    (tmp_path / "model.json").write_text(json.dumps({"channels": [16, 32, 48], "strides": [2, 2]}))
    aff = osp.itk_geometry_to_ras_affine((1.0, 1.0, 1.0), (0.0, 0.0, 0.0), np.eye(3).flatten())
    imgs, labs = [], []
    for i in range(2):
        raw = (normalized_volume((36, 32, 24), seed=50 + i) * 40.0 + 10.0)[0].numpy()
        nifti.write(tmp_path / f"img{i}.nii.gz", raw, aff)
        imgs.append(tmp_path / f"img{i}.nii.gz")
    out = tmp_path / "pass1"
    predict(tmp_path / "model.ckpt", imgs, None, out, {"Bone": 1, "Fat": 2})
    # ground truth = the prediction itself for image 0 (Dice 1), a shifted copy for image 1 (Dice < 1)
    for i in range(2):
        lab, _, _ = nifti.read(out / f"img{i}.nii.gz")
        gt = lab[0] if i == 0 else np.roll(lab[0], 2, axis=0)
        nifti.write(tmp_path / f"lab{i}.nii.gz", gt.astype(np.float32), aff)
        labs.append(tmp_path / f"lab{i}.nii.gz")
    capsys.readouterr()
    out2 = tmp_path / "pass2"
    predict(tmp_path / "model.ckpt", imgs, labs, out2, {"Bone": 1, "Fat": 2})
    text = capsys.readouterr().out
    assert "Class Dice:" in text and "Total Conf. Matrix Metrics:" in text and "Bone" in text and "sensitivity" in text
    scores = np.atleast_1d(np.loadtxt(out2 / "mean_dice_model_generalized_score.txt", delimiter=","))
    assert scores.shape == (2,)
    pred0, _, _ = nifti.read(out2 / "img0.nii.gz")
    pred1, _, _ = nifti.read(out2 / "img1.nii.gz")
    gt1, _, _ = nifti.read(labs[1])
    d0 = oe.class_dice(oe.confusion_matrix(3, pred0[0], pred0[0]))
    d1 = oe.class_dice(oe.confusion_matrix(3, pred1[0], gt1[0]))
    assert np.isclose(scores[0], np.nanmean(d0)) and np.isclose(scores[1], np.nanmean([np.nanmean(d0), np.nanmean(d1)]))
    cm1 = np.loadtxt(out2 / "img1_confusion.csv", delimiter=",", dtype=np.int64)
    assert np.array_equal(cm1, oe.confusion_matrix(3, pred1[0], gt1[0]))
