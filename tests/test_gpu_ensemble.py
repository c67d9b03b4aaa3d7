"""GPU parity of the ensemble combinations (mean / vote / select_best) against the oracle restatement of MONAI's
MeanEnsemble / VoteEnsemble and the reference's SelectBestEnsemble, and of ``ensemble_creator`` end to end
(reference: seg/monai_unet.py:834-1004, seg/transforms.py:15-61)."""
import json

import numpy as np
import pytest
import torch

from oracle import ensemble as oens
from oracle import spacing as osp
from oracle.predict import predict_volume as oracle_predict
from tests.helpers import make_oracle_net, normalized_volume

pytestmark = pytest.mark.gpu
SMALL = dict(channels=(16, 32, 48), strides=(2, 2))


@pytest.mark.parametrize("models,classes", [(1, 3), (3, 10), (5, 4)])
def test_combination_kernels_match_oracle(cuda_device, models, classes):
    from segmantic_b200.seg import ensemble as E
    g = torch.Generator().manual_seed(models * 10 + classes)
    shape = (17, 9, 31)
    logits = torch.randn((models, classes) + shape, generator=g)
    logits[:, :, 0] = logits[:, :1, 0]          # exact ties between classes -> lowest class
    weights = [0.5 + 0.1 * m for m in range(models)]
    for w in (None, weights):
        ref = oens.mean_ensemble(logits, w)
        lab, mean = E.mean_ensemble_argmax(logits.to(cuda_device), w, return_mean=True)
        assert float((mean.cpu() - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
        assert torch.equal(lab.cpu().long(), mean.cpu().argmax(0))
        gap = ref.topk(2, dim=0).values
        clear = (gap[0] - gap[1]) > 1e-5
        assert torch.equal(lab.cpu().long()[clear], ref.argmax(0)[clear])
    labels = torch.randint(0, classes, (models,) + shape, generator=g, dtype=torch.uint8)
    vote = E.vote_ensemble(labels.to(cuda_device), classes)
    assert torch.equal(vote.cpu().long(), oens.vote_ensemble(labels, classes))
    pairs = [(c, int(torch.randint(0, models, (1,), generator=g))) for c in range(1, classes)]
    sel = E.select_best_ensemble(labels.to(cuda_device), pairs)
    assert torch.equal(sel.cpu().long(), oens.select_best_ensemble(labels, dict(pairs)))


def test_ensemble_creator_end_to_end(cuda_device, tmp_path):
    """Three synthetic checkpoints, vote / mean / select_best: <image>_seg.nii.gz equals the oracle's combination of the
    three oracle predictions (outside argmax near-ties)."""
    from segmantic_b200.image import nifti
    from segmantic_b200.seg.monai_unet import ensemble_creator
    from segmantic_b200.synthetic import synthetic_lightning_checkpoint
    mdir = tmp_path / "models"
    mdir.mkdir()
    seeds, files = (12, 13, 14), []
    for i, sd_seed in enumerate(seeds):
        ck = synthetic_lightning_checkpoint(num_classes=3, num_channels=1, spatial_dims=3, spatial_size=[16, 16, 16],
                                            seed=sd_seed, **SMALL)
        f = mdir / f"epoch={i}-val_dice=0.{7 + i}.ckpt"
        torch.save(ck, f)
        f.with_suffix(".json").write_text(json.dumps({"channels": [16, 32, 48], "strides": [2, 2]}))
        files.append(f)
    aff = osp.itk_geometry_to_ras_affine((1.0, 1.0, 1.0), (3.0, -4.0, 5.0), np.eye(3).flatten())
    raw = (normalized_volume((36, 32, 24), seed=60) * 40.0 + 10.0)[0].numpy()
    nifti.write(tmp_path / "img.nii.gz", raw, aff)
    (tmp_path / "best.json").write_text(json.dumps({"Bone": 2, "Fat": 0}))
    tissue = {"Background": 0, "Bone": 1, "Fat": 2}
    # oracle: the three models' final label maps (vote / select_best are voxel-wise on label maps and commute with the
    # un-crop / un-flip), and for "mean" one oracle prediction whose predictor is the weighted mean of the three networks
    # (overlap blending is linear in the window logits)
    onets = [make_oracle_net(3, 1, 3, seed=sd_seed, **SMALL)[0] for sd_seed in seeds]
    x = torch.from_numpy(raw)[None]
    kw = dict(roi=(16, 16, 16), overlap=0.5, mode="constant")
    labels = torch.stack([oracle_predict(n, x, aff, (), **kw)[0] for n in onets])
    mean_lab, _ = oracle_predict(None, x, aff, (), predictor=lambda w: oens.mean_ensemble(
        torch.stack([n(w) for n in onets]), [0.7, 0.8, 0.9]), **kw)
    expect = {
        "vote": oens.vote_ensemble(labels, 3),
        "mean": mean_lab.long(),
        "select_best": oens.select_best_ensemble(labels, {1: 2, 2: 0}),
    }
    for cm in ("vote", "mean", "select_best"):
        out = tmp_path / cm
        ensemble_creator(files, [tmp_path / "img.nii.gz"], None, out, tissue, [], cm,
                         tmp_path / "best.json" if cm == "select_best" else None)
        lab, aff2, _ = nifti.read(out / "img_seg.nii.gz")
        assert np.allclose(aff2, aff)
        got = torch.from_numpy(lab[0]).long()
        assert tuple(got.shape) == tuple(expect[cm].shape)
        assert float((got != expect[cm]).float().mean()) < 2e-3, cm
    with pytest.raises(ValueError):
        ensemble_creator(files, [tmp_path / "img.nii.gz"], None, tmp_path / "x", tissue, [], "select_best", None)


def test_reference_fixtures_on_device(cuda_device):
    """The reference's own golden vectors (tests/seg/test_transforms.py:8-27 SelectBestEnsembled, tests/seg/
    test_evaluation.py:8-22 confusion matrix) through the device kernels."""
    from segmantic_b200.seg import ensemble as E
    from segmantic_b200.seg import evaluation as EV
    preds = torch.stack([torch.ones(3, dtype=torch.uint8), torch.tensor([2, 0, 2], dtype=torch.uint8),
                         torch.tensor([2, 1, 0], dtype=torch.uint8)]).to(cuda_device)
    out = E.select_best_ensemble(preds, [(1, 0), (2, 1), (0, 2)])      # label_model_dict={1: 0, 2: 1, 0: 2}, in order
    assert out.cpu().tolist() == [2, 1, 0]
    field = np.zeros((10, 10), np.uint8)
    field[2:3, 2:4] = 1
    field[3:5, 3:4] = 2
    view = field.T.flatten()
    cm = EV.confusion_matrix(3, view, view)
    assert np.all(np.diagonal(cm) == np.bincount(view))
    assert np.all(np.diagonal(cm, offset=1) == 0) and np.all(np.diagonal(cm, offset=-1) == 0)
