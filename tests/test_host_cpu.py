"""CPU tests of the host side: schedule, slab partition, layer schema / BN folding, checkpoint loading,
NIfTI I/O, tissue lists, CLI wiring, and that the C-ABI library exports what include/*.h declares."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sliding_window as osw
from oracle.unet import UNet, load_checkpoint_into
from segmantic_b200.seg import sliding_window as sw
from segmantic_b200.seg import unet_spec
from segmantic_b200.synthetic import synthetic_lightning_checkpoint, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("size,roi,overlap", [((256, 256, 256), (96, 96, 96), 0.5), ((100, 97, 300), (96, 96, 96), 0.25),
                                              ((40, 50, 60), (32, 32, 32), 0.5), ((1, 512, 512), (1, 96, 96), 0.25)])
def test_schedule_equals_oracle(size, roi, overlap):
    s = sw.make_schedule(size, roi, overlap, "gaussian")
    assert s.windows() == [tuple(w) for w in osw.window_starts(s.padded_size, roi, overlap)]
    imap = osw.importance_map(roi, "gaussian")
    mine = torch.clamp((s.tables[0][:, None, None] * s.tables[1][None, :, None]) * s.tables[2][None, None, :], min=s.floor)
    assert torch.equal(mine, imap)  # bit-identical importance map


def test_schedule_padding_and_errors():
    s = sw.make_schedule((20, 100, 100), (32, 32, 32), 0.25)
    assert s.padded_size == (32, 100, 100) and s.pad_lo == (6, 0, 0)
    with pytest.raises(ValueError):
        sw.make_schedule((10, 10, 10), (8, 8, 8), 1.0)
    with pytest.raises(ValueError):
        sw.importance_tables((8, 8, 8), "nope")


@pytest.mark.parametrize("size,world", [((256, 64, 64), 2), ((512, 64, 64), 4), ((2048, 96, 96), 8), ((150, 48, 64), 3),
                                        ((96, 96, 96), 4)])
def test_slab_partition_covers_and_orders(size, world):
    roi = (96, 96, 96) if size[0] >= 96 else (32, 32, 32)
    if size == (150, 48, 64):
        roi = (32, 32, 32)
    s = sw.make_schedule(size, roi, 0.5, "gaussian")
    parts = sw.slab_partition(s, world)
    assert len(parts) == world and parts[0]["x0"] == 0 and parts[-1]["x1"] == s.padded_size[0]
    for a, b in zip(parts, parts[1:]):
        assert a["x1"] == b["x0"]
    s0 = s.starts[0]
    for p in parts:
        if p["x1"] == p["x0"]:
            continue
        rows = [j for j in range(len(s0)) if s0[j] < p["x1"] and s0[j] + roi[0] > p["x0"]]
        assert rows == list(range(p["a0_begin"], p["a0_end"]))       # every intersecting row, in order
        assert p["vol_x0"] == s0[rows[0]] and p["vol_x1"] == s0[rows[-1]] + roi[0]


def test_spec_matches_oracle_module_tree():
    for dims, cin, cout, ch, st in ((3, 1, 3, (16, 32, 64, 128, 256), (2, 2, 2, 2)), (2, 2, 10, (16, 32, 64), (2, 2)),
                                    (3, 2, 4, (8, 16, 24), (2, 1))):
        net = UNet(dims, cin, cout, ch, st)
        schema = unet_spec.state_dict_schema(dims, cin, cout, ch, st)
        sd = net.state_dict()
        assert list(sd.keys()) == list(schema.keys())
        assert all(tuple(sd[k].shape) == tuple(v) for k, v in schema.items())


def test_fold_batchnorm_reproduces_eval_forward():
    """Folded conv (+PReLU) == conv -> BatchNorm(eval) -> PReLU of the oracle, layer by layer."""
    sd = synthetic_state_dict(3, 1, 3, (16, 32, 48), (2, 2), seed=3)
    net = UNet(3, 1, 3, (16, 32, 48), (2, 2))
    load_checkpoint_into(net, sd)
    net.eval()
    specs = unet_spec.unet_conv_specs(1, 3, (16, 32, 48), (2, 2))
    folded = unet_spec.fold_batchnorm(sd, specs)
    x = torch.randn(1, 1, 16, 16, 16)
    f0 = folded[0]  # down0.unit0
    y = F.conv3d(x, f0.weight, f0.bias, stride=2, padding=1)
    y = torch.where(y > 0, y, y * f0.alpha)
    with torch.no_grad():
        ref = net.model[0].conv.unit0(x)
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-5)
    up = next(f for f in folded if f.spec.role == "up0.convT")
    xin = torch.randn(1, up.spec.cin, 8, 8, 8)
    y = F.conv_transpose3d(xin, up.weight, up.bias, stride=2, padding=1, output_padding=1)
    y = torch.where(y > 0, y, y * up.alpha)
    with torch.no_grad():
        ref = net.model[2][0](xin)
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-5)
    roles = [s.role for s in specs]
    assert roles[:3] == ["down0.unit0", "down0.unit1", "down0.residual"] and roles[-1] == "up0.ru"
    assert len(specs) == 5 * 2 + 3


def test_fold_rejects_wrong_checkpoints():
    sd = synthetic_state_dict(3, 1, 3, seed=0)
    specs = unet_spec.unet_conv_specs(1, 5)
    with pytest.raises(ValueError):
        unet_spec.fold_batchnorm(sd, specs)
    del sd["model.0.conv.unit0.conv.weight"]
    with pytest.raises(KeyError):
        unet_spec.fold_batchnorm(sd, unet_spec.unet_conv_specs(1, 3))
    with pytest.raises(ValueError):
        unet_spec.unet_conv_specs(1, 3, (16, 32, 64), (2, 3))


def test_checkpoint_loading_lightning_and_plain(tmp_path):
    from segmantic_b200.seg.monai_unet import Net
    ck = synthetic_lightning_checkpoint(num_classes=4, num_channels=2, spatial_dims=3, channels=(16, 32, 48),
                                        strides=(2, 2), seed=1)
    p = tmp_path / "model.ckpt"
    torch.save(ck, p)
    net = Net.load_from_checkpoint(p)
    assert net.num_classes == 4 and net.hparams.num_channels == 2 and net.hparams.channels == (16, 32, 48)
    assert "model.0.conv.unit0.conv.weight" in {k.replace("_model.", "") for k in net._state_dict}
    # keyword overrides win over saved hyper-parameters (monai_unet.py:571-573)
    net2 = Net.load_from_checkpoint(p, dropout=0.5)
    assert net2.hparams.dropout == 0.5
    # plain MONAI state_dict as written by scripts/extract_unet.py
    plain = {k[len("_model."):]: v for k, v in ck["state_dict"].items()}
    p2 = tmp_path / "model.pth"
    torch.save(plain, p2)
    net3 = Net.load_from_checkpoint(p2, channels=(16, 32, 48), strides=(2, 2))
    assert net3.num_classes == 4 and net3.hparams.num_channels == 2 and net3.spatial_dims == 3


def test_no_cpu_fallback():
    from segmantic_b200.seg.engine import UNetB200
    from segmantic_b200.seg.monai_unet import predict
    from segmantic_b200.seg.utils import make_device
    assert make_device([-1]) == torch.device("cpu")
    with pytest.raises(RuntimeError):
        UNetB200(synthetic_state_dict(3, 1, 3, seed=0), out_channels=3, device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            UNetB200(synthetic_state_dict(3, 1, 3, seed=0), out_channels=3, device="cuda:0")


def test_nifti_roundtrip_and_geometry(tmp_path):
    from segmantic_b200.image import nifti
    from segmantic_b200.seg.transforms import itk_geometry_to_ras_affine, ras_affine_to_itk_geometry
    arr = (np.random.default_rng(0).random((5, 6, 7)) * 100).astype(np.float32)
    aff = itk_geometry_to_ras_affine((0.5, 0.6, 0.7), (10.0, 20.0, 30.0), np.eye(3).flatten())
    for name in ("a.nii", "a.nii.gz"):
        nifti.write(tmp_path / name, arr, aff)
        back, aff2, _ = nifti.read(tmp_path / name)
        assert back.shape == (1, 5, 6, 7) and np.array_equal(back[0], arr) and np.allclose(aff, aff2)
    sp, org, direction = ras_affine_to_itk_geometry(aff)
    assert np.allclose(sp, (0.5, 0.6, 0.7)) and np.allclose(org, (10, 20, 30)) and np.allclose(direction, np.eye(3).flatten())
    nifti.write(tmp_path / "l.nii.gz", np.arange(24, dtype=np.uint8).reshape(2, 3, 4), np.eye(4))
    lab, _, _ = nifti.read(tmp_path / "l.nii.gz")
    assert lab.dtype == np.float32 and lab[0, 1, 2, 3] == 23


def test_tissue_lists(tmp_path):
    from segmantic_b200.image import labels
    p = tmp_path / "tissues.txt"
    p.write_text("V7\nN3\nC0.00 0.00 1.00 0.50 Bone\nC0.00 1.00 0.00 0.50 Fat\nC1.00 0.00 0.00 0.50 Skin\n")
    assert labels.load_tissue_list(p) == {"Background": 0, "Bone": 1, "Fat": 2, "Skin": 3}
    p.write_text("V7\nN2\nC0 0 1 0.5 Bone\nC0 1 0 0.5 Bone\n")
    with pytest.raises(KeyError):
        labels.load_tissue_list(p)
    d = tmp_path / "datalist.json"
    d.write_text(json.dumps({"labels": {"1": "Bone", "2": "Fat"}, "test": ["a.nii.gz", {"image": "b.nii.gz", "label": "bl.nii.gz"}]}))
    assert labels.load_decathlon_tissuelist(d) == {"Bone": 1, "Fat": 2, "Background": 0}
    labels.save_tissue_list({"Bone": 1, "Fat": 2}, tmp_path / "out.txt")
    assert labels.load_tissue_list(tmp_path / "out.txt") == {"Background": 0, "Bone": 1, "Fat": 2}
    from segmantic_b200.commands.monai_unet_cli import load_decathlon_datalist
    items = load_decathlon_datalist(d, "test")
    assert items[0]["image"] == str(tmp_path / "a.nii.gz") and items[1]["label"] == str(tmp_path / "bl.nii.gz")
    with pytest.raises(ValueError):
        load_decathlon_datalist(d, "validation")


def test_cli_exposes_reference_options():
    out = subprocess.run([sys.executable, "-m", "segmantic_b200.commands.monai_unet_cli", "predict", "--help"],
                         capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    for opt in ("--datalist", "-d", "--model-file", "-m", "--tissue-list", "-t", "--results-dir", "-r", "--spacing",
                "--gpu-ids", "--datalist-key"):
        assert opt in out.stdout, opt


def test_image_helpers_match_reference_fixture():
    """tests/image/test_image.py:7-30 of the reference (extract_slices, pad + crop_center round trip)."""
    from segmantic_b200.image import processing as P
    lab = P.make_image(shape=(5, 5, 5), spacing=(0.5, 0.6, 0.7))
    for i in range(5):
        lab.array[..., i] = i
    slices = P.extract_slices(lab, axis=2)
    assert slices[0].GetSpacing() == (0.5, 0.6)
    for k, sl in enumerate(slices):
        assert np.all(sl.array == k)
    padded = P.pad(lab, target_size=(9, 9, 9))
    cropped = P.crop_center(padded, target_size=(5, 5, 5))
    assert cropped.GetSpacing() == lab.GetSpacing() and cropped.GetOrigin() == lab.GetOrigin()
    assert np.array_equal(cropped.array, lab.array)
    assert P.crop_center(lab, target_size=(5, 5, 1)).GetSize()[2] == 1
    with pytest.raises(ValueError):
        P.make_image((4, 4), spacing=(1.0, 1.0, 1.0))


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports exactly the entry points include/segmantic_b200.h declares,
    and the ctypes table mirrors the header one to one (no compute calls here: there is no GPU)."""
    from segmantic_b200 import _lib
    header = open(os.path.join(ROOT, "include", "segmantic_b200.h")).read()
    declared = set(re.findall(r"SGM_API\s+[\w\s\*]+?\b(sgm_\w+)\s*\(", header))
    assert declared, "no SGM_API declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(str(_lib.lib_path()))
    for name in declared:
        assert hasattr(lib, name), f"{name} missing from the shared library"
    assert _lib.load().sgm_version() >= 100
    if not torch.cuda.is_available():  # compute entry points fail loudly without a device
        desc = _lib.UnetDesc()
        desc.spatial_dims, desc.n_levels, desc.n_convs = 3, 1, 8
        desc.in_channels = desc.out_channels = 1
        desc.channels[0] = desc.channels[1] = 16
        handle = ctypes.c_void_p()
        rc = _lib.load().sgm_unet_create(ctypes.byref(desc), ctypes.byref(handle))
        assert rc == -2 and b"no CUDA device" in _lib.load().sgm_last_error()


def test_ensemble_cli_and_argument_validation(tmp_path):
    """`ensemble-predict` exposes the reference's options (commands/monai_unet_cli.py:212-246); ensemble_creator rejects what
    the reference rejects (select_best without a candidate file, monai_unet.py:859-865) before touching a device."""
    out = subprocess.run([sys.executable, "-m", "segmantic_b200.commands.monai_unet_cli", "ensemble-predict", "--help"],
                         capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    for opt in ("--datalist", "-d", "--models-dir", "-m", "--tissue-list", "-t", "--results-dir", "-r",
                "--combination-mode", "-cm", "--candidate-yaml", "-cy", "--spacing", "--gpu-ids", "--datalist-key"):
        assert opt in out.stdout, opt
    from segmantic_b200.seg.monai_unet import ensemble_creator
    with pytest.raises(ValueError):
        ensemble_creator([tmp_path / "a.ckpt"], [], None, None, None, [], "select_best", None)
    with pytest.raises(ValueError):
        ensemble_creator([tmp_path / "a.ckpt"], [], None, None, None, [], "median", None)
    with pytest.raises(RuntimeError):   # gpu_ids=[-1] is the reference's CPU switch: no CPU path here
        ensemble_creator([tmp_path / "a.ckpt"], [], None, None, None, [], "vote", None, gpu_ids=[-1])


def test_new_entry_points_fail_loudly_without_a_device():
    """Evaluation / ensemble entry points: no CPU fallback (host wrappers raise, the C ABI returns SGM_ERR_CUDA)."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from segmantic_b200 import _lib
    from segmantic_b200.seg import ensemble as E
    from segmantic_b200.seg import evaluation as EV
    with pytest.raises(RuntimeError):
        EV.confusion_matrix(3, np.zeros(8, np.uint8), np.zeros(8, np.uint8))
    with pytest.raises(ValueError):
        E.vote_ensemble(torch.zeros((2, 4), dtype=torch.uint8), 3)       # host tensor: not a CUDA tensor
    lib = _lib.load()
    buf = (ctypes.c_int64 * 16)()
    rc = lib.sgm_confusion_matrix(ctypes.cast(buf, ctypes.c_void_p), ctypes.cast(buf, ctypes.c_void_p), 8, 3,
                                  ctypes.cast(buf, ctypes.c_void_p), None, None)
    assert rc == -2
    # pure host arithmetic of the evaluation module needs no device
    cm = np.array([[1, 1, 0], [0, 2, 1], [1, 0, 2]])
    assert np.allclose(EV.class_dice(cm), [2 / 3, 2 / 3])
    assert EV.confusion_metrics([EV.confusion_counts(cm)])["accuracy"] == 0.75


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores; needs no GPU): one JSON line with the
    contract's keys, `impl: reference`, a cpu_baseline describing the run and an e2e object repeating the value."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-windows", "1"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "Mvoxel/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_reference_label_fixtures(tmp_path):
    """tests/image/test_labels.py of the reference replayed: tissue-list round trip (incl. colours) and tissue merging."""
    from segmantic_b200.image import labels
    tissue_map = {"Background": 0, "Bone": 1, "Fat": 2, "Skin": 3}
    p = tmp_path / "tissue.txt"
    labels.save_tissue_list(tissue_map, p)
    assert labels.load_tissue_list(p) == tissue_map
    assert len(labels.load_tissue_colors(p)) == len(tissue_map)

    def _map_name(name):
        return name if name in ("Background", "Bone") else "Other_tissue"

    omap, i2o = labels.build_tissue_mapping(tissue_map, _map_name)
    assert len(omap) == 3
    assert omap == {_map_name(n1): i2o[i1] for n1, i1 in tissue_map.items()}
    assert omap == {"Background": 0, "Bone": 1, "Other_tissue": 2} and i2o.tolist() == [0, 1, 2, 2]


def test_segmantic_import_shim_and_console_script():
    """`segmantic.*` imports of the reference resolve to the drop-in, and pyproject.toml declares the reference's
    console script (`segmantic-unet`, /root/reference/pyproject.toml:64-65)."""
    import segmantic_b200.seg.monai_unet as impl
    from segmantic.commands.monai_unet_cli import main
    from segmantic.image.processing import resample, resample_to_ref
    from segmantic.seg import monai_unet
    assert monai_unet is impl and callable(monai_unet.predict) and callable(main)
    assert callable(resample) and callable(resample_to_ref)
    text = open(os.path.join(ROOT, "pyproject.toml")).read()
    assert 'segmantic-unet = "segmantic_b200.commands.monai_unet_cli:main"' in text
    with pytest.raises(ImportError):
        import segmantic.seg.dataset  # noqa: F401  (not on the prediction path: absent, as documented)


def test_plan_chunks_for_volumes_beyond_hbm():
    """Single-GPU chunking of BASELINE configs[3] (2100 windows x 20 classes x 96^3 x 4 B = 148 GB of deferred-blend
    buffer): the smallest chunk count whose two alternating buffers fit the budget; chunks tile the window list."""
    from segmantic_b200.seg.sliding_window import make_schedule, plan_chunks
    sched = make_schedule((1024, 512, 512), (96, 96, 96), 0.5, "gaussian")
    per_window = 20 * 96 ** 3 * 4
    assert sched.n_windows == 2100 and sched.n_windows * per_window > 140e9
    parts = plan_chunks(sched, per_window, 8 << 30, 140 << 30)
    assert parts is not None and len(parts) == 3
    assert parts[0]["w_lo"] == 0 and parts[-1]["w_hi"] == 2100
    assert all(a["w_hi"] == b["w_lo"] and b["wb"] == a["send_lo"] for a, b in zip(parts, parts[1:]))
    held = sorted((p["w_hi"] - p["wb"]) * per_window for p in parts)
    assert held[-1] + held[-2] + (8 << 30) <= 140 << 30
    assert len(plan_chunks(sched, per_window, 8 << 30, 60 << 30)) > 3          # a smaller budget -> more chunks
    assert plan_chunks(sched, per_window, 8 << 30, 10 << 30) is None           # nothing fits: caller falls back


def test_bench_block_is_thread_count_independent():
    """bench.make_block normalises the synthetic block with float64, digit-rounded statistics: a float32 parallel
    mean() depends on the number of host threads (torchrun sets OMP_NUM_THREADS=1), which fed the N = 1 and N > 1 runs
    volumes that differed in the last bit."""
    import bench
    old = torch.get_num_threads()
    try:
        outs = []
        for t in (1, max(2, old)):
            torch.set_num_threads(t)
            small = bench.VOL
            bench.VOL = (64, 64, 64)
            try:
                outs.append(bench.make_block(seed=3)[1])
            finally:
                bench.VOL = small
        assert torch.equal(outs[0], outs[1])
    finally:
        torch.set_num_threads(old)


def test_predict_volumes_needs_a_cuda_device():
    from segmantic_b200.seg.monai_unet import Net, predict_volumes
    from segmantic_b200.synthetic import synthetic_state_dict
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    net = Net(num_classes=3, num_channels=1, spatial_dims=3, channels=(16, 32, 48), strides=(2, 2))
    net.load_state_dict(synthetic_state_dict(3, 1, 3, (16, 32, 48), (2, 2), seed=1))
    with pytest.raises(RuntimeError):
        list(predict_volumes(net, [torch.zeros(1, 16, 16, 16)]))
