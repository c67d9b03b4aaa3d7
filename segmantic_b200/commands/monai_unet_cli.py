"""``segmantic-unet predict`` on the B200-native engine.

Mirrors the ``predict`` command of ``/root/reference/src/segmantic/commands/monai_unet_cli.py:165-209``
(same options: ``--datalist/-d``, ``--model-file/-m``, ``--tissue-list/-t``, ``--results-dir/-r``,
``--spacing``, ``--gpu-ids``, ``--datalist-key``), plus keyword extras whose defaults reproduce the
reference (``--overlap 0.25 --mode constant --precision fp32 --invert logits``), and the ``ensemble-predict``
command (``:212-264``).  The training / cross-validation commands of the reference are out of scope (SURVEY.md
section 8).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import List, Optional

import typer

from ..image.labels import load_decathlon_tissuelist, load_tissue_list
from ..seg import monai_unet

app = typer.Typer()


def load_decathlon_datalist(datalist_file: Path, data_list_key: str = "test") -> List[dict]:
    """MONAI ``load_decathlon_datalist`` semantics: items are file names or ``{"image":..,"label":..}``
    dicts; relative paths are resolved against the json's directory."""
    datalist_file = Path(datalist_file)
    content = json.loads(datalist_file.read_text())
    if data_list_key not in content:
        raise ValueError(f'Data list {data_list_key} not specified in "{datalist_file}".')
    base = datalist_file.parent
    out = []
    for item in content[data_list_key]:
        if not isinstance(item, dict):
            item = {"image": item}
        fixed = {}
        for k, v in item.items():
            if isinstance(v, str) and k in ("image", "label") and not Path(v).is_absolute():
                v = str(base / v)
            fixed[k] = v
        out.append(fixed)
    return out


@app.command()
def predict(
    datalist_file: Path = typer.Option(..., "--datalist", "-d", help="decathlon style datalist json file"),
    model_file: Path = typer.Option(..., "--model-file", "-m", help="saved model checkpoint"),
    tissue_list: Optional[Path] = typer.Option(None, "--tissue-list", "-t", help="label descriptors in iSEG format"),
    results_dir: Optional[Path] = typer.Option(None, "--results-dir", "-r", help="output directory"),
    spacing: List[float] = typer.Option([], "--spacing", help="if specified, the image is first resampled"),
    gpu_ids: List[int] = [0],
    datalist_key: str = "test",
    overlap: float = typer.Option(0.25, help="sliding-window overlap (reference default 0.25)"),
    mode: str = typer.Option("constant", help="blend mode: constant | gaussian"),
    precision: str = typer.Option("fp32", help="fp32 (reference numerics) | bf16 (tcgen05 tensor cores)"),
    invert: str = typer.Option("logits", help="logits (reference: Invertd then argmax) | labels (nearest)"),
) -> None:
    """Predict segmentations

    Example invocation:

        -d ./datalist.json -m model.ckpt --results-dir ./results
    """
    datalist = load_decathlon_datalist(datalist_file, data_list_key=datalist_key)
    test_images = [Path(d["image"]) for d in datalist]
    test_labels = [Path(d["label"]) for d in datalist if "label" in d]
    if tissue_list is not None:
        tissue_dict = load_tissue_list(tissue_list)
    else:
        tissue_dict = load_decathlon_tissuelist(datalist_file)
    monai_unet.predict(model_file=model_file, test_images=test_images, test_labels=test_labels,
                       tissue_dict=tissue_dict, output_dir=results_dir, spacing=spacing, gpu_ids=gpu_ids,
                       overlap=overlap, mode=mode, precision=precision, invert=invert)


@app.command()
def ensemble_predict(
    datalist_file: Path = typer.Option(..., "--datalist", "-d", help="decathlon style datalist json file"),
    models_dir: Path = typer.Option(..., "--models-dir", "-m", help="saved model checkpoints"),
    tissue_list: Optional[Path] = typer.Option(None, "--tissue-list", "-t", help="label descriptors in iSEG format"),
    results_dir: Optional[Path] = typer.Option(None, "--results-dir", "-r", help="output directory"),
    combination_mode: str = typer.Option(..., "--combination-mode", "-cm", help="mean | vote | select_best"),
    candidate_per_tissue_path: Optional[Path] = typer.Option(None, "--candidate-yaml", "-cy",
                                                             help="yaml with best model for tissues"),
    spacing: List[float] = typer.Option([], "--spacing", help="if specified, the image is first resampled"),
    gpu_ids: List[int] = [0],
    datalist_key: str = "test",
    precision: str = typer.Option("fp32", help="fp32 (reference numerics) | bf16 (tcgen05 tensor cores)"),
) -> None:
    """Ensemble-based prediction

    Example invocation:

        -d ./datalist.json -m ./training_01 -cm vote --results-dir ./results --tissue-list ./dataset/labels.txt
    """
    datalist = load_decathlon_datalist(datalist_file, data_list_key=datalist_key)
    test_images = [Path(d["image"]) for d in datalist]
    test_labels = [Path(d["label"]) for d in datalist if "label" in d]
    if tissue_list is not None:
        tissue_dict = load_tissue_list(tissue_list)
    else:
        tissue_dict = load_decathlon_tissuelist(datalist_file)
    monai_unet.ensemble_creator(model_files=sorted(f for f in Path(models_dir).glob("*.ckpt")), test_images=test_images,
                                test_labels=test_labels, tissue_dict=tissue_dict, output_dir=results_dir,
                                combination_mode=combination_mode, candidate_per_tissue_path=candidate_per_tissue_path,
                                spacing=spacing, gpu_ids=gpu_ids, precision=precision)


@app.callback()
def _main() -> None:
    """segmantic-unet (B200-native prediction path)."""


def main() -> None:
    app()


if __name__ == "__main__":
    main()
