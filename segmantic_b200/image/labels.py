"""Tissue-list helpers used by the ``predict`` command.

Same names and file formats as ``/root/reference/src/segmantic/image/labels.py``:
``load_tissue_list`` (iSEG format, ``:89-109``) and ``load_decathlon_tissuelist`` (``:112-117``) are what
``segmantic-unet predict`` calls (``commands/monai_unet_cli.py:196-199``); ``save_tissue_list`` is the
writer of the same format.  Pure host-side parsing -- nothing here touches the GPU.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict


def load_tissue_list(file_name: Path) -> Dict[str, int]:
    """Parse an iSEG tissue list (``V7`` / ``N<count>`` / ``C<r> <g> <b> <a> <name>`` lines).
    Label 0 is the implicit ``Background``; tissues are numbered in file order from 1."""
    names = {"Background": 0}
    with open(file_name) as f:
        for raw in f:
            if not raw.startswith("C"):
                continue
            tissue = raw.strip().rsplit(" ", 1)[-1].rstrip()
            if tissue in names:
                raise KeyError(f"duplicate label '{tissue}' found in '{file_name}'")
            names[tissue] = len(names)
    return names


def load_decathlon_tissuelist(file_name: Path) -> Dict[str, int]:
    """Tissue names from the ``labels`` object of a decathlon-style datalist json."""
    print(f"Reading {file_name}")
    labels = dict(json.loads(Path(file_name).read_text())["labels"])
    labels["0"] = "Background"
    return {name: int(idx) for idx, name in labels.items()}


def save_tissue_list(tissue_label_map: Dict[str, int], tissue_list_file_name: Path) -> None:
    """Write an iSEG tissue list (grey-scale colours; label 0 / Background is implicit)."""
    by_label = {}
    for name, label in tissue_label_map.items():
        if label in by_label:
            raise KeyError("duplicate labels found in 'tissue_label_map'")
        by_label[label] = name
    count = max(by_label) if by_label else 0
    with open(tissue_list_file_name, "w") as f:
        print("V7", file=f)
        print(f"N{count}", file=f)
        for label in range(1, count + 1):
            g = label / max(count, 1)
            print(f"C{g:.2f} {1 - g:.2f} {0.5:.2f} {0.5:.2f} {by_label[label]}", file=f)


def load_tissue_colors(file_name: Path) -> Dict[int, tuple]:
    """``{label: (r, g, b)}`` from the ``C<r> <g> <b> <a> <name>`` lines of an iSEG tissue list (reference
    ``image/labels.py:120-140``); label 0 (Background) is black."""
    colors = {0: (0.0, 0.0, 0.0)}
    with open(file_name) as f:
        for raw in f:
            if raw.startswith("C"):
                r, g, b = (float(tok) for tok in raw[1:].split()[:3])
                colors[len(colors)] = (r, g, b)
    return colors


def build_tissue_mapping(input_label_map: Dict[str, int], mapper):
    """Merge / rename tissues (reference ``image/labels.py:13-37``): ``mapper`` maps a tissue name to its new name.
    Returns the new ``{name: label}`` dict (Background first, the other names sorted) and the uint16 lookup table
    ``old label -> new label`` to apply to a label field."""
    import numpy as np

    new_names = sorted({mapper(name) for name in input_label_map} - {"Background"})
    output_label_map = {name: i for i, name in enumerate(["Background"] + new_names)}
    lut = np.zeros(len(input_label_map), dtype=np.uint16)
    for name, label in input_label_map.items():
        lut[label] = output_label_map[mapper(name)]
    return output_label_map, lut
