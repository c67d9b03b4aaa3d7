"""Device selection with the semantics of ``segmantic.seg.utils.make_device``
(``/root/reference/src/segmantic/seg/utils.py:4-12``): only the FIRST id of ``gpu_ids`` counts; an empty list means
``cuda:0`` when CUDA is present; a negative id (or no CUDA at all) means the CPU -- which every compute entry point of
this package then refuses (there is no CPU path)."""
from typing import Sequence

import torch


def make_device(gpu_ids: Sequence[int]) -> torch.device:
    first = gpu_ids[0] if len(gpu_ids) else (0 if torch.cuda.is_available() else -1)
    return torch.device("cpu") if first < 0 else torch.device("cuda", int(first))
