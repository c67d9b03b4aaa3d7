"""``segmantic.seg.utils`` (``/root/reference/src/segmantic/seg/utils.py:4-12``)."""
import torch


def make_device(gpu_ids: list) -> torch.device:
    # use by default if none specified
    if not gpu_ids and torch.cuda.is_available():
        gpu_ids = [0]
    # negative index means no gpu
    if not gpu_ids or gpu_ids[0] < 0:
        return torch.device("cpu")
    # use gpu
    return torch.device(f"cuda:{gpu_ids[0]}")
