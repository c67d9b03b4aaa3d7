"""Import shim: ``segmantic.*`` resolves to the B200-native drop-in ``segmantic_b200.*`` for the modules on the
prediction path, so code written against the reference keeps its imports --

    from segmantic.seg import monai_unet          # /root/reference/src/segmantic/seg/monai_unet.py
    from segmantic.image.processing import resample_to_ref
    from segmantic.commands.monai_unet_cli import main

Only the modules the drop-in implements are aliased (``seg.monai_unet``, ``seg.utils``, ``seg.evaluation``,
``seg.transforms``, ``image.processing``, ``image.labels``, ``commands.monai_unet_cli``); anything
else of the reference (training, datasets, plotting ...) raises ImportError as an absent module would.
"""
import importlib
import sys

_ALIASED = ("seg", "seg.monai_unet", "seg.utils", "seg.evaluation", "seg.transforms", "image", "image.processing",
            "image.labels", "commands", "commands.monai_unet_cli")

for _name in _ALIASED:
    _mod = importlib.import_module("segmantic_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
