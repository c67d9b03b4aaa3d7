"""``segmantic.image.utils.array_view_reverse_ordering`` (``/root/reference/src/segmantic/image/utils.py:13-14``): the
zero-copy view that turns a numpy array in SimpleITK's ``[z, y, x]`` order into ITK index order ``[x, y, z]`` (and
back) -- the convention ``image/processing.Image`` and the NIfTI reader / writer use.  (The VTK converters of the
reference module are visualisation helpers: out of scope.)"""
import numpy as np


def array_view_reverse_ordering(x: np.ndarray) -> np.ndarray:
    """View of ``x`` with its axes in reverse order (no copy)."""
    return np.asarray(x).T
