"""Profiling driver for ncu: the three resample kernels on the shapes of BASELINE configs[2] (512x512x120 @ 0.5x0.5x3 mm
<-> 256x256x358 @ 1 mm), `reps` launches each:  python tests/prof_resample.py [reps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmantic_b200 import _lib  # noqa: E402
from segmantic_b200.seg import transforms as T  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
src_shape, dst_shape = (512, 512, 120), (256, 256, 358)
aff = np.diag([0.5, 0.5, 3.0, 1.0])
g = torch.Generator().manual_seed(3)
new_aff = T.zoom_affine(aff, (1.0, 1.0, 1.0))
shp, offset = T.compute_shape_offset(src_shape, aff, new_aff)
new_aff[:3, -1] = offset
xf, xinv = np.linalg.solve(aff, new_aff), np.linalg.solve(new_aff, aff)
img = torch.randn((1,) + src_shape, generator=g).to(dev)
logits = torch.randn((10,) + dst_shape, generator=g).to(dev)
lab = torch.randint(0, 10, tuple(reversed(dst_shape)), generator=g, dtype=torch.uint8).to(dev)
lout = torch.empty(tuple(reversed(src_shape)), dtype=torch.uint8, device=dev)
lib = _lib.load()


def dbl(a):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    return (C.c_double * a.size)(*a.tolist())


for _ in range(reps):
    a = T.resample_index_affine(img, xf, dst_shape)
    b = T.resample_index_affine_argmax(logits, xinv, src_shape)
    with torch.cuda.device(dev):
        _lib.check(lib.sgm_resample_itk(lab.data_ptr(), 0, _lib.i3(dst_shape), lout.data_ptr(), _lib.i3(src_shape),
                                        dbl(np.diag([0.5, 0.5, 3.0])), dbl(np.zeros(3)), dbl(np.eye(3)), dbl(np.zeros(3)), 1, 0.0,
                                        int(torch.cuda.current_stream(dev).cuda_stream)), "sgm_resample_itk")
torch.cuda.synchronize()
print("ok", float(a.mean()), int(b.sum()), int(lout.sum()))
