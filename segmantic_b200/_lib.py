"""ctypes binding of ``libsegmantic_b200.so`` (the C ABI declared in ``include/segmantic_b200.h``).

The library is built in-tree by ``segmantic_b200/csrc/build.sh`` (``__graft_entry__.build()``).
There is no CPU fallback: if the shared library is missing the import of any compute entry point
raises, and every compute call returns ``SGM_ERR_CUDA`` without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

SGM_MAX_LEVELS = 8
SGM_MAX_STARTS = 128

KIND_CONV, KIND_CONV_TRANSPOSE, KIND_IDENTITY = 0, 1, 2
PRECISION_FP32, PRECISION_BF16 = 0, 1

_LIB_PATH = Path(__file__).resolve().parent / "libsegmantic_b200.so"


class SgmError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("kernel", C.c_int32),
        ("stride", C.c_int32), ("has_act", C.c_int32), ("alpha", C.c_float),
        ("weight", C.POINTER(C.c_float)), ("bias", C.POINTER(C.c_float)),
    ]


class UnetDesc(C.Structure):
    _fields_ = [
        ("spatial_dims", C.c_int32), ("in_channels", C.c_int32), ("out_channels", C.c_int32),
        ("n_levels", C.c_int32), ("channels", C.c_int32 * SGM_MAX_LEVELS),
        ("strides", C.c_int32 * SGM_MAX_LEVELS), ("precision", C.c_int32), ("n_convs", C.c_int32),
        ("convs", C.POINTER(ConvDesc)),
    ]


class SwCfg(C.Structure):
    _fields_ = [
        ("dims", C.c_int32 * 3), ("roi", C.c_int32 * 3), ("n_starts", C.c_int32 * 3),
        ("starts", (C.c_int32 * SGM_MAX_STARTS) * 3), ("sw_batch", C.c_int32),
        ("imap", C.POINTER(C.c_float) * 3), ("imap_floor", C.c_float),
        ("a0_begin", C.c_int32), ("a0_end", C.c_int32), ("vol_x0", C.c_int32), ("vol_nx", C.c_int32),
        ("acc_x0", C.c_int32), ("acc_nx", C.c_int32),
    ]


# name -> (restype, argtypes); mirrors include/segmantic_b200.h one to one
_I3 = C.POINTER(C.c_int32)
_D = C.POINTER(C.c_double)
SIGNATURES = {
    "sgm_last_error": (C.c_char_p, []),
    "sgm_version": (C.c_int32, []),
    "sgm_unet_create": (C.c_int32, [C.POINTER(UnetDesc), C.POINTER(C.c_void_p)]),
    "sgm_unet_destroy": (None, [C.c_void_p]),
    "sgm_unet_workspace_bytes": (C.c_int64, [C.c_void_p, _I3, C.c_int32]),
    "sgm_unet_forward": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, _I3, C.c_void_p,
                                     C.c_int64, C.c_void_p]),
    "sgm_unet_last_launch_count": (C.c_int64, [C.c_void_p]),
    "sgm_unet_check": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "sgm_unet_set_profiling": (C.c_int32, [C.c_void_p, C.c_int32]),
    "sgm_unet_get_profile": (C.c_int32, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32,
                                         C.c_void_p]),
    "sgm_debug_conv": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, _I3,
                                   _I3, C.c_void_p]),
    "sgm_sw_accumulate": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(SwCfg), C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_void_p]),
    "sgm_sw_workspace_bytes": (C.c_int64, [C.c_void_p, C.POINTER(SwCfg)]),
    "sgm_sw_windows_workspace_bytes": (C.c_int64, [C.c_void_p, C.POINTER(SwCfg)]),
    "sgm_sw_windows": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(SwCfg), C.c_int64, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p]),
    "sgm_sw_blend": (C.c_int32, [C.POINTER(SwCfg), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "sgm_sw_predict_workspace_bytes": (C.c_int64, [C.c_void_p, C.POINTER(SwCfg)]),
    "sgm_sw_predict": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(SwCfg), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p]),
    "sgm_sw_finalize": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(SwCfg), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "sgm_resample_trilinear": (C.c_int32, [C.c_void_p, _I3, C.c_int32, C.c_void_p, _I3, _D, C.c_void_p]),
    "sgm_resample_trilinear_argmax": (C.c_int32, [C.c_void_p, _I3, C.c_int32, C.c_void_p, _I3, _D,
                                                  C.c_void_p]),
    "sgm_resample_itk": (C.c_int32, [C.c_void_p, C.c_int32, _I3, C.c_void_p, _I3, _D, _D, _D, _D,
                                     C.c_int32, C.c_double, C.c_void_p]),
    "sgm_normalize_intensity": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                            C.c_void_p]),
    "sgm_foreground_bbox": (C.c_int32, [C.c_void_p, C.c_int32, _I3, C.c_void_p, C.c_void_p]),
    "sgm_ensemble_mean_argmax": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_float), C.c_void_p,
                                             C.c_void_p, C.c_void_p]),
    "sgm_ensemble_vote": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "sgm_ensemble_select_best": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                             C.c_int32, C.c_void_p, C.c_void_p]),
    "sgm_p2p_export": (C.c_int32, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "sgm_p2p_open": (C.c_int32, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "sgm_p2p_close": (C.c_int32, [C.c_void_p]),
    "sgm_p2p_put": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "sgm_p2p_signal": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "sgm_p2p_wait": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_double, C.c_void_p]),
    "sgm_confusion_matrix": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
}

_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("SEGMANTIC_B200_LIB", _LIB_PATH))


def load() -> C.CDLL:
    """Load the shared library (once) and bind every declared symbol; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise SgmError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or segmantic_b200/csrc/build.sh). segmantic_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = load().sgm_last_error()
        raise SgmError(f"{what or 'segmantic_b200'} failed ({rc}): {msg.decode() if msg else '?'}")
    return rc


def i3(vals) -> C.Array:
    return (C.c_int32 * 3)(*[int(v) for v in vals])
