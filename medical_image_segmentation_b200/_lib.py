"""ctypes binding of libmis_b200.so (the C ABI declared in include/mis_b200.h).

There is NO fallback: if the library is missing the import of any product module fails with
an explicit message (build it with ``python -m medical_image_segmentation_b200._build`` or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libmis_b200.so")

MIS_OK, MIS_ERR_INVALID_ARG, MIS_ERR_UNSUPPORTED, MIS_ERR_CUDA = 0, 1, 2, 3
MIS_DTYPE_BF16, MIS_DTYPE_F32 = 0, 1
MIS_VIEW_FLIP, MIS_VIEW_JITTER, MIS_VIEW_GRAY, MIS_VIEW_BLUR, MIS_VIEW_SOLARIZE = 1, 2, 4, 8, 16

# numpy mirror of struct MisViewParams (48 bytes)
VIEW_PARAMS_DTYPE = np.dtype([
    ("img", "<i4"), ("top", "<i4"), ("left", "<i4"), ("h", "<i4"), ("w", "<i4"), ("flags", "<u4"),
    ("order", "u1", (4,)), ("brightness", "<f4"), ("contrast", "<f4"), ("saturation", "<f4"), ("hue", "<f4"),
    ("blur_sigma", "<f4"),
], align=False)
assert VIEW_PARAMS_DTYPE.itemsize == 48

# numpy mirror of struct MisEmaEntry (32 bytes)
EMA_ENTRY_DTYPE = np.dtype([("online", "<u8"), ("momentum", "<u8"), ("n", "<i8"), ("chunk0", "<i8")], align=False)

EXPORTS = {
    "mis_version": (C.c_int, []),
    "mis_last_error": (C.c_char_p, []),
    "mis_draw_two_view_params": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mis_draw_two_view_params_cb": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mis_params_to_view_major": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mis_view_params_check": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_int)]),
    "mis_draw_resize_jitter_params": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                C.c_float, C.c_void_p]),
    "mis_aug_two_view": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                   C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p]),
    "mis_aug_two_view_ordered": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int, C.c_void_p]),
    "mis_aug_two_view_staged": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int), C.c_void_p]),
    "mis_view_cost_order": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mis_aug_kernel_variant": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int]),
    "mis_aug_blur_views": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "mis_h2d_needed_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                      C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "mis_aug_algorithmic_bytes": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mis_ntxent_padded_rows": (C.c_int, [C.c_int]),
    "mis_ntxent_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "mis_ntxent_prep": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mis_ntxent_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int64, C.c_void_p]),
    "mis_ntxent_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p]),
    "mis_ntxent_fwd_peer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p]),
    "mis_ntxent_bwd_peer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_double, C.c_void_p, C.c_int64, C.c_void_p]),
    "mis_ntxent_fwd_bwd_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "mis_ntxent_fwd_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    "mis_ema_chunks": (C.c_int64, [C.c_int64]),
    "mis_ema_update": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "mis_knn_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "mis_knn_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "mis_u16_moments": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]),
    "mis_byol_loss_fwd_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
}


class MisError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. This package has no CPU/PyTorch fallback: build the CUDA library first "
            "(python -m medical_image_segmentation_b200._build, or __graft_entry__.build()).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError here == the .so does not export the declared ABI
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(code: int, what: str) -> None:
    """Map ABI return codes to the exceptions torchvision raises for the same mistakes."""
    if code == MIS_OK:
        return
    msg = lib.mis_last_error().decode("utf-8", "replace")
    if code == MIS_ERR_INVALID_ARG:
        raise ValueError(f"{what}: {msg}")
    if code == MIS_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    raise MisError(f"{what}: {msg}")
