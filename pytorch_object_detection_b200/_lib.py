"""ctypes binding of libb200det.so — the only way this package reaches the GPU.

The prototypes below are a 1:1 transcription of ``include/b200det.h``.  There is no CPU
implementation behind them: if the library is missing it is built with nvcc (``build.py``),
and if that fails the import raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

from . import build as _build

MAX_LEVELS = 8
MAX_BOX = 8192
ABI_VERSION = 6


class Level(C.Structure):
    """``b200det_level``."""
    _fields_ = [("cls", C.c_void_p), ("cnt", C.c_void_p), ("reg", C.c_void_p),
                ("h", C.c_int32), ("w", C.c_int32), ("stride", C.c_int32), ("dtypes", C.c_int32),
                ("reg_scale", C.c_void_p)]


_P = C.c_void_p
_LV = C.POINTER(Level)
PROTOTYPES = {
    "b200det_abi_version": (C.c_int, []),
    "b200det_status_string": (C.c_char_p, [C.c_int]),
    "b200det_last_cuda_error": (C.c_char_p, []),
    "b200det_score_points": (C.c_int, [_LV, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "b200det_select_topk": (C.c_int, [_LV, C.c_int, C.c_int, _P, _P, C.c_float, C.c_int, _P, _P, _P, _P, _P, _P]),
    "b200det_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "b200det_batched_nms": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, C.c_float, C.c_double, C.c_int, C.c_int,
                                      _P, C.c_size_t, _P, _P, _P, _P, _P, _P]),
    "b200det_postprocess_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "b200det_postprocess": (C.c_int, [_LV, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double, C.c_int, C.c_int,
                                      C.c_int, _P, C.c_size_t, _P, _P, _P, _P, _P, _P]),
    "b200det_clip_boxes": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P]),
    "b200det_assign_targets": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "b200det_box_loss_fwd": (C.c_int, [_LV, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "b200det_box_loss_bwd": (C.c_int, [_LV, _P, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "b200det_cnt_loss_fwd": (C.c_int, [_LV, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "b200det_cnt_loss_bwd": (C.c_int, [_LV, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "b200det_cls_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "b200det_cls_loss_fwd": (C.c_int, [_LV, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P, _P, _P]),
    "b200det_cls_loss_bwd": (C.c_int, [_LV, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "b200det_cls_loss_step": (C.c_int, [_LV, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P,
                                        C.c_size_t,
                                        _P, _P, _P, _P]),
    "b200det_assign_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "b200det_assign_loss_fused": (C.c_int, [_LV, _P, _P, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P, _P, C.c_int,
                                            _P, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "b200det_scale_maps": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "b200det_rescale_maps": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P]),
    "b200det_eval_ap_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "b200det_eval_ap": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.c_double, _P, C.c_size_t,
                                  _P, _P]),
    "b200det_coco_boxes": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, C.c_float, _P, _P, _P]),
    "b200det_pack_gt": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "b200det_collate_images": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
}

_lib = None


class B200DetError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if the in-tree .so is missing or stale).  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("B200DET_LIB") or _build.LIB      # (B200DET_LIB: a variant build for A/B timing / tracing)
    if not os.path.exists(path):     # staleness is handled by __graft_entry__.build(), not at import
        try:
            _build.build()
        except Exception as e:
            raise B200DetError(f"libb200det.so is missing and could not be built: {e}") from e
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.b200det_abi_version() != ABI_VERSION:
        raise B200DetError(f"libb200det.so ABI {lib.b200det_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        lib = load()
        msg = lib.b200det_status_string(status).decode()
        if status == 4:
            msg += ": " + lib.b200det_last_cuda_error().decode()
        raise B200DetError(f"{what}: {msg}")


def make_levels(entries: Sequence[tuple], dtypes: int = 0) -> C.Array:
    """entries: (cls_ptr, cnt_ptr, reg_ptr, h, w, stride[, reg_scale_ptr]) per level; pointers may be 0/None.
    ``dtypes``: B200DET_LEVEL_DTYPES(cls_cnt, reg) = cls_cnt | reg << 4 (0 = fp32 maps)."""
    if not 0 < len(entries) <= MAX_LEVELS:
        raise B200DetError(f"between 1 and {MAX_LEVELS} levels are supported, got {len(entries)}")
    arr = (Level * len(entries))()
    for i, e in enumerate(entries):
        a, b, c, h, w, s = e[:6]
        arr[i] = Level(a or None, b or None, c or None, h, w, s, int(dtypes), (e[6] if len(e) > 6 else None) or None)
    return arr
