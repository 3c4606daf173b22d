"""CPU oracle — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this package; the product path (``vision-inspection-system_b200``) never does and fails loudly
when its CUDA library is missing.

Contents
    resample_oracle.c   Pillow 8bpc resampler + Qwen2-VL normalize/patchify, restated in plain C
    cvdraw_oracle.c     OpenCV 4.13 drawing primitives used by draw_bounding_boxes, restated in plain C
    cvresize_oracle.c   OpenCV 4.13 cv2.resize (INTER_LINEAR, 8-bit) used by create_side_by_side_comparison
    qwen2vl.py          smart_resize / thumbnail size rules / whole-frame preprocess (numpy + the C code)
    overlay.py          draw_bounding_boxes box logic on top of cvdraw_oracle.c

Parity pin: the reference has no tests or golden vectors for this path, so the
oracle is pinned against the installed third-party binaries the reference calls (Pillow 12.2.0,
opencv-python-headless 4.13.0.92, transformers 5.5.0) and against ``tests/golden/`` fixtures generated
in the build container from the real reference (``tests/golden/make_goldens.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the C restatements with gcc (oracle/Makefile)."""
    srcs = [_DIR / "resample_oracle.c", _DIR / "cvdraw_oracle.c", _DIR / "cvresize_oracle.c"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(_DIR)] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (building when the sources are newer) the oracle shared object."""
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not _LIB_PATH.exists():
                raise
        _lib = ctypes.CDLL(os.fspath(_LIB_PATH))
        _declare(_lib)
    return _lib


def _declare(L: ctypes.CDLL) -> None:
    c = ctypes
    u8p, i32p, f32p = c.POINTER(c.c_uint8), c.POINTER(c.c_int32), c.POINTER(c.c_float)
    L.orc_ksize.argtypes = [c.c_int, c.c_int, c.c_int]
    L.orc_ksize.restype = c.c_int
    L.orc_coeffs.argtypes = [c.c_int, c.c_int, c.c_int, i32p, i32p]
    L.orc_coeffs.restype = c.c_int
    L.orc_resize.argtypes = [u8p, c.c_int, c.c_int, c.c_int, u8p, c.c_int, c.c_int, c.c_int]
    L.orc_resize.restype = c.c_int
    L.orc_ksize_box.argtypes = [c.c_float, c.c_float, c.c_int, c.c_int]
    L.orc_ksize_box.restype = c.c_int
    L.orc_coeffs_box.argtypes = [c.c_int, c.c_float, c.c_float, c.c_int, c.c_int, i32p, i32p]
    L.orc_coeffs_box.restype = c.c_int
    L.orc_resize_box.argtypes = [u8p, c.c_int, c.c_int, c.c_int, u8p, c.c_int, c.c_int, c.c_int, f32p]
    L.orc_resize_box.restype = c.c_int
    L.orc_reduce.argtypes = [u8p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, i32p, u8p]
    L.orc_reduce.restype = c.c_int
    L.orc_nearest_table.argtypes = [c.c_int, c.c_float, c.c_float, c.c_int, i32p]
    L.orc_nearest_table.restype = c.c_int
    L.orc_lut.argtypes = [f32p, f32p, c.c_double, f32p]
    L.orc_lut.restype = None
    L.orc_patchify.argtypes = [u8p, c.c_int, c.c_int, f32p, f32p]
    L.orc_patchify.restype = c.c_int
    # drawing oracle (cvdraw_oracle.c)
    L.ocv_rectangle.argtypes = [u8p, c.c_int, c.c_int, c.c_int64, c.c_int, c.c_int, c.c_int, c.c_int,
                                c.c_int, c.c_int, c.c_int, c.c_int, c.c_int]
    L.ocv_rectangle.restype = None
    L.ocv_line.argtypes = L.ocv_rectangle.argtypes
    L.ocv_line.restype = None
    L.ocv_circle.argtypes = [u8p, c.c_int, c.c_int, c.c_int64, c.c_int, c.c_int, c.c_int,
                             c.c_int, c.c_int, c.c_int, c.c_int, c.c_int]
    L.ocv_circle.restype = None
    L.ocv_put_text.argtypes = [u8p, c.c_int, c.c_int, c.c_int64, c.c_char_p, c.c_int, c.c_int, c.c_double,
                               c.c_int, c.c_int, c.c_int, c.c_int]
    L.ocv_put_text.restype = c.c_int
    L.ocv_get_text_size.argtypes = [c.c_char_p, c.c_double, c.c_int, c.POINTER(c.c_int), c.POINTER(c.c_int)]
    L.ocv_get_text_size.restype = c.c_int
    # cv2.resize oracle (cvresize_oracle.c)
    L.ocv_resize_linear_mode.argtypes = [c.c_int, c.c_int, c.c_int, c.c_int]
    L.ocv_resize_linear_mode.restype = c.c_int
    L.ocv_resize_linear_u8.argtypes = [u8p, c.c_int, c.c_int, c.c_int64, c.c_int, u8p, c.c_int, c.c_int, c.c_int64]
    L.ocv_resize_linear_u8.restype = c.c_int
