"""ctypes binding of libvis_b200.so — the C ABI declared in include/vis_b200.h.

There is no CPU fallback: if the library is missing the import of anything that needs it raises, and every
device entry point requires CUDA pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["VIS_B200_LIB"]) if os.environ.get("VIS_B200_LIB") else PKG_DIR / "libvis_b200.so"   # override: developer A/B builds

VIS_OK = 0
VIS_E_INVALID = -1
VIS_E_CUDA = -2
VIS_E_UNSUPPORTED = -3
VIS_E_CAPACITY = -4

FILTER_LANCZOS = 1
FILTER_BICUBIC = 3

LEAF_WORDS = 12

# struct layouts of include/vis_b200.h as numpy dtypes (C alignment)
FRAME_DTYPE = np.dtype([
    ("src", np.uint64), ("src_pitch", np.int64),
    ("src_h", np.int32), ("src_w", np.int32), ("dst_h", np.int32), ("dst_w", np.int32),
    ("hrec", np.uint64), ("vrec", np.uint64), ("row0", np.int64)], align=True)
STRIP_DTYPE = np.dtype([("frame", np.int32), ("x0", np.int32), ("x1", np.int32), ("y0", np.int32), ("y1", np.int32)],
                       align=True)
BOX_DTYPE = np.dtype([("x", np.int32), ("y", np.int32), ("w", np.int32), ("h", np.int32),
                      ("b", np.uint8), ("g", np.uint8), ("r", np.uint8), ("dashed", np.uint8),
                      ("label", np.uint64)], align=True)      # host pointer to NUL-terminated bytes (see HostRecords)
LEAF_DTYPE = np.dtype([("w", np.int32, (LEAF_WORDS,))], align=True)
OVERLAY_FRAME_DTYPE = np.dtype([
    ("src", np.uint64), ("dst", np.uint64), ("src_pitch", np.int64), ("dst_pitch", np.int64),
    ("h", np.int32), ("w", np.int32), ("group_begin", np.int32), ("group_end", np.int32)], align=True)

FRAME_REF_DTYPE = np.dtype([("src", np.uint64), ("row0", np.int64)], align=True)
SCHED_HEAD_DTYPE = np.dtype([("src_h", np.int32), ("src_w", np.int32), ("dst_h", np.int32), ("dst_w", np.int32),
                             ("src_pitch", np.int64), ("kt", np.int32), ("n_strips", np.int32), ("n_segs", np.int32),
                             ("stage_pitch", np.int32), ("max_strip_w", np.int32), ("per_index", np.int32), ("ring", np.int32),
                             ("n_subs", np.int32), ("out_mode", np.int32), ("h_pull", np.int32), ("n_vwarps", np.int32), ("dp_words", np.int32), ("mma_ks", np.int32), ("chunk_rows", np.int32)], align=True)
SCHED_OUT_PIXEL_VALUES, SCHED_OUT_U8 = 0, 1
SCHED_FLAG_DP4A = 0x100
SCHED_FLAG_MMA = 0x200
RESIZE_REF_DTYPE = np.dtype([("src", np.uint64), ("dst", np.uint64)], align=True)

assert FRAME_DTYPE.itemsize == 56 and STRIP_DTYPE.itemsize == 20 and BOX_DTYPE.itemsize == 32
OVERLAY_TILE_DTYPE = np.dtype([("frame", np.int32), ("txy", np.int32), ("ref_begin", np.int32), ("ref_end", np.int32)], align=True)
OVERLAY_REF_DTYPE = np.dtype([("leaf_begin", np.int32), ("leaf_end", np.int32)], align=True)
QUALITY_FRAME_DTYPE = np.dtype([("src", np.uint64), ("pitch", np.int64), ("h", np.int32), ("w", np.int32)], align=True)
assert LEAF_DTYPE.itemsize == 48 and OVERLAY_FRAME_DTYPE.itemsize == 48
SPRITE_DTYPE = np.dtype([("radius", np.int32), ("b", np.uint8), ("g", np.uint8), ("r", np.uint8), ("pad", np.uint8),
                         ("label", np.uint64), ("pixels", np.uint64), ("w", np.int32), ("h", np.int32), ("ox", np.int32),
                         ("oy", np.int32)], align=True)
assert SPRITE_DTYPE.itemsize == 40
REPITCH_DTYPE = np.dtype([("src", np.uint64), ("dst", np.uint64), ("src_pitch", np.int64), ("dst_pitch", np.int64),
                          ("rows", np.int32), ("row_bytes", np.int32)], align=True)
assert REPITCH_DTYPE.itemsize == 40
HP_U16LE, HP_U16BE, HP_I32, HP_F32 = 0, 1, 2, 3
HEAT_FRAME_DTYPE = np.dtype([("src", np.uint64), ("dst", np.uint64), ("src_pitch", np.int64), ("dst_pitch", np.int64),
                             ("h", np.int32), ("w", np.int32), ("plane_off", np.int64), ("final_ksize", np.int32),
                             ("final_koff", np.int32)], align=True)
assert HEAT_FRAME_DTYPE.itemsize == 56
RESIZE_COPY, RESIZE_AREA2, RESIZE_BILINEAR = 0, 1, 2
JPEG_BACKEND_DEFAULT, JPEG_BACKEND_HYBRID, JPEG_BACKEND_GPU_HYBRID, JPEG_BACKEND_HARDWARE = 0, 1, 2, 3
JPEG_CSS_444, JPEG_CSS_422, JPEG_CSS_420, JPEG_CSS_GRAY = 0, 1, 2, 6
PANEL_DTYPE = np.dtype([("src", np.uint64), ("src_pitch", np.int64), ("src_h", np.int32), ("src_w", np.int32),
                        ("dst_h", np.int32), ("dst_w", np.int32), ("org_x", np.int32), ("org_y", np.int32),
                        ("mode", np.int32), ("pad", np.int32), ("xofs", np.uint64), ("alpha", np.uint64),
                        ("yofs", np.uint64), ("beta", np.uint64)], align=True)
DRAW_LINE, DRAW_RECTANGLE, DRAW_CIRCLE, DRAW_TEXT = 1, 2, 3, 4
DRAW_CMD_DTYPE = np.dtype([("kind", np.int32), ("x1", np.int32), ("y1", np.int32), ("x2", np.int32), ("y2", np.int32),
                           ("thickness", np.int32), ("line_type", np.int32), ("color", np.uint8, (4,)),
                           ("font_scale", np.float64), ("text", np.uint64)], align=True)
PANEL_CANVAS_DTYPE = np.dtype([("canvas", np.uint64), ("pitch", np.int64), ("h", np.int32), ("w", np.int32),
                               ("fill", np.int32), ("n_panels", np.int32), ("panels", PANEL_DTYPE, (4,))], align=True)
assert PANEL_DTYPE.itemsize == 80 and DRAW_CMD_DTYPE.itemsize == 48 and PANEL_CANVAS_DTYPE.itemsize == 352

EXPORTS = [
    "vis_abi_version", "vis_last_error", "vis_source_hash", "vis_coeff_ksize", "vis_build_coeffs", "vis_build_lut",
    "vis_resample_h_u8", "vis_resample_v_u8", "vis_normalize_patchify",
    "vis_max_taps", "vis_fused_kt_class", "vis_record_stride", "vis_pack_records", "vis_fused_supported",
    "vis_plan_strips_max", "vis_plan_strips", "vis_preprocess_fused",
    "vis_sched_sizeof", "vis_sched_build", "vis_sched_pack_records", "vis_sched_record_stride_dp", "vis_sched_pack_records_dp", "vis_sched_record_stride_mma", "vis_sched_pack_records_mma", "vis_preprocess_fused_sched", "vis_preprocess_fused_sched_dup",
    "vis_resize_fused_sched",
    "vis_overlay_expand", "vis_overlay_tiles", "vis_overlay_plan_batch", "vis_overlay_draw", "vis_quality_stats", "vis_heatmap_batch",
    "vis_resize_linear_mode", "vis_linear_table", "vis_compose_panels", "vis_compose_panels_batch", "vis_text_size", "vis_draw_expand", "vis_overlay_draw_cn", "vis_overlay_sprite_expand", "vis_overlay_stamp_expand", "vis_overlay_plan_batch_sprites",
    "vis_coeff_ksize_box", "vis_build_coeffs_box", "vis_reduce_u8", "vis_nearest_table", "vis_gather_u8", "vis_alpha_premultiply_u8", "vis_repitch_u8", "vis_build_coeffs_f64", "vis_resample_hp",
    "vis_jpeg_create", "vis_jpeg_destroy", "vis_jpeg_info", "vis_jpeg_decode", "vis_jpeg_decode_batch",
    "vis_jpeg_encode_bound", "vis_jpeg_encode",
]


class HostRecords(np.ndarray):
    """Structured records that hold HOST pointers to strings (``VisBox.label``, ``VisSprite.label``,
    ``VisDrawCmd.text``): the array keeps the buffers it points into alive (``keep``), views and slices inherit
    them.  ``np.concatenate`` returns a plain array: keep the parts alive while the result is in use."""

    def __array_finalize__(self, obj):
        self.keep = getattr(obj, "keep", None)
        self.strings = None                 # the strings of a freshly built array, in record order (not kept by views)


def cv_text(text) -> bytes:
    """The bytes cv2.putText receives for a Python string: UTF-8, cut at the first NUL."""
    if isinstance(text, (bytes, bytearray)):
        b = bytes(text)
    else:
        b = str(text).encode("utf-8", "replace")
    return b.split(b"\0")[0] if b"\0" in b else b


# Short strings ('1' .. '99', the default panel labels) recur on every frame: their buffers are interned once (bounded),
# everything else gets a buffer owned by the array that points at it.
_INTERNED: dict = {}
_INTERN_MAX, _INTERN_LEN = 4096, 32


def _string_address(b: bytes, keep: list) -> int:
    hit = _INTERNED.get(b)
    if hit is not None:
        return hit[1]
    buf = C.create_string_buffer(b)
    addr = C.addressof(buf)
    if len(b) <= _INTERN_LEN and len(_INTERNED) < _INTERN_MAX:
        _INTERNED[b] = (buf, addr)
    else:
        keep.append(buf)
    return addr


def host_records(rows: list, dtype: np.dtype, string_field: str) -> "HostRecords":
    """Records from tuples whose ``string_field`` entry is ``bytes``: the strings are stored in ctypes buffers owned
    by the returned array (or interned) and the field receives their addresses."""
    k = dtype.names.index(string_field)
    keep: list = []
    fixed = [row[:k] + (_string_address(row[k], keep),) + row[k + 1:] for row in rows]
    out = (np.array(fixed, dtype) if fixed else np.zeros(0, dtype)).view(HostRecords)
    out.keep = keep
    out.strings = [row[k] for row in rows]
    return out


def record_string(rec, field: str) -> bytes:
    """The bytes a record's pointer field points to (``b""`` for a null pointer)."""
    addr = int(rec[field])
    return C.string_at(addr) if addr else b""


class VisError(RuntimeError):
    def __init__(self, code: int, where: str, text: str):
        super().__init__(f"{where} failed ({code}): {text}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """The loaded C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python {PKG_DIR / 'build.py'}` (nvcc, sm_100a). "
                "This engine has no CPU fallback.")
        L = C.CDLL(os.fspath(LIB_PATH))
        _declare(L)
        if L.vis_abi_version() != 20:
            raise RuntimeError("libvis_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def _declare(L: C.CDLL) -> None:
    i32p, f32p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_void_p
    ip = C.POINTER(C.c_int)
    L.vis_abi_version.restype = C.c_int
    L.vis_last_error.restype = C.c_char_p
    L.vis_source_hash.restype = C.c_char_p
    L.vis_coeff_ksize.argtypes = [C.c_int, C.c_int, C.c_int]
    L.vis_build_coeffs.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, ip]
    L.vis_build_lut.argtypes = [f32p, f32p, C.c_double, f32p]
    L.vis_resample_h_u8.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, C.c_int64, C.c_int,
                                    vp, vp, C.c_int, vp]
    L.vis_resample_v_u8.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp, C.c_int64, C.c_int, vp, vp, C.c_int, vp]
    L.vis_normalize_patchify.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp, vp, C.c_int64, vp]
    L.vis_max_taps.argtypes = [i32p, C.c_int]
    L.vis_fused_kt_class.argtypes = [C.c_int]
    L.vis_record_stride.argtypes = [C.c_int]
    L.vis_pack_records.argtypes = [C.c_int, i32p, i32p, C.c_int, C.c_int, i32p, C.c_int64]
    L.vis_fused_supported.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.vis_plan_strips_max.argtypes = [C.c_int, C.c_int]
    L.vis_plan_strips.argtypes = [C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.c_int, vp, C.c_int, ip, ip]
    L.vis_preprocess_fused.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.vis_sched_sizeof.argtypes = []
    L.vis_sched_build.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, i32p, i32p, C.c_int, C.c_int, vp]
    L.vis_resize_fused_sched.argtypes = [vp, vp, C.c_int, C.c_int64, vp, vp, vp]
    L.vis_sched_pack_records.argtypes = [C.c_int, i32p, i32p, C.c_int, C.c_int, C.c_int, i32p, C.c_int64]
    L.vis_sched_record_stride_dp.argtypes = [C.c_int]
    L.vis_sched_pack_records_dp.argtypes = [C.c_int, i32p, i32p, C.c_int, C.c_int, i32p, C.c_int64]
    L.vis_sched_record_stride_mma.argtypes = [C.c_int]
    L.vis_sched_pack_records_mma.argtypes = [C.c_int, i32p, i32p, C.c_int, C.c_int, i32p, C.c_int64]
    L.vis_preprocess_fused_sched.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.vis_preprocess_fused_sched_dup.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.vis_overlay_expand.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, ip]
    L.vis_overlay_tiles.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, ip, ip]
    L.vis_overlay_plan_batch.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int64, vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int]
    L.vis_overlay_draw.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp]
    L.vis_heatmap_batch.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, vp,
                                    vp, vp, vp, vp]
    L.vis_quality_stats.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.vis_resize_linear_mode.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.vis_linear_table.argtypes = [C.c_int, C.c_int, C.c_int, i32p, vp]
    L.vis_compose_panels.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
    L.vis_compose_panels_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.vis_text_size.argtypes = [C.c_char_p, C.c_double, C.c_int, ip, ip]
    L.vis_draw_expand.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, ip]
    L.vis_overlay_draw_cn.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp]
    L.vis_coeff_ksize_box.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int]
    L.vis_build_coeffs_box.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, i32p, i32p, ip]
    L.vis_reduce_u8.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, vp, C.c_int64, vp]
    L.vis_nearest_table.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, i32p]
    L.vis_gather_u8.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, C.c_int64, C.c_int, C.c_int, vp, vp, vp]
    L.vis_alpha_premultiply_u8.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.vis_repitch_u8.argtypes = [vp, C.c_int, C.c_int64, vp]
    L.vis_build_coeffs_f64.argtypes = [C.c_int, C.c_int, C.c_int, vp, i32p, ip]
    L.vis_resample_hp.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int64, C.c_int, vp, vp, C.c_int, vp]
    L.vis_overlay_sprite_expand.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, vp, C.c_int, ip, ip, ip, ip, ip]
    L.vis_overlay_stamp_expand.argtypes = [C.c_int, C.c_int, vp, C.c_int, ip, ip, ip, ip, ip]
    L.vis_overlay_plan_batch_sprites.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int64, vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int,
                                                 vp, C.c_int]
    L.vis_jpeg_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.vis_jpeg_destroy.argtypes = [vp]
    L.vis_jpeg_info.argtypes = [vp, vp, C.c_int64, ip, ip, ip, ip]
    L.vis_jpeg_decode.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp]
    L.vis_jpeg_decode_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp]
    L.vis_jpeg_encode_bound.argtypes = [C.c_int, C.c_int]
    L.vis_jpeg_encode.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int64,
                                  C.POINTER(C.c_int64), vp]
    for name in EXPORTS:
        if name in ("vis_jpeg_destroy",):
            getattr(L, name).restype = None
        elif name == "vis_jpeg_encode_bound":
            getattr(L, name).restype = C.c_int64
        elif name not in ("vis_last_error", "vis_source_hash"):
            getattr(L, name).restype = C.c_int


def check(code: int, where: str) -> int:
    if code < 0:
        raise VisError(code, where, lib().vis_last_error().decode("utf-8", "replace"))
    return code


def i32ptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def f32ptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))
