"""Host-side tables for the kernels, computed by the [host] entry points of libvis_b200.so (no GPU needed).

Resampling coefficients (Pillow ``precompute_coeffs`` / ``normalize_coeffs_8bpc`` semantics), the 768-entry
normalisation table, push-order records and strip plans for the fused kernel.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

from . import _native as N
from .geometry import IMAGE_MEAN, IMAGE_STD, RESCALE_FACTOR


@dataclass(frozen=True)
class CoeffTable:
    in_size: int
    out_size: int
    filt: int
    ksize: int
    k: np.ndarray          # int32 [out_size, ksize]
    bounds: np.ndarray     # int32 [out_size, 2]  (first input index, tap count)
    max_taps: int


@lru_cache(maxsize=256)
def coeff_table(in_size: int, out_size: int, filt: int) -> CoeffTable:
    L = N.lib()
    ksize = N.check(L.vis_coeff_ksize(in_size, out_size, filt), "vis_coeff_ksize")
    k = np.zeros((out_size, ksize), np.int32)
    b = np.zeros((out_size, 2), np.int32)
    ks = C.c_int(0)
    N.check(L.vis_build_coeffs(in_size, out_size, filt, N.i32ptr(k), N.i32ptr(b), C.byref(ks)), "vis_build_coeffs")
    taps = N.check(L.vis_max_taps(N.i32ptr(b), out_size), "vis_max_taps")
    k.setflags(write=False)
    b.setflags(write=False)
    return CoeffTable(in_size, out_size, filt, ksize, k, b, taps)


@lru_cache(maxsize=64)
def coeff_table_box(in_size: int, in0: float, in1: float, out_size: int, filt: int) -> CoeffTable:
    """Same table for the fractional source box [in0, in1) of the axis (floats, as ImagingResample's box)."""
    L = N.lib()
    ksize = N.check(L.vis_coeff_ksize_box(in0, in1, out_size, filt), "vis_coeff_ksize_box")
    k = np.zeros((out_size, ksize), np.int32)
    b = np.zeros((out_size, 2), np.int32)
    ks = C.c_int(0)
    N.check(L.vis_build_coeffs_box(in_size, in0, in1, out_size, filt, N.i32ptr(k), N.i32ptr(b), C.byref(ks)),
            "vis_build_coeffs_box")
    taps = N.check(L.vis_max_taps(N.i32ptr(b), out_size), "vis_max_taps")
    return CoeffTable(in_size, out_size, filt, ksize, k, b, taps)


@lru_cache(maxsize=64)
def coeff_table_f64(in_size: int, out_size: int, filt: int):
    """(k float64 [out_size, ksize], bounds int32 [out_size, 2]) — Pillow's double-precision weights, for the modes it
    does not resample as 8 bits per channel ("I;16", "I", "F")."""
    L = N.lib()
    ksize = N.check(L.vis_coeff_ksize(in_size, out_size, filt), "vis_coeff_ksize")
    k = np.zeros((out_size, ksize), np.float64)
    b = np.zeros((out_size, 2), np.int32)
    N.check(L.vis_build_coeffs_f64(in_size, out_size, filt, k.ctypes.data_as(C.c_void_p), N.i32ptr(b), None),
            "vis_build_coeffs_f64")
    return k, b


def normalize_lut(mean=IMAGE_MEAN, std=IMAGE_STD, rescale: float = RESCALE_FACTOR) -> np.ndarray:
    m = np.asarray(mean, np.float32)
    s = np.asarray(std, np.float32)
    lut = np.empty(768, np.float32)
    N.check(N.lib().vis_build_lut(N.f32ptr(m), N.f32ptr(s), float(rescale), N.f32ptr(lut)), "vis_build_lut")
    return lut


def kt_class(taps: int) -> int:
    """Tap class of the fused kernel (6/8/12/16) or 0 when the geometry needs the generic passes."""
    return N.lib().vis_fused_kt_class(int(taps))


def pack_records(table: CoeffTable, kt: int) -> np.ndarray:
    """Push-order records ``[out_size + 1, stride]`` int32 for the fused kernel."""
    L = N.lib()
    stride = N.check(L.vis_record_stride(kt), "vis_record_stride")
    rec = np.zeros((table.out_size + 1, stride), np.int32)
    N.check(L.vis_pack_records(table.out_size, N.i32ptr(table.k), N.i32ptr(table.bounds), table.ksize, kt,
                               N.i32ptr(rec), rec.size), "vis_pack_records")
    return rec


@dataclass(frozen=True)
class StripPlan:
    strips: np.ndarray     # STRIP_DTYPE, frame index 0
    span_bytes: int
    strip_w: int


def plan_strips(dst_h: int, dst_w: int, htable: CoeffTable, kt: int, vsplit: int = 1) -> StripPlan:
    L = N.lib()
    cap = N.check(L.vis_plan_strips_max(dst_h, dst_w), "vis_plan_strips_max")
    strips = np.zeros(cap, N.STRIP_DTYPE)
    span, sw = C.c_int(0), C.c_int(0)
    n = N.check(L.vis_plan_strips(0, dst_h, dst_w, N.i32ptr(htable.bounds), kt, vsplit,
                                  strips.ctypes.data_as(C.c_void_p), cap, C.byref(span), C.byref(sw)),
                "vis_plan_strips")
    return StripPlan(strips[:n].copy(), span.value, sw.value)
