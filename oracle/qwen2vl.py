"""Oracle (TEST INFRASTRUCTURE ONLY): frame -> Qwen2-VL ``pixel_values`` / ``image_grid_thw`` on the CPU.

Restates, for the path BASELINE.json names:
  * ``smart_resize``            tf:models/qwen2_vl/image_processing_pil_qwen2_vl.py:57-83
  * resize / rescale / normalize / patchify of ``Qwen2VLImageProcessorPil._preprocess`` (same file :143-224)
    through ``resample_oracle.c``
  * the thumbnail size rule of ``Image.thumbnail`` (PIL:Image.py:2878-2893) used by
    ``_encode_image_optimized`` (src/agents/vlm_inspector.py:63-64, src/agents/vlm_auditor.py:90-91)
  * the size rule of ``resize_image`` (utils/image_utils.py:62-73)
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import lib

LANCZOS = 1
BICUBIC = 3

# tf:utils/constants.py:5-6 (OPENAI_CLIP_MEAN / OPENAI_CLIP_STD), tf:image_processing_utils.py:192 (1/255)
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
RESCALE = 1 / 255

PATCH = 14
MERGE = 2
TEMPORAL = 2
FACTOR = PATCH * MERGE
DEFAULT_MIN_PIXELS = 56 * 56            # tf:...pil_qwen2_vl.py:90  size["shortest_edge"]
DEFAULT_MAX_PIXELS = 28 * 28 * 1280     # size["longest_edge"]


def _u8p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def _f32p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i32p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def smart_resize(height: int, width: int, factor: int = FACTOR, min_pixels: int = DEFAULT_MIN_PIXELS,
                 max_pixels: int = DEFAULT_MAX_PIXELS) -> tuple[int, int]:
    """tf:models/qwen2_vl/image_processing_pil_qwen2_vl.py:57-83 (Python ``round`` = half-to-even)."""
    if max(height, width) / min(height, width) > 200:
        raise ValueError(
            f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}")
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def thumbnail_size(width: int, height: int, max_size: int) -> tuple[int, int] | None:
    """Size ``img.thumbnail((S, S))`` resizes to, or None when the image already fits (PIL:Image.py:2873-2893)."""
    x = y = math.floor(max_size)
    if x >= width and y >= height:
        return None

    def round_aspect(number, key):
        return max(min(math.floor(number), math.ceil(number), key=key), 1)

    aspect = width / height
    if x / y >= aspect:
        x = round_aspect(y * aspect, key=lambda n: abs(aspect - n / y))
    else:
        y = round_aspect(x / aspect, key=lambda n: 0 if n == 0 else abs(aspect - x / n))
    return x, y


def resize_image_size(width: int, height: int, max_dimension: int) -> tuple[int, int] | None:
    """Size rule of the reference ``resize_image`` (utils/image_utils.py:62-73); None = returned unchanged."""
    if width <= max_dimension and height <= max_dimension:
        return None
    if width > height:
        return max_dimension, int(height * (max_dimension / width))
    return int(width * (max_dimension / height)), max_dimension


def coeffs(in_size: int, out_size: int, filt: int):
    L = lib()
    ks = L.orc_ksize(in_size, out_size, filt)
    if ks < 0:
        raise ValueError("bad coefficient request")
    k = np.zeros((out_size, ks), np.int32)
    b = np.zeros((out_size, 2), np.int32)
    rc = L.orc_coeffs(in_size, out_size, filt, _i32p(k), _i32p(b))
    if rc < 0:
        raise RuntimeError("orc_coeffs failed")
    return k, b, ks


def resize(img: np.ndarray, out_h: int, out_w: int, filt: int = BICUBIC) -> np.ndarray:
    """``PIL.Image.fromarray(img).resize((out_w, out_h), filt, reducing_gap=None)`` for uint8 HWC input."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    h, w, ch = img.shape
    out = np.empty((out_h, out_w, ch), np.uint8)
    rc = lib().orc_resize(_u8p(img), h, w, ch, _u8p(out), out_h, out_w, filt)
    if rc:
        raise RuntimeError(f"orc_resize failed ({rc})")
    return out


def normalize_lut(mean=CLIP_MEAN, std=CLIP_STD, rescale: float = RESCALE) -> np.ndarray:
    m = np.asarray(mean, np.float32)
    s = np.asarray(std, np.float32)
    lut = np.empty(768, np.float32)
    lib().orc_lut(_f32p(m), _f32p(s), float(rescale), _f32p(lut))
    return lut


def patchify(resized: np.ndarray, lut: np.ndarray | None = None) -> np.ndarray:
    resized = np.ascontiguousarray(resized, dtype=np.uint8)
    h, w, c = resized.shape
    assert c == 3
    lut = normalize_lut() if lut is None else np.ascontiguousarray(lut, np.float32)
    out = np.empty(((h // PATCH) * (w // PATCH), 3 * TEMPORAL * PATCH * PATCH), np.float32)
    rc = lib().orc_patchify(_u8p(resized), h, w, _f32p(lut), _f32p(out))
    if rc:
        raise ValueError("patchify needs dimensions that are multiples of 28")
    return out


def preprocess(frames, min_pixels: int = DEFAULT_MIN_PIXELS, max_pixels: int = DEFAULT_MAX_PIXELS):
    """``Qwen2VLImageProcessorPil(min/max)(images=frames, return_tensors="np")`` for RGB uint8 HWC frames.

    Returns (pixel_values float32 [sum N_i, 1176], image_grid_thw int64 [B, 3]).
    """
    lut = normalize_lut()
    rows, grids = [], []
    for f in frames:
        f = np.asarray(f)
        h, w = f.shape[:2]
        rh, rw = smart_resize(h, w, FACTOR, min_pixels, max_pixels)
        r = resize(f, rh, rw, BICUBIC)
        rows.append(patchify(r, lut))
        grids.append((1, rh // PATCH, rw // PATCH))
    return np.concatenate(rows, axis=0), np.asarray(grids, np.int64)


def agent_thumbnail(frame: np.ndarray, max_size: int) -> np.ndarray:
    """Geometry half of ``_encode_image_optimized`` (src/agents/vlm_inspector.py:59-69): LANCZOS thumbnail to fit
    ``max_size`` (2048 Inspector, 1024 Auditor); the JPEG round trip is a codec step and is excluded."""
    h, w = frame.shape[:2]
    if max(w, h) <= max_size:
        return frame
    tw, th = thumbnail_size(w, h, max_size)
    if int(w / tw / 2.0) > 1 or int(h / th / 2.0) > 1:
        raise NotImplementedError("thumbnail box-reduce pre-pass (>= 4x downscale) is outside the oracle")
    return resize(frame, th, tw, LANCZOS)
