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


def reduce(img: np.ndarray, factor, box=None) -> np.ndarray:
    """``PIL.Image.fromarray(img).reduce(factor, box)`` for uint8 HWC input (libImaging/Reduce.c)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, ch = img.shape
    fx, fy = factor if isinstance(factor, (tuple, list)) else (factor, factor)
    box = np.asarray((0, 0, w, h) if box is None else box, np.int32)
    out = np.empty((-(-(int(box[3]) - int(box[1])) // fy), -(-(int(box[2]) - int(box[0])) // fx), ch), np.uint8)
    rc = lib().orc_reduce(_u8p(img), h, w, ch, fx, fy, _i32p(box), _u8p(out))
    if rc:
        raise RuntimeError(f"orc_reduce failed ({rc})")
    return out


def resize_box(img: np.ndarray, out_h: int, out_w: int, filt: int, box) -> np.ndarray:
    """``Image.resize((out_w, out_h), filt, box=box, reducing_gap=None)`` (the core ImagingResample with a float box)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, ch = img.shape
    if h > w * 100 and out_h < h:                           # tall image: vertical pass first (PIL:Image.py:2431-2435)
        tmp = resize_box(img, out_h, w, filt, (0, box[1], w, box[3]))
        return resize_box(tmp, out_h, out_w, filt, (box[0], 0, box[2], out_h))
    out = np.empty((out_h, out_w, ch), np.uint8)
    fbox = np.asarray(box, np.float32)                      # ImagingResample takes float box[4]
    rc = lib().orc_resize_box(_u8p(img), h, w, ch, _u8p(out), out_h, out_w, filt, _f32p(fbox))
    if rc:
        raise RuntimeError(f"orc_resize_box failed ({rc})")
    return out


def resize_nearest(img: np.ndarray, out_h: int, out_w: int, box=None) -> np.ndarray:
    """``Image.resize((out_w, out_h), NEAREST, box)`` — the path Pillow takes for "P" and "1" images (uint8 HW or HWC)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    box = np.asarray((0, 0, w, h) if box is None else box, np.float32)
    xt, yt = np.empty(out_w, np.int32), np.empty(out_h, np.int32)
    lib().orc_nearest_table(w, float(box[0]), float(box[2]), out_w, _i32p(xt))
    lib().orc_nearest_table(h, float(box[1]), float(box[3]), out_h, _i32p(yt))
    out = img[np.clip(yt, 0, h - 1)][:, np.clip(xt, 0, w - 1)].copy()
    out[yt < 0] = 0
    out[:, xt < 0] = 0
    return out


_FILTER_SUPPORT = {BICUBIC: 2.0, LANCZOS: 3.0}


def reducing_plan(width: int, height: int, out_w: int, out_h: int, filt: int, box=None, reducing_gap: float = 2.0):
    """The pre-pass ``Image.resize(..., reducing_gap)`` inserts (PIL:Image.py:2407-2424): returns None when no
    reduction applies, else ((factor_x, factor_y), reduce_box ints, box floats relative to the reduced image)."""
    box = (0, 0, width, height) if box is None else box
    factor_x = int((box[2] - box[0]) / out_w / reducing_gap) or 1
    factor_y = int((box[3] - box[1]) / out_h / reducing_gap) or 1
    if factor_x <= 1 and factor_y <= 1:
        return None
    support = _FILTER_SUPPORT[filt] - 0.5                               # Image._get_safe_box
    sx, sy = support * (box[2] - box[0]) / out_w, support * (box[3] - box[1]) / out_h
    rb = (max(0, int(box[0] - sx)), max(0, int(box[1] - sy)),
          min(width, math.ceil(box[2] + sx)), min(height, math.ceil(box[3] + sy)))
    new_box = ((box[0] - rb[0]) / factor_x, (box[1] - rb[1]) / factor_y,
               (box[2] - rb[0]) / factor_x, (box[3] - rb[1]) / factor_y)
    return (factor_x, factor_y), rb, new_box


def resize_reducing(img: np.ndarray, out_h: int, out_w: int, filt: int, box=None, reducing_gap: float = 2.0) -> np.ndarray:
    """``Image.resize((out_w, out_h), filt, box, reducing_gap)`` for uint8 HWC input: reduce + boxed resample."""
    h, w = img.shape[:2]
    plan = reducing_plan(w, h, out_w, out_h, filt, box, reducing_gap)
    if plan is None:
        return resize(img, out_h, out_w, filt) if box is None else resize_box(img, out_h, out_w, filt, box)
    factor, rb, new_box = plan
    return resize_box(reduce(img, factor, rb), out_h, out_w, filt, new_box)


def agent_thumbnail(frame: np.ndarray, max_size: int, box=None) -> np.ndarray:
    """Geometry half of ``_encode_image_optimized`` (src/agents/vlm_inspector.py:59-69): ``thumbnail((S, S), LANCZOS)``
    with PIL's default ``reducing_gap=2.0`` (box-reduce pre-pass from 4x downscales on; ``box`` = what a JPEG draft
    decode hands over); the JPEG round trip is a codec step and is excluded."""
    h, w = frame.shape[:2]
    if max(w, h) <= max_size:
        return frame
    tw, th = thumbnail_size(w, h, max_size)
    return resize_reducing(frame, th, tw, LANCZOS, box)


def resize_hp(arr: np.ndarray, out_h: int, out_w: int, filt: int = LANCZOS) -> np.ndarray:
    """Test infrastructure.  ``Image.resize((out_w, out_h), filt)`` for the single-channel modes Pillow resamples in
    double precision — uint16 ("I;16"), int32 ("I"), float32 ("F") arrays — restating libImaging/Resample.c
    ImagingResampleHorizontal/Vertical_16bpc and _32bpc: normalised double weights (precompute_coeffs without the
    fixed-point conversion), ss += pixel * k in tap order, ROUND_UP (x86 cvttsd2si: INT_MIN when out of range), and for
    the 16-bit path Pillow's two CLIP8 byte writes (low = CLIP8(ss % 256) with C remainder, high = CLIP8(ss >> 8))."""
    import math

    def table(in_size, out_size):
        support0 = 3.0 if filt == LANCZOS else 2.0
        scale = in_size / out_size
        fs = max(scale, 1.0)
        support = support0 * fs
        rows = []
        for o in range(out_size):
            center = (o + 0.5) * scale
            first = max(int(center - support + 0.5), 0)
            last = min(int(center + support + 0.5), in_size)
            w = []
            for t in range(last - first):
                x = (t + first - center + 0.5) * (1.0 / fs)
                if filt == LANCZOS:
                    def sinc(v):
                        if v == 0.0:
                            return 1.0
                        v = v * math.pi
                        return math.sin(v) / v
                    w.append(sinc(x) * sinc(x / 3) if -3.0 <= x < 3.0 else 0.0)
                else:
                    a, x = -0.5, abs(x)
                    w.append(((a + 2.0) * x - (a + 3.0)) * x * x + 1 if x < 1.0 else
                             (((x - 5) * x + 8) * x - 4) * a if x < 2.0 else 0.0)
            total = 0.0
            for v in w:
                total += v
            rows.append((first, [v / total if total != 0.0 else v for v in w]))
        return rows

    def round_up(ss):
        v = np.where(ss >= 0.0, ss + 0.5, ss - 0.5)
        ok = (v > -2147483649.0) & (v < 2147483648.0)
        return np.where(ok, np.trunc(np.where(ok, v, 0.0)), -2147483648.0).astype(np.int64)

    def one_pass(a, out_size, axis):
        a = np.moveaxis(a, axis, 0)
        out = np.empty((out_size,) + a.shape[1:], a.dtype)
        for o, (first, w) in enumerate(table(a.shape[0], out_size)):
            ss = np.zeros(a.shape[1:], np.float64)
            for t, k in enumerate(w):
                ss = ss + a[first + t].astype(np.float64) * k
            if a.dtype == np.uint16:
                r = round_up(ss)
                lo = np.clip(np.fmod(r, 256), 0, 255)            # C remainder keeps the sign of r
                hi = np.clip(r >> 8, 0, 255)
                out[o] = (lo + 256 * hi).astype(np.uint16)
            elif a.dtype == np.int32:
                out[o] = round_up(ss).astype(np.int32)
            else:
                out[o] = ss.astype(np.float32)
        return np.moveaxis(out, 0, axis)

    h, w = arr.shape
    cur = arr
    need_h, need_v = out_w != w, out_h != h                    # PIL:Image.py:2431-2435: tall frames go vertical first
    order = ("vh" if (h > w * 100 and out_h < h) else "hv") if need_h and need_v else ("h" if need_h else "") + ("v" if need_v else "")
    for axis in order:
        cur = one_pass(cur, out_w, 1) if axis == "h" else one_pass(cur, out_h, 0)
    return cur.copy()
