"""Host-side size rules of the hot path (pure Python, as in the reference; no GPU work here).

* ``smart_resize``       — Qwen2-VL target size: tf:models/qwen2_vl/image_processing_pil_qwen2_vl.py:57-83
* ``thumbnail_size``     — ``Image.thumbnail`` aspect rule (PIL:Image.py:2873-2893) used by the agents'
                           ``_encode_image_optimized`` (src/agents/vlm_inspector.py:63-64, vlm_auditor.py:90-91)
* ``resize_image_size``  — size rule of ``resize_image`` (utils/image_utils.py:62-73)
* ``pil_pass_order``     — which resample passes Pillow runs and in which order (PIL:Image.py:2400-2435)
"""
from __future__ import annotations

import math

PATCH_SIZE = 14
MERGE_SIZE = 2
TEMPORAL_PATCH_SIZE = 2
FACTOR = PATCH_SIZE * MERGE_SIZE
ROW_FLOATS = 3 * TEMPORAL_PATCH_SIZE * PATCH_SIZE * PATCH_SIZE      # 1176

# Qwen2VLImageProcessorPil class defaults (tf:...image_processing_pil_qwen2_vl.py:90)
DEFAULT_MIN_PIXELS = 56 * 56
DEFAULT_MAX_PIXELS = 28 * 28 * 1280
# Qwen2-VL-7B-Instruct hub preprocessor_config.json
HUB_MAX_PIXELS = 12845056

# OPENAI_CLIP_MEAN / OPENAI_CLIP_STD (tf:utils/constants.py:5-6), rescale_factor 1/255
IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)
RESCALE_FACTOR = 1 / 255

INSPECTOR_MAX_SIZE = 2048      # config.max_image_dimension default (utils/config.py:184)
AUDITOR_MAX_SIZE = 1024        # src/agents/vlm_auditor.py:85


def smart_resize(height: int, width: int, factor: int = FACTOR, min_pixels: int = DEFAULT_MIN_PIXELS,
                 max_pixels: int = DEFAULT_MAX_PIXELS) -> tuple[int, int]:
    """(h_bar, w_bar): both divisible by ``factor``, area within [min_pixels, max_pixels], aspect kept."""
    ratio = max(height, width) / min(height, width)
    if ratio > 200:
        raise ValueError(f"absolute aspect ratio must be smaller than 200, got {ratio}")
    h_bar = round(height / factor) * factor          # Python round: half to even
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


def grid_thw(height: int, width: int) -> tuple[int, int, int]:
    return 1, height // PATCH_SIZE, width // PATCH_SIZE


def thumbnail_size(width: int, height: int, max_size: int) -> tuple[int, int] | None:
    """Target (w, h) of ``img.thumbnail((max_size, max_size))`` or None when the image already fits."""
    x = y = math.floor(max_size)
    if x >= width and y >= height:
        return None
    aspect = width / height

    def pick(number, err):
        lo, hi = math.floor(number), math.ceil(number)
        return max(lo if err(lo) <= err(hi) else hi, 1)

    if x / y >= aspect:
        x = pick(y * aspect, lambda n: abs(aspect - n / y))
    else:
        y = pick(x / aspect, lambda n: 0 if n == 0 else abs(aspect - x / n))
    return x, y


def thumbnail_needs_reduce(width: int, height: int, tw: int, th: int, reducing_gap: float = 2.0) -> bool:
    """True when ``Image.thumbnail`` would run its box-reduce pre-pass (>= 4x downscale; PIL:Image.py:2413-2429)."""
    return (int(width / tw / reducing_gap) or 1) > 1 or (int(height / th / reducing_gap) or 1) > 1


_FILTER_SUPPORT = {1: 3.0, 3: 2.0}                 # LANCZOS, BICUBIC (PIL:Image.py _filters_support)


def reducing_plan(width: int, height: int, out_w: int, out_h: int, filt: int, box=None, reducing_gap: float = 2.0):
    """The pre-pass ``Image.resize(..., reducing_gap)`` inserts (PIL:Image.py:2407-2424; ``Image.thumbnail`` passes
    2.0): None when no reduction applies, else ((factor_x, factor_y), reduce_box of ints — ``Image._get_safe_box`` —,
    the source box as floats relative to the reduced image)."""
    box = (0, 0, width, height) if box is None else box
    factor_x = int((box[2] - box[0]) / out_w / reducing_gap) or 1
    factor_y = int((box[3] - box[1]) / out_h / reducing_gap) or 1
    if factor_x <= 1 and factor_y <= 1:
        return None
    support = _FILTER_SUPPORT[int(filt)] - 0.5
    sx, sy = support * (box[2] - box[0]) / out_w, support * (box[3] - box[1]) / out_h
    rb = (max(0, int(box[0] - sx)), max(0, int(box[1] - sy)),
          min(width, math.ceil(box[2] + sx)), min(height, math.ceil(box[3] + sy)))
    new_box = ((box[0] - rb[0]) / factor_x, (box[1] - rb[1]) / factor_y,
               (box[2] - rb[0]) / factor_x, (box[3] - rb[1]) / factor_y)
    return (factor_x, factor_y), rb, new_box


def resize_image_size(width: int, height: int, max_dimension: int) -> tuple[int, int] | None:
    """Target (w, h) of the reference ``resize_image``; None when it returns the input unchanged."""
    if width <= max_dimension and height <= max_dimension:
        return None
    if width > height:
        return max_dimension, int(height * (max_dimension / width))
    return int(width * (max_dimension / height)), max_dimension


def pil_pass_order(height: int, width: int, out_h: int, out_w: int) -> str:
    """'' (copy), 'h', 'v', 'hv' or 'vh' — the passes ``Image.resize`` runs for this geometry."""
    need_h, need_v = out_w != width, out_h != height
    if need_h and need_v:
        return "vh" if (height > width * 100 and out_h < height) else "hv"
    return ("h" if need_h else "") + ("v" if need_v else "")
