"""Host half of ``create_heatmap_overlay`` (utils/image_utils.py:320-604 of the reference; SURVEY.md 8f "next" row 2):
defect dicts -> per-defect heat parameters (intensity, centre, sigma, region, blur kernel) and the Gaussian kernels.

All of this is scalar Python arithmetic exactly as the reference does it (Python floats, ``int()`` truncation); the
per-pixel work — analytic heat of every defect, its separable blur, the max-composite, the final blur, the global
max, JET colouring and the 0.6/0.4 blend — runs in the CUDA library (vis_heatmap_*).
"""
from __future__ import annotations

import base64
import functools
import logging
import zlib

import numpy as np

logger = logging.getLogger("vision_inspection_system_b200.image_utils")

# cv2.COLORMAP_JET as applyColorMap produces it (BGR, 256 x 3), captured from opencv-python-headless 4.13.0
# (tests/golden/make_goldens.py prints it; tests/test_oracle_heatmap.py checks it against cv2 when importable)
JET_BGR = np.frombuffer(zlib.decompress(base64.b64decode(
    "eNod0gFHnQEAQNG7STKZJJkkmUmSZJIkSZIkSZJMkplkMkmSJEmSJEmSJEkmeZI8ySRJkskkSfIk8yTJJEmS7vJxfsLph0EYhlEYh0mYhlmYh0VYghCsQhg2"
    "YBO2YRf24QAO4RhOIQIXEIUruIFbuIdHeAZfxWAcxmMCJuEHTMV0/ISZmI25mIcFWIQlWIYVWIU1WIcN2IjN+A1bsQ3bsRO7sRf7cRCHcRTHcRKncRbncRGX"
    "MISrGMYN3MRt3MV9PMBDPMZTjOAFRvEKb/AW7/ERn3nx7ZOxD7678/0/E69NvjTlr2nnfjwz48SsI3P++Pm3+XsW7li8Zekvy9etXLN6xdpl63/6ZcGmOb/O"
    "2DLl9wl/jNkxYteQPQP29TnQ41CXIx2O/XDiu1Mtznx1rsmFL/6sd7nWlWrXKl0v91epW8XuFLqX7+/P/snxKMuTDM8+ep7m3xQvk71O9N977975EOvTG194"
    "lke5l1u5kSuJyoVE5FSO5VAOZF92ZVs2ZUPCsiohWZJFmZdZmZZJGZdRGZZB6Zde6ZZOaZc2aZVv0iyN0iB1UiNVUiFlUiJFUiB5kivZkimfJF1S5YMkSYLE"
    "S5zEGPR5DirdB61ugmLRoFskqHccNDwISu4GPTeDquGgbSgovBh0ng1qTwbNR4Pyr/H7/wN46G3P")), np.uint8).reshape(256, 3)

HEAT_DTYPE = np.dtype([
    ("kind", np.int32),                      # 0: box defect (Gaussian + boosts, own blur), 1: widespread (whole image)
    ("x", np.int32), ("y", np.int32), ("w", np.int32), ("h", np.int32),             # pixel box
    ("x1", np.int32), ("y1", np.int32), ("x2", np.int32), ("y2", np.int32),         # region [x1,x2) x [y1,y2)
    ("ksize", np.int32), ("koff", np.int32), ("pad", np.int32),                     # blur kernel size, offset into kernels[]
    ("intensity", np.float64), ("cx", np.float64), ("cy", np.float64), ("sigma", np.float64)], align=True)

# VisHeatItem: a defect of a batch (include/vis_b200.h)
ITEM_DTYPE = np.dtype([("d", HEAT_DTYPE), ("frame", np.int32), ("pad", np.int32), ("tmp_off", np.int64), ("tab_off", np.int64)],
                      align=True)

_SEVERITY_WEIGHT = {"CRITICAL": 1.0, "MODERATE": 0.75, "COSMETIC": 0.5, "MINOR": 0.5}        # image_utils.py:376-381
_CONFIDENCE_FACTOR = {"high": 1.0, "medium": 0.75, "low": 0.55}                               # :386
_WIDESPREAD = ("entire surface", "everywhere", "whole component", "complete surface")         # :398


@functools.lru_cache(maxsize=1024)
def _gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    t = np.exp((-0.5 / (sigma * sigma)) * x * x)
    k = (t * (1.0 / t.sum())).astype(np.float32)
    k.setflags(write=False)
    return k


def gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma, CV_32F) for sigma > 0: exp in double, normalised in double, cast to float."""
    return _gaussian_kernel(int(ksize), float(sigma))


def blur_kernel_size(sigma: float, cap: int) -> int:
    k = min(int(2 * np.ceil(3 * sigma) + 1), cap)
    return k + 1 if k % 2 == 0 else k


def defect_params(defects: list, width: int, height: int):
    """(HEAT_DTYPE records in list order, concatenated float32 blur kernels, had_defects).

    Mirrors utils/image_utils.py:364-571: intensity :376-395, widespread :398-418, box validation :420-441,
    percent->pixel + clamp :443-461, centre/sigma :470-494, region :501-505, per-defect blur :556-562."""
    recs, kernels, koff = [], [], 0
    had = False
    for defect in defects:
        had = True
        bbox = defect.get("bbox", {})
        severity = defect.get("safety_impact", "MODERATE")
        conf = defect.get("confidence", "medium")
        intensity = max(_SEVERITY_WEIGHT.get(severity, 0.6) * _CONFIDENCE_FACTOR.get(conf, 0.65), 0.35)
        if severity == "CRITICAL" and conf == "high":
            intensity = min(1.0, intensity * 1.2)
        location = defect.get("location", "").lower()
        has_valid_bbox = bool(bbox and bbox.get("x") is not None and bbox.get("y") is not None
                              and bbox.get("width", 0) > 0 and bbox.get("height", 0) > 0)
        if bbox is None and any(kw in location for kw in _WIDESPREAD):
            # (kind, x, y, w, h, x1, y1, x2, y2, ksize, koff, pad, intensity, cx, cy, sigma)
            recs.append((1, 0, 0, 0, 0, 0, 0, width, height, 1, 0, 0, intensity, width // 2, height // 2,
                         (max(width, height) // 2) * 0.7))
            continue
        if not has_valid_bbox:
            continue
        raw_x, raw_y = bbox.get("x", 0), bbox.get("y", 0)
        raw_w, raw_h = bbox.get("width", 10), bbox.get("height", 10)
        if not (0 <= raw_x <= 100 and 0 <= raw_y <= 100 and 0 < raw_w <= 100 and 0 < raw_h <= 100):
            logger.warning("Invalid bbox in heatmap (out of 0-100 range): %s", bbox)
            continue
        if raw_x + raw_w > 100 or raw_y + raw_h > 100:
            logger.warning("Bbox exceeds bounds in heatmap: %s", bbox)
            continue
        area = (raw_w * raw_h) / 100.0
        if area < 0.05 or area > 50.0:
            logger.warning("Bbox unreasonable size in heatmap (area=%.2f%%), skipping: %s", area, bbox)
            continue
        x, y = int((raw_x / 100.0) * width), int((raw_y / 100.0) * height)
        w, h = int((raw_w / 100.0) * width), int((raw_h / 100.0) * height)
        if x < 0:
            w += x
            x = 0
        if y < 0:
            h += y
            y = 0
        w, h = min(w, width - x), min(h, height - y)
        if w <= 0 or h <= 0:
            logger.warning("Invalid bbox size after conversion: w=%s, h=%s", w, h)
            continue
        cx, cy = float(x + w / 2.0), float(y + h / 2.0)
        sigma = max((w / 2.0) * 1.8, (h / 2.0) * 1.8)
        sigma = max(sigma, max(w, h) * 0.6, 20)
        sigma = min(sigma, min(width, height) * 0.15)
        margin = int(4 * sigma) + 15
        x1, y1 = int(max(0, cx - margin)), int(max(0, cy - margin))
        x2, y2 = int(min(width, cx + margin + 1)), int(min(height, cy + margin + 1))
        if x2 <= x1 or y2 <= y1:
            continue
        bsig = sigma * 0.4
        ksize = blur_kernel_size(bsig, 51)
        recs.append((0, x, y, w, h, x1, y1, x2, y2, ksize if ksize > 1 else 1, koff, 0, intensity, cx, cy, sigma))
        if ksize > 1:
            k = gaussian_kernel(ksize, bsig)
            kernels.append(k)
            koff += len(k)
    rec_arr = np.array(recs, HEAT_DTYPE) if recs else np.zeros(0, HEAT_DTYPE)
    kern = np.concatenate(kernels) if kernels else np.zeros(1, np.float32)
    return rec_arr, kern, had


def final_blur(width: int, height: int):
    """(ksize, kernel) of the whole-mask blur (:583-593): sigma = 1 % of the smaller side, at most 31 taps."""
    sigma = min(width, height) * 0.01
    ksize = blur_kernel_size(sigma, 31)
    if ksize <= 1:
        return 1, np.ones(1, np.float32)
    return ksize, gaussian_kernel(ksize, sigma)
