"""Drop-in mirror of the reference's ``src/safety/image_quality.py`` ("next" row 3).

Same entry point, result keys, thresholds and error behaviour:

    assess_image_quality(image_path) -> dict            src/safety/image_quality.py:182-185 (-> assess_quality :30-103)

The file is decoded on the host (``cv2.imread``, as in the reference); gray conversion, the 3x3 Laplacian and the three
reductions run in ONE CUDA pass (``vis_quality_stats``) that returns exact int64 sums; variance, mean and the scores
are finished here in float64.  ``laplacian_variance`` equals ``cv2.Laplacian(gray, CV_64F).var()`` to float64
rounding (the sums are exact; numpy's two-pass variance rounds differently in the last bits).  No CPU fallback.
"""
from __future__ import annotations

import logging
from fractions import Fraction
from pathlib import Path
from typing import Any, Dict

import numpy as np

logger = logging.getLogger("vision_inspection_system_b200.image_quality")

MIN_SHARPNESS = 100.0      # Laplacian variance threshold        (image_quality.py:24)
MIN_BRIGHTNESS = 30.0      # mean pixel value                     (:25)
MAX_BRIGHTNESS = 220.0     # avoid overexposed                    (:26)
MIN_RESOLUTION = 100       # minimum width or height              (:27)
MIN_PIXELS = 10000         # minimum total pixels                 (:28)


def _sharpness_score(var: float) -> float:
    if var < MIN_SHARPNESS:
        return var / MIN_SHARPNESS * 0.5
    return min(1.0, 0.5 + (var - MIN_SHARPNESS) / 400.0)


def _brightness_score(mean: float) -> float:
    if MIN_BRIGHTNESS <= mean <= MAX_BRIGHTNESS:
        ideal_center = (MIN_BRIGHTNESS + MAX_BRIGHTNESS) / 2
        return 1.0 - (abs(mean - ideal_center) / ((MAX_BRIGHTNESS - MIN_BRIGHTNESS) / 2)) * 0.3
    if mean < MIN_BRIGHTNESS:
        return max(0.0, mean / MIN_BRIGHTNESS * 0.6)
    return max(0.0, 1.0 - ((mean - MAX_BRIGHTNESS) / (255 - MAX_BRIGHTNESS)) * 0.8)


def _resolution_score(width: int, height: int) -> float:
    if min(width, height) < MIN_RESOLUTION:
        return 0.3
    if width * height < MIN_PIXELS:
        return 0.5
    return min(1.0, width * height / 2000000.0)


def _quality_failed(reason: str) -> Dict[str, Any]:
    return {"quality_score": 0.0, "quality_passed": False,
            "sharpness": {"score": 0.0, "passed": False}, "brightness": {"score": 0.0, "passed": False},
            "resolution": {"score": 0.0, "passed": False}, "error": reason}


def result_from_sums(width: int, height: int, sum_gray: int, sum_lap: int, sum_lap2: int) -> Dict[str, Any]:
    """The reference's result dict (without ``image_path``) from the exact sums of ``vis_quality_stats``."""
    n = width * height
    var = float((Fraction(sum_lap2) - Fraction(sum_lap * sum_lap, n)) / n)      # exact, rounded once
    mean = sum_gray / n
    s, b, r = _sharpness_score(var), _brightness_score(mean), _resolution_score(width, height)
    overall = 0.4 * s + 0.3 * b + 0.3 * r
    return {"quality_score": round(overall, 3), "quality_passed": overall >= 0.6,
            "sharpness": {"score": round(s, 3), "laplacian_variance": var, "passed": s >= 0.6},
            "brightness": {"score": round(b, 3), "mean_value": round(mean, 1), "passed": b >= 0.6},
            "resolution": {"score": round(r, 3), "width": width, "height": height, "total_pixels": n,
                           "passed": r >= 0.6}}


def assess_frames(frames) -> list:
    """Batch form: BGR uint8 HWC CUDA tensors (``[B,H,W,3]`` or a list) -> one result dict per frame."""
    from .engine import get_engine
    eng = get_engine()
    sums, shapes = eng.quality_stats(frames)
    host = sums.cpu().numpy()
    return [result_from_sums(w, h, int(host[i, 0]), int(host[i, 1]), int(host[i, 2])) for i, (h, w) in enumerate(shapes)]


def assess_image_quality(image_path: Path) -> Dict[str, Any]:
    """Assess image quality; never raises (a failed result dict with ``error`` instead, like the reference)."""
    try:
        import cv2
        import torch
        img = cv2.imread(str(image_path))
        if img is None:
            return _quality_failed(f"Failed to load image: {image_path}")
        result = assess_frames([torch.from_numpy(img).cuda()])[0]
        result["image_path"] = str(image_path)
        logger.info("Image quality assessment: score=%.2f", result["quality_score"])
        return result
    except Exception as e:                                  # image_quality.py:101-103
        logger.error("Image quality assessment failed: %s", e, exc_info=True)
        return _quality_failed(f"Assessment error: {str(e)}")
