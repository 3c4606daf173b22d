"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's image-quality assessment.

Follows /root/reference/src/safety/image_quality.py:
  assess_quality :30-103, _assess_sharpness :105-119, _compute_laplacian_variance :121-124,
  _assess_brightness :126-149, _assess_resolution :151-168, thresholds :23-28
and the two OpenCV 4.13 routines it calls (third-party, opencv-python>=4.11.0.86; installed 4.13.0):
  cv2.cvtColor(BGR2GRAY)  — RGB2Gray<uchar>: (3735*B + 19235*G + 9798*R + 2^14) >> 15
  cv2.Laplacian(CV_64F)   — ksize 1: kernel [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101
Pinned by tests/test_oracle_quality.py against cv2 itself (when importable) and against results captured from the
reference's own ImageQualityAssessment (tests/golden/goldens.json, section "quality").
Only tests/, __graft_entry__.smoke() and bench tools may import this module; the product never does.
"""
from __future__ import annotations

import numpy as np

MIN_SHARPNESS, MIN_BRIGHTNESS, MAX_BRIGHTNESS, MIN_RESOLUTION, MIN_PIXELS = 100.0, 30.0, 220.0, 100, 10000


def gray_from_bgr(bgr: np.ndarray) -> np.ndarray:
    b, g, r = (bgr[:, :, i].astype(np.int64) for i in range(3))
    return ((3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15).astype(np.uint8)


def laplacian(gray: np.ndarray) -> np.ndarray:
    g = gray.astype(np.int64)
    h, w = g.shape
    p = np.pad(g, 1, mode="reflect") if h > 1 and w > 1 else np.pad(g, 1, mode="edge" if min(h, w) == 1 else "reflect")
    if h == 1 or w == 1:                      # reflect-101 of a single row/column is the row/column itself
        p = g
        if h == 1:
            p = np.concatenate([p, p, p], 0)
        else:
            p = np.pad(p, ((1, 1), (0, 0)), mode="reflect")
        if w == 1:
            p = np.concatenate([p, p, p], 1)
        else:
            p = np.pad(p, ((0, 0), (1, 1)), mode="reflect")
    return p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:] - 4 * p[1:-1, 1:-1]


def stats(bgr: np.ndarray):
    """(sum gray, sum laplacian, sum laplacian^2) as Python ints — what the CUDA kernel returns."""
    g = gray_from_bgr(bgr)
    lap = laplacian(g)
    return int(g.astype(np.int64).sum()), int(lap.sum()), int((lap * lap).sum())


def laplacian_variance(bgr: np.ndarray) -> float:
    return float(laplacian(gray_from_bgr(bgr)).astype(np.float64).var())


def sharpness_score(var: float) -> float:
    if var < MIN_SHARPNESS:
        return var / MIN_SHARPNESS * 0.5
    return min(1.0, 0.5 + (var - MIN_SHARPNESS) / 400.0)


def brightness_score(mean: float) -> float:
    if MIN_BRIGHTNESS <= mean <= MAX_BRIGHTNESS:
        ideal = (MIN_BRIGHTNESS + MAX_BRIGHTNESS) / 2
        return 1.0 - (abs(mean - ideal) / ((MAX_BRIGHTNESS - MIN_BRIGHTNESS) / 2)) * 0.3
    if mean < MIN_BRIGHTNESS:
        return max(0.0, mean / MIN_BRIGHTNESS * 0.6)
    return max(0.0, 1.0 - ((mean - MAX_BRIGHTNESS) / (255 - MAX_BRIGHTNESS)) * 0.8)


def resolution_score(width: int, height: int) -> float:
    if min(width, height) < MIN_RESOLUTION:
        return 0.3
    if width * height < MIN_PIXELS:
        return 0.5
    return min(1.0, width * height / 2000000.0)


def assess(bgr: np.ndarray) -> dict:
    """The result dict of assess_quality (without image_path) for an in-memory BGR frame."""
    h, w = bgr.shape[:2]
    var = laplacian_variance(bgr)
    mean = float(np.mean(gray_from_bgr(bgr)))
    s, b, r = sharpness_score(var), brightness_score(mean), resolution_score(w, h)
    overall = 0.4 * s + 0.3 * b + 0.3 * r
    return {"quality_score": round(overall, 3), "quality_passed": overall >= 0.6,
            "sharpness": {"score": round(s, 3), "laplacian_variance": var, "passed": s >= 0.6},
            "brightness": {"score": round(b, 3), "mean_value": round(mean, 1), "passed": b >= 0.6},
            "resolution": {"score": round(r, 3), "width": w, "height": h, "total_pixels": w * h, "passed": r >= 0.6}}
