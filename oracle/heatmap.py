"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's create_heatmap_overlay.

Follows /root/reference/utils/image_utils.py:320-604 line by line (same numpy expressions, same dtypes) and restates
the three OpenCV 4.13 routines it calls (third-party, opencv-python>=4.11.0.86; installed 4.13.0):
  cv2.GaussianBlur(float32, (k,k), sigma)   separable, kernel = getGaussianKernel(k, sigma, CV_32F) (double exp,
                                            normalised in double, cast to float), BORDER_REFLECT_101, float32 accumulate
  cv2.applyColorMap(COLORMAP_JET)           256 x 3 table (captured from the binary, stored in the package)
  cv2.addWeighted(a, 0.6, b, 0.4, 0)        float32 multiply-add, round half to even, saturate
PARITY NOTE: OpenCV's separable filter accumulates in float32 in a SIMD-dependent order that is not reproduced here
(nor on the GPU), so this path is pinned with a TOLERANCE: tests/test_oracle_heatmap.py requires this restatement to
be within 2 levels of the reference's own output (tests/golden, section "heatmap") with >= 99.5 % of the bytes equal.
Only tests/ and bench tools may import this module; the product never does.
"""
from __future__ import annotations

import numpy as np

SEVERITY_WEIGHT = {"CRITICAL": 1.0, "MODERATE": 0.75, "COSMETIC": 0.5, "MINOR": 0.5}
CONFIDENCE_FACTOR = {"high": 1.0, "medium": 0.75, "low": 0.55}
WIDESPREAD = ["entire surface", "everywhere", "whole component", "complete surface"]


def gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    t = np.exp((-0.5 / (sigma * sigma)) * x * x)
    return (t * (1.0 / t.sum())).astype(np.float32)


def gaussian_blur(a: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(a float32, (ksize, ksize), sigma) up to float32 summation order."""
    k = gaussian_kernel(ksize, sigma)
    r = ksize // 2
    mode = "reflect"

    def pad(v, n, axis):
        # BORDER_REFLECT_101 also when the array is shorter than the radius (np.pad 'reflect' handles repeats)
        widths = [(0, 0), (0, 0)]
        widths[axis] = (n, n)
        return np.pad(v, widths, mode=mode) if v.shape[axis] > 1 else np.repeat(v, 2 * n + 1, axis=axis)
    p = pad(a, r, 1)
    h = np.zeros_like(a)
    for i in range(ksize):
        h += k[i] * p[:, i:i + a.shape[1]]
    p = pad(h, r, 0)
    v = np.zeros_like(a)
    for i in range(ksize):
        v += k[i] * p[i:i + a.shape[0], :]
    return v


def heat_mask(defects: list, width: int, height: int):
    """(float32 heat mask after the final blur, has_defects) — utils/image_utils.py:361-593."""
    heat = np.zeros((height, width), dtype=np.float32)
    has_defects = False
    for defect in defects:
        has_defects = True
        bbox = defect.get("bbox", {})
        severity = defect.get("safety_impact", "MODERATE")
        conf = defect.get("confidence", "medium")
        intensity = max(SEVERITY_WEIGHT.get(severity, 0.6) * CONFIDENCE_FACTOR.get(conf, 0.65), 0.35)
        if severity == "CRITICAL" and conf == "high":
            intensity = min(1.0, intensity * 1.2)
        location_lower = defect.get("location", "").lower()
        has_valid_bbox = (bbox and bbox.get("x") is not None and bbox.get("y") is not None and
                          bbox.get("width", 0) > 0 and bbox.get("height", 0) > 0)
        if bbox is None and any(kw in location_lower for kw in WIDESPREAD):
            cx, cy = width // 2, height // 2
            radius = max(width, height) // 2
            ys, xs = np.ogrid[:height, :width]
            dist_sq = (xs - cx) ** 2 + (ys - cy) ** 2
            g = intensity * np.exp(-dist_sq / (2 * (radius * 0.7) ** 2))
            heat = np.maximum(heat, g.astype(np.float32))
            continue
        if not has_valid_bbox:
            continue
        rx, ry, rw, rh = bbox.get("x", 0), bbox.get("y", 0), bbox.get("width", 10), bbox.get("height", 10)
        if not (0 <= rx <= 100 and 0 <= ry <= 100 and 0 < rw <= 100 and 0 < rh <= 100):
            continue
        if rx + rw > 100 or ry + rh > 100:
            continue
        area = (rw * rh) / 100.0
        if area < 0.05 or area > 50.0:
            continue
        x, y = int((rx / 100.0) * width), int((ry / 100.0) * height)
        w, h = int((rw / 100.0) * width), int((rh / 100.0) * height)
        if x < 0:
            w += x
            x = 0
        if y < 0:
            h += y
            y = 0
        w, h = min(w, width - x), min(h, height - y)
        if w <= 0 or h <= 0:
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
        ys, xs = np.ogrid[y1:y2, x1:x2]
        dist_sq = (xs - cx) ** 2 + (ys - cy) ** 2
        g = intensity * np.exp(-dist_sq / (2 * sigma ** 2))
        in_bbox = ((xs >= x) & (xs < (x + w))) & ((ys >= y) & (ys < (y + h)))
        ddx = (xs - cx) / max(w / 2.0, 1)
        ddy = (ys - cy) / max(h / 2.0, 1)
        strong = (ddx ** 2 + ddy ** 2) < 1.2 ** 2
        g = np.minimum(1.0, g * np.where(strong, 1.8, np.where(in_bbox, 1.4, 1.0)))
        local = np.where(dist_sq < (4.0 * sigma) ** 2, g.astype(np.float32), 0).astype(np.float32)
        bsig = sigma * 0.4
        ksize = min(int(2 * np.ceil(3 * bsig) + 1), 51)
        if ksize % 2 == 0:
            ksize += 1
        if ksize > 1 and local.size > 0:
            local = gaussian_blur(local, ksize, bsig)
        heat[y1:y2, x1:x2] = np.maximum(heat[y1:y2, x1:x2], local)
    if has_defects and heat.max() > 0:
        bsig = min(width, height) * 0.01
        ksize = min(int(2 * np.ceil(3 * bsig) + 1), 31)
        if ksize % 2 == 0:
            ksize += 1
        if ksize > 1:
            heat = gaussian_blur(heat, ksize, bsig)
    return heat, has_defects


def create_heatmap_overlay(bgr: np.ndarray, defects: list, jet_bgr: np.ndarray) -> np.ndarray:
    """The BGR array the reference hands to cv2.imwrite."""
    h, w = bgr.shape[:2]
    heat, has_defects = heat_mask(defects, w, h)
    if not has_defects:
        return bgr.copy()
    if heat.max() > 0:
        norm = (heat / heat.max() * 255).astype(np.uint8)
    else:
        norm = (heat * 255).astype(np.uint8)
    color = jet_bgr[norm]
    t = bgr.astype(np.float32) * np.float32(0.6) + color.astype(np.float32) * np.float32(0.4)
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)
