"""Seeded synthetic workloads of BASELINE.json's configs.  Host-side numpy only.

Uniform-noise frames are the parity worst case: they exercise clip8 and every rounding boundary of the
fixed-point resampler, and they do not compress in any cache.
"""
from __future__ import annotations

import numpy as np

FRAME_1080P = (1080, 1920)
FRAME_4K = (2160, 3840)
MIXED_RESOLUTIONS = [(480, 640), (720, 1280), (1080, 1920), (1536, 2048), (2160, 3840), (100, 502)]


def noise_frame(seed: int, h: int, w: int) -> np.ndarray:
    """``np.random.default_rng(seed).integers(0, 256, (h, w, 3), uint8)``."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def frames_1080p(n: int, first_seed: int = 1234) -> np.ndarray:
    """Config 2: n frames 1920x1080, seeds first_seed + i."""
    return np.stack([noise_frame(first_seed + i, *FRAME_1080P) for i in range(n)])


def frames_4k(n: int, first_seed: int = 4000) -> np.ndarray:
    """Config 3: n frames 3840x2160, seeds first_seed + i."""
    return np.stack([noise_frame(first_seed + i, *FRAME_4K) for i in range(n)])


def pattern_frames(h: int = 1080, w: int = 1920) -> dict:
    """Parity-only side set: gradients, constants, 1-px checkerboard, low-pass noise."""
    yy, xx = np.mgrid[0:h, 0:w]
    out = {
        "hgrad": np.repeat(((xx * 255) // max(w - 1, 1)).astype(np.uint8)[:, :, None], 3, 2),
        "vgrad": np.repeat(((yy * 255) // max(h - 1, 1)).astype(np.uint8)[:, :, None], 3, 2),
        "zeros": np.zeros((h, w, 3), np.uint8),
        "full": np.full((h, w, 3), 255, np.uint8),
        "checker": np.repeat((((xx + yy) & 1) * 255).astype(np.uint8)[:, :, None], 3, 2),
    }
    rng = np.random.default_rng(99)
    small = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3)).astype(np.float32)
    up = np.kron(small, np.ones((8, 8, 1), np.float32))[:h, :w]
    out["lowpass"] = np.clip(up + rng.normal(0, 4, up.shape), 0, 255).astype(np.uint8)
    return out


def random_boxes(rng: np.random.Generator, k: int) -> list:
    """k boxes per SURVEY 8(d) config 4: w,h ~ U(3,40)%, x ~ U(0,100-w), y ~ U(0,100-h), one decimal, area 0.1..50 %."""
    sev = ["CRITICAL", "MODERATE", "COSMETIC"]
    conf = ["high", "medium", "low"]
    out = []
    while len(out) < k:
        w = round(float(rng.uniform(3, 40)), 1)
        h = round(float(rng.uniform(3, 40)), 1)
        x = round(float(rng.uniform(0, 100 - w)), 1)
        y = round(float(rng.uniform(0, 100 - h)), 1)
        if not (0.1 <= w * h / 100 <= 50):
            continue
        out.append({"x": x, "y": y, "width": w, "height": h, "label": f"#{len(out) + 1}",
                    "severity": sev[int(rng.integers(0, 3))], "confidence": conf[int(rng.integers(0, 3))]})
    return out


def annotated_frame(seed: int, h: int = 1080, w: int = 1920):
    """Config 4: one BGR noise frame + K ~ U{1..8} boxes, all drawn from ``default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    return frame, random_boxes(rng, int(rng.integers(1, 9)))


def mixed_resolution_shapes(n: int, seed: int = 9000) -> list:
    """Config 5: per-frame (h, w) drawn uniformly from MIXED_RESOLUTIONS."""
    rng = np.random.default_rng(seed)
    return [MIXED_RESOLUTIONS[int(i)] for i in rng.integers(0, len(MIXED_RESOLUTIONS), n)]


def random_defects(rng: np.random.Generator, k: int) -> list:
    """k defect dicts as src/reporting/pdf_generator.py:1262-1279 hands them to create_heatmap_overlay."""
    sev = ["CRITICAL", "MODERATE", "COSMETIC", "MINOR"]
    conf = ["high", "medium", "low"]
    out = []
    for b in random_boxes(rng, k):
        out.append({"type": "crack", "bbox": {"x": b["x"], "y": b["y"], "width": b["width"], "height": b["height"]},
                    "safety_impact": sev[int(rng.integers(0, 4))], "confidence": conf[int(rng.integers(0, 3))],
                    "location": "surface"})
    return out



def heatmap_cases() -> list:
    """(name, BGR frame, defect list, subsample step) of the heat-map goldens (tests/golden, section "heatmap")."""
    small, vga = pattern_frames(360, 480), pattern_frames(480, 640)
    widespread = {"type": "corrosion", "bbox": None, "safety_impact": "MODERATE", "confidence": "medium",
                  "location": "Entire surface of the panel"}
    invalid = [{"type": "x", "bbox": {"x": 90, "y": 10, "width": 30, "height": 10}, "safety_impact": "CRITICAL", "confidence": "high"},
               {"type": "y", "bbox": {"x": 10, "y": 10, "width": 0.1, "height": 0.1}, "safety_impact": "MINOR", "confidence": "low"},
               {"type": "z", "bbox": {}, "safety_impact": "MINOR", "confidence": "low"}]
    return [
        ("hgrad_360x480_3", small["hgrad"], random_defects(np.random.default_rng(8000), 3), 1),
        ("vgrad_480x640_5", vga["vgrad"], random_defects(np.random.default_rng(8001), 5) + invalid[:1], 1),
        ("checker_360x480_widespread", small["checker"], [widespread] + random_defects(np.random.default_rng(8002), 1), 1),
        ("noise_1080p_4", noise_frame(8003, 1080, 1920), random_defects(np.random.default_rng(8003), 4), 8),
        ("zeros_360x480_critical_high", small["zeros"],
         [{"type": "crack", "bbox": {"x": 40, "y": 40, "width": 20, "height": 20}, "safety_impact": "CRITICAL",
           "confidence": "high", "location": "centre"},
          {"type": "edge", "bbox": {"x": 0, "y": 0, "width": 8, "height": 6}, "safety_impact": "COSMETIC",
           "confidence": "low", "location": "corner"},
          {"type": "edge2", "bbox": {"x": 92, "y": 90, "width": 8, "height": 10}, "safety_impact": "UNKNOWN",
           "confidence": "unsure", "location": "corner"}], 1),
        ("full_360x480_all_invalid", small["full"], invalid, 1),
        ("hgrad_360x480_empty", small["hgrad"], [], 1),
    ]


def agent_input_image(seed: int, shape, mode: str):
    """PIL input of an ``encode_image_optimized`` case: low-passed pattern (even seeds) or noise (odd seeds) in mode
    RGB / L / RGBA (noise alpha).  Shared by tests/golden/make_goldens.py and the tests."""
    from PIL import Image
    h, w = shape
    rgb = pattern_frames(h, w)["lowpass"] if seed % 2 == 0 else noise_frame(seed, h, w)
    if mode == "RGB":
        return Image.fromarray(rgb)
    if mode == "L":
        return Image.fromarray(rgb[:, :, 0].copy())
    alpha = noise_frame(seed + 1000, h, w)[:, :, 0]
    return Image.fromarray(np.dstack([rgb, alpha]))


def write_agent_input(img, path) -> None:
    """Lossless PNG, or the JPEG (quality 92) the case names."""
    if str(path).endswith(".jpg"):
        img.save(path, format="JPEG", quality=92)
    else:
        img.save(path, format="PNG", compress_level=1)
