"""Host logic of the defect overlay: the box handling of the reference's ``draw_bounding_boxes``
(utils/image_utils.py:176-257) up to the point where it starts calling cv2, then expansion into leaf primitives
through the C ABI (``vis_overlay_expand``).  Rasterisation itself is the CUDA kernel behind ``vis_overlay_draw``.
"""
from __future__ import annotations

import ctypes as C
import logging

import threading

import numpy as np

from . import _native as N

logger = logging.getLogger("vision_inspection_system_b200.overlay")

_CONFIDENCE_LEVELS = {"low": 1, "medium": 2, "high": 3}
_RED_BGR = (0, 0, 255)               # CRITICAL / MODERATE (utils/image_utils.py:250)
_COSMETIC_BGR = (0, 200, 255)        # COSMETIC (utils/image_utils.py:252)


def filter_by_confidence(boxes, confidence_threshold: str = "low", criticality: str = "medium") -> list:
    """utils/image_utils.py:177-189: keep boxes at or above the threshold, or everything when criticality is high."""
    threshold = _CONFIDENCE_LEVELS.get(confidence_threshold, 1)
    if threshold <= 1 or criticality == "high":      # nothing can fall below "low": the reference keeps every box
        return boxes
    kept = []
    for box in boxes:
        level = _CONFIDENCE_LEVELS.get(box.get("confidence", "medium"), 2)
        if level >= threshold or criticality == "high":
            kept.append(box)
        else:
            logger.debug("Skipping low-confidence defect: %s (confidence=%s)", box.get("type", "unknown"),
                         box.get("confidence", "medium"))
    return kept


def boxes_to_pixels(boxes, img_width: int, img_height: int, confidence_threshold: str = "low",
                    criticality: str = "medium") -> np.ndarray:
    """Percent boxes -> ``VisBox`` records, skipping (with a warning) exactly the boxes the reference skips.  Like
    the reference (utils/image_utils.py:192-313) this never raises on the content of a box: labels may be any text."""
    return N.host_records(box_rows(boxes, img_width, img_height, confidence_threshold, criticality), N.BOX_DTYPE, "label")


def box_rows(boxes, img_width: int, img_height: int, confidence_threshold: str = "low", criticality: str = "medium") -> list:
    """The same as plain tuples ``(x, y, w, h, b, g, r, dashed, label bytes)`` — what a batch planner concatenates before it
    builds ONE record array (1024 small structured arrays cost more to concatenate than to fill)."""
    out = []
    cv_text = N.cv_text
    for i, box in enumerate(filter_by_confidence(boxes, confidence_threshold, criticality)):
        get = box.get
        raw_x, raw_y = get("x", 0), get("y", 0)
        raw_w, raw_h = get("width", 10), get("height", 10)
        if not (0 <= raw_x <= 100 and 0 <= raw_y <= 100 and 0 < raw_w <= 100 and 0 < raw_h <= 100):
            logger.warning("Invalid bbox coordinates (out of 0-100 range): %s", box)
            continue
        if raw_x + raw_w > 100 or raw_y + raw_h > 100:
            logger.warning("Bbox exceeds image bounds: x+width=%s, y+height=%s", raw_x + raw_w, raw_y + raw_h)
            continue
        area_percent = (raw_w * raw_h) / 100.0
        if area_percent < 0.1:
            logger.warning("Bbox too small (area=%.2f%%) - skipping: %s", area_percent, box)
            continue
        if area_percent > 50.0:
            logger.warning("Bbox too large (area=%.2f%%) - likely error, skipping: %s", area_percent, box)
            continue
        x = int((raw_x / 100.0) * img_width)          # int() truncation, as the reference
        y = int((raw_y / 100.0) * img_height)
        w = int((raw_w / 100.0) * img_width)
        h = int((raw_h / 100.0) * img_height)
        x = min(max(0, x), img_width - 1)
        y = min(max(0, y), img_height - 1)
        w = min(w, img_width - x)
        h = min(h, img_height - y)
        if w <= 0 or h <= 0:
            logger.warning("Bbox invalid after clamping, skipping: %s", box)
            continue
        label_full = get("label", f"#{i + 1}")        # index over the confidence-filtered list
        try:
            label_text = label_full.replace("#", "")
        except Exception:
            label_text = str(i + 1)
        color = _COSMETIC_BGR if get("severity", "MODERATE") == "COSMETIC" else _RED_BGR
        # any text, any length: cv2.putText receives the UTF-8 bytes and draws '?' for every byte outside 32..126
        out.append((x, y, w, h, color[0], color[1], color[2], 1 if get("confidence", "medium") == "low" else 0,
                    cv_text(label_text)))
    return out


def box_label(rec) -> bytes:
    """The label bytes of one ``VisBox`` record."""
    return N.record_string(rec, "label")


_tls = threading.local()


def _scratch_array(name: str, count: int, dtype) -> np.ndarray:
    """Grow-only host scratch, one set per thread (batch planning runs frames on a thread pool; the C entry points
    release the GIL), so that per-frame planning does not allocate and zero a worst-case buffer every time."""
    pool = _tls.__dict__.setdefault("scratch", {})
    a = pool.get(name)
    if a is None or len(a) < count:
        a = pool[name] = np.empty(max(count, 2 * (len(a) if a is not None else 0)), dtype)
    return a


def expand_leaves(pixel_boxes: np.ndarray, img_width: int, img_height: int) -> np.ndarray:
    """``VisBox`` records of one frame -> its leaf array (group headers first) via ``vis_overlay_expand``."""
    L = N.lib()
    n = len(pixel_boxes)
    if n == 0:
        return np.zeros(0, N.LEAF_DTYPE)
    boxes = np.ascontiguousarray(pixel_boxes)
    cap = 2048 * n
    while True:
        leaves = _scratch_array("leaves", cap, N.LEAF_DTYPE)
        needed = C.c_int(0)
        rc = L.vis_overlay_expand(img_height, img_width, boxes.ctypes.data_as(C.c_void_p), n,
                                  leaves.ctypes.data_as(C.c_void_p), len(leaves), C.byref(needed))
        if rc == N.VIS_E_CAPACITY:
            cap = needed.value
            continue
        N.check(rc, "vis_overlay_expand")
        return leaves[:rc].copy()


def touched_tiles(leaves: np.ndarray, n_boxes: int, img_width: int, img_height: int):
    """Bins one frame's leaves into the 64x16 tiles of the draw kernel via ``vis_overlay_tiles``.

    Returns (tiles int32 [n, 3] = (tx | ty << 16, first ref, one past last ref), refs int32 [m, 2] = leaf ranges)."""
    if n_boxes == 0 or len(leaves) == 0:
        return np.zeros((0, 3), np.int32), np.zeros((0, 2), np.int32)
    L = N.lib()
    tcap = ((img_width + 63) // 64) * ((img_height + 15) // 16)
    rcap = 4 * tcap
    while True:
        tiles = _scratch_array("tiles", tcap * 3, np.int32)
        refs = _scratch_array("refs", rcap * 2, np.int32)
        nt, nr = C.c_int(0), C.c_int(0)
        rc = L.vis_overlay_tiles(img_height, img_width, leaves.ctypes.data_as(C.c_void_p), n_boxes,
                                 tiles.ctypes.data_as(C.c_void_p), len(tiles) // 3, refs.ctypes.data_as(C.c_void_p),
                                 len(refs) // 2, C.byref(nt), C.byref(nr))
        if rc == N.VIS_E_CAPACITY:
            tcap, rcap = max(tcap, nt.value), max(rcap, nr.value)
            continue
        N.check(rc, "vis_overlay_tiles")
        return tiles[:3 * nt.value].reshape(-1, 3).copy(), refs[:2 * nr.value].reshape(-1, 2).copy()
