"""Oracle (TEST INFRASTRUCTURE ONLY): the defect overlay of ``draw_bounding_boxes`` on the CPU.

Restates utils/image_utils.py:176-313 of the reference in two steps:
  * ``select_boxes``  — confidence filter (:177-189), percent validation (:200-217), percent->pixel conversion with
    ``int()`` truncation (:220-223), clamping (:229-237), label / colour / dash style (:240-257);
  * ``render``        — the drawing calls (:259-313) issued against ``cvdraw_oracle.c`` (a plain-C restatement of
    OpenCV 4.13 ``drawing.cpp``) instead of cv2.
``render_cv2`` issues the very same calls against the installed cv2 binary; tests use it to pin the C restatement.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import lib

LINE_8 = 8
LINE_AA = 16

_LEVEL = {"low": 1, "medium": 2, "high": 3}
RED = (0, 0, 255)            # BGR, utils/image_utils.py:250
COSMETIC = (0, 200, 255)     # utils/image_utils.py:252
DASH, GAP = 10, 5            # utils/image_utils.py:262-263


@dataclass
class PixelBox:
    x: int
    y: int
    w: int
    h: int
    color: tuple
    dashed: bool
    label: str


def select_boxes(boxes, img_w: int, img_h: int, confidence_threshold: str = "low", criticality: str = "medium"):
    """Boxes that the reference would draw, in drawing order, in pixels."""
    need = _LEVEL.get(confidence_threshold, 1)
    kept = [b for b in boxes
            if _LEVEL.get(b.get("confidence", "medium"), 2) >= need or criticality == "high"]
    out = []
    for i, b in enumerate(kept):
        px, py = b.get("x", 0), b.get("y", 0)
        pw, ph = b.get("width", 10), b.get("height", 10)
        if not (0 <= px <= 100 and 0 <= py <= 100 and 0 < pw <= 100 and 0 < ph <= 100):
            continue
        if px + pw > 100 or py + ph > 100:
            continue
        area = (pw * ph) / 100.0
        if area < 0.1 or area > 50.0:
            continue
        x = int((px / 100.0) * img_w)
        y = int((py / 100.0) * img_h)
        w = int((pw / 100.0) * img_w)
        h = int((ph / 100.0) * img_h)
        x = min(max(0, x), img_w - 1)
        y = min(max(0, y), img_h - 1)
        w = min(w, img_w - x)
        h = min(h, img_h - y)
        if w <= 0 or h <= 0:
            continue
        label = b.get("label", f"#{i + 1}")
        try:
            text = label.replace("#", "")
        except Exception:
            text = str(i + 1)
        color = COSMETIC if b.get("severity", "MODERATE") == "COSMETIC" else RED
        out.append(PixelBox(x, y, w, h, color, b.get("confidence", "medium") == "low", text))
    return out


def marker_geometry(box: PixelBox, img_w: int, img_h: int):
    """(cx, cy, radius, font_scale, text_thickness) — utils/image_utils.py:290-306."""
    r = int(max(img_w, img_h) * 0.04)
    r = max(25, min(r, 60))
    cx = max(r + 5, min(box.x + r + 5, img_w - r - 5))
    cy = max(r + 5, min(box.y + r + 5, img_h - r - 5))
    fs = r / 20.0 * 0.7
    return cx, cy, r, fs, max(2, int(fs * 2))


class _CDraw:
    """cv2-shaped facade over cvdraw_oracle.c."""

    def __init__(self, img: np.ndarray):
        assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3 and img.flags.c_contiguous
        self.img = img
        self.L = lib()
        self.p = img.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        self.h, self.w = img.shape[:2]
        self.step = img.strides[0]

    def line(self, p1, p2, color, thickness, line_type):
        self.L.ocv_line(self.p, self.h, self.w, self.step, p1[0], p1[1], p2[0], p2[1], *color, thickness, line_type)

    def rectangle(self, p1, p2, color, thickness, line_type):
        self.L.ocv_rectangle(self.p, self.h, self.w, self.step, p1[0], p1[1], p2[0], p2[1], *color, thickness,
                             line_type)

    def circle(self, c, radius, color, thickness, line_type=LINE_8):
        self.L.ocv_circle(self.p, self.h, self.w, self.step, c[0], c[1], radius, *color, thickness, line_type)

    def text_size(self, text, fs, thickness):
        w, h = ctypes.c_int(), ctypes.c_int()
        if self.L.ocv_get_text_size(_cv_bytes(text), fs, thickness, ctypes.byref(w), ctypes.byref(h)):
            raise ValueError(f"glyph outside the oracle's Hershey table: {text!r}")
        return w.value, h.value

    def put_text(self, text, org, fs, color, thickness):
        if self.L.ocv_put_text(self.p, self.h, self.w, self.step, _cv_bytes(text), org[0], org[1], fs,
                               *color, thickness):
            raise ValueError(f"glyph outside the oracle's Hershey table: {text!r}")


def _cv_bytes(text: str) -> bytes:
    """what cv2.putText receives from Python: the UTF-8 bytes, cut at the first NUL"""
    return text.encode("utf-8", "replace").split(b"\0")[0]


class _Cv2Draw:
    """The same facade over the installed cv2 binary (used to pin the C restatement)."""

    def __init__(self, img: np.ndarray):
        import cv2
        self.cv2 = cv2
        self.img = img

    def line(self, p1, p2, color, thickness, line_type):
        self.cv2.line(self.img, p1, p2, color, thickness, line_type)

    def rectangle(self, p1, p2, color, thickness, line_type):
        self.cv2.rectangle(self.img, p1, p2, color, thickness, line_type)

    def circle(self, c, radius, color, thickness, line_type=LINE_8):
        self.cv2.circle(self.img, c, radius, color, thickness, line_type)

    def text_size(self, text, fs, thickness):
        return self.cv2.getTextSize(text, self.cv2.FONT_HERSHEY_SIMPLEX, fs, thickness)[0]

    def put_text(self, text, org, fs, color, thickness):
        self.cv2.putText(self.img, text, org, self.cv2.FONT_HERSHEY_SIMPLEX, fs, color, thickness)


def _draw(d, pixel_boxes, img_w, img_h):
    for b in pixel_boxes:
        x, y, w, h, col = b.x, b.y, b.w, b.h, b.color
        if b.dashed:
            for yy in (y, y + h):                                   # top edge, then bottom edge
                for px in range(x, x + w, DASH + GAP):
                    ex = min(px + DASH, x + w)
                    if ex > px:
                        d.line((px, yy), (ex, yy), col, 2, LINE_AA)
            for xx in (x, x + w):                                   # left edge, then right edge
                for py in range(y, y + h, DASH + GAP):
                    ey = min(py + DASH, y + h)
                    if ey > py:
                        d.line((xx, py), (xx, ey), col, 2, LINE_AA)
        else:
            d.rectangle((x, y), (x + w, y + h), col, 2, LINE_AA)
        cx, cy, r, fs, tt = marker_geometry(b, img_w, img_h)
        d.circle((cx, cy), r, (255, 255, 255), -1)
        d.circle((cx, cy), r, col, 3)
        tw, th = d.text_size(b.label, fs, tt)
        d.put_text(b.label, (int(cx - tw / 2), int(cy + th / 2)), fs, (0, 0, 0), tt)


def render(img_bgr: np.ndarray, pixel_boxes) -> np.ndarray:
    """Draw on a copy of ``img_bgr`` with the plain-C restatement."""
    out = np.ascontiguousarray(img_bgr).copy()
    _draw(_CDraw(out), pixel_boxes, out.shape[1], out.shape[0])
    return out


def render_cv2(img_bgr: np.ndarray, pixel_boxes) -> np.ndarray:
    """Draw on a copy of ``img_bgr`` with the installed cv2 binary."""
    out = np.ascontiguousarray(img_bgr).copy()
    _draw(_Cv2Draw(out), pixel_boxes, out.shape[1], out.shape[0])
    return out


def draw_bounding_boxes(img_bgr: np.ndarray, boxes, confidence_threshold: str = "low",
                        criticality: str = "medium") -> np.ndarray:
    """In-memory form of the reference function (file decode/encode excluded): returns the annotated BGR array."""
    h, w = img_bgr.shape[:2]
    return render(img_bgr, select_boxes(boxes, w, h, confidence_threshold, criticality))
