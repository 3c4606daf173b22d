"""Oracle (TEST INFRASTRUCTURE ONLY): ``create_side_by_side_comparison`` and ``create_status_stamp`` on the CPU.

Restates utils/image_utils.py:608-739 of the reference on top of the plain-C restatements of OpenCV
(``cvresize_oracle.c`` for ``cv2.resize``, ``cvdraw_oracle.c`` for the drawing calls); file decode / encode excluded.
``*_cv2`` variants issue the same calls against the installed cv2 binary and pin the C code.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import lib
from .overlay import LINE_8, _CDraw, _Cv2Draw

TARGET_HEIGHT = 800          # utils/image_utils.py:635
HEADER_HEIGHT = 40           # :647
DIVIDER = 10                 # :648
GRAY = 45                    # :650, :670
LABELS = ("Original Input", "AI Analysis Layer")

_U8P = ctypes.POINTER(ctypes.c_uint8)


def resize_linear(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """``cv2.resize(img, (dw, dh))`` (INTER_LINEAR, uint8 HWC) through cvresize_oracle.c."""
    img = np.ascontiguousarray(img)
    h, w, cn = img.shape
    out = np.empty((dh, dw, cn), np.uint8)
    rc = lib().ocv_resize_linear_u8(img.ctypes.data_as(_U8P), h, w, img.strides[0], cn,
                                    out.ctypes.data_as(_U8P), dh, dw, out.strides[0])
    if rc:
        raise ValueError(f"ocv_resize_linear_u8 failed ({rc})")
    return out


def panel_width(h: int, w: int, target_h: int = TARGET_HEIGHT) -> int:
    """utils/image_utils.py:637-641: ``int(w * (target_h / h))``."""
    return int(w * (target_h / h))


def _compose(original, annotated, labels, resize, make_draw):
    left = resize(original, panel_width(*original.shape[:2]), TARGET_HEIGHT)
    right = resize(annotated, panel_width(*annotated.shape[:2]), TARGET_HEIGHT)
    total = left.shape[1] + right.shape[1] + DIVIDER
    header = np.full((HEADER_HEIGHT, total, 3), GRAY, np.uint8)
    d = make_draw(header)
    left_label, right_label = labels
    tw, _ = d.text_size(left_label, 0.7, 2)
    d.put_text(left_label, (left.shape[1] // 2 - tw // 2, 28), 0.7, (255, 255, 255), 2)
    tw, _ = d.text_size(right_label, 0.7, 2)
    d.put_text(right_label, (left.shape[1] + DIVIDER + right.shape[1] // 2 - tw // 2, 28), 0.7, (255, 255, 255), 2)
    divider = np.full((TARGET_HEIGHT, DIVIDER, 3), GRAY, np.uint8)
    return np.vstack([header, np.hstack([left, divider, right])])


def side_by_side(original_bgr: np.ndarray, annotated_bgr: np.ndarray, labels=LABELS) -> np.ndarray:
    """The array the reference hands to ``cv2.imwrite`` (:683), computed with the C restatements."""
    return _compose(original_bgr, annotated_bgr, labels, resize_linear, _CDraw)


def side_by_side_cv2(original_bgr: np.ndarray, annotated_bgr: np.ndarray, labels=LABELS) -> np.ndarray:
    import cv2
    return _compose(original_bgr, annotated_bgr, labels, lambda im, dw, dh: cv2.resize(im, (dw, dh)), _Cv2Draw)


def stamp_style(verdict: str):
    """(text, colour BGRA, border colour BGRA) — utils/image_utils.py:711-722."""
    if verdict == "SAFE":
        return "PASSED", (0, 200, 0, 255), (0, 150, 0, 255)
    if verdict == "UNSAFE":
        return "REJECTED", (0, 0, 200, 255), (0, 0, 150, 255)
    return "REVIEW", (0, 140, 255, 255), (0, 100, 200, 255)


def status_stamp(verdict: str, size=(300, 100)) -> np.ndarray:
    """The BGRA array of ``create_status_stamp`` (:706-733).  Every primitive is LINE_8 (no blending): a touched pixel
    receives all four channels of the colour, so the C oracle (3-channel) draws the BGR planes with the colour and the
    alpha plane with (A, A, A) over the same geometry."""
    width, height = size
    text, color, border = stamp_style(verdict)
    planes = []
    for pick in (lambda c: c[:3], lambda c: (c[3],) * 3):
        img = np.zeros((height, width, 3), np.uint8)
        d = _CDraw(img)
        d.rectangle((5, 5), (width - 5, height - 5), pick(border), 4, LINE_8)
        tw, th = d.text_size(text, 1.5, 4)
        d.put_text(text, ((width - tw) // 2, (height + th) // 2), 1.5, pick(color), 4)
        planes.append(img)
    return np.concatenate([planes[0], planes[1][:, :, :1]], axis=2)


def status_stamp_cv2(verdict: str, size=(300, 100)) -> np.ndarray:
    import cv2
    width, height = size
    text, color, border = stamp_style(verdict)
    stamp = np.zeros((height, width, 4), np.uint8)
    cv2.rectangle(stamp, (5, 5), (width - 5, height - 5), border, 4)
    tw, th = cv2.getTextSize(text, cv2.FONT_HERSHEY_SIMPLEX, 1.5, 4)[0]
    cv2.putText(stamp, text, ((width - tw) // 2, (height + th) // 2), cv2.FONT_HERSHEY_SIMPLEX, 1.5, color, 4)
    return stamp
