"""Host logic of the comparison panel and the status stamp (utils/image_utils.py:608-739 in the reference): panel
geometry, label placement and stamp styling, expressed as a panel list (``vis_compose_panels``) and a draw list
(``vis_draw_expand``).  Every pixel is produced by the CUDA library; nothing here touches image data.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

TARGET_HEIGHT = 800              # utils/image_utils.py:635
HEADER_HEIGHT = 40               # :647
DIVIDER_WIDTH = 10               # :648
BAR_GRAY = 45                    # :650, :670
DEFAULT_LABELS = ("Original Input", "AI Analysis Layer")
LINE_8 = 8


def panel_width(height: int, width: int, target_h: int = TARGET_HEIGHT) -> int:
    """``int(w * (target_h / h))`` — utils/image_utils.py:637-641."""
    return int(width * (target_h / height))


def text_size(text: str, font_scale: float, thickness: int) -> tuple[int, int]:
    """``cv2.getTextSize(text, FONT_HERSHEY_SIMPLEX, font_scale, thickness)[0]`` through the C ABI."""
    w, h = C.c_int(0), C.c_int(0)
    N.check(N.lib().vis_text_size(N.cv_text(text), float(font_scale), int(thickness),
                                  C.byref(w), C.byref(h)), "vis_text_size")
    return w.value, h.value


def _cmd(kind, x1, y1, x2, y2, thickness, color, line_type=LINE_8, font_scale=0.0, text=""):
    color = tuple(color) + (0,) * (4 - len(color))
    return (kind, x1, y1, x2, y2, thickness, line_type, color, font_scale, N.cv_text(text))


def commands(rows: list) -> np.ndarray:
    """``VisDrawCmd`` records from ``_cmd`` tuples (the texts live in buffers the array keeps alive)."""
    return N.host_records(rows, N.DRAW_CMD_DTYPE, "text")


def commands_key(cmds: np.ndarray) -> tuple:
    """A hashable value identifying a draw list by CONTENT (the records hold pointers, so their bytes do not)."""
    return tuple((int(c["kind"]), int(c["x1"]), int(c["y1"]), int(c["x2"]), int(c["y2"]), int(c["thickness"]),
                  int(c["line_type"]), bytes(c["color"]), float(c["font_scale"]), N.record_string(c, "text")) for c in cmds)


def header_commands(left_w: int, right_w: int, labels=DEFAULT_LABELS) -> np.ndarray:
    """The two ``cv2.putText`` calls of the header bar (utils/image_utils.py:652-668)."""
    left_label, right_label = labels
    cmds = []
    tw, _ = text_size(left_label, 0.7, 2)
    cmds.append(_cmd(N.DRAW_TEXT, left_w // 2 - tw // 2, 28, 0, 0, 2, (255, 255, 255), font_scale=0.7, text=left_label))
    tw, _ = text_size(right_label, 0.7, 2)
    cmds.append(_cmd(N.DRAW_TEXT, left_w + DIVIDER_WIDTH + right_w // 2 - tw // 2, 28, 0, 0, 2, (255, 255, 255),
                     font_scale=0.7, text=right_label))
    return commands(cmds)


def stamp_style(verdict: str):
    """(text, BGRA colour, BGRA border colour) — utils/image_utils.py:711-722."""
    if verdict == "SAFE":
        return "PASSED", (0, 200, 0, 255), (0, 150, 0, 255)
    if verdict == "UNSAFE":
        return "REJECTED", (0, 0, 200, 255), (0, 0, 150, 255)
    return "REVIEW", (0, 140, 255, 255), (0, 100, 200, 255)


def stamp_commands(verdict: str, width: int, height: int) -> np.ndarray:
    """``cv2.rectangle(.., border, 4)`` then ``cv2.putText(.., 1.5, colour, 4)`` (utils/image_utils.py:725-733)."""
    text, color, border = stamp_style(verdict)
    tw, th = text_size(text, 1.5, 4)
    return commands([
        _cmd(N.DRAW_RECTANGLE, 5, 5, width - 5, height - 5, 4, border),
        _cmd(N.DRAW_TEXT, (width - tw) // 2, (height + th) // 2, 0, 0, 4, color, font_scale=1.5, text=text),
    ])


def expand_commands(cmds: np.ndarray, img_width: int, img_height: int) -> np.ndarray:
    """Draw list of one canvas -> its leaf array (group headers first) via ``vis_draw_expand``."""
    L = N.lib()
    n = len(cmds)
    if n == 0:
        return np.zeros(0, N.LEAF_DTYPE)
    owner = cmds                               # the records point into buffers this object keeps alive
    cmds = np.ascontiguousarray(cmds)
    cap = 4096
    while True:
        leaves = np.empty(cap, N.LEAF_DTYPE)
        needed = C.c_int(0)
        rc = L.vis_draw_expand(img_height, img_width, cmds.ctypes.data_as(C.c_void_p), n,
                               leaves.ctypes.data_as(C.c_void_p), cap, C.byref(needed))
        if rc == N.VIS_E_CAPACITY:
            cap = needed.value
            continue
        N.check(rc, "vis_draw_expand")
        del owner
        return leaves[:rc].copy()


def linear_tables(src_size: int, dst_size: int, is_x: bool):
    """(ofs int32 [dst], coef int16 [dst, 2]) of ``vis_linear_table``."""
    ofs = np.empty(dst_size, np.int32)
    coef = np.empty((dst_size, 2), np.int16)
    N.check(N.lib().vis_linear_table(src_size, dst_size, 1 if is_x else 0, N.i32ptr(ofs),
                                     coef.ctypes.data_as(C.c_void_p)), "vis_linear_table")
    return ofs, coef
