"""Pins the drawing oracle (oracle/cvdraw_oracle.c + oracle/overlay.py) against the installed cv2 binary and against
overlays captured from the reference's own draw_bounding_boxes (tests/golden/)."""
import ctypes
import hashlib

import numpy as np
import pytest

from oracle import lib, overlay as OV


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def P(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def test_overlay_goldens(goldens):
    for rec in goldens["overlay"]:
        h, w = rec["shape"]
        frame = np.random.default_rng(rec["seed"]).integers(0, 256, (h, w, 3), dtype=np.uint8)
        got = OV.draw_bounding_boxes(frame, rec["boxes"], rec["confidence_threshold"], rec["criticality"])
        assert int((got != frame).any(2).sum()) == rec["changed_pixels"], rec["name"]
        assert sha(got) == rec["sha256"], rec["name"]


def test_primitives_against_cv2():
    cv2 = pytest.importorskip("cv2")
    L = lib()
    rng = np.random.default_rng(1)
    for it in range(480):
        H, W = int(rng.integers(40, 260)), int(rng.integers(40, 260))
        base = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        a, b = base.copy(), base.copy()
        col = tuple(int(v) for v in rng.integers(0, 256, 3))
        x1, y1, x2, y2 = [int(v) for v in rng.integers(-30, 290, 4)]
        kind = it % 6
        t = int(rng.integers(2, 5))
        if kind == 0:
            cv2.line(a, (x1, y1), (x2, y2), col, t, cv2.LINE_AA)
            L.ocv_line(P(b), H, W, b.strides[0], x1, y1, x2, y2, *col, t, 16)
        elif kind == 1:
            cv2.line(a, (x1, y1), (x2, y2), col, t, cv2.LINE_8)
            L.ocv_line(P(b), H, W, b.strides[0], x1, y1, x2, y2, *col, t, 8)
        elif kind == 2:
            cv2.rectangle(a, (x1, y1), (x2, y2), col, t, cv2.LINE_AA)
            L.ocv_rectangle(P(b), H, W, b.strides[0], x1, y1, x2, y2, *col, t, 16)
        elif kind == 3:
            r = int(rng.integers(0, 80))
            cv2.circle(a, (x1, y1), r, col, -1)
            L.ocv_circle(P(b), H, W, b.strides[0], x1, y1, r, *col, -1, 8)
        elif kind == 4:
            r = int(rng.integers(0, 80))
            cv2.circle(a, (x1, y1), r, col, t)
            L.ocv_circle(P(b), H, W, b.strides[0], x1, y1, r, *col, t, 8)
        else:
            fs = float(rng.uniform(0.4, 2.5))
            if it % 12 == 5:
                txt = str(int(rng.integers(0, 1000)))
            else:                                        # any printable ASCII (the full Hershey simplex table)
                txt = "".join(chr(int(c)) for c in rng.integers(32, 127, int(rng.integers(1, 6))))
            cv2.putText(a, txt, (x1, y1), cv2.FONT_HERSHEY_SIMPLEX, fs, col, t)
            assert L.ocv_put_text(P(b), H, W, b.strides[0], txt.encode(), x1, y1, fs, *col, t) == 0
            tw, th = ctypes.c_int(), ctypes.c_int()
            L.ocv_get_text_size(txt.encode(), fs, t, ctypes.byref(tw), ctypes.byref(th))
            assert (tw.value, th.value) == cv2.getTextSize(txt, cv2.FONT_HERSHEY_SIMPLEX, fs, t)[0]
        assert np.array_equal(a, b), (kind, (H, W), (x1, y1, x2, y2), t)


def test_overlay_against_cv2_calls():
    pytest.importorskip("cv2")
    from vision_inspection_system_b200 import synth
    for seed, shape in ((1, (480, 640)), (2, (720, 1280)), (3, (300, 500)), (4, (1080, 1920))):
        frame, boxes = synth.annotated_frame(seed, *shape)
        px = OV.select_boxes(boxes, shape[1], shape[0])
        assert np.array_equal(OV.render(frame, px), OV.render_cv2(frame, px)), seed


def test_every_printable_ascii_glyph_against_cv2():
    cv2 = pytest.importorskip("cv2")
    L = lib()
    for c in range(32, 127):
        for fs, t in ((0.7, 2), (2.1, 2), (1.3, 3)):
            a = np.zeros((120, 160, 3), np.uint8)
            b = a.copy()
            cv2.putText(a, chr(c), (30, 90), cv2.FONT_HERSHEY_SIMPLEX, fs, (255, 200, 50), t)
            assert L.ocv_put_text(P(b), 120, 160, b.strides[0], chr(c).encode(), 30, 90, fs, 255, 200, 50, t) == 0
            assert np.array_equal(a, b), (chr(c), fs, t)
            tw, th = ctypes.c_int(), ctypes.c_int()
            L.ocv_get_text_size(chr(c).encode(), fs, t, ctypes.byref(tw), ctypes.byref(th))
            assert (tw.value, th.value) == cv2.getTextSize(chr(c), cv2.FONT_HERSHEY_SIMPLEX, fs, t)[0], chr(c)
    # outside printable ASCII: cv2 draws one '?' per BYTE of the UTF-8 text (cv: readCheck), the reference never raises
    for text in ("\xe9", "\u65e5\u672c", "a\tb", "\x7f", "caf\xe9 #3"):
        a = np.zeros((120, 400, 3), np.uint8)
        b = a.copy()
        cv2.putText(a, text, (10, 90), cv2.FONT_HERSHEY_SIMPLEX, 1.0, (255, 200, 50), 2)
        assert L.ocv_put_text(P(b), 120, 400, b.strides[0], text.encode("utf-8"), 10, 90, 1.0, 255, 200, 50, 2) == 0
        assert np.array_equal(a, b), text
        tw, th = ctypes.c_int(), ctypes.c_int()
        assert L.ocv_get_text_size(text.encode("utf-8"), 1.0, 2, ctypes.byref(tw), ctypes.byref(th)) == 0
        assert (tw.value, th.value) == cv2.getTextSize(text, cv2.FONT_HERSHEY_SIMPLEX, 1.0, 2)[0], text


def test_text_labels_against_cv2_calls():
    """draw_bounding_boxes with labels other than '#<int>' (any printable ASCII): oracle render == the same cv2 calls."""
    pytest.importorskip("cv2")
    from vision_inspection_system_b200 import synth
    labels = ["#A7", "crack-12", "Z", "a|b", "#(x)", "Q9%", "~", "a label far longer than eleven bytes", "d\xe9faut #2",
              "\u6b20\u9665", ""]
    for seed, shape in ((11, (480, 640)), (12, (1080, 1920))):
        frame, boxes = synth.annotated_frame(seed, *shape)
        for i, b in enumerate(boxes):
            b["label"] = labels[(seed + i) % len(labels)]
        px = OV.select_boxes(boxes, shape[1], shape[0])
        assert np.array_equal(OV.render(frame, px), OV.render_cv2(frame, px)), seed


def test_sine_table_landmarks():
    # the table the oracle derives must hold the 7-decimal values OpenCV ships
    L = lib()
    img = np.zeros((64, 64, 3), np.uint8)
    L.ocv_circle(P(img), 64, 64, img.strides[0], 32, 32, 20, 255, 255, 255, 3, 8)     # forces init_sin
    tab = (ctypes.c_float * 451).in_dll(L, "SinTable") if hasattr(L, "SinTable") else None
    assert tab is None or abs(tab[30] - 0.5) < 1e-7
