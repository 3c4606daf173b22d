"""CPU: the comparison-panel / status-stamp oracle (oracle/compare.py, cvresize_oracle.c) against the installed cv2
binary and against arrays captured from the reference's own create_side_by_side_comparison / create_status_stamp;
the product's host logic (tables, text size, draw-list expansion) against the same.  No GPU."""
import ctypes
import hashlib

import numpy as np
import pytest

from leaf_model import apply_leaves
from oracle import compare as OC
from oracle import lib as oracle_lib
from vision_inspection_system_b200 import _native as N
from vision_inspection_system_b200 import compare as CP
from vision_inspection_system_b200 import synth

cv2 = pytest.importorskip("cv2")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_resize_linear_matches_cv2():
    rng = np.random.default_rng(11)
    cases = [(1080, 1920, 800, 1422), (1600, 1200, 800, 600), (1600, 1201, 800, 600), (480, 640, 800, 1066),
             (800, 600, 800, 600), (100, 502, 800, 4016), (7, 5, 800, 571), (1, 1, 800, 800), (3200, 100, 800, 25),
             (4, 4, 2, 2), (5, 4, 2, 2), (2, 2, 1, 1)]
    cases += [tuple(int(v) for v in rng.integers(1, 200, 4)) for _ in range(60)]
    for sh, sw, dh, dw in cases:
        for cn in (3, 4):
            img = rng.integers(0, 256, (sh, sw, cn), dtype=np.uint8)
            want = cv2.resize(img, (dw, dh))
            assert np.array_equal(OC.resize_linear(img, dw, dh), want), (sh, sw, dh, dw, cn)


def test_side_by_side_matches_cv2_calls():
    rng = np.random.default_rng(12)
    for (h1, w1), (h2, w2), labels in [((480, 640), (480, 640), OC.LABELS), ((333, 517), (1080, 1920), ("a|~Q", "R{[x]}")),
                                       ((1600, 1200), (800, 900), ("", "W"))]:
        a = rng.integers(0, 256, (h1, w1, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (h2, w2, 3), dtype=np.uint8)
        assert np.array_equal(OC.side_by_side(a, b, labels), OC.side_by_side_cv2(a, b, labels))


def test_stamp_matches_cv2_calls():
    for verdict in ("SAFE", "UNSAFE", "REQUIRES_HUMAN_REVIEW"):
        for size in ((300, 100), (200, 80), (640, 200), (97, 41)):
            assert np.array_equal(OC.status_stamp(verdict, size), OC.status_stamp_cv2(verdict, size)), (verdict, size)


def test_oracle_matches_reference_goldens(goldens):
    for rec in goldens["compare"]["side_by_side"]:
        a = synth.noise_frame(*rec["original"])
        b = synth.noise_frame(*rec["annotated"])
        got = OC.side_by_side(a, b, tuple(rec["labels"]) if rec["labels"] else OC.LABELS)
        assert list(got.shape) == rec["shape"] and sha(got) == rec["sha256"], rec["name"]
    for rec in goldens["compare"]["stamp"]:
        got = OC.status_stamp(rec["verdict"], tuple(rec["size"]))
        assert sha(got) == rec["sha256"] and int((got[:, :, 3] > 0).sum()) == rec["opaque_pixels"], rec


def test_host_tables_and_modes_match_the_oracle():
    """vis_resize_linear_mode against the oracle's; vis_linear_table through a numpy evaluation of the bilinear
    formula against cv2 itself."""
    L, V = oracle_lib(), N.lib()
    rng = np.random.default_rng(13)
    for i in range(1500):
        sh, sw, dh, dw = (int(v) for v in rng.integers(1, 60, 4))
        if i % 3 == 0:
            dh, dw = max(sh // 2, 1), max(sw // 2, 1)
        assert L.ocv_resize_linear_mode(sh, sw, dh, dw) == V.vis_resize_linear_mode(sh, sw, dh, dw)
    assert V.vis_resize_linear_mode(0, 4, 2, 2) == N.VIS_E_INVALID
    for sh, sw, dh, dw in [(1080, 1920, 800, 1422), (100, 502, 800, 4016), (37, 53, 91, 17), (5, 3, 9, 200)]:
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8).astype(np.int64)
        xofs, alpha = CP.linear_tables(sw, dw, True)
        yofs, beta = CP.linear_tables(sh, dh, False)
        a = alpha.astype(np.int64)
        x1 = np.minimum(xofs + 1, sw - 1)
        rows = img[:, xofs] * a[None, :, :1] + img[:, x1] * a[None, :, 1:]
        r0 = rows[np.clip(yofs, 0, sh - 1)] >> 4
        r1 = rows[np.clip(yofs + 1, 0, sh - 1)] >> 4
        b = beta.astype(np.int64)
        got = (((b[:, None, :1] * r0) >> 16) + ((b[:, None, 1:] * r1) >> 16) + 2) >> 2
        assert np.array_equal(got.astype(np.uint8), cv2.resize(img.astype(np.uint8), (dw, dh))), (sh, sw, dh, dw)


def test_text_size_matches_cv2():
    for text, fs, th in [("Original Input", 0.7, 2), ("AI Analysis Layer", 0.7, 2), ("PASSED", 1.5, 4),
                         ("REJECTED", 1.5, 4), ("REVIEW", 1.5, 4), ("", 1.0, 1), ("a|~Q{", 2.3, 3),
                         ("caf\xe9", 1.0, 1), ("\u6b20\u9665 panel", 0.7, 2), ("y" * 200, 0.7, 2)]:     # '?' per non-ASCII byte, any length
        assert CP.text_size(text, fs, th) == cv2.getTextSize(text, cv2.FONT_HERSHEY_SIMPLEX, fs, th)[0], text


def test_draw_list_expansion_reproduces_the_oracle():
    """vis_draw_expand + the numpy leaf model (what the CUDA kernel computes per pixel) against the oracle."""
    for verdict in ("SAFE", "UNSAFE", "other"):
        for size in ((300, 100), (200, 80), (97, 41)):
            w, h = size
            leaves = CP.expand_commands(CP.stamp_commands(verdict, w, h), w, h)
            got = apply_leaves(np.zeros((h, w, 4), np.uint8), leaves)
            assert np.array_equal(got, OC.status_stamp(verdict, size)), (verdict, size)
    left_w, right_w = 1422, 1066
    header = np.full((CP.HEADER_HEIGHT, left_w + CP.DIVIDER_WIDTH + right_w, 3), CP.BAR_GRAY, np.uint8)
    leaves = CP.expand_commands(CP.header_commands(left_w, right_w), header.shape[1], header.shape[0])
    want = OC.side_by_side(np.zeros((1080, 1920, 3), np.uint8), np.zeros((480, 640, 3), np.uint8))[:CP.HEADER_HEIGHT]
    assert np.array_equal(apply_leaves(header, leaves), want)


def test_draw_list_rejects_bad_commands():
    bad = np.zeros(1, N.DRAW_CMD_DTYPE)
    bad["kind"] = 9
    with pytest.raises(N.VisError):
        CP.expand_commands(bad, 64, 64)
    # texts of any length are drawable (the reference's cv2.putText takes any label)
    long_cmd = CP.commands([CP._cmd(N.DRAW_TEXT, 2, 40, 0, 0, 2, (255, 255, 255), font_scale=0.5, text="x" * 100)])
    got = apply_leaves(np.zeros((64, 900, 3), np.uint8), CP.expand_commands(long_cmd, 900, 64))
    want = np.zeros((64, 900, 3), np.uint8)
    cv2.putText(want, "x" * 100, (2, 40), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 255, 255), 2)
    assert np.array_equal(got, want)
