"""Recovers the Hershey simplex glyph strings of printable ASCII from the installed OpenCV binary (build container only).

cv2.putText draws g_HersheyGlyphs[HersheySimplex[c - ' ' + 1]]; the glyph strings are plain string literals in the
shared object.  For every character this script renders each candidate literal with putText's own geometry through
cv2.polylines (shift = 16) and keeps the literal that reproduces cv2.putText at four scale / thickness / line-type
settings and matches cv2.getTextSize.  Output: the 95-entry table embedded in oracle/cvdraw_oracle.c
(kSimplexGlyphs) and vision-inspection-system_b200/csrc/vis_overlay_host.cpp (glyph_for).

    python tests/golden/find_glyphs.py > glyphs.json
"""
import json
import os
import re

import cv2
import numpy as np

ONE = 1 << 16


def candidates():
    so = os.path.join(os.path.dirname(cv2.__file__), [f for f in os.listdir(os.path.dirname(cv2.__file__)) if f.endswith(".so")][0])
    data = open(so, "rb").read()
    out = []
    for probe in (b"H\\NJPISFS[", b"H\\QFNGLJKOKRLWNZQ[S[VZXWYRYOXJVGSFQF"):     # short / long literal pools ('1', '0')
        at = data.find(probe)
        seg = data[max(0, at - 120000):at + 120000]
        out += [m.group()[:-1] for m in re.finditer(rb"[ -~]{2,}\x00", seg)]
    return list(dict.fromkeys(out))


def render_glyph(g, fs, t, lt):
    img = np.zeros((160, 160), np.uint8)
    hs = int(np.rint(fs * ONE))
    vx, vy = (30 << 16) - (g[0] - ord("R")) * hs, (120 << 16) - 9 * hs
    pts, i = [], 2
    while True:
        if i >= len(g) or g[i] == 32:
            if len(pts) > 1:
                cv2.polylines(img, [np.array(pts, np.int32).reshape(-1, 1, 2)], False, 255, t, lt, 16)
            pts = []
            if i >= len(g):
                break
            i += 1
        else:
            if i + 1 >= len(g):
                break
            pts.append(((g[i] - ord("R")) * hs + vx, (g[i + 1] - ord("R")) * hs + vy))
            i += 2
    return img


def main():
    cand = candidates()
    settings = [(1.7, 2, cv2.LINE_8), (3.3, 1, cv2.LINE_8), (0.9, 1, cv2.LINE_AA), (4.1, 1, cv2.LINE_AA)]
    table = {}
    for c in range(32, 127):
        (tw, _), _ = cv2.getTextSize(chr(c), cv2.FONT_HERSHEY_SIMPLEX, 1.0, 1)
        hits = [g for g in cand if len(g) >= 2 and (g[1] - g[0]) == tw - 1]
        for fs, t, lt in settings:
            ref = np.zeros((160, 160), np.uint8)
            cv2.putText(ref, chr(c), (30, 120), cv2.FONT_HERSHEY_SIMPLEX, fs, 255, t, lt)
            hits = [g for g in hits if np.array_equal(render_glyph(g, fs, t, lt), ref)]
        if c == 32:
            hits = [b"JZ"]                                   # the blank glyph: bounds only (advance 16), merged into another literal
        assert hits, chr(c)
        table[c] = hits[0].decode()                          # several literals can render identically ('X': stroke direction)
    print(json.dumps({chr(c): g for c, g in table.items()}, indent=0))


if __name__ == "__main__":
    main()
