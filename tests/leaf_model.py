"""Test helper: a numpy model of how vis_overlay.cu applies leaf primitives (per pixel, in list order).

It lets the CPU suite check the host expansion (vis_overlay_expand) against the drawing oracle without a GPU;
the CUDA kernel itself is checked against the oracle by the -m gpu tests.
"""
import numpy as np

FILTER = np.array([
    168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
    254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
    158, 149, 140, 131, 122, 114, 105, 97, 89, 82, 75, 68, 62, 56, 50, 45,
    40, 36, 32, 28, 25, 22, 19, 16, 14, 12, 11, 9, 8, 7, 5, 5], np.int64)

GROUP, LINE8, LINEAA, TRAP, SPANS = 1, 2, 3, 4, 5
XMAJOR, AA = 0x100, 0x200


def apply_leaves(img: np.ndarray, leaves: np.ndarray) -> np.ndarray:
    out = img.copy()
    W = leaves["w"].astype(np.int64)
    n_groups = 0
    while n_groups < len(W) and (W[n_groups, 0] & 0xff) == GROUP:
        n_groups += 1
    for g in range(n_groups):
        for li in range(W[g, 2], W[g, 3]):
            _apply(out, W[li])
    return out


def _apply(out, w):
    kind = int(w[0] & 0xff)
    if kind <= GROUP:
        return
    col = np.array([(w[1] >> (8 * k)) & 0xff for k in range(out.shape[2])], np.int64)     # B, G, R (, A)
    x0, x1 = int(w[10] & 0xffff), int((w[10] >> 16) & 0xffff)
    y0, y1 = int(w[11] & 0xffff), int((w[11] >> 16) & 0xffff)
    ys, xs = np.mgrid[y0:y1 + 1, x0:x1 + 1]
    region = out[y0:y1 + 1, x0:x1 + 1]
    if kind == TRAP:
        d = ys - w[2]
        xa, xb = w[4] + w[5] * d, w[6] + w[7] * d
        lo, hi = np.minimum(xa, xb), np.maximum(xa, xb)
        aa = bool(w[0] & AA)
        xx1, xx2 = (lo + (65535 if aa else 32768)) >> 16, (hi + (0 if aa else 32768)) >> 16
        m = (d >= 0) & (ys <= w[3]) & (xs >= xx1) & (xs <= xx2)
        region[m] = col
    elif kind == SPANS:
        r = ys - w[3]
        ok = (r >= 0) & (r < w[4])
        rr = np.clip(r, 0, 15)
        hw = (w[5 + (rr >> 2)] >> (8 * (rr & 3))) & 0xff
        m = ok & (hw != 0xff) & (np.abs(xs - w[2]) <= hw)
        region[m] = col
    elif kind == LINE8:
        maj, mnr = (xs, ys) if w[0] & XMAJOR else (ys, xs)
        i = maj - w[2]
        m = (i >= 0) & (i <= w[3]) & (((w[4] + w[5] * i) >> 16) == mnr)
        m |= (xs == w[6]) & (ys == w[7])
        region[m] = col
    elif kind == LINEAA:
        maj, mnr = (xs, ys) if w[0] & XMAJOR else (ys, xs)
        s = maj - w[2]
        e = w[3] - s
        minor = w[4] + w[5] * s
        d = mnr - ((minor >> 16) - 1)
        m = (s >= 0) & (s <= w[3]) & (d >= 0) & (d <= 2)
        idx = ((((s >= 2).astype(np.int64) + 1) & (s | 2)) * 3 + (((e >= 2).astype(np.int64) + 1) & (e | 2)))
        idx = np.clip(idx, 0, 8)
        ep = (w[6 + idx // 3] >> (10 * (idx % 3))) & 0x3ff
        dist = (minor >> 11) & 31
        f = np.where(d == 0, FILTER[np.clip(dist + 32, 0, 63)], np.where(d == 1, FILTER[dist], FILTER[63 - dist]))
        a = ((ep * f) >> 8) & 0xff
        c = region.astype(np.int64)
        for _ in range(2):
            c = c + (((col[None, None, :] - c) * a[:, :, None] + 127) >> 8)
        region[m] = c[m].astype(np.uint8)
