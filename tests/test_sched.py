"""Host logic of the statically scheduled kernel: vis_sched_build's masks must say exactly where Pillow's tap windows
end (utils: tf:image_transforms.py:367 -> Pillow precompute_coeffs bounds), checked by replaying the device loops on the CPU."""
import ctypes as C

import numpy as np
import pytest

from vision_inspection_system_b200 import _native as N
from vision_inspection_system_b200 import geometry as G
from vision_inspection_system_b200 import tables as T

SUB = np.dtype([("xa", np.uint16), ("xb", np.uint16), ("p0", np.uint16), ("nsteps", np.uint16), ("mask_off", np.uint16),
                ("pad", np.uint16)])
STRIP = np.dtype([("x0", np.int32), ("x1", np.int32), ("px0", np.int32), ("row_bytes", np.int32)])
SEG = np.dtype([("y0", np.int32), ("y1", np.int32), ("r_first", np.int32), ("r_end", np.int32), ("mask_off", np.int32),
                ("pad", np.int32)])
SCHED = np.dtype([("head", N.SCHED_HEAD_DTYPE), ("strip", STRIP, (16,)), ("sub", SUB, (16, 20)), ("seg", SEG, (16,)),
                  ("mask", np.uint8, (6144,))], align=True)
SEG_COUNT = 16


def sched_ends_and_records(table, kt, per_index=1):
    """vis_sched_pack_records, checked: window ends strictly increase, stay within kt slots of the first tap, differ
    from Pillow's only where the far border clamps, and the shifted coefficients are Pillow's."""
    L = N.lib()
    stride = L.vis_record_stride(kt)
    rec = np.zeros((table.out_size + 1, stride), np.int32)
    assert L.vis_sched_pack_records(table.out_size, N.i32ptr(table.k), N.i32ptr(table.bounds), table.ksize, kt,
                                    per_index, N.i32ptr(rec), rec.size) == N.VIS_OK
    first, taps = table.bounds[:, 0], table.bounds[:, 1]
    ends = rec[:-1, stride - 1]
    assert (np.diff(ends) >= 0).all() and (ends - first + 1 <= kt).all() and (rec[:-1, stride - 2] == first).all()
    assert np.unique(ends, return_counts=True)[1].max() <= per_index
    true_last = first + taps - 1
    moved = ends != true_last
    assert (true_last[moved] == table.in_size - 1).all() and (ends >= true_last).all()
    for o in range(table.out_size):
        sh = ends[o] - true_last[o]
        want = np.zeros(kt, np.int32)
        want[sh:sh + taps[o]] = table.k[o, :taps[o]][::-1]
        assert (rec[o, :kt] == want).all()
    return ends, rec


def build(src_h, src_w, dst_h, dst_w, pitch, vsplit, filt=N.FILTER_BICUBIC, out_mode=N.SCHED_OUT_PIXEL_VALUES):
    L = N.lib()
    assert SCHED.itemsize == L.vis_sched_sizeof()
    ht = T.coeff_table(src_w, dst_w, filt)
    vt = T.coeff_table(src_h, dst_h, filt)
    buf = np.zeros(1, SCHED)
    rc = L.vis_sched_build(src_h, src_w, dst_h, dst_w, pitch, N.i32ptr(ht.bounds), N.i32ptr(vt.bounds), vsplit, out_mode,
                           buf.ctypes.data_as(C.c_void_p))
    return rc, buf[0], ht, vt


def read_mask(mask, off, index, ring):
    """(first-sample mask, second-sample mask) of step / group `index`: ring / 8 bytes each, little endian."""
    mb = ring // 8
    at = int(off) + 2 * mb * index
    return (int.from_bytes(bytes(mask[at:at + mb]), "little"), int.from_bytes(bytes(mask[at + mb:at + 2 * mb]), "little"))


def replay(s, ht, vt, w, pitch, want_per, want_ring, unit):
    hd = s["head"]
    ring, per, n_subs = int(hd["ring"]), int(hd["per_index"]), int(hd["n_subs"])
    assert (per, ring) == (want_per, want_ring) and n_subs == 12
    dh, dw = int(hd["dst_h"]), int(hd["dst_w"])
    hlast, _ = sched_ends_and_records(ht, int(hd["kt"]), per)
    vlast, _ = sched_ends_and_records(vt, int(hd["kt"]), per)
    # horizontal: every output column of every strip is emitted exactly once, at the pixel where its window ends,
    # after all of its taps have been read (p0 <= first tap), and the staged row segment covers the window
    covered = np.zeros(dw, np.int32)
    h_pull = int(hd["h_pull"])
    assert h_pull == (1 if hd["kt"] > 16 and not hd["dp_words"] else 0)       # the packed-byte kernel pushes every class
    if h_pull:                                            # pull-order H: no step masks, only the sub-range edges
        covered[:] = 1
    for st in range(hd["n_strips"]):
        S = s["strip"][st]
        assert S["x0"] % unit == 0 and S["x1"] % unit == 0 and S["px0"] % 16 == 0 and S["row_bytes"] % 16 == 0
        assert S["x1"] - S["x0"] <= hd["max_strip_w"] <= (336 if ring == 8 else 256)
        assert S["row_bytes"] <= hd["stage_pitch"] and S["px0"] * 3 + S["row_bytes"] <= pitch
        assert [int(s["sub"][st][u]["xa"]) for u in range(1, n_subs)] == [int(s["sub"][st][u]["xb"]) for u in range(n_subs - 1)]
        assert s["sub"][st][0]["xa"] == S["x0"] and s["sub"][st][n_subs - 1]["xb"] == S["x1"]
        for u in range(n_subs):
            U = s["sub"][st][u]
            xo = int(U["xa"])
            if h_pull:
                assert U["nsteps"] == 0
                for x in range(int(U["xa"]), int(U["xb"])):      # the pulled window stays inside the staged row (+ pad in front)
                    assert (hlast[x] - (int(hd["kt"]) - 1) - S["px0"]) * 3 >= -128
                continue
            assert U["p0"] % ring == 0 and U["p0"] >= S["px0"] and U["p0"] <= ht.bounds[xo, 0]
            for i in range(U["nsteps"]):
                m1, m2 = read_mask(s["mask"], U["mask_off"], i, ring)
                assert m2 & ~m1 == 0 and (per == 2 or m2 == 0)
                for jj in range(ring):
                    for m in (m1, m2):
                        if m >> jj & 1:
                            assert hlast[xo] == U["p0"] + ring * i + jj
                            assert (min(hlast[xo], w - 1) + 1) * 3 <= S["px0"] * 3 + S["row_bytes"]
                            covered[xo] += 1
                            xo += 1
            assert xo == U["xb"]
    assert (covered == 1).all()
    # vertical: every output row once, at the input row where its window ends; segments tile [0, dst_h)
    rows = np.zeros(dh, np.int32)
    assert s["seg"][0]["y0"] == 0 and s["seg"][hd["n_segs"] - 1]["y1"] == dh
    for sg in range(hd["n_segs"]):
        Gs = s["seg"][sg]
        yo = int(Gs["y0"])
        assert Gs["y0"] % 14 == 0 and (Gs["y1"] % 14 == 0 or Gs["y1"] == dh) and Gs["r_first"] % 16 == 0 and Gs["mask_off"] % 8 == 0
        assert Gs["r_first"] <= vt.bounds[yo, 0]
        n_chunks = -(-(Gs["r_end"] - Gs["r_first"]) // 32)
        for c in range(n_chunks):
            emitted_here = 0
            for g in range(32 // ring):
                m1, m2 = read_mask(s["mask"], Gs["mask_off"], c * (32 // ring) + g, ring)
                assert m2 & ~m1 == 0 and (per == 2 or m2 == 0)
                for u in range(ring):
                    for m in (m1, m2):
                        if m >> u & 1:
                            assert vlast[yo] == Gs["r_first"] + c * 32 + g * ring + u
                            rows[yo] += 1
                            yo += 1
                            emitted_here += 1
            assert emitted_here <= 32 * per
        assert yo == Gs["y1"]
    assert (rows == 1).all()


@pytest.mark.parametrize("shape,vsplit,max_pixels,want_per,want_ring", [
    ((1080, 1920), 1, G.DEFAULT_MAX_PIXELS, 1, 8), ((1080, 1920), 3, G.DEFAULT_MAX_PIXELS, 1, 8), ((2160, 3840), 2, G.HUB_MAX_PIXELS, 1, 8),
    ((1536, 2048), 8, G.DEFAULT_MAX_PIXELS, 1, 8), ((1152, 2048), 1, G.DEFAULT_MAX_PIXELS, 1, 8), ((2048, 1536), 4, G.DEFAULT_MAX_PIXELS, 1, 8),
    ((600, 5000), 2, G.DEFAULT_MAX_PIXELS, 1, 8), ((1080, 1920), 2, G.HUB_MAX_PIXELS, 2, 8), ((720, 1280), 1, G.DEFAULT_MAX_PIXELS, 2, 8),
    ((480, 640), 3, G.DEFAULT_MAX_PIXELS, 2, 8), ((560, 1000), 1, G.DEFAULT_MAX_PIXELS, 2, 8),
    ((2160, 3840), 1, G.DEFAULT_MAX_PIXELS, 1, 16), ((2160, 3840), 16, G.DEFAULT_MAX_PIXELS, 1, 16), ((1080, 1920), 2, 250000, 1, 16),
    ((2160, 3840), 2, 250000, 1, 16), ((3100, 5500), 1, G.DEFAULT_MAX_PIXELS, 1, 16)])
def test_schedule_replays_the_tap_windows(shape, vsplit, max_pixels, want_per, want_ring):
    h, w = shape
    dh, dw = G.smart_resize(h, w, G.FACTOR, G.DEFAULT_MIN_PIXELS, max_pixels)
    pitch = (w * 3 + 15) // 16 * 16
    rc, s, ht, vt = build(h, w, dh, dw, pitch, vsplit)
    assert rc == N.VIS_OK, N.lib().vis_last_error()
    hd = s["head"]
    want_kt = max(ht.max_taps, vt.max_taps)
    assert (hd["dst_h"], hd["dst_w"]) == (dh, dw) and hd["kt"] in ((T.kt_class(want_kt),) if want_ring == 8 else ((12,) if want_kt <= 12 else (13, 14, 16) if want_kt <= 13 else (14, 16) if want_kt <= 14 else (16,) if want_kt <= 16 else (24,) if want_kt <= 24 else (32,)))
    replay(s, ht, vt, w, pitch, want_per, want_ring, 28)


@pytest.mark.parametrize("shape,out,vsplit", [((2160, 3840), (1152, 2048), 1), ((1080, 1920), (576, 1024), 5),
                                              ((1600, 1200), (1024, 768), 2), ((1536, 2048), (768, 1024), 16),
                                              ((300, 500), (153, 256), 1), ((1365, 2048), (683, 1024), 3),
                                              ((2160, 3840), (576, 1024), 4), ((2160, 3840), (864, 1536), 1),
                                              ((3000, 4000), (768, 1024), 2), ((1080, 1920), (360, 640), 1)])
def test_uint8_resize_schedule(shape, out, vsplit):
    """VIS_SCHED_OUT_U8 (LANCZOS thumbnails / resize_image): 16-slot kernel, strips in units of 4 columns, last segment
    ends at dst_h."""
    (h, w), (dh, dw) = shape, out
    pitch = (w * 3 + 15) // 16 * 16
    rc, s, ht, vt = build(h, w, dh, dw, pitch, vsplit, N.FILTER_LANCZOS, N.SCHED_OUT_U8)
    assert rc == N.VIS_OK, N.lib().vis_last_error()
    want_kt = max(ht.max_taps, vt.max_taps)
    assert s["head"]["out_mode"] == N.SCHED_OUT_U8 and s["head"]["kt"] in ((want_kt,) if want_kt in (13, 14) else (max(12, (want_kt + 3) // 4 * 4),))
    replay(s, ht, vt, w, pitch, 1, 16, 4)


@pytest.mark.parametrize("shape,max_pixels,why", [((20, 30), G.DEFAULT_MAX_PIXELS, b"upscale"), ((7000, 10000), G.DEFAULT_MAX_PIXELS, b"taps"),
                                                  ((100, 502), G.DEFAULT_MAX_PIXELS, b"pitch")])
def test_schedule_declines_what_it_cannot_express(shape, max_pixels, why):
    h, w = shape
    dh, dw = G.smart_resize(h, w, G.FACTOR, G.DEFAULT_MIN_PIXELS, max_pixels)
    pitch = w * 3 if why == b"pitch" else (w * 3 + 15) // 16 * 16
    rc, _, _, _ = build(h, w, dh, dw, pitch, 1)
    assert rc == N.VIS_E_UNSUPPORTED and why in N.lib().vis_last_error()


@pytest.mark.parametrize("shape,out,filt,mode,want_words", [
    ((2160, 3840), (728, 1316), N.FILTER_BICUBIC, N.SCHED_OUT_PIXEL_VALUES, 4),      # 13 taps
    ((2160, 3840), (1152, 2048), N.FILTER_LANCZOS, N.SCHED_OUT_U8, 4),               # 13 taps
    ((2160, 3840), (576, 1024), N.FILTER_LANCZOS, N.SCHED_OUT_U8, 7),                # 25 taps
    ((1400, 1424), (1024, 1040), N.FILTER_LANCZOS, N.SCHED_OUT_U8, 4),               # 9..10 taps
    ((1600, 2560), (640, 1024), N.FILTER_LANCZOS, N.SCHED_OUT_U8, 5),                # 17 taps
    ((3000, 4000), (768, 1024), N.FILTER_LANCZOS, N.SCHED_OUT_U8, 7)])
def test_packed_byte_records(shape, out, filt, mode, want_words):
    """VIS_SCHED_FLAG_DP4A: the schedule names the packed-byte kernel with W words per window; its records hold every
    coefficient as three byte limbs at the byte the kernel multiplies with that input sample, zero elsewhere — replayed
    here with the kernel's arithmetic (three dot products of packed words, recombined mod 2^32) on random rows."""
    import ctypes as C
    (h, w), (dh, dw) = shape, out
    pitch = (w * 3 + 15) // 16 * 16
    rc, s, ht, vt = build(h, w, dh, dw, pitch, 2, filt, mode | N.SCHED_FLAG_DP4A)
    assert rc == N.VIS_OK, N.lib().vis_last_error()
    hd = s["head"]
    words = int(hd["dp_words"])
    assert words == want_words and hd["ring"] == 16 and hd["kt"] == 4 * words - 3 and hd["h_pull"] == 0
    replay(s, ht, vt, w, pitch, 1, 16, 28 if mode == N.SCHED_OUT_PIXEL_VALUES else 4)
    L = N.lib()
    stride = L.vis_sched_record_stride_dp(words)
    assert stride >= 3 * words and stride % 4 == 0
    rng = np.random.default_rng(5)
    for t in (ht, vt):
        rec = np.zeros((t.out_size + 1, stride), np.int32)
        N.check(L.vis_sched_pack_records_dp(t.out_size, N.i32ptr(t.k), N.i32ptr(t.bounds), t.ksize, words, N.i32ptr(rec),
                                            rec.size), "vis_sched_pack_records_dp")
        by = rec.view(np.uint8).reshape(t.out_size + 1, stride * 4)
        k0 = by[:, :4 * words].astype(np.int64)
        k1 = by[:, 4 * words:8 * words].astype(np.int64)
        k2 = by[:, 8 * words:12 * words].view(np.int8).astype(np.int64)
        coeff = k0 + 256 * k1 + 65536 * k2                       # [out + 1, 4 * words] coefficient per window byte
        assert not coeff[t.out_size].any()                       # sentinel record
        ends = np.zeros(t.out_size, np.int64)                    # the schedule's window ends (same rule as the builder)
        e_prev = -1
        for o in range(t.out_size):
            e = max(int(t.bounds[o, 0] + t.bounds[o, 1] - 1), e_prev)
            if e == e_prev:
                e += 1
            ends[o], e_prev = e, e
        line = rng.integers(0, 256, t.in_size + 64, dtype=np.int64)          # samples past the border: weight 0
        for o in list(range(0, t.out_size, max(1, t.out_size // 97))) + [t.out_size - 1, t.out_size - 2]:
            first, taps = (int(v) for v in t.bounds[o])
            base = 4 * ((int(ends[o]) >> 2) - (words - 1))
            assert base <= first and first + taps - 1 <= ends[o] < base + 4 * words
            full = np.zeros(4 * words, np.int64)
            full[first - base:first - base + taps] = t.k[o, :taps]
            assert np.array_equal(coeff[o], full), (o, first, taps)
            px = np.array([line[base + j] if base + j >= 0 else 255 for j in range(4 * words)], np.int64)
            acc = (1 << 21) + int((px * k0[o]).sum()) + (int((px * k1[o]).sum()) << 8) + (int((px * k2[o]).sum()) << 16)
            want = (1 << 21) + int((line[first:first + taps] * t.k[o, :taps]).sum())
            assert acc == want


# ---------------------------------------------------------------------------------------------------------------------
# Integer tensor-path kernel (vis_fused_mma.cu): replay of its fragment indexing on the CPU.  The device gathers the B
# operand of the horizontal pass and the A operand of the vertical pass from compact per-sample records
# (vis_sched_pack_records_mma) by word arithmetic; this replays exactly that arithmetic in numpy and compares the banded
# products with Pillow's definition  clip8((sum k * p + 2^21) >> 22).
def _mma_records(table, words):
    L = N.lib()
    stride = L.vis_sched_record_stride_mma(words)
    assert stride >= 3 * words + 2 and stride % 4 == 0
    rec = np.zeros((table.out_size + 1, stride), np.int32)
    assert L.vis_sched_pack_records_mma(table.out_size, N.i32ptr(table.k), N.i32ptr(table.bounds), table.ksize, words,
                                        N.i32ptr(rec), rec.size) == N.VIS_OK
    return rec


def _limb_rows(rec_row, words):
    """(3, 4 * words) int64: the limb bytes of one record; limb 2 is signed."""
    b = rec_row[:3 * words].view(np.uint8).reshape(words, 3, 4).transpose(1, 0, 2).reshape(3, 4 * words).astype(np.int64)
    b[2] = b[2].astype(np.uint8).astype(np.int8)
    return b


def _direct_pass(table, line):
    """Pillow's 8bpc pass along the last axis of `line` ([..., in_size] uint8) by definition."""
    out = np.zeros(line.shape[:-1] + (table.out_size,), np.uint8)
    for o in range(table.out_size):
        f, n = table.bounds[o]
        acc = (line[..., f:f + n].astype(np.int64) * table.k[o, :n].astype(np.int64)).sum(-1) + (1 << 21)
        out[..., o] = np.clip(acc >> 22, 0, 255)
    return out


@pytest.mark.parametrize("src_h,src_w,dst_h,dst_w,filt,mode", [
    (2160, 3840, 756, 1316, N.FILTER_BICUBIC, N.SCHED_OUT_PIXEL_VALUES),      # 4K at the default max_pixels
    (2160, 3840, 576, 1024, N.FILTER_LANCZOS, N.SCHED_OUT_U8),                # Auditor thumbnail, 25 taps
    (1600, 2560, 640, 1024, N.FILTER_LANCZOS, N.SCHED_OUT_U8),
    (1080, 1920, 576, 1024, N.FILTER_LANCZOS, N.SCHED_OUT_U8),                # 17 rows per 32-row chunk: 28-row chunks
    (2160, 3840, 1152, 2048, N.FILTER_LANCZOS, N.SCHED_OUT_U8),
    (700, 1000, 252, 364, N.FILTER_BICUBIC, N.SCHED_OUT_PIXEL_VALUES),
])
def test_mma_fragment_indexing_replay(src_h, src_w, dst_h, dst_w, filt, mode):
    pitch = (src_w * 3 + 15) // 16 * 16
    rc, sc, ht, vt = build(src_h, src_w, dst_h, dst_w, pitch, 3, filt, mode | N.SCHED_FLAG_MMA)
    assert rc == N.VIS_OK
    head = sc["head"]
    W, KS = int(head["dp_words"]), int(head["mma_ks"])
    assert 4 <= W <= 9 and 1 <= KS <= 3 and head["ring"] == 16
    hrec, vrec = _mma_records(ht, W), _mma_records(vt, W)
    rng = np.random.default_rng(5)

    # ---- horizontal pass: tiles of 16 outputs per strip, A gathered from the records, K = 32 * KS pixels from word kw ----
    row = rng.integers(0, 256, src_w + 4 * 8 * KS + 64, dtype=np.uint8)       # reads past the row meet zero coefficients
    row[src_w:] = 255
    want = _direct_pass(ht, row[:src_w])
    got = np.zeros(dst_w, np.uint8)
    for S in sc["strip"][:head["n_strips"]]:
        x0, x1, px0 = int(S["x0"]), int(S["x1"]), int(S["px0"])
        sw = x1 - x0
        for jt in range((sw + 15) // 16):
            kw = int(hrec[x0 + 16 * jt, 3 * W + 1])
            assert 4 * kw >= px0
            for m in range(16):
                xr = x0 + min(16 * jt + m, sw - 1)
                bw = int(hrec[xr, 3 * W])
                limbs = _limb_rows(hrec[xr], W)
                col = np.zeros((3, 32 * KS), np.int64)
                for wd in range(8 * KS):                                        # word of the k window: kw + 8s + 4h + t
                    q = kw + wd - bw
                    if 0 <= q < W:
                        col[:, 4 * wd:4 * wd + 4] = limbs[:, 4 * q:4 * q + 4]
                px = row[4 * kw:4 * kw + 32 * KS].astype(np.int64)
                acc = (px * col[0]).sum() + ((px * col[1]).sum() << 8) + ((px * col[2]).sum() << 16) + (1 << 21)
                if 16 * jt + m < sw:
                    got[xr] = np.clip(acc >> 22, 0, 255)
    assert np.array_equal(got, want)

    # ---- vertical pass: chunks of 32 ring rows + carry, A gathered from the chunk's records relative to ring byte 0 ----
    carry = 4 * (W - 1)
    CR = int(head["chunk_rows"])
    assert CR in (24, 28, 32)
    col_px = rng.integers(0, 256, (5, src_h), dtype=np.uint8)
    want_v = _direct_pass(vt, col_px)
    got_v = np.zeros((5, dst_h), np.uint8)
    for G_ in sc["seg"][:head["n_segs"]]:
        y0, y1, r_first, r_end, moff = (int(G_[k]) for k in ("y0", "y1", "r_first", "r_end", "mask_off"))
        yo = y0
        for c in range((r_end - r_first + CR - 1) // CR):
            r0 = r_first + CR * c
            n = sum(bin(read_mask(sc["mask"], moff, 2 * c + i, 16)[0]).count("1") for i in range(2))
            assert n <= 32
            cbw = (r0 - carry) >> 2
            ring = np.zeros((5, 64), np.int64)                                   # ring bytes of a column: rows r0 - carry ...
            for b in range(64):
                rr = r0 - carry + b
                ring[:, b] = col_px[:, rr] if 0 <= rr < src_h else 255           # garbage wherever no real row sits
            for i in range(n):
                rel = int(vrec[yo + i, 3 * W]) - cbw
                limbs = _limb_rows(vrec[yo + i], W)
                a = np.zeros((3, 64), np.int64)
                for wd in range(16):                                             # ring word 8s + 4h + t
                    q = wd - rel
                    if 0 <= q < W:
                        a[:, 4 * wd:4 * wd + 4] = limbs[:, 4 * q:4 * q + 4]
                acc = (ring * a[0]).sum(1) + ((ring * a[1]).sum(1) << 8) + ((ring * a[2]).sum(1) << 16) + (1 << 21)
                got_v[:, yo + i] = np.clip(acc >> 22, 0, 255)
            yo += n
        assert yo == y1
    assert np.array_equal(got_v, want_v)


def test_mma_chunk_rows_follow_the_vertical_scale():
    """One M-tile of 16 output rows per chunk: 32-row chunks from vertical scale 2 on, 28 / 24 below; under 24 the
    packed-byte kernel keeps the geometry (dp_words > 0, mma_ks == 0, IDP.4A records)."""
    for (sh, swd, dh, dw, want) in ((2160, 3840, 576, 1024, 32), (1080, 1920, 576, 1024, 28), (900, 1600, 576, 1024, 24),
                                    (800, 1400, 576, 1008, 0)):
        rc, sc, _, _ = build(sh, swd, dh, dw, (swd * 3 + 15) // 16 * 16, 1, N.FILTER_LANCZOS, N.SCHED_OUT_U8 | N.SCHED_FLAG_MMA)
        assert rc == N.VIS_OK and sc["head"]["dp_words"] >= 4
        if want:
            assert sc["head"]["mma_ks"] in (1, 2, 3) and sc["head"]["chunk_rows"] == want, (sh, sc["head"])
        else:
            assert sc["head"]["mma_ks"] == 0 and sc["head"]["chunk_rows"] == 32, (sh, sc["head"])


def test_kernel_family_is_a_property_of_the_geometry():
    """The engine packs the records once per geometry: the kernel family and chunk advance that vis_sched_build picks may
    depend neither on the segment count nor on the row pitch."""
    for (sh, swd, dh, dw, filt, mode) in ((64, 96, 56, 84, N.FILTER_BICUBIC, N.SCHED_OUT_U8),
                                          (1080, 1920, 576, 1024, N.FILTER_LANCZOS, N.SCHED_OUT_U8),
                                          (2160, 3840, 756, 1316, N.FILTER_BICUBIC, N.SCHED_OUT_PIXEL_VALUES),
                                          (1536, 2048, 840, 1148, N.FILTER_BICUBIC, N.SCHED_OUT_PIXEL_VALUES)):
        seen = set()
        for pitch in ((swd * 3 + 15) // 16 * 16, (swd * 3 + 15) // 16 * 16 + 64):
            for vs in (1, 2, 3, 4, 7):
                rc, sc, _, _ = build(sh, swd, dh, dw, pitch, vs, filt, mode | N.SCHED_FLAG_MMA)
                assert rc == N.VIS_OK
                seen.add((int(sc["head"]["dp_words"]), int(sc["head"]["mma_ks"]), int(sc["head"]["chunk_rows"])))
        assert len(seen) == 1, (sh, swd, seen)


def test_mma_declines_windows_wider_than_three_k_steps():
    """A tile of 16 output columns whose window spans more than 96 input pixels (scale ~5 and up with LANCZOS / ~6 with
    BICUBIC) stays with the packed-byte kernel: vis_sched_build answers the tensor-path request with an IDP.4A schedule."""
    seen = {}
    for scale_w, filt in ((4.6, N.FILTER_LANCZOS), (5.3, N.FILTER_LANCZOS), (5.0, N.FILTER_BICUBIC), (7.4, N.FILTER_BICUBIC)):
        dw, dh = 256, 128
        sw_, sh_ = int(dw * scale_w) // 16 * 16, int(dh * 3.0)
        rc, sc, ht, _ = build(sh_, sw_, dh, dw, sw_ * 3, 2, filt, N.SCHED_OUT_U8 | N.SCHED_FLAG_MMA)
        if rc != N.VIS_OK:
            continue                                   # more than 33 taps: generic passes
        h = sc["head"]
        span = max(int(ht.bounds[min(x + 15, dw - 1), 0] + ht.bounds[min(x + 15, dw - 1), 1] - (ht.bounds[x, 0] & ~3))
                   for x in range(0, dw, 4))
        seen[(scale_w, filt)] = (int(h["mma_ks"]), span)
        assert h["dp_words"] >= 4
        assert (h["mma_ks"] > 0) == (span <= 96), (scale_w, filt, h, span)
        if h["mma_ks"]:
            assert h["mma_ks"] == (span + 31) // 32
    assert any(v[0] == 0 for v in seen.values()) and any(v[0] > 0 for v in seen.values()), seen
