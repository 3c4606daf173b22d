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
