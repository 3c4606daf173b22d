"""Host side of the product (the [host] entry points of libvis_b200.so + Python size rules) against the oracle.
No GPU needed: these are independent implementations of the same Pillow / OpenCV / transformers rules."""
import hashlib

import numpy as np
import pytest

from leaf_model import apply_leaves
from oracle import overlay as OV
from oracle import qwen2vl as Q
from vision_inspection_system_b200 import _native as N
from vision_inspection_system_b200 import geometry as G
from vision_inspection_system_b200 import overlay as PO
from vision_inspection_system_b200 import synth
from vision_inspection_system_b200 import tables as T

PAIRS = [(1920, 1316), (1080, 728), (3840, 1316), (2160, 728), (1920, 1932), (1080, 1092), (502, 504), (100, 112),
         (3840, 2048), (2160, 1152), (3840, 1024), (640, 644), (480, 476), (20, 28), (3000, 2800), (1, 56), (56, 56),
         (10, 56), (2048, 1316), (1152, 728), (517, 700), (333, 200)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("filt", [N.FILTER_BICUBIC, N.FILTER_LANCZOS])
def test_coefficients_match_oracle(filt):
    for i, o in PAIRS:
        t = T.coeff_table(i, o, filt)
        k, b, ks = Q.coeffs(i, o, filt)
        assert t.ksize == ks, (i, o)
        assert np.array_equal(t.bounds, b), (i, o)
        assert np.array_equal(t.k, k), (i, o)
        assert t.max_taps == int(b[:, 1].max())


def test_coefficient_edge_cases():
    assert N.lib().vis_coeff_ksize(0, 10, N.FILTER_BICUBIC) == N.VIS_E_INVALID
    assert N.lib().vis_coeff_ksize(10, 10, 7) == N.VIS_E_INVALID
    assert b"bad arguments" in N.lib().vis_last_error()
    t = T.coeff_table(56, 56, N.FILTER_BICUBIC)          # identity geometry: a single unit tap per output
    assert all(np.count_nonzero(row) == 1 and row.max() == 1 << 22 for row in t.k)


def test_lut_matches_oracle_and_numpy():
    lut = T.normalize_lut()
    assert np.array_equal(lut, Q.normalize_lut())
    v = np.arange(256, dtype=np.uint8)
    want = ((v.astype(np.float64) * (1 / 255)).astype(np.float32)[:, None] - np.asarray(G.IMAGE_MEAN, np.float32)) \
        / np.asarray(G.IMAGE_STD, np.float32)
    assert np.array_equal(lut.reshape(256, 3), want.astype(np.float32))


def test_size_rules_match_oracle(goldens):
    rng = np.random.default_rng(3)
    for _ in range(3000):
        h, w = int(rng.integers(1, 5000)), int(rng.integers(1, 5000))
        for mx in (G.DEFAULT_MAX_PIXELS, G.HUB_MAX_PIXELS):
            try:
                want = Q.smart_resize(h, w, max_pixels=mx)
            except ValueError:
                with pytest.raises(ValueError):
                    G.smart_resize(h, w, max_pixels=mx)
                continue
            assert G.smart_resize(h, w, max_pixels=mx) == want
        for limit in (1024, 2048, 256):
            assert G.thumbnail_size(w, h, limit) == Q.thumbnail_size(w, h, limit)
            assert G.resize_image_size(w, h, limit) == Q.resize_image_size(w, h, limit)
    for rec in goldens["thumbnail"]:
        h, w = rec["shape"]
        assert list(G.thumbnail_size(w, h, rec["limit"])) == rec["thumb_size"]
        assert list(G.resize_image_size(w, h, rec["limit"])) == rec["resize_image_size"]
    assert G.pil_pass_order(3000, 20, 2800, 28) == "vh" and G.pil_pass_order(1080, 1920, 728, 1316) == "hv"
    assert G.pil_pass_order(56, 56, 56, 56) == "" and G.pil_pass_order(56, 60, 56, 56) == "h"


def test_records_and_strip_plans():
    for (sh, sw), (dh, dw) in [((1080, 1920), (728, 1316)), ((2160, 3840), (728, 1316)), ((100, 502), (112, 504)),
                               ((1080, 1920), (1092, 1932)), ((2160, 3840), (2156, 3836))]:
        ht, vt = T.coeff_table(sw, dw, N.FILTER_BICUBIC), T.coeff_table(sh, dh, N.FILTER_BICUBIC)
        kt = T.kt_class(max(ht.max_taps, vt.max_taps))
        assert kt in (0, 6, 8)
        if kt == 0:                               # 9+ taps: scheduled 16-slot kernel or generic passes (tests/test_sched.py)
            assert max(ht.max_taps, vt.max_taps) > 8
            continue
        rec = T.pack_records(ht, kt)
        stride = rec.shape[1]
        assert rec.shape[0] == dw + 1 and stride % 4 == 0 and stride >= kt + 2
        for o in (0, 1, dw // 2, dw - 1):
            first, taps = ht.bounds[o]
            assert rec[o, stride - 2] == first and rec[o, stride - 1] == first + taps - 1
            assert np.array_equal(rec[o, :taps], ht.k[o, :taps][::-1])          # newest tap first
            assert not rec[o, taps:kt].any()
        assert rec[dw, stride - 1] == np.iinfo(np.int32).max                      # sentinel
        assert (np.diff(rec[:dw, stride - 1]) >= 0).all()                         # push order needs monotone ends
        for vs in (1, 3):
            plan = T.plan_strips(dh, dw, ht, kt, vs)
            s = plan.strips
            cover = np.zeros((dh // 14, dw // 28), np.int32)
            for r in s:
                assert r["x0"] % 28 == 0 and r["x1"] % 28 == 0 and r["y0"] % 14 == 0 and r["y1"] % 14 == 0
                assert 0 < r["x1"] - r["x0"] <= plan.strip_w <= 336
                cover[r["y0"] // 14:r["y1"] // 14, r["x0"] // 28:r["x1"] // 28] += 1
            assert (cover == 1).all()                                              # exact tiling, no overlap
    assert T.kt_class(17) == 0 and T.kt_class(9) == 0 and T.kt_class(8) == 8


def test_boxes_to_pixels_matches_oracle(goldens):
    for rec in goldens["overlay"]:
        h, w = rec["shape"]
        want = OV.select_boxes(rec["boxes"], w, h, rec["confidence_threshold"], rec["criticality"])
        got = PO.boxes_to_pixels(rec["boxes"], w, h, rec["confidence_threshold"], rec["criticality"])
        assert len(got) == len(want), rec["name"]
        for g, e in zip(got, want):
            assert (g["x"], g["y"], g["w"], g["h"]) == (e.x, e.y, e.w, e.h)
            assert (g["b"], g["g"], g["r"]) == e.color and bool(g["dashed"]) == e.dashed
            assert PO.box_label(g).decode() == e.label


def test_leaf_expansion_reproduces_reference_overlays(goldens):
    """vis_overlay_expand + the per-pixel in-order leaf semantics == the reference's cv2 drawing (golden sha256)."""
    for rec in goldens["overlay"]:
        h, w = rec["shape"]
        if h * w > 1080 * 1920:
            continue                                   # keep the CPU suite short; the GPU suite covers 4K
        frame = np.random.default_rng(rec["seed"]).integers(0, 256, (h, w, 3), dtype=np.uint8)
        px = PO.boxes_to_pixels(rec["boxes"], w, h, rec["confidence_threshold"], rec["criticality"])
        leaves = PO.expand_leaves(px, w, h)
        got = apply_leaves(frame, leaves)
        assert sha(got) == rec["sha256"], rec["name"]


def test_leaf_expansion_edge_cases():
    assert len(PO.expand_leaves(np.zeros(0, N.BOX_DTYPE), 640, 480)) == 0
    frame, _ = synth.annotated_frame(5, 300, 500)
    boxes = [{"x": 0, "y": 0, "width": 100, "height": 50, "label": "#1", "confidence": "low"},       # area 50: kept
             {"x": 90, "y": 90, "width": 10, "height": 10, "label": "#99"}]                             # ends at 100 %
    px = PO.boxes_to_pixels(boxes, 500, 300)
    assert len(px) == 2
    got = apply_leaves(frame, PO.expand_leaves(px, 500, 300))
    assert np.array_equal(got, OV.draw_bounding_boxes(frame, boxes))
    # labels are any printable ASCII (full Hershey simplex table); anything else is refused, not guessed
    texts = ["#A7", "crack-12", "Z", "a|b", "#(x)", "Q9%", "~", "ok!", "{[<>]}"]
    for k, shape in enumerate(((300, 500), (480, 640), (1080, 1920))):
        frame, boxes = synth.annotated_frame(50 + k, *shape)
        for i, b in enumerate(boxes):
            b["label"] = texts[(3 * k + i) % len(texts)]
        px = PO.boxes_to_pixels(boxes, shape[1], shape[0])
        got = apply_leaves(frame, PO.expand_leaves(px, shape[1], shape[0]))
        assert np.array_equal(got, OV.draw_bounding_boxes(frame, boxes)), shape
    # like the reference, nothing about a label raises: any length, any text ('?' per byte outside ASCII, as cv2), even a
    # label wider than the frame (clipped) or a non-string (utils/image_utils.py:240-244 falls back to the index)
    frame, _ = synth.annotated_frame(77, 300, 500)
    odd = ["#\xe9", "a label far longer than eleven bytes", "\u6b20\u9665 #4", "", "x" * 300, 17]
    boxes = [{"x": 5 + 14 * i, "y": 8 + 12 * i, "width": 20, "height": 20, "label": t,
              "confidence": "low" if i % 2 else "high"} for i, t in enumerate(odd)]
    px = PO.boxes_to_pixels(boxes, 500, 300)
    assert len(px) == len(odd) and PO.box_label(px[0]) == "\xe9".encode("utf-8") and PO.box_label(px[5]) == b"6"
    got = apply_leaves(frame, PO.expand_leaves(px, 500, 300))
    assert np.array_equal(got, OV.draw_bounding_boxes(frame, boxes))


def test_overlay_plan_batch_equals_per_frame_planning():
    """vis_overlay_plan_batch (threaded) = vis_overlay_expand + vis_overlay_tiles per frame, concatenated."""
    import ctypes as C
    from vision_inspection_system_b200 import _native as N
    from vision_inspection_system_b200 import overlay as O
    from vision_inspection_system_b200 import synth
    shapes = [(1080, 1920), (480, 640), (333, 517), (1080, 1920), (97, 211), (2160, 3840)] * 4
    items = [synth.annotated_frame(900 + i, *s)[1] for i, s in enumerate(shapes)]
    items[2] = []                                               # a frame without boxes
    px = [O.boxes_to_pixels(b, w, h, "low", "medium") for (h, w), b in zip(shapes, items)]
    box_begin = np.concatenate([[0], np.cumsum([len(p) for p in px])]).astype(np.int32)
    boxes = np.concatenate([p for p in px if len(p)])
    hw = np.asarray(shapes, np.int32)
    n = len(shapes)
    needed = np.zeros(3, np.int64)
    leaf_begin = np.zeros(n + 1, np.int32)
    L = N.lib()
    rc = L.vis_overlay_plan_batch(n, hw.ctypes.data_as(C.c_void_p), boxes.ctypes.data_as(C.c_void_p),
                                  box_begin.ctypes.data_as(C.c_void_p), None, 0, leaf_begin.ctypes.data_as(C.c_void_p),
                                  None, 0, None, 0, needed.ctypes.data_as(C.c_void_p), 3)
    assert rc == N.VIS_E_CAPACITY and (needed > 0).all()
    leaves = np.zeros(int(needed[0]), N.LEAF_DTYPE)
    tiles = np.zeros(int(needed[1]), N.OVERLAY_TILE_DTYPE)
    refs = np.zeros(int(needed[2]), N.OVERLAY_REF_DTYPE)
    rc = L.vis_overlay_plan_batch(n, hw.ctypes.data_as(C.c_void_p), boxes.ctypes.data_as(C.c_void_p),
                                  box_begin.ctypes.data_as(C.c_void_p), leaves.ctypes.data_as(C.c_void_p), len(leaves),
                                  leaf_begin.ctypes.data_as(C.c_void_p), tiles.ctypes.data_as(C.c_void_p), len(tiles),
                                  refs.ctypes.data_as(C.c_void_p), len(refs), needed.ctypes.data_as(C.c_void_p), 5)
    assert rc == len(tiles)
    at_t = at_r = 0
    for i, ((h, w), p) in enumerate(zip(shapes, px)):
        want = O.expand_leaves(p, w, h)
        assert np.array_equal(leaves[leaf_begin[i]:leaf_begin[i + 1]]["w"], want["w"]), i
        t3, r2 = O.touched_tiles(want, len(p), w, h)
        got_t = tiles[at_t:at_t + len(t3)]
        assert (got_t["frame"] == i).all() and np.array_equal(got_t["txy"], t3[:, 0])
        assert np.array_equal(got_t["ref_begin"] - at_r, t3[:, 1]) and np.array_equal(got_t["ref_end"] - at_r, t3[:, 2])
        got_r = refs[at_r:at_r + len(r2)]
        assert np.array_equal(got_r["leaf_begin"], r2[:, 0]) and np.array_equal(got_r["leaf_end"], r2[:, 1])
        at_t += len(t3)
        at_r += len(r2)
    assert at_t == len(tiles) and at_r == len(refs) and leaf_begin[n] == len(leaves)


def test_host_only_image_utils_helpers(tmp_path):
    """load_image / get_image_info / validate_image: same results and messages as the reference's host code
    (utils/image_utils.py:20-43, 81-145); no GPU involved."""
    from PIL import Image
    from vision_inspection_system_b200 import image_utils as IU
    p = tmp_path / "frame.png"
    Image.fromarray(np.zeros((20, 30, 3), np.uint8)).save(p)
    info = IU.get_image_info(p)
    assert (info["width"], info["height"], info["mode"], info["format"], info["filename"]) == (30, 20, "RGB", "PNG", "frame.png")
    assert info["size_bytes"] == p.stat().st_size and info["path"] == str(p)
    assert IU.validate_image(p) == (True, None)
    assert IU.validate_image(tmp_path / "none.png") == (False, "File does not exist")
    bad_ext = tmp_path / "frame.gif"
    bad_ext.write_bytes(p.read_bytes())
    ok, msg = IU.validate_image(bad_ext)
    assert not ok and msg.startswith("Invalid extension 'gif'")
    tiny = tmp_path / "tiny.png"
    Image.fromarray(np.zeros((5, 30, 3), np.uint8)).save(tiny)
    assert IU.validate_image(tiny) == (False, "Image too small (minimum 10x10 pixels)")
    junk = tmp_path / "junk.jpg"
    junk.write_bytes(b"not an image")
    ok, msg = IU.validate_image(junk)
    assert not ok and msg.startswith("Invalid image file: Failed to load image")
    assert IU.validate_image(p, max_size_mb=1e-9)[1].startswith("File too large")
    with pytest.raises(FileNotFoundError):
        IU.load_image(tmp_path / "none.png")
    with pytest.raises(ValueError, match="Failed to load image"):
        IU.load_image(junk)


def test_exif_orientation_parser():
    """jpeg.exif_orientation: the tag cv2.imread honours; 1 for files without EXIF or with damaged segments."""
    import io
    from PIL import Image
    from vision_inspection_system_b200.jpeg import exif_orientation
    im = Image.fromarray(np.random.default_rng(0).integers(0, 255, (40, 60, 3), dtype=np.uint8))
    plain = io.BytesIO()
    im.save(plain, "JPEG")
    assert exif_orientation(plain.getvalue()) == 1
    for o in range(1, 9):
        b = io.BytesIO()
        ex = Image.Exif()
        ex[0x0112] = o
        im.save(b, "JPEG", exif=ex)
        assert exif_orientation(b.getvalue()) == o
    data = b.getvalue()
    assert exif_orientation(data[:30]) == 1 and exif_orientation(b"") == 1 and exif_orientation(b"\xff\xd8\xff\xe1\x00") == 1
