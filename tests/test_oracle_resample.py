"""Pins the CPU oracle (oracle/resample_oracle.c + oracle/qwen2vl.py) against the golden vectors produced by the real
transformers PIL processor / Pillow / the reference's resize_image, and against the installed binaries directly."""
import hashlib

import numpy as np
import pytest

from conftest import golden_frame
from oracle import qwen2vl as Q


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_smart_resize_table():
    # SURVEY.md appendix D (values probed from transformers 5.5.0)
    table = {(1080, 1920): ((728, 1316), (1092, 1932)), (2160, 3840): ((728, 1316), (2156, 3836)),
             (720, 1280): ((728, 1288), (728, 1288)), (480, 640): ((476, 644), (476, 644)),
             (1536, 2048): ((840, 1148), (1540, 2044)), (1152, 2048): ((728, 1316), (1148, 2044)),
             (100, 502): ((112, 504), (112, 504))}
    for (h, w), (default, hub) in table.items():
        assert Q.smart_resize(h, w) == default
        assert Q.smart_resize(h, w, max_pixels=12845056) == hub
    with pytest.raises(ValueError):
        Q.smart_resize(10, 2001)
    assert Q.smart_resize(10, 2000) is not None        # ratio exactly 200 is allowed


def test_qwen_goldens(goldens, arrays):
    for rec in goldens["qwen"]:
        frame = golden_frame(rec, arrays)
        assert sha(frame) == rec["input_sha256"], rec["name"]
        kw = {} if rec["max_pixels"] is None else {"max_pixels": rec["max_pixels"]}
        pv, grid = Q.preprocess([frame], **kw)
        assert grid[0].tolist() == rec["grid_thw"], rec["name"]
        assert pv.shape == (rec["rows"], 1176)
        assert np.array_equal(pv[rec["sample_rows"]], arrays[f"qwen_{rec['name']}_samples"]), rec["name"]
        assert sha(pv) == rec["sha256"], rec["name"]


def test_mouri_known_answer(goldens, arrays):
    # Mouri.jpg -> grid [1,8,36], (288,1176), sha256[:16] 01e0d93015585789
    pv, grid = Q.preprocess([arrays["mouri_rgb"]])
    assert grid.tolist() == [[1, 8, 36]] and pv.shape == (288, 1176)
    assert sha(pv)[:16] == "01e0d93015585789"
    assert sha(arrays["mouri_rgb"])[:16] == "13b0ebd773d68238"
    assert abs(float(pv.min()) - -1.7922626) < 1e-6 and abs(float(pv.max()) - 2.145897) < 1e-6
    # the same through the FILE (tests/golden/Mouri.jpg is the reference's own fixture, 9 KB): host decode -> oracle
    from pathlib import Path
    from PIL import Image
    decoded = np.asarray(Image.open(Path(__file__).parent / "golden" / "Mouri.jpg").convert("RGB"))
    assert np.array_equal(decoded, arrays["mouri_rgb"])


def test_thumbnail_goldens(goldens, arrays):
    from vision_inspection_system_b200 import synth
    for rec in goldens["thumbnail"]:
        h, w = rec["shape"]
        frame = synth.noise_frame(rec["seed"], h, w)
        tw, th = Q.thumbnail_size(w, h, rec["limit"])
        assert [tw, th] == rec["thumb_size"], rec["name"]
        assert sha(Q.resize(frame, th, tw, Q.LANCZOS)) == rec["thumb_sha256"], rec["name"]
        rw, rh = Q.resize_image_size(w, h, rec["limit"])
        assert [rw, rh] == rec["resize_image_size"], rec["name"]
        assert sha(Q.resize(frame, rh, rw, Q.LANCZOS)) == rec["resize_image_sha256"], rec["name"]


def test_against_installed_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    cases = [((100, 502), (112, 504)), ((333, 517), (200, 700)), ((3000, 20), (2800, 28)), ((2000, 11), (1000, 28)),
             ((10, 10), (56, 56)), ((1, 1), (56, 56)), ((56, 56), (56, 56)), ((97, 1), (28, 28)), ((480, 640), (476, 644))]
    for (h, w), (oh, ow) in cases:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for filt in (Q.BICUBIC, Q.LANCZOS):
            want = np.asarray(Image.fromarray(a).resize((ow, oh), Image.Resampling(filt), reducing_gap=None))
            assert np.array_equal(Q.resize(a, oh, ow, filt), want), ((h, w), (oh, ow), filt)
    g = rng.integers(0, 256, (77, 131), dtype=np.uint8)                 # single channel
    want = np.asarray(Image.fromarray(g).resize((60, 40), Image.Resampling.LANCZOS))
    assert np.array_equal(Q.resize(g, 40, 60, Q.LANCZOS)[:, :, 0], want)


def test_against_installed_transformers():
    pytest.importorskip("transformers")
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil
    rng = np.random.default_rng(5)
    proc = Qwen2VLImageProcessorPil()
    frames = [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in [(120, 200), (64, 64), (300, 90)]]
    want = proc(images=[Image.fromarray(f) for f in frames], return_tensors="np")
    pv, grid = Q.preprocess(frames)
    assert np.array_equal(grid, want["image_grid_thw"])
    assert np.array_equal(pv, want["pixel_values"])


# ---------------------------------------------------------------- thumbnails of very large frames (reduce + boxed resample)
def test_reduce_matches_pillow():
    from PIL import Image
    rng = np.random.default_rng(41)
    for t in range(40):
        h, w = (int(v) for v in rng.integers(1, 70, 2))
        fx, fy = (int(v) for v in rng.integers(1, 8, 2))
        ch = int(rng.choice([1, 3, 4]))
        a = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
        box = None
        if t % 2:
            x0 = int(rng.integers(0, w)); x1 = int(rng.integers(x0 + 1, w + 1))
            y0 = int(rng.integers(0, h)); y1 = int(rng.integers(y0 + 1, h + 1))
            box = (x0, y0, x1, y1)
        im = Image.fromarray(a if ch != 1 else a[:, :, 0], {1: "L", 3: "RGB", 4: "RGBX"}[ch])
        want = np.asarray(im.reduce((fx, fy), box=box))
        want = want.reshape(want.shape[0], want.shape[1], -1)
        assert np.array_equal(Q.reduce(a, (fx, fy), box), want), (h, w, fx, fy, ch, box)


def test_boxed_resize_matches_pillow():
    from PIL import Image
    rng = np.random.default_rng(42)
    for t in range(40):
        h, w = (int(v) for v in rng.integers(4, 120, 2))
        oh, ow = (int(v) for v in rng.integers(1, 80, 2))
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        x0 = float(rng.uniform(0, w - 2)); x1 = float(rng.uniform(x0 + 1, w))
        y0 = float(rng.uniform(0, h - 2)); y1 = float(rng.uniform(y0 + 1, h))
        if t % 3 == 0:
            x0, y0, x1, y1 = 0, 0, w, h
        filt = (Q.LANCZOS, Q.BICUBIC)[t % 2]
        want = np.asarray(Image.fromarray(a).resize((ow, oh), filt, box=(x0, y0, x1, y1), reducing_gap=None))
        assert np.array_equal(Q.resize_box(a, oh, ow, filt, (x0, y0, x1, y1)), want), (h, w, oh, ow, x0, y0, x1, y1)


@pytest.mark.parametrize("h,w,limit", [(2160, 4096, 1024), (3000, 5000, 512), (4000, 300, 256), (1200, 1600, 128),
                                       (777, 4001, 500), (2304, 2304, 512)])
def test_thumbnail_with_reduce_prepass_matches_pillow(h, w, limit):
    """``Image.thumbnail`` from 4x downscales on: reduce by the integer factor, then LANCZOS over a fractional box."""
    from PIL import Image
    a = np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    im = Image.fromarray(a)
    im.thumbnail((limit, limit), Image.Resampling.LANCZOS)
    tw, th = Q.thumbnail_size(w, h, limit)
    assert Q.reducing_plan(w, h, tw, th, Q.LANCZOS) is not None
    assert np.array_equal(Q.agent_thumbnail(a, limit), np.asarray(im))


def test_product_boxed_coefficients_and_plan_match_the_oracle():
    """vis_build_coeffs_box (product, host) against orc_coeffs_box (oracle); geometry.reducing_plan against the oracle's."""
    import ctypes as C
    from oracle import lib as oracle_lib
    from vision_inspection_system_b200 import geometry as G
    from vision_inspection_system_b200 import tables as T
    L = oracle_lib()
    rng = np.random.default_rng(43)
    for t in range(60):
        in_size, out_size = int(rng.integers(2, 3000)), int(rng.integers(1, 900))
        in0 = float(np.float32(rng.uniform(0, in_size - 1.5)))
        in1 = float(np.float32(rng.uniform(in0 + 1, in_size)))
        filt = (Q.LANCZOS, Q.BICUBIC)[t % 2]
        tab = T.coeff_table_box(in_size, in0, in1, out_size, filt)
        ks = L.orc_ksize_box(in0, in1, out_size, filt)
        k = np.zeros((out_size, ks), np.int32)
        b = np.zeros((out_size, 2), np.int32)
        assert L.orc_coeffs_box(in_size, in0, in1, out_size, filt, k.ctypes.data_as(C.POINTER(C.c_int32)),
                                b.ctypes.data_as(C.POINTER(C.c_int32))) == ks
        assert tab.ksize == ks and np.array_equal(tab.k, k) and np.array_equal(tab.bounds, b), (in_size, in0, in1, out_size)
    for w, h, limit in [(4096, 2160, 1024), (5000, 3000, 512), (300, 9000, 256), (9001, 777, 500), (3840, 2160, 1024)]:
        tw, th = G.thumbnail_size(w, h, limit)
        assert G.reducing_plan(w, h, tw, th, Q.LANCZOS) == Q.reducing_plan(w, h, tw, th, Q.LANCZOS)
    assert G.reducing_plan(3840, 2160, 1024, 576, Q.LANCZOS) is None


def test_nearest_resize_matches_pillow_for_palette_images():
    from PIL import Image
    rng = np.random.default_rng(44)
    for t in range(40):
        h, w = (int(v) for v in rng.integers(1, 300, 2))
        oh, ow = (int(v) for v in rng.integers(1, 300, 2))
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        box = None
        if t % 3 == 0:
            x0 = float(rng.uniform(0, w - 1)); x1 = float(rng.uniform(x0 + 0.5, w))
            y0 = float(rng.uniform(0, h - 1)); y1 = float(rng.uniform(y0 + 0.5, h))
            box = (x0, y0, x1, y1)
        want = np.asarray(Image.fromarray(a, "P").resize((ow, oh), Image.Resampling.LANCZOS, box=box))
        assert np.array_equal(Q.resize_nearest(a, oh, ow, box), want), (h, w, oh, ow, box)
    # product host table against the oracle's
    import ctypes as C
    from oracle import lib as oracle_lib
    from vision_inspection_system_b200 import _native as N
    for in_size, in0, in1, out in [(5000, 0.0, 5000.0, 1024), (300, 3.25, 250.5, 517), (7, 0.0, 7.0, 100), (100, 50.0, 100.0, 3)]:
        a, b = np.empty(out, np.int32), np.empty(out, np.int32)
        assert N.lib().vis_nearest_table(in_size, in0, in1, out, N.i32ptr(a)) == 0
        assert oracle_lib().orc_nearest_table(in_size, in0, in1, out, b.ctypes.data_as(C.POINTER(C.c_int32))) == 0
        assert np.array_equal(a, b)


def test_double_precision_modes_against_pillow():
    """"I;16", "I" and "F" frames: Pillow resamples them in double precision (ImagingResample*_16bpc / _32bpc), which is
    what the reference's resize_image (utils/image_utils.py:75) runs for them.  The oracle's restatement is bit-exact with
    the installed Pillow, overshoot and the 16-bit path's two CLIP8 byte writes included."""
    from PIL import Image
    rng = np.random.default_rng(11)
    cases = [("I;16", rng.integers(0, 65536, (300, 400), dtype=np.uint16)),
             ("I;16", np.where(rng.random((120, 500)) < 0.5, 0, 65535).astype(np.uint16)),      # overshoot both ways
             ("I", rng.integers(-2 ** 31, 2 ** 31 - 1, (200, 300), dtype=np.int32)),              # out of range -> INT_MIN
             ("I", rng.integers(-50000, 50000, (257, 333), dtype=np.int32)),
             ("F", rng.normal(0, 1000, (300, 400)).astype(np.float32))]
    for mode, arr in cases:
        h, w = arr.shape
        im = Image.fromarray(arr)
        assert im.mode == mode
        for (oh, ow), filt in (((h // 2, w // 2), Q.LANCZOS), ((h * 2 // 3, w - 7), Q.BICUBIC), ((h, w // 3), Q.LANCZOS),
                               ((h + 40, w + 60), Q.LANCZOS)):
            want = np.asarray(im.resize((ow, oh), Image.Resampling(filt)))
            got = Q.resize_hp(arr, oh, ow, filt)
            assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=True), (mode, oh, ow, filt)
