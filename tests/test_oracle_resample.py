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
