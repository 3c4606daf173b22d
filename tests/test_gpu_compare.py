"""-m gpu: create_side_by_side_comparison / create_status_stamp on the GPU (vis_compose_panels + the draw list through
vis_overlay_draw_cn) against the oracle and arrays captured from the reference's own functions.  Bar: bit-exact."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import compare as OC
from vision_inspection_system_b200 import image_utils as IU
from vision_inspection_system_b200 import synth

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_side_by_side_goldens(engine, goldens):
    for rec in goldens["compare"]["side_by_side"]:
        a = synth.noise_frame(*rec["original"])
        b = synth.noise_frame(*rec["annotated"])
        labels = tuple(rec["labels"]) if rec["labels"] else None
        got = engine.side_by_side(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), labels).cpu().numpy()
        assert list(got.shape) == rec["shape"], rec["name"]
        assert sha(got) == rec["sha256"], rec["name"]
        assert engine.last_launches == 2


def test_side_by_side_against_oracle_odd_geometry(engine):
    rng = np.random.default_rng(21)
    shapes = [((97, 211), (1601, 333)), ((1600, 2), (801, 1000)), ((799, 1203), (65, 130)), ((2, 2), (4000, 30))]
    for (h1, w1), (h2, w2) in shapes:
        a = rng.integers(0, 256, (h1, w1, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (h2, w2, 3), dtype=np.uint8)
        got = engine.side_by_side(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), ("L", "R")).cpu().numpy()
        assert np.array_equal(got, OC.side_by_side(a, b, ("L", "R"))), (h1, w1, h2, w2)


def test_side_by_side_takes_strided_frames(engine):
    """Row-padded inputs (a crop of a wider tensor) are read through their pitch."""
    rng = np.random.default_rng(22)
    wide = rng.integers(0, 256, (480, 700, 3), dtype=np.uint8)
    dev = torch.from_numpy(wide).cuda()
    got = engine.side_by_side(dev[:, :640], dev[:, 60:]).cpu().numpy()
    assert np.array_equal(got, OC.side_by_side(wide[:, :640], wide[:, 60:]))


def test_status_stamp_goldens(engine, goldens):
    for rec in goldens["compare"]["stamp"]:
        got = engine.status_stamp(rec["verdict"], tuple(rec["size"])).cpu().numpy()
        assert list(got.shape) == rec["shape"] and sha(got) == rec["sha256"], rec
        assert np.array_equal(got, OC.status_stamp(rec["verdict"], tuple(rec["size"])))


def test_file_level_functions(engine, tmp_path):
    """The drop-in signatures: PNG files in, files out, same errors as the reference."""
    import cv2
    a, b = synth.noise_frame(31, 480, 640), synth.noise_frame(32, 333, 517)
    cv2.imwrite(str(tmp_path / "a.png"), a)
    cv2.imwrite(str(tmp_path / "b.png"), b)
    out = IU.create_side_by_side_comparison(tmp_path / "a.png", tmp_path / "b.png", tmp_path / "sub" / "cmp.png")
    assert out == tmp_path / "sub" / "cmp.png"
    assert np.array_equal(cv2.imread(str(out)), OC.side_by_side(a, b))
    with pytest.raises(ValueError, match="Failed to load images for comparison"):
        IU.create_side_by_side_comparison(tmp_path / "missing.png", tmp_path / "b.png", tmp_path / "x.png")
    out = IU.create_status_stamp("UNSAFE", tmp_path / "stamps" / "s.png")
    assert np.array_equal(cv2.imread(str(out), cv2.IMREAD_UNCHANGED), OC.status_stamp("UNSAFE"))


def test_side_by_side_batch_equals_single_calls(engine):
    """A batch of pairs in two launches (vis_compose_panels_batch + one label draw over every canvas): mixed geometries
    (bilinear, the exact-half area path, a copy-sized panel, strided inputs), custom labels; every canvas equals the
    single-pair call and the oracle."""
    rng = np.random.default_rng(23)
    shapes = [((480, 640), (480, 640)), ((1600, 400), (800, 300)), ((97, 211), (333, 517)), ((480, 640), (480, 640)),
              ((1080, 1920), (1080, 1920)), ((800, 120), (799, 1203))]
    pairs = []
    for (h1, w1), (h2, w2) in shapes:
        pairs.append((rng.integers(0, 256, (h1, w1, 3), dtype=np.uint8), rng.integers(0, 256, (h2, w2, 3), dtype=np.uint8)))
    wide = torch.from_numpy(rng.integers(0, 256, (480, 700, 3), dtype=np.uint8)).cuda()
    dev = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in pairs] + [(wide[:, :640], wide[:, 60:])]
    for labels in (None, ("Before", "After rework #2")):
        outs = engine.side_by_side_batch([a for a, _ in dev], [b for _, b in dev], labels)
        assert engine.last_launches == 2 and len(outs) == len(dev)
        for i, (a, b) in enumerate(dev):
            single = engine.side_by_side(a, b, labels)
            assert torch.equal(outs[i], single), (i, labels)
        for i in (0, 1, 4):
            want = OC.side_by_side(*pairs[i]) if labels is None else OC.side_by_side(*pairs[i], labels)
            assert np.array_equal(outs[i].cpu().numpy(), want), i
    with pytest.raises(ValueError):
        engine.side_by_side_batch([], [])
