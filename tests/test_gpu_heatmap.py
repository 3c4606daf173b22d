"""-m gpu: vis_heatmap_batch (through the C ABI) against the arrays captured from the reference's own
create_heatmap_overlay and against the oracle.  Floating-point path: tolerance stated in test_oracle_heatmap.py
(<= 2 levels, >= 99.5 % of the bytes identical)."""
import numpy as np
import pytest
import torch

from oracle import heatmap as OH
from test_oracle_heatmap import close_enough
from vision_inspection_system_b200 import heatmap as H
from vision_inspection_system_b200 import synth

pytestmark = pytest.mark.gpu


def test_goldens_from_the_reference(engine, goldens, arrays):
    cases = {name: (frame, defects, step) for name, frame, defects, step in synth.heatmap_cases()}
    for rec in goldens["heatmap"]:
        frame, defects, step = cases[rec["name"]]
        dev = torch.from_numpy(np.ascontiguousarray(frame)).cuda()
        got = engine.heatmap(dev, defects).cpu().numpy()
        close_enough(got[::step, ::step], arrays[f"heatmap_{rec['name']}"], rec["name"])
        assert torch.equal(dev.cpu(), torch.from_numpy(np.ascontiguousarray(frame)))        # input untouched
    empty = cases["hgrad_360x480_empty"][0]
    out = engine.heatmap(torch.from_numpy(np.ascontiguousarray(empty)).cuda(), [])
    assert np.array_equal(out.cpu().numpy(), empty) and engine.last_launches == 0            # no defects: plain copy


def test_against_oracle_random_defects(engine):
    for seed, shape in ((1, (333, 517)), (2, (720, 1280)), (3, (97, 211))):
        rng = np.random.default_rng(seed)
        frame = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        defects = synth.random_defects(rng, int(rng.integers(1, 7)))
        got = engine.heatmap(torch.from_numpy(frame).cuda(), defects).cpu().numpy()
        close_enough(got, OH.create_heatmap_overlay(frame, defects, H.JET_BGR), (seed, shape))


def test_create_heatmap_overlay_files(engine, tmp_path):
    import cv2
    from vision_inspection_system_b200 import image_utils as IU
    frame = synth.pattern_frames(360, 480)["hgrad"]
    defects = synth.random_defects(np.random.default_rng(8000), 3)
    src, dst = tmp_path / "in.png", tmp_path / "heat.png"
    cv2.imwrite(str(src), frame)
    assert IU.create_heatmap_overlay(src, defects, dst) == dst
    close_enough(cv2.imread(str(dst)), OH.create_heatmap_overlay(frame, defects, H.JET_BGR), "file round trip")
    with pytest.raises(ValueError, match="Failed to load image"):
        IU.create_heatmap_overlay(tmp_path / "missing.png", defects, dst)


def test_batch_equals_per_frame_and_oracle(engine):
    """heatmap_batch: mixed sizes, a frame without defects (plain copy), a widespread defect, many defects per frame —
    six launches for the whole batch, every frame within the tolerance of the oracle and IDENTICAL to its own per-frame call."""
    shapes = [(333, 517), (720, 1280), (97, 211), (1080, 1920), (480, 640), (600, 50)]
    frames, defects = [], []
    for i, shape in enumerate(shapes):
        rng = np.random.default_rng(40 + i)
        frames.append(rng.integers(0, 256, shape + (3,), dtype=np.uint8))
        defects.append(synth.random_defects(rng, int(rng.integers(1, 7))))
    defects[2] = []                                                                   # untouched copy
    defects[4] = defects[4] + [{"bbox": None, "location": "corrosion on the entire surface", "safety_impact": "MODERATE",
                                "confidence": "high"}]
    dev = [torch.from_numpy(f).cuda() for f in frames]
    outs = engine.heatmap_batch(dev, defects)
    assert engine.last_launches == 6
    for f, d, o, dv in zip(frames, defects, outs, dev):
        got = o.cpu().numpy()
        close_enough(got, OH.create_heatmap_overlay(f, d, H.JET_BGR), f.shape)
        assert np.array_equal(engine.heatmap(dv, d).cpu().numpy(), got), f.shape          # batching changes nothing
    assert np.array_equal(outs[2].cpu().numpy(), frames[2])
    same = torch.from_numpy(np.stack([frames[3]] * 3)).cuda()                              # a [B, H, W, 3] tensor
    outs3 = engine.heatmap_batch(same, [defects[3]] * 3)
    assert all(torch.equal(o, outs[3]) for o in outs3)
