"""-m gpu: vis_quality_stats (through the C ABI) against the oracle's exact sums and the reference's own results.

Bar: the three int64 sums are bit-exact; the float64 variance finished on the host is within 1e-9 relative of what
the reference computed (numpy's two-pass variance of the same integers), every score and flag equal."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import quality as Q
from test_oracle_quality import quality_frames, same_result
from vision_inspection_system_b200 import synth

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))


def test_sums_are_exact(engine):
    rng = np.random.default_rng(3)
    # whole strips through the bulk-copy ring (16-byte aligned rows), rows that are only 4-byte aligned, byte-aligned rows,
    # partial last strips with >= 6 / < 6 pixels (the right halo of the last ring strip), one-row / one-column frames
    shapes = [(1080, 1920), (480, 640), (333, 517), (1, 1), (1, 9), (9, 1), (2, 2), (17, 129), (16, 128), (2160, 3840),
              (1, 128), (3, 256), (36, 256), (37, 256), (75, 2048), (100, 1152), (100, 1030), (40, 1157), (40, 1158), (181, 133),
              (200, 132), (500, 3), (64, 127), (19, 40)]
    frames = [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in shapes]
    sums, got_shapes = engine.quality_stats([torch.from_numpy(f).cuda() for f in frames])
    assert got_shapes == shapes and engine.last_launches == 2
    host = sums.cpu().numpy()
    for i, f in enumerate(frames):
        assert tuple(int(v) for v in host[i]) == Q.stats(f), shapes[i]


def test_goldens_from_the_reference(engine, goldens, arrays):
    from vision_inspection_system_b200 import image_quality as IQ
    frames = quality_frames(arrays)
    names = [r["name"] for r in goldens["quality"]]
    results = IQ.assess_frames([torch.from_numpy(np.ascontiguousarray(frames[n])).cuda() for n in names])
    for rec, got in zip(goldens["quality"], results):
        same_result(got, rec["result"], 1e-9)


def test_uniform_batch_and_padded_pitch(engine):
    base = synth.frames_1080p(4)
    sums, _ = engine.quality_stats(torch.from_numpy(base).cuda())
    for i in range(4):
        assert tuple(int(v) for v in sums[i].cpu().numpy()) == Q.stats(base[i])
    padded = torch.zeros((300, 700, 3), dtype=torch.uint8, device="cuda")
    padded[:, :640] = torch.from_numpy(synth.noise_frame(5, 300, 640)).cuda()
    s2, _ = engine.quality_stats([padded[:, :640]])
    assert tuple(int(v) for v in s2[0].cpu().numpy()) == Q.stats(synth.noise_frame(5, 300, 640))


def test_assess_image_quality_file_semantics(engine, tmp_path):
    import cv2
    from vision_inspection_system_b200 import image_quality as IQ
    frame = synth.noise_frame(7, 480, 640)
    p = tmp_path / "frame.png"
    cv2.imwrite(str(p), frame)
    res = IQ.assess_image_quality(p)
    assert res["image_path"] == str(p)
    res.pop("image_path")
    same_result(res, Q.assess(frame), 1e-9)
    bad = IQ.assess_image_quality(tmp_path / "missing.png")          # never raises: failed-result dict
    assert bad["quality_passed"] is False and bad["quality_score"] == 0.0 and "Failed to load image" in bad["error"]


def test_ring_strips_next_to_partial_strips(engine):
    """Rows that ARE 16-byte aligned (the bulk-copy ring) with a partial last strip: >= 6 pixels left (the last ring strip
    takes its right halo from the row), < 6 left (that strip falls back to per-lane loads), and row-padded views whose
    pitch, not width, carries the alignment."""
    rng = np.random.default_rng(5)
    cases = [(90, 1168, 1168), (75, 1028, 1040), (300, 640, 704), (41, 134, 144), (200, 2560, 2560), (33, 1024 + 6, 1040)]
    frames, views = [], []
    for h, w, wp in cases:
        assert (wp * 3) % 16 == 0
        buf = rng.integers(0, 256, (h, wp, 3), dtype=np.uint8)
        frames.append(np.ascontiguousarray(buf[:, :w]))
        views.append(torch.from_numpy(buf).cuda()[:, :w])
    for v, f, c in zip(views, frames, cases):                      # one by one (tall bands) ...
        sums, _ = engine.quality_stats([v])
        assert tuple(int(x) for x in sums[0].cpu().numpy()) == Q.stats(f), c
    sums, _ = engine.quality_stats(views)                           # ... and as one mixed batch (short bands)
    host = sums.cpu().numpy()
    for i, f in enumerate(frames):
        assert tuple(int(x) for x in host[i]) == Q.stats(f), cases[i]
