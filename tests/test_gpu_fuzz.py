"""-m gpu: randomised geometries through every resampling path (scheduled 8-slot / 16-slot kernels with their tap
classes, vertical-warp tiers and row segments, general kernels, generic passes, reduce pre-pass) against the oracle.
Seeded, so a failure names its geometry."""
import numpy as np
import pytest
import torch

from oracle import qwen2vl as Q
from vision_inspection_system_b200 import geometry as G
from vision_inspection_system_b200 import synth

pytestmark = pytest.mark.gpu


def random_shapes(seed, n, lo=(120, 160), hi=(2400, 4200)):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        h = int(rng.integers(lo[0], hi[0]))
        w = int(rng.integers(lo[1], hi[1]))
        if rng.random() < 0.5:
            w = (w + 15) // 16 * 16                      # rows of whole 16-byte vectors take the fused kernels
        out.append((h, w))
    return out


def test_preprocess_random_geometries(engine):
    shapes = random_shapes(101, 14)
    frames = [synth.noise_frame(500 + i, h, w) for i, (h, w) in enumerate(shapes)]
    for max_pixels in (G.DEFAULT_MAX_PIXELS, G.HUB_MAX_PIXELS):
        pv, grid = engine.preprocess([torch.from_numpy(f).cuda() for f in frames], max_pixels=max_pixels)
        want, wgrid = Q.preprocess(frames, max_pixels=max_pixels)
        assert np.array_equal(grid.numpy(), wgrid), shapes
        got = pv.cpu().numpy()
        if not np.array_equal(got, want):
            rows = np.cumsum([0] + [int(g[1] * g[2]) for g in wgrid])
            bad = [shapes[i] for i in range(len(shapes)) if not np.array_equal(got[rows[i]:rows[i + 1]], want[rows[i]:rows[i + 1]])]
            raise AssertionError(f"pixel_values differ for {bad} at max_pixels={max_pixels}")


def test_same_geometry_batches_random(engine):
    """Uniform batches (one scheduled launch, several row segments) for a few random geometries."""
    for k, (h, w) in enumerate(random_shapes(102, 4, lo=(600, 800))):
        w = (w + 15) // 16 * 16
        frames = np.stack([synth.noise_frame(700 + 10 * k + i, h, w) for i in range(5)])
        pv, grid = engine.preprocess(torch.from_numpy(frames).cuda())
        want, wgrid = Q.preprocess(list(frames))
        assert np.array_equal(grid.numpy(), wgrid) and np.array_equal(pv.cpu().numpy(), want), (h, w)


@pytest.mark.parametrize("role,limit", [("inspector", 2048), ("auditor", 1024)])
def test_agent_thumbnails_random_geometries(engine, role, limit):
    shapes = random_shapes(103 + limit, 10, lo=(700, 1100), hi=(3200, 6400))
    frames = [synth.noise_frame(900 + i, h, w) for i, (h, w) in enumerate(shapes)]
    outs = engine.agent_inputs([torch.from_numpy(f).cuda() for f in frames], role)
    for f, o, s in zip(frames, outs, shapes):
        assert np.array_equal(o.cpu().numpy(), Q.agent_thumbnail(f, limit)), (role, s)


def test_overlay_random_frames_and_boxes(engine):
    """Sprites, stamps, translated templates and in-place expansion all mixed: random frame sizes (markers and dashes
    touching the borders on the small ones), random boxes pushed to the edges, free-text labels; batch and in place."""
    from oracle import overlay as OV
    rng = np.random.default_rng(105)
    shapes = [(97, 211), (130, 150), (300, 500), (480, 640), (333, 517), (720, 1280), (1080, 1920), (65, 400), (400, 66)]
    items = []
    for k in range(18):
        h, w = shapes[k % len(shapes)]
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        boxes = synth.random_boxes(rng, int(rng.integers(1, 7)))
        for b in boxes[:3]:
            if rng.random() < 0.5:
                b["x"] = 0.0
            if rng.random() < 0.5:
                b["y"] = 0.0
            if rng.random() < 0.3:
                b["x"] = round(100 - b["width"], 1)
            if rng.random() < 0.3:
                b["y"] = round(100 - b["height"], 1)
            if rng.random() < 0.3:
                b["label"] = ["wide-label!", "#123", "Q", "#7"][int(rng.integers(0, 4))]
            if rng.random() < 0.5:
                b["confidence"] = "low"
        items.append((frame, boxes))
    outs = engine.annotate([torch.from_numpy(f).cuda() for f, _ in items], [b for _, b in items])
    for k, ((f, b), o) in enumerate(zip(items, outs)):
        assert np.array_equal(o.cpu().numpy(), OV.draw_bounding_boxes(f, b)), (k, f.shape)
    devs = [torch.from_numpy(f.copy()).cuda() for f, _ in items]
    engine.annotate(devs, [b for _, b in items], inplace=True)
    for k, ((f, b), d) in enumerate(zip(items, devs)):
        assert np.array_equal(d.cpu().numpy(), OV.draw_bounding_boxes(f, b)), (k, f.shape, "in place")


def test_all_kernel_families_for_long_windows(engine):
    """9+ tap geometries are served by the integer tensor-path (IMMA) kernel by default, by the packed-byte (IDP.4A) kernel
    and by the 16-slot IMAD kernel on request: all bit-exact against the oracle — 4K at the default max_pixels (13 taps), bicubic 2x downscales, the
    agents' LANCZOS thumbnails (13 and 25 taps), odd segment counts."""
    frames4k = [synth.noise_frame(4000 + i, 2160, 3840) for i in range(3)]
    want4k, wgrid = Q.preprocess(frames4k)
    mid = [synth.noise_frame(4100 + i, 1536, 2048) for i in range(5)]
    wantmid, _ = Q.preprocess(mid)
    thumbs = [(2160, 3840, 2048), (2160, 3840, 1024), (1080, 1920, 1024), (1600, 2560, 1024), (1234, 3008, 1024)]
    tframes = [synth.noise_frame(4200 + i, h, w) for i, (h, w, _) in enumerate(thumbs)]
    try:
        for dp in ((True, True), (True, False), (False, False)):
            engine.use_dp4a(*dp)
            pv, grid = engine.preprocess(torch.from_numpy(np.stack(frames4k)).cuda())
            assert np.array_equal(grid.numpy(), wgrid) and np.array_equal(pv.cpu().numpy(), want4k), ("4k", dp)
            plan = engine.plan_batch(torch.from_numpy(np.stack(frames4k)).cuda())
            assert not plan.generic and len(plan.fused) == 1
            pv, _ = engine.preprocess(torch.from_numpy(np.stack(mid)).cuda(), vsplit=3)
            assert np.array_equal(pv.cpu().numpy(), wantmid), ("1536x2048", dp)
            for f, (h, w, limit) in zip(tframes, thumbs):
                got = engine.agent_inputs([torch.from_numpy(f).cuda()], max_size=limit)[0]
                assert engine.last_launches == 1, (h, w, limit, dp)
                assert np.array_equal(got.cpu().numpy(), Q.agent_thumbnail(f, limit)), (h, w, limit, dp)
    finally:
        engine.use_dp4a(True, True)


def test_tensor_path_kernel_random_scales(engine):
    """The integer tensor-path kernel (k_fused_mma) over its whole envelope: vertical scales 1.5 .. 5.5 (chunk advance 24 /
    28 / 32 rows), 9 .. 33 taps (4 .. 9 window words, 1 .. 3 k-steps), strip widths that are no multiple of 16, batches that
    take several row segments — LANCZOS and BICUBIC uint8 resizes against the oracle (the pixel_values form of such
    geometries is in test_preprocess_random_geometries and tests/test_gpu_preprocess.py)."""
    from vision_inspection_system_b200 import _native as N
    rng = np.random.default_rng(777)
    taken = 0
    for k in range(64):
        filt = Q.LANCZOS if k % 2 else Q.BICUBIC
        sx = float(rng.uniform(1.5, 5.4 if filt == Q.LANCZOS else 7.5))
        sy = float(rng.uniform(1.5, 5.4 if filt == Q.LANCZOS else 7.5))
        out_w = int(rng.integers(16, 160)) * 4
        out_h = int(rng.integers(30, 400))
        w = (int(out_w * sx) + 15) // 16 * 16
        h = int(out_h * sy)
        if h * w > 3000 * 6000:
            continue
        n = int(rng.integers(1, 4))
        frames = [synth.noise_frame(8000 + 10 * k + i, h, w) for i in range(n)]
        dev = [torch.from_numpy(f).cuda() for f in frames]
        outs = engine.resize_batch_u8(dev, out_h, out_w, filt)
        hit = engine._resize_sched(h, w, out_h, out_w, filt, w * 3, 1)
        if hit is not None:
            head = np.frombuffer(hit[0][:N.SCHED_HEAD_DTYPE.itemsize], N.SCHED_HEAD_DTYPE)[0]
            taken += int(head["mma_ks"]) > 0
        for f, o in zip(frames, outs):
            assert np.array_equal(o.cpu().numpy(), Q.resize(f, out_h, out_w, filt)), (h, w, out_h, out_w, filt)
    assert taken >= 36, taken        # most of these geometries are the tensor-path kernel's
