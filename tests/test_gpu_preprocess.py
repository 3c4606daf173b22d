"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Bar: image_grid_thw, patch order and every pixel_values bit identical (integer resample + exact LUT), which is
stricter than BASELINE.json's 1e-5 post-normalisation tolerance; the tolerance is asserted as well.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden_frame
from oracle import qwen2vl as Q
from vision_inspection_system_b200 import geometry as G
from vision_inspection_system_b200 import synth

pytestmark = pytest.mark.gpu
POST_NORM_TOL = 1e-5          # BASELINE.json north_star: max-abs on fp32 pixel_values


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(engine, frames, **kw):
    dev = [torch.from_numpy(np.ascontiguousarray(f)).cuda() for f in frames]
    pv, grid = engine.preprocess(dev, **kw)
    torch.cuda.synchronize()
    return pv.cpu().numpy(), grid.numpy()


def check_equal(got, want, what):
    assert got.shape == want.shape, what
    assert float(np.abs(got - want).max()) <= POST_NORM_TOL, what
    assert np.array_equal(got, want), f"{what}: within tolerance but not bit-exact"


@pytest.mark.parametrize("force_generic", [False, True])
def test_goldens(engine, goldens, arrays, force_generic):
    for rec in goldens["qwen"]:
        frame = golden_frame(rec, arrays)
        kw = {} if rec["max_pixels"] is None else {"max_pixels": rec["max_pixels"]}
        pv, grid = run(engine, [frame], force_generic=force_generic, **kw)
        assert grid[0].tolist() == rec["grid_thw"], rec["name"]
        assert np.array_equal(pv[rec["sample_rows"]], arrays[f"qwen_{rec['name']}_samples"]), rec["name"]
        assert sha(pv) == rec["sha256"], rec["name"]


def test_fused_path_is_taken_for_the_benchmark_geometry(engine):
    f = torch.from_numpy(synth.noise_frame(1234, 1080, 1920)).cuda()
    from vision_inspection_system_b200.engine import _SchedLaunch
    plan = engine.plan_batch([f])
    assert len(plan.fused) == 1 and not plan.generic and isinstance(plan.fused[0], _SchedLaunch)
    engine.preprocess([f])
    assert engine.last_launches == 1
    plan = engine.plan_batch([f], path="general")
    assert len(plan.fused) == 1 and not isinstance(plan.fused[0], _SchedLaunch)


@pytest.mark.parametrize("shape,max_pixels", [
    ((1080, 1920), G.DEFAULT_MAX_PIXELS), ((1080, 1920), G.HUB_MAX_PIXELS), ((2160, 3840), G.DEFAULT_MAX_PIXELS),
    ((720, 1280), G.DEFAULT_MAX_PIXELS), ((480, 640), G.DEFAULT_MAX_PIXELS), ((1536, 2048), G.DEFAULT_MAX_PIXELS),
    ((1152, 2048), G.DEFAULT_MAX_PIXELS), ((576, 1024), G.DEFAULT_MAX_PIXELS), ((100, 502), G.DEFAULT_MAX_PIXELS),
    ((64, 96), G.DEFAULT_MAX_PIXELS), ((28, 5600), G.DEFAULT_MAX_PIXELS), ((3000, 20), G.DEFAULT_MAX_PIXELS),
    ((2048, 1536), G.DEFAULT_MAX_PIXELS), ((1080, 1920), 400000), ((2160, 3840), 250000), ((3100, 5500), G.DEFAULT_MAX_PIXELS), ((1200, 1600), G.DEFAULT_MAX_PIXELS), ((600, 5000), G.DEFAULT_MAX_PIXELS),
    ((1, 1), G.DEFAULT_MAX_PIXELS), ((2160, 3840), G.HUB_MAX_PIXELS)])
def test_against_oracle(engine, shape, max_pixels):
    frame = synth.noise_frame(hash(shape) % 1000, *shape)
    want, wgrid = Q.preprocess([frame], max_pixels=max_pixels)
    for path in ("auto", "general"):             # scheduled kernel where it applies / general fused kernel everywhere
        for vsplit in (None, 1, 3):
            pv, grid = run(engine, [frame], max_pixels=max_pixels, vsplit=vsplit, path=path)
            assert np.array_equal(grid, wgrid)
            check_equal(pv, want, (shape, max_pixels, vsplit, path))


def test_padded_pitch_and_unaligned_views(engine):
    base = synth.noise_frame(77, 300, 640)
    want, _ = Q.preprocess([base])
    padded = torch.zeros((300, 656, 3), dtype=torch.uint8, device="cuda")          # pitch 1968 = 16 * 123
    padded[:, :640] = torch.from_numpy(base).cuda()
    pv, _ = engine.preprocess([padded[:, :640]])
    check_equal(pv.cpu().numpy(), want, "padded pitch (fused)")
    odd = torch.zeros((300, 641, 3), dtype=torch.uint8, device="cuda")             # pitch 1923: not bulk-copyable
    odd[:, :640] = torch.from_numpy(base).cuda()
    plan = engine.plan_batch([odd[:, :640]])
    assert plan.generic and not plan.fused                                          # as submitted: generic passes
    pv, _ = engine.preprocess([odd[:, :640]], force_generic=True)
    check_equal(pv.cpu().numpy(), want, "unaligned pitch (generic)")
    pv, _ = engine.preprocess([odd[:, :640]])                                       # default: repacked, fused kernel
    assert engine.last_launches == 2     # vis_repitch_u8 + the fused launch
    check_equal(pv.cpu().numpy(), want, "unaligned pitch (repacked)")
    mouri_like = [synth.noise_frame(80 + i, 100, 502) for i in range(3)]            # 1506-byte rows (BASELINE config 1 shape)
    wantm, _ = Q.preprocess(mouri_like)
    pv, _ = engine.preprocess([torch.from_numpy(f).cuda() for f in mouri_like])
    assert engine.last_launches == 2     # vis_repitch_u8 + the fused launch
    check_equal(pv.cpu().numpy(), wantm, "502-pixel rows (repacked)")
    pv, _ = engine.preprocess(torch.from_numpy(np.stack(mouri_like)).cuda())        # the same as one [B, H, W, 3] tensor
    assert engine.last_launches == 2     # vis_repitch_u8 + the fused launch
    check_equal(pv.cpu().numpy(), wantm, "502-pixel rows, batch tensor (repacked)")


def test_mixed_resolution_batch(engine):
    shapes = synth.mixed_resolution_shapes(10, seed=9000) + [(3000, 20), (56, 56)]
    frames = [synth.noise_frame(500 + i, *s) for i, s in enumerate(shapes)]
    want, wgrid = Q.preprocess(frames)
    pv, grid = run(engine, frames)
    assert np.array_equal(grid, wgrid)
    check_equal(pv, want, "mixed batch")


def test_uniform_batch_tensor_and_out_buffer(engine):
    frames = synth.frames_1080p(5)
    want, wgrid = Q.preprocess(list(frames))
    batch = torch.from_numpy(frames).cuda()
    out = torch.empty((want.shape[0], 1176), dtype=torch.float32, device="cuda")
    pv, grid = engine.preprocess(batch, out=out)
    assert pv.data_ptr() == out.data_ptr()
    assert np.array_equal(grid.numpy(), wgrid)
    check_equal(pv.cpu().numpy(), want, "uniform batch")
    pv2, _ = engine.preprocess(batch)                      # cached plan, fresh output
    assert torch.equal(pv, pv2)
    with pytest.raises(ValueError):
        engine.preprocess(batch, out=torch.empty((3, 1176), dtype=torch.float32, device="cuda"))


def test_full_size_batch_properties(engine):
    """BASELINE config 2 at full size (256 x 1080p): size-independent properties instead of a full oracle run."""
    n = 256
    base = synth.frames_1080p(8)
    batch = torch.from_numpy(base).cuda().repeat(n // 8, 1, 1, 1)          # frame i == frame i % 8
    pv, grid = engine.preprocess(batch)
    torch.cuda.synchronize()
    rows = 52 * 94
    assert pv.shape == (n * rows, 1176) and grid.shape == (n, 3) and (grid == torch.tensor([1, 52, 94])).all()
    ref8, _ = Q.preprocess(list(base))
    got = pv.view(n, rows, 1176)
    assert torch.equal(got[:8].cpu(), torch.from_numpy(ref8).view(8, rows, 1176))           # oracle on the first 8
    assert torch.equal(got, got[:8].repeat(n // 8, 1, 1))                                    # replicas identical
    v = pv.view(-1, 3, 2, 14, 14)
    assert torch.equal(v[:, :, 0], v[:, :, 1])                                               # temporal duplicate
    lut = torch.from_numpy(Q.normalize_lut()).cuda().view(256, 3)
    for c in range(3):                                                                       # only table values occur
        assert torch.isin(v[::97, c].reshape(-1), lut[:, c]).all()


def test_full_size_4k_batch_properties(engine):
    """BASELINE config 3 (4K frames at the Qwen2-VL hub max_pixels, one GPU's shard): properties + oracle on two frames."""
    n = 16
    base = synth.frames_4k(2)
    batch = torch.from_numpy(base).cuda().repeat(n // 2, 1, 1, 1)
    pv, grid = engine.preprocess(batch, max_pixels=G.HUB_MAX_PIXELS)
    torch.cuda.synchronize()
    rows = (2156 // 14) * (3836 // 14)
    assert pv.shape == (n * rows, 1176) and (grid == torch.tensor([1, 154, 274])).all()
    got = pv.view(n, rows, 1176)
    assert torch.equal(got, got[:2].repeat(n // 2, 1, 1))                                    # replicas identical
    want, _ = Q.preprocess([base[0]], max_pixels=G.HUB_MAX_PIXELS)
    assert torch.equal(got[0].cpu(), torch.from_numpy(want))
    v = pv.view(-1, 3, 2, 14, 14)
    assert torch.equal(v[:, :, 0], v[:, :, 1])                                               # temporal duplicate
    # default max_pixels on the same frames: the 16-slot kernel (13 taps)
    pv2, grid2 = engine.preprocess(batch)
    want2, wgrid2 = Q.preprocess([base[1]])
    assert (grid2 == torch.tensor(wgrid2[0])).all()
    r2 = want2.shape[0]
    assert torch.equal(pv2.view(n, r2, 1176)[1].cpu(), torch.from_numpy(want2))
    assert torch.equal(pv2.view(n, r2, 1176), pv2.view(n, r2, 1176)[:2].repeat(n // 2, 1, 1))


def test_constant_and_extreme_frames(engine):
    lut = Q.normalize_lut().reshape(256, 3)
    for value in (0, 255, 128):
        f = np.full((480, 640, 3), value, np.uint8)
        pv, _ = run(engine, [f])
        for c in range(3):
            assert np.all(pv.reshape(-1, 3, 392)[:, c] == lut[value, c])


def test_error_behaviour(engine):
    with pytest.raises(TypeError):
        engine.preprocess([torch.zeros((10, 10, 3), dtype=torch.uint8)])                    # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        engine.preprocess([torch.zeros((10, 2001, 3), dtype=torch.uint8, device="cuda")])   # aspect ratio > 200
    with pytest.raises(ValueError):
        engine.preprocess([])


def test_config1_jpeg_file_known_answer(engine):
    """BASELINE config 1 as written: the reference's own Mouri.jpg FILE -> load_image -> preprocess_for_vlm must give the
    known answer of SURVEY.md 8(c) (grid [1, 8, 36], 288 rows, sha256[:16] 01e0d93015585789, decoded RGB 13b0ebd773d68238)."""
    from pathlib import Path
    from vision_inspection_system_b200 import image_utils as IU
    path = Path(__file__).parent / "golden" / "Mouri.jpg"
    img = IU.load_image(path)
    assert img.size == (502, 100) and sha(np.asarray(img.convert("RGB")))[:16] == "13b0ebd773d68238"
    for source in (img, path, str(path)):
        pv, grid = IU.preprocess_for_vlm(source)
        assert grid.tolist() == [[1, 8, 36]] and tuple(pv.shape) == (288, 1176)
        assert sha(pv.cpu().numpy())[:16] == "01e0d93015585789"
    pv, grid = IU.preprocess_for_vlm([path, img], role="inspector")          # 502 px: no thumbnail for either role
    assert grid.tolist() == [[1, 8, 36]] * 2 and sha(pv[:288].cpu().numpy())[:16] == "01e0d93015585789"
    assert torch.equal(pv[:288], pv[288:])
    with pytest.raises(FileNotFoundError):
        IU.load_image(path.with_name("missing.jpg"))


def test_image_utils_preprocess_for_vlm(engine, arrays, goldens):
    from PIL import Image
    from vision_inspection_system_b200 import image_utils as IU
    rec = next(r for r in goldens["qwen"] if r["name"] == "mouri")
    pv, grid = IU.preprocess_for_vlm(Image.fromarray(arrays["mouri_rgb"]))
    assert grid.tolist() == [rec["grid_thw"]] and sha(pv.cpu().numpy()) == rec["sha256"]
    pv2, _ = IU.normalize([arrays["mouri_rgb"]])
    assert torch.equal(pv, pv2)
    # dual Inspector / Auditor inputs (BASELINE config 5): agent thumbnail, then the Qwen2-VL processor
    frame = synth.noise_frame(31, 2160, 3840)
    for role, limit in (("inspector", 2048), ("auditor", 1024)):
        want, wgrid = Q.preprocess([Q.agent_thumbnail(frame, limit)])
        got, ggrid = IU.preprocess_for_vlm(frame, role=role)
        assert np.array_equal(ggrid.numpy(), wgrid)
        check_equal(got.cpu().numpy(), want, role)


def test_preprocess_dual_equals_two_role_passes(engine):
    """BASELINE config 5 entry point: both agents' inputs in one pass == thumbnail(2048 / 1024, LANCZOS) -> processor per
    role (src/agents/vlm_inspector.py:59-69, vlm_auditor.py:87-96), bit-exact; frames neither role thumbnails are
    resampled once and stored to both tensors; frames of one geometry share a launch across roles and sources."""
    shapes = [(480, 640), (720, 1280), (1080, 1920), (1536, 2048), (2160, 3840), (100, 502), (1080, 1920), (480, 640),
              (2160, 3840), (100, 502)]
    frames = [synth.noise_frame(9100 + i, h, w) for i, (h, w) in enumerate(shapes)]
    dev = [torch.from_numpy(f).cuda() for f in frames]
    res = engine.preprocess_dual(dev)
    launches = engine.last_launches
    torch.cuda.synchronize()
    for role, limit in (("inspector", 2048), ("auditor", 1024)):
        want, wgrid = Q.preprocess([Q.agent_thumbnail(f, limit) for f in frames])
        pv, grid = res[role]
        assert np.array_equal(grid.numpy(), wgrid), role
        check_equal(pv.cpu().numpy(), want, role)
    assert res["inspector"][0].data_ptr() + res["inspector"][0].numel() * 4 == res["auditor"][0].data_ptr()   # one allocation
    # 5 thumbnail launches (720p/1080p/1536 Auditor, 4K both) + 1 re-pitch (502-px rows) + one processor launch per
    # distinct input geometry (8: the Auditor's 1024x576 thumbnails of 720p, 1080p and 4K frames share one)
    assert launches <= 14, launches
    again = engine.preprocess_dual(dev, out=torch.empty_like(torch.cat([res["inspector"][0], res["auditor"][0]])))
    assert torch.equal(again["auditor"][0], res["auditor"][0]) and torch.equal(again["inspector"][0], res["inspector"][0])
    # the cached pass spreads its independent launches over several CUDA streams (event dependencies between a thumbnail
    # and the processor launch that reads it): one stream after the other and any stream count give the same tensors,
    # also when the result buffer was just overwritten on the caller's stream
    assert engine.dual_streams > 1
    keep = engine.dual_streams
    try:
        for n_streams in (1, 2, 5):
            engine.dual_streams = n_streams
            scratch = torch.full_like(torch.cat([res["inspector"][0], res["auditor"][0]]), float("nan"))
            got = engine.preprocess_dual(dev, out=scratch)
            torch.cuda.synchronize()
            assert torch.equal(got["auditor"][0], res["auditor"][0]) and torch.equal(got["inspector"][0], res["inspector"][0]), n_streams
    finally:
        engine.dual_streams = keep
    # a frame >= 4x the Auditor's limit takes Pillow's reduce pre-pass (not cacheable as one fused launch): still exact,
    # next to an unaligned 502-pixel frame in the same call; twice, because the second call reuses the cached plan
    big = [synth.noise_frame(9200, 2200, 4400), synth.noise_frame(9201, 100, 502), synth.noise_frame(9202, 1080, 1920)]
    dev_big = [torch.from_numpy(f).cuda() for f in big]
    for _ in range(2):
        res = engine.preprocess_dual(dev_big)
        for role, limit in (("inspector", 2048), ("auditor", 1024)):
            want, wgrid = Q.preprocess([Q.agent_thumbnail(f, limit) for f in big])
            assert np.array_equal(res[role][1].numpy(), wgrid), role
            check_equal(res[role][0].cpu().numpy(), want, ("reduce pre-pass", role))
