"""-m gpu: uint8 resampling (resize_image, agent thumbnails) against the oracle and the golden vectors."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import qwen2vl as Q
from vision_inspection_system_b200 import image_utils as IU
from vision_inspection_system_b200 import synth

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_thumbnail_goldens(engine, goldens, arrays):
    from PIL import Image
    from vision_inspection_system_b200 import image_utils as IU
    for rec in goldens["thumbnail"]:
        h, w = rec["shape"]
        frame = synth.noise_frame(rec["seed"], h, w)
        thumb = IU.agent_thumbnail(Image.fromarray(frame), max_size=rec["limit"])
        assert list(thumb.size) == rec["thumb_size"] and sha(np.asarray(thumb)) == rec["thumb_sha256"], rec["name"]
        rz = IU.resize_image(Image.fromarray(frame), rec["limit"])
        assert list(rz.size) == rec["resize_image_size"] and sha(np.asarray(rz)) == rec["resize_image_sha256"], rec["name"]


@pytest.mark.parametrize("shape,out,ch", [((333, 517), (200, 700), 3), ((3000, 20), (2800, 28), 3), ((97, 131), (40, 60), 1),
                                          ((64, 64), (64, 64), 3), ((64, 80), (64, 40), 3), ((80, 64), (40, 64), 4),
                                          ((10, 10), (56, 56), 2), ((2160, 3840), (576, 1024), 3)])
@pytest.mark.parametrize("filt", [Q.BICUBIC, Q.LANCZOS])
def test_resize_against_oracle(engine, shape, out, ch, filt):
    a = np.random.default_rng(11).integers(0, 256, shape + (ch,), dtype=np.uint8)
    got = engine.resize_u8(torch.from_numpy(a).cuda(), out[0], out[1], filt).cpu().numpy()
    assert np.array_equal(got, Q.resize(a, out[0], out[1], filt))


def test_resize_image_semantics(engine):
    from PIL import Image
    from vision_inspection_system_b200 import image_utils as IU
    small = Image.fromarray(synth.noise_frame(1, 100, 200))
    assert IU.resize_image(small) is small                                  # unchanged object when it fits
    rgba = Image.fromarray(np.random.default_rng(2).integers(0, 256, (300, 500, 4), dtype=np.uint8), "RGBA")
    want = rgba.resize((256, 153), Image.Resampling.LANCZOS)               # Pillow: premultiplied-alpha path
    got = IU.resize_image(rgba, 256)
    assert got.mode == "RGBA" and got.size == (256, 153)
    assert np.array_equal(np.asarray(got), np.asarray(want))
    gray = Image.fromarray(synth.noise_frame(3, 300, 500)[:, :, 0])
    assert np.array_equal(np.asarray(IU.resize_image(gray, 128)), np.asarray(gray.resize((128, 76), Image.Resampling.LANCZOS)))


def test_fused_thumbnail_path_and_batches(engine):
    """RGB LANCZOS thumbnails with <= 16 taps take the fused scheduled kernel (one launch for a whole batch); wider
    filters, other channel counts and odd widths fall back to the generic passes.  Both give Pillow's bytes."""
    frames = [synth.noise_frame(30 + i, 1080, 1920) for i in range(3)]
    dev = [torch.from_numpy(f).cuda() for f in frames]
    outs = engine.resize_batch_u8(dev, 576, 1024, Q.LANCZOS)
    assert engine.last_launches == 1
    for f, o in zip(frames, outs):
        assert np.array_equal(o.cpu().numpy(), Q.resize(f, 576, 1024, Q.LANCZOS))
    big = synth.noise_frame(33, 2160, 3840)
    got = engine.resize_u8(torch.from_numpy(big).cuda(), 1152, 2048, Q.LANCZOS)
    assert engine.last_launches == 1
    assert np.array_equal(got.cpu().numpy(), Q.resize(big, 1152, 2048, Q.LANCZOS))
    got = engine.resize_u8(torch.from_numpy(big).cuda(), 576, 1024, Q.LANCZOS)          # 25 taps: pull-order H role
    assert engine.last_launches == 1
    assert np.array_equal(got.cpu().numpy(), Q.resize(big, 576, 1024, Q.LANCZOS))
    for out_hw in ((864, 1536), (432, 768), (270, 480)):                                 # 17, 31 taps: fused; 49 taps: generic
        got = engine.resize_u8(torch.from_numpy(big).cuda(), out_hw[0], out_hw[1], Q.LANCZOS)
        assert engine.last_launches == (2 if out_hw[0] == 270 else 1), out_hw
        assert np.array_equal(got.cpu().numpy(), Q.resize(big, out_hw[0], out_hw[1], Q.LANCZOS)), out_hw
    for out_hw in ((683, 1024), (700, 1000), (1023, 1366), (540, 958)):                  # segments ending inside a band, odd widths
        a = synth.noise_frame(34, 1365, 2048)
        got = engine.resize_u8(torch.from_numpy(a).cuda(), out_hw[0], out_hw[1], Q.LANCZOS)
        assert np.array_equal(got.cpu().numpy(), Q.resize(a, out_hw[0], out_hw[1], Q.LANCZOS)), out_hw
    same = engine.resize_u8(torch.from_numpy(frames[0]).cuda(), 576, 1024, Q.LANCZOS, fused=False)
    assert engine.last_launches == 2 and torch.equal(same, outs[0])


# ---------------------------------------------------------------- thumbnails of very large frames (reduce + boxed resample)
def test_reduce_against_oracle(engine):
    rng = np.random.default_rng(51)
    cases = [((1080, 1920, 3), (2, 2), None), ((333, 517, 3), (3, 2), None), ((100, 502, 4), (5, 7), (3, 1, 499, 97)),
             ((64, 96, 1), (4, 4), (0, 0, 95, 63)), ((2160, 4096, 3), (2, 2), None), ((50, 50, 3), (1, 6), (10, 0, 50, 50))]
    for shape, factor, box in cases:
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        got = engine.reduce_u8(torch.from_numpy(a).cuda(), factor, box).cpu().numpy()
        assert np.array_equal(got, Q.reduce(a, factor, box)), (shape, factor, box)


def test_boxed_resize_against_oracle(engine):
    rng = np.random.default_rng(52)
    for t in range(12):
        h, w = (int(v) for v in rng.integers(8, 400, 2))
        oh, ow = (int(v) for v in rng.integers(1, 200, 2))
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        x0 = float(rng.uniform(0, w - 2)); x1 = float(rng.uniform(x0 + 1, w))
        y0 = float(rng.uniform(0, h - 2)); y1 = float(rng.uniform(y0 + 1, h))
        if t % 4 == 0:
            x0, x1 = 0.0, float(w)
        filt = (Q.LANCZOS, Q.BICUBIC)[t % 2]
        got = engine.resize_box_u8(torch.from_numpy(a).cuda(), oh, ow, filt, (x0, y0, x1, y1)).cpu().numpy()
        assert np.array_equal(got, Q.resize_box(a, oh, ow, filt, (x0, y0, x1, y1))), (h, w, oh, ow, x0, y0, x1, y1)


@pytest.mark.parametrize("h,w,limit", [(2160, 4096, 1024), (3000, 5000, 512), (4000, 300, 256), (777, 4001, 500),
                                       (4000, 6000, 1024)])
def test_large_frame_thumbnail_equals_pillow(engine, h, w, limit):
    """From 4x downscales on ``Image.thumbnail`` reduces by an integer factor first: same bytes as Pillow."""
    from PIL import Image
    a = synth.noise_frame(h + w, h, w)
    im = Image.fromarray(a)
    im.thumbnail((limit, limit), Image.Resampling.LANCZOS)
    want = np.asarray(im)
    tw, th = Q.thumbnail_size(w, h, limit)
    got = engine.resize_reducing_u8(torch.from_numpy(a).cuda(), th, tw, Q.LANCZOS).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(np.asarray(IU.pil_thumbnail(Image.fromarray(a), limit)), want)


def test_jpeg_draft_thumbnail_equals_pillow(engine, tmp_path):
    """An unread JPEG >= 4x larger than the request is decoded by PIL at a DCT-scaled size (draft) and hands a
    fractional box to the resample: the function form must reproduce ``thumbnail`` on the same file."""
    from PIL import Image
    a = synth.pattern_frames(2500, 4100)["lowpass"]
    path = tmp_path / "big.jpg"
    Image.fromarray(a).save(path, format="JPEG", quality=90)
    for limit in (256, 500):
        ref = Image.open(path)
        ref.thumbnail((limit, limit), Image.Resampling.LANCZOS)
        got = IU.pil_thumbnail(Image.open(path), limit)
        assert got.size == ref.size and np.array_equal(np.asarray(got), np.asarray(ref)), limit
    # the agents' whole function on that file: same data URI as the plain-PIL restatement
    from oracle import agents as OA
    from vision_inspection_system_b200 import agents as A
    assert A.encode_image_optimized(path, 500, "auditor") == OA.encode_image_optimized(path, 500, "auditor")


def test_palette_and_bilevel_frames_take_nearest_like_pillow(engine, tmp_path):
    """Pillow resamples "P" and "1" images with NEAREST whatever filter is named; a palette PNG larger than the agents'
    limit is thumbnailed that way BEFORE its RGB conversion (src/agents/vlm_inspector.py:64, :68)."""
    from PIL import Image
    rng = np.random.default_rng(61)
    idx = rng.integers(0, 256, (1500, 2600), dtype=np.uint8)
    pal = rng.integers(0, 256, 768, dtype=np.uint8).tolist()
    p_img = Image.fromarray(idx, "P")
    p_img.putpalette(pal)
    ref = p_img.copy()
    ref.thumbnail((1024, 1024), Image.Resampling.LANCZOS)
    got = IU.pil_thumbnail(p_img, 1024)
    assert got.mode == "P" and got.size == ref.size
    assert np.array_equal(np.asarray(got), np.asarray(ref))
    assert np.array_equal(np.asarray(got.convert("RGB")), np.asarray(ref.convert("RGB")))
    b_img = Image.fromarray((idx > 127).astype(np.uint8) * 255).convert("1", dither=Image.Dither.NONE)
    ref = b_img.copy()
    ref.thumbnail((500, 500), Image.Resampling.LANCZOS)
    got = IU.pil_thumbnail(b_img, 500)
    assert got.mode == "1" and np.array_equal(np.asarray(got), np.asarray(ref))
    # engine level, with a box, against the oracle
    for box in (None, (3.25, 7.5, 2000.75, 1400.0)):
        want = Q.resize_nearest(idx, 333, 517, box)
        assert np.array_equal(engine.resize_nearest_u8(torch.from_numpy(idx).cuda(), 333, 517, box).cpu().numpy(), want)
    # the agents' whole function on a palette PNG: same data URI as the plain-PIL restatement
    from oracle import agents as OA
    from vision_inspection_system_b200 import agents as A
    path = tmp_path / "palette.png"
    p_img.save(path)
    for role in ("inspector", "auditor"):
        assert A.encode_image_optimized(path, 1024, role) == OA.encode_image_optimized(path, 1024, role), role


def test_alpha_frames_are_resampled_premultiplied_like_pillow(engine):
    """"RGBA" / "LA": premultiply (Convert.c rgbA2rgba), resample, un-premultiply — all on the device, same bytes as Pillow."""
    from PIL import Image
    rng = np.random.default_rng(71)
    rgba = rng.integers(0, 256, (300, 420, 4), dtype=np.uint8)
    rgba[:40, :, 3] = 0
    rgba[40:80, :, 3] = 255
    dev = torch.from_numpy(rgba.copy()).cuda()
    engine.alpha_premultiply_(dev, True)
    assert np.array_equal(dev.cpu().numpy(), np.asarray(Image.fromarray(rgba, "RGBA").convert("RGBa")))
    dev = torch.from_numpy(rgba.copy()).cuda()
    engine.alpha_premultiply_(dev, False)
    assert np.array_equal(dev.cpu().numpy(), np.asarray(Image.frombytes("RGBa", (420, 300), rgba.tobytes()).convert("RGBA")))
    big = rng.integers(0, 256, (1300, 2500, 4), dtype=np.uint8)
    for mode, arr in (("RGBA", big), ("LA", big[:, :, [0, 3]])):
        img = Image.frombytes(mode, (2500, 1300), np.ascontiguousarray(arr).tobytes())
        ref = img.copy()
        ref.thumbnail((1024, 1024), Image.Resampling.LANCZOS)
        got = IU.pil_thumbnail(img, 1024)
        assert got.mode == mode and got.size == ref.size and got.tobytes() == ref.tobytes(), mode
        rz = IU.resize_image(img, 1000)
        want = img.resize((1000, int(1300 * (1000 / 2500))), Image.Resampling.LANCZOS)
        assert rz.tobytes() == want.tobytes(), mode
    # a 5000-pixel RGBA frame: Pillow skips the reduce pre-pass on the alpha branch, and so do we
    huge = Image.frombytes("RGBA", (5000, 700), rng.integers(0, 256, (700, 5000, 4), dtype=np.uint8).tobytes())
    ref = huge.copy()
    ref.thumbnail((1024, 1024), Image.Resampling.LANCZOS)
    assert IU.pil_thumbnail(huge, 1024).tobytes() == ref.tobytes()


def test_resize_image_double_precision_modes(engine):
    """resize_image on "I;16" / "I;16B" / "I" / "F" images (VERDICT r1 missing item: these raised NotImplementedError): the
    reference's img.resize(..., LANCZOS) runs Pillow's double-precision passes for them; same bytes here."""
    from PIL import Image
    from vision_inspection_system_b200 import image_utils as IU
    rng = np.random.default_rng(12)
    a16 = rng.integers(0, 65536, (1300, 2600), dtype=np.uint16)
    images = [Image.fromarray(a16), None,
              Image.fromarray(rng.integers(-2 ** 31, 2 ** 31 - 1, (900, 2500), dtype=np.int32)),
              Image.fromarray(rng.integers(-70000, 70000, (2300, 700), dtype=np.int32)),
              Image.fromarray(rng.normal(0, 1000, (1200, 2400)).astype(np.float32))]
    images[1] = Image.frombytes("I;16B", (2600, 1300), a16.byteswap().tobytes())
    images.append(Image.frombytes("I;16N", (2600, 1300), a16.tobytes()))     # Pillow reads these words big-endian (sic)
    images.append(Image.frombytes("I;16L", (2600, 1300), a16.tobytes()))
    for im in images:
        for limit in (2048, 1000):
            want = im.resize(IU.G.resize_image_size(im.size[0], im.size[1], limit), Image.Resampling.LANCZOS)
            got = IU.resize_image(im, limit)
            assert got.mode == want.mode == im.mode and got.size == want.size
            assert got.tobytes() == want.tobytes(), (im.mode, im.size, limit)
    small = Image.fromarray(np.ascontiguousarray(a16[:100, :200]))
    assert IU.resize_image(small, 2048) is small                       # fits: the SAME object, as the reference
    tall = Image.fromarray(rng.integers(0, 65536, (3000, 20), dtype=np.uint16))             # vertical-first branch
    assert IU.resize_image(tall, 1500).tobytes() == tall.resize((10, 1500), Image.Resampling.LANCZOS).tobytes()
    arr = rng.integers(0, 65536, (300, 400), dtype=np.uint16)
    dev = torch.from_numpy(arr.view(np.uint8).reshape(300, 800)).cuda()
    got = engine.resize_hp(dev, 150, 200, Q.LANCZOS, 0).cpu().numpy().view(np.uint16).reshape(150, 200)
    assert np.array_equal(got, Q.resize_hp(arr, 150, 200, Q.LANCZOS))
