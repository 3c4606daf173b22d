"""-m gpu: uint8 resampling (resize_image, agent thumbnails) against the oracle and the golden vectors."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import qwen2vl as Q
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
