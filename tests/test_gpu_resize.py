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
