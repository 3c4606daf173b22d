"""-m gpu: the nvJPEG codec stage (vis_jpeg_* through the C ABI).  nvJPEG is not libjpeg-turbo: the decoded pixels are
specified with a TOLERANCE against the reference's decoders (PIL / cv2), measured on this GPU pool (tools/jpeg_probe.py,
profiles/r01_jpeg_probe.jsonl: max 5 levels, mean 0.70 with interpolated chroma upsampling):

    max |nvJPEG - libjpeg-turbo| <= 6 levels, mean <= 0.8 levels        (4:4:4, 4:2:2, 4:2:0, progressive, gray)

Encodes are checked through a round trip: the PSNR of the nvJPEG stream is within 0.6 dB of cv2.imwrite's at the same
quality / subsampling.  Everything downstream of the decode stays bit-exact (checked by feeding the decoded frame to both
the GPU path and the oracle)."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import overlay as OV
from oracle import qwen2vl as Q
from vision_inspection_system_b200 import image_utils as IU
from vision_inspection_system_b200 import synth

pytestmark = pytest.mark.gpu

MAX_ABS, MEAN_ABS = 6, 0.8


def jpeg_bytes(rgb, quality=90, **kw):
    buf = io.BytesIO()
    Image.fromarray(rgb).save(buf, format="JPEG", quality=quality, **kw)
    return buf.getvalue()


def frames():
    return {"lowpass_1080p": synth.pattern_frames(1080, 1920)["lowpass"], "noise_vga": synth.noise_frame(2, 480, 640),
            "hgrad_odd": synth.pattern_frames(333, 517)["hgrad"], "checker_small": synth.pattern_frames(64, 96)["checker"]}


def variants(rgb):
    yield "444", jpeg_bytes(rgb, subsampling=0)
    yield "422", jpeg_bytes(rgb, subsampling=1)
    yield "420", jpeg_bytes(rgb, subsampling=2)
    yield "progressive", jpeg_bytes(rgb, 85, progressive=True)
    yield "gray", jpeg_bytes(np.asarray(Image.fromarray(rgb).convert("L")))


@pytest.mark.parametrize("backend", ["gpu_hybrid", "hybrid"])
def test_decode_within_tolerance_of_the_host_decoders(engine, backend):
    import cv2
    codec = engine.jpeg_codec(backend)
    for name, rgb in frames().items():
        for tag, data in variants(rgb):
            want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
            got_t = codec.decode(data)
            assert got_t.stride(0) % 16 == 0 and got_t.stride(1) == 3            # ready for the fused kernels
            got = got_t.cpu().numpy().astype(np.int32)
            d = np.abs(got - want)
            assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS, (name, tag, int(d.max()), float(d.mean()))
            bgr = codec.decode(data, bgr=True).cpu().numpy()
            assert np.array_equal(bgr[:, :, ::-1], got.astype(np.uint8)), (name, tag)
            cv = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR).astype(np.int32)
            assert np.abs(bgr.astype(np.int32) - cv).max() <= MAX_ABS, (name, tag)


def test_batched_decode_equals_single_decodes(engine):
    codec = engine.jpeg_codec()
    streams = [d for rgb in frames().values() for _, d in variants(rgb)]
    outs = codec.decode_batch(streams)
    assert len(outs) == len(streams)
    for s, o in zip(streams, outs):
        assert torch.equal(o.cpu(), codec.decode(s).cpu())
    assert codec.decode_batch([]) == []


def test_decode_rejects_garbage(engine):
    from vision_inspection_system_b200._native import VisError
    codec = engine.jpeg_codec()
    with pytest.raises(VisError):
        codec.decode(b"\xff\xd8\xff\xe0 not a jpeg at all")
    good = jpeg_bytes(synth.noise_frame(1, 64, 96))
    with pytest.raises(VisError):
        codec.decode(good[:200])


def test_encode_round_trip_matches_cv2_quality(engine):
    import cv2
    codec = engine.jpeg_codec()
    bgr = np.ascontiguousarray(synth.pattern_frames(1080, 1920)["lowpass"][:, :, ::-1])
    dev = torch.from_numpy(bgr).cuda()
    src = bgr.astype(np.float64)

    def psnr(a):
        return 10 * np.log10(255.0 ** 2 / np.mean((a.astype(np.float64) - src) ** 2))

    for q, sub, flag in ((95, "4:2:0", []), (85, "4:2:0", []),
                         (95, "4:4:4", [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])):
        data = codec.encode(dev, q, sub)
        back = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert back.shape == bgr.shape
        ok, ref = cv2.imencode(".jpg", bgr, [cv2.IMWRITE_JPEG_QUALITY, q] + flag)
        ref_back = cv2.imdecode(ref, cv2.IMREAD_COLOR)
        assert psnr(back) >= psnr(ref_back) - 0.6, (q, sub, psnr(back), psnr(ref_back))
        assert 0.8 * len(ref) <= len(data) <= 1.2 * len(ref)
        pil = np.asarray(Image.open(io.BytesIO(data)))                       # PIL reads it too (RGB order)
        assert np.array_equal(pil[:, :, ::-1], back)
    # a pitched view (a crop of a wider tensor) encodes the crop
    wide = torch.from_numpy(np.ascontiguousarray(np.concatenate([bgr, bgr], 1))).cuda()
    crop = cv2.imdecode(np.frombuffer(codec.encode(wide[:, :1920]), np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(crop, cv2.imdecode(np.frombuffer(codec.encode(dev), np.uint8), cv2.IMREAD_COLOR))


def test_pipeline_after_the_decode_stays_bit_exact(engine, tmp_path):
    """codec="nvjpeg": the decoded frame differs from the host decode by the tolerance above, but pixel_values and the
    overlay computed FROM that decoded frame are bit-exact against the oracle."""
    import cv2
    rgb = synth.pattern_frames(1080, 1920)["lowpass"]
    path = tmp_path / "frame.jpg"
    path.write_bytes(jpeg_bytes(rgb, subsampling=2))
    dec = IU.decode_image(path)                                             # RGB, CUDA, nvJPEG
    host = np.asarray(Image.open(path).convert("RGB")).astype(np.int32)
    assert np.abs(dec.cpu().numpy().astype(np.int32) - host).max() <= MAX_ABS
    pv, grid = IU.preprocess_for_vlm([path, path], codec="nvjpeg")
    want, wgrid = Q.preprocess([dec.cpu().numpy()] * 2)
    assert np.array_equal(grid.numpy(), wgrid) and np.array_equal(pv.cpu().numpy(), want)
    pv_host, _ = IU.preprocess_for_vlm([path])                               # default codec: PIL decode, as the reference
    assert np.array_equal(pv_host.cpu().numpy(), Q.preprocess([host.astype(np.uint8)])[0])
    # overlay: nvJPEG decode -> draw -> nvJPEG encode; compare before the encode through the PNG route
    _, boxes = synth.annotated_frame(7000)
    out_png = IU.draw_bounding_boxes(path, boxes, tmp_path / "annot.png", codec="nvjpeg")
    bgr = IU.decode_image(path, bgr=True).cpu().numpy()
    assert np.array_equal(cv2.imread(str(out_png)), OV.draw_bounding_boxes(bgr, boxes))
    out_jpg = IU.draw_bounding_boxes(path, boxes, tmp_path / "annot.jpg", codec="nvjpeg")
    back = cv2.imread(str(out_jpg)).astype(np.int32)
    ref_jpg = tmp_path / "annot_ref.jpg"
    cv2.imwrite(str(ref_jpg), OV.draw_bounding_boxes(bgr, boxes))
    assert np.abs(back - cv2.imread(str(ref_jpg)).astype(np.int32)).mean() <= 2 * MEAN_ABS
    with pytest.raises(FileNotFoundError):
        IU.decode_image(tmp_path / "missing.jpg")
    (tmp_path / "bad.jpg").write_bytes(b"\xff\xd8\xff garbage")
    with pytest.raises(ValueError, match="Failed to load image"):
        IU.decode_image(tmp_path / "bad.jpg")
    with pytest.raises(ValueError):
        IU.draw_bounding_boxes(path, boxes, tmp_path / "x.jpg", codec="turbo")


def test_exif_orientation_follows_cv2(tmp_path):
    """A phone-camera JPEG with EXIF Orientation 6 / 8 / 3: cv2.imread (the reference's decoder for the overlay, heat map
    and comparison panel, utils/image_utils.py:170) rotates it, nvJPEG does not — such files keep the host decoder, so
    the frame, its H/W and the pixels the percent boxes land on equal the reference's exactly."""
    import cv2
    from vision_inspection_system_b200.jpeg import exif_orientation
    rgb = synth.pattern_frames(480, 640)["lowpass"]
    _, boxes = synth.annotated_frame(7000, 480, 640)
    for orientation in (6, 8, 3):
        buf = io.BytesIO()
        ex = Image.Exif()
        ex[0x0112] = orientation
        Image.fromarray(rgb).save(buf, "JPEG", quality=90, exif=ex)
        path = tmp_path / f"o{orientation}.jpg"
        path.write_bytes(buf.getvalue())
        assert exif_orientation(buf.getvalue()) == orientation
        want = cv2.imread(str(path))
        assert want.shape == ((640, 480, 3) if orientation in (6, 8) else (480, 640, 3))
        got = IU.decode_image(path, bgr=True, codec="nvjpeg")
        assert np.array_equal(got.cpu().numpy(), want)
        out = IU.draw_bounding_boxes(path, boxes, tmp_path / f"o{orientation}.png", codec="nvjpeg")
        assert np.array_equal(cv2.imread(str(out)), OV.draw_bounding_boxes(want, boxes))
    assert exif_orientation(jpeg_bytes(rgb)) == 1 and exif_orientation(b"\xff\xd8\xff\xe1") == 1
