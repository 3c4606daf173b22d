"""-m gpu: the overlay rasteriser (through the C ABI) against the drawing oracle and overlays captured from the
reference's own draw_bounding_boxes.  Bar: every pixel identical."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import overlay as OV
from vision_inspection_system_b200 import synth

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_goldens(engine, goldens):
    for rec in goldens["overlay"]:
        h, w = rec["shape"]
        frame = np.random.default_rng(rec["seed"]).integers(0, 256, (h, w, 3), dtype=np.uint8)
        dev = torch.from_numpy(frame).cuda()
        out = engine.annotate([dev], [rec["boxes"]], rec["confidence_threshold"], rec["criticality"])[0]
        got = out.cpu().numpy()
        assert int((got != frame).any(2).sum()) == rec["changed_pixels"], rec["name"]
        assert sha(got) == rec["sha256"], rec["name"]
        assert torch.equal(dev.cpu(), torch.from_numpy(frame))              # out of place: input untouched
        engine.annotate([dev], [rec["boxes"]], rec["confidence_threshold"], rec["criticality"], inplace=True)
        assert sha(dev.cpu().numpy()) == rec["sha256"], rec["name"] + " (in place)"


def test_batch_against_oracle(engine):
    items = [synth.annotated_frame(7000 + i) for i in range(24)]
    frames = np.stack([f for f, _ in items])
    boxes = [b for _, b in items]
    out = engine.annotate(torch.from_numpy(frames).cuda(), boxes).cpu().numpy()
    for i, (f, b) in enumerate(items):
        assert np.array_equal(out[i], OV.draw_bounding_boxes(f, b)), i


def test_mixed_sizes_odd_widths_and_empty_lists(engine):
    shapes = [(480, 640), (333, 517), (97, 211), (1080, 1920), (65, 130)]
    items = [synth.annotated_frame(40 + i, *s) for i, s in enumerate(shapes)]
    items[2] = (items[2][0], [])                                            # no boxes: plain copy
    outs = engine.annotate([torch.from_numpy(f).cuda() for f, _ in items], [b for _, b in items])
    for (f, b), o in zip(items, outs):
        assert np.array_equal(o.cpu().numpy(), OV.draw_bounding_boxes(f, b))


def test_text_labels_beyond_digits(engine):
    """Labels are any text of any length (the reference passes '#<int>', cv2.putText accepts anything and draws '?'
    for every byte outside printable ASCII; labels wider than the frame are clipped; non-strings fall back to the index)."""
    texts = ["#A7", "crack-12", "Z", "a|b", "#(x)", "Q9%", "~", "ok!", "{[<>]}", "W_m",
             "a label far longer than eleven bytes", "d\xe9faut #2", "\u6b20\u9665", "", "x" * 300, 17]   # never raises
    items = []
    for k, shape in enumerate(((480, 640), (1080, 1920), (333, 517))):
        frame, boxes = synth.annotated_frame(60 + k, *shape)
        for i, b in enumerate(boxes):
            b["label"] = texts[(2 * k + i) % len(texts)]
        items.append((frame, boxes))
    outs = engine.annotate([torch.from_numpy(f).cuda() for f, _ in items], [b for _, b in items])
    for (f, b), o in zip(items, outs):
        assert np.array_equal(o.cpu().numpy(), OV.draw_bounding_boxes(f, b))


def test_draw_bounding_boxes_files(engine, tmp_path):
    import cv2
    from vision_inspection_system_b200 import image_utils as IU
    frame, boxes = synth.annotated_frame(7003, 480, 640)
    src, dst = tmp_path / "in.png", tmp_path / "annotated.png"             # lossless codec: pixels comparable
    cv2.imwrite(str(src), frame)
    assert IU.draw_bounding_boxes(src, boxes, dst, confidence_threshold="low", criticality="medium") == dst
    assert np.array_equal(cv2.imread(str(dst)), OV.draw_bounding_boxes(frame, boxes))
    with pytest.raises(ValueError, match="Failed to load image"):
        IU.draw_bounding_boxes(tmp_path / "missing.png", boxes, dst)


def test_config4_full_size_properties(engine):
    """BASELINE config 4 at full size (1024 annotated 1080p frames): size-independent properties + oracle on a few."""
    n, distinct = 1024, 32
    items = [synth.annotated_frame(7000 + i) for i in range(distinct)]
    items[5] = (items[5][0], [])                                             # a frame without boxes: plain copy
    frames = torch.from_numpy(np.stack([f for f, _ in items])).cuda().repeat(n // distinct, 1, 1, 1).contiguous()
    boxes = [items[i % distinct][1] for i in range(n)]
    out = engine.annotate(frames, boxes)
    torch.cuda.synchronize()
    assert torch.equal(out, out[:distinct].repeat(n // distinct, 1, 1, 1))   # replicas identical
    assert torch.equal(out[5], frames[5])                                    # no boxes -> untouched copy
    for i in (0, 7, 31):
        assert np.array_equal(out[i].cpu().numpy(), OV.draw_bounding_boxes(*items[i])), i
    again = engine.annotate(out, [[] for _ in range(n)])                     # drawing nothing is the identity
    assert torch.equal(again, out)
    work = frames.clone()
    engine.annotate(work, boxes, inplace=True)                               # in place == out of place
    assert torch.equal(work, out)
    changed = (out[:distinct] != frames[:distinct]).any(3).sum().item()
    assert 0 < changed < 0.06 * distinct * 1080 * 1920                       # overlays touch a few percent of the pixels
