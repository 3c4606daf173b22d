"""-m gpu: the callers either side of the kernels.  ``encode_image_optimized`` (GPU thumbnail + PIL's JPEG encode)
returns the SAME data URI as the reference's ``_encode_image_optimized`` — byte for byte — on every golden case; the
nvJPEG encode variant is a valid stream of the same geometry close to it; ``build_visual_evidence_images`` writes the
report's two derived panels with the overlay pixels of the oracle."""
import base64
import hashlib
import io

import numpy as np
import pytest
from PIL import Image

from oracle import agents as OA
from oracle import overlay as OV
from vision_inspection_system_b200 import agents as A
from vision_inspection_system_b200 import synth

from test_oracle_agents import write_case

pytestmark = pytest.mark.gpu


def test_data_uri_equals_the_reference(engine, goldens, tmp_path):
    for rec in goldens["agents"]:
        path = write_case(rec, tmp_path)
        uri = A.encode_image_optimized(path, rec["max_size"], rec["role"])
        assert hashlib.sha256(uri.encode()).hexdigest() == rec["uri_sha256"], rec["name"]
        assert uri == OA.encode_image_optimized(path, rec["max_size"], rec["role"])


def test_nvjpeg_encode_variant(engine, goldens, tmp_path):
    rec = next(r for r in goldens["agents"] if r["name"] == "auditor_1080p_png")
    path = write_case(rec, tmp_path)
    uri = A.encode_image_optimized(path, role="auditor", codec="nvjpeg")
    assert uri.startswith("data:image/jpeg;base64,")
    got = np.asarray(Image.open(io.BytesIO(base64.b64decode(uri.split(",", 1)[1]))).convert("RGB")).astype(np.int32)
    ref = OA.encode_image_optimized(path, role="auditor")
    want = np.asarray(Image.open(io.BytesIO(base64.b64decode(ref.split(",", 1)[1]))).convert("RGB")).astype(np.int32)
    assert got.shape == want.shape == (576, 1024, 3)
    thumb = Image.open(path)
    thumb.thumbnail((1024, 1024), Image.Resampling.LANCZOS)
    src = np.asarray(thumb).astype(np.float64)

    def psnr(a):
        return 10 * np.log10(255.0 ** 2 / np.mean((a - src) ** 2))

    # two JPEG quality-85 encodes of the same (bit-exact) thumbnail: equally close to it
    assert psnr(got) >= psnr(want) - 0.6, (psnr(got), psnr(want))


def test_visual_evidence_images(engine, tmp_path):
    import cv2
    frame, _ = synth.annotated_frame(7000, 480, 640)
    src = tmp_path / "frame.png"
    cv2.imwrite(str(src), frame)
    defects = [
        {"type": "crack", "bbox": {"x": 10, "y": 20, "width": 30, "height": 25}, "safety_impact": "CRITICAL", "confidence": "high"},
        {"type": "stain", "bbox": None, "safety_impact": "COSMETIC"},
        {"type": "dent", "bbox": {"x": 55, "y": 40, "width": 20, "height": 30}, "safety_impact": "COSMETIC", "confidence": "low"},
    ]
    state = {"image_path": str(src), "consensus": {"combined_defects": defects}, "context": {"criticality": "high"}}
    heat, annot = A.build_visual_evidence_images(state, tmp_path / "reports")
    assert heat.name == "heatmap_frame.jpg" and annot.name == "annotated_frame.jpg" and heat.exists() and annot.exists()
    # the annotated panel, through a lossless suffix: exactly the oracle's overlay of the boxes the builder derives
    want = OV.draw_bounding_boxes(frame, A.evidence_boxes(defects), "low", "high")
    from vision_inspection_system_b200 import image_utils as IU
    png = IU.draw_bounding_boxes(src, A.evidence_boxes(defects), tmp_path / "a.png", "low", "high")
    assert np.array_equal(cv2.imread(str(png)), want)
    ref_jpg = tmp_path / "ref.jpg"                                           # host codec + bit-exact pixels: the same FILE
    cv2.imwrite(str(ref_jpg), want)
    assert annot.read_bytes() == ref_jpg.read_bytes()
    # no defects: the original file is copied
    state["consensus"]["combined_defects"] = []
    _, annot2 = A.build_visual_evidence_images(state, tmp_path / "reports2")
    assert annot2.read_bytes() == src.read_bytes()
