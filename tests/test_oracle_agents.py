"""CPU: the agents' ``_encode_image_optimized`` oracle (oracle/agents.py, plain PIL) against data URIs produced by the
reference's own function bodies (tests/golden, "agents"), and the host-only pieces of the product's agents module."""
import hashlib

import pytest

from oracle import agents as OA
from vision_inspection_system_b200 import agents as A
from vision_inspection_system_b200 import synth


def write_case(rec, tmp_path):
    path = tmp_path / f"{rec['name']}{rec['suffix']}"
    synth.write_agent_input(synth.agent_input_image(rec["seed"], tuple(rec["shape"]), rec["mode"]), path)
    assert hashlib.sha256(path.read_bytes()).hexdigest() == rec["input_sha256"], "input file differs from the golden run"
    return path


def test_oracle_matches_reference_data_uris(goldens, tmp_path):
    for rec in goldens["agents"]:
        path = write_case(rec, tmp_path)
        uri = OA.encode_image_optimized(path, rec["max_size"], rec["role"])
        assert hashlib.sha256(uri.encode()).hexdigest() == rec["uri_sha256"], rec["name"]


def test_evidence_boxes_follow_the_report_builder():
    defects = [
        {"bbox": {"x": 10, "y": 20, "width": 30, "height": 15}, "safety_impact": "CRITICAL", "confidence": "high"},
        {"bbox": None},
        {"bbox": "not a dict"},
        {"bbox": {"x": 1.5}, "confidence": "low"},
    ]
    boxes = A.evidence_boxes(defects)
    assert boxes == [
        {"x": 10, "y": 20, "width": 30, "height": 15, "label": "#1", "severity": "CRITICAL", "confidence": "high"},
        {"x": 1.5, "y": 0, "width": 0, "height": 0, "label": "#4", "severity": "MODERATE", "confidence": "low"},
    ]
    assert A.evidence_boxes([]) == []


def test_argument_errors():
    with pytest.raises(ValueError):
        A.encode_image_optimized("x.png", role="explainer")
    with pytest.raises(ValueError):
        A.encode_image_optimized("x.png", codec="turbo")
    assert A.build_visual_evidence_images({"image_path": "/nonexistent/frame.jpg"}, "/tmp") is None
