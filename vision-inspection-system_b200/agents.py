"""Drop-in replacements for the image halves of the reference's callers either side of the kernels:

    encode_image_optimized(image_path, max_size, role)     VLMInspectorAgent._encode_image_optimized
                                                           (src/agents/vlm_inspector.py:46-88) and
                                                           VLMAuditorAgent._encode_image_optimized
                                                           (src/agents/vlm_auditor.py:85-108)
    build_visual_evidence_images(state, report_dir)        the image work of InspectionReport._build_visual_evidence
                                                           (src/reporting/pdf_generator.py:1262-1330)

``encode_image_optimized`` keeps the reference's steps — open, LANCZOS thumbnail when the frame exceeds ``max_size``, RGB
conversion for palette / alpha modes, JPEG quality 85 (60 when the stream exceeds 5 MB), base64 data URI — with the
thumbnail computed on the GPU (bit-exact against Pillow).  With ``codec="host"`` (default) the JPEG encode is PIL's own,
so the returned string EQUALS the reference's, byte for byte; ``codec="nvjpeg"`` encodes on the GPU (same quality and
subsampling, different Huffman tables / rounding: a valid stream, not the same bytes).
"""
from __future__ import annotations

import base64
import io
import logging
import shutil
from pathlib import Path

from PIL import Image

from . import geometry as G
from . import image_utils as IU

logger = logging.getLogger("vision_inspection_system_b200.agents")

# modes converted to RGB before the JPEG encode: the Inspector lists LA, the Auditor does not
_RGB_MODES = {"inspector": ("RGBA", "P", "LA"), "auditor": ("RGBA", "P")}
_DEFAULT_MAX_SIZE = {"inspector": G.INSPECTOR_MAX_SIZE, "auditor": G.AUDITOR_MAX_SIZE}


def _jpeg_bytes(img: Image.Image, quality: int, codec: str) -> bytes:
    if codec == "nvjpeg" and img.mode == "RGB":
        import numpy as np
        import torch
        from .engine import get_engine
        frame = torch.from_numpy(np.asarray(img).copy()).cuda()
        # PIL's defaults for quality 85 / 60: 4:2:0 chroma subsampling; optimize=True = optimised Huffman tables
        return get_engine().jpeg_codec().encode(frame, quality=quality, subsampling="4:2:0", bgr=False, optimize=True)
    buffer = io.BytesIO()
    img.save(buffer, format="JPEG", quality=quality, optimize=True)
    return buffer.getvalue()


def encode_image_optimized(image_path: Path, max_size: int | None = None, role: str = "inspector",
                           codec: str = "host") -> str:
    """Base64 JPEG data URI of the frame at ``image_path``, resized and compressed like the agents do.

    ``role``: "inspector" (default ``max_size`` 2048, LA converted, ``ValueError`` above 10 MB — vlm_inspector.py:46-88)
    or "auditor" (default 1024 — vlm_auditor.py:85-108)."""
    if role not in _RGB_MODES:
        raise ValueError(f"role must be 'inspector' or 'auditor', got {role!r}")
    IU._check_codec(codec)
    max_size = max_size or _DEFAULT_MAX_SIZE[role]
    img = Image.open(image_path)
    original_size = img.size
    if max(img.size) > max_size:
        img = IU.pil_thumbnail(img, max_size)
        logger.debug("Resized image from %s to %s", original_size, img.size)
    if img.mode in _RGB_MODES[role]:
        img = img.convert("RGB")
    data = _jpeg_bytes(img, 85, codec)
    if len(data) > 5_000_000:
        logger.debug("Image still large (%d bytes), reducing quality", len(data))
        data = _jpeg_bytes(img, 60, codec)
    if role == "inspector" and len(data) > 10_000_000:
        raise ValueError(f"Image too large even after optimization: {len(data)} bytes")
    logger.debug("Encoded image: %d bytes", len(data))
    return f"data:image/jpeg;base64,{base64.b64encode(data).decode()}"


def evidence_boxes(defects: list) -> list:
    """Box dicts for ``draw_bounding_boxes`` from consensus defects (src/reporting/pdf_generator.py:1291-1306): label
    ``#i`` counts every defect (1-based), defects without a dict ``bbox`` are skipped with a warning."""
    boxes = []
    for i, defect in enumerate(defects, 1):
        bbox = defect.get("bbox")
        if bbox is None or not isinstance(bbox, dict):
            logger.warning("Defect #%d has no valid bbox, skipping annotation", i)
            continue
        boxes.append({"x": bbox.get("x", 0), "y": bbox.get("y", 0), "width": bbox.get("width", 0),
                      "height": bbox.get("height", 0), "label": f"#{i}",
                      "severity": defect.get("safety_impact", "MODERATE"), "confidence": defect.get("confidence", "medium")})
    return boxes


def build_visual_evidence_images(state: dict, report_dir: Path, codec: str = "host"):
    """The two derived panels of the report's visual evidence: ``heatmap_<stem>.jpg`` and ``annotated_<stem>.jpg`` in
    ``report_dir`` (src/reporting/pdf_generator.py:1262-1330).  Returns (heatmap_path, annotated_path), or None when
    the state's image does not exist (the reference prints "Image not available")."""
    image_path = Path(state.get("image_path", ""))
    if not image_path.exists():
        return None
    defects = state.get("consensus", {}).get("combined_defects", [])
    criticality = state.get("context", {}).get("criticality", "medium")
    report_dir = Path(report_dir)
    report_dir.mkdir(parents=True, exist_ok=True)     # the reference creates REPORT_DIR at import (utils/config.py:354-356)
    heatmap_path = report_dir / f"heatmap_{image_path.stem}.jpg"
    annotated_path = report_dir / f"annotated_{image_path.stem}.jpg"
    IU.create_heatmap_overlay(image_path, defects, heatmap_path, actual_model_size=IU.MAX_IMAGE_DIMENSION,
                              confidence_threshold="low", criticality=criticality, codec=codec)
    boxes = evidence_boxes(defects) if defects else []
    if boxes:
        IU.draw_bounding_boxes(image_path, boxes, annotated_path, confidence_threshold="low", criticality=criticality,
                               codec=codec)
    else:                                     # no defects / no valid boxes: the original is copied (:1318-1323)
        shutil.copy(image_path, annotated_path)
    return heatmap_path, annotated_path
