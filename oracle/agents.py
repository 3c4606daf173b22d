"""Oracle (TEST INFRASTRUCTURE ONLY): the agents' ``_encode_image_optimized`` restated with plain PIL calls.

Follows src/agents/vlm_inspector.py:59-88 (Inspector: converts RGBA / P / LA, raises above 10 MB) and
src/agents/vlm_auditor.py:87-108 (Auditor: converts RGBA / P, no final check).  Pinned by tests/golden: the reference's
own function bodies, extracted unmodified from those files and executed (tests/golden/make_goldens.py, "agents").
"""
from __future__ import annotations

import base64
import io

from PIL import Image


def encode_image_optimized(image_path, max_size=None, role="inspector") -> str:
    max_size = max_size or (2048 if role == "inspector" else 1024)
    img = Image.open(image_path)
    if max(img.size) > max_size:
        img.thumbnail((max_size, max_size), Image.Resampling.LANCZOS)
    if img.mode in (("RGBA", "P", "LA") if role == "inspector" else ("RGBA", "P")):
        img = img.convert("RGB")
    buffer = io.BytesIO()
    img.save(buffer, format="JPEG", quality=85, optimize=True)
    if buffer.tell() > 5_000_000:
        buffer = io.BytesIO()
        img.save(buffer, format="JPEG", quality=60, optimize=True)
    if role == "inspector" and buffer.tell() > 10_000_000:
        raise ValueError(f"Image too large even after optimization: {buffer.tell()} bytes")
    return f"data:image/jpeg;base64,{base64.b64encode(buffer.getvalue()).decode()}"
