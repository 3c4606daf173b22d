"""The product path never routes through the oracle or a CPU fallback."""
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "vision-inspection-system_b200"


def test_product_never_imports_oracle():
    for path in list(PKG.rglob("*.py")) + list(PKG.rglob("*.cu")) + list(PKG.rglob("*.cpp")) + list(PKG.rglob("*.h")):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), path
        assert "oracle/" not in text and "liboracle" not in text, path


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the CPU-only behaviour")
def test_engine_refuses_to_run_without_cuda():
    from vision_inspection_system_b200 import engine, image_utils
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.get_engine()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.Engine()
    from PIL import Image
    big = Image.new("RGB", (4000, 100))
    with pytest.raises(RuntimeError):
        image_utils.resize_image(big, 2048)
    small = Image.new("RGB", (100, 100))
    assert image_utils.resize_image(small, 2048) is small          # reference returns the same object when it fits


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from vision_inspection_system_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", tmp_path / "libvis_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.lib()


def test_load_image_errors(tmp_path):
    from vision_inspection_system_b200.image_utils import load_image
    with pytest.raises(FileNotFoundError):
        load_image(tmp_path / "absent.png")
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not an image")
    with pytest.raises(ValueError, match="Failed to load image"):
        load_image(bad)
