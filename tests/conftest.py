"""Shared fixtures.  GPU tests carry @pytest.mark.gpu; everything else runs on a CPU-only box."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure the C-ABI library and the oracle are built (nvcc cross-compiles without a GPU)."""
    from vision_inspection_system_b200 import build
    build.build()
    import oracle
    oracle.build()


@pytest.fixture(scope="session")
def goldens():
    return json.loads((GOLDEN_DIR / "goldens.json").read_text())


@pytest.fixture(scope="session")
def arrays():
    return dict(np.load(GOLDEN_DIR / "arrays.npz"))


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vision_inspection_system_b200.engine import get_engine
    return get_engine()


def golden_frame(rec, arrays):
    """Rebuild the input frame of a ``qwen`` golden record."""
    from vision_inspection_system_b200 import synth
    name = rec["name"]
    h, w = rec["shape"]
    if name == "mouri":
        return arrays["mouri_rgb"]
    if name.startswith("noise_"):
        seed = int(name.split("seed")[1].split("_")[0])
        return synth.noise_frame(seed, h, w)
    if name.startswith("pattern_"):
        kind = name.split("_")[1]
        return synth.pattern_frames(h, w)[kind]
    raise KeyError(name)
