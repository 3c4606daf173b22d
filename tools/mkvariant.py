"""Build an A/B variant of libvis_b200.so with extra nvcc flags:  python tools/mkvariant.py NAME [-DFOO=1 ...]
The library goes to variants/NAME.so (git-ignored, travels to the GPU box); select it with VIS_B200_LIB."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import build as B  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
B.NVCC_FLAGS[:0] = extra
(ROOT / "variants").mkdir(exist_ok=True)
B.LIB_PATH = ROOT / "variants" / f"{name}.so"
B.build(force=True)
print(B.LIB_PATH)
