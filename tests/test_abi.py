"""The C-ABI library loads without a GPU and exports every symbol include/vis_b200.h declares."""
import ctypes
import re
from pathlib import Path

from vision_inspection_system_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "vis_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vis_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_header():
    L = ctypes.CDLL(str(N.LIB_PATH))
    names = declared_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/vis_b200.h but not exported"
    assert set(names) == set(N.EXPORTS), "binding list and header disagree"
    assert N.lib().vis_abi_version() == 20


def test_struct_layouts_match_header():
    # sizes implied by the field lists in include/vis_b200.h (natural C alignment)
    assert N.FRAME_DTYPE.itemsize == 8 + 8 + 4 * 4 + 8 + 8 + 8
    assert N.STRIP_DTYPE.itemsize == 5 * 4
    assert N.BOX_DTYPE.itemsize == 4 * 4 + 4 + 12
    assert N.LEAF_DTYPE.itemsize == 12 * 4
    assert N.OVERLAY_FRAME_DTYPE.itemsize == 8 + 8 + 8 + 8 + 4 * 4


def test_host_entry_points_report_errors():
    L = N.lib()
    assert L.vis_build_coeffs(0, 0, 3, None, None, None) == N.VIS_E_INVALID
    assert L.vis_last_error().startswith(b"vis_build_coeffs")
    assert L.vis_fused_supported(0, 5760, 1080, 1920, 728, 1316, 6, 6) == N.VIS_OK
    assert L.vis_fused_supported(8, 5760, 1080, 1920, 728, 1316, 6, 6) == N.VIS_E_UNSUPPORTED      # misaligned base
    assert L.vis_fused_supported(0, 1506, 100, 502, 112, 504, 5, 5) == N.VIS_E_UNSUPPORTED         # pitch % 16
    assert L.vis_fused_supported(0, 64, 3000, 20, 2800, 28, 6, 6) == N.VIS_E_UNSUPPORTED           # vertical-first
    assert L.vis_fused_supported(0, 5760, 1080, 1920, 728, 1316, 25, 25) == N.VIS_E_UNSUPPORTED    # too many taps
