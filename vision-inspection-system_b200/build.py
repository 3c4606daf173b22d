"""Builds libvis_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python vision-inspection-system_b200/build.py [--force]

The shared object is git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libvis_b200.so"

SOURCES = ["vis_host.cpp", "vis_generic.cu", "vis_fused.cu", "vis_fused_ws.cu", "vis_fused_sched.cu", "vis_fused_sched16.cu", "vis_overlay_host.cpp", "vis_overlay.cu", "vis_quality.cu", "vis_heatmap.cu", "vis_compose.cu", "vis_jpeg.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-shared",
    "-lnvjpeg",                      # CUDA toolkit library (codec stage, csrc/vis_jpeg.cpp)
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*")) + [INCLUDE / "vis_b200.h", Path(__file__)])
    return newest > LIB_PATH.stat().st_mtime


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into one shared library; no-op when up to date."""
    if not force and not is_stale():
        return LIB_PATH
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(INCLUDE), "-I", str(CSRC), "-o", str(LIB_PATH), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(out)
