"""Builds libvis_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python vision-inspection-system_b200/build.py [--force] [-v]

Every source is compiled to its own object (in parallel; objects are cached under build/ by content hash), then linked.
The library carries a hash of everything it was built from (sources, header, flags): ``vis_source_hash()`` returns it
and the bytes ``VIS_SOURCE_HASH=<hex16>`` sit in its data segment, so ``build()`` decides "up to date" by CONTENT, never
by mtimes (meaningless after a checkout or a snapshot copy): a stale binary can neither be tested nor benchmarked.
The shared object is git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libvis_b200.so"
OBJ_DIR = PKG_DIR.parent / "build" / "obj"

SOURCES = ["vis_host.cpp", "vis_generic.cu", "vis_fused_ws.cu", "vis_fused_sched.cu", "vis_fused_sched16.cu",
           "vis_fused_dp.cu", "vis_fused_mma.cu", "vis_overlay_host.cpp", "vis_overlay.cu", "vis_quality.cu", "vis_heatmap.cu", "vis_compose.cu",
           "vis_jpeg.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
]
LINK_FLAGS = ["-shared", "-lnvjpeg"]      # nvJPEG: CUDA toolkit library (codec stage, csrc/vis_jpeg.cpp)
_MARK = b"VIS_SOURCE_HASH="


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _sources() -> list:
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def _headers() -> list:
    return sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc"))) + [INCLUDE / "vis_b200.h"]


def source_hash() -> str:
    """hex16 of every byte the library is built from: sources (in link order), headers, compiler and linker flags."""
    h = hashlib.sha256()
    for p in _sources() + _headers():
        h.update(p.name.encode() + b"\0" + p.read_bytes() + b"\0")
    h.update(" ".join(NVCC_FLAGS + LINK_FLAGS).encode())
    return h.hexdigest()[:16]


def built_hash(path: Path = LIB_PATH):
    """The hash stamped into an existing library (read from its bytes: no dlopen), or None."""
    try:
        data = path.read_bytes()
    except OSError:
        return None
    m = re.search(re.escape(_MARK) + rb"([0-9a-f]{16})", data)
    return m.group(1).decode() if m else None


def is_stale() -> bool:
    return built_hash() != source_hash()


def _compile_one(nvcc: str, src: Path, stamp: str, verbose: bool, defines=(), obj_dir: Path = OBJ_DIR) -> Path:
    h = hashlib.sha256(src.read_bytes())
    for p in _headers():
        h.update(p.read_bytes())
    h.update(" ".join([*NVCC_FLAGS, *defines]).encode())
    extra = list(defines)
    if src.name == "vis_host.cpp":                      # the one translation unit that carries the stamp
        extra.append(f'-DVIS_SOURCE_HASH_VALUE="{stamp}"')
        h.update(stamp.encode())
    obj = obj_dir / f"{src.stem}.{h.hexdigest()[:16]}.o"
    if obj.exists():
        return obj
    for old in obj_dir.glob(f"{src.stem}.*.o"):
        old.unlink()
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", str(INCLUDE), "-I", str(CSRC), "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(f"---- {src.name}\n{res.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into one shared library; no-op when the stamped hash matches the tree."""
    stamp = source_hash()
    if not force and built_hash() == stamp:
        return LIB_PATH
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    if force:
        for old in OBJ_DIR.glob("*.o"):
            old.unlink()
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(lambda s: _compile_one(nvcc, s, stamp, verbose), _sources()))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", *LINK_FLAGS, "-o", str(tmp), *map(str, objs)],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    if built_hash() != stamp:
        raise RuntimeError("the built library does not carry the source hash it was built with")
    return LIB_PATH


def build_variant(name: str, defines: list) -> Path:
    """Developer A/B builds: the same sources with extra -D flags -> variants/libvis_<name>.so (git-ignored; selected at
    run time with VIS_B200_LIB=...).  Never what build() / the tests / bench.py load."""
    nvcc = _nvcc()
    out_dir = PKG_DIR.parent / "variants"
    obj_dir = PKG_DIR.parent / "build" / f"obj_{name}"
    out_dir.mkdir(exist_ok=True)
    obj_dir.mkdir(parents=True, exist_ok=True)
    stamp = source_hash()
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(lambda s: _compile_one(nvcc, s, stamp, False, defines, obj_dir), _sources()))
    out = out_dir / f"libvis_{name}.so"
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", *LINK_FLAGS, "-o", str(out), *map(str, objs)],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":      # build.py --variant name -DX=1 -DY=2
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(out, built_hash())
