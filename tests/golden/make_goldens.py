"""Generates tests/golden/goldens.json + arrays.npz from the REAL reference and its dependencies.

Run in the build container only (needs /root/reference, Pillow, opencv, transformers):

    python tests/golden/make_goldens.py

What is recorded (versions of every binary are stored next to the vectors):
  * ``qwen``       Qwen2VLImageProcessorPil (transformers, PIL backend) outputs — grid_thw, sha256 of pixel_values,
                   three sample rows — for seeded frames, pattern frames, Mouri.jpg and edge geometries.
  * ``thumbnail``  PIL ``Image.thumbnail(LANCZOS)`` as called by ``_encode_image_optimized`` (src/agents/*.py) and the
                   reference's own ``utils.image_utils.resize_image`` — sha256 of the resized bytes.
  * ``overlay``    the reference's own ``utils.image_utils.draw_bounding_boxes`` executed unmodified, with
                   ``cv2.imread`` / ``cv2.imwrite`` intercepted so the BGR array it holds right before the
                   JPEG encode is captured — sha256 + changed-pixel count.
  * ``heatmap``    the reference's own ``utils.image_utils.create_heatmap_overlay`` executed unmodified (imread / imwrite
                   intercepted): the BGR array before the JPEG encode, stored in arrays.npz (tolerance-compared).
  * ``quality``    the reference's own ``src.safety.image_quality.assess_image_quality`` on lossless PNG files of seeded
                   frames — the full result dict.
  * ``compare``    the reference's own ``create_side_by_side_comparison`` and ``create_status_stamp`` executed unmodified
                   (imread / imwrite intercepted): sha256 of the array handed to ``cv2.imwrite``.
  * ``agents``     the reference's own ``_encode_image_optimized`` bodies (Inspector and Auditor), extracted unmodified from
                   src/agents/*.py and executed on seeded input files: sha256 of the returned data URI.
The reference has no tests or vectors of its own for this path; these fixtures are the pin.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path("/root/reference")
sys.path.insert(0, str(REPO))

from vision_inspection_system_b200 import synth  # noqa: E402  (seeded workload definitions, numpy only)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def import_reference_image_utils():
    """Import /root/reference/utils/image_utils.py unmodified (colorlog shim, dummy key, temp cwd)."""
    shim = types.ModuleType("colorlog")
    import logging

    class ColoredFormatter(logging.Formatter):
        def __init__(self, fmt=None, datefmt=None, *a, **k):
            super().__init__("%(message)s", datefmt)

    shim.ColoredFormatter = ColoredFormatter
    shim.StreamHandler = logging.StreamHandler
    sys.modules.setdefault("colorlog", shim)
    os.environ.setdefault("HUGGINGFACE_API_KEY", "hf_dummy_key_for_goldens_0000000000")
    tmp = tempfile.mkdtemp(prefix="ref_cwd_")
    os.chdir(tmp)                       # utils/config.py creates uploads/ reports/ logs/ at import
    sys.path.insert(0, str(REFERENCE))
    return importlib.import_module("utils.image_utils")


def qwen_cases():
    pats = synth.pattern_frames(1080, 1920)
    small = synth.pattern_frames(140, 252)
    cases = [
        ("noise_1080p_seed1234", synth.noise_frame(1234, 1080, 1920), None),
        ("noise_1080p_seed1235", synth.noise_frame(1235, 1080, 1920), None),
        ("noise_1080p_seed1234_hub", synth.noise_frame(1234, 1080, 1920), 12845056),
        ("noise_4k_seed4000", synth.noise_frame(4000, 2160, 3840), None),
        ("noise_720p_seed1", synth.noise_frame(1, 720, 1280), None),
        ("noise_480p_seed2", synth.noise_frame(2, 480, 640), None),
        ("noise_2048x1536_seed3", synth.noise_frame(3, 1536, 2048), None),
        ("noise_thumb2048_seed4", synth.noise_frame(4, 1152, 2048), None),
        ("noise_thumb1024_seed5", synth.noise_frame(5, 576, 1024), None),
        ("noise_odd_333x517_seed6", synth.noise_frame(6, 333, 517), None),
        ("noise_tall_3000x20_seed7", synth.noise_frame(7, 3000, 20), None),
        ("noise_wide_20x3000_seed8", synth.noise_frame(8, 20, 3000), None),
        ("noise_tiny_10x10_seed9", synth.noise_frame(9, 10, 10), None),
        ("noise_1x1_seed10", synth.noise_frame(10, 1, 1), None),
        ("noise_56x56_seed11", synth.noise_frame(11, 56, 56), None),
        ("noise_small_64x96_seed12", synth.noise_frame(12, 64, 96), None),
    ]
    cases += [(f"pattern_{k}_1080p", v, None) for k, v in pats.items()]
    cases += [(f"pattern_{k}_140x252", v, None) for k, v in small.items()]
    return cases


def quality_cases():
    """BGR frames for the image-quality goldens (name -> array); regenerated from seeds by the tests."""
    pats = synth.pattern_frames(1080, 1920)
    return {
        "noise_1080p_seed1234": synth.noise_frame(1234, 1080, 1920),
        "lowpass_1080p": pats["lowpass"], "vgrad_1080p": pats["vgrad"], "zeros_1080p": pats["zeros"],
        "full_1080p": pats["full"], "checker_1080p": pats["checker"],
        "noise_vga_seed2": synth.noise_frame(2, 480, 640), "noise_odd_333x517_seed6": synth.noise_frame(6, 333, 517),
        "noise_small_64x96_seed12": synth.noise_frame(12, 64, 96), "noise_99x101_seed13": synth.noise_frame(13, 99, 101),
        "noise_100x100_seed14": synth.noise_frame(14, 100, 100), "noise_1x1_seed10": synth.noise_frame(10, 1, 1),
        "noise_1x7_seed15": synth.noise_frame(15, 1, 7), "dark_vga": (synth.noise_frame(16, 480, 640) // 12).astype(np.uint8),
        "bright_vga": (255 - synth.noise_frame(17, 480, 640) // 12).astype(np.uint8),
    }


def quality_goldens():
    """The reference's own ImageQualityAssessment (src/safety/image_quality.py) on lossless files of quality_cases()."""
    import cv2
    import_reference_image_utils()                      # shims + sys.path for the reference tree
    iq = importlib.import_module("src.safety.image_quality")
    tmp = Path(tempfile.mkdtemp(prefix="iq_"))
    out = []
    cases = dict(quality_cases())
    mouri = np.load(HERE / "arrays.npz")["mouri_rgb"]
    cases["mouri_bgr"] = np.ascontiguousarray(mouri[:, :, ::-1])
    for name, bgr in cases.items():
        path = tmp / f"{name}.png"
        cv2.imwrite(str(path), bgr)
        res = iq.assess_image_quality(path)
        res.pop("image_path", None)
        out.append({"name": name, "shape": list(bgr.shape[:2]), "input_sha256": sha(bgr), "result": res})
        print("quality", name, res.get("quality_score"), res.get("sharpness", {}).get("laplacian_variance"))
    return out


def heatmap_goldens(arrays: dict):
    """The reference's own create_heatmap_overlay (utils/image_utils.py:320-604), cv2.imread / imwrite intercepted; the
    BGR array it hands to imwrite is stored (subsampled for the 1080p case) in arrays.npz as heatmap_<name>."""
    ref = import_reference_image_utils()
    captured = {}
    real_imread, real_imwrite = ref.cv2.imread, ref.cv2.imwrite
    out = []
    for name, frame, defects, step in synth.heatmap_cases():
        captured["input"] = frame
        ref.cv2.imread = lambda path, *a: captured["input"].copy()
        ref.cv2.imwrite = lambda path, img, *a: captured.__setitem__("output", img.copy()) or True
        try:
            ref.create_heatmap_overlay(Path("in.png"), defects, Path("heat.jpg"))
        finally:
            ref.cv2.imread, ref.cv2.imwrite = real_imread, real_imwrite
        res = captured["output"]
        arrays[f"heatmap_{name}"] = res[::step, ::step].copy()
        out.append({"name": name, "shape": list(frame.shape[:2]), "input_sha256": sha(frame), "n_defects": len(defects),
                    "subsample": step, "full_sha256": sha(res)})
        print("heatmap", name, sha(res)[:16], int((res != frame).any(2).sum()))
    return out


def label_overlay_cases():
    """draw_bounding_boxes cases whose labels are not '#<int>' (any printable ASCII)."""
    texts = ["#A7", "crack-12", "Z", "a|b", "#(x)", "Q9%", "~", "ok!", "{[<>]}", "W_m", "#", ""]
    cases = []
    for k, (seed, shape) in enumerate(((7400, (1080, 1920)), (7401, (480, 640)), (7402, (333, 517)))):
        _, boxes = synth.annotated_frame(seed, *shape)
        for i, b in enumerate(boxes):
            b["label"] = texts[(5 * k + i) % len(texts)]
        cases.append((f"labels_seed{seed}", seed, shape, boxes, "low", "medium"))
    return cases


def overlay_label_goldens():
    ref = import_reference_image_utils()
    captured = {}
    real_imread, real_imwrite = ref.cv2.imread, ref.cv2.imwrite
    out = []
    for name, seed, shape, boxes, thr, crit in label_overlay_cases():
        frame = np.random.default_rng(seed).integers(0, 256, (*shape, 3), dtype=np.uint8)
        captured["input"] = frame
        ref.cv2.imread = lambda path, *a: captured["input"].copy()
        ref.cv2.imwrite = lambda path, img, *a: captured.__setitem__("output", img.copy()) or True
        try:
            ref.draw_bounding_boxes(Path("in.png"), boxes, Path("out.jpg"), thr, crit)
        finally:
            ref.cv2.imread, ref.cv2.imwrite = real_imread, real_imwrite
        res = captured["output"]
        out.append({"name": name, "seed": seed, "shape": list(shape), "boxes": boxes, "confidence_threshold": thr,
                    "criticality": crit, "sha256": sha(res), "changed_pixels": int((res != frame).any(2).sum())})
        print("overlay", name, sha(res)[:16], out[-1]["changed_pixels"])
    return out


def compare_cases():
    """(name, (seed, h, w) of the original, (seed, h, w) of the annotated frame, labels or None)."""
    return [
        ("pair_1080p", (7500, 1080, 1920), (7501, 1080, 1920), None),
        ("pair_vga", (7502, 480, 640), (7503, 480, 640), None),
        ("pair_1600x1200_area", (7504, 1600, 1200), (7505, 1600, 1200), None),
        ("pair_odd_and_uhd", (7506, 333, 517), (7507, 2160, 3840), None),
        ("pair_800_copy", (7508, 800, 600), (7509, 800, 1000), None),
        ("pair_logo_upscale", (7510, 100, 502), (7511, 100, 502), ("Logo (raw)", "Logo #2 ~ {annotated}")),
        ("pair_mixed_area_bilinear", (7512, 1600, 1201), (7513, 1600, 900), ("A", "B")),
    ]


def stamp_cases():
    return [("SAFE", (300, 100)), ("UNSAFE", (300, 100)), ("REQUIRES_HUMAN_REVIEW", (300, 100)),
            ("SAFE", (200, 80)), ("UNSAFE", (640, 200)), ("anything", (97, 41))]


def compare_goldens():
    ref = import_reference_image_utils()
    real_imread, real_imwrite = ref.cv2.imread, ref.cv2.imwrite
    captured, out = {}, {"side_by_side": [], "stamp": []}
    ref.cv2.imwrite = lambda path, img, *a: captured.__setitem__("output", img.copy()) or True
    try:
        for name, (s1, h1, w1), (s2, h2, w2), labels in compare_cases():
            files = {"orig.png": synth.noise_frame(s1, h1, w1), "annot.png": synth.noise_frame(s2, h2, w2)}
            ref.cv2.imread = lambda path, *a: files[Path(path).name].copy()
            kw = {} if labels is None else {"labels": labels}
            ref.create_side_by_side_comparison(Path("orig.png"), Path("annot.png"), Path("cmp/out.jpg"), **kw)
            res = captured["output"]
            out["side_by_side"].append({"name": name, "original": [s1, h1, w1], "annotated": [s2, h2, w2],
                                        "labels": labels, "shape": list(res.shape), "sha256": sha(res)})
            print("compare", name, res.shape, sha(res)[:16])
        for verdict, size in stamp_cases():
            ref.create_status_stamp(verdict, Path("cmp/stamp.png"), size)
            res = captured["output"]
            out["stamp"].append({"verdict": verdict, "size": list(size), "shape": list(res.shape), "sha256": sha(res),
                                 "opaque_pixels": int((res[:, :, 3] > 0).sum())})
            print("stamp", verdict, size, sha(res)[:16])
    finally:
        ref.cv2.imread, ref.cv2.imwrite = real_imread, real_imwrite
    return out


def agent_cases():
    """(name, seed, (h, w), mode, file suffix, role, max_size or None) — inputs are written losslessly (PNG) or as the
    JPEG the case names, then handed to the reference's function as a path."""
    return [
        ("inspector_1080p_png", 7600, (1080, 1920), "RGB", ".png", "inspector", None),        # fits: no thumbnail
        ("inspector_4k_png", 7601, (2160, 3840), "RGB", ".png", "inspector", None),            # -> 2048x1152
        ("auditor_1080p_png", 7602, (1080, 1920), "RGB", ".png", "auditor", None),             # -> 1024x576
        ("auditor_4k_png", 7603, (2160, 3840), "RGB", ".png", "auditor", None),                # -> 1024x576, 25 taps
        ("inspector_rgba_png", 7604, (1200, 1600), "RGBA", ".png", "inspector", 1024),         # alpha -> premultiplied resample -> RGB
        ("auditor_gray_png", 7605, (900, 1500), "L", ".png", "auditor", None),
        ("inspector_jpeg_input", 7606, (1536, 2048), "RGB", ".jpg", "inspector", 1024),        # 2x: no draft mode
        ("auditor_small_png", 7607, (480, 640), "RGB", ".png", "auditor", None),
    ]


def agent_goldens():
    """The reference's own ``_encode_image_optimized`` bodies (src/agents/vlm_inspector.py, vlm_auditor.py), extracted
    UNMODIFIED from the files with ``ast`` (the modules themselves import langchain / groq, absent here) and executed."""
    import ast
    import base64
    import io
    import logging
    from typing import Optional
    from PIL import Image
    funcs = {}
    for role, rel in (("inspector", "src/agents/vlm_inspector.py"), ("auditor", "src/agents/vlm_auditor.py")):
        src = (REFERENCE / rel).read_text()
        tree = ast.parse(src)
        fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "_encode_image_optimized")
        ns = {"Image": Image, "io": io, "base64": base64, "Path": Path, "Optional": Optional}
        exec(compile(ast.Module([fn], []), rel, "exec"), ns)
        funcs[role] = ns["_encode_image_optimized"]

    class Self:
        max_image_size = 2048                     # config.max_image_dimension (utils/config.py:184)
        logger = logging.getLogger("ref_agent")

    tmp = Path(tempfile.mkdtemp(prefix="agent_"))
    out = []
    for name, seed, shape, mode, suffix, role, max_size in agent_cases():
        path = tmp / f"{name}{suffix}"
        synth.write_agent_input(synth.agent_input_image(seed, shape, mode), path)
        args = (path,) if max_size is None else (path, max_size)
        uri = funcs[role](Self(), *args)
        data = base64.b64decode(uri.split(",", 1)[1])
        size = Image.open(io.BytesIO(data)).size
        out.append({"name": name, "seed": seed, "shape": list(shape), "mode": mode, "suffix": suffix, "role": role,
                    "max_size": max_size, "input_sha256": hashlib.sha256(path.read_bytes()).hexdigest(),
                    "uri_sha256": hashlib.sha256(uri.encode()).hexdigest(), "jpeg_bytes": len(data), "jpeg_size": list(size)})
        print("agent", name, size, len(data), out[-1]["uri_sha256"][:16])
    return out


def main():
    if "--only-overlay-labels" in sys.argv:             # append / refresh the text-label overlay cases
        out = json.loads((HERE / "goldens.json").read_text())
        out["overlay"] = [r for r in out["overlay"] if not r["name"].startswith("labels_")] + overlay_label_goldens()
        (HERE / "goldens.json").write_text(json.dumps(out, indent=1))
        print("updated", HERE / "goldens.json")
        return
    if "--only-agents" in sys.argv:                     # add / refresh the "agents" section of an existing file
        out = json.loads((HERE / "goldens.json").read_text())
        out["agents"] = agent_goldens()
        (HERE / "goldens.json").write_text(json.dumps(out, indent=1))
        print("updated", HERE / "goldens.json")
        return
    if "--only-compare" in sys.argv:                    # add / refresh the "compare" section of an existing file
        out = json.loads((HERE / "goldens.json").read_text())
        out["compare"] = compare_goldens()
        (HERE / "goldens.json").write_text(json.dumps(out, indent=1))
        print("updated", HERE / "goldens.json")
        return
    if "--only-heatmap" in sys.argv:                    # add / refresh the "heatmap" section of existing files
        out = json.loads((HERE / "goldens.json").read_text())
        arrays = dict(np.load(HERE / "arrays.npz"))
        out["heatmap"] = heatmap_goldens(arrays)
        import cv2
        arrays["jet_bgr"] = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)[0]
        np.savez_compressed(HERE / "arrays.npz", **arrays)
        (HERE / "goldens.json").write_text(json.dumps(out, indent=1))
        print("updated", HERE / "goldens.json", HERE / "arrays.npz")
        return
    if "--only-quality" in sys.argv:                    # add / refresh the "quality" section of an existing file
        out = json.loads((HERE / "goldens.json").read_text())
        out["quality"] = quality_goldens()
        (HERE / "goldens.json").write_text(json.dumps(out, indent=1))
        print("updated", HERE / "goldens.json")
        return
    import cv2
    import PIL
    import transformers
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil

    out = {"versions": {"pillow": PIL.__version__, "opencv": cv2.__version__, "transformers": transformers.__version__,
                        "numpy": np.__version__}, "qwen": [], "thumbnail": [], "overlay": []}
    arrays = {}

    # ---------------- Mouri.jpg (BASELINE config 1) ----------------
    mouri = np.asarray(Image.open(REFERENCE / "Mouri.jpg").convert("RGB"))
    arrays["mouri_rgb"] = mouri

    # ---------------- qwen ----------------
    def run_proc(frame, max_pixels):
        proc = Qwen2VLImageProcessorPil() if max_pixels is None else Qwen2VLImageProcessorPil(
            size={"shortest_edge": 56 * 56, "longest_edge": max_pixels})
        r = proc(images=[Image.fromarray(frame)], return_tensors="np")
        return r["pixel_values"], r["image_grid_thw"]

    for name, frame, max_pixels in [("mouri", mouri, None)] + qwen_cases():
        pv, grid = run_proc(frame, max_pixels)
        n = pv.shape[0]
        rows = sorted({0, n // 2, n - 1})
        out["qwen"].append({
            "name": name, "shape": list(frame.shape[:2]), "input_sha256": sha(frame),
            "max_pixels": max_pixels, "grid_thw": grid[0].tolist(), "rows": n, "sha256": sha(pv),
            "min": float(pv.min()), "max": float(pv.max()), "sample_rows": rows})
        arrays[f"qwen_{name}_samples"] = pv[rows]
        if pv.nbytes <= 200_000:
            arrays[f"qwen_{name}_full"] = pv
        print("qwen", name, grid[0].tolist(), sha(pv)[:16])

    # ---------------- thumbnails / resize_image ----------------
    ref = import_reference_image_utils()
    for name, frame, limit in [
        ("thumb_4k_to_2048", synth.noise_frame(20, 2160, 3840), 2048),
        ("thumb_4k_to_1024", synth.noise_frame(21, 2160, 3840), 1024),
        ("thumb_1080p_to_1024", synth.noise_frame(22, 1080, 1920), 1024),
        ("thumb_portrait_1600x1200_to_1024", synth.noise_frame(23, 1600, 1200), 1024),
        ("thumb_2048x1536_to_1024", synth.noise_frame(24, 1536, 2048), 1024),
        ("thumb_small_300x500_to_256", synth.noise_frame(25, 300, 500), 256),
    ]:
        im = Image.fromarray(frame)
        im.thumbnail((limit, limit), Image.Resampling.LANCZOS)      # exactly the agents' call
        t = np.asarray(im)
        rz = np.asarray(ref.resize_image(Image.fromarray(frame), limit))   # the reference's own function
        out["thumbnail"].append({"name": name, "shape": list(frame.shape[:2]), "seed": int(name and 0),
                                 "limit": limit, "thumb_size": [im.size[0], im.size[1]], "thumb_sha256": sha(t),
                                 "resize_image_size": [rz.shape[1], rz.shape[0]], "resize_image_sha256": sha(rz)})
        if t.nbytes <= 200_000:
            arrays[f"thumb_{name}_full"] = t
        print("thumb", name, im.size, sha(t)[:16], rz.shape, sha(rz)[:16])
    seeds = {"thumb_4k_to_2048": 20, "thumb_4k_to_1024": 21, "thumb_1080p_to_1024": 22,
             "thumb_portrait_1600x1200_to_1024": 23, "thumb_2048x1536_to_1024": 24, "thumb_small_300x500_to_256": 25}
    for rec in out["thumbnail"]:
        rec["seed"] = seeds[rec["name"]]

    # ---------------- overlay: the reference function itself ----------------
    captured = {}
    real_imread, real_imwrite = ref.cv2.imread, ref.cv2.imwrite

    def fake_imread(path, *a):
        return captured["input"].copy()

    def fake_imwrite(path, img, *a):
        captured["output"] = img.copy()
        return True

    def run_ref(frame, boxes, thr="low", crit="medium"):
        captured["input"] = frame
        ref.cv2.imread, ref.cv2.imwrite = fake_imread, fake_imwrite
        try:
            ref.draw_bounding_boxes(Path("in.png"), boxes, Path("out.jpg"), thr, crit)
        finally:
            ref.cv2.imread, ref.cv2.imwrite = real_imread, real_imwrite
        return captured["output"]

    overlay_cases = []
    for i in range(6):
        frame, boxes = synth.annotated_frame(7000 + i)
        overlay_cases.append((f"cfg4_seed{7000 + i}", 7000 + i, (1080, 1920), boxes, "low", "medium"))
    for i in range(6):
        frame, boxes = synth.annotated_frame(7100 + i, 480, 640)
        overlay_cases.append((f"vga_seed{7100 + i}", 7100 + i, (480, 640), boxes, "low", "medium"))
    frame, boxes = synth.annotated_frame(7200, 2160, 3840)
    overlay_cases.append(("uhd_seed7200", 7200, (2160, 3840), boxes, "low", "medium"))
    edge = [
        {"x": 0, "y": 0, "width": 30, "height": 30, "label": "#10", "confidence": "low"},
        {"x": 70, "y": 60, "width": 30, "height": 40, "label": "#12", "severity": "COSMETIC", "confidence": "low"},
        {"x": 60.0, "y": 70.0, "width": 40.0, "height": 30.0, "label": "#3"},
        {"x": 99.5, "y": 99.5, "width": 0.5, "height": 0.5, "label": "#4"},          # too small -> skipped
        {"x": 10, "y": 10, "width": 95, "height": 20, "label": "#5"},                 # exceeds bounds -> skipped
        {"x": 5, "y": 50, "width": 90, "height": 60, "label": "#6"},                  # exceeds / too large -> skipped
        {"x": -1, "y": 5, "width": 10, "height": 10, "label": "#7"},                  # invalid -> skipped
        {"x": 45.5, "y": 2.2, "width": 3.3, "height": 3.1, "label": "#8", "severity": "CRITICAL", "confidence": "high"},
        {"x": 2, "y": 80, "width": 20, "height": 20, "confidence": "medium"},         # default label from index
    ]
    for shape, tag in (((1080, 1920), "1080p"), ((480, 640), "vga"), ((2160, 3840), "uhd"), ((333, 517), "odd")):
        overlay_cases.append((f"edge_{tag}", 7300, shape, edge, "low", "medium"))
    overlay_cases.append(("edge_1080p_thr_medium", 7300, (1080, 1920), edge, "medium", "medium"))
    overlay_cases.append(("edge_1080p_thr_high_crit_high", 7300, (1080, 1920), edge, "high", "high"))
    overlay_cases.append(("edge_1080p_thr_high", 7300, (1080, 1920), edge, "high", "low"))
    for name, seed, shape, boxes, thr, crit in overlay_cases:
        frame = np.random.default_rng(seed).integers(0, 256, (*shape, 3), dtype=np.uint8)
        res = run_ref(frame, boxes, thr, crit)
        out["overlay"].append({"name": name, "seed": seed, "shape": list(shape), "boxes": boxes,
                               "confidence_threshold": thr, "criticality": crit, "sha256": sha(res),
                               "changed_pixels": int((res != frame).any(2).sum())})
        print("overlay", name, sha(res)[:16], out["overlay"][-1]["changed_pixels"])

    out["overlay"] += overlay_label_goldens()
    out["heatmap"] = heatmap_goldens(arrays)
    arrays["jet_bgr"] = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)[0]
    np.savez_compressed(HERE / "arrays.npz", **arrays)
    out["quality"] = quality_goldens()
    out["compare"] = compare_goldens()
    out["agents"] = agent_goldens()
    (HERE / "goldens.json").write_text(json.dumps(out, indent=1))
    print("wrote", HERE / "goldens.json", HERE / "arrays.npz")


if __name__ == "__main__":
    main()
