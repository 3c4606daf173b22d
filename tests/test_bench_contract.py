"""bench.py's contract with the driver, as far as a box without a GPU can check it: the reference arm (the reference's own
CPU implementation of the path, transformers' Qwen2VLImageProcessorPil on forked workers, or the oracle port) prints ONE
JSON line with the same metric / unit / config as our arm, and our arm refuses to run without CUDA instead of falling
back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def run(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-images", "8")
    assert r.returncode == 0, r.stderr[-800:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout, nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/s 1080p->Qwen2-VL pixel_values" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["frames_per_step"] == 8
    # the SAME config object as our arm prints (the driver compares them): no arm-specific keys inside it
    cfg = d["config"]
    assert set(cfg) == {"workload", "frames_per_gpu_per_step", "frame", "pixel_values_rows_per_frame", "parallelism", "l2"}
    assert cfg["frame"] == [1080, 1920, 3] and cfg["pixel_values_rows_per_frame"] == 4888 and cfg["frames_per_gpu_per_step"] == 256


def test_our_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the arm runs (covered by the driver)")
    r = run("--no-cpu-baseline", "--steps", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]       # no JSON line pretending to be a result
