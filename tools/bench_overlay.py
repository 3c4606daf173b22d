"""GPU benchmark of the overlay rasteriser on BASELINE config 4 (1024 annotated 1080p frames, K ~ U{1..8} boxes).

Prints one JSON line: device-resident images/s out of place and in place, fraction of the measured HBM copy peak
(2 * H * W * 3 algorithmic bytes per image out of place), host expansion cost, and the reference's own
draw-only CPU time (cv2 through the oracle's cv2 facade, in-memory arrays) beside it.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    distinct = min(n, 64)
    eng = get_engine()
    items = [synth.annotated_frame(7000 + i) for i in range(distinct)]
    frames = torch.from_numpy(np.stack([f for f, _ in items])).cuda().repeat(n // distinct, 1, 1, 1).contiguous()
    boxes = [items[i % distinct][1] for i in range(n)]
    shapes = [(1080, 1920)] * n
    t0 = time.perf_counter()
    plan = eng.plan_overlay(shapes, boxes)
    t_plan_first = time.perf_counter() - t0              # includes the one-time pinned-buffer allocation
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = eng.plan_overlay(shapes, boxes)
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0                    # steady state: box rules, expansion, tile binning, upload
    n_leaves = plan[0].numel() // 48
    out = eng.annotate(frames, boxes, plan=plan)
    torch.cuda.synchronize()

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms_oop = timed(lambda: eng.annotate(frames, boxes, plan=plan))
    work = frames.clone()
    ms_inp = timed(lambda: eng.annotate(work, boxes, plan=plan, inplace=True))
    peak = 6539.9
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    bytes_img = 2 * 1080 * 1920 * 3
    # CPU reference beside it: the same cv2 calls on in-memory arrays (draw only, one thread)
    cpu_ms = None
    try:
        from oracle import overlay as OV
        t0 = time.perf_counter()
        for f, b in items[:32]:
            OV.render_cv2(f, OV.select_boxes(b, 1920, 1080))
        cpu_ms = (time.perf_counter() - t0) / 32 * 1e3
    except Exception:
        pass
    print(json.dumps({
        "workload": f"{n} annotated 1080p BGR frames, {sum(len(b) for b in boxes)} boxes, {n_leaves} leaves",
        "out_of_place": {"ms": ms_oop, "images_per_s": n / ms_oop * 1e3, "hbm_frac": n * bytes_img / ms_oop / 1e6 / peak},
        "in_place": {"ms": ms_inp, "images_per_s": n / ms_inp * 1e3},
        "host_plan_ms_per_frame": t_plan / n * 1e3, "host_plan_first_call_ms_per_frame": t_plan_first / n * 1e3, "leaf_bytes_per_frame": n_leaves * 48 / n,
        "cpu_cv2_draw_only_ms_per_frame": cpu_ms, "peak_gbs": peak}))
    assert torch.equal(out[:distinct], out[distinct:2 * distinct]) if n >= 2 * distinct else True


if __name__ == "__main__":
    main()
