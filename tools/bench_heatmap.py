"""GPU benchmark of the heat-map overlay: 1080p frames with K ~ U{1..6} defects, device resident, ONE batch call (six
launches) and, beside it, one call per frame (the reference API is per image).  Reports ms per frame, kernels launched, and the oracle port on the CPU beside it."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import heatmap as H  # noqa: E402
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    eng = get_engine()
    items = []
    for i in range(n):
        rng = np.random.default_rng(8100 + i)
        frame = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
        items.append((torch.from_numpy(frame).cuda(), synth.random_defects(rng, int(rng.integers(1, 7))), frame))
    frames, defects = [f for f, _, _ in items], [d for _, d, _ in items]
    eng.heatmap_batch(frames[:8], defects[:8])
    shapes = [(1080, 1920)] * n
    t0 = time.perf_counter()
    plan = eng.plan_heatmap(shapes, defects)             # host half: the reference's per-defect scalar code + tables
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(3):
        eng.heatmap_batch(frames, plan=plan)             # first full calls: the scratch planes (GBs) are allocated
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        eng.heatmap_batch(frames, plan=plan)             # ONE batch call: six launches for all frames
    launches = eng.last_launches
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    t0 = time.perf_counter()
    eng.heatmap_batch(frames, defects)                   # the public call, host planning included
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    # the same frames one call each (the reference API is per image)
    a.record()
    for f, d in zip(frames, defects):
        eng.heatmap(f, d)
    b.record()
    torch.cuda.synchronize()
    ms_single = a.elapsed_time(b)
    cpu_ms = None
    try:
        from oracle import heatmap as OH
        t0 = time.perf_counter()
        for _, d, frame in items[:3]:
            OH.create_heatmap_overlay(frame, d, H.JET_BGR)
        cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    except Exception:
        pass
    print(json.dumps({"workload": f"{n} 1080p BGR frames, {sum(len(d) for _, d, _ in items)} defects", "gpu_ms_per_frame": ms / n,
                      "wall_ms_per_frame_incl_host_params": wall / n * 1e3, "host_plan_ms_per_frame": plan_ms / n, "images_per_s": n / ms * 1e3,
                      "launches_per_batch": launches, "per_image_calls_ms_per_frame": ms_single / n,
                      "per_image_calls_images_per_s": n / ms_single * 1e3, "cpu_oracle_port_ms_per_frame": cpu_ms}))


if __name__ == "__main__":
    main()
