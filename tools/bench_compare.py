"""GPU benchmark of the comparison panel (vis_compose_panels + the header draw list): two device-resident 1080p frames ->
the [840, 2854, 3] canvas of create_side_by_side_comparison; ms per panel, fraction of the measured HBM copy peak
(2 * H*W*3 read + canvas written) and the reference's cv2 calls (one thread, in memory) beside it.  Also the compose
kernel alone (CUDA events around the C-ABI call with tables already resident is what ``ms_kernel_only`` approximates by
subtracting nothing: the host work is reported separately as ``ms_wall``)."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    eng = get_engine()
    a = torch.from_numpy(synth.noise_frame(7500, 1080, 1920)).cuda()
    b = torch.from_numpy(synth.noise_frame(7501, 1080, 1920)).cuda()
    for _ in range(3):
        out = eng.side_by_side(a, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        out = eng.side_by_side(a, b)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) / iters * 1e3
    ms = e0.elapsed_time(e1) / iters
    peak = 6539.9
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    nbytes = 2 * 1080 * 1920 * 3 + out.numel()
    cpu_ms = None
    try:
        import cv2
        cv2.setNumThreads(1)
        from oracle import compare as OC
        fa, fb = a.cpu().numpy(), b.cpu().numpy()
        t0 = time.perf_counter()
        for _ in range(5):
            OC.side_by_side_cv2(fa, fb)
        cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    except Exception:
        pass
    # a report run: one panel per inspected image, composed as ONE batch (two launches)
    n = 256
    from vision_inspection_system_b200 import synth as S
    base = torch.from_numpy(S.frames_1080p(16)).cuda()
    originals = base.repeat(n // 16, 1, 1, 1).contiguous()                   # [256, 1080, 1920, 3]
    annotated = originals.roll(5, 0).contiguous()
    for _ in range(2):
        outs = eng.side_by_side_batch(originals, annotated)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    reps = 5
    for _ in range(reps):
        outs = eng.side_by_side_batch(originals, annotated)
    e1.record()
    torch.cuda.synchronize()
    batch_wall = (time.perf_counter() - t0) / reps * 1e3
    batch_ms = e0.elapsed_time(e1) / reps
    batch = {"pairs": n, "ms_device": batch_ms, "ms_wall": batch_wall, "panels_per_s": n / batch_ms * 1e3,
             "hbm_frac_device_time": n * nbytes / batch_ms / 1e6 / peak, "launches": eng.last_launches,
             "equals_single": bool(torch.equal(outs[0], eng.side_by_side(originals[0], annotated[0])))}
    print(json.dumps({"workload": "side-by-side panel of two 1080p frames", "ms_device": ms, "ms_wall": wall_ms, "batch": batch,
                      "panels_per_s": 1e3 / ms, "bytes": nbytes, "hbm_frac_device_time": nbytes / ms / 1e6 / peak,
                      "cpu_cv2_ms_per_panel_1thread": cpu_ms, "peak_gbs": peak, "launches": eng.last_launches}))


if __name__ == "__main__":
    main()
