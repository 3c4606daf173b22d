"""GPU micro-benchmark of the fused kernel: ms per 256x1080p batch for several row-split factors."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    eng = get_engine()
    base = torch.from_numpy(synth.frames_1080p(16)).cuda()
    frames = base.repeat(n // 16, 1, 1, 1).contiguous()
    out = torch.empty((n * 4888, 1176), dtype=torch.float32, device="cuda")
    for vs in (None, 1, 2, 3, 4, 6):
        for _ in range(3):
            eng.preprocess(frames, out=out, vsplit=vs)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            eng.preprocess(frames, out=out, vsplit=vs)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        strips = eng.plan_batch(frames, vsplit=vs).fused[0].n_strips
        print(f"vsplit={vs}: {ms:.3f} ms/batch  {n / ms * 1e3:.0f} img/s  {n * 29213952 / ms / 1e6 / 6539.9:.3f} of HBM peak  ({strips} CTAs)")


if __name__ == "__main__":
    main()
