"""Small invocations of every kernel for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py` (GPU box).
Results are also checked against the oracle, so a clean run means: no invalid access AND bit-exact output."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import overlay as OV  # noqa: E402
from oracle import quality as QL  # noqa: E402
from oracle import qwen2vl as Q  # noqa: E402
from vision_inspection_system_b200 import geometry as G  # noqa: E402
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def main():
    eng = get_engine()
    cases = [((270, 480), G.DEFAULT_MAX_PIXELS, "auto"),      # scheduled, 8-slot, downscale? (upscale -> two per index)
             ((1080, 1920), G.DEFAULT_MAX_PIXELS, "auto"),    # scheduled, 8-slot
             ((1080, 1920), 250000, "auto"),                  # scheduled, 16-slot
             ((300, 640), G.DEFAULT_MAX_PIXELS, "general"),   # general warp-specialised kernel
             ((2160, 3840), G.DEFAULT_MAX_PIXELS, "general"), # phase-synchronous 16-tap kernel
             ((100, 502), G.DEFAULT_MAX_PIXELS, "auto")]      # generic passes (unaligned pitch)
    for shape, mp, path in cases:
        f = synth.noise_frame(3, *shape)
        pv, grid = eng.preprocess([torch.from_numpy(f).cuda()], max_pixels=mp, path=path)
        want, wgrid = Q.preprocess([f], max_pixels=mp)
        assert np.array_equal(grid.numpy(), wgrid) and np.array_equal(pv.cpu().numpy(), want), (shape, mp, path)
        print("preprocess ok", shape, mp, path, flush=True)
    f = synth.noise_frame(4, 1080, 1920)
    for out in ((576, 1024), (540, 958)):
        got = eng.resize_u8(torch.from_numpy(f).cuda(), out[0], out[1], Q.LANCZOS)
        assert np.array_equal(got.cpu().numpy(), Q.resize(f, out[0], out[1], Q.LANCZOS)), out
    print("resize ok", flush=True)
    frame, boxes = synth.annotated_frame(7001, 480, 640)
    got = eng.annotate([torch.from_numpy(frame).cuda()], [boxes])[0].cpu().numpy()
    assert np.array_equal(got, OV.draw_bounding_boxes(frame, boxes))
    print("overlay ok", flush=True)
    sums, _ = eng.quality_stats([torch.from_numpy(frame).cuda(), torch.from_numpy(f).cuda()])
    assert tuple(int(v) for v in sums[0].cpu().numpy()) == QL.stats(frame)
    assert tuple(int(v) for v in sums[1].cpu().numpy()) == QL.stats(f)
    print("quality ok", flush=True)


if __name__ == "__main__":
    main()
