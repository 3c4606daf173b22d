"""Device-resident timings of every BASELINE.json config besides the headline one (which bench.py measures):
hub max_pixels, 4K frames, agent thumbnails, the dual Inspector+Auditor stream on mixed resolutions.
One JSON object per line: ms, images/s, algorithmic bytes, fraction of the measured HBM copy peak, kernels launched.

    python tools/bench_configs.py [name ...]
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import _native as N  # noqa: E402
from vision_inspection_system_b200 import geometry as G  # noqa: E402
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402

PEAK = 6539.9
p = ROOT / "MEASURED_PEAKS.json"
if p.exists():
    PEAK = float(json.loads(p.read_text())["hbm_gbs"])


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, n, ms, nbytes, launches, note=""):
    print(json.dumps({"config": name, "frames": n, "ms": round(ms, 4), "images_per_s": round(n / ms * 1e3, 1),
                      "algorithmic_bytes": int(nbytes), "hbm_frac": round(nbytes / ms / 1e6 / PEAK, 4),
                      "launches": launches, "note": note}), flush=True)


def out_bytes(h, w, max_pixels):
    dh, dw = G.smart_resize(h, w, G.FACTOR, G.DEFAULT_MIN_PIXELS, max_pixels)
    return (dh // 14) * (dw // 14) * 1176 * 4, (dh, dw)


def uniform(eng, name, shape, n, max_pixels, distinct=8):
    h, w = shape
    base = torch.from_numpy(np.stack([synth.noise_frame(4000 + i, h, w) for i in range(distinct)])).cuda()
    frames = base.repeat(n // distinct, 1, 1, 1).contiguous()
    ob, dst = out_bytes(h, w, max_pixels)
    out = torch.empty((n * ob // 4704, 1176), dtype=torch.float32, device="cuda")
    ms = timed(lambda: eng.preprocess(frames, max_pixels=max_pixels, out=out))
    plan = eng.plan_batch(frames, max_pixels=max_pixels)
    kinds = ",".join(type(f).__name__ for f in plan.fused) + ("+generic" if plan.generic else "")
    report(name, n, ms, n * (h * w * 3 + ob), eng.last_launches, f"{h}x{w} -> {dst[0]}x{dst[1]} [{kinds}]")


def thumbs(eng, name, shape, limit, n):
    h, w = shape
    tw, th = G.thumbnail_size(w, h, limit)
    frames = [torch.from_numpy(synth.noise_frame(4000 + i % 4, h, w)).cuda() for i in range(n)]
    ms = timed(lambda: eng.resize_batch_u8(frames, th, tw, N.FILTER_LANCZOS), reps=5)
    launches = eng.last_launches
    nbytes = n * (h * w * 3 + th * tw * 3)          # frame read + thumbnail written (intermediates not credited)
    report(name, n, ms, nbytes, launches, f"{h}x{w} -> {th}x{tw} LANCZOS uint8 "
           + ("[fused scheduled kernel]" if launches == 1 else "[two generic passes per frame]"))


def dual_stream(eng, name, n):
    from PIL import Image  # noqa: F401
    shapes = synth.mixed_resolution_shapes(n, seed=9000)
    cache = {}
    frames = []
    for i, s in enumerate(shapes):
        if s not in cache:
            cache[s] = [torch.from_numpy(synth.noise_frame(9000 + k, *s)).cuda() for k in range(2)]
        frames.append(cache[s][i % 2])
    nbytes = 0
    for (h, w) in shapes:
        nbytes += h * w * 3
        for limit in (G.INSPECTOR_MAX_SIZE, G.AUDITOR_MAX_SIZE):
            hh, ww = h, w
            if max(h, w) > limit:
                ww, hh = G.thumbnail_size(w, h, limit)
            nbytes += out_bytes(hh, ww, G.DEFAULT_MAX_PIXELS)[0]
    launches = [0]

    def run():
        eng.preprocess_dual(frames)                     # cached plan: thumbnails per source geometry, one plan over both roles
        launches[0] = eng.last_launches
    ms = timed(run, reps=3, warm=2)
    report(name, n, ms, nbytes, launches[0], "mixed resolutions, Inspector (2048) + Auditor (1024) inputs per frame; "
           "bytes = frame once + both pixel_values (thumbnails not credited)")


def main():
    eng = get_engine()
    want = set(sys.argv[1:])
    cases = {
        "1080p_default": lambda: uniform(eng, "1080p_default", (1080, 1920), 256, G.DEFAULT_MAX_PIXELS),
        "1080p_hub": lambda: uniform(eng, "1080p_hub", (1080, 1920), 96, G.HUB_MAX_PIXELS),
        "4k_default": lambda: uniform(eng, "4k_default", (2160, 3840), 64, G.DEFAULT_MAX_PIXELS),
        "4k_hub": lambda: uniform(eng, "4k_hub", (2160, 3840), 24, G.HUB_MAX_PIXELS),
        "720p_default": lambda: uniform(eng, "720p_default", (720, 1280), 256, G.DEFAULT_MAX_PIXELS),
        "thumb_4k_2048": lambda: thumbs(eng, "thumb_4k_2048", (2160, 3840), 2048, 32),
        "thumb_4k_1024": lambda: thumbs(eng, "thumb_4k_1024", (2160, 3840), 1024, 32),
        "thumb_1080p_1024": lambda: thumbs(eng, "thumb_1080p_1024", (1080, 1920), 1024, 64),
        "dual_mixed": lambda: dual_stream(eng, "dual_mixed", 192),
    }
    for name, fn in cases.items():
        if not want or name in want:
            fn()
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
