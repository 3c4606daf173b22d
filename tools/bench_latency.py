"""Single-call latencies of the drop-in entry points (what the reference's one-image-at-a-time callers see): wall time per
call, synchronised, for a fresh host image each time.  One JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import image_utils as IU  # noqa: E402
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    eng = get_engine()
    frames = [synth.noise_frame(10 + i, 1080, 1920) for i in range(4)]
    pil = [Image.fromarray(f) for f in frames]
    big = Image.fromarray(synth.noise_frame(20, 2160, 3840))
    _, boxes = synth.annotated_frame(7000)
    k = [0]

    def nxt(seq):
        k[0] += 1
        return seq[k[0] % len(seq)]

    out = {
        "preprocess_for_vlm_1080p_pil_ms": timed(lambda: IU.preprocess_for_vlm(nxt(pil))[0][0, 0].item()),
        "preprocess_for_vlm_1080p_auditor_ms": timed(lambda: IU.preprocess_for_vlm(nxt(pil), role="auditor")[0][0, 0].item()),
        "engine_preprocess_1080p_device_frame_ms": timed(lambda: eng.preprocess([torch.from_numpy(nxt(frames)).cuda()])),
        "resize_image_4k_to_2048_ms": timed(lambda: IU.resize_image(big, 2048), reps=10),
        "annotate_1080p_inplace_device_frame_ms": timed(
            lambda: eng.annotate([torch.from_numpy(nxt(frames)).cuda()], [boxes], inplace=True)),
    }
    try:
        from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil
        proc = Qwen2VLImageProcessorPil()
        t0 = time.perf_counter()
        for i in range(3):
            proc(images=[pil[i]], return_tensors="np")
        out["cpu_reference_preprocess_1080p_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    except Exception:
        pass
    print(json.dumps({k_: round(v, 3) for k_, v in out.items()}))


if __name__ == "__main__":
    main()
