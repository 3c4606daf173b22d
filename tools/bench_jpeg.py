"""JPEG files -> pixel_values with the decode on the GPU (nvJPEG batched, Huffman on the GPU) against the reference's
host route (PIL decode + transformers PIL processor): images/s for a batch of 1080p JPEG streams held in host memory.
The decode is the tolerance-specified stage (see tests/test_gpu_jpeg.py); the number is reported beside, not instead of,
bench.py's headline (raw frames).  One JSON line."""
import io
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def smooth_frame(seed, h=1080, w=1920):
    """Natural-image-like content (heavily low-passed noise + gradient): ~250-350 KB at q90 4:2:0, like a photo."""
    import cv2
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 8, w // 8, 3), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    fine = rng.integers(0, 7, (h, w, 3), dtype=np.uint8)
    return np.clip(img.astype(np.int32) + fine - 3, 0, 255).astype(np.uint8)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    eng = get_engine()
    codec = eng.jpeg_codec("gpu_hybrid")
    streams = []
    for i in range(8):
        buf = io.BytesIO()
        Image.fromarray(smooth_frame(100 + i)).save(buf, format="JPEG", quality=90)
        streams.append(buf.getvalue())
    batch = [streams[i % 8] for i in range(n)]

    def gpu_route():
        frames = codec.decode_batch(batch)
        return eng.preprocess(frames)

    for _ in range(2):
        gpu_route()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        pv, grid = gpu_route()
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        codec.decode_batch(batch)
    torch.cuda.synchronize()
    dec_s = (time.perf_counter() - t0) / reps

    cpu = None
    try:
        from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil
        proc = Qwen2VLImageProcessorPil()
        t0 = time.perf_counter()
        k = 8
        for s in batch[:k]:
            proc(images=[Image.open(io.BytesIO(s)).convert("RGB")], return_tensors="np")
        cpu = k / (time.perf_counter() - t0)
    except Exception as e:                                   # transformers absent: report the decode alone
        cpu = None
        print("cpu route unavailable:", e, file=sys.stderr)
    t0 = time.perf_counter()
    for s in batch[:16]:
        np.asarray(Image.open(io.BytesIO(s)).convert("RGB"))
    pil_dec = 16 / (time.perf_counter() - t0)
    print(json.dumps({"workload": f"{n} 1080p JPEG streams (q90 4:2:0, {sum(len(s) for s in streams) // 8} B avg) -> pixel_values",
                      "images_per_s": round(n / gpu_s, 1), "decode_only_images_per_s": round(n / dec_s, 1),
                      "cpu_reference_images_per_s_1thread": None if cpu is None else round(cpu, 2),
                      "pil_decode_images_per_s_1thread": round(pil_dec, 1), "host_cores": os.cpu_count(),
                      "backend": "nvJPEG gpu_hybrid, interpolated chroma", "rows": int(pv.shape[0])}))


if __name__ == "__main__":
    main()
