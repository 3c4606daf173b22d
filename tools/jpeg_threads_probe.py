"""How far does the nvJPEG decode stage scale with host threads?  T threads, each with its own nvJPEG handle (the
engine's per-thread codec) and its own CUDA stream, decode batches of 1080p JPEG streams concurrently; images/s in total.
One JSON line per T."""
import io
import json
import sys
import threading
import time
from pathlib import Path

import torch
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from bench_jpeg import smooth_frame  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def main():
    eng = get_engine()
    streams = []
    for i in range(8):
        buf = io.BytesIO()
        Image.fromarray(smooth_frame(100 + i)).save(buf, format="JPEG", quality=90)
        streams.append(buf.getvalue())
    per_batch, reps = 64, 6
    batch = [streams[i % 8] for i in range(per_batch)]
    for backend in ("gpu_hybrid", "hybrid", "default"):
        for T in (1, 2, 4, 8):
            ready, go = threading.Barrier(T + 1), threading.Barrier(T + 1)
            errs = []

            def work():
                try:
                    st = torch.cuda.Stream()
                    with torch.cuda.stream(st):
                        codec = eng.jpeg_codec(backend)
                        codec.decode_batch(batch, cpu_threads=max(1, 32 // T))
                        st.synchronize()
                        ready.wait()
                        go.wait()
                        for _ in range(reps):
                            codec.decode_batch(batch, cpu_threads=max(1, 32 // T))
                        st.synchronize()
                except Exception as e:                      # noqa: BLE001
                    errs.append(repr(e)[:200])
                    try:
                        ready.abort(); go.abort()
                    except Exception:
                        pass

            th = [threading.Thread(target=work) for _ in range(T)]
            for t in th:
                t.start()
            try:
                ready.wait()
                t0 = time.perf_counter()
                go.wait()
                for t in th:
                    t.join()
                dt = time.perf_counter() - t0
                print(json.dumps({"backend": backend, "threads": T, "images_per_s": round(T * reps * per_batch / dt, 1), "errors": errs}), flush=True)
            except threading.BrokenBarrierError:
                for t in th:
                    t.join()
                print(json.dumps({"backend": backend, "threads": T, "errors": errs}), flush=True)


if __name__ == "__main__":
    main()
