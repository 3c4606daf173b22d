"""Probe of the nvJPEG stage on this GPU: which backends exist, how far the decoded pixels are from the reference's
decoders (PIL / cv2 = libjpeg-turbo) with and without interpolated chroma upsampling, encode round trips, throughput of
single and batched decodes.  One JSON object per line."""
import io
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.jpeg import BACKENDS, JpegCodec  # noqa: E402


def main():
    import cv2
    from PIL import Image
    dev = torch.device("cuda", 0)
    frames = {"lowpass_1080p": synth.pattern_frames(1080, 1920)["lowpass"], "noise_vga": synth.noise_frame(2, 480, 640),
              "odd_333x517": synth.pattern_frames(333, 517)["lowpass"]}
    streams = {}
    for name, f in frames.items():
        for tag, sub in (("444", 0), ("420", 2), ("422", 1)):
            buf = io.BytesIO()
            Image.fromarray(f).save(buf, format="JPEG", quality=90, subsampling=sub)
            streams[f"{name}_{tag}"] = buf.getvalue()
        buf = io.BytesIO()
        Image.fromarray(f).save(buf, format="JPEG", quality=85, progressive=True)
        streams[f"{name}_progressive"] = buf.getvalue()
        buf = io.BytesIO()
        Image.fromarray(f).convert("L").save(buf, format="JPEG", quality=90)
        streams[f"{name}_gray"] = buf.getvalue()
    for backend in BACKENDS:
        for interp in (True, False):
            try:
                codec = JpegCodec(dev, backend, interp)
            except Exception as e:
                print(json.dumps({"backend": backend, "interpolate": interp, "create_error": str(e)[:200]}), flush=True)
                continue
            for name, s in streams.items():
                want = np.asarray(Image.open(io.BytesIO(s)).convert("RGB")).astype(np.int32)
                rec = {"backend": backend, "interpolate": interp, "stream": name, "bytes": len(s)}
                try:
                    got = codec.decode(s).cpu().numpy().astype(np.int32)
                    d = np.abs(got - want)
                    rec.update(max_abs=int(d.max()), mean_abs=round(float(d.mean()), 4), frac_gt2=round(float((d > 2).mean()), 5))
                    bgr = codec.decode(s, bgr=True).cpu().numpy()
                    rec["bgr_is_flipped_rgb"] = bool(np.array_equal(bgr[:, :, ::-1], got.astype(np.uint8)))
                except Exception as e:
                    rec["decode_error"] = str(e)[:200]
                try:
                    got = codec.decode_batch([s, s])[1].cpu().numpy().astype(np.int32)
                    d = np.abs(got - want)
                    rec.update(batch_max_abs=int(d.max()), batch_mean_abs=round(float(d.mean()), 4))
                except Exception as e:
                    rec["batch_error"] = str(e)[:200]
                print(json.dumps(rec), flush=True)
            # throughput: 64 x 1080p 4:2:0 q90
            s = streams["lowpass_1080p_420"]
            try:
                for n, fn in (("single", lambda: [codec.decode(s) for _ in range(64)]),
                              ("batch64", lambda: codec.decode_batch([s] * 64))):
                    fn()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    fn()
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    print(json.dumps({"backend": backend, "interpolate": interp, "decode": n, "images_per_s": round(64 / dt, 1)}), flush=True)
            except Exception as e:
                print(json.dumps({"backend": backend, "interpolate": interp, "throughput_error": str(e)[:200]}), flush=True)
            # encode round trip
            try:
                f = frames["lowpass_1080p"]
                dev_f = torch.from_numpy(np.ascontiguousarray(f[:, :, ::-1])).cuda()
                for q, sub in ((95, "4:2:0"), (85, "4:2:0"), (95, "4:4:4")):
                    data = codec.encode(dev_f, q, sub)
                    back = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
                    ok, ref = cv2.imencode(".jpg", dev_f.cpu().numpy(), [cv2.IMWRITE_JPEG_QUALITY, q] + ([cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444] if sub == "4:4:4" else []))
                    refback = cv2.imdecode(ref, cv2.IMREAD_COLOR)
                    src = dev_f.cpu().numpy().astype(np.float64)
                    psnr = lambda a: round(float(10 * np.log10(255 ** 2 / np.mean((a.astype(np.float64) - src) ** 2))), 2)  # noqa: E731
                    print(json.dumps({"backend": backend, "encode_q": q, "sub": sub, "bytes": len(data), "cv2_bytes": int(len(ref)),
                                      "psnr": psnr(back), "cv2_psnr": psnr(refback)}), flush=True)
            except Exception as e:
                print(json.dumps({"backend": backend, "encode_error": str(e)[:200]}), flush=True)
            codec.close()


if __name__ == "__main__":
    main()
