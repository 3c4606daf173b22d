#!/usr/bin/env python
"""Per-launch breakdown of the dual Inspector + Auditor stream slice of bench.py (config 5): which launches the cached pass
consists of, their grids and device times (CUDA events, each launch repeated).  python tools/dual_breakdown.py"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vision_inspection_system_b200 import _native as N, synth          # noqa: E402
from vision_inspection_system_b200.engine import get_engine, _SchedLaunch, _stream_ptr   # noqa: E402


def timed(fn, reps=20):
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


def head_of(sched):
    return np.frombuffer(sched[:N.SCHED_HEAD_DTYPE.itemsize], N.SCHED_HEAD_DTYPE)[0]


def main():
    eng = get_engine()
    shapes = synth.mixed_resolution_shapes(192, seed=9000)
    cache, fr = {}, []
    for i, s in enumerate(shapes):
        if s not in cache:
            cache[s] = [torch.from_numpy(synth.noise_frame(9000 + k, *s)).cuda() for k in range(2)]
        fr.append(cache[s][i % 2])
    res = eng.preprocess_dual(fr)
    out = torch.empty((res["inspector"][0].shape[0] + res["auditor"][0].shape[0], 1176), dtype=torch.float32, device="cuda")
    print(json.dumps({"whole_pass_ms": timed(lambda: eng.preprocess_dual(fr, out=out))}))
    dp = next(iter(eng._dual_plans.values()))
    rows = []
    for rp in dp["thumbs"]:
        h = head_of(rp[0])
        ms = timed(lambda rp=rp: eng._run_resize(rp))
        rows.append({"stage": "thumbnail", "src": [int(h["src_h"]), int(h["src_w"])], "dst": [int(h["dst_h"]), int(h["dst_w"])],
                     "frames": int(rp[4]), "items": int(rp[4]) * int(h["n_strips"]) * int(h["n_segs"]),
                     "kernel": "mma" if h["mma_ks"] else "dp" if h["dp_words"] else f"ring{int(h['ring'])}", "ms": round(ms, 4)})
    plan = dp["plan"]
    sp = _stream_ptr()
    for fl in plan.fused:
        if isinstance(fl, _SchedLaunch):
            h = head_of(fl.sched)
            fn = lambda fl=fl: N.check(eng.L.vis_preprocess_fused_sched_dup(
                fl.sched.ctypes.data_as(C.c_void_p), fl.frames.data_ptr(), fl.n_frames, fl.hrec.data_ptr(), fl.vrec.data_ptr(),
                eng.lut.data_ptr(), out.data_ptr(), fl.dup.data_ptr() if fl.dup is not None else None, sp), "sched")
            rows.append({"stage": "processor", "src": [int(h["src_h"]), int(h["src_w"])], "dst": [int(h["dst_h"]), int(h["dst_w"])],
                         "frames": fl.n_frames, "items": fl.n_items,
                         "kernel": "mma" if h["mma_ks"] else "dp" if h["dp_words"] else f"ring{int(h['ring'])}", "ms": round(timed(fn), 4)})
        else:
            rows.append({"stage": "processor", "kernel": "general", "frames": fl.n_frames})
    rows.append({"generic_frames": len(plan.generic)})
    for r in rows:
        print(json.dumps(r))
    print(json.dumps({"sum_of_launches_ms": round(sum(r.get("ms", 0) for r in rows), 4)}))


if __name__ == "__main__":
    main()
