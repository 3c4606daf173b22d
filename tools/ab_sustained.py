"""A/B library builds of the hot kernel under SUSTAINED load: for every .so given (VIS_B200_LIB), in a subprocess:
parity of one 1080p frame against the oracle, the 20-launch burst, then the same launch back to back for 3 s with the
SM clock / power / throttle reasons sampled during it (the kernel draws ~1 kW: its sustained speed is set by the power
cap, so instructions and shared-memory wavefronts per image matter beyond what the burst number shows).

    python tools/ab_sustained.py variants/libvis_a.so variants/libvis_b.so ...
"""
import json
import os
import subprocess
import sys

CODE = r'''
import json, sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
from vision_inspection_system_b200 import synth
from vision_inspection_system_b200.engine import get_engine
from oracle import qwen2vl as Q
eng = get_engine()
f = synth.noise_frame(1234, 1080, 1920)
pv, _ = eng.preprocess([torch.from_numpy(f).cuda()])
want, _ = Q.preprocess([f])
ok = bool(np.array_equal(pv.cpu().numpy(), want))
n = 256
base = torch.from_numpy(synth.frames_1080p(16)).cuda()
frames = base.repeat(n // 16, 1, 1, 1).contiguous()
out = torch.empty((n * 4888, 1176), dtype=torch.float32, device="cuda")
sampler = bench.ClockSampler(0); sampler.start()
def timed(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.time(); a.record()
    for _ in range(reps):
        eng.preprocess(frames, out=out)
    b.record(); torch.cuda.synchronize(); t1 = time.time()
    return a.elapsed_time(b) / reps, (t0, t1)
for _ in range(5):
    eng.preprocess(frames, out=out)
time.sleep(1.0)                                        # start every variant from an idle, cool-ish chip
burst, _ = timed(20)
time.sleep(1.0)
sus, win = timed(int(3000 / burst))
ck = sampler.window(*win); sampler.stop()
frac = lambda ms: n * 29213952 / ms / 1e6 / 6539.9
print(json.dumps({"exact": ok, "burst_ms": round(burst, 4), "burst_frac": round(frac(burst), 4), "sustained_ms": round(sus, 4),
                  "sustained_frac": round(frac(sus), 4), "sm_mhz": ck["sm_mhz"], "power_w_max": ck["power_w_max"], "reasons": ck["reasons"]}))
'''

for lib in sys.argv[1:]:
    env = dict(os.environ, VIS_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"error": r.stderr.strip()[-400:]})
    print(json.dumps({"lib": os.path.basename(lib), **json.loads(line)}), flush=True)
