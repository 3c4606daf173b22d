"""A/B the fused kernel across library builds: for every .so given (VIS_B200_LIB), run the parity check on one
1080p frame + the 256-frame timing in a subprocess.   python tools/ab_fused.py variants/a.so variants/b.so ..."""
import os
import subprocess
import sys

CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from vision_inspection_system_b200 import synth
from vision_inspection_system_b200.engine import get_engine
from oracle import qwen2vl as Q
eng = get_engine()
f = synth.noise_frame(1234, 1080, 1920)
pv, _ = eng.preprocess([torch.from_numpy(f).cuda()])
want, _ = Q.preprocess([f])
ok = np.array_equal(pv.cpu().numpy(), want)
n = 256
base = torch.from_numpy(synth.frames_1080p(16)).cuda()
frames = base.repeat(n // 16, 1, 1, 1).contiguous()
out = torch.empty((n * 4888, 1176), dtype=torch.float32, device="cuda")
for _ in range(3):
    eng.preprocess(frames, out=out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(20):
    eng.preprocess(frames, out=out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"exact={ok}  {ms:.3f} ms/batch  {n / ms * 1e3:.0f} img/s  {n * 29213952 / ms / 1e6 / 6539.9:.3f} of HBM peak")
'''

for lib in sys.argv[1:]:
    env = dict(os.environ, VIS_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(f"{lib}: {r.stdout.strip() or r.stderr.strip()[-400:]}", flush=True)
