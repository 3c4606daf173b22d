import sys, time, cProfile, pstats, io
sys.path.insert(0, ".")
import numpy as np, torch
from vision_inspection_system_b200 import synth
from vision_inspection_system_b200.engine import get_engine
eng = get_engine()
items = [synth.annotated_frame(7000 + i) for i in range(64)]
boxes = [items[i % 64][1] for i in range(1024)]
shapes = [(1080, 1920)] * 1024
for _ in range(2):
    eng.plan_overlay(shapes, boxes); torch.cuda.synchronize()
t = time.perf_counter(); eng.plan_overlay(shapes, boxes); torch.cuda.synchronize(); print("plan ms", (time.perf_counter() - t) * 1e3)
pr = cProfile.Profile(); pr.enable(); eng.plan_overlay(shapes, boxes); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
