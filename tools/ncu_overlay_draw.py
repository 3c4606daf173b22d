"""One in-place overlay draw of BASELINE config 4 (1024 annotated 1080p frames) between cudaProfilerStart / Stop:

    ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_overlay_tiles -c 1 \
        -f -o gpurun_out/overlay_draw python tools/ncu_overlay_draw.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402

eng = get_engine()
n, distinct = 1024, 64
items = [synth.annotated_frame(7000 + i) for i in range(distinct)]
frames = torch.from_numpy(np.stack([f for f, _ in items])).cuda().repeat(n // distinct, 1, 1, 1).contiguous()
boxes = [items[i % distinct][1] for i in range(n)]
plan = eng.plan_overlay([(1080, 1920)] * n, boxes)
work = frames.clone()
eng.annotate(work, boxes, plan=plan, inplace=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.annotate(work, boxes, plan=plan, inplace=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
