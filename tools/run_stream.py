"""BASELINE config 5 on N GPUs: a mixed-resolution inspection stream, image-sharded (greedy balance on algorithmic bytes,
no collective on the data path), two inputs per frame (Inspector: thumbnail 2048 -> Qwen2-VL processor; Auditor: thumbnail
1024 -> processor).  Optionally gathers the Inspector patch rows on rank 0 over NCCL and checks them.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_stream.py [--frames F] [--gather]

Prints one JSON line on rank 0: whole-job frames/s (device resident, CUDA events, max over ranks), per-rank byte balance,
gather bandwidth.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vision_inspection_system_b200 import _native as N  # noqa: E402
from vision_inspection_system_b200 import geometry as G  # noqa: E402
from vision_inspection_system_b200 import sharding as S  # noqa: E402
from vision_inspection_system_b200 import synth  # noqa: E402
from vision_inspection_system_b200.engine import get_engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512, help="frames of the whole job")
    ap.add_argument("--gather", action="store_true")
    ap.add_argument("--gather-frames", type=int, default=512,
                    help="the optional gather collects the Inspector rows of the first G frames of the job on rank 0 "
                         "(the consumer's chunk; all 8192 frames would be 127 GiB on one GPU)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = get_engine()
    shapes = synth.mixed_resolution_shapes(args.frames, seed=9000)
    costs = [S.frame_bytes(h, w) for h, w in shapes]
    mine = S.balanced_shards(costs, world)[rank]
    cache = {}
    frames = []
    for i in mine:                                        # two distinct noise frames per resolution, resident in HBM
        s = shapes[i]
        if s not in cache:
            cache[s] = [torch.from_numpy(synth.noise_frame(9000 + k, *s)).cuda() for k in range(2)]
        frames.append(cache[s][i % 2])

    def step():
        out = {}
        for role, limit in (("inspector", G.INSPECTOR_MAX_SIZE), ("auditor", G.AUDITOR_MAX_SIZE)):
            out[role] = eng.preprocess(eng.agent_inputs(frames, role))
        return out

    for _ in range(2):
        res = step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        res = step()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / 3], device="cuda")
    load = torch.tensor([float(sum(costs[i] for i in mine))], device="cuda")
    loads = [load.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_gather(loads, load)
    line = {"config": "mixed-resolution dual Inspector+Auditor stream", "frames": args.frames, "n_gpus": world,
            "ms_per_pass": float(ms.item()), "frames_per_s": args.frames / float(ms.item()) * 1e3,
            "rank_bytes_min_over_max": float(min(x.item() for x in loads) / max(x.item() for x in loads))}
    if args.gather and world > 1:
        pv, grid = res["inspector"]
        n_gather = min(args.gather_frames, args.frames)
        k_local = sum(1 for i in mine if i < n_gather)               # `mine` is sorted: the chunk is a prefix of it
        rows_local = int((grid[:k_local, 0] * grid[:k_local, 1] * grid[:k_local, 2]).sum())
        pv, grid = pv[:rows_local], grid[:k_local]
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        S.gather_patches(pv[:1], grid[:1], mine[:1], dst=0)           # opens the peer connections (not timed)
        dist.barrier()
        torch.cuda.synchronize()
        g0.record()
        all_pv, all_grid = S.gather_patches(pv, grid, mine[:k_local], dst=0)
        g1.record()
        torch.cuda.synchronize()
        if rank == 0:
            # rank 0 recomputes two remote frames itself and compares the gathered rows bit for bit
            rows = (all_grid[:, 0] * all_grid[:, 1] * all_grid[:, 2]).tolist()
            starts = np.concatenate([[0], np.cumsum(rows)])
            ok = len(rows) == n_gather
            remote = [i for i in range(n_gather) if i not in set(mine)][:2]
            for i in remote:
                f = torch.from_numpy(synth.noise_frame(9000 + i % 2, *shapes[i])).cuda()
                want, _ = eng.preprocess(eng.agent_inputs([f], "inspector"))
                ok = ok and torch.equal(all_pv[starts[i]:starts[i + 1]], want)
            sec = g0.elapsed_time(g1) / 1e3
            recv = (all_pv.shape[0] - pv.shape[0]) * 1176 * 4
            line["gather"] = {"ok": bool(ok), "frames": n_gather, "rows_total": int(all_pv.shape[0]), "bytes_received": int(recv),
                              "seconds": sec, "receiver_GBps": recv / sec / 1e9}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
