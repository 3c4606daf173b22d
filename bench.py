#!/usr/bin/env python
"""bench.py — images/s for 1080p -> Qwen2-VL pixel_values (+ % of the HBM roofline) at N GPUs of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of 256 synthetic 1920x1080 RGB frames per GPU
(BASELINE.json configs[1]); batches shard by image across ranks with no collective (weak scaling).
Prints ONE JSON line on rank 0:
  value        whole-job images/s with the frames already resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the public host-buffer API: pinned host frames -> device pixel_values,
               H2D copies inside the timed region, plus a small result read-back (see DESIGN.md "Measurement")
  roofline     algorithmic HBM bytes of the fused kernel / its measured duration vs the measured copy peak
  cpu_baseline the reference CPU path (transformers Qwen2VLImageProcessorPil) timed on this box's host cores
`--impl reference` times only that CPU path, with every host core, on the same workload definition.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BATCH = 256
H, W = 1080, 1920
DST_H, DST_W = 728, 1316                       # smart_resize(1080, 1920) at the processor's default max_pixels
ROWS = (DST_H // 14) * (DST_W // 14)           # 4888 patch rows per frame
BYTES_PER_IMAGE = H * W * 3 + ROWS * 1176 * 4  # algorithmic HBM bytes: uint8 read + fp32 write = 29 213 952
METRIC = "images/s 1080p->Qwen2-VL pixel_values"
WORKLOAD = "256 synthetic 1920x1080 RGB frames per GPU -> Qwen2-VL pixel_values (min_pixels 3136, max_pixels 1003520)"


# ------------------------------------------------------------------------------------------ CPU reference arm
_CPU_FRAMES = None


def _cpu_worker(idx_range):
    os.environ["OMP_NUM_THREADS"] = "1"
    kind = _CPU_KIND
    lo, hi = idx_range
    n = 0
    if kind == "reference":
        from PIL import Image
        from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil
        proc = Qwen2VLImageProcessorPil()
        for i in range(lo, hi):
            r = proc(images=[Image.fromarray(_CPU_FRAMES[i % len(_CPU_FRAMES)])], return_tensors="np")
            n += int(r["pixel_values"].shape[0] == ROWS)
    else:
        from oracle import qwen2vl as Q
        for i in range(lo, hi):
            pv, _ = Q.preprocess([_CPU_FRAMES[i % len(_CPU_FRAMES)]])
            n += int(pv.shape[0] == ROWS)
    return n


_CPU_KIND = "reference"


class CpuArm:
    """The reference's CPU implementation of the path on `cores` forked workers (one thread each).

    kind "reference": the installed transformers Qwen2VLImageProcessorPil (the processor the Inspector/Auditor
    inputs go through server-side; Pillow does the resampling) — kind "port": the plain-C oracle restatement.
    The pool is created once (imports + processor construction happen in the warm-up), then `run(n)` times n frames.
    """

    def __init__(self, cores: int):
        global _CPU_FRAMES, _CPU_KIND
        import multiprocessing as mp
        from vision_inspection_system_b200 import synth
        try:
            import transformers  # noqa: F401
            from PIL import Image  # noqa: F401
            _CPU_KIND = "reference"
        except Exception:
            _CPU_KIND = "port"
        self.kind = _CPU_KIND
        self.cores = max(1, cores)
        _CPU_FRAMES = [synth.noise_frame(1234 + i, H, W) for i in range(32)]
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.pool.map(_cpu_worker, [(0, 1)] * self.cores)               # warm-up: imports, processor construction

    def run(self, n_images: int) -> float:
        """seconds of wall clock for n_images frames spread over the workers"""
        k = min(self.cores, n_images)
        bounds = [(n_images * i // k, n_images * (i + 1) // k) for i in range(k)]
        t0 = time.perf_counter()
        done = sum(self.pool.map(_cpu_worker, bounds))
        dt = time.perf_counter() - t0
        assert done == n_images
        return dt

    def calibrated(self, target_s: float):
        """(images/s, n, seconds): a sample sized from a short probe so that it takes about target_s seconds"""
        probe_n = 2 * self.cores
        rate = probe_n / self.run(probe_n)
        n = int(max(probe_n, min(rate * target_s, 20000)))
        dt = self.run(n)
        return n / dt, n, dt

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_model() -> str:
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ main
def emit(line: dict) -> None:
    """The ONE JSON line, on the real stdout (libraries such as NCCL print banners to fd 1; see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def bind_to_gpu_numa_node(local_rank: int):
    """One process per GPU on a two-socket box: run this rank (and first-touch its pinned staging buffers) on the CPUs
    NVML names as local to its GPU, so that every rank's host->device copies leave through its own socket instead of
    all of them crossing to the socket the launcher happened to start on.  Only the `e2e` leg moves host data; the
    device-resident `value` is unaffected.  Returns the CPU count bound to, or None when NVML cannot say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    global _REAL_STDOUT
    # keep stdout clean for the driver: everything any library prints to fd 1 goes to stderr instead
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=BATCH, help="frames per GPU per step (default 256 = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-images", type=int, default=0, help="size of the bounded CPU sample (0 = calibrated to ~15 s)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = world if world > 1 else 1
    cores = os.cpu_count() or 1
    config = {"workload": WORKLOAD, "frames_per_gpu_per_step": args.batch, "frame": [H, W, 3],
              "pixel_values_rows_per_frame": ROWS, "parallelism": f"image-sharded x{n_gpus}, no collective",
              "l2": "inputs (1.6 GB) and outputs (5.9 GB) per step exceed the 126 MB L2"}

    # ---------------- reference arm: CPU only, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        arm = CpuArm(cores)
        steps = max(1, args.steps)
        budget_s = 150.0                               # the whole run stays within a few minutes
        probe_n = 2 * cores
        rate = probe_n / arm.run(probe_n)
        n = args.cpu_images or int(max(cores, min(rate * budget_s / (steps + max(args.warmup, 0)), 4096)))
        for _ in range(max(args.warmup, 0)):
            arm.run(n)
        t = [arm.run(n) for _ in range(steps)]
        arm.close()
        total = float(sum(t))
        v = n * steps / total
        line = {"metric": METRIC, "value": v, "unit": "images/s", "n_gpus": n_gpus, "steps": steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8->int32 fixed point->f32", "data": "synthetic", "impl": "reference",
                "config": dict(config, frames_per_step=n),
                "cpu_baseline": {"value": v, "unit": "images/s", "cores": arm.cores, "kind": arm.kind,
                                 "sample": f"{n} seeded 1080p noise frames per step x {steps} steps over {arm.cores} "
                                           f"forked workers (one thread each), {total:.1f} s, on {cpu_model()}"},
                "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ---------------- CPU baseline (rank 0, N=1) BEFORE CUDA is initialised (fork safety) ----------------
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        arm = CpuArm(cores)
        if args.cpu_images:
            n = args.cpu_images
            dt = arm.run(n)
            v = n / dt
        else:
            v, n, dt = arm.calibrated(target_s=15.0)   # ~15 s of wall clock on every core: a bounded sample
        arm.close()
        single = CpuArm(1)                             # one process, one thread
        t1 = single.run(8)
        single.close()
        cpu_baseline = {"value": v, "unit": "images/s", "cores": arm.cores, "kind": arm.kind,
                        "single_thread_value": 8 / t1,
                        "sample": f"{n} seeded 1080p noise frames over {arm.cores} forked workers (one thread each), "
                                  f"{dt:.1f} s, on {cpu_model()} ({cores} logical cores)"}

    import torch
    from vision_inspection_system_b200 import synth
    from vision_inspection_system_b200.engine import get_engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = get_engine()

    # ---------------- inputs: seeded noise, distinct per rank, resident in HBM ----------------
    distinct = min(args.batch, 32)
    base = synth.frames_1080p(distinct, first_seed=1234 + 100000 * rank)
    host = torch.from_numpy(base).repeat((args.batch + distinct - 1) // distinct, 1, 1, 1)[:args.batch].contiguous()
    host_pinned = host.pin_memory()
    frames = host_pinned.cuda(non_blocking=True)
    out = torch.empty((args.batch * ROWS, 1176), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for _ in range(steps):
            fn()
        end.record()
        barrier()
        ms = start.elapsed_time(end)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---------------- device-resident throughput ----------------
    def step_device():
        eng.preprocess(frames, out=out)

    for _ in range(args.warmup):
        step_device()
    launches_per_step = eng.last_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, args.steps)
    value = n_gpus * args.batch * args.steps / (ms / 1e3)

    # ---------------- end to end: pinned host frames -> device pixel_values (+ small read-back) ----------------
    probe = torch.empty((args.batch, 1176), dtype=torch.float32).pin_memory()

    def step_e2e():
        pv, _grid = eng.preprocess_host(host_pinned, out=out)
        probe.copy_(pv.view(args.batch, ROWS, 1176)[:, 0], non_blocking=True)     # first patch row of every frame

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    e2e_launches = eng.last_launches
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e_value = n_gpus * args.batch * e2e_steps / (ms_e2e / 1e3)

    # strict variant: the whole pixel_values tensor copied back to pinned host memory as well (reported separately)
    full_value = None
    try:
        if world > 1:                      # N x 5.9 GB of pinned host memory: measured at N=1 only
            raise RuntimeError("skipped")
        host_out = torch.empty((args.batch * ROWS, 1176), dtype=torch.float32).pin_memory()

        def step_full():
            pv, _grid = eng.preprocess_host(host_pinned, out=out)
            host_out.copy_(pv, non_blocking=True)

        step_full()
        ms_full = timed(step_full, 2)
        full_value = n_gpus * args.batch * 2 / (ms_full / 1e3)
    except Exception:
        pass
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
    kernel_s = (ms / 1e3) / args.steps                       # one fused launch per step on this rank
    achieved = args.batch * BYTES_PER_IMAGE / kernel_s / 1e9
    traffic, traffic_src = None, None                        # measured DRAM bytes per launch, from the committed ncu capture
    try:
        rec = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(f"k_fused_sched@{args.batch}x1080p")
        if rec:
            traffic, traffic_src = rec["dram_bytes_read"] + rec["dram_bytes_write"], rec["source"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8->int32 fixed point->f32", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": args.batch * H * W * 3,
                "d2h_bytes_per_step": args.batch * 1176 * 4,
                "cpus_bound_per_rank": numa,
                "note": "pinned host frames -> device pixel_values (consumer is on the GPU); first patch row of "
                        "every frame read back"},
        "e2e_full_readback": {"value": full_value, "unit": "images/s",
                              "d2h_bytes_per_step": args.batch * ROWS * 1176 * 4},
        "gpu_launches": launches_per_step * args.steps + e2e_launches * e2e_steps,
        "roofline": {"bound": "hbm", "kernel": "k_fused_sched (vis_preprocess_fused_sched)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                     "traffic_source": traffic_src, "algorithmic_bytes_per_launch": args.batch * BYTES_PER_IMAGE,
                     "peak_source": peak_src, "bytes_per_image": BYTES_PER_IMAGE,
                     "launches_timed": launches_per_step * args.steps},
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
