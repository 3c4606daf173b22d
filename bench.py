#!/usr/bin/env python
"""bench.py — images/s for 1080p -> Qwen2-VL pixel_values (+ % of the HBM roofline) at N GPUs of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs all|headline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of 256 synthetic 1920x1080 RGB frames per GPU
(BASELINE.json configs[1]); batches shard by image across ranks with no collective (weak scaling).
Prints ONE JSON line on rank 0:
  value         whole-job images/s with the frames already resident in HBM (CUDA events, max over ranks): the headline
  sustained     the same launch repeated back to back for >= 3 s, with the median SM clock seen during it
  e2e           same metric through the public host-buffer API: pinned host frames -> device pixel_values, H2D copies
                inside the timed region, plus a small result read-back (the consumer is the vision tower on the GPU)
  e2e_full_readback   ... with the WHOLE pixel_values tensor copied back to pinned host memory as well (full duplex)
  e2e_jpeg      ... with the frames crossing PCIe as JPEG streams and decoded on the GPU (nvJPEG)
  configs       every other BASELINE.json config (4K at both max_pixels, overlay, mixed dual stream, thumbnails), each
                with its own algorithmic bytes, roofline fraction and clock sample
  roofline      algorithmic HBM bytes of the fused kernel / its measured duration vs the measured copy peak
  cpu_baseline  the reference CPU path (transformers Qwen2VLImageProcessorPil) timed on this box's host cores
`--impl reference` times only that CPU path, with every host core, on the same workload definition.
Every device timing uses a CACHED batch plan (host planning happens in the warm-up, as in a streaming loop that
refills the same staging buffers): `plan_cached` says so in the line.
"""
from __future__ import annotations

import argparse
import io
import json
import math
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
INT_MAC_LANES_PER_CLK_PER_SM = 62.0            # measured IMAD / IDP.4A issue rate (profiles/r02_ubench_pipes.jsonl)


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
    """SM clock / power / throttle reasons for the whole run, each sample stamped on receipt, so that every leg reports
    the clocks seen during ITS timed region (`window`).  NVML is polled every ~4 ms from a thread (the 20-launch
    headline lasts ~26 ms: it gets its own samples); if NVML is unavailable, `nvidia-smi -lms 100` is pumped instead."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    NVML_BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.samples = []           # (t, sm_mhz, max_mhz, power_w, [reasons])
        self.proc = None
        self.source = None
        self._stop = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            reasons_fn(h)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)

            def poll():
                while not self._stop:
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        bits = int(reasons_fn(h))
                        try:
                            power = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
                        except Exception:
                            power = None
                        self.samples.append((time.time(), sm, mx, power, [n for n, b in self.NVML_BITS.items() if bits & b]))
                    except Exception:
                        pass
                    time.sleep(0.004)
            threading.Thread(target=poll, daemon=True).start()
            self.source = "nvml, ~4 ms period"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            self.source = "nvidia-smi -lms 100"
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            t = time.time()
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) < 9:
                continue
            try:
                sm, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            try:
                power = float(parts[3])
            except ValueError:
                power = None
            reasons = [name for name, val in zip(self.NAMES, parts[5:9]) if val.lower().startswith("active")]
            self.samples.append((t, sm, mx, power, reasons))

    def window(self, t0: float, t1: float, pad: float = 0.0) -> dict:
        """clocks seen in [t0, t1] (wall clock); a leg shorter than the sampling period borrows the nearest sample"""
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML and nvidia-smi unavailable)"], "samples": 0}
        if self.proc is not None and t1 - t0 < 0.35:
            time.sleep(0.12)
        rows = [s for s in self.samples if t0 - pad <= s[0] <= t1 + pad]
        borrowed = False
        if not rows and self.samples:
            mid = 0.5 * (t0 + t1)
            rows, borrowed = [min(self.samples, key=lambda s: abs(s[0] - mid))], True
        sm = [r[1] for r in rows]
        reasons = sorted({x for r in rows for x in r[4]})
        power = [r[3] for r in rows if r[3] is not None]
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
               "sm_max_mhz": max(r[2] for r in rows) if rows else None, "power_w_max": max(power) if power else None,
               "reasons": reasons, "samples": len(rows), "source": self.source}
        if borrowed:
            out["note"] = "leg shorter than the sampling period: nearest sample"
        return out

    def stop(self):
        self._stop = True
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


# ------------------------------------------------------------------------------------------ main
def emit(line: dict) -> None:
    """The ONE JSON line, on the real stdout (libraries such as NCCL print banners to fd 1; see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def bind_to_gpu_numa_node(local_rank: int):
    """One process per GPU on a two-socket box: run this rank (and first-touch its pinned staging buffers) on the CPUs
    NVML names as local to its GPU, so that every rank's host->device copies leave through its own socket instead of
    all of them crossing to the socket the launcher happened to start on.  Only the `e2e` legs move host data; the
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


def resample_macs(src_h, src_w, dst_h, dst_w, filt) -> int:
    """integer multiply-accumulates of one Pillow 8bpc two-pass resample (3 channels): the compute side of the roofline"""
    from vision_inspection_system_b200 import tables as T
    ht, vt = T.coeff_table(src_w, dst_w, filt), T.coeff_table(src_h, dst_h, filt)
    h_macs = int(ht.bounds[:, 1].sum()) * src_h * 3 if dst_w != src_w else 0
    v_macs = int(vt.bounds[:, 1].sum()) * dst_w * 3 if dst_h != src_h else 0
    return h_macs + v_macs


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
    ap.add_argument("--legs", choices=["all", "headline"], default="all",
                    help="headline: value + e2e only; all: also sustained, e2e_full_readback, e2e_jpeg and every other config")
    ap.add_argument("--sustain-s", type=float, default=3.0)
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
                "config": config,                      # the SAME config object as our arm (the step is a bounded sample of it)
                "cpu_baseline": {"value": v, "unit": "images/s", "cores": arm.cores, "kind": arm.kind,
                                 "frames_per_step": n,
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
    from vision_inspection_system_b200 import _native as N
    from vision_inspection_system_b200 import geometry as G
    from vision_inspection_system_b200 import sharding as S
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
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count

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
        """ms for `steps` calls: barrier + synchronize on both sides, CUDA events, MAX over ranks; also the wall window"""
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        start.record()
        for _ in range(steps):
            fn()
        end.record()
        barrier()
        t1 = time.time()
        ms = start.elapsed_time(end)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (t0, t1)

    def clocks_of(win):
        return sampler.window(*win) if rank == 0 else None

    def all_ok(flag: bool) -> bool:
        """True iff `flag` holds on EVERY rank (a leg one rank cannot run is skipped by all: the timed loops hold barriers)"""
        if dist is None:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def leg(fn, units_per_call, bytes_per_call, min_s=0.6, warm=2, macs_per_call=0, note=None, max_reps=2000):
        """One device-resident config leg: warm-up (the batch plan is built and cached there), a calibration call, then
        as many calls as fill ~min_s seconds, timed as above; `units` are whole-job (all ranks)."""
        for _ in range(warm):
            fn()
        launches = eng.last_launches
        ms1, _ = timed(fn, 1)
        reps = int(min(max(3, math.ceil(min_s * 1e3 / max(ms1, 1e-3))), max_reps))
        ms, win = timed(fn, reps)
        per = ms / reps
        ck = clocks_of(win)
        rec = {"ms": per, "reps": reps, "images_per_s": n_gpus * units_per_call / per * 1e3,
               "images_per_s_per_gpu": units_per_call / per * 1e3,
               "algorithmic_bytes": int(bytes_per_call), "achieved_gbs": bytes_per_call / per / 1e6,
               "frac": bytes_per_call / per / 1e6 / peak, "launches_per_call": launches, "clocks": ck}
        if macs_per_call:
            mhz = (ck or {}).get("sm_mhz") or 1965.0
            rec["int_macs"] = int(macs_per_call)
            # Pillow taps per second against ONE tap per CUDA-core integer-MAC lane and clock: the yardstick of the IMAD /
            # IDP.4A kernels.  Since the 9+ tap geometries run on the integer tensor path (k_fused_mma, IMMA.16832:
            # 2037 byte MACs / clk / SM measured, three limb MACs per tap) it is a comparison figure, not a utilisation.
            rec["int_mac_issue_frac"] = macs_per_call / (per * 1e-3) / (sm_count * INT_MAC_LANES_PER_CLK_PER_SM * mhz * 1e6)
            rec["int_mac_yardstick"] = "taps/s per CUDA-core integer-MAC lane (62 / clk / SM); 9+ tap legs use IMMA"
        if note:
            rec["note"] = note
        return rec

    # ---------------- device-resident throughput: the headline ----------------
    def step_device():
        eng.preprocess(frames, out=out)

    for _ in range(args.warmup):
        step_device()
    launches_per_step = eng.last_launches
    ms, win = timed(step_device, args.steps)
    value = n_gpus * args.batch * args.steps / (ms / 1e3)
    gpu_launches = launches_per_step * args.steps

    # ---------------- sustained: the same launch back to back for >= sustain_s seconds ----------------
    sustained = None
    if args.legs == "all":
        per_ms = ms / args.steps
        reps = int(max(args.steps, math.ceil(args.sustain_s * 1e3 / per_ms)))
        ms_s, win_s = timed(step_device, reps)
        ach = args.batch * BYTES_PER_IMAGE / (ms_s / reps / 1e3) / 1e9
        sustained = {"seconds": ms_s / 1e3, "launches": reps * launches_per_step, "ms_per_step": ms_s / reps,
                     "value": n_gpus * args.batch * reps / (ms_s / 1e3), "unit": "images/s",
                     "achieved_gbs": ach, "frac": ach / peak, "clocks": clocks_of(win_s)}
        gpu_launches += reps * launches_per_step
    # clocks DURING the headline's own timed region (NVML polled every ~4 ms: the 20 launches take ~26 ms); the sustained
    # leg, which follows immediately with the same launch, carries its own sample (`sustained.clocks`)
    clocks = clocks_of(win)

    # ---------------- end to end: pinned host frames -> device pixel_values (+ small read-back) ----------------
    probe = torch.empty((args.batch, 1176), dtype=torch.float32).pin_memory()

    def step_e2e():
        pv, _grid = eng.preprocess_host(host_pinned, out=out)
        probe.copy_(pv.view(args.batch, ROWS, 1176)[:, 0], non_blocking=True)     # first patch row of every frame

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    e2e_launches = eng.last_launches
    ms_e2e, win_e2e = timed(step_e2e, e2e_steps)
    e2e_value = n_gpus * args.batch * e2e_steps / (ms_e2e / 1e3)
    gpu_launches += e2e_launches * e2e_steps

    # strict variant: the whole pixel_values tensor copied back to pinned host memory as well, chunk by chunk on a
    # third stream while the next chunk uploads (PCIe is full duplex)
    full = {"value": None, "unit": "images/s", "d2h_bytes_per_step": args.batch * ROWS * 1176 * 4,
            "h2d_bytes_per_step": args.batch * H * W * 3}
    host_out = None
    if args.legs == "all":
        try:
            try:
                host_out = torch.empty((args.batch * ROWS, 1176), dtype=torch.float32).pin_memory()
            except Exception:
                host_out = None
            if not all_ok(host_out is not None):
                raise RuntimeError("could not pin the 5.9 GB result buffer on every rank")

            def step_full():
                eng.preprocess_host(host_pinned, out=out, host_out=host_out)

            step_full()
            ms_full, win_full = timed(step_full, 3)
            full["value"] = n_gpus * args.batch * 3 / (ms_full / 1e3)
            full["note"] = ("pinned host frames -> device -> pinned host pixel_values; the read-back of chunk i runs on a "
                            "third stream while chunk i+1 uploads")
            full["clocks"] = clocks_of(win_full)
            gpu_launches += eng.last_launches * 4
            torch.cuda.synchronize()
            check = host_out[:ROWS].clone()
            full["readback_equals_device"] = bool(torch.equal(check, out[:ROWS].cpu()))
        except Exception as e:                         # e.g. not enough pinnable host memory for N x 5.9 GB
            full["note"] = f"skipped: {type(e).__name__}: {e}"
        host_out = None

    # JPEG streams over PCIe, decoded on the GPU (nvJPEG): ~9x fewer bytes per frame than raw RGB
    e2e_jpeg = None
    if args.legs == "all":
        try:
            from PIL import Image
            natural = synth.pattern_frames(H, W)["lowpass"]
            rng = np.random.default_rng(77 + rank)
            streams = []
            for i in range(16):
                buf = io.BytesIO()
                Image.fromarray(np.roll(natural, int(rng.integers(0, W)), axis=1)).save(buf, "JPEG", quality=90, subsampling=2)
                streams.append(buf.getvalue())
            streams = [streams[i % 16] for i in range(args.batch)]

            def step_jpeg():
                pv, _grid = eng.preprocess_jpeg(streams, out=out)
                probe.copy_(pv.view(args.batch, ROWS, 1176)[:, 0], non_blocking=True)

            ok = True
            try:
                step_jpeg()
            except Exception as e:
                ok, why = False, f"{type(e).__name__}: {e}"
            if not all_ok(ok):
                raise RuntimeError("nvJPEG leg failed on a rank" if ok else why)
            jl = eng.last_launches
            ms_j, win_j = timed(step_jpeg, 2)
            e2e_jpeg = {"value": n_gpus * args.batch * 2 / (ms_j / 1e3), "unit": "images/s",
                        "h2d_bytes_per_step": int(sum(len(s) for s in streams)), "d2h_bytes_per_step": args.batch * 1176 * 4,
                        "clocks": clocks_of(win_j),
                        "note": "host JPEG streams (1080p, q90, 4:2:0, low-pass synthetic content) -> nvJPEG batched decode "
                                "on the GPU -> the same kernels; bounded by the nvJPEG decode stage (library, GPU Huffman)"}
            gpu_launches += jl * 3
        except Exception as e:
            e2e_jpeg = {"value": None, "note": f"skipped: {type(e).__name__}: {e}"}

    # ---------------- every other BASELINE config, device resident ----------------
    configs = None
    if args.legs == "all":
        configs = {}
        del frames, out
        torch.cuda.empty_cache()

        def out_rows(h, w, max_pixels):
            dh, dw = G.smart_resize(h, w, G.FACTOR, G.DEFAULT_MIN_PIXELS, max_pixels)
            return (dh // 14) * (dw // 14), (dh, dw)

        def uniform_leg(name, shape, n, max_pixels, seed0, distinct_n=8):
            h, w = shape
            basef = torch.from_numpy(np.stack([synth.noise_frame(seed0 + 1000 * rank + i, h, w) for i in range(distinct_n)])).cuda()
            fr = basef.repeat(n // distinct_n, 1, 1, 1).contiguous()
            rows, dst = out_rows(h, w, max_pixels)
            o = torch.empty((n * rows, 1176), dtype=torch.float32, device="cuda")
            macs = n * resample_macs(h, w, dst[0], dst[1], N.FILTER_BICUBIC)
            configs[name] = leg(lambda: eng.preprocess(fr, max_pixels=max_pixels, out=o), n, n * (h * w * 3 + rows * 4704),
                                macs_per_call=macs, note=f"{n} frames per GPU, {h}x{w} -> {dst[0]}x{dst[1]} "
                                f"(max_pixels {max_pixels}), bytes = frame read + pixel_values written")
            del fr, o, basef
            torch.cuda.empty_cache()

        # config 2, second line: the Qwen2-VL-7B hub max_pixels; config 3: 4K at both settings
        uniform_leg("1080p_hub_max_pixels", (1080, 1920), 128, G.HUB_MAX_PIXELS, 1234)
        uniform_leg("4k_default_max_pixels", (2160, 3840), 64, G.DEFAULT_MAX_PIXELS, 4000)
        uniform_leg("4k_hub_max_pixels", (2160, 3840), 32, G.HUB_MAX_PIXELS, 4000)

        # the agents' LANCZOS thumbnails (src/agents/vlm_inspector.py:64, vlm_auditor.py:91), uint8 in -> uint8 out
        def thumb_leg(name, shape, limit, n):
            h, w = shape
            tw, th = G.thumbnail_size(w, h, limit)
            fr = [torch.from_numpy(synth.noise_frame(4000 + i % 4, h, w)).cuda() for i in range(4)]
            fr = [fr[i % 4] for i in range(n)]
            macs = n * resample_macs(h, w, th, tw, N.FILTER_LANCZOS)
            configs[name] = leg(lambda: eng.resize_batch_u8(fr, th, tw, N.FILTER_LANCZOS), n, n * (h * w * 3 + th * tw * 3),
                                macs_per_call=macs, note=f"{n} frames per GPU, {h}x{w} -> {th}x{tw} LANCZOS uint8, one fused "
                                "launch; bytes = frame read + thumbnail written; compute bound (see int_mac_issue_frac)")
            del fr
            torch.cuda.empty_cache()

        thumb_leg("thumbnail_4k_to_2048", (2160, 3840), 2048, 32)
        thumb_leg("thumbnail_4k_to_1024", (2160, 3840), 1024, 32)
        thumb_leg("thumbnail_1080p_to_1024", (1080, 1920), 1024, 64)

        # config 5: a slice of the mixed-resolution dual Inspector + Auditor stream, byte-balanced over the ranks
        per_gpu = 192
        shapes = synth.mixed_resolution_shapes(per_gpu * n_gpus, seed=9000)
        costs = [S.frame_bytes(h, w) for h, w in shapes]
        mine = S.balanced_shards(costs, n_gpus)[rank]
        cache, fr = {}, []
        for i in mine:
            s = shapes[i]
            if s not in cache:
                cache[s] = [torch.from_numpy(synth.noise_frame(9000 + k, *s)).cuda() for k in range(2)]
            fr.append(cache[s][i % 2])
        job_bytes, job_macs = 0, 0
        for (h, w) in shapes:                                   # whole job: frame once + both roles' pixel_values
            job_bytes += h * w * 3
            for limit in (G.INSPECTOR_MAX_SIZE, G.AUDITOR_MAX_SIZE):
                hh, ww = h, w
                if max(h, w) > limit:
                    ww, hh = G.thumbnail_size(w, h, limit)
                    job_macs += resample_macs(h, w, hh, ww, N.FILTER_LANCZOS)
                rows, dst = out_rows(hh, ww, G.DEFAULT_MAX_PIXELS)
                job_bytes += rows * 4704
                if max(h, w) > limit or limit == G.INSPECTOR_MAX_SIZE:      # frames no role thumbnails are resampled once
                    job_macs += resample_macs(hh, ww, dst[0], dst[1], N.FILTER_BICUBIC)
        res = eng.preprocess_dual(fr)
        total_rows = res["inspector"][0].shape[0] + res["auditor"][0].shape[0]
        del res
        dual_out = torch.empty((total_rows, 1176), dtype=torch.float32, device="cuda")
        rec = leg(lambda: eng.preprocess_dual(fr, out=dual_out), len(shapes) / n_gpus, job_bytes / n_gpus,
                  macs_per_call=job_macs / n_gpus,
                  note=f"{per_gpu} frames per GPU of the seed-9000 mixed-resolution stream, both agents' inputs per frame "
                       "(thumbnail 2048 / 1024 LANCZOS -> processor); bytes = frame once + both pixel_values, the "
                       "thumbnails in between are not credited; frames <= 1024 px are resampled once and stored twice")
        rec["frames_per_s"] = rec.pop("images_per_s")
        rec["frames_per_s_per_gpu"] = rec.pop("images_per_s_per_gpu")
        configs["dual_inspector_auditor_stream"] = rec
        del fr, cache, dual_out
        torch.cuda.empty_cache()

        # config 4: defect overlay on 1024 annotated 1080p BGR frames per GPU
        n_ov, d_ov = 1024, 64
        items = [synth.annotated_frame(7000 + i) for i in range(d_ov)]
        ofr = torch.from_numpy(np.stack([f for f, _ in items])).cuda().repeat(n_ov // d_ov, 1, 1, 1).contiguous()
        boxes = [items[i % d_ov][1] for i in range(n_ov)]
        oshapes = [(1080, 1920)] * n_ov
        plan = eng.plan_overlay(oshapes, boxes)                 # first call: pinned-buffer allocation, sprites, stamps
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan = eng.plan_overlay(oshapes, boxes)
        torch.cuda.synchronize()
        plan_ms = (time.perf_counter() - t0) * 1e3              # steady state: box rules, expansion, tile binning, upload
        ov_bytes = n_ov * 2 * 1080 * 1920 * 3
        configs["overlay_out_of_place"] = leg(lambda: eng.annotate(ofr, boxes, plan=plan), n_ov, ov_bytes,
                                              note=f"{n_ov} annotated 1080p BGR frames per GPU, {sum(len(b) for b in boxes)} boxes; "
                                              "frame copy + in-order tile draw, plan (host) cached; bytes = 2*H*W*3 per frame")
        work = ofr.clone()
        rec = leg(lambda: eng.annotate(work, boxes, plan=plan, inplace=True), n_ov, ov_bytes, note="drawn in place: only the touched 64x16 tiles move; "
                  "frac is quoted against the out-of-place bytes for comparison and may exceed 1")
        configs["overlay_in_place"] = rec
        t0 = time.perf_counter()
        for _ in range(2):
            eng.annotate(ofr, boxes)
        torch.cuda.synchronize()
        api_ms = (time.perf_counter() - t0) / 2 * 1e3
        configs["overlay_host_plan"] = {"plan_overlay_ms_per_call": plan_ms, "plan_overlay_ms_per_frame": plan_ms / n_ov,
                                        "annotate_api_ms_per_call": api_ms,
                                        "annotate_api_images_per_s_per_gpu": n_ov / api_ms * 1e3,
                                        "note": "wall clock on this rank, host planning (box rules in Python, expansion + tile "
                                                "binning in C++, upload) INCLUDED: what one un-planned annotate() call costs"}
        del ofr, work, plan
        torch.cuda.empty_cache()

        # "next" rows of the scope table, on the same footing: image-quality statistics (three exact sums per frame) and
        # the defect heat map (tolerance-specified), both as ONE batch call on device-resident 1080p BGR frames
        n_q = 256
        qfr = torch.from_numpy(synth.frames_1080p(16)).cuda().repeat(n_q // 16, 1, 1, 1).contiguous()
        configs["quality_stats_1080p"] = leg(lambda: eng.quality_stats(qfr), n_q, n_q * 1080 * 1920 * 3,
                                             note=f"{n_q} 1080p BGR frames per GPU -> sum(gray), sum(lap), sum(lap^2) per frame, exact int64; "
                                             "frame descriptors cached for the repeated batch tensor (like the batch plans); bytes = H*W*3 per frame (read only)")
        # the report's comparison panels, one per inspected image, as one batch (two launches)
        n_p = 256
        pa = qfr
        pb = qfr.roll(5, 0).contiguous()
        panel_bytes = 2 * 1080 * 1920 * 3 + 840 * (1422 + 10 + 1422) * 3
        configs["comparison_panels_1080p"] = leg(lambda: eng.side_by_side_batch(pa, pb), n_p, n_p * panel_bytes,
                                                 note=f"{n_p} pairs of 1080p BGR frames -> [840, 2854, 3] side-by-side canvases with header labels "
                                                 "(cv2.resize arithmetic, bit-exact); host record assembly inside the call; bytes = both frames + canvas")
        del qfr, pa, pb
        n_h = 64
        hfr, hdef = [], []
        for i in range(n_h):
            rng = np.random.default_rng(8100 + i)
            hfr.append(torch.from_numpy(rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)).cuda())
            hdef.append(synth.random_defects(rng, int(rng.integers(1, 7))))
        hplan = eng.plan_heatmap([(1080, 1920)] * n_h, hdef)
        rec = leg(lambda: eng.heatmap_batch(hfr, plan=hplan), n_h, n_h * 2 * 1080 * 1920 * 3, warm=3,
                  note=f"{n_h} 1080p BGR frames per GPU, {sum(len(d) for d in hdef)} defects; analytic heat, two Gaussian blurs per defect and "
                  "per frame (float32), max-composite, JET blend; per-defect host parameters cached (plan); bytes = 2*H*W*3 per "
                  "frame; the bound of this leg is the fp32 FMA pipe (<= ~56 k images/s), not HBM")
        configs["heatmap_overlay_1080p"] = rec
        del hfr, hplan
        torch.cuda.empty_cache()
        gpu_launches += sum(int(v.get("launches_per_call", 0)) * int(v.get("reps", 0)) for v in configs.values())

    if rank == 0:
        sampler.stop()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

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
        "plan_cached": True,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": args.batch * H * W * 3,
                "d2h_bytes_per_step": args.batch * 1176 * 4,
                "cpus_bound_per_rank": numa, "clocks": clocks_of(win_e2e),
                "note": "pinned host frames -> device pixel_values (consumer is on the GPU); first patch row of "
                        "every frame read back; see e2e_full_readback for the whole tensor brought back"},
        "e2e_full_readback": full,
        "e2e_jpeg": e2e_jpeg,
        "gpu_launches": int(gpu_launches),
        "roofline": {"bound": "hbm", "kernel": "k_fused_sched (vis_preprocess_fused_sched)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                     "traffic_source": traffic_src, "algorithmic_bytes_per_launch": args.batch * BYTES_PER_IMAGE,
                     "peak_source": peak_src, "bytes_per_image": BYTES_PER_IMAGE,
                     "launches_timed": launches_per_step * args.steps,
                     "sustained_frac": sustained["frac"] if sustained else None},
        "sustained": sustained,
        "configs": configs,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
