#!/usr/bin/env python
"""SIFT keypoints/s on synthetic 4K frames (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole SIFT chain (Gaussian pyramid -> DoG -> extrema ->
orientation -> descriptor) over one batch of FRAMES_PER_STEP distinct synthetic
3840x2160 frames per GPU (first octave 0, default SIFT arguments, all 10 octaves).

 * value  : keypoints/s with the frames already resident in HBM (device pointers
            passed through the C ABI), CUDA-event timed, max over ranks.
 * e2e    : the same metric through the host-facing call with HOST buffers: every
            step copies the frames from pinned host memory and reads keypoints +
            descriptors back to the host inside the timed region.
 * roofline: the pyramid's dominant kernel (the 25-tap stage launch on octave 0:
            12 B per pixel algorithmic) timed with CUDA events on its
            stream inside the library, against MEASURED_PEAKS.json's HBM copy bandwidth;
            `pyramid` next to it is the whole Gaussian-pyramid + DoG stage at 48 B per
            octave pixel.
 * cpu_baseline: the CPU oracle (a restatement of the reference's CPU path; the
            reference itself cannot be compiled here) on the box's host cores.
 * --impl reference: the same CPU path timed as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160  # the metric's configuration; --frame 1920x1080 measures config C4 instead
FRAMES_PER_STEP = int(os.environ.get("SARA_B200_BENCH_FRAMES", "4"))
METRIC = "sift_keypoints_per_sec_4k"
UNIT = "keypoints/s"
ALGO_BYTES_PER_OCTAVE_PIXEL = 48  # 1 fp32 read + 6 Gaussian + 5 DoG fp32 writes


def octave_pixels(w, h, n_oct):
    tot = 0
    for _ in range(n_oct):
        tot += w * h
        w //= 2
        h //= 2
    return tot


def make_frames(n, w=None, h=None, seed=1234):
    from sara_b200 import synthetic as S

    return [S.tex(w or W4K, h or H4K, seed + i) for i in range(n)]


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        # NVML is polled every few milliseconds (the timed region lasts a fraction of a second);
        # nvidia-smi is the fallback when the binding is missing.
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            while not self.stop_flag.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = 0.0
                self.samples.append([str(sm), str(mx), str(pw)] +
                                    ["Active" if r & bits[k] else "Not Active"
                                     for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                self.stop_flag.wait(0.005)
            return
        except Exception:
            pass
        while not self.stop_flag.is_set():
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    text=True, timeout=5)
                self.samples.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_run(frames, steps, warmup, mode=1):
    """Times the CPU oracle (all host threads when mode == 1)."""
    from oracle import oracle as O

    O.set_threading(mode, 0)
    pp = O.PyramidParams(first_octave_index=0)
    n_kp, times = 0, []
    for i in range(warmup + steps):
        img = frames[i % len(frames)]
        t0 = time.perf_counter()
        r = O.compute_sift_keypoints(img, pp, parallel=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            n_kp += len(r.keypoints)
    return n_kp, float(sum(times)), O.num_threads()


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (the oracle port:
    the reference cannot be compiled in this image) with all host threads.
    Each step is one 3840x2160 frame."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = make_frames(min(2, max(1, args.steps)))
    n_kp, secs, cores = cpu_oracle_run(frames, args.steps, min(args.warmup, 1), mode=1)
    value = n_kp / secs if secs > 0 else 0.0
    sample = f"{args.steps} synthetic {W4K}x{H4K} frames (tex seeds 1234..), one per step, all {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{W4K}x{H4K} synthetic frame, full SIFT (first octave 0, all octaves, 6 scales/octave)",
                   "frames_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import sara_b200 as sb
    from sara_b200 import parallel as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    try:  # keep this rank's host thread and its pinned buffers on the CPUs next to its GPU
        import pynvml as nv

        nv.nvmlInit()
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    F = FRAMES_PER_STEP
    # Per-GPU work is fixed (weak scaling): every rank owns F distinct frames per step.
    frames = make_frames(F, seed=1234 + 100 * rank)
    pp = sb.ImagePyramidParams(first_octave_index=0)
    ctx = sb.SiftContext(W4K, H4K, device=local_rank, max_keypoints=131072, num_slots=F, min_first_octave_index=0)
    L = sb.load_library()
    import ctypes as C
    from sara_b200.api import _SiftArgs, KEYPOINT_DTYPE

    sargs = _SiftArgs(pp._c(), 4.0, 0.01, 10.0, 5)

    d_frames = [torch.from_numpy(f).to(dev) for f in frames]
    streams = [torch.cuda.Stream(device=dev) for _ in range(F)]
    main = torch.cuda.current_stream(dev)

    # Frames are pipelined across steps: slot i is re-armed with the next step's frame as
    # soon as its result has been taken, so copies, kernels and read-backs of neighbouring
    # frames overlap.  `run(steps)` processes exactly steps * F frames, all inside the timed region.
    def run_resident(steps):
        n = 0
        for i in range(F):
            ctx.enqueue_raw(i, d_frames[i].data_ptr(), W4K, H4K, True, sargs, streams[i].cuda_stream)
        for s in range(steps):
            for i in range(F):
                n += ctx.wait(i)
                if s + 1 < steps:
                    ctx.enqueue_raw(i, d_frames[i].data_ptr(), W4K, H4K, True, sargs, streams[i].cuda_stream)
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(main)
        for s in streams:
            s.wait_stream(main)
        t0 = time.perf_counter()
        n = fn(steps)
        for s in streams:
            main.wait_stream(s)
        end.record(main)
        barrier()
        wall = time.perf_counter() - t0
        return n, start.elapsed_time(end) * 1e-3, wall

    # ---- warm-up, then the resident (kernel-side) measurement ------------------------
    run_resident(max(args.warmup, 3))
    launches_per_frame = ctx.timings(0)["total_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    n_kp, secs, wall = timed(run_resident, args.steps)
    clocks = sampler.summary()
    tot_kp, max_secs = P.reduce_throughput(n_kp, secs, device=dev)
    value = tot_kp / max_secs

    # ---- end to end: pinned host frames in, host keypoints + descriptors out ----------
    h_frames = [torch.from_numpy(f).pin_memory() for f in frames]
    cap = 131072
    h_kps = [torch.empty(cap * 52, dtype=torch.uint8).pin_memory() for _ in range(F)]
    h_desc = [torch.empty(cap * 128, dtype=torch.float32).pin_memory() for _ in range(F)]
    d2h = [0]

    def run_e2e(steps):
        n = 0
        for i in range(F):
            ctx.enqueue_raw(i, h_frames[i].data_ptr(), W4K, H4K, False, sargs, streams[i].cuda_stream)
        for s in range(steps):
            m_step = 0
            for i in range(F):
                m_step += ctx.collect_into(i, h_kps[i].data_ptr(), h_desc[i].data_ptr(), cap)
                if s + 1 < steps:
                    ctx.enqueue_raw(i, h_frames[i].data_ptr(), W4K, H4K, False, sargs, streams[i].cuda_stream)
            d2h[0] = m_step * (52 + 512) + F * 16
            n += m_step
        return n

    run_e2e(2)
    n_e2e, secs_e2e, wall_e2e = timed(run_e2e, args.steps)
    tot_e2e, max_e2e = P.reduce_throughput(n_e2e, max(secs_e2e, wall_e2e), device=dev)
    e2e_value = tot_e2e / max_e2e

    # ---- roofline of the pyramid stage (rank 0 only needs it) -----------------------------
    roofline, stage_ms, cpu_baseline = None, None, None
    if rank == 0:
        ctx.set_profiling(True)
        pyr_ms, top_ms, stage_acc = [], [], {}
        reps = max(args.steps, 5)
        top_mb = 0.0
        for r in range(3 + reps):
            i = r % F
            ctx.enqueue_raw(0, d_frames[i].data_ptr(), W4K, H4K, True, sargs, streams[0].cuda_stream)
            ctx.wait(0)
            if r >= 3:
                t = ctx.timings(0)
                pyr_ms.append(t["pyramid"])
                for k in ("pyramid", "extrema", "orientation", "descriptor", "total"):
                    stage_acc.setdefault(k, []).append(t[k])
        # The dominant kernel is timed ALONE (octaves serialised), as the roofline asks.
        ctx.set_octave_overlap(False)
        for r in range(3 + reps):
            i = r % F
            ctx.enqueue_raw(0, d_frames[i].data_ptr(), W4K, H4K, True, sargs, streams[0].cuda_stream)
            ctx.wait(0)
            if r >= 3:
                t = ctx.timings(0)
                top_ms.append(t["pyramid_top_kernel"])
                top_mb = t["pyramid_top_kernel_mbytes"]
        ctx.set_octave_overlap(True)
        ctx.set_profiling(False)
        n_oct = ctx.num_octaves(0)
        algo = ALGO_BYTES_PER_OCTAVE_PIXEL * octave_pixels(W4K, H4K, n_oct)
        mean_ms = float(np.mean(pyr_ms))
        peak, peak_src = 6650.0, "fallback"
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            peak_src = "measured"
        except Exception:
            pass
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "pyramid_traffic.json")))
            traffic, traffic_src = tr.get("top_kernel_dram_bytes"), tr.get("source")
        except Exception:
            pass
        t0 = ctx.timings(0)
        top_mean = float(np.mean(top_ms)) if top_ms and np.mean(top_ms) > 0 else None
        achieved = (top_mb * 1e6) / (top_mean * 1e-3) / 1e9 if top_mean else algo / (mean_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "kernel": f"stage_kernel<25>, octave 0 ({W4K}x{H4K}): reads G(4), writes G(5) and D(4), 12 B/px",
                    "algorithmic_bytes": top_mb * 1e6, "ms": top_mean, "peak_source": peak_src,
                    "note": "bit-exact reference arithmetic (separate fp32 multiply and add per tap) makes this kernel "
                            "fp32-pipe bound, not HBM bound: DESIGN.md section 3",
                    "pyramid": {"achieved": algo / (mean_ms * 1e-3) / 1e9, "frac": algo / (mean_ms * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes": algo, "ms": mean_ms, "launches": t0["pyramid_launches"],
                                "bytes_per_octave_pixel": ALGO_BYTES_PER_OCTAVE_PIXEL}}
        stage_ms = {k: float(np.mean(v)) for k, v in stage_acc.items()}

        # ---- CPU baseline on this box's host cores (bounded sample; reported at N = 1 only) --
        try:
            if world > 1:
                raise RuntimeError("reported at N = 1 only")
            n_cpu, secs_cpu, cores = cpu_oracle_run(frames, 2, 1, mode=1)
            cpu_baseline = {"value": n_cpu / secs_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"2 of the benchmark's {W4K}x{H4K} frames after 1 warm-up, every stage OpenMP over all host threads",
                            "ms_per_frame": 1e3 * secs_cpu / 2}
            n_a, secs_a, _ = cpu_oracle_run(frames, 1, 0, mode=0)
            cpu_baseline["reference_threading"] = {"value": n_a / secs_a, "ms_per_frame": 1e3 * secs_a,
                                                   "note": "pyramid/gradient/orientation serial as in the reference's default build"}
        except Exception as e:  # the oracle is test infrastructure; its absence must not break the bench
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"unavailable: {e}"}

    n_oct_all = ctx.num_octaves(0)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * max_secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{W4K}x{H4K} synthetic frame, full SIFT (first octave 0, {n_oct_all} octaves, 6 scales/octave)",
                       "frames_per_step_per_gpu": F, "keypoints_per_step": tot_kp // args.steps,
                       "l2": f"inputs larger than L2: {F} distinct {W4K * H4K * 4 / 1e6:.0f} MB frames per step, "
                             f"{48 * octave_pixels(W4K, H4K, n_oct_all) / 1e6:.0f} MB of pyramid written per frame",
                       "parallelism": f"frames sharded {F}/GPU/step, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * F * W4K * H4K * 4,
                    "d2h_bytes_per_step": world * d2h[0],
                    "ms_per_step": 1e3 * max_e2e / args.steps},
            "gpu_launches": int(launches_per_frame) * F * args.steps * world,
            "clocks": clocks, "roofline": roofline, "stage_ms_per_frame": stage_ms, "cpu_baseline": cpu_baseline,
            "wall_s": wall,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frame", default="3840x2160", help="frame size WxH (default: the metric's 4K configuration)")
    args = ap.parse_args()
    global W4K, H4K
    W4K, H4K = (int(v) for v in args.frame.lower().split("x"))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
