#!/usr/bin/env python
"""SIFT keypoints/s on synthetic 4K frames (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frame WxH]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole SIFT chain (Gaussian pyramid -> DoG -> extrema ->
orientation -> descriptor) over one batch of FRAMES_PER_STEP (16) distinct synthetic
3840x2160 frames per GPU (first octave 0, default SIFT arguments, all 10 octaves), SLOTS (8)
frames in flight.

 * value   : keypoints/s with the frames already resident in HBM (device pointers passed
             through the C ABI), CUDA-event timed, max over ranks.
 * e2e     : the same metric through the host-facing C-ABI call with HOST buffers: every
             step copies the float32 frames from pinned host memory and reads keypoints +
             descriptors back to the host inside the timed region.  Beside it:
             `rgb8` / `gray8` -- the ingest entry points (8-bit frames in, the reference's
             from_rgb8_to_gray32f done on the device), `pageable` -- the synchronous drop-in
             call (sara_b200_sift) on ordinary pageable memory, one frame at a time.
 * roofline: the pyramid's dominant kernel (the 25-tap launch on octave 0: 12 B per pixel
             algorithmic) timed alone with CUDA events on its stream
             inside the library, against MEASURED_PEAKS.json's HBM copy bandwidth; `pyramid`
             next to it is the whole Gaussian-pyramid + DoG stage at 48 B per octave pixel.
 * cpu_baseline / --impl reference: the CPU oracle (a restatement of the reference's CPU path:
             the reference itself needs Eigen/Boost/HDF5 and cannot be compiled in this image)
             with an EXPLICIT OpenMP thread count = the cores this process may run on (torchrun
             exports OMP_NUM_THREADS=1, which must not reach this arm), one 4K frame per step.
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
FRAMES_PER_STEP = int(os.environ.get("SARA_B200_BENCH_FRAMES", "16"))
SLOTS = int(os.environ.get("SARA_B200_BENCH_SLOTS", "8"))
METRIC = "sift_keypoints_per_sec_4k"
UNIT = "keypoints/s"
ALGO_BYTES_PER_OCTAVE_PIXEL = 48  # 1 fp32 read + 6 Gaussian + 5 DoG fp32 writes
CAP = 131072                      # keypoint capacity per frame


def octave_pixels(w, h, n_oct):
    tot = 0
    for _ in range(n_oct):
        tot += w * h
        w //= 2
        h //= 2
    return tot


def make_frames(n, w=None, h=None, seed=1234):
    """n distinct frames: `tex` scenes (4 per seed group) and their mirror images."""
    from sara_b200 import synthetic as S

    w, h = w or W4K, h or H4K
    base = [S.tex(w, h, seed + i) for i in range(min(n, 4))]
    flips = [lambda a: a, lambda a: a[:, ::-1], lambda a: a[::-1, :], lambda a: a[::-1, ::-1]]
    out = []
    for i in range(n):
        out.append(np.ascontiguousarray(flips[(i // 4) % 4](base[i % len(base)])))
    return out


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (NVML, nvidia-smi as fallback)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
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
        sm, mx, pw, reasons = [], 0.0, [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                pw.append(float(s[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_run(frames, steps, warmup, mode=1, budget_s=None):
    """Times the CPU oracle, one frame per step, on an explicit number of OpenMP threads
    (mode 1: every stage threaded; mode 0: the reference's default build, whose pyramid,
    gradient and orientation stages are serial).  Returns (keypoints, seconds, threads, steps)."""
    from oracle import oracle as O

    threads = host_threads()
    O.set_threading(mode, threads)
    pp = O.PyramidParams(first_octave_index=0)
    n_kp, secs, done = 0, 0.0, 0
    for i in range(warmup + steps):
        img = frames[i % len(frames)]
        t0 = time.perf_counter()
        r = O.compute_sift_keypoints(img, pp, parallel=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            secs += dt
            n_kp += len(r.keypoints)
            done += 1
            if budget_s is not None and secs >= budget_s and done >= 2:
                break
    return n_kp, secs, O.num_threads(), done


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (the oracle port: the
    reference cannot be compiled in this image) on ALL the host threads this process
    may use; rank 0 alone works under torchrun.  Each step is one 3840x2160 frame."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warmup = min(max(args.warmup, 1), 2)
    frames = make_frames(4)
    n_kp, secs, cores, done = cpu_oracle_run(frames, args.steps, warmup, mode=1, budget_s=150.0)
    value = n_kp / secs if secs > 0 else 0.0
    sample = (f"{done} synthetic {W4K}x{H4K} frames (tex seeds 1234..1237), one per step, after {warmup} warm-up; "
              f"every stage OpenMP on {cores} host threads (explicit count; OMP_NUM_THREADS ignored)")
    if done < args.steps:
        sample += f"; bounded to {done} of the {args.steps} requested steps (150 s budget)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": warmup, "ms_per_step": 1e3 * secs / max(done, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{W4K}x{H4K} synthetic frame, full SIFT (first octave 0, all octaves, 6 scales/octave)",
                   "frames_per_step": 1, "reference_kind": "oracle port of the reference CPU path (oracle/sift_oracle.cpp)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    F, NS = FRAMES_PER_STEP, SLOTS
    assert F % NS == 0

    # Per-GPU work is fixed (weak scaling): every rank owns F distinct frames per step.
    frames = make_frames(F, seed=1234 + 100 * rank)

    # ---- CPU baseline first (N = 1 only): before CUDA, NCCL or any CPU-affinity change -------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref_frames = make_frames(4) if rank else frames[:4]
            n_cpu, secs_cpu, cores, done = cpu_oracle_run(ref_frames, 12, 1, mode=1, budget_s=12.0)
            cpu_baseline = {"value": n_cpu / secs_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{done} of the benchmark's {W4K}x{H4K} frames, one per step, after 1 warm-up; every "
                                      f"stage OpenMP on {cores} host threads (same procedure as --impl reference)",
                            "ms_per_frame": 1e3 * secs_cpu / done}
            n_a, secs_a, _, done_a = cpu_oracle_run(ref_frames, 2, 0, mode=0)
            cpu_baseline["reference_threading"] = {
                "value": n_a / secs_a, "ms_per_frame": 1e3 * secs_a / done_a,
                "note": "pyramid / gradient / orientation serial as in the reference's default (non-Halide) build"}
        except Exception as e:  # the oracle is test infrastructure; its absence must not break the bench
            cpu_baseline = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port",
                            "sample": f"unavailable: {e}"}

    import torch
    import torch.distributed as dist

    import sara_b200 as sb
    from sara_b200 import parallel as P
    from sara_b200.api import _SiftArgs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        try:  # keep this rank's host thread and its pinned buffers on the CPUs next to its GPU
            import pynvml as nv

            nv.nvmlInit()
            nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=dev)

    pp = sb.ImagePyramidParams(first_octave_index=0)
    ctx = sb.SiftContext(W4K, H4K, device=local_rank, max_keypoints=CAP, num_slots=NS, min_first_octave_index=0)
    sargs = _SiftArgs(pp._c(), 4.0, 0.01, 10.0, 5)

    d_frames = [torch.from_numpy(f).to(dev) for f in frames]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    main = torch.cuda.current_stream(dev)
    enqueue_s = [0.0, 0]

    # Frames are pipelined: slot i is re-armed with the next frame as soon as its result has been
    # taken, so copies, kernels and read-backs of neighbouring frames overlap.  run(steps)
    # processes exactly steps * F frames, all inside the timed region.  ctx.wait raises on a
    # keypoint-capacity overflow, so a counted keypoint is always a described keypoint.
    def pipeline(steps, enqueue, take):
        total, n = steps * F, 0
        for j in range(min(NS, total)):
            enqueue(j % NS, j % F)
        for j in range(total):
            n += take(j % NS)
            nxt = j + NS
            if nxt < total:
                t0 = time.perf_counter()
                enqueue(nxt % NS, nxt % F)
                enqueue_s[0] += time.perf_counter() - t0
                enqueue_s[1] += 1
        return n

    def run_resident(steps):
        return pipeline(steps,
                        lambda s, f: ctx.enqueue_raw(s, d_frames[f].data_ptr(), W4K, H4K, True, sargs,
                                                     streams[s].cuda_stream),
                        lambda s: ctx.wait(s))

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
    enqueue_s[:] = [0.0, 0]
    sampler = ClockSampler(local_rank)
    sampler.start()
    n_kp, secs, wall = timed(run_resident, args.steps)
    clocks = sampler.summary()
    host_enqueue_us = 1e6 * enqueue_s[0] / max(enqueue_s[1], 1)
    tot_kp, max_secs = P.reduce_throughput(n_kp, secs, device=dev)
    value = tot_kp / max_secs

    # ---- end to end: host frames in, host keypoints + descriptors out ------------------
    h_kps = [torch.empty(CAP * 52, dtype=torch.uint8).pin_memory() for _ in range(NS)]
    h_desc = [torch.empty(CAP * 128, dtype=torch.float32).pin_memory() for _ in range(NS)]

    def e2e_leg(h_in, channels):
        """channels: 0 float32 gray, 1 gray8, 3 rgb8 (pinned host tensors)."""
        d2h = [0]

        def enqueue(s, f):
            if channels == 0:
                ctx.enqueue_raw(s, h_in[f].data_ptr(), W4K, H4K, False, sargs, streams[s].cuda_stream)
            else:
                ctx.enqueue_raw_u8(s, h_in[f].data_ptr(), W4K, H4K, channels, False, sargs, streams[s].cuda_stream)

        def take(s):
            m = ctx.collect_into(s, h_kps[s].data_ptr(), h_desc[s].data_ptr(), CAP)
            d2h[0] += m * (52 + 512) + 32
            return m

        def run(steps):
            return pipeline(steps, enqueue, take)

        run(1)
        d2h[0] = 0
        n, s_ev, s_wall = timed(run, args.steps)
        tot, mx = P.reduce_throughput(n, max(s_ev, s_wall), device=dev)
        bpp = 4 if channels == 0 else channels
        return {"value": tot / mx, "unit": UNIT, "h2d_bytes_per_step": world * F * W4K * H4K * bpp,
                "d2h_bytes_per_step": world * d2h[0] // max(args.steps, 1), "ms_per_step": 1e3 * mx / args.steps}

    h_frames = [torch.from_numpy(f).pin_memory() for f in frames]
    e2e = e2e_leg(h_frames, 0)
    e2e["input"] = "float32 gray frames in pinned host memory (the C ABI's ImageView<float> contract)"
    # What the host link alone allows: the step's float frames copied host -> device on the same streams, no
    # compute.  When this takes as long as the e2e step, the e2e figure is the PCIe figure, not the kernels'.
    d_sink = [torch.empty(H4K, W4K, dtype=torch.float32, device=dev) for _ in range(NS)]

    def copies_only(steps):
        for j in range(steps * F):
            with torch.cuda.stream(streams[j % NS]):
                d_sink[j % NS].copy_(h_frames[j % F], non_blocking=True)
        return 0

    copies_only(1)
    _, c_ev, c_wall = timed(copies_only, args.steps)
    c_ms = 1e3 * max(c_ev, c_wall) / args.steps
    e2e["h2d_only"] = {"ms_per_step": c_ms, "GB/s_per_gpu": F * W4K * H4K * 4 / (c_ms * 1e-3) / 1e9,
                       "how": "the step's pinned float frames copied to the device on the same streams, nothing else running"}
    del d_sink
    u8_gray = [np.clip(np.rint(f * 255.0), 0, 255).astype(np.uint8) for f in frames]
    h_gray8 = [torch.from_numpy(g).pin_memory() for g in u8_gray]
    e2e_gray8 = e2e_leg(h_gray8, 1)
    h_rgb8 = [torch.from_numpy(np.ascontiguousarray(np.repeat(g[:, :, None], 3, axis=2))).pin_memory() for g in u8_gray]
    e2e_rgb8 = e2e_leg(h_rgb8, 3)
    del h_rgb8, h_gray8
    e2e["rgb8"] = dict(e2e_rgb8, input="interleaved RGB8 frames (what the reference's video loop decodes), "
                                        "from_rgb8_to_gray32f on the device")
    e2e["gray8"] = dict(e2e_gray8, input="gray8 frames, converted on the device")

    # the synchronous drop-in call on pageable memory (what include/sara_b200.hpp does), N = 1 view
    if rank == 0:
        import ctypes as C

        from sara_b200.api import KEYPOINT_DTYPE

        L = sb.load_library()
        kps = np.empty(CAP, KEYPOINT_DTYPE)
        desc = np.empty((CAP, 128), np.float32)
        m = C.c_int()
        n_pg, t_pg = 0, 0.0
        for i in range(2 + 8):
            f = frames[i % F]
            t0 = time.perf_counter()
            rc = L.sara_b200_sift(ctx._ctx, f.ctypes.data, W4K, H4K, 0, C.byref(sargs), kps.ctypes.data,
                                  desc.ctypes.data, CAP, C.byref(m))
            dt = time.perf_counter() - t0
            assert rc == 0, rc
            if i >= 2:
                n_pg += m.value
                t_pg += dt
        e2e["pageable"] = {"value": n_pg / t_pg, "unit": UNIT, "ms_per_frame": 1e3 * t_pg / 8, "per_gpu": True,
                           "input": "sara_b200_sift: one frame at a time, pageable host memory in and out, wall clock"}

    # ---- roofline of the pyramid stage (rank 0 only needs it) -----------------------------
    roofline, stage_ms = None, None
    if rank == 0:
        ctx.set_profiling(True)  # under profiling the stages do not overlap (no early classify)
        pyr_ms, top_ms, stage_acc = [], [], {}
        reps = 10
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
                    "kernel": f"march_kernel<25>, octave 0 ({W4K}x{H4K}): reads G(4), writes G(5) and D(4), 12 B/px",
                    "algorithmic_bytes": top_mb * 1e6, "ms": top_mean, "peak_source": peak_src,
                    "note": "bit-exact reference arithmetic (separate fp32 multiply and add per tap) makes the pyramid "
                            "fp32-pipe bound, not HBM bound: DESIGN.md section 3",
                    "fp32_pipe": None,
                    "pyramid": {"achieved": algo / (mean_ms * 1e-3) / 1e9, "frac": algo / (mean_ms * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes": algo, "ms": mean_ms, "launches": t0["pyramid_launches"],
                                "bytes_per_octave_pixel": ALGO_BYTES_PER_OCTAVE_PIXEL,
                                "how": "CUDA events around the pyramid stage of single frames (one frame in flight, "
                                       "octaves overlapped on side streams), mean of 10"}}
        stage_ms = {k: float(np.mean(v)) for k, v in stage_acc.items()}
        if top_mean:
            # what actually bounds that launch: un-fused fp32 multiplies and adds.  Per pixel and
            # pass the contract needs K additions and c + 1 multiplies (symmetric taps share their products).
            K, c = 25, 12
            lane_ops = 2.0 * (K + c + 1) * (top_mb * 1e6 / 12.0)
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            pipe_peak = 148 * 128 * sm_mhz * 1e6
            roofline["fp32_pipe"] = {"achieved_lane_ops_per_s": lane_ops / (top_mean * 1e-3), "peak_lane_ops_per_s": pipe_peak,
                                     "frac": lane_ops / (top_mean * 1e-3) / pipe_peak,
                                     "how": "2 passes x (25 adds + 13 multiplies) per pixel of the launch / its duration, against "
                                            "148 SMs x 128 fp32 lanes x the SM clock sampled during the run (no FMA: 1 op per lane-clock)"}

    # ---- matching row: AnnMatcher between two frames of the batch, N = 1 view -------------
    matching = None
    if rank == 0:
        try:
            # a frame and the same scene moved by (3, 2) pixels with fresh sensor noise
            moved = np.roll(frames[0], (2, 3), axis=(0, 1)) + np.random.default_rng(7).normal(0, 0.01, frames[0].shape)
            d_pair = [d_frames[0], torch.from_numpy(np.clip(moved, 0, 1).astype(np.float32)).to(dev)]
            kl = []
            for i in (0, 1):
                ctx.enqueue_raw(0, d_pair[i].data_ptr(), W4K, H4K, True, sargs, streams[0].cuda_stream)
                kl.append(ctx.collect(0))
            g = [torch.from_numpy(k.descriptors).to(dev) for k in kl]
            ms = []
            for _ in range(6):
                _, _, st = ctx.knn(g[0], g[1], 3, mode="tensor")
                ms.append(st["gpu_ms"])
            knn_ms = float(np.median(ms[1:]))
            t = []
            for _ in range(5):
                t0 = time.perf_counter()
                m = ctx.compute_matches(g[0], g[1], 0.6, kl[0].features, kl[1].features)
                t.append(time.perf_counter() - t0)
            n1, n2 = len(kl[0]), len(kl[1])
            flops = 2.0 * n1 * n2 * 384
            bf16_peak = None
            try:
                bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
            except Exception:
                pass
            matching = {"n1": n1, "n2": n2, "knn3_gpu_ms": knn_ms, "stats": st,
                        "ann_match_0.6": {"wall_ms": 1e3 * float(np.median(t[1:])), "matches": int(len(m)),
                                          "input": "descriptors resident on the device, both directions, match list to the host"},
                        "tensor": {"bound": "tensor", "achieved": flops / (knn_ms * 1e-3) / 1e12, "peak": bf16_peak,
                                   "unit": "TFLOP/s", "frac": (flops / (knn_ms * 1e-3) / 1e12 / bf16_peak) if bf16_peak else None,
                                   "how": "2 n1 n2 x 384 (bf16 split: hi.hi + lo.hi + hi.lo) / the WHOLE search (split, tcgen05 "
                                          "candidates, selection, exact re-ranking), CUDA events inside the library"}}
            if cpu_baseline is not None and cpu_baseline.get("value"):
                from oracle import match as OM

                t0 = time.perf_counter()
                i0, d0 = OM.knn_linear(kl[1].descriptors, kl[0].descriptors, 3)
                matching["cpu_exact_knn3_ms"] = 1e3 * (time.perf_counter() - t0)
                idx, dist, _ = ctx.knn(g[0], g[1], 3, mode="tensor")
                matching["gpu_equals_cpu_exact"] = bool(np.array_equal(idx, i0) and
                                                        np.array_equal(dist.view(np.uint32), d0.view(np.uint32)))
                if OM.have_ref():
                    t0 = time.perf_counter()
                    kd = OM.FlannRef(kl[1].descriptors, "kdtree")
                    ik, _ = kd.knn(kl[0].descriptors, 3)
                    matching["reference_flann_kdtree_ms"] = 1e3 * (time.perf_counter() - t0)
                    matching["reference_flann_kdtree_recall_nn"] = float((ik[:, 0] == i0[:, 0]).mean())
        except Exception as e:
            matching = {"unavailable": repr(e)}

    n_oct_all = ctx.num_octaves(0)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * max_secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{W4K}x{H4K} synthetic frame, full SIFT (first octave 0, {n_oct_all} octaves, 6 scales/octave)",
                       "frames_per_step_per_gpu": F, "frames_in_flight_per_gpu": NS,
                       "keypoints_per_step": tot_kp // args.steps, "keypoint_capacity_per_frame": CAP,
                       "l2": f"inputs larger than L2: {F} distinct {W4K * H4K * 4 / 1e6:.0f} MB frames per step, "
                             f"{48 * octave_pixels(W4K, H4K, n_oct_all) / 1e6:.0f} MB of pyramid written per frame",
                       "parallelism": f"frames sharded {F}/GPU/step, no data-path collective"},
            "e2e": e2e,
            "gpu_launches": int(launches_per_frame) * F * args.steps * world,
            "gpu_launches_per_frame": int(launches_per_frame),
            "host_enqueue_us_per_frame": host_enqueue_us,
            "clocks": clocks, "roofline": roofline, "stage_ms_per_frame": stage_ms, "cpu_baseline": cpu_baseline,
            "matching": matching, "timed_region_s": max_secs, "wall_s": wall,
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # Nothing but the JSON line may reach stdout: NCCL prints its version banner there (at NCCL_DEBUG =
    # VERSION, WARN or INFO), libraries may chat.  File descriptor 1 is pointed at stderr for the whole run
    # and the line is written to a duplicate of the original stdout.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frame", default="3840x2160", help="frame size WxH (default: the metric's 4K configuration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global W4K, H4K
    W4K, H4K = (int(v) for v in args.frame.lower().split("x"))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
