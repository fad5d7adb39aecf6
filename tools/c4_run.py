"""Config C4 (BASELINE.json): a batch of synthetic 1080p SfM frames sharded over the GPUs of one
node, keypoint buffers gathered to rank 0 over NCCL straight from device memory.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/c4_run.py [--frames 256]

Frame i belongs to rank i mod N (sara_b200.parallel.shard_frames).  Every rank ALSO extracts the
shard of its right neighbour, and both copies are gathered: rank 0 checks that the two results of
every frame -- computed on different GPUs, travelled through NCCL -- are identical byte for byte
(cross-GPU determinism + integrity of the variable-length gather).  Timings: CUDA events, max
over ranks.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sara_b200 as sb  # noqa: E402
from sara_b200 import parallel as P, synthetic as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--size", default="1920x1080")
args = ap.parse_args()
W, H = (int(v) for v in args.size.split("x"))
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

mine = P.shard_frames(args.frames, rank, world)
theirs = P.shard_frames(args.frames, (rank + 1) % world, world) if world > 1 else []
t0 = time.perf_counter()
host = {i: S.sequence_frame(W, H, i) for i in sorted(set(mine) | set(theirs))}
t_gen = time.perf_counter() - t0
pinned = {i: torch.from_numpy(f).pin_memory() for i, f in host.items()}

NS = 4
ctx = sb.SiftContext(W, H, device=local_rank, max_keypoints=65536, num_slots=NS, min_first_octave_index=0)
pp = sb.ImagePyramidParams(first_octave_index=0)


def extract(frames):
    """Pipelined over NS slots; results stay on the device."""
    out = {}
    for j, f in enumerate(frames[:NS]):
        ctx.enqueue(j, pinned[f], pp)
    for j, f in enumerate(frames):
        out[f] = ctx.collect_device(j % NS)
        if j + NS < len(frames):
            ctx.enqueue(j % NS, pinned[frames[j + NS]], pp)
    return out


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


warm = extract(mine[:NS])  # warm-up: kernels, and the NCCL communicator / its first connections
if world > 1:
    P.gather_keypoint_tensors(warm, dst=0)
sync()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
res = extract(mine)
e1.record()
merged = P.gather_keypoint_tensors(res, dst=0) if world > 1 else res
e2.record()
sync()
t_extract, t_gather = e0.elapsed_time(e1) * 1e-3, e1.elapsed_time(e2) * 1e-3
n_local = sum(int(k.shape[0]) for k, _ in res.values())
tot, mx_extract = P.reduce_throughput(n_local, t_extract, device=dev)
_, mx_total = P.reduce_throughput(n_local, t_extract + t_gather, device=dev)

check = None
if world > 1:
    res2 = extract(theirs)
    merged2 = P.gather_keypoint_tensors(res2, dst=0)
    if rank == 0:
        assert sorted(merged) == sorted(merged2) == list(range(args.frames))
        bad = [f for f in merged if not (torch.equal(merged[f][0], merged2[f][0]) and torch.equal(merged[f][1], merged2[f][1]))]
        assert not bad, f"frames {bad[:8]} differ between the owner GPU and its neighbour"
        check = f"{args.frames} frames: owner GPU == neighbour GPU, byte for byte, after the NCCL gather"
if rank == 0:
    gathered_bytes = sum(k.numel() + d.numel() * 4 for k, d in merged.values())
    print(json.dumps({
        "config": f"C4: {args.frames} synthetic {W}x{H} SfM frames (sequence_frame), sharded i mod {world}, full SIFT, "
                  "keypoint buffers gathered to rank 0 over NCCL from device memory",
        "n_gpus": world, "keypoints": tot, "extract_s": mx_extract, "extract_plus_gather_s": mx_total,
        "gather_s_rank0": t_gather, "gathered_MB": gathered_bytes / 1e6,
        "keypoints_per_s_extract": tot / mx_extract, "keypoints_per_s_with_gather": tot / mx_total,
        "frames_per_s_with_gather": args.frames / mx_total, "check": check, "frame_generation_s_rank0": t_gen}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
