"""Multi-GPU plumbing for the per-frame extraction step.

Frames are independent units: frame i is owned by rank i mod world.  There is no
exchange step inside extraction, hence no data-path collective; the only
communication is the result gather (variable-length keypoint buffers) and the
count reduction the throughput metric needs.  One process per GPU,
torch.distributed (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_frames(n_frames: int, rank: int, world: int) -> list[int]:
    """Indices of the frames rank `rank` owns (round-robin, like a video stream
    dealt to the GPUs)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def gather_keypoint_lists(local: dict, dst: int = 0, group=None, device=None):
    """Gathers {frame_index: (features, descriptors)} from every rank onto `dst`.

    features: structured array (sara_b200.KEYPOINT_DTYPE, 52 bytes per keypoint),
    descriptors: (n, 128) float32.  Counts travel first (all_gather of one int64
    vector), then every rank sends one padded byte buffer; `dst` slices it back.
    Returns the merged dict on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = device if device is not None else torch.device("cpu")

    frames = sorted(local)
    counts = [len(local[f][0]) for f in frames]
    meta = torch.tensor([len(frames)] + [v for fc in zip(frames, counts) for v in fc], dtype=torch.int64)
    n_meta = torch.tensor([meta.numel()], dtype=torch.int64, device=dev)
    all_n = [torch.zeros_like(n_meta) for _ in range(world)]
    dist.all_gather(all_n, n_meta, group=group)
    max_meta = int(max(int(t.item()) for t in all_n))
    meta_pad = torch.zeros(max_meta, dtype=torch.int64, device=dev)
    meta_pad[: meta.numel()] = meta.to(dev)
    all_meta = [torch.zeros_like(meta_pad) for _ in range(world)]
    dist.all_gather(all_meta, meta_pad, group=group)

    rec = 52 + 512
    per_rank_bytes = []
    for m in all_meta:
        m = m.cpu().numpy()
        nf = int(m[0])
        per_rank_bytes.append(int(sum(m[2 + 2 * i] for i in range(nf))) * rec)
    max_bytes = max(max(per_rank_bytes), 1)

    payload = np.zeros(max_bytes, np.uint8)
    off = 0
    for f in frames:
        feats, desc = local[f]
        n = len(feats)
        payload[off: off + 52 * n] = np.frombuffer(np.ascontiguousarray(feats).tobytes(), np.uint8)
        off += 52 * n
        payload[off: off + 512 * n] = np.frombuffer(np.ascontiguousarray(desc, np.float32).tobytes(), np.uint8)
        off += 512 * n
    buf = torch.from_numpy(payload).to(dev)
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gathered, dst=dst, group=group)
    if rank != dst:
        return None

    from .api import KEYPOINT_DTYPE

    out = {}
    for r in range(world):
        m = all_meta[r].cpu().numpy()
        nf = int(m[0])
        raw = gathered[r].cpu().numpy()
        off = 0
        for i in range(nf):
            f, n = int(m[1 + 2 * i]), int(m[2 + 2 * i])
            feats = np.frombuffer(raw[off: off + 52 * n].tobytes(), KEYPOINT_DTYPE).copy()
            off += 52 * n
            desc = np.frombuffer(raw[off: off + 512 * n].tobytes(), np.float32).reshape(n, 128).copy()
            off += 512 * n
            out[f] = (feats, desc)
    return out


def reduce_throughput(n_keypoints: int, seconds: float, device=None, group=None):
    """(sum of keypoints over ranks, max of the elapsed time over ranks)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return n_keypoints, seconds
    dev = device if device is not None else torch.device("cpu")
    tot = torch.tensor([float(n_keypoints)], dtype=torch.float64, device=dev)
    mx = torch.tensor([float(seconds)], dtype=torch.float64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return int(round(tot.item())), float(mx.item())
