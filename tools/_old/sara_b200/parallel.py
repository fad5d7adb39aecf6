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


def gather_keypoint_tensors(local: dict, dst: int = 0, group=None):
    """Gathers {frame_index: (keypoints, descriptors)} onto `dst` WITHOUT leaving the device:
    keypoints are (n, 52) uint8 tensors (KEYPOINT_DTYPE records), descriptors (n, 128) float32
    tensors, both on this rank's device (CUDA + NCCL in production, CPU + gloo in the tests).
    Counts travel first (one all_gather of a fixed-size int64 table), then every rank
    contributes one flat byte buffer padded to the longest (dist.gather); `dst` slices the
    frames back out as views of the gathered buffers.  Returns the merged dict on `dst`, None
    elsewhere.  SURVEY.md section 8e."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    frames = sorted(local)
    dev = local[frames[0]][1].device if frames else torch.device("cpu")

    n_local = torch.tensor([len(frames)], dtype=torch.int64, device=dev)
    all_n = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(all_n, n_local, group=group)
    max_frames = max(int(t.item()) for t in all_n)
    table = torch.zeros((max(max_frames, 1), 2), dtype=torch.int64, device=dev)  # (frame, count)
    for i, f in enumerate(frames):
        table[i, 0], table[i, 1] = f, local[f][0].shape[0]
    all_tab = [torch.zeros_like(table) for _ in range(world)]
    dist.all_gather(all_tab, table, group=group)
    all_tab = [t.cpu().numpy() for t in all_tab]
    rec = 52 + 512
    totals = [int(all_tab[r][: int(all_n[r].item()), 1].sum()) * rec for r in range(world)]
    max_bytes = max(max(totals), 1)

    buf = torch.zeros(max_bytes, dtype=torch.uint8, device=dev)
    off = 0
    for f in frames:
        kp, desc = local[f]
        n = kp.shape[0]
        buf[off: off + 52 * n] = kp.reshape(-1)
        off += 52 * n
        buf[off: off + 512 * n] = desc.contiguous().view(torch.uint8).reshape(-1)
        off += 512 * n
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = {}
    for r in range(world):
        off = 0
        for i in range(int(all_n[r].item())):
            f, n = int(all_tab[r][i, 0]), int(all_tab[r][i, 1])
            kp = gathered[r][off: off + 52 * n].view(n, 52)
            off += 52 * n
            desc = gathered[r][off: off + 512 * n].view(torch.float32).view(n, 128)
            off += 512 * n
            out[f] = (kp, desc)
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
