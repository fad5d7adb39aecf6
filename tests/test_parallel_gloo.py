"""World-size-2 gloo test of the multi-GPU host logic (frame sharding, the
variable-length keypoint gather, the throughput reduction)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_result(frame: int):
    from sara_b200.api import KEYPOINT_DTYPE

    rng = np.random.default_rng(frame)
    n = int(rng.integers(0, 40))
    k = np.zeros(n, KEYPOINT_DTYPE)
    k["x"] = rng.random(n)
    k["o"] = frame
    k["xi"] = np.arange(n)
    return k, rng.random((n, 128)).astype(np.float32)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from sara_b200 import parallel as P

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = P.shard_frames(7, rank, world)
    local = {f: _fake_result(f) for f in mine}
    merged = P.gather_keypoint_lists(local, dst=0)
    # the tensor path (NCCL from device buffers in production; CPU tensors over gloo here)
    import torch

    local_t = {f: (torch.from_numpy(np.frombuffer(k.tobytes(), np.uint8).reshape(-1, 52).copy()), torch.from_numpy(d))
               for f, (k, d) in local.items()}
    merged_t = P.gather_keypoint_tensors(local_t, dst=0)
    n_local = sum(len(v[0]) for v in local.values())
    tot, mx = P.reduce_throughput(n_local, 1.0 + rank)
    ok = True
    if rank == 0:
        ok = sorted(merged) == list(range(7))
        for f in range(7):
            k, d = _fake_result(f)
            ok = ok and merged[f][0].tobytes() == k.tobytes() and merged[f][1].tobytes() == d.tobytes()
            ok = ok and merged_t[f][0].numpy().tobytes() == k.tobytes() and merged_t[f][1].numpy().tobytes() == d.tobytes()
        ok = ok and tot == sum(len(_fake_result(f)[0]) for f in range(7)) and mx == float(world)
    else:
        ok = merged is None and merged_t is None
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_frames():
    from sara_b200 import parallel as P

    assert P.shard_frames(7, 0, 2) == [0, 2, 4, 6] and P.shard_frames(7, 1, 2) == [1, 3, 5]
    cover = sorted(i for r in range(8) for i in P.shard_frames(256, r, 8))
    assert cover == list(range(256))
    assert P.shard_frames(0, 0, 4) == [] and P.shard_frames(2, 3, 4) == []


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
