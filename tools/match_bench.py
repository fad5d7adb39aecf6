"""Matching row benchmark: two 4K frames of the synthetic sequence -> SIFT on the GPU -> AnnMatcher.

    python tools/match_bench.py [WxH]

Prints one JSON line: the k-NN search on the tensor cores and on the CUDA cores (CUDA events inside the
library), the whole compute_matches call (wall clock, host descriptors in, matches out), and the CPU side:
the oracle's exact search (OpenMP) and -- when oracle/_ref is present -- the reference's own FLANN KD-tree
forest (single thread, as AnnMatcher.cpp:242-254 calls it)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sara_b200 as sb  # noqa: E402
from sara_b200 import synthetic as S  # noqa: E402

W, H = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160").split("x"))
ctx = sb.SiftContext(W, H, device=0, min_first_octave_index=0)
pp = sb.ImagePyramidParams(first_octave_index=0)
kls = [ctx.compute_sift_keypoints(S.sequence_frame(W, H, i), pp) for i in (0, 1)]
d1, d2 = kls[0].descriptors, kls[1].descriptors
n1, n2 = len(d1), len(d2)
out = {"frames": f"{W}x{H} sequence frames 0 and 1", "n1": n1, "n2": n2}

import torch  # noqa: E402

g1, g2 = torch.from_numpy(d1).cuda(), torch.from_numpy(d2).cuda()
for mode in ("tensor", "scalar"):
    ms = []
    for _ in range(6):
        _, _, st = ctx.knn(g1, g2, 3, mode=mode)
        ms.append(st["gpu_ms"])
    out[f"knn_{mode}_ms"] = float(np.median(ms[1:]))
    out[f"knn_{mode}_stats"] = st
flops = 2.0 * n1 * n2 * 384
out["tensor_pass_note"] = "bf16 split: K = 384 per pair; gpu_ms covers split + MMA + re-rank (+ fallback)"
out["tensor_tflops_whole_search"] = flops / (out["knn_tensor_ms"] * 1e-3) / 1e12
for thr in (0.6, 1.2):
    t = []
    for _ in range(4):
        t0 = time.perf_counter()
        m, st = ctx.compute_matches(d1, d2, thr, kls[0].features, kls[1].features, return_stats=True)
        t.append(time.perf_counter() - t0)
    out[f"compute_matches_{thr}"] = {"wall_ms": float(np.median(t[1:]) * 1e3), "matches": int(len(m)), "gpu_ms": st["gpu_ms"],
                                     "redone": st["n_redone"]}

from oracle import match as M  # noqa: E402
from oracle import oracle as O  # noqa: E402

t0 = time.perf_counter()
i0, dd0 = M.knn_linear(d2, d1, 3)
out["cpu_exact_port_ms"] = (time.perf_counter() - t0) * 1e3
out["cpu_threads"] = O.num_threads()
idx, dist, _ = ctx.knn(g1, g2, 3, mode="tensor")
out["gpu_equals_cpu_exact"] = bool(np.array_equal(idx, i0) and np.array_equal(dist.view(np.uint32), dd0.view(np.uint32)))
if M.have_ref():
    t0 = time.perf_counter()
    kd = M.FlannRef(d2, "kdtree")
    out["flann_kdtree_build_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    ik, dk = kd.knn(d1, 3)
    out["flann_kdtree_search_ms"] = (time.perf_counter() - t0) * 1e3
    out["flann_kdtree_recall_nn"] = float((ik[:, 0] == i0[:, 0]).mean())
    t0 = time.perf_counter()
    mk = M.ann_match(d1, d2, 0.6, kls[0].features, kls[1].features, backend="kdtree")
    out["reference_flann_ann_match_0.6_ms"] = (time.perf_counter() - t0) * 1e3
    me = ctx.compute_matches(d1, d2, 0.6, kls[0].features, kls[1].features)
    a = {(int(x["x_index"]), int(x["y_index"])) for x in me}
    b = {(int(x["x_index"]), int(x["y_index"])) for x in mk}
    out["matches_exact_vs_kdtree_0.6"] = {"exact": len(a), "kdtree": len(b), "common": len(a & b)}
print(json.dumps(out))
