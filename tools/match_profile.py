"""One tensor-core k-NN search on SIFT-like descriptor sets (for ncu): python tools/match_profile.py [n1 n2]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sara_b200 as sb  # noqa: E402

n1, n2 = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (11800, 11759)
rng = np.random.default_rng(0)


def sift_like(n):
    d = rng.gamma(0.6, 1.0, (n, 128)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = np.minimum(d, 0.2)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.minimum(d * 512, 255).astype(np.float32)


import torch  # noqa: E402

a = sift_like(n1)
b = np.vstack([a[: n2 // 2] + rng.normal(0, 6, (n2 // 2, 128)), sift_like(n2 - n2 // 2)]).astype(np.float32)
ctx = sb.SiftContext(64, 64, device=0)
ga, gb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
for _ in range(3):
    idx, dist, st = ctx.knn(ga, gb, 3, mode="tensor")
print(st)
