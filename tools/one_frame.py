"""Runs the whole SIFT chain on a few synthetic frames, one at a time, with direct launches (no CUDA
graph) -- the target of `ncu --metrics gpu__time_duration.sum` launch lists and `--set full` captures.
    python tools/one_frame.py [WxH] [n_frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sara_b200 as sb  # noqa: E402
from sara_b200 import synthetic as S  # noqa: E402

size = next((a for a in sys.argv[1:] if "x" in a), "3840x2160")
W, H = (int(v) for v in size.split("x"))
n = next((int(a) for a in sys.argv[1:] if a.isdigit()), 3)
ctx = sb.SiftContext(W, H, max_keypoints=131072, min_first_octave_index=0)
ctx.set_graphs(False)
pp = sb.ImagePyramidParams(first_octave_index=0)
img = torch.from_numpy(S.tex(W, H, 1234)).cuda()
for i in range(n):
    ctx.enqueue(0, img, pp)
    print("frame", i, "keypoints", ctx.wait(0), "launches", ctx.timings(0)["total_launches"])
ctx.close()
