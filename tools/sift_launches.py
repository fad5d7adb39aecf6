"""One full-SIFT 4K frame per iteration on device-resident input (for ncu launch lists)."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
import sara_b200 as sb
from sara_b200 import synthetic as S
w, h = 3840, 2160
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = [torch.from_numpy(S.tex(w, h, 1234 + i)).cuda() for i in range(2)]
ctx = sb.SiftContext(w, h, max_keypoints=131072)
ctx.set_profiling(True)
pp = sb.ImagePyramidParams(first_octave_index=0)
for i in range(iters):
    ctx.enqueue(0, frames[i % 2], pp)
    n = ctx.wait(0)
    print(n, ctx.timings(0))
