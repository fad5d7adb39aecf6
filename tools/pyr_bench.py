"""Times the Gaussian pyramid + DoG stage alone on device-resident synthetic frames.
    python tools/pyr_bench.py [w h [iters]]"""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
import sara_b200 as sb
from sara_b200 import synthetic as S

w = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
h = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mode = sys.argv[4] if len(sys.argv) > 4 else "auto"
rng = np.random.default_rng(0)
frames = [torch.from_numpy(rng.random((h, w), dtype=np.float32)).cuda() for _ in range(4)]
ctx = sb.SiftContext(w, h)
pp = sb.ImagePyramidParams(first_octave_index=0)
ctx.set_profiling(True)
ctx.set_pyramid_mode(mode)
ts = []
for i in range(iters + 3):
    ctx.pyramid_enqueue(0, frames[i % 4], pp)
    ctx.wait(0)
    if i >= 3:
        ts.append(ctx.timings(0)["pyramid"])
px = 0
ww, hh = w, h
for o in range(ctx.num_octaves()):
    px += ww * hh; ww //= 2; hh //= 2
ms = float(np.median(ts))
print(f"[{mode}] {w}x{h}: pyramid+DoG median {ms*1e3:.1f} us over {iters} frames (min {min(ts)*1e3:.1f}); "
      f"{48*px/ms/1e6:.0f} GB/s algorithmic ({48*px/1e6:.1f} MB), octaves {ctx.num_octaves()}")
