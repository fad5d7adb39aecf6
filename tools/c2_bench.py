"""Config C2 of BASELINE.json: 1920x1080 synthetic gradient image, 4 octaves, pyramid + DoG only;
GPU (CUDA events) and the CPU oracle (reference threading = serial pyramid, and all threads)."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
import sara_b200 as sb
from sara_b200 import synthetic as S
from oracle import oracle as O

img = S.grad(1920, 1080)
d = torch.from_numpy(img).cuda()
ctx = sb.SiftContext(1920, 1080)
ctx.set_profiling(True)
pp = sb.ImagePyramidParams(first_octave_index=0, num_octaves_max=4)
ts = []
for i in range(13):
    ctx.pyramid_enqueue(0, d, pp); ctx.wait(0)
    if i >= 3: ts.append(ctx.timings(0)["pyramid"])
px = sum((1920 >> o) * (1080 >> o) for o in range(4))
ms = float(np.median(ts))
print(f"C2 GPU: pyramid+DoG {ms*1e3:.1f} us, {48*px/ms/1e6:.0f} GB/s algorithmic ({48*px/1e6:.2f} MB)")
for mode, name in ((0, "reference threading"), (1, "all threads")):
    O.set_threading(mode, 0)
    t = []
    for i in range(4):
        t0 = time.perf_counter(); O.compute_dog_extrema(img, O.PyramidParams(first_octave_index=0, num_octaves_max=4)); t.append(time.perf_counter() - t0)
    print(f"C2 CPU ({name}, {O.num_threads()} threads): pyramid+DoG+extrema {1e3*np.median(t[1:]):.1f} ms")
