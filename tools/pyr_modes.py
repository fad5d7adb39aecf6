"""Times the Gaussian + DoG pyramid of one 4K (or WxH) frame in each pyramid mode (CUDA events
inside the library; octaves serialised and overlapped) and checks the modes agree bit for bit.

    python tools/pyr_modes.py [WxH] [mode ...]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sara_b200 as sb  # noqa: E402

size = next((a for a in sys.argv[1:] if "x" in a), "3840x2160")
W, H = (int(v) for v in size.split("x"))
modes = [a for a in sys.argv[1:] if "x" not in a] or ["stage", "march"]
rng = np.random.default_rng(3)
img = torch.from_numpy(rng.random((H, W), dtype=np.float32)).cuda()
pp = sb.ImagePyramidParams(first_octave_index=0)
ctx = sb.SiftContext(W, H, max_keypoints=65536, min_first_octave_index=0)
ctx.set_profiling(True)
ref, out = None, {}
for mode in modes:
    ctx.set_pyramid_mode(mode)
    res = {}
    for overlap in (False, True):
        ctx.set_octave_overlap(overlap)
        ts, top = [], []
        for r in range(8):
            ctx.pyramid_enqueue(0, img, pp)
            ctx.wait(0)
            if r >= 3:
                t = ctx.timings(0)
                ts.append(t["pyramid"])
                top.append(t["pyramid_top_kernel"])
        res["overlap" if overlap else "serial"] = {"pyramid_ms": float(np.mean(ts)), "min": float(np.min(ts)),
                                                   "top_kernel_ms": float(np.mean(top)),
                                                   "launches": ctx.timings(0)["pyramid_launches"]}
    layers = [ctx.dog_layer(s, o).tobytes() for o in range(ctx.num_octaves()) for s in (0, 2, 4)] + \
             [ctx.gaussian_layer(5, o).tobytes() for o in range(ctx.num_octaves())]
    if ref is None:
        ref = layers
    res["same_bits_as_first_mode"] = layers == ref
    out[mode] = res
print(json.dumps({"size": size, "modes": out}))
ctx.close()
