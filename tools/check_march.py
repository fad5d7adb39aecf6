"""Debug aid: pyramid in mode `stage` vs `march`, every layer, reports where they differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sara_b200 as sb
from sara_b200 import synthetic as S

cases = [("tex", 3840, 2160), ("tex", 1920, 1080), ("rand", 3840, 2160), ("tex", 1300, 420)]
pp = sb.ImagePyramidParams(first_octave_index=0)
for name, w, h in cases:
    img = S.tex(w, h, 1234) if name == "tex" else np.random.default_rng(3).random((h, w), dtype=np.float32)
    ctx = sb.SiftContext(w, h, max_keypoints=65536, min_first_octave_index=0)
    for rep in range(3):
        layers = {}
        for mode in ("stage", "march"):
            ctx.set_pyramid_mode("stage")
            ctx.pyramid_enqueue(0, 1.0 - img, pp)  # poison the arena: stale rows must not look right
            ctx.wait(0)
            ctx.set_pyramid_mode(mode)
            ctx.pyramid_enqueue(0, img, pp)
            ctx.wait(0)
            layers[mode] = [[ctx.gaussian_layer(s, o) for s in range(6)] for o in range(ctx.num_octaves())]
        bad = 0
        for o in range(len(layers["stage"])):
            for s in range(6):
                a, b = layers["stage"][o][s], layers["march"][o][s]
                if a.tobytes() != b.tobytes():
                    d = np.argwhere(a != b)
                    ys, xs = d[:, 0], d[:, 1]
                    print(f"{name} {w}x{h} rep {rep}: G({s},{o}) differs at {len(d)} px; rows {ys.min()}..{ys.max()} "
                          f"cols {xs.min()}..{xs.max()}; distinct rows {len(set(ys))} distinct cols {len(set(xs))}; "
                          f"first {d[:3].tolist()}")
                    bad += 1
                    break
            if bad:
                break
        print(f"{name} {w}x{h} rep {rep}: {'OK' if not bad else 'MISMATCH'}")
    ctx.close()
