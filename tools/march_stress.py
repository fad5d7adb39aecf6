"""Debug aid: the march pyramid repeated many times against the stage pyramid; reports where layers differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sara_b200 as sb
from sara_b200 import synthetic as S

size = next((a for a in sys.argv[1:] if "x" in a), "3840x2160")
w, h = (int(v) for v in size.split("x"))
reps = next((int(a) for a in sys.argv[1:] if a.isdigit()), 30)
pp = sb.ImagePyramidParams(first_octave_index=0)
img = torch.from_numpy(S.tex(w, h, 1234)).cuda()
poison = torch.from_numpy(1.0 - S.tex(w, h, 1234)).cuda()
ctx = sb.SiftContext(w, h, max_keypoints=65536, min_first_octave_index=0)
ctx.set_pyramid_mode("stage")
ctx.pyramid_enqueue(0, img, pp); ctx.wait(0)
no = ctx.num_octaves()
ref = [[(ctx.gaussian_layer(s, o), ctx.dog_layer(s, o) if s < 5 else None) for s in range(6)] for o in range(no)]
nbad = 0
for rep in range(reps):
    ctx.set_pyramid_mode("stage"); ctx.set_octave_overlap(True)
    ctx.pyramid_enqueue(0, poison, pp); ctx.wait(0)
    ctx.set_pyramid_mode("march"); ctx.set_octave_overlap(rep % 2 == 1)
    ctx.pyramid_enqueue(0, img, pp); ctx.wait(0)
    for o in range(no):
        for s in range(6):
            for name, a, b in (("G", ref[o][s][0], ctx.gaussian_layer(s, o)),) + ((("D", ref[o][s][1], ctx.dog_layer(s, o)),) if s < 5 else ()):
                if a.tobytes() != b.tobytes():
                    d = np.argwhere(a != b); ys, xs = d[:, 0], d[:, 1]
                    print(f"rep {rep} overlap {rep % 2}: {name}({s},{o}) {a.shape} differs at {len(d)} px; rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}; "
                          f"distinct rows {len(set(ys))} cols {len(set(xs))}; first {d[:4].tolist()}")
                    nbad += 1
                    x0 = (xs.min() // 248) * 248
                    print("   cols-x0:", sorted(set((xs - x0).tolist())))
                    print("   rows:", sorted(set(ys.tolist())), "per-row counts", [int((ys == r).sum()) for r in sorted(set(ys.tolist()))])
                    break
            else:
                continue
            break
        else:
            continue
        break
print("reps", reps, "mismatching layers", nbad)
ctx.close()
