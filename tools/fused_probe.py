import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import sara_b200 as sb
from sara_b200 import synthetic as S
from oracle import oracle as O
from parity import assert_pyramids_identical
w, h = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else 'auto'
img = S.tex(w, h, 7)
ctx = sb.SiftContext(max(w, 64), max(h, 64))
ctx.set_pyramid_mode(mode)
ctx.pyramid_enqueue(0, img, sb.ImagePyramidParams(first_octave_index=0))
ctx.wait(0)
ref = O.compute_dog_extrema(img, O.PyramidParams(first_octave_index=0))
for o in range(ref.num_octaves):
    for s in range(ref.num_scales):
        g, r = ctx.gaussian_layer(s, o), ref.gaussian(s, o)
        bad = np.argwhere(g != r)
        print("G", s, o, g.shape, "mismatch", len(bad), bad[:4].tolist(), bad[-2:].tolist() if len(bad) else "")
    for s in range(ref.num_scales - 1):
        g, r = ctx.dog_layer(s, o), ref.dog(s, o)
        bad = np.argwhere(g != r)
        print("D", s, o, "mismatch", len(bad), bad[:4].tolist())
