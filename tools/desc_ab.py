"""A/B aid: hashes of keypoints + descriptors of several frames, computed with the package found under the
directory given as argv[1] (e.g. `.` and `tools/_old`): identical hashes <=> bit-identical results."""
import hashlib, os, sys
import numpy as np
root = os.path.abspath(sys.argv[1])
sys.path.insert(0, root)
import sara_b200 as sb
from sara_b200 import synthetic as S
assert os.path.dirname(os.path.dirname(sb.__file__)) == root, sb.__file__
cases = [("tex", 3840, 2160, 0, 1.6), ("tex", 1920, 1080, -1, 1.6), ("tex", 1300, 420, 0, 1.0), ("tex", 640, 480, -1, 1.2)]
for name, w, h, fo, s0 in cases:
    img = S.tex(w, h, 1234)
    ctx = sb.SiftContext(w * (2 if fo < 0 else 1), h * (2 if fo < 0 else 1), max_keypoints=262144, min_first_octave_index=fo)
    pp = sb.ImagePyramidParams(first_octave_index=fo, scale_initial=s0)
    kl = ctx.compute_sift_keypoints(img, pp)
    hk = hashlib.sha256(np.ascontiguousarray(kl.features).tobytes()).hexdigest()[:16]
    hd = hashlib.sha256(np.ascontiguousarray(kl.descriptors).tobytes()).hexdigest()[:16]
    print(name, w, h, fo, s0, len(kl), hk, hd)
    ctx.close()
