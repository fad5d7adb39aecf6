"""Generates tests/golden/stinkbug_gray.npy and stinkbug_oracle.npz (config C1).

Run in the build container only (needs /root/reference/data/stinkbug.png and PIL):
    python tests/golden/make_stinkbug_fixture.py
Grey conversion restates DO::Sara rgb8 -> gray32f (Core/Pixel/ColorConversion.hpp:27-33,
ChannelConversion.hpp:41-53): channel / 255 in double, 0.2125 R + 0.7154 G + 0.0721 B in
double, cast to float.  The oracle output stored beside it pins the oracle against itself
over time (the reference holds no golden keypoints).
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

rgb = np.asarray(Image.open("/root/reference/data/stinkbug.png").convert("RGB"), dtype=np.float64) / 255.0
gray = (0.2125 * rgb[..., 0] + 0.7154 * rgb[..., 1] + 0.0721 * rgb[..., 2]).astype(np.float32)
assert gray.shape == (375, 500)
np.save(os.path.join(HERE, "stinkbug_gray.npy"), gray)

r = O.compute_sift_keypoints(gray, O.PyramidParams(), parallel=True)  # default params, fo = -1
kp = r.keypoints
np.savez_compressed(
    os.path.join(HERE, "stinkbug_oracle.npz"),
    keypoints=kp, extrema=r.extrema, descriptors=r.descriptors.astype(np.float32),
    num_octaves=r.num_octaves)
print("stinkbug:", gray.shape, "octaves", r.num_octaves, "extrema", len(r.extrema), "keypoints", len(kp))
