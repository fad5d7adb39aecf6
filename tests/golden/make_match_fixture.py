"""Golden vectors of the matching row, produced by the REFERENCE's own search library.

    python tests/golden/make_match_fixture.py        (in the build container: needs /root/reference)

Inputs: SIFT descriptors of two frames of the synthetic translated-scene sequence (oracle SIFT,
640 x 480).  Outputs, all computed by the reference's vendored FLANN (oracle/_ref/libflann_ref.so,
compiled from /root/reference/cpp/third-party/flann by oracle/Makefile):
  * lin_idx / lin_dist   : flann::Index<L2<float>>(LinearIndexParams).knnSearch(k = 3), both directions
  * kd_idx / kd_dist     : the same through KDTreeIndexParams{8} (what AnnMatcher.cpp:228-237 builds)
  * matches_*            : AnnMatcher::compute_matches logic (oracle/match_oracle.cpp) over the real
                           LinearIndex for ratios 0.6, 1.0 and 1.2 (radius branch)
The fixture travels to the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import match as M, oracle as O  # noqa: E402
from sara_b200 import synthetic as S  # noqa: E402

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "match_flann.npz")
frames = [S.sequence_frame(640, 480, i) for i in (0, 2)]
res = [O.compute_sift_keypoints(f, O.PyramidParams(first_octave_index=0), parallel=True) for f in frames]
d1, d2 = (np.ascontiguousarray(r.descriptors, np.float32) for r in res)
f1, f2 = (np.ascontiguousarray(r.keypoints) for r in res)
print("descriptors", d1.shape, d2.shape)

lin1, lin2 = M.FlannRef(d1, "linear"), M.FlannRef(d2, "linear")
kd1, kd2 = M.FlannRef(d1, "kdtree"), M.FlannRef(d2, "kdtree")
data = dict(d1=d1, d2=d2, f1=f1, f2=f2)
data["lin_idx_12"], data["lin_dist_12"] = lin2.knn(d1, 3)
data["lin_idx_21"], data["lin_dist_21"] = lin1.knn(d2, 3)
data["kd_idx_12"], data["kd_dist_12"] = kd2.knn(d1, 3)
data["kd_idx_21"], data["kd_dist_21"] = kd1.knn(d2, 3)
for thr in (0.6, 1.0, 1.2):
    data[f"matches_lin_{thr}"] = M.ann_match(d1, d2, thr, f1, f2, backend="linear")
    data[f"matches_kd_{thr}"] = M.ann_match(d1, d2, thr, f1, f2, backend="kdtree")
    print(thr, len(data[f"matches_lin_{thr}"]), len(data[f"matches_kd_{thr}"]))
data["self_matches_lin_1.2"] = M.ann_match(d1, d1, 1.2, f1, f1, self_matching=True, backend="linear")
print("self", len(data["self_matches_lin_1.2"]))
print("kd-tree recall of the nearest neighbour:", float((data["kd_idx_12"][:, 0] == data["lin_idx_12"][:, 0]).mean()))
np.savez_compressed(out, **data)
print("wrote", out, os.path.getsize(out), "bytes")
