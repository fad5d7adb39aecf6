"""Pins the oracle's Eigen-boundary restatements against an INDEPENDENT float64 solver.

The reference calls Eigen 3.4 (absent here) for two things inside refine_extremum
(FeatureDetectors/RefineExtremum.cpp:74-85):
  * SelfAdjointEigenSolver<Matrix3f>: only the sign pattern of the eigenvalues is used
    (`(lambda * type).maxCoeff() >= 0` => do not refine);
  * Matrix3f::inverse(): h = -H^-1 g.
The oracle restates them as a cyclic Jacobi iteration and the cofactor inverse in fp32, and
the CUDA kernel repeats the same operations bit for bit -- so a mistake shared by both would
pass every GPU-vs-oracle test.  Here every Newton iteration of every candidate of config C1
(stinkbug, default parameters) and of a 4K frame (config C3) is re-decided with numpy's LAPACK
routines in float64: no decision may flip, eigenvalues must agree to fp32 accuracy and the
Newton offsets to the accuracy the conditioning of H allows."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from sara_b200 import synthetic as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _trace(img, pp, pad):
    O.set_threading(1, 0)
    O.trace_refinement(True)
    try:
        res = O.compute_dog_extrema(img, pp, 4.0, 0.01, 10.0, pad, 5)
        return O.refinement_trace(), res
    finally:
        O.trace_refinement(False)


@pytest.mark.parametrize("case", ["c1_stinkbug", "c3_4k", "c4_1080p_sequence"])
def test_eigen_sign_and_inverse_against_float64(case):
    if case == "c1_stinkbug":
        img, pp = np.load(os.path.join(GOLDEN, "stinkbug_gray.npy")), O.PyramidParams()
    elif case == "c3_4k":
        img, pp = S.tex(3840, 2160, 1234), O.PyramidParams(first_octave_index=0)
    else:
        img, pp = S.sequence_frame(1920, 1080, 2), O.PyramidParams(first_octave_index=0)
    tr, res = _trace(img, pp, 5)  # padding 5 = what compute_sift_keypoints passes (quirk N1)
    assert len(tr) >= len(res.extrema) > 100
    H = tr[:, 6:15].reshape(-1, 3, 3).astype(np.float64)
    g = tr[:, 15:18].astype(np.float64)
    typ = tr[:, 4].astype(np.float64)  # 1 or 255 (uint8 map value, quirk N2)
    assert set(np.unique(typ)) <= {1.0, 255.0}
    assert np.array_equal(H, H.transpose(0, 2, 1))

    # (1) eigenvalues and the sign decision, RefineExtremum.cpp:74-81
    lam64 = np.linalg.eigvalsh(H)
    lam32 = np.sort(tr[:, 18:21].astype(np.float64), axis=1)
    scale = np.abs(lam64).max(axis=1, keepdims=True)
    assert (np.abs(lam32 - lam64) / scale).max() < 2e-6
    newton64 = (lam64 * typ[:, None]).max(axis=1) < 0
    newton32 = tr[:, 24] == 1
    flips = int((newton32 != newton64).sum())
    assert flips == 0, f"{flips} of {len(tr)} definiteness decisions differ from float64"
    # minima (type 255) are only refined when H is NEGATIVE definite -- i.e. practically never
    assert newton32[typ == 255].mean() < 0.01

    # (2) h = -H^-1 g, RefineExtremum.cpp:85
    m = newton32
    assert m.sum() > 50
    h64 = -np.linalg.solve(H[m], g[m][:, :, None])[:, :, 0]
    h32 = tr[m, 21:24].astype(np.float64)
    err = np.abs(h64 - h32).max(axis=1)
    cond = np.linalg.cond(H[m])
    hmag = np.maximum(np.abs(h64).max(axis=1), 1.0)
    # fp32 cofactor inverse: error <= a few eps32 * cond(H) * |h|
    assert (err <= 4 * 1.2e-7 * cond * hmag + 1e-6).all(), float((err / (cond * hmag)).max())
    assert (err[cond < 50] < 1e-4).all() and np.mean(err < 1e-4) >= 0.995
    # and the decisions taken on h (|h_xy| > 1.5 reject, > 0.6 step) do not flip either
    hxy64, hxy32 = np.abs(h64[:, :2]), np.abs(h32[:, :2])
    assert np.array_equal(hxy64.max(axis=1) > 1.5, hxy32.max(axis=1) > 1.5)
    assert np.array_equal(hxy64.min(axis=1) > 0.6, hxy32.min(axis=1) > 0.6)
