"""Keypoint wire format and de-duplication, host side.

Mirrors, on the arrays the C ABI returns (KEYPOINT_DTYPE records + N x 128 float32):
  * write_keypoints / read_keypoints   cpp/src/DO/Sara/Features/IO.hpp:78-134 (text format)
  * remove_redundant_features          cpp/src/DO/Sara/Features/Utilities.cpp:23-82
The HDF5 variants (IO.hpp:139-164) store the same two arrays as datasets `<group>/features`
and `<group>/descriptors`; HDF5 is not available in this image, so only the text format is
provided here.

Text format (one keypoint per line after the "N dim" header):
    x y <shape_matrix in memory order, 4 floats> orientation int(type) <descriptor, dim floats>
Numbers are printed like a C++ ostream at its default precision (6 significant digits, %g);
the two Eigen row vectors are printed with Eigen's default IOFormat, which right-aligns every
coefficient of a vector to the width of its widest coefficient.
"""
from __future__ import annotations

import functools

import numpy as np

from .api import KEYPOINT_DTYPE, KeypointList


def _g(v) -> str:
    return "%g" % float(v)


def _eigen_row(values) -> str:
    strs = [_g(v) for v in values]
    width = max(len(s) for s in strs) if strs else 0
    return " ".join(s.rjust(width) for s in strs)


def write_keypoints(keys: KeypointList, path: str) -> bool:
    """write_keypoints(features, descriptors, name), IO.hpp:107-134."""
    feats, desc = keys.features, np.asarray(keys.descriptors, dtype=np.float32)
    try:
        f = open(path, "w")
    except OSError:
        return False
    with f:
        f.write(f"{len(feats)} {desc.shape[1] if desc.ndim == 2 else 0}\n")
        for i in range(len(feats)):
            k = feats[i]
            f.write(f"{_g(k['x'])} {_g(k['y'])} {_eigen_row(k['shape'])} {_g(k['orientation'])} {int(k['type'])} "
                    f"{_eigen_row(desc[i])}\n")
    return True


def read_keypoints(path: str) -> KeypointList | None:
    """read_keypoints(features, descriptors, name), IO.hpp:78-105.  The shape matrix is read
    element by element, rows first (Core/EigenExtension.hpp:162-170), into a column-major
    matrix; the other fields of OERegion keep their defaults (extremum_value 0, types undefined)."""
    try:
        tokens = open(path).read().split()
    except OSError:
        return None
    n, dim = int(tokens[0]), int(tokens[1])
    per = 2 + 4 + 1 + 1 + dim
    body = np.asarray(tokens[2: 2 + n * per], dtype=np.float64).reshape(n, per)
    feats = np.zeros(n, KEYPOINT_DTYPE)
    feats["x"], feats["y"] = body[:, 0], body[:, 1]
    m = body[:, 2:6]  # m(0,0) m(0,1) m(1,0) m(1,1) -> column-major storage (0,0) (1,0) (0,1) (1,1)
    feats["shape"] = np.stack([m[:, 0], m[:, 2], m[:, 1], m[:, 3]], axis=1)
    feats["orientation"] = body[:, 6]
    feats["type"] = body[:, 7].astype(np.uint8)
    feats["extremum_type"] = -2  # ExtremumType::Undefined
    return KeypointList(feats, body[:, 8:].astype(np.float32))


def remove_redundant_features(keys: KeypointList) -> KeypointList:
    """remove_redundant_features, Utilities.cpp:23-82: sort the keypoints lexicographically by
    descriptor (ties: larger extremum_value first), then drop every keypoint whose descriptor is
    within 1e-6 (squared L2) of the last one kept."""
    feats, desc = keys.features, np.asarray(keys.descriptors, dtype=np.float32)
    if len(feats) != len(desc):
        raise RuntimeError("Fatal: the number of features and descriptors are not equal")
    n = len(feats)
    if n == 0:
        return KeypointList(feats.copy(), desc.copy())

    def equal(i, j):
        d = desc[i] - desc[j]
        return float(np.dot(d, d)) < 1e-6

    def less(i, j):  # compare_less
        a, b = desc[i], desc[j]
        ne = np.nonzero(a != b)[0]
        if len(ne) and a[ne[0]] < b[ne[0]]:
            return True
        return equal(i, j) and feats["extremum_value"][i] > feats["extremum_value"][j]

    # std::sort with compare_less; a stable lexicographic pre-sort keeps the Python comparisons few
    pre = np.lexsort(desc.T[::-1])
    order = sorted(pre.tolist(), key=functools.cmp_to_key(lambda i, j: -1 if less(i, j) else (1 if less(j, i) else 0)))
    kept = [order[0]]
    for i in order[1:]:  # std::unique with compare_equal
        if not equal(kept[-1], i):
            kept.append(i)
    kept = np.asarray(kept, dtype=np.int64)
    return KeypointList(feats[kept].copy(), desc[kept].copy())
