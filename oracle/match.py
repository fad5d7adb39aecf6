"""ctypes loader for the matching oracle (oracle/match_oracle.cpp) and for the real FLANN of
the reference (oracle/_ref/libflann_ref.so, built from /root/reference's vendored sources).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under sara_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as _O

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_PATH = os.path.join(_HERE, "_ref", "libflann_ref.so")

MATCH_DTYPE = np.dtype([("x_index", "<i4"), ("y_index", "<i4"), ("rank", "<i4"), ("score", "<f4"), ("direction", "<i4")])

KNN_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float))
RADIUS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int)

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_ready = False


def _lib():
    global _ready
    L = _O.lib()
    if not _ready:
        L.oracle_l2_flann.restype = C.c_float
        L.oracle_l2_flann.argtypes = [_fp, _fp, C.c_int]
        L.oracle_knn_linear.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _ip, _fp]
        L.oracle_ann_match.restype = C.c_int
        L.oracle_ann_match.argtypes = [_fp, C.c_void_p, C.c_int, _fp, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int,
                                       C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int]
        _ready = True
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def l2_flann(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return float(_lib().oracle_l2_flann(a.ctypes.data_as(_fp), b.ctypes.data_as(_fp), a.size))


def knn_linear(data, queries, k: int = 3):
    """Exact k-NN with FLANN's L2 functor and result-set rules; returns (idx (nq, k) int32, dist (nq, k) float32)."""
    data, queries = _f32(data), _f32(queries)
    nq, dim = queries.shape
    idx = np.empty((nq, k), np.int32)
    dist = np.empty((nq, k), np.float32)
    _lib().oracle_knn_linear(data.ctypes.data_as(_fp), data.shape[0], dim, queries.ctypes.data_as(_fp), nq, k,
                             idx.ctypes.data_as(_ip), dist.ctypes.data_as(_fp))
    return idx, dist


def have_ref() -> bool:
    if not os.path.exists(_REF_PATH) and os.path.isdir("/root/reference/cpp/third-party/flann/src/cpp/flann"):
        subprocess.call(["make", "-C", _HERE, "-s", "ref"])
    return os.path.exists(_REF_PATH)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libflann_ref.so is not built (needs /root/reference)")
        R = C.CDLL(_REF_PATH)
        R.flannref_build.restype = C.c_void_p
        R.flannref_build.argtypes = [_fp, C.c_int, C.c_int, C.c_int]
        R.flannref_free.argtypes = [C.c_void_p]
        R.flannref_knn_batch.argtypes = [C.c_void_p, _fp, C.c_int, C.c_int, _ip, _fp]
        _ref = R
    return _ref


class FlannRef:
    """The reference's vendored FLANN: kind 'linear' (exact) or 'kdtree' (KDTreeIndexParams{8}, the
    index AnnMatcher builds).  Keeps the data array alive (FLANN does not copy it)."""

    def __init__(self, data, kind: str = "kdtree"):
        self.data = _f32(data)
        self.R = ref_lib()
        self.h = self.R.flannref_build(self.data.ctypes.data_as(_fp), self.data.shape[0], self.data.shape[1],
                                       0 if kind == "linear" else 1)

    def knn(self, queries, k: int = 3):
        q = _f32(queries)
        idx = np.empty((q.shape[0], k), np.int32)
        dist = np.empty((q.shape[0], k), np.float32)
        self.R.flannref_knn_batch(self.h, q.ctypes.data_as(_fp), q.shape[0], k, idx.ctypes.data_as(_ip),
                                  dist.ctypes.data_as(_fp))
        return idx, dist

    def close(self):
        if self.h:
            self.R.flannref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ann_match(desc1, desc2, sift_ratio_thres: float = 1.2, feat1=None, feat2=None, self_matching: bool = False,
              min_max_metric_dist_thres: float = 0.5, pixel_dist_thres: float = 10.0, backend: str = "port"):
    """AnnMatcher(keys1, keys2, sift_ratio_thres).compute_matches() (AnnMatcher.cpp:219-282).

    backend 'port': the restated exact search; 'linear' / 'kdtree': the reference's FLANN
    (oracle/_ref) underneath the same matching logic.  Returns a MATCH_DTYPE array."""
    d1, d2 = _f32(desc1), _f32(desc2)
    n1, n2 = d1.shape[0], d2.shape[0]
    dim = d1.shape[1] if d1.ndim == 2 else d2.shape[1]
    f1 = np.ascontiguousarray(feat1, dtype=_O.KEYPOINT_DTYPE) if feat1 is not None else None
    f2 = np.ascontiguousarray(feat2, dtype=_O.KEYPOINT_DTYPE) if feat2 is not None else None
    knn = radius = None
    i1 = i2 = None
    h1 = h2 = None
    if backend != "port" and n1 and n2:
        R = ref_lib()
        i1, i2 = FlannRef(d1, backend), FlannRef(d2, backend)
        knn = C.cast(R.flannref_knn, C.c_void_p)
        radius = C.cast(R.flannref_radius, C.c_void_p)
        h1, h2 = i1.h, i2.h
    cap = max(4 * (n1 + n2), 1024)
    while True:
        out = np.empty(cap, MATCH_DTYPE)
        n = _lib().oracle_ann_match(d1.ctypes.data_as(_fp), f1.ctypes.data if f1 is not None else None, n1,
                                    d2.ctypes.data_as(_fp), f2.ctypes.data if f2 is not None else None, n2, dim,
                                    sift_ratio_thres, int(self_matching), min_max_metric_dist_thres, pixel_dist_thres,
                                    knn, radius, h1, h2, out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("Error: the list of key-points is empty!")
        if n <= cap:
            return out[:n].copy()
        cap = n
