"""ctypes binding of include/sara_b200.h and the pysara-shaped host API.

Reference interfaces mirrored here (paths relative to /root/reference):
  * pysara.ImagePyramidParams        python/oddkiva/sara/pybind11/FeatureDetectors.cpp:70-88
  * pysara.compute_sift_keypoints    python/oddkiva/sara/pybind11/FeatureDetectors.cpp:116-124
  * pysara.features / descriptors    python/oddkiva/sara/pybind11/FeatureDetectors.cpp:57-68
  * DO::Sara::ComputeDoGExtrema      cpp/src/DO/Sara/FeatureDetectors/DoG.hpp:72-165
Error behaviour follows the reference: bad sizes raise ValueError
(std::domain_error / std::range_error), fewer than 4 scales raises RuntimeError
(DoG.hpp:86-89).  This module never touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KEYPOINT_DTYPE = np.dtype(
    [
        ("x", "<f4"), ("y", "<f4"), ("shape", "<f4", (4,)), ("orientation", "<f4"),
        ("extremum_value", "<f4"), ("type", "u1"), ("extremum_type", "i1"), ("reserved", "<i2"),
        ("s", "<i4"), ("o", "<i4"), ("xi", "<i4"), ("yi", "<i4"),
    ]
)
assert KEYPOINT_DTYPE.itemsize == 52


class SaraB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sara_b200 error {code}: {msg}")
        self.code = code


class _PyrParams(C.Structure):
    _fields_ = [
        ("first_octave_index", C.c_int32), ("scale_count_per_octave", C.c_int32),
        ("scale_geometric_factor", C.c_float), ("image_padding_size", C.c_int32),
        ("scale_camera", C.c_float), ("scale_initial", C.c_float), ("num_octaves_max", C.c_int32),
    ]


class _Limits(C.Structure):
    _fields_ = [
        ("max_width", C.c_int32), ("max_height", C.c_int32), ("max_keypoints", C.c_int32),
        ("num_slots", C.c_int32), ("min_first_octave_index", C.c_int32),
    ]


class _SiftArgs(C.Structure):
    _fields_ = [
        ("pyramid_params", _PyrParams), ("gauss_truncate", C.c_float), ("extremum_thres", C.c_float),
        ("edge_ratio_thres", C.c_float), ("extremum_refinement_iter", C.c_int32),
    ]


class _DogArgs(C.Structure):
    _fields_ = [
        ("pyramid_params", _PyrParams), ("gauss_truncate", C.c_float), ("extremum_thres", C.c_float),
        ("edge_ratio_thres", C.c_float), ("img_padding_sz", C.c_int32),
        ("extremum_refinement_iter", C.c_int32),
    ]


class _MatchArgs(C.Structure):
    _fields_ = [
        ("sift_ratio_thres", C.c_float), ("self_matching", C.c_int32), ("min_max_metric_dist_thres", C.c_float),
        ("pixel_dist_thres", C.c_float), ("knn_mode", C.c_int32),
    ]


class KnnStats(C.Structure):
    _fields_ = [
        ("used_tensor_cores", C.c_int32), ("n_redone", C.c_int32), ("launches", C.c_int32), ("splits", C.c_int32),
        ("gpu_ms", C.c_float),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


MATCH_DTYPE = np.dtype([("x_index", "<i4"), ("y_index", "<i4"), ("rank", "<i4"), ("score", "<f4"), ("direction", "<i4")])
KNN_MODES = {"auto": 0, "scalar": 1, "tensor": 2}


class Timings(C.Structure):
    _fields_ = [
        ("upload", C.c_float), ("pyramid", C.c_float), ("extrema", C.c_float),
        ("orientation", C.c_float), ("descriptor", C.c_float), ("total", C.c_float),
        ("pyramid_launches", C.c_int32), ("total_launches", C.c_int32),
        ("pyramid_top_kernel", C.c_float), ("pyramid_top_kernel_mbytes", C.c_float),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# Every symbol include/sara_b200.h declares (tests check the library exports all).
SYMBOLS = [
    "sara_b200_version", "sara_b200_last_error", "sara_b200_default_pyramid_params",
    "sara_b200_default_sift_args", "sara_b200_default_dog_args", "sara_b200_create",
    "sara_b200_destroy", "sara_b200_host_alloc", "sara_b200_host_free", "sara_b200_set_profiling",
    "sara_b200_set_pyramid_mode", "sara_b200_set_octave_overlap",
    "sara_b200_last_timings", "sara_b200_sift", "sara_b200_sift_enqueue", "sara_b200_collect",
    "sara_b200_device_results", "sara_b200_wait", "sara_b200_dog_extrema",
    "sara_b200_pyramid_enqueue", "sara_b200_num_octaves", "sara_b200_num_scales",
    "sara_b200_layer_size", "sara_b200_octave_scaling_factor", "sara_b200_copy_layer",
    "sara_b200_copy_extrema", "sara_b200_copy_oriented", "sara_b200_gaussian",
    "sara_b200_make_gaussian_kernel", "sara_b200_sift_u8", "sara_b200_sift_enqueue_u8", "sara_b200_to_gray32f",
    "sara_b200_collect_device", "sara_b200_set_graphs",
    "sara_b200_default_match_args", "sara_b200_knn", "sara_b200_compute_matches",
    "sara_b200_log_extrema", "sara_b200_doh_extrema", "sara_b200_describe_extrema", "sara_b200_hessian_laplace",
    "sara_b200_harris_laplace",
]


def library_path() -> str:
    return os.path.join(_HERE, "libsara_b200.so")


def load_library() -> C.CDLL:
    """Loads (building first when the sources are newer) the C-ABI library.

    Fails loudly when it is missing: there is no other implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    from . import build as _build

    path = _build.build()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m sara_b200.build`")
    L = C.CDLL(path)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.sara_b200_last_error.restype = C.c_char_p
    L.sara_b200_last_error.argtypes = [vp]
    L.sara_b200_create.argtypes = [C.c_int, C.POINTER(_Limits), C.POINTER(vp)]
    L.sara_b200_destroy.argtypes = [vp]
    L.sara_b200_destroy.restype = None
    L.sara_b200_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.sara_b200_host_free.argtypes = [vp]
    L.sara_b200_host_free.restype = None
    L.sara_b200_set_profiling.argtypes = [vp, C.c_int]
    L.sara_b200_set_pyramid_mode.argtypes = [vp, C.c_int]
    L.sara_b200_set_octave_overlap.argtypes = [vp, C.c_int]
    L.sara_b200_set_graphs.argtypes = [vp, C.c_int]
    L.sara_b200_last_timings.argtypes = [vp, C.c_int, C.POINTER(Timings)]
    L.sara_b200_sift.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(_SiftArgs), vp, vp, C.c_int, ip]
    L.sara_b200_sift_enqueue.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(_SiftArgs), vp]
    L.sara_b200_sift_u8.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_SiftArgs), vp, vp, C.c_int, ip]
    L.sara_b200_sift_enqueue_u8.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_SiftArgs), vp]
    L.sara_b200_to_gray32f.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.sara_b200_collect.argtypes = [vp, C.c_int, vp, vp, C.c_int, ip]
    L.sara_b200_collect_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, ip]
    L.sara_b200_device_results.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), ip]
    L.sara_b200_wait.argtypes = [vp, C.c_int, ip]
    L.sara_b200_dog_extrema.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(_DogArgs)]
    L.sara_b200_log_extrema.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(_DogArgs)]
    L.sara_b200_doh_extrema.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(_DogArgs)]
    L.sara_b200_hessian_laplace.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(_DogArgs), C.c_int]
    L.sara_b200_harris_laplace.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(_DogArgs), C.c_float,
                                           C.c_int]
    L.sara_b200_describe_extrema.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, ip]
    L.sara_b200_pyramid_enqueue.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(_PyrParams), C.c_float, vp]
    L.sara_b200_num_octaves.argtypes = [vp, C.c_int]
    L.sara_b200_num_scales.argtypes = [vp, C.c_int]
    L.sara_b200_layer_size.argtypes = [vp, C.c_int, C.c_int, ip, ip]
    L.sara_b200_octave_scaling_factor.argtypes = [vp, C.c_int, C.c_int]
    L.sara_b200_octave_scaling_factor.restype = C.c_float
    L.sara_b200_copy_layer.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.sara_b200_copy_extrema.argtypes = [vp, C.c_int, vp, C.c_int, ip]
    L.sara_b200_copy_oriented.argtypes = [vp, C.c_int, vp, C.c_int, ip]
    L.sara_b200_gaussian.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, vp]
    L.sara_b200_make_gaussian_kernel.argtypes = [C.c_float, C.c_float, fp, C.c_int]
    L.sara_b200_default_pyramid_params.argtypes = [C.POINTER(_PyrParams)]
    L.sara_b200_default_sift_args.argtypes = [C.POINTER(_SiftArgs)]
    L.sara_b200_default_dog_args.argtypes = [C.POINTER(_DogArgs)]
    L.sara_b200_default_match_args.argtypes = [C.POINTER(_MatchArgs)]
    L.sara_b200_default_match_args.restype = None
    L.sara_b200_knn.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                C.POINTER(KnnStats)]
    L.sara_b200_compute_matches.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(_MatchArgs), vp, C.c_int, ip, C.POINTER(KnnStats)]
    _LIB = L
    return L


@dataclass
class ImagePyramidParams:
    """ImagePyramidParams (cpp/src/DO/Sara/ImageProcessing/ImagePyramid.hpp:33-49).

    Defaults are the C++ ones (first_octave_index = -1); the pybind11 wrapper
    defaults first_octave_index to 1 (FeatureDetectors.cpp:72), callers of the
    SfM path pass 0 (SfM/BuildingBlocks/FeatureParams.hpp:10)."""

    first_octave_index: int = -1
    scale_count_per_octave: int = 6
    scale_geometric_factor: float = float(np.float32(2.0) ** np.float32(1.0 / 3.0))
    image_padding_size: int = 1
    scale_camera: float = 0.5
    scale_initial: float = 1.6
    num_octaves_max: int = 2**31 - 1

    def _c(self) -> _PyrParams:
        return _PyrParams(
            int(self.first_octave_index), int(self.scale_count_per_octave),
            float(self.scale_geometric_factor), int(self.image_padding_size),
            float(self.scale_camera), float(self.scale_initial), int(self.num_octaves_max))


class KeypointList:
    """KeypointList<OERegion, float> = (features, descriptors)
    (cpp/src/DO/Sara/Features/KeypointList.hpp:35-36)."""

    def __init__(self, feats: np.ndarray, descs: np.ndarray):
        self.features = feats
        self.descriptors = descs

    def __len__(self):
        return len(self.features)


def features(kl: KeypointList) -> np.ndarray:
    return kl.features


def descriptors(kl: KeypointList) -> np.ndarray:
    return kl.descriptors


def _raise(L, ctx, rc: int):
    msg = L.sara_b200_last_error(ctx).decode(errors="replace")
    if rc == -1:
        raise ValueError(msg)  # std::domain_error / range_error
    if rc == -2:
        raise RuntimeError(msg)  # DoG.hpp:86-89
    raise SaraB200Error(rc, msg)


def _as_image(image):
    """Returns (pointer, w, h, on_device, keepalive)."""
    if hasattr(image, "data_ptr"):  # torch tensor
        t = image
        if t.dim() != 2 or str(t.dtype) != "torch.float32" or not t.is_contiguous():
            raise ValueError("image tensor must be 2-D contiguous float32")
        return t.data_ptr(), int(t.shape[1]), int(t.shape[0]), bool(t.is_cuda), t
    a = np.ascontiguousarray(image, dtype=np.float32)
    if a.ndim != 2:
        raise ValueError("image must be HxW float32")
    return a.ctypes.data, int(a.shape[1]), int(a.shape[0]), False, a


def _as_u8_image(image):
    """Returns (pointer, w, h, channels, on_device, keepalive) of an 8-bit frame."""
    if hasattr(image, "data_ptr"):  # torch tensor
        t = image
        if str(t.dtype) != "torch.uint8" or not t.is_contiguous() or t.dim() not in (2, 3) or \
                (t.dim() == 3 and t.shape[2] != 3):
            raise ValueError("8-bit frame tensor must be contiguous uint8, HxW or HxWx3")
        return t.data_ptr(), int(t.shape[1]), int(t.shape[0]), 3 if t.dim() == 3 else 1, bool(t.is_cuda), t
    a = np.ascontiguousarray(image, dtype=np.uint8)
    if a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] != 3):
        raise ValueError("8-bit frame must be HxW or HxWx3 uint8")
    return a.ctypes.data, int(a.shape[1]), int(a.shape[0]), 3 if a.ndim == 3 else 1, False, a


class SiftContext:
    """One GPU context (see the threading contract in include/sara_b200.h)."""

    def __init__(self, max_width: int, max_height: int, device: int = 0, max_keypoints: int = 262144,
                 num_slots: int = 1, min_first_octave_index: int = -1):
        self._L = load_library()
        self._ctx = C.c_void_p()
        lim = _Limits(max_width, max_height, max_keypoints, num_slots, min_first_octave_index)
        rc = self._L.sara_b200_create(device, C.byref(lim), C.byref(self._ctx))
        if rc != 0:
            _raise(self._L, None, rc)
        self.max_keypoints = max_keypoints
        self.num_slots = num_slots
        self._keep = {}

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.sara_b200_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            _raise(self._L, self._ctx, rc)

    # ---- descriptor matching (AnnMatcher) ----------------------------------------
    @staticmethod
    def _as_desc(d):
        """(pointer, n, dim, on_device, keepalive) of an n x dim float32 matrix (numpy or torch)."""
        if hasattr(d, "data_ptr"):
            if d.dim() != 2 or str(d.dtype) != "torch.float32" or not d.is_contiguous():
                raise ValueError("descriptor tensor must be 2-D contiguous float32")
            return d.data_ptr(), int(d.shape[0]), int(d.shape[1]), bool(d.is_cuda), d
        a = np.ascontiguousarray(d, dtype=np.float32)
        if a.ndim != 2:
            raise ValueError("descriptors must be an n x dim float32 matrix")
        return a.ctypes.data, int(a.shape[0]), int(a.shape[1]), False, a

    def knn(self, queries, data, k: int = 3, mode="auto"):
        """flann::Index::knnSearch for every query row, exact (see include/sara_b200.h).
        Returns (idx (nq, k) int32, dist (nq, k) float32, stats dict)."""
        qp, nq, dim, q_dev, qk = self._as_desc(queries)
        dp, nd, dim2, d_dev, dk = self._as_desc(data)
        if dim != dim2 or q_dev != d_dev:
            raise ValueError("queries and data must have the same dimension and live on the same side")
        idx = np.empty((nq, k), np.int32)
        dist = np.empty((nq, k), np.float32)
        st = KnnStats()
        self._check(self._L.sara_b200_knn(self._ctx, qp, nq, dp, nd, dim, k, int(q_dev), KNN_MODES.get(mode, mode),
                                          idx.ctypes.data, dist.ctypes.data, C.byref(st)))
        return idx, dist, st.asdict()

    def compute_matches(self, desc1, desc2, sift_ratio_thres: float = 1.2, feat1=None, feat2=None,
                        self_matching: bool = False, min_max_metric_dist_thres: float = 0.5,
                        pixel_dist_thres: float = 10.0, mode="auto", return_stats: bool = False):
        """AnnMatcher(keys1, keys2, sift_ratio_thres).compute_matches() (AnnMatcher.cpp:219-282)."""
        p1, n1, dim, dev1, k1 = self._as_desc(desc1)
        p2, n2, dim2, dev2, k2 = self._as_desc(desc2)
        if n1 and n2 and (dim != dim2 or dev1 != dev2):
            raise ValueError("the two descriptor matrices must have the same dimension and live on the same side")
        f1 = np.ascontiguousarray(feat1, dtype=KEYPOINT_DTYPE) if feat1 is not None else None
        f2 = np.ascontiguousarray(feat2, dtype=KEYPOINT_DTYPE) if feat2 is not None else None
        args = _MatchArgs(sift_ratio_thres, int(self_matching), min_max_metric_dist_thres, pixel_dist_thres,
                          KNN_MODES.get(mode, mode))
        cap = max(4 * (n1 + n2), 1024)
        st = KnnStats()
        while True:
            out = np.empty(cap, MATCH_DTYPE)
            n = C.c_int(0)
            rc = self._L.sara_b200_compute_matches(self._ctx, p1, f1.ctypes.data if f1 is not None else None, n1, p2,
                                                   f2.ctypes.data if f2 is not None else None, n2, max(dim, 1),
                                                   int(dev1), C.byref(args), out.ctypes.data, cap, C.byref(n),
                                                   C.byref(st))
            if rc == -5 and n.value > cap:  # OVERFLOW: the full count came back
                cap = n.value
                continue
            self._check(rc)
            res = out[: n.value].copy()
            return (res, st.asdict()) if return_stats else res

    # ---- compute_sift_keypoints ------------------------------------------------
    @staticmethod
    def _sift_args(pp, gauss_truncate, extremum_thres, edge_ratio_thres, extremum_refinement_iter):
        return _SiftArgs((pp or ImagePyramidParams())._c(), gauss_truncate, extremum_thres,
                         edge_ratio_thres, int(extremum_refinement_iter))

    def set_pyramid_mode(self, mode):
        """0 auto, 1 generic, 2 gather-form stage kernel, 3 fused octave kernel, 4 scatter-form
        marching kernel (same bits)."""
        mode = {"auto": 0, "generic": 1, "stage": 2, "fused": 3, "march": 4}.get(mode, mode)
        self._check(self._L.sara_b200_set_pyramid_mode(self._ctx, int(mode)))

    def set_octave_overlap(self, on: bool):
        self._check(self._L.sara_b200_set_octave_overlap(self._ctx, int(on)))

    def set_graphs(self, on: bool):
        self._check(self._L.sara_b200_set_graphs(self._ctx, int(on)))

    def set_profiling(self, on: bool):
        self._check(self._L.sara_b200_set_profiling(self._ctx, int(on)))

    def timings(self, slot: int = 0) -> dict:
        t = Timings()
        self._check(self._L.sara_b200_last_timings(self._ctx, slot, C.byref(t)))
        return t.asdict()

    def enqueue(self, slot, image, pyramid_params=None, gauss_truncate=4.0, extremum_thres=0.01,
                edge_ratio_thres=10.0, extremum_refinement_iter=5, stream=None):
        ptr, w, h, on_dev, keep = _as_image(image)
        args = self._sift_args(pyramid_params, gauss_truncate, extremum_thres, edge_ratio_thres,
                               extremum_refinement_iter)
        self._check(self._L.sara_b200_sift_enqueue(self._ctx, slot, ptr, w, h, int(on_dev), C.byref(args),
                                                   C.c_void_p(stream) if stream else None))
        self._keep[slot] = keep  # only once the enqueue succeeded: a BUSY slot keeps its own frame alive

    def enqueue_u8(self, slot, image, pyramid_params=None, gauss_truncate=4.0, extremum_thres=0.01,
                   edge_ratio_thres=10.0, extremum_refinement_iter=5, stream=None):
        """8-bit frame in: HxW (gray8) or HxWx3 (interleaved RGB8); converted on the device as
        from_rgb8_to_gray32f / ImageView<uint8_t>::convert<float>() would (FastColorConversion.cpp:42-67)."""
        ptr, w, h, ch, on_dev, keep = _as_u8_image(image)
        args = self._sift_args(pyramid_params, gauss_truncate, extremum_thres, edge_ratio_thres,
                               extremum_refinement_iter)
        self._check(self._L.sara_b200_sift_enqueue_u8(self._ctx, slot, ptr, w, h, ch, int(on_dev), C.byref(args),
                                                      C.c_void_p(stream) if stream else None))
        self._keep[slot] = keep

    def enqueue_raw_u8(self, slot, ptr, w, h, channels, on_device, args, stream=None):
        self._check(self._L.sara_b200_sift_enqueue_u8(self._ctx, slot, ptr, w, h, channels, int(on_device),
                                                      C.byref(args), C.c_void_p(stream) if stream else None))

    def compute_sift_keypoints_u8(self, image, pyramid_params=None, gauss_truncate=4.0, extremum_thres=0.01,
                                  edge_ratio_thres=10.0, extremum_refinement_iter=5) -> KeypointList:
        self.enqueue_u8(0, image, pyramid_params, gauss_truncate, extremum_thres, edge_ratio_thres,
                        extremum_refinement_iter)
        return self.collect(0)

    def to_gray32f(self, image) -> np.ndarray:
        a = np.ascontiguousarray(image, dtype=np.uint8)
        if a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] != 3):
            raise ValueError("image must be HxW or HxWx3 uint8")
        out = np.empty(a.shape[:2], np.float32)
        self._check(self._L.sara_b200_to_gray32f(self._ctx, a.ctypes.data, a.shape[1], a.shape[0],
                                                 3 if a.ndim == 3 else 1, out.ctypes.data))
        return out

    def enqueue_raw(self, slot, ptr, w, h, on_device, args, stream=None):
        self._check(self._L.sara_b200_sift_enqueue(self._ctx, slot, ptr, w, h, int(on_device), C.byref(args),
                                                   C.c_void_p(stream) if stream else None))

    def wait(self, slot=0) -> int:
        n = C.c_int()
        self._check(self._L.sara_b200_wait(self._ctx, slot, C.byref(n)))
        return n.value

    def collect(self, slot=0, out_keypoints=None, out_descriptors=None) -> KeypointList:
        n = self.wait(slot)
        self._keep.pop(slot, None)
        kps = out_keypoints if out_keypoints is not None else np.empty(max(n, 1), KEYPOINT_DTYPE)
        desc = out_descriptors if out_descriptors is not None else np.empty((max(n, 1), 128), np.float32)
        m = C.c_int()
        self._check(self._L.sara_b200_collect(self._ctx, slot, kps.ctypes.data, desc.ctypes.data,
                                              len(kps), C.byref(m)))
        return KeypointList(kps[: m.value], desc[: m.value])

    def collect_device(self, slot=0):
        """Results of the slot as CUDA tensors (keypoints: (n, 52) uint8 records of KEYPOINT_DTYPE,
        descriptors: (n, 128) float32), copied device to device; nothing crosses PCIe."""
        import torch

        n = self.wait(slot)
        self._keep.pop(slot, None)
        dev = torch.device("cuda", torch.cuda.current_device())
        kps = torch.empty((max(n, 1), 52), dtype=torch.uint8, device=dev)
        desc = torch.empty((max(n, 1), 128), dtype=torch.float32, device=dev)
        m = C.c_int()
        self._check(self._L.sara_b200_collect_device(self._ctx, slot, kps.data_ptr(), desc.data_ptr(), max(n, 1),
                                                     C.byref(m)))
        return kps[: m.value], desc[: m.value]

    def collect_into(self, slot, kps_ptr, desc_ptr, capacity) -> int:
        m = C.c_int()
        self._check(self._L.sara_b200_collect(self._ctx, slot, kps_ptr, desc_ptr, capacity, C.byref(m)))
        return m.value

    def compute_sift_keypoints(self, image, pyramid_params=None, gauss_truncate=4.0, extremum_thres=0.01,
                               edge_ratio_thres=10.0, extremum_refinement_iter=5, parallel=True) -> KeypointList:
        self.enqueue(0, image, pyramid_params, gauss_truncate, extremum_thres, edge_ratio_thres,
                     extremum_refinement_iter)
        return self.collect(0)

    # ---- pyramid only ---------------------------------------------------------------
    def pyramid_enqueue(self, slot, image, pyramid_params=None, gauss_truncate=4.0, stream=None):
        ptr, w, h, on_dev, keep = _as_image(image)
        pp = (pyramid_params or ImagePyramidParams())._c()
        self._check(self._L.sara_b200_pyramid_enqueue(self._ctx, slot, ptr, w, h, int(on_dev), C.byref(pp),
                                                      gauss_truncate, C.c_void_p(stream) if stream else None))
        self._keep[slot] = keep

    # ---- ComputeDoGExtrema ------------------------------------------------------------
    def dog_extrema(self, image, pyramid_params=None, gauss_truncate=4.0, extremum_thres=0.01,
                    edge_ratio_thres=10.0, img_padding_sz=1, extremum_refinement_iter=5, slot=0) -> np.ndarray:
        ptr, w, h, on_dev, keep = _as_image(image)
        args = _DogArgs((pyramid_params or ImagePyramidParams())._c(), gauss_truncate, extremum_thres,
                        edge_ratio_thres, int(img_padding_sz), int(extremum_refinement_iter))
        self._check(self._L.sara_b200_dog_extrema(self._ctx, slot, ptr, w, h, int(on_dev), C.byref(args)))
        del keep
        return self.extrema(slot)

    def function_extrema(self, which: str, image, pyramid_params=None, extremum_thres=0.01, edge_ratio_thres=10.0,
                         img_padding_sz=1, extremum_refinement_iter=5, slot=0) -> np.ndarray:
        """ComputeLoGExtrema (which = "log") / ComputeDoHExtrema ("doh"): the function pyramid is then read with
        dog_layer(s, o) for s < num_scales()."""
        ptr, w, h, on_dev, keep = _as_image(image)
        args = _DogArgs((pyramid_params or ImagePyramidParams(scale_count_per_octave=5))._c(), 4.0, extremum_thres,
                        edge_ratio_thres, int(img_padding_sz), int(extremum_refinement_iter))
        fn = {"log": self._L.sara_b200_log_extrema, "doh": self._L.sara_b200_doh_extrema}[which]
        self._check(fn(self._ctx, slot, ptr, w, h, int(on_dev), C.byref(args)))
        del keep
        return self.extrema(slot)

    def hessian_laplace(self, image, pyramid_params=None, extremum_thres=1e-5, img_padding_sz=1, num_scales=10,
                        extremum_refinement_iter=5, slot=0) -> np.ndarray:
        """ComputeHessianLaplaceMaxima (Hessian.hpp:84-94): det-of-Hessian maxima with Laplace scale selection."""
        ptr, w, h, on_dev, keep = _as_image(image)
        args = _DogArgs((pyramid_params or ImagePyramidParams(scale_count_per_octave=4))._c(), 4.0, extremum_thres, 10.0,
                        int(img_padding_sz), int(extremum_refinement_iter))
        self._check(self._L.sara_b200_hessian_laplace(self._ctx, slot, ptr, w, h, int(on_dev), C.byref(args),
                                                      int(num_scales)))
        del keep
        return self.extrema(slot)

    def harris_laplace(self, image, pyramid_params=None, kappa=0.04, extremum_thres=1e-6, img_padding_sz=1,
                       num_scales=10, extremum_refinement_iter=5, slot=0) -> np.ndarray:
        """ComputeHarrisLaplaceCorners (Harris.hpp:125-138)."""
        ptr, w, h, on_dev, keep = _as_image(image)
        pp = pyramid_params or ImagePyramidParams(-1, 3, float(np.sqrt(np.float32(2.0))), 1)
        args = _DogArgs(pp._c(), 4.0, extremum_thres, 10.0, int(img_padding_sz), int(extremum_refinement_iter))
        self._check(self._L.sara_b200_harris_laplace(self._ctx, slot, ptr, w, h, int(on_dev), C.byref(args), kappa,
                                                     int(num_scales)))
        del keep
        return self.extrema(slot)

    def describe_extrema(self, extrema, slot=0):
        """ComputeDominantOrientations + ComputeSIFTDescriptor<4, 8> (+ rescale) on the given extrema against the
        pyramid the slot holds.  Returns (oriented keypoints in octave coordinates, KeypointList in image
        coordinates)."""
        e = np.ascontiguousarray(extrema, dtype=KEYPOINT_DTYPE)
        cap = max(4 * len(e), 16)
        while True:
            ori = np.empty(cap, KEYPOINT_DTYPE)
            kps = np.empty(cap, KEYPOINT_DTYPE)
            desc = np.empty((cap, 128), np.float32)
            n = C.c_int(0)
            rc = self._L.sara_b200_describe_extrema(self._ctx, slot, e.ctypes.data, len(e), ori.ctypes.data,
                                                    kps.ctypes.data, desc.ctypes.data, cap, C.byref(n))
            if rc == -5 and n.value > cap:
                cap = n.value
                continue
            self._check(rc)
            return ori[: n.value].copy(), KeypointList(kps[: n.value].copy(), desc[: n.value].copy())

    # ---- stage accessors -----------------------------------------------------------------
    def num_octaves(self, slot=0) -> int:
        return self._L.sara_b200_num_octaves(self._ctx, slot)

    def num_scales(self, slot=0) -> int:
        return self._L.sara_b200_num_scales(self._ctx, slot)

    def layer_size(self, o, slot=0):
        w, h = C.c_int(), C.c_int()
        self._check(self._L.sara_b200_layer_size(self._ctx, slot, o, C.byref(w), C.byref(h)))
        return w.value, h.value

    def octave_scaling_factor(self, o, slot=0) -> float:
        return float(self._L.sara_b200_octave_scaling_factor(self._ctx, slot, o))

    def _layer(self, which, s, o, slot):
        w, h = self.layer_size(o, slot)
        a = np.empty((h, w), np.float32)
        self._check(self._L.sara_b200_copy_layer(self._ctx, slot, which, s, o, a.ctypes.data))
        return a

    def gaussian_layer(self, s, o, slot=0) -> np.ndarray:
        return self._layer(0, s, o, slot)

    def dog_layer(self, s, o, slot=0) -> np.ndarray:
        return self._layer(1, s, o, slot)

    def _kps(self, fn, slot):
        n = C.c_int()
        rc = fn(self._ctx, slot, None, 0, C.byref(n))
        if n.value == 0:
            if rc not in (0, -5):
                self._check(rc)
            return np.empty(0, KEYPOINT_DTYPE)
        a = np.empty(n.value, KEYPOINT_DTYPE)
        self._check(fn(self._ctx, slot, a.ctypes.data, len(a), C.byref(n)))
        return a

    def extrema(self, slot=0) -> np.ndarray:
        return self._kps(self._L.sara_b200_copy_extrema, slot)

    def oriented(self, slot=0) -> np.ndarray:
        return self._kps(self._L.sara_b200_copy_oriented, slot)

    def device_results(self, slot=0):
        kp, ds, n = C.c_void_p(), C.c_void_p(), C.c_int()
        self._check(self._L.sara_b200_device_results(self._ctx, slot, C.byref(kp), C.byref(ds), C.byref(n)))
        return kp.value, ds.value, n.value

    # ---- building blocks -------------------------------------------------------------------
    def gaussian(self, image, sigma: float, gauss_truncate: float = 4.0) -> np.ndarray:
        a = np.ascontiguousarray(image, dtype=np.float32)
        out = np.empty_like(a)
        self._check(self._L.sara_b200_gaussian(self._ctx, a.ctypes.data, a.shape[1], a.shape[0], sigma,
                                               gauss_truncate, out.ctypes.data))
        return out


def make_gaussian_kernel(sigma: float, gauss_truncate: float = 4.0) -> np.ndarray:
    L = load_library()
    buf = np.zeros(256, np.float32)
    n = L.sara_b200_make_gaussian_kernel(sigma, gauss_truncate, buf.ctypes.data_as(C.POINTER(C.c_float)), 256)
    if n < 0:
        raise ValueError(f"kernel needs {-n} taps")
    return buf[:n].copy()


class ComputeDoGExtrema:
    """DO::Sara::ComputeDoGExtrema (FeatureDetectors/DoG.hpp:72-165): same constructor
    arguments, call operator and accessors."""

    def __init__(self, pyramid_params: ImagePyramidParams | None = None, gauss_truncate=4.0,
                 extremum_thres=0.01, edge_ratio_thres=10.0, img_padding_sz=1, extremum_refinement_iter=5,
                 device: int = 0):
        self.params = pyramid_params or ImagePyramidParams()
        if self.params.scale_count_per_octave < 4:
            raise RuntimeError("Error: The extraction of DoG extrema needs (1 + 3) = 4 scales per octave "
                               "at the very minimum!")
        self.args = (gauss_truncate, extremum_thres, edge_ratio_thres, img_padding_sz, extremum_refinement_iter)
        self.device = device
        self._ctx = None

    def __call__(self, image):
        """Returns (extrema, scale_octave_pairs) as DoG.cpp:23-87."""
        _, w, h, _, _ = _as_image(image)
        if self._ctx is None:
            self._ctx = SiftContext(w, h, self.device, min_first_octave_index=min(self.params.first_octave_index, 0))
        e = self._ctx.dog_extrema(image, self.params, *self.args)
        return e, np.stack([e["s"], e["o"]], axis=1)

    def gaussians(self, s, o):
        return self._ctx.gaussian_layer(s, o)

    def diff_of_gaussians(self, s, o):
        return self._ctx.dog_layer(s, o)

    def extrema(self, s, o):
        e = self._ctx.extrema()
        return e[(e["s"] == s) & (e["o"] == o)]


class _ComputeFunctionExtrema:
    _which = ""
    _defaults = (0.01, 10.0, 1, 5)
    _params = dict(first_octave_index=-1, scale_count_per_octave=5)

    def __init__(self, pyramid_params: ImagePyramidParams | None = None, extremum_thres=None, edge_ratio_thres=None,
                 img_padding_sz=None, extremum_refinement_iter=None, device: int = 0):
        self.params = pyramid_params or ImagePyramidParams(**self._params)
        given = (extremum_thres, edge_ratio_thres, img_padding_sz, extremum_refinement_iter)
        self.args = tuple(d if g is None else g for g, d in zip(given, self._defaults))
        self.device = device
        self._ctx = None

    def __call__(self, image):
        """Returns (extrema, scale_octave_pairs)."""
        _, w, h, _, _ = _as_image(image)
        if self._ctx is None:
            self._ctx = SiftContext(w, h, self.device, min_first_octave_index=min(self.params.first_octave_index, 0))
        e = self._ctx.function_extrema(self._which, image, self.params, *self.args)
        return e, np.stack([e["s"], e["o"]], axis=1)

    def gaussians(self, s, o):
        return self._ctx.gaussian_layer(s, o)

    def _function(self, s, o):
        return self._ctx.dog_layer(s, o)


class ComputeLoGExtrema(_ComputeFunctionExtrema):
    """DO::Sara::ComputeLoGExtrema (FeatureDetectors/LoG.hpp:71-117): constructor defaults
    ImagePyramidParams(-1, 3 + 2), 0.01, 10, 1, 5; call operator; laplacians_of_gaussians(s, o)."""
    _which = "log"

    def laplacians_of_gaussians(self, s, o):
        return self._function(s, o)


class ComputeDoHExtrema(_ComputeFunctionExtrema):
    """DO::Sara::ComputeDoHExtrema (FeatureDetectors/Hessian.hpp:195-240): constructor defaults
    ImagePyramidParams(-1, 3 + 2, 2^(1/3), 2), 1e-6, 10, 1, 2; call operator; det_of_hessians(s, o)."""
    _which = "doh"
    _defaults = (1e-6, 10.0, 1, 2)
    _params = dict(first_octave_index=-1, scale_count_per_octave=5, image_padding_size=2)

    def det_of_hessians(self, s, o):
        return self._function(s, o)


class ComputeHessianLaplaceMaxima:
    """DO::Sara::ComputeHessianLaplaceMaxima (FeatureDetectors/Hessian.hpp:60-127): constructor defaults
    ImagePyramidParams(-1, 3 + 1), 1e-5, 1, 10 scales, 5 iterations; call operator; gaussians / det_of_hessians."""

    def __init__(self, pyramid_params: ImagePyramidParams | None = None, extremum_thres=1e-5, img_padding_sz=1,
                 num_scales=10, extremum_refinement_iter=5, device: int = 0):
        self.params = pyramid_params or ImagePyramidParams(first_octave_index=-1, scale_count_per_octave=4)
        self.args = (extremum_thres, img_padding_sz, num_scales, extremum_refinement_iter)
        self.device = device
        self._ctx = None

    def __call__(self, image):
        _, w, h, _, _ = _as_image(image)
        if self._ctx is None:
            self._ctx = SiftContext(w, h, self.device, min_first_octave_index=min(self.params.first_octave_index, 0))
        e = self._ctx.hessian_laplace(image, self.params, *self.args)
        return e, np.stack([e["s"], e["o"]], axis=1)

    def gaussians(self, s, o):
        return self._ctx.gaussian_layer(s, o)

    def det_of_hessians(self, s, o):
        return self._ctx.dog_layer(s, o)


class ComputeHarrisLaplaceCorners:
    """DO::Sara::ComputeHarrisLaplaceCorners (FeatureDetectors/Harris.hpp:95-175): constructor defaults
    ImagePyramidParams(-1, 2 + 1, sqrt(2), 1), kappa 0.04, 1e-6, 1, 10 scales, 5 iterations; call operator;
    gaussians / harris."""

    def __init__(self, pyramid_params: ImagePyramidParams | None = None, kappa=0.04, extremum_thres=1e-6,
                 img_padding_sz=1, scale_count=10, extremum_refinement_iter=5, device: int = 0):
        self.params = pyramid_params or ImagePyramidParams(-1, 3, float(np.sqrt(np.float32(2.0))), 1)
        self.args = (kappa, extremum_thres, img_padding_sz, scale_count, extremum_refinement_iter)
        self.device = device
        self._ctx = None

    def __call__(self, image):
        _, w, h, _, _ = _as_image(image)
        if self._ctx is None:
            self._ctx = SiftContext(w, h, self.device, min_first_octave_index=min(self.params.first_octave_index, 0))
        e = self._ctx.harris_laplace(image, self.params, *self.args)
        return e, np.stack([e["s"], e["o"]], axis=1)

    def gaussians(self, s, o):
        return self._ctx.gaussian_layer(s, o)

    def harris(self, s, o):
        return self._ctx.dog_layer(s, o)


_DEFAULT_CTX: dict = {}


def compute_sift_keypoints(image, pyramid_params: ImagePyramidParams | None = None, gauss_truncate: float = 4.0,
                           extremum_thres: float = 0.01, edge_ratio_thres: float = 10.0,
                           extremum_refinement_iter: int = 5, parallel: bool = True, device: int = 0) -> KeypointList:
    """pysara.compute_sift_keypoints (pybind11/FeatureDetectors.cpp:116-124), same
    argument order and defaults; `parallel` is accepted and ignored (the GPU path is
    always parallel).  One context per device is kept, sized for the largest width and the
    largest height seen so far (frames of alternating orientation do not recreate it)."""
    ptr, w, h, on_dev, keep = _as_image(image)
    ctx = _DEFAULT_CTX.get(device)
    if ctx is None or ctx._w < w or ctx._h < h:
        mw, mh = (max(w, ctx._w), max(h, ctx._h)) if ctx is not None else (w, h)
        if ctx is not None:
            ctx.close()
        ctx = SiftContext(mw, mh, device)
        ctx._w, ctx._h = mw, mh
        _DEFAULT_CTX[device] = ctx
    # `keep` is the converted array / tensor: handed on as is, no second conversion
    return ctx.compute_sift_keypoints(keep, pyramid_params, gauss_truncate, extremum_thres, edge_ratio_thres,
                                      extremum_refinement_iter)


class AnnMatcher:
    """DO::Sara::AnnMatcher (FeatureMatching/AnnMatcher.hpp:33-84): same constructors, `compute_matches()`
    returns the matches ordered by score as a MATCH_DTYPE array (Match::x_index, y_index, rank, score,
    matching_direction).  The search is exact (FLANN's LinearIndex answer), not the reference's KD-tree forest."""

    def __init__(self, keys1: KeypointList, keys2: KeypointList | float | None = None, sift_ratio_thres: float = 1.2,
                 min_max_metric_dist_thres: float = 0.5, pixel_dist_thres: float = 10.0, ctx: SiftContext | None = None,
                 device: int = 0):
        if keys2 is None or isinstance(keys2, (int, float)):  # AnnMatcher(keys, ratio, metric, pixel): self matching
            if isinstance(keys2, (int, float)):
                sift_ratio_thres = float(keys2)
            keys2, self._self = keys1, True
        else:
            self._self = False
        for k in (keys1, keys2):  # size_consistency_predicate (Features/KeypointList.hpp)
            if len(k.features) != len(k.descriptors):
                raise RuntimeError("The list of keypoints are inconsistent in size!")
        self.keys1, self.keys2 = keys1, keys2
        self.sift_ratio_thres = sift_ratio_thres
        self.metric, self.pixel = min_max_metric_dist_thres, pixel_dist_thres
        self._ctx, self._device = ctx, device

    def compute_matches(self, mode="auto") -> np.ndarray:
        ctx = self._ctx
        if ctx is None:
            ctx = _DEFAULT_CTX.get(self._device)
            if ctx is None:
                ctx = SiftContext(64, 64, self._device)
                ctx._w, ctx._h = 64, 64
                _DEFAULT_CTX[self._device] = ctx
        return ctx.compute_matches(self.keys1.descriptors, self.keys2.descriptors, self.sift_ratio_thres,
                                   self.keys1.features, self.keys2.features, self._self, self.metric, self.pixel, mode)

    compute_self_matches = compute_matches


def match(keys1: KeypointList, keys2: KeypointList, lowe_ratio: float = 0.6) -> np.ndarray:
    """DO::Sara::match (SfM/Helpers/KeypointMatching.cpp:19-25)."""
    return AnnMatcher(keys1, keys2, lowe_ratio).compute_matches()
