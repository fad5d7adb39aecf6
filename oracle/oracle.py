"""ctypes loader for the CPU oracle (oracle/sift_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under sara_b200/
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsara_oracle.so")

KEYPOINT_DTYPE = np.dtype(
    [
        ("x", "<f4"),
        ("y", "<f4"),
        ("shape", "<f4", (4,)),
        ("orientation", "<f4"),
        ("extremum_value", "<f4"),
        ("type", "u1"),
        ("extremum_type", "i1"),
        ("pad_", "<i2"),
        ("s", "<i4"),
        ("o", "<i4"),
        ("xi", "<i4"),
        ("yi", "<i4"),
    ]
)


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc only, no GPU)."""
    src = os.path.join(_HERE, "sift_oracle.cpp")
    srcs = [src, os.path.join(_HERE, "match_oracle.cpp")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in srcs
    )
    if force or stale:
        if not os.path.exists(src):
            raise RuntimeError("oracle source missing and no prebuilt library")
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_refine_peak.restype = C.c_float
        L.oracle_octave_scaling.restype = C.c_float
        L.oracle_octave_scaling.argtypes = [C.c_void_p, C.c_int]
        L.oracle_sift.argtypes = [
            fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
            C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
            C.POINTER(C.c_void_p),
        ]
        L.oracle_dog_extrema.argtypes = [
            fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
            C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
            C.POINTER(C.c_void_p),
        ]
        L.oracle_function_extrema.argtypes = [
            fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
            C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
            C.POINTER(C.c_void_p),
        ]
        L.oracle_hessian_laplace.argtypes = [
            fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
            C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
        ]
        L.oracle_harris_laplace.argtypes = [
            fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
            C.c_float, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
        ]
        for name in (
            "oracle_free", "oracle_num_octaves", "oracle_num_scales",
            "oracle_num_extrema", "oracle_num_keypoints",
        ):
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_layer_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_copy_layer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, fp]
        L.oracle_copy_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_copy_descriptors.argtypes = [C.c_void_p, fp]
        L.oracle_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        assert L.oracle_keypoint_stride() == KEYPOINT_DTYPE.itemsize
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _check(rc: int):
    if rc != 0:
        raise RuntimeError(lib().oracle_last_error().decode())


@dataclass
class PyramidParams:
    """ImagePyramidParams (ImageProcessing/ImagePyramid.hpp:33-49) defaults."""

    first_octave_index: int = -1
    scale_count_per_octave: int = 6
    scale_geometric_factor: float = float(np.float32(2.0) ** np.float32(1.0 / 3.0))
    image_padding_size: int = 1
    scale_camera: float = 0.5
    scale_initial: float = 1.6
    num_octaves_max: int = 2**31 - 1

    def astuple(self):
        return (
            int(self.first_octave_index), int(self.scale_count_per_octave),
            float(self.scale_geometric_factor), int(self.image_padding_size),
            float(self.scale_camera), float(self.scale_initial),
            int(self.num_octaves_max),
        )


def rgb8_to_gray32f(rgb) -> np.ndarray:
    """from_rgb8_to_gray32f (FastColorConversion.cpp:42-67, non-Halide branch): HxWx3 uint8 -> HxW float32."""
    a = np.ascontiguousarray(rgb, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3
    out = np.empty(a.shape[:2], np.float32)
    lib().oracle_rgb8_to_gray32f(a.ctypes.data_as(C.c_void_p), a.shape[0] * a.shape[1], _p(out))
    return out


def gray8_to_gray32f(gray) -> np.ndarray:
    a = np.ascontiguousarray(gray, dtype=np.uint8)
    out = np.empty(a.shape, np.float32)
    lib().oracle_gray8_to_gray32f(a.ctypes.data_as(C.c_void_p), a.size, _p(out))
    return out


TRACE_FIELDS = 25  # x, y, s, o, type, iteration, H[9], g[3], lambda[3], h[3], decision


def trace_refinement(on: bool) -> None:
    """Starts (and clears) / stops the per-iteration trace of refine_extremum."""
    lib().oracle_trace_refinement(int(on))


def refinement_trace() -> np.ndarray:
    """(n, 25) float32: one record per Newton iteration that reached the eigenvalue test."""
    n = lib().oracle_trace_size()
    out = np.zeros((n, TRACE_FIELDS), np.float32)
    if n:
        lib().oracle_trace_copy(_p(out))
    return out


def set_threading(mode: int, threads: int = 0) -> None:
    lib().oracle_set_threading(int(mode), int(threads))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ---- unit-level entry points (images are HxW float32 arrays) ---------------
def make_gaussian_kernel(sigma: float, truncate: float = 4.0) -> np.ndarray:
    out = np.zeros(4096, np.float32)
    n = lib().oracle_make_gaussian_kernel(C.c_float(sigma), C.c_float(truncate), _p(out), out.size)
    assert n > 0
    return out[:n].copy()


def convolve_array(signal, kernel) -> np.ndarray:
    sig = _f32(signal).copy()
    ker = _f32(kernel)
    n = sig.size - ker.size + 1
    lib().oracle_convolve_array(_p(sig), _p(ker), n, ker.size)
    return sig


def row_filter(img, kernel) -> np.ndarray:
    img = _f32(img); ker = _f32(kernel)
    h, w = img.shape
    dst = np.empty_like(img)
    _check(lib().oracle_row_filter(_p(img), w, h, _p(ker), ker.size, _p(dst)))
    return dst


def column_filter(img, kernel) -> np.ndarray:
    img = _f32(img); ker = _f32(kernel)
    h, w = img.shape
    dst = np.empty_like(img)
    _check(lib().oracle_column_filter(_p(img), w, h, _p(ker), ker.size, _p(dst)))
    return dst


def gaussian(img, sigma: float, truncate: float = 4.0) -> np.ndarray:
    img = _f32(img)
    h, w = img.shape
    dst = np.empty_like(img)
    _check(lib().oracle_gaussian(_p(img), w, h, C.c_float(sigma), C.c_float(truncate), _p(dst)))
    return dst


def interpolate(img, x: float, y: float) -> float:
    img = _f32(img)
    h, w = img.shape
    out = C.c_double()
    _check(lib().oracle_interpolate(_p(img), w, h, C.c_double(x), C.c_double(y), C.byref(out)))
    return out.value


def enlarge(img, new_w: int, new_h: int) -> np.ndarray:
    img = _f32(img)
    h, w = img.shape
    dst = np.empty((new_h, new_w), np.float32)
    _check(lib().oracle_enlarge(_p(img), w, h, _p(dst), new_w, new_h))
    return dst


def downscale(img, fact: int) -> np.ndarray:
    img = _f32(img)
    h, w = img.shape
    dst = np.empty((h // fact, w // fact), np.float32)
    _check(lib().oracle_downscale(_p(img), w, h, fact, _p(dst)))
    return dst


def local_scale_space_extremum(stack3, x: int, y: int, kind: int) -> bool:
    st = _f32(stack3)
    _, h, w = st.shape
    return bool(lib().oracle_local_scale_space_extremum(_p(st), w, h, x, y, kind))


def local_extremum(img, x: int, y: int, kind: int) -> bool:
    img = _f32(img)
    h, w = img.shape
    return bool(lib().oracle_local_extremum(_p(img), w, h, x, y, kind))


def on_edge(img, x: int, y: int, ratio: float) -> bool:
    img = _f32(img)
    h, w = img.shape
    return bool(lib().oracle_on_edge(_p(img), w, h, x, y, C.c_float(ratio)))


def hessian2(img, x: int, y: int) -> np.ndarray:
    img = _f32(img)
    h, w = img.shape
    H = np.zeros(4, np.float32)
    lib().oracle_hessian2(_p(img), w, h, x, y, _p(H))
    return H.reshape(2, 2)


def sym3_eigenvalues(H) -> np.ndarray:
    H = _f32(H).reshape(9)
    lam = np.zeros(3, np.float32)
    lib().oracle_sym3_eigenvalues(_p(H), _p(lam))
    return lam


def inverse3(M) -> np.ndarray:
    M = _f32(M).reshape(9)
    R = np.zeros(9, np.float32)
    lib().oracle_inverse3(_p(M), _p(R))
    return R.reshape(3, 3)


def gradient_polar(img) -> np.ndarray:
    """HxWx2 (mag = 2*|grad|, ori = atan2) as Orientation.cpp:24-56."""
    img = _f32(img)
    h, w = img.shape
    out = np.empty((h, w, 2), np.float32)
    lib().oracle_gradient_polar(_p(img), w, h, _p(out))
    return out


def orientation_histogram(grad, x, y, s, trunc=3.0, blur=1.5) -> np.ndarray:
    g = _f32(grad)
    h, w, _ = g.shape
    hist = np.zeros(36, np.float32)
    lib().oracle_orientation_histogram(
        _p(g), w, h, C.c_float(x), C.c_float(y), C.c_float(s), C.c_float(trunc),
        C.c_float(blur), _p(hist))
    return hist


def lowe_smooth_histogram(hist, iters: int = 6) -> np.ndarray:
    hh = _f32(hist).copy()
    lib().oracle_lowe_smooth_histogram(_p(hh), iters)
    return hh


def find_peaks(hist, ratio: float = 0.8) -> list[int]:
    hh = _f32(hist)
    peaks = (C.c_int * 36)()
    n = lib().oracle_find_peaks(_p(hh), C.c_float(ratio), peaks)
    return [peaks[i] for i in range(n)]


def refine_peak(hist, i: int) -> float:
    return float(lib().oracle_refine_peak(_p(_f32(hist)), i))


def dominant_orientations(grad, x, y, sigma) -> np.ndarray:
    g = _f32(grad)
    h, w, _ = g.shape
    out = np.zeros(36, np.float32)
    n = lib().oracle_dominant_orientations(
        _p(g), w, h, C.c_float(x), C.c_float(y), C.c_float(sigma), _p(out))
    return out[:n].copy()


def sift_descriptor(grad, x, y, s, theta, normalize: bool = True) -> np.ndarray:
    g = _f32(grad)
    h, w, _ = g.shape
    d = np.zeros(128, np.float32)
    lib().oracle_sift_descriptor(
        _p(g), w, h, C.c_float(x), C.c_float(y), C.c_float(s), C.c_float(theta),
        int(normalize), _p(d))
    return d


# ---- whole pipeline ---------------------------------------------------------
class SiftResult:
    """Owns an oracle Result handle; exposes pyramids, extrema, keypoints."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_free(self._h)
            self._h = None

    @property
    def num_octaves(self) -> int:
        return lib().oracle_num_octaves(self._h)

    @property
    def num_scales(self) -> int:
        return lib().oracle_num_scales(self._h)

    def octave_scaling(self, o: int) -> float:
        return float(lib().oracle_octave_scaling(self._h, o))

    def layer_size(self, o: int):
        w, h = C.c_int(), C.c_int()
        lib().oracle_layer_size(self._h, o, C.byref(w), C.byref(h))
        return w.value, h.value

    def gaussian(self, s: int, o: int) -> np.ndarray:
        w, h = self.layer_size(o)
        a = np.empty((h, w), np.float32)
        lib().oracle_copy_layer(self._h, 0, s, o, _p(a))
        return a

    def dog(self, s: int, o: int) -> np.ndarray:
        w, h = self.layer_size(o)
        a = np.empty((h, w), np.float32)
        lib().oracle_copy_layer(self._h, 1, s, o, _p(a))
        return a

    def _kps(self, which: int, n: int) -> np.ndarray:
        a = np.zeros(n, KEYPOINT_DTYPE)
        if n:
            lib().oracle_copy_keypoints(self._h, which, a.ctypes.data_as(C.c_void_p))
        return a

    @property
    def extrema(self) -> np.ndarray:
        return self._kps(0, lib().oracle_num_extrema(self._h))

    @property
    def oriented(self) -> np.ndarray:
        return self._kps(1, lib().oracle_num_keypoints(self._h))

    @property
    def keypoints(self) -> np.ndarray:
        return self._kps(2, lib().oracle_num_keypoints(self._h))

    @property
    def descriptors(self) -> np.ndarray:
        n = lib().oracle_num_keypoints(self._h)
        d = np.zeros((n, 128), np.float32)
        if n:
            lib().oracle_copy_descriptors(self._h, _p(d))
        return d

    @property
    def stage_ms(self) -> dict:
        ms = (C.c_double * 4)()
        lib().oracle_stage_ms(self._h, ms)
        return {"dog": ms[0], "gradient": ms[1], "orientation": ms[2], "descriptors": ms[3]}


def compute_sift_keypoints(
    image,
    pyramid_params: PyramidParams | None = None,
    gauss_truncate: float = 4.0,
    extremum_thres: float = 0.01,
    edge_ratio_thres: float = 10.0,
    extremum_refinement_iter: int = 5,
    parallel: bool = False,
) -> SiftResult:
    """FeatureDetectors/SIFT.hpp:24-33 (same argument order and defaults)."""
    img = _f32(image)
    h, w = img.shape
    pp = (pyramid_params or PyramidParams()).astuple()
    out = C.c_void_p()
    _check(lib().oracle_sift(
        _p(img), w, h, pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6],
        gauss_truncate, extremum_thres, edge_ratio_thres,
        int(extremum_refinement_iter), int(parallel), C.byref(out)))
    return SiftResult(out)


def compute_dog_extrema(
    image,
    pyramid_params: PyramidParams | None = None,
    gauss_truncate: float = 4.0,
    extremum_thres: float = 0.01,
    edge_ratio_thres: float = 10.0,
    img_padding_sz: int = 1,
    extremum_refinement_iter: int = 5,
) -> SiftResult:
    """ComputeDoGExtrema (FeatureDetectors/DoG.hpp:72-78) ctor + operator()."""
    img = _f32(image)
    h, w = img.shape
    pp = (pyramid_params or PyramidParams()).astuple()
    out = C.c_void_p()
    _check(lib().oracle_dog_extrema(
        _p(img), w, h, pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6],
        gauss_truncate, extremum_thres, edge_ratio_thres, int(img_padding_sz),
        int(extremum_refinement_iter), C.byref(out)))
    return SiftResult(out)


def compute_function_extrema(
    image,
    which: str,
    pyramid_params: PyramidParams | None = None,
    extremum_thres: float = 0.01,
    edge_ratio_thres: float = 10.0,
    img_padding_sz: int = 1,
    extremum_refinement_iter: int = 5,
) -> SiftResult:
    """ComputeLoGExtrema (FeatureDetectors/LoG.hpp:71-83, LoG.cpp:20-58; which = "log") or
    ComputeDoHExtrema (FeatureDetectors/Hessian.cpp:59-98; which = "doh").  Defaults of both
    constructors: ImagePyramidParams(-1, 3 + 2).  The function pyramid is layer kind 1."""
    img = _f32(image)
    h, w = img.shape
    pp = (pyramid_params or PyramidParams(scale_count_per_octave=5)).astuple()
    out = C.c_void_p()
    _check(lib().oracle_function_extrema(
        _p(img), w, h, pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6], {"log": 1, "doh": 2}[which],
        extremum_thres, edge_ratio_thres, int(img_padding_sz), int(extremum_refinement_iter), C.byref(out)))
    return SiftResult(out)


def compute_hessian_laplace(
    image,
    pyramid_params: PyramidParams | None = None,
    extremum_thres: float = 1e-5,
    img_padding_sz: int = 1,
    num_scales: int = 10,
    extremum_refinement_iter: int = 5,
) -> SiftResult:
    """ComputeHessianLaplaceMaxima (FeatureDetectors/Hessian.hpp:84-94, Hessian.cpp:19-57): det-of-Hessian
    maxima with Laplace scale selection (laplace_maxima, RefineExtremum.cpp:659-709)."""
    img = _f32(image)
    h, w = img.shape
    pp = (pyramid_params or PyramidParams(scale_count_per_octave=4)).astuple()
    out = C.c_void_p()
    _check(lib().oracle_hessian_laplace(
        _p(img), w, h, pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6], extremum_thres, int(img_padding_sz),
        int(num_scales), int(extremum_refinement_iter), C.byref(out)))
    return SiftResult(out)


def compute_harris_laplace(
    image,
    pyramid_params: PyramidParams | None = None,
    kappa: float = 0.04,
    extremum_thres: float = 1e-6,
    img_padding_sz: int = 1,
    num_scales: int = 10,
    extremum_refinement_iter: int = 5,
) -> SiftResult:
    """ComputeHarrisLaplaceCorners (FeatureDetectors/Harris.hpp:125-138, Harris.cpp:165-230)."""
    img = _f32(image)
    h, w = img.shape
    pp = (pyramid_params or PyramidParams(-1, 3, float(np.sqrt(np.float32(2.0))), 1)).astuple()
    out = C.c_void_p()
    _check(lib().oracle_harris_laplace(
        _p(img), w, h, pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6], kappa, extremum_thres, int(img_padding_sz),
        int(num_scales), int(extremum_refinement_iter), C.byref(out)))
    return SiftResult(out)
