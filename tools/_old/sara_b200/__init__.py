"""B200-native SIFT path behind the API surface of oddkiva/sara.

Host-side mirror of the reference's Python binding
(python/oddkiva/sara/pybind11/FeatureDetectors.cpp:57-124) on top of the C ABI
in include/sara_b200.h.  Every call runs hand-written sm_100a CUDA kernels;
there is no CPU fallback (creation raises when the library or a device is
missing).
"""
from .api import (  # noqa: F401
    AnnMatcher,
    MATCH_DTYPE,
    match,
    ComputeDoGExtrema,
    ComputeDoHExtrema,
    ComputeHarrisLaplaceCorners,
    ComputeHessianLaplaceMaxima,
    ComputeLoGExtrema,
    ImagePyramidParams,
    KEYPOINT_DTYPE,
    KeypointList,
    SaraB200Error,
    SiftContext,
    compute_sift_keypoints,
    descriptors,
    features,
    library_path,
    load_library,
)
from .features_io import read_keypoints, remove_redundant_features, write_keypoints  # noqa: F401,E402
