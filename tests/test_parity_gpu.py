"""Parity of the CUDA path (called through the C ABI) against the CPU oracle.

All of these need a B200 (`-m gpu`).  Sizes are chosen so the oracle finishes in
seconds; the 4K / 1080p configurations are covered by size-independent
properties in test_properties_gpu.py.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from sara_b200 import synthetic as S
import sara_b200 as sb
from parity import assert_extrema_identical, assert_pyramids_identical, compare_keypoints

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    c = sb.SiftContext(1024, 768, max_keypoints=65536, num_slots=2, min_first_octave_index=-1)
    yield c
    c.close()


def _pp(fo=0, **kw):
    return sb.ImagePyramidParams(first_octave_index=fo, **kw), O.PyramidParams(first_octave_index=fo, **kw)


# ---- apply_gaussian_filter: LinearFiltering.cpp:30-68 -------------------------
@pytest.mark.parametrize("w,h", [(1, 1), (3, 3), (37, 23), (64, 32), (65, 33), (130, 70), (640, 480)])
@pytest.mark.parametrize("sigma", [0.5, 1.2262735, 1.5198684, 3.0900156, 6.0])
def test_gaussian_bit_exact(ctx, w, h, sigma):
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.random((h, w), dtype=np.float32)
    got, ref = ctx.gaussian(img, sigma), O.gaussian(img, sigma)
    assert got.tobytes() == ref.tobytes(), f"max |d| = {np.abs(got - ref).max()}"


def test_gaussian_of_dirac(ctx):
    # test_imageprocessing_linear_filtering.cpp:136-187, on the GPU
    for n, trunc in [(3, 1.0), (9, 4.0), (65, 4.0)]:
        img = np.zeros((n, n), np.float32)
        img[n // 2, n // 2] = 1
        c = n // 2
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        true = np.exp(-((i - c) ** 2 + (j - c) ** 2) / 2.0)
        true = (true / true.sum()).astype(np.float32)
        assert np.linalg.norm(true - ctx.gaussian(img, 1.0, trunc)) < 1e-5


def test_gaussian_kernel_bits():
    from sara_b200.api import make_gaussian_kernel

    for sigma in (0.3, 1.2262735, 1.5450078, 1.9465878, 2.4525470, 3.0900156, 1.5198684):
        assert make_gaussian_kernel(sigma).tobytes() == O.make_gaussian_kernel(sigma).tobytes()


# ---- gaussian_pyramid + difference_of_gaussians_pyramid ------------------------
@pytest.mark.parametrize("name,w,h,fo", [
    ("tex", 640, 480, 0), ("tex", 333, 251, 0), ("tex", 320, 240, -1), ("tex", 200, 150, 1),
    ("grad", 480, 270, 0), ("tex", 64, 48, 0),
])
def test_pyramid_bit_exact(ctx, name, w, h, fo):
    img = S.tex(w, h, 11) if name == "tex" else S.grad(w, h)
    p_gpu, p_ref = _pp(fo)
    ref = O.compute_dog_extrema(img, p_ref)
    ctx.pyramid_enqueue(0, img, p_gpu)
    ctx.wait(0)
    assert_pyramids_identical(ctx, ref)


@pytest.fixture(scope="module")
def wide_ctx():
    c = sb.SiftContext(2048, 1200, max_keypoints=65536)
    yield c
    c.close()


# Sizes that exercise the fused octave kernel's strips and segments: several strips
# of both tile widths, odd widths/heights (ragged last strip, straddling column
# pair), segment seams, and images narrower than the cascade halo.
@pytest.mark.parametrize("mode", ["stage", "fused", "auto"])
@pytest.mark.parametrize("w,h", [(1300, 420), (1281, 333), (2000, 300), (1025, 1100), (1920, 1080),
                                 (129, 700), (257, 97), (90, 1000), (1000, 41), (31, 31)])
def test_pyramid_kernels_bit_exact_on_ragged_shapes(wide_ctx, w, h, mode):
    """Every pyramid implementation, selected EXPLICITLY, on shapes that hit strip / segment
    seams, ragged last strips and images narrower than the cascade halo."""
    img = S.tex(w, h, 7)
    p_gpu, p_ref = _pp(0)
    ref = O.compute_dog_extrema(img, p_ref)
    wide_ctx.set_pyramid_mode(mode)
    try:
        wide_ctx.pyramid_enqueue(0, img, p_gpu)
        wide_ctx.wait(0)
        assert_pyramids_identical(wide_ctx, ref)
    finally:
        wide_ctx.set_pyramid_mode("auto")


@pytest.mark.parametrize("mode", ["stage", "fused", "auto"])
def test_pyramid_kernels_noise_and_negative_values(wide_ctx, mode):
    # white noise (every tap matters) and signed data (signed zeros / cancellation)
    rng = np.random.default_rng(5)
    wide_ctx.set_pyramid_mode(mode)
    try:
        for img in (rng.random((300, 1111), dtype=np.float32),
                    (rng.standard_normal((257, 640)) * 3).astype(np.float32)):
            p_gpu, p_ref = _pp(0)
            ref = O.compute_dog_extrema(img, p_ref)
            wide_ctx.pyramid_enqueue(0, img, p_gpu)
            wide_ctx.wait(0)
            assert_pyramids_identical(wide_ctx, ref)
    finally:
        wide_ctx.set_pyramid_mode("auto")


def test_pyramid_other_schedules(ctx):
    img = S.tex(256, 192, 5)
    for kw in (dict(scale_count_per_octave=5, scale_geometric_factor=float(np.float32(2.0) ** np.float32(0.5))),
               dict(scale_count_per_octave=4, scale_geometric_factor=2.0),
               dict(scale_count_per_octave=7, scale_geometric_factor=float(np.float32(2.0) ** np.float32(0.25))),
               dict(num_octaves_max=2), dict(image_padding_size=4), dict(scale_camera=1.7)):
        p_gpu, p_ref = _pp(0, **kw)
        ref = O.compute_dog_extrema(img, p_ref)
        ctx.pyramid_enqueue(0, img, p_gpu)
        ctx.wait(0)
        assert_pyramids_identical(ctx, ref)


# ---- ComputeDoGExtrema -----------------------------------------------------------
@pytest.mark.parametrize("w,h,fo,pad,it", [(640, 480, 0, 1, 5), (640, 480, 0, 5, 5), (400, 300, -1, 1, 5),
                                           (333, 251, 0, 2, 1), (512, 384, 0, 1, 0)])
def test_extrema_identical(ctx, w, h, fo, pad, it):
    img = S.tex(w, h, 21)
    p_gpu, p_ref = _pp(fo)
    ref = O.compute_dog_extrema(img, p_ref, 4.0, 0.01, 10.0, pad, it)
    e = ctx.dog_extrema(img, p_gpu, 4.0, 0.01, 10.0, pad, it)
    assert len(ref.extrema) > 20
    assert_extrema_identical(e, ref.extrema)
    assert_pyramids_identical(ctx, ref)


def test_dog_plateau(ctx):
    # FeatureDetectors/test_featuredetectors_dog.cpp:45-100 through the ComputeDoGExtrema mirror
    N = 11
    I = np.zeros((N, N), np.float32)
    I[3:8, 3:8] = 1
    pp = sb.ImagePyramidParams(0, 6, float(np.float32(2.0) ** np.float32(1.0 / 3)), 1, 1.0, 1.6)
    det = sb.ComputeDoGExtrema(pp, 1e-6, 1e-6)
    feats, so = det(I)
    assert len(feats) > 0
    z = det._ctx.octave_scaling_factor(int(so[0][1]))
    assert abs(feats[0]["x"] * z - 5) < 1e-2 and abs(feats[0]["y"] * z - 5) < 1e-2
    ref = O.compute_dog_extrema(I, O.PyramidParams(0, 6, pp.scale_geometric_factor, 1, 1.0, 1.6), 1e-6, 1e-6)
    assert_extrema_identical(feats, ref.extrema)


def test_too_few_scales_raises(ctx):
    with pytest.raises(RuntimeError):
        ctx.dog_extrema(np.zeros((32, 32), np.float32), sb.ImagePyramidParams(0, 3))
    with pytest.raises(RuntimeError):
        sb.ComputeDoGExtrema(sb.ImagePyramidParams(0, 3))


# ---- compute_sift_keypoints ---------------------------------------------------------
@pytest.mark.parametrize("w,h,fo,seed", [(640, 480, 0, 31), (800, 600, 0, 32), (400, 300, -1, 33), (1024, 768, 0, 34)])
def test_sift_keypoints(ctx, w, h, fo, seed):
    img = S.tex(w, h, seed)
    p_gpu, p_ref = _pp(fo)
    ref = O.compute_sift_keypoints(img, p_ref, parallel=True)
    kl = ctx.compute_sift_keypoints(img, p_gpu)
    assert_pyramids_identical(ctx, ref)
    assert_extrema_identical(ctx.extrema(), ref.extrema)
    assert len(ref.keypoints) > 50
    stats = compare_keypoints(kl.features, kl.descriptors, ref.keypoints, ref.descriptors, ref, ctx.oriented())
    print(stats)
    # octave-coordinate list (before SIFT.cpp:92-98) matches too
    ko = ctx.oriented()
    assert len(ko) == len(kl.features)
    assert abs(len(kl.features) - len(ref.keypoints)) <= max(1, len(ref.keypoints) // 1000)


def test_stinkbug_golden(ctx):
    """Config C1: data/stinkbug.png (500x375), default parameters (first octave -1)."""
    gray = np.load(os.path.join(GOLDEN, "stinkbug_gray.npy"))
    gold = np.load(os.path.join(GOLDEN, "stinkbug_oracle.npz"))
    kl = ctx.compute_sift_keypoints(gray, sb.ImagePyramidParams())
    assert ctx.num_octaves() == int(gold["num_octaves"])
    assert_extrema_identical(ctx.extrema(), gold["extrema"])
    stats = compare_keypoints(kl.features, kl.descriptors, gold["keypoints"], gold["descriptors"])
    print(stats)
    # and the live oracle still agrees with its committed output
    ref = O.compute_sift_keypoints(gray, O.PyramidParams(), parallel=True)
    assert ref.keypoints.tobytes() == gold["keypoints"].tobytes()


def test_module_level_api_and_edge_cases():
    # python/oddkiva/sara/pybind11/test/test_sfm.py:16-21
    kl = sb.compute_sift_keypoints(np.zeros((24, 32), np.float32), sb.ImagePyramidParams(first_octave_index=0))
    assert len(sb.features(kl)) == 0 and sb.descriptors(kl).shape == (0, 128)
    # constant image, tiny images, and an image smaller than one octave
    for shape in [(1, 1), (2, 3), (5, 5), (16, 16)]:
        kl = sb.compute_sift_keypoints(np.full(shape, 0.5, np.float32), sb.ImagePyramidParams(first_octave_index=0))
        ref = O.compute_sift_keypoints(np.full(shape, 0.5, np.float32), O.PyramidParams(first_octave_index=0))
        assert len(kl) == len(ref.keypoints)
    with pytest.raises(ValueError):
        sb.compute_sift_keypoints(np.zeros((4, 4, 3), np.float32))


def test_device_input_and_slots(ctx):
    import torch

    imgs = [S.tex(512, 384, 40 + i) for i in range(2)]
    p_gpu, p_ref = _pp(0)
    dev = [torch.from_numpy(i).cuda() for i in imgs]
    for slot in range(2):
        ctx.enqueue(slot, dev[slot], p_gpu)
    outs = [ctx.collect(slot) for slot in range(2)]
    for img, kl in zip(imgs, outs):
        host = ctx.compute_sift_keypoints(img, p_gpu)
        assert host.features.tobytes() == kl.features.tobytes()       # deterministic, host == device input
        assert host.descriptors.tobytes() == kl.descriptors.tobytes()


def test_capacity_overflow_is_reported():
    c = sb.SiftContext(640, 480, max_keypoints=64)
    with pytest.raises(sb.SaraB200Error) as ei:
        c.compute_sift_keypoints(S.tex(640, 480, 3), sb.ImagePyramidParams(first_octave_index=0))
    assert ei.value.code == -5
    c.close()


# ---- full SIFT on other parameter sets (generic kernels, other thresholds) -------------
@pytest.mark.parametrize("kw,args", [
    (dict(scale_count_per_octave=5, scale_geometric_factor=float(np.float32(2.0) ** np.float32(0.5))), {}),
    (dict(scale_count_per_octave=7, scale_geometric_factor=float(np.float32(2.0) ** np.float32(0.25))), {}),
    (dict(num_octaves_max=3), dict(extremum_thres=0.02, edge_ratio_thres=5.0)),
    (dict(), dict(extremum_thres=0.003, extremum_refinement_iter=2)),
    (dict(scale_initial=0.8), {}), (dict(scale_initial=1.0), {}), (dict(scale_initial=1.2), {}),
    (dict(scale_initial=1.0, scale_camera=0.25), dict(extremum_thres=0.005)),
])
def test_sift_other_parameters(ctx, kw, args):
    img = S.tex(512, 384, 91)
    p_gpu, p_ref = _pp(0, **kw)
    a = dict(gauss_truncate=4.0, extremum_thres=0.01, edge_ratio_thres=10.0, extremum_refinement_iter=5)
    a.update(args)
    ref = O.compute_sift_keypoints(img, p_ref, a["gauss_truncate"], a["extremum_thres"], a["edge_ratio_thres"],
                                   a["extremum_refinement_iter"], parallel=True)
    kl = ctx.compute_sift_keypoints(img, p_gpu, a["gauss_truncate"], a["extremum_thres"], a["edge_ratio_thres"],
                                    a["extremum_refinement_iter"])
    assert_pyramids_identical(ctx, ref)
    assert_extrema_identical(ctx.extrema(), ref.extrema)
    assert len(ref.keypoints) > 20
    compare_keypoints(kl.features, kl.descriptors, ref.keypoints, ref.descriptors, ref, ctx.oriented())


@pytest.mark.parametrize("fo,s0", [(0, 1.0), (-1, 1.0), (0, 0.8), (-1, 1.6)])
def test_descriptors_of_small_scale_keypoints(ctx, fo, s0):
    """Keypoints whose descriptor window is narrower than a warp (round(10.6 sigma) <= 15, i.e.
    refined sigma < 1.46): EVERY one of them must be inside the descriptor tolerance (the
    window walk of descriptor_kernel used to assume side >= 33)."""
    img = S.tex(640, 480, 123)
    p_gpu, p_ref = _pp(fo, scale_initial=s0)
    ref = O.compute_sift_keypoints(img, p_ref, parallel=True)
    kl = ctx.compute_sift_keypoints(img, p_gpu)
    assert_extrema_identical(ctx.extrema(), ref.extrema)
    from parity import match_oriented, DESC_REL, DESC_ABS

    ko_gpu, ko_ref = ctx.oriented(), ref.oriented
    ig, ir, bad, n_ext = match_oriented(ko_gpu, ko_ref)
    sigma = 1.0 / np.sqrt(ko_ref["shape"][ir, 0])
    small = sigma < 1.46
    if s0 <= 1.0:
        assert small.sum() > 100, small.sum()
    if small.sum() == 0:
        return
    dth = np.abs(ko_gpu["orientation"][ig] - ko_ref["orientation"][ir])
    dth = np.minimum(dth, 2 * np.pi - dth)
    da, db = kl.descriptors[ig].astype(np.float64), ref.descriptors[ir].astype(np.float64)
    err = np.linalg.norm(da - db, axis=1)
    tol = DESC_REL * np.linalg.norm(db, axis=1) + DESC_ABS
    ok = (err <= tol) | (dth > 1e-3)
    assert ok[small].all(), f"{(~ok[small]).sum()} of {small.sum()} small-scale descriptors off, max {err[small].max()}"


@pytest.mark.parametrize("i", [0, 7, 100])
def test_c4_sequence_frames_vs_oracle(i):
    """Configs C4 / C5: frames of the translated-scene sequence, full SIFT against the oracle."""
    img = S.sequence_frame(1920, 1080, i)
    c = sb.SiftContext(1920, 1080, max_keypoints=65536, min_first_octave_index=0)
    try:
        p_gpu, p_ref = _pp(0)
        ref = O.compute_sift_keypoints(img, p_ref, parallel=True)
        kl = c.compute_sift_keypoints(img, p_gpu)
        assert_pyramids_identical(c, ref)
        assert_extrema_identical(c.extrema(), ref.extrema)
        assert len(ref.keypoints) > 1000
        print(compare_keypoints(kl.features, kl.descriptors, ref.keypoints, ref.descriptors, ref, c.oriented()))
    finally:
        c.close()


@pytest.mark.parametrize("w,h", [(2000, 40), (40, 1500), (641, 479), (1023, 767)])
def test_sift_ragged_shapes(ctx, w, h):
    c = sb.SiftContext(max(w, 64), max(h, 64), max_keypoints=65536)
    try:
        img = S.tex(w, h, 17)
        p_gpu, p_ref = _pp(0)
        ref = O.compute_sift_keypoints(img, p_ref, parallel=True)
        kl = c.compute_sift_keypoints(img, p_gpu)
        assert_pyramids_identical(c, ref)
        assert_extrema_identical(c.extrema(), ref.extrema)
        if len(ref.keypoints):
            compare_keypoints(kl.features, kl.descriptors, ref.keypoints, ref.descriptors, ref, c.oriented())
    finally:
        c.close()


def test_misaligned_device_image_takes_the_fallback(ctx):
    """A device image that is not 16-byte aligned cannot be staged by TMA: the pre-blur must fall
    back to the generic kernel and still give the same bits."""
    import torch

    img = S.tex(640, 480, 55)
    buf = torch.zeros(640 * 480 + 1, dtype=torch.float32, device="cuda")
    buf[1:] = torch.from_numpy(img).cuda().flatten()
    view = buf[1:].view(480, 640)
    assert view.data_ptr() % 16 == 4
    p_gpu, p_ref = _pp(0)
    a = ctx.compute_sift_keypoints(view, p_gpu)
    b = ctx.compute_sift_keypoints(img, p_gpu)
    assert a.features.tobytes() == b.features.tobytes() and a.descriptors.tobytes() == b.descriptors.tobytes()


def test_many_slots_mixed_sizes_repeatable():
    """Frames of different sizes in flight on 6 slots, three rounds: every result must equal the
    single-slot result of the same frame (arena reuse, side streams, device-side queues)."""
    sizes = [(640, 480), (333, 251), (800, 600), (1024, 768), (200, 150), (641, 479)]
    imgs = [S.tex(w, h, 60 + i) for i, (w, h) in enumerate(sizes)]
    pp = sb.ImagePyramidParams(first_octave_index=0)
    c = sb.SiftContext(1024, 768, max_keypoints=65536, num_slots=6)
    try:
        single = [c.compute_sift_keypoints(im, pp) for im in imgs]
        for _ in range(3):
            for slot, im in enumerate(imgs):
                c.enqueue(slot, im, pp)
            for slot in reversed(range(len(imgs))):
                kl = c.collect(slot)
                assert kl.features.tobytes() == single[slot].features.tobytes()
                assert kl.descriptors.tobytes() == single[slot].descriptors.tobytes()
    finally:
        c.close()


def test_stage_functors_on_supplied_extrema():
    """ComputeDominantOrientations / ComputeSIFTDescriptor entry point: the extrema of ComputeDoGExtrema handed
    back by the caller give exactly the keypoints and descriptors of the whole compute_sift_keypoints call;
    a subset gives the matching subset; a bad (s, o) pair is rejected."""
    img = S.tex(800, 600, 4321)
    ctx = sb.SiftContext(800, 600, device=0)
    pp = sb.ImagePyramidParams(first_octave_index=0)
    kl = ctx.compute_sift_keypoints(img, pp)
    ext = ctx.extrema()
    # N1: compute_sift_keypoints passes its refinement-iteration argument as the padding
    e2 = ctx.dog_extrema(img, pp, 4.0, 0.01, 10.0, 5, 5)
    assert e2.tobytes() == ext.tobytes()
    ori, kl2 = ctx.describe_extrema(e2)
    assert len(kl2) == len(kl) > 300
    assert kl2.features.tobytes() == kl.features.tobytes()
    assert kl2.descriptors.tobytes() == kl.descriptors.tobytes()
    assert np.array_equal(ori["orientation"], kl.features["orientation"])
    sub = e2[::7]
    _, kl3 = ctx.describe_extrema(sub)
    keep = np.isin(kl.features["xi"] * 100000 + kl.features["yi"] + 1e7 * kl.features["s"] + 1e8 * kl.features["o"],
                   sub["xi"] * 100000 + sub["yi"] + 1e7 * sub["s"] + 1e8 * sub["o"])
    assert kl3.descriptors.tobytes() == kl.descriptors[keep].tobytes()
    bad = e2[:3].copy()
    bad["o"][1] = 99
    with pytest.raises(ValueError):
        ctx.describe_extrema(bad)
    assert len(ctx.describe_extrema(e2[:0])[1]) == 0
    ctx.close()
