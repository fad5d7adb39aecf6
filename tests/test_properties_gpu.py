"""Full-size checks (BASELINE.json configurations C2, C3): parity against the CPU oracle at
4K / 1080p where it still finishes in seconds, plus size-independent properties of the
CUDA path (the three pyramid implementations agree bit for bit, D = G(s+1) - G(s), the
next octave is the even sub-sampling of G(2), extrema come out in raster order, the whole
chain is deterministic)."""
import numpy as np
import pytest

from oracle import oracle as O
from sara_b200 import synthetic as S
import sara_b200 as sb
from parity import assert_extrema_identical, assert_pyramids_identical, compare_keypoints

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx4k():
    c = sb.SiftContext(3840, 2160, max_keypoints=131072)
    yield c
    c.close()


def _layers(ctx):
    return [[ctx.gaussian_layer(s, o).copy() for s in range(ctx.num_scales())] for o in range(ctx.num_octaves())], \
           [[ctx.dog_layer(s, o).copy() for s in range(ctx.num_scales() - 1)] for o in range(ctx.num_octaves())]


def test_c2_1080p_gradient_pyramid_vs_oracle(ctx4k):
    """C2: 1920x1080 synthetic gradient image, 4 octaves, 3 + 3 scales per octave: pyramid + DoG."""
    img = S.grad(1920, 1080)
    kw = dict(first_octave_index=0, num_octaves_max=4)
    ref = O.compute_dog_extrema(img, O.PyramidParams(**kw))
    ctx4k.pyramid_enqueue(0, img, sb.ImagePyramidParams(**kw))
    ctx4k.wait(0)
    assert ctx4k.num_octaves() == 4
    assert_pyramids_identical(ctx4k, ref)


def test_c3_4k_full_sift_vs_oracle(ctx4k):
    """C3: 3840x2160 synthetic frame, full SIFT; descriptor L2 tolerance as stated in parity.py."""
    img = S.tex(3840, 2160, 1234)
    ref = O.compute_sift_keypoints(img, O.PyramidParams(first_octave_index=0), parallel=True)
    kl = ctx4k.compute_sift_keypoints(img, sb.ImagePyramidParams(first_octave_index=0))
    assert ctx4k.num_octaves() == 10
    assert_pyramids_identical(ctx4k, ref)
    assert_extrema_identical(ctx4k.extrema(), ref.extrema)
    assert len(ref.keypoints) > 5000
    stats = compare_keypoints(kl.features, kl.descriptors, ref.keypoints, ref.descriptors, ref, ctx4k.oriented())
    print(stats)


def test_pyramid_implementations_agree_at_4k(ctx4k):
    """The generic per-scale kernel (oracle-checked at every small size), the TMA marching
    kernel and the fused octave kernel must produce the same bits on a full 4K frame."""
    rng = np.random.default_rng(3)
    img = rng.random((2160, 3840), dtype=np.float32)
    pp = sb.ImagePyramidParams(first_octave_index=0)
    got = {}
    for mode in ("generic", "stage", "fused"):
        ctx4k.set_pyramid_mode(mode)
        ctx4k.pyramid_enqueue(0, img, pp)
        ctx4k.wait(0)
        got[mode] = _layers(ctx4k)
    ctx4k.set_pyramid_mode("auto")
    for mode in ("stage", "fused"):
        for o in range(len(got["generic"][0])):
            for s, (a, b) in enumerate(zip(got["generic"][0][o], got[mode][0][o])):
                assert a.tobytes() == b.tobytes(), f"{mode}: G({s},{o}) differs from the generic kernel"
            for s, (a, b) in enumerate(zip(got["generic"][1][o], got[mode][1][o])):
                assert a.tobytes() == b.tobytes(), f"{mode}: D({s},{o}) differs from the generic kernel"


def test_pyramid_structure_properties_at_4k(ctx4k):
    img = S.tex(3840, 2160, 77)
    ctx4k.pyramid_enqueue(0, img, sb.ImagePyramidParams(first_octave_index=0))
    ctx4k.wait(0)
    G, D = _layers(ctx4k)
    for o in range(len(G)):
        for s in range(len(D[o])):
            # difference_of_gaussians_pyramid, GaussianPyramid.cpp:23-51: an exact fp32 subtraction
            assert (G[o][s + 1] - G[o][s]).tobytes() == D[o][s].tobytes()
        if o + 1 < len(G):
            # downscale(G(2, o), 2), GaussianPyramid.hpp:114
            h, w = G[o + 1][0].shape
            assert G[o][2][: 2 * h : 2, : 2 * w : 2].tobytes() == G[o + 1][0].tobytes()
        # smoothing a [0, 1] image keeps it inside [0, 1] up to rounding
        for g in G[o]:
            assert g.min() >= -1e-6 and g.max() <= 1 + 1e-5


def test_constant_image_has_flat_pyramid_and_no_keypoints(ctx4k):
    img = np.full((1080, 1920), 0.25, np.float32)
    kl = ctx4k.compute_sift_keypoints(img, sb.ImagePyramidParams(first_octave_index=0))
    assert len(kl) == 0
    for o in range(ctx4k.num_octaves()):
        for s in range(ctx4k.num_scales() - 1):
            assert np.abs(ctx4k.dog_layer(s, o)).max() < 1e-6


def test_extrema_order_and_determinism_at_4k(ctx4k):
    img = S.tex(3840, 2160, 4321)
    pp = sb.ImagePyramidParams(first_octave_index=0)
    a = ctx4k.compute_sift_keypoints(img, pp)
    e = ctx4k.extrema().copy()
    b = ctx4k.compute_sift_keypoints(img, pp)
    # deterministic: same bits on every run (no floating-point atomics anywhere)
    assert a.features.tobytes() == b.features.tobytes()
    assert a.descriptors.tobytes() == b.descriptors.tobytes()
    # reference order: octave-major, scale-minor, raster (DoG.cpp:70-82, RefineExtremum.cpp:495-515)
    key = (e["o"].astype(np.int64) << 40) | (e["s"].astype(np.int64) << 32) | (e["yi"].astype(np.int64) << 16) | e["xi"]
    assert np.all(np.diff(key) > 0)
    # descriptors: finite and capped at 255 (SIFT.hpp:128).  They may be NEGATIVE: the modf
    # truncation of SIFT.hpp:204-238 (quirk N6) gives negative trilinear weights.
    assert np.isfinite(a.descriptors).all() and a.descriptors.max() <= 255
    assert len(a) > 5000


def test_8k_frame_pyramid_implementations_agree():
    """Largest configuration the reference benchmarks (7680x4320): grid sizes,
    arena offsets and 32-bit index arithmetic at 33 M pixels per layer."""
    c = sb.SiftContext(7680, 4320, max_keypoints=262144)
    try:
        rng = np.random.default_rng(11)
        img = rng.random((4320, 7680), dtype=np.float32)
        pp = sb.ImagePyramidParams(first_octave_index=0)
        ref = None
        for mode in ("generic", "stage"):
            c.set_pyramid_mode(mode)
            c.pyramid_enqueue(0, img, pp)
            c.wait(0)
            got = [c.dog_layer(s, o).copy() for o in (0, 1, c.num_octaves() - 1) for s in (0, 4)] + \
                  [c.gaussian_layer(5, 0).copy()]
            if ref is None:
                ref = got
            else:
                for a, b in zip(ref, got):
                    assert a.tobytes() == b.tobytes()
        kl = c.compute_sift_keypoints(img[:, :], pp)
        assert c.num_octaves() == 11 and len(kl) >= 0
    finally:
        c.close()


@pytest.mark.parametrize("size", [(3840, 2160), (1920, 1080), (1300, 420)])
def test_latency_and_throughput_schedules_agree(size):
    """A context made for several frames in flight cuts the layers that cannot fill the machine into
    fewer, taller segments (set_march_schedule in csrc/pyramid_march.cu); the segment height must
    not show in a single bit of the pyramid, nor in the keypoints and descriptors."""
    w, h = size
    img = S.tex(w, h, 97)
    pp = sb.ImagePyramidParams(first_octave_index=0)
    res = []
    for slots in (1, 4):
        c = sb.SiftContext(w, h, max_keypoints=131072, num_slots=slots)
        try:
            kl = c.compute_sift_keypoints(img, pp)
            G, D = _layers(c)
            res.append((kl, G, D))
        finally:
            c.close()
    (ka, Ga, Da), (kb, Gb, Db) = res
    assert len(ka) == len(kb) > 0
    for o in range(len(Ga)):
        for a, b in zip(Ga[o] + Da[o], Gb[o] + Db[o]):
            assert a.tobytes() == b.tobytes()
    assert ka.features.tobytes() == kb.features.tobytes()
    assert ka.descriptors.tobytes() == kb.descriptors.tobytes()


def test_eight_4k_frames_in_flight_equal_the_lone_frames():
    """The bench's regime: eight 4K frames in flight on eight streams (kernels of different frames share the
    SMs, the side streams of all slots are busy).  Every frame must come out exactly as it does alone."""
    pp = sb.ImagePyramidParams(first_octave_index=0)
    imgs = [S.tex(3840, 2160, 500 + i) for i in range(4)]
    c = sb.SiftContext(3840, 2160, max_keypoints=131072, num_slots=8)
    try:
        lone = [c.compute_sift_keypoints(im, pp) for im in imgs]
        for _ in range(3):
            for slot in range(8):
                c.enqueue(slot, imgs[slot % 4], pp)
            for slot in range(8):
                kl = c.collect(slot)
                assert kl.features.tobytes() == lone[slot % 4].features.tobytes()
                assert kl.descriptors.tobytes() == lone[slot % 4].descriptors.tobytes()
    finally:
        c.close()
