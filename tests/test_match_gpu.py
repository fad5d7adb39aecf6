"""Matching row on the GPU, through the C ABI: sara_b200_knn / sara_b200_compute_matches
against the oracle (bit-identical distances, identical indices and match lists) and against the golden vectors
of the reference's vendored FLANN."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "match_flann.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ctx():
    import sara_b200 as sb

    c = sb.SiftContext(64, 64, device=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def sift_like(rng, n, dim=128):
    """Non-negative, clipped, norm 512 rows: the statistics of ComputeSIFTDescriptor's output."""
    d = rng.gamma(0.6, 1.0, (n, dim)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = np.minimum(d, 0.2)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.minimum(d * 512, 255).astype(np.float32)


def assert_matches_equal(m, g):
    assert len(m) == len(g)
    for k in ("x_index", "y_index", "rank", "direction"):
        assert np.array_equal(m[k], g[k]), k
    assert np.array_equal(bits(m["score"]), bits(g["score"]))


def test_reference_kat_on_gpu(ctx):
    """test_featurematching_matching.cpp:27-57 through the C ABI (2-D features: the scalar kernel)."""
    d1 = np.zeros((1, 2), np.float32)
    d2 = np.repeat(np.arange(10, dtype=np.float32)[:, None], 2, axis=1)
    m = ctx.compute_matches(d1, d2, 0.6)
    assert len(m) == 1 and (m[0]["x_index"], m[0]["y_index"]) == (0, 0) and m[0]["score"] == 0.0


@pytest.mark.parametrize("mode", ["scalar", "tensor"])
def test_knn_equals_golden_flann_linear(ctx, gold, mode):
    for q, data, tag in ((gold["d1"], gold["d2"], "12"), (gold["d2"], gold["d1"], "21")):
        idx, dist, st = ctx.knn(q, data, 3, mode=mode)
        assert st["used_tensor_cores"] == (1 if mode == "tensor" else 0)
        assert np.array_equal(idx, gold[f"lin_idx_{tag}"])
        assert np.array_equal(bits(dist), bits(gold[f"lin_dist_{tag}"]))


@pytest.mark.parametrize("mode", ["scalar", "tensor"])
@pytest.mark.parametrize("thr", [0.6, 1.0, 1.2])
def test_matches_equal_golden(ctx, gold, thr, mode):
    m = ctx.compute_matches(gold["d1"], gold["d2"], thr, gold["f1"], gold["f2"], mode=mode)
    assert_matches_equal(m, gold[f"matches_lin_{thr}"])


def test_self_matches_equal_golden(ctx, gold):
    m = ctx.compute_matches(gold["d1"], gold["d1"], 1.2, gold["f1"], gold["f1"], self_matching=True)
    assert_matches_equal(m, gold["self_matches_lin_1.2"])
    assert len(ctx.compute_matches(gold["d1"], gold["d1"], 0.9, gold["f1"], gold["f1"], self_matching=True)) == 0


@pytest.mark.parametrize("n1,n2", [(1, 3), (127, 129), (128, 128), (1000, 777), (5000, 4100)])
def test_tensor_path_vs_oracle(ctx, n1, n2):
    """Descriptor-like random sets of ragged sizes: the tcgen05 candidates + exact re-ranking give the oracle's
    neighbours and distance bits; the scalar kernel agrees."""
    from oracle import match as M

    rng = np.random.default_rng(n1 * 7 + n2)
    a = sift_like(rng, n1)
    m = min(n1, n2 // 2)
    b = np.vstack([a[:m] + rng.normal(0, 6, (m, 128)), sift_like(rng, n2 - m)]).astype(np.float32)
    i0, d0 = M.knn_linear(b, a, 3)
    for mode in ("tensor", "scalar"):
        idx, dist, st = ctx.knn(a, b, 3, mode=mode)
        assert np.array_equal(idx, i0), mode
        assert np.array_equal(bits(dist), bits(d0)), mode
    # few queries should need the exact fallback on data like this
    _, _, st = ctx.knn(a, b, 3, mode="tensor")
    if n2 >= 4000:
        assert st["n_redone"] <= max(2, n1 // 100), st


def test_degenerate_sets_fall_back_and_stay_exact(ctx):
    """Many equal distances (duplicates, zero rows): the certificate cannot hold, the exact kernel takes over;
    equal distances come out in index order like KNNSimpleResultSet."""
    from oracle import match as M

    rng = np.random.default_rng(5)
    base = sift_like(rng, 10)
    b = np.vstack([base] * 48 + [np.zeros((30, 128), np.float32)]).astype(np.float32)  # every row 48 times
    a = np.vstack([base, sift_like(rng, 15)])
    i0, d0 = M.knn_linear(b, a, 3)
    idx, dist, st = ctx.knn(a, b, 3, mode="tensor")
    assert np.array_equal(idx, i0) and np.array_equal(bits(dist), bits(d0))
    assert st["n_redone"] > 0
    idx, dist, _ = ctx.knn(a, b, 8, mode="tensor")
    i8, d8 = M.knn_linear(b, a, 8)
    assert np.array_equal(idx, i8) and np.array_equal(bits(dist), bits(d8))


@pytest.mark.parametrize("dim", [1, 2, 3, 64, 130, 256])
def test_scalar_kernel_any_dimension(ctx, dim):
    from oracle import match as M

    rng = np.random.default_rng(dim)
    a = rng.normal(0, 30, (150, dim)).astype(np.float32)
    b = rng.normal(0, 30, (333, dim)).astype(np.float32)
    i0, d0 = M.knn_linear(b, a, 3)
    idx, dist, _ = ctx.knn(a, b, 3)
    assert np.array_equal(idx, i0) and np.array_equal(bits(dist), bits(d0))
    assert_matches_equal(ctx.compute_matches(a, b, 0.9), M.ann_match(a, b, 0.9))


def test_fewer_points_than_k(ctx):
    rng = np.random.default_rng(1)
    a, b = sift_like(rng, 10), sift_like(rng, 2)
    idx, dist, _ = ctx.knn(a, b, 3)
    assert np.all(idx[:, 2] == -1) and np.all(dist[:, 2] == np.finfo(np.float32).max)
    assert set(idx[0, :2].tolist()) == {0, 1}


def test_boundary_cases_and_errors(ctx):
    from oracle import match as M

    rng = np.random.default_rng(0)
    d1 = sift_like(rng, 5)
    one = sift_like(rng, 1)
    assert len(ctx.compute_matches(d1, one, 0.6)) == 0
    assert_matches_equal(ctx.compute_matches(d1, one, 1.2), M.ann_match(d1, one, 1.2))
    two = sift_like(rng, 2)
    assert_matches_equal(ctx.compute_matches(d1, two, 0.95), M.ann_match(d1, two, 0.95))
    with pytest.raises(ValueError):  # "the list of key-points is empty" (AnnMatcher.cpp:45-46)
        ctx.compute_matches(d1, np.zeros((0, 128), np.float32), 0.6)


def test_device_resident_descriptors_and_sift_frames(ctx):
    """Two frames of the synthetic sequence: SIFT on the GPU, descriptors never leave the device, matches equal
    the oracle's AnnMatcher on the downloaded copies; AnnMatcher class mirrors the reference's constructor."""
    import torch

    import sara_b200 as sb
    from oracle import match as M
    from sara_b200 import synthetic as S

    c = sb.SiftContext(1280, 720, device=0, num_slots=2)
    pp = sb.ImagePyramidParams(first_octave_index=0)
    kls = [c.compute_sift_keypoints(S.sequence_frame(1280, 720, i), pp) for i in (0, 1)]
    dev = [torch.from_numpy(k.descriptors).cuda() for k in kls]
    m_dev = c.compute_matches(dev[0], dev[1], 0.6, kls[0].features, kls[1].features)
    m_ref = M.ann_match(kls[0].descriptors, kls[1].descriptors, 0.6, kls[0].features, kls[1].features)
    assert len(m_ref) > 100
    assert_matches_equal(m_dev, m_ref)
    m_cls = sb.AnnMatcher(kls[0], kls[1], 0.6, ctx=c).compute_matches()
    assert_matches_equal(m_cls, m_ref)
    # the sequence is a pure translation by (3, 1.5) px per frame: the best matches must say so
    best = m_ref[:100]
    dx = kls[1].features["x"][best["y_index"]] - kls[0].features["x"][best["x_index"]]
    dy = kls[1].features["y"][best["y_index"]] - kls[0].features["y"][best["x_index"]]
    assert abs(np.median(dx) - 3.0) < 0.5 and abs(np.median(dy) - 1.5) < 0.5
    c.close()
