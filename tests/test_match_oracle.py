"""Matching row, CPU side: the restated AnnMatcher search against the reference's own
known-answer test, against the golden vectors produced by the reference's vendored FLANN
(tests/golden/match_flann.npz, tests/golden/make_match_fixture.py) and -- when oracle/_ref is built --
against that library directly."""
import os

import numpy as np
import pytest

from oracle import match as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "match_flann.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_reference_kat_ann_matching():
    """cpp/test/Sara/FeatureMatching/test_featurematching_matching.cpp:27-57: one key (0, 0) against ten keys
    (i, i), ratio 0.6 -> exactly one match {0, 0} with score 0."""
    d1 = np.zeros((1, 2), np.float32)
    d2 = np.repeat(np.arange(10, dtype=np.float32)[:, None], 2, axis=1)
    m = M.ann_match(d1, d2, 0.6)
    assert len(m) == 1
    assert (m[0]["x_index"], m[0]["y_index"]) == (0, 0) and m[0]["score"] == 0.0


def test_l2_functor_groups_of_four():
    """flann::L2 (dist.h:151-178): result += (d0^2 + d1^2 + d2^2 + d3^2) per group, then the tail one by one."""
    rng = np.random.default_rng(3)
    for dim in (1, 2, 3, 4, 7, 128, 130):
        a, b = rng.normal(0, 50, dim).astype(np.float32), rng.normal(0, 50, dim).astype(np.float32)
        r = np.float32(0)
        i = 0
        while i + 3 < dim:
            d = (a[i:i + 4] - b[i:i + 4]).astype(np.float32)
            g = np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2]))
            g = np.float32(g + np.float32(d[3] * d[3]))
            r = np.float32(r + g)
            i += 4
        while i < dim:
            d = np.float32(a[i] - b[i])
            r = np.float32(r + np.float32(d * d))
            i += 1
        assert np.float32(M.l2_flann(a, b)).view(np.uint32) == r.view(np.uint32)


def test_port_equals_golden_flann_linear(gold):
    d1, d2 = gold["d1"], gold["d2"]
    for q, data, tag in ((d1, d2, "12"), (d2, d1, "21")):
        idx, dist = M.knn_linear(data, q, 3)
        assert np.array_equal(idx, gold[f"lin_idx_{tag}"])
        assert np.array_equal(bits(dist), bits(gold[f"lin_dist_{tag}"]))


@pytest.mark.parametrize("thr", [0.6, 1.0, 1.2])
def test_port_matches_equal_golden(gold, thr):
    m = M.ann_match(gold["d1"], gold["d2"], thr, gold["f1"], gold["f2"])
    g = gold[f"matches_lin_{thr}"]
    assert len(m) == len(g)
    for k in ("x_index", "y_index", "rank", "direction"):
        assert np.array_equal(m[k], g[k]), k
    assert np.array_equal(bits(m["score"]), bits(g["score"]))


def test_port_self_matches_equal_golden(gold):
    m = M.ann_match(gold["d1"], gold["d1"], 1.2, gold["f1"], gold["f1"], self_matching=True)
    g = gold["self_matches_lin_1.2"]
    assert len(m) == len(g) and np.array_equal(m["x_index"], g["x_index"]) and np.array_equal(m["y_index"], g["y_index"])
    assert np.array_equal(bits(m["score"]), bits(g["score"]))
    # ratio <= 1: the loop of AnnMatcher.cpp:149 starts at rank 1 with K = 1 -> no self match at all
    assert len(M.ann_match(gold["d1"], gold["d1"], 0.9, gold["f1"], gold["f1"], self_matching=True)) == 0


def test_kdtree_forest_is_an_approximation_of_the_exact_search(gold):
    """What the reference really runs (KDTreeIndexParams{8}, 32 checks) finds the exact nearest neighbour for
    most keys; wherever it does, its distance carries the same bits as the exact search."""
    same = gold["kd_idx_12"][:, 0] == gold["lin_idx_12"][:, 0]
    assert same.mean() > 0.9
    assert np.array_equal(bits(gold["kd_dist_12"][same, 0]), bits(gold["lin_dist_12"][same, 0]))
    # and the match lists at ratio 0.6 agree on almost every pair
    a = {(int(m["x_index"]), int(m["y_index"])) for m in gold["matches_lin_0.6"]}
    b = {(int(m["x_index"]), int(m["y_index"])) for m in gold["matches_kd_0.6"]}
    assert len(a & b) >= 0.95 * len(a)


def test_boundary_cases():
    rng = np.random.default_rng(0)
    d1 = rng.normal(0, 1, (5, 8)).astype(np.float32)
    # one indexed key: score 1 kept only when 1 < ratio^2 (AnnMatcher.cpp:88-103), both directions
    one = rng.normal(0, 1, (1, 8)).astype(np.float32)
    assert len(M.ann_match(d1, one, 0.6)) == 0
    m = M.ann_match(d1, one, 1.2)
    assert sorted(m["x_index"].tolist()) == [0, 1, 2, 3, 4] and np.all(m["score"] <= 1.0)
    with pytest.raises(RuntimeError):
        M.ann_match(d1, np.zeros((0, 8), np.float32), 0.6)
    # equal distances: the lower index comes first (KNNSimpleResultSet::addPoint)
    data = np.zeros((4, 8), np.float32)
    idx, dist = M.knn_linear(data, d1[:1], 3)
    assert idx.tolist() == [[0, 1, 2]]


@pytest.mark.skipif(not M.have_ref(), reason="oracle/_ref/libflann_ref.so not built (needs /root/reference)")
def test_port_equals_real_flann_linear_on_random_sets():
    rng = np.random.default_rng(11)
    for n1, n2, dim in ((300, 257, 128), (64, 500, 128), (50, 40, 2), (33, 70, 131)):
        a = (rng.random((n1, dim)) * 255).astype(np.float32)
        m = min(n1, n2 // 2)
        b = np.vstack([a[:m] + rng.normal(0, 4, (m, dim)), rng.random((n2 - m, dim)) * 255]).astype(np.float32)
        ref = M.FlannRef(b, "linear")
        i0, d0 = M.knn_linear(b, a, 3)
        i1, d1 = ref.knn(a, 3)
        assert np.array_equal(i0, i1) and np.array_equal(bits(d0), bits(d1))
        for thr in (0.6, 1.2):
            assert np.array_equal(M.ann_match(a, b, thr), M.ann_match(a, b, thr, backend="linear"))
