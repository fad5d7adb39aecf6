"""Sibling detectors on the same pyramid: ComputeLoGExtrema / ComputeDoHExtrema on the GPU against
the oracle -- bit-identical function pyramids, identical ordered extrema -- and the reference's own tests
(test_featuredetectors_log.cpp, test_featuredetectors_hessian.cpp)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def k3():
    return float(np.float32(2.0) ** np.float32(1.0 / 3.0))


def compare(ctx, ref, n_layers):
    import sara_b200 as sb  # noqa: F401

    assert ctx.num_octaves() == ref.num_octaves
    for o in range(ref.num_octaves):
        for s in range(n_layers):
            a, b = ctx.dog_layer(s, o), ref.dog(s, o)
            assert a.tobytes() == b.tobytes(), (s, o)
    e, r = ctx.extrema(), ref.extrema
    assert len(e) == len(r)
    for k in ("xi", "yi", "s", "o", "extremum_type"):
        assert np.array_equal(e[k], r[k]), k
    for k in ("x", "y", "extremum_value", "shape"):
        assert e[k].tobytes() == r[k].tobytes(), k


def test_reference_log_kat():
    """test_featuredetectors_log.cpp:24-78: a 5 x 5 block of ones in an 11 x 11 image; the first LoG extremum
    sits at the centre."""
    import sara_b200 as sb

    N = 11
    img = np.zeros((N, N), np.float32)
    img[3:8, 3:8] = 1
    pp = sb.ImagePyramidParams(0, 6, k3(), 1, 1.0, 1.6)
    det = sb.ComputeLoGExtrema(pp)
    feats, so = det(img)
    assert len(feats) >= 1
    z = det._ctx.octave_scaling_factor(int(so[0][1]))
    assert abs(feats[0]["x"] * z - 5) < 1e-2 and abs(feats[0]["y"] * z - 5) < 1e-2


def test_reference_doh_smoke_kat():
    """test_featuredetectors_hessian.cpp:35-46: a single bright pixel, default ComputeDoHExtrema: must run; the
    result equals the oracle's."""
    import sara_b200 as sb
    from oracle import oracle as O

    N = 21
    img = np.zeros((N, N), np.float32)
    img[1, 1] = 1
    det = sb.ComputeDoHExtrema()
    feats, _ = det(img)
    ref = O.compute_function_extrema(img, "doh", O.PyramidParams(-1, 5, k3(), 2), 1e-6, 10.0, 1, 2)
    compare(det._ctx, ref, 5)


@pytest.mark.parametrize("which", ["log", "doh"])
@pytest.mark.parametrize("size,fo,ns", [((640, 480), 0, 5), ((517, 389), 0, 6), ((300, 200), -1, 5)])
def test_function_extrema_vs_oracle(which, size, fo, ns):
    import sara_b200 as sb
    from oracle import oracle as O
    from sara_b200 import synthetic as S

    w, h = size
    img = S.tex(w, h, 77)
    thres = 0.01 if which == "log" else 1e-4
    ctx = sb.SiftContext(w, h, device=0)
    pp = sb.ImagePyramidParams(first_octave_index=fo, scale_count_per_octave=ns)
    e = ctx.function_extrema(which, img, pp, thres, 10.0, 2, 5)
    ref = O.compute_function_extrema(img, which, O.PyramidParams(first_octave_index=fo, scale_count_per_octave=ns),
                                     thres, 10.0, 2, 5)
    assert len(e) > 20
    compare(ctx, ref, ns)
    # the DoG detector on the same context afterwards is unaffected
    d = ctx.dog_extrema(img, sb.ImagePyramidParams(first_octave_index=0))
    rd = O.compute_dog_extrema(img, O.PyramidParams(first_octave_index=0))
    assert len(d) == len(rd.extrema) and np.array_equal(d["xi"], rd.extrema["xi"])
    ctx.close()


def compare_hessian_laplace(ctx, ref, n_layers):
    for o in range(ref.num_octaves):
        for s in range(n_layers):
            assert ctx.dog_layer(s, o).tobytes() == ref.dog(s, o).tobytes(), (s, o)
    e, r = ctx.extrema(), ref.extrema
    assert len(e) == len(r)
    for k in ("xi", "yi", "s", "o", "extremum_type"):
        assert np.array_equal(e[k], r[k]), k
    for k in ("x", "y", "extremum_value"):  # 2-D refinement: bit-identical
        assert e[k].tobytes() == r[k].tobytes(), k
    # the selected scale goes through pow(ratio, h) (libm vs CUDA): shape = scale^-2 within 1e-5 relative
    assert np.allclose(e["shape"], r["shape"], rtol=1e-5, atol=0)


def test_reference_hessian_laplace_smoke_kat():
    """test_featuredetectors_hessian.cpp:23-33: default ComputeHessianLaplaceMaxima on a single bright pixel."""
    import sara_b200 as sb
    from oracle import oracle as O

    img = np.zeros((21, 21), np.float32)
    img[1, 1] = 1
    det = sb.ComputeHessianLaplaceMaxima()
    feats, _ = det(img)
    ref = O.compute_hessian_laplace(img)
    compare_hessian_laplace(det._ctx, ref, 4)


@pytest.mark.parametrize("size,fo,ns,num_scales,thres", [((640, 480), 0, 4, 10, 1e-5), ((517, 389), 0, 5, 6, 1e-4),
                                                         ((300, 200), -1, 4, 10, 1e-5)])
def test_hessian_laplace_vs_oracle(size, fo, ns, num_scales, thres):
    import sara_b200 as sb
    from oracle import oracle as O
    from sara_b200 import synthetic as S

    w, h = size
    img = S.tex(w, h, 78)
    ctx = sb.SiftContext(w, h, device=0)
    pp = sb.ImagePyramidParams(first_octave_index=fo, scale_count_per_octave=ns)
    e = ctx.hessian_laplace(img, pp, thres, 2, num_scales, 5)
    ref = O.compute_hessian_laplace(img, O.PyramidParams(first_octave_index=fo, scale_count_per_octave=ns), thres, 2,
                                    num_scales, 5)
    assert len(e) > 50
    compare_hessian_laplace(ctx, ref, ns)
    # SIFT on the same context afterwards is unaffected
    kl = ctx.compute_sift_keypoints(img, sb.ImagePyramidParams(first_octave_index=0))
    rk = O.compute_sift_keypoints(img, O.PyramidParams(first_octave_index=0), parallel=True)
    assert len(kl) == len(rk.keypoints)
    ctx.close()


def test_reference_harris_laplace_smoke_kat():
    """test_featuredetectors_harris.cpp:35-45: default ComputeHarrisLaplaceCorners on a single bright pixel."""
    import sara_b200 as sb
    from oracle import oracle as O

    img = np.zeros((21, 21), np.float32)
    img[1, 1] = 1
    det = sb.ComputeHarrisLaplaceCorners()
    det(img)
    compare_hessian_laplace(det._ctx, O.compute_harris_laplace(img), 3)


@pytest.mark.parametrize("size,fo,ns,k,thres", [((640, 480), 0, 3, None, 1e-9), ((517, 389), 0, 4, 2.0 ** (1 / 3), 1e-8),
                                                ((300, 200), -1, 3, None, 1e-9)])
def test_harris_laplace_vs_oracle(size, fo, ns, k, thres):
    import sara_b200 as sb
    from oracle import oracle as O
    from sara_b200 import synthetic as S

    w, h = size
    img = S.tex(w, h, 79)
    kk = float(np.sqrt(np.float32(2.0))) if k is None else float(np.float32(k))
    ctx = sb.SiftContext(w, h, device=0)
    e = ctx.harris_laplace(img, sb.ImagePyramidParams(fo, ns, kk, 1), 0.04, thres, 2, 10, 5)
    ref = O.compute_harris_laplace(img, O.PyramidParams(fo, ns, kk, 1), 0.04, thres, 2, 10, 5)
    assert len(e) > 50
    compare_hessian_laplace(ctx, ref, ns)
    ctx.close()
