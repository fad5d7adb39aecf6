"""A second, independent restatement of the gradient / orientation-histogram / descriptor stage, written in
vectorised float64 numpy straight from the reference sources (not from oracle/sift_oracle.cpp), checked against the
oracle.  The reference's own tests pin this stage only by shape and by a single-gradient histogram
(test_featuredescriptors_sift.cpp, test_featuredescriptors_orientation.cpp), so without this file a slip in the
oracle's bin indexing, weights or window geometry would go unnoticed by every parity test (the CUDA kernels are
compared with the oracle).  float64 against float32: values agree to rounding, except that a sample sitting on the
edge of the rotated grid or on a bin boundary can fall on the other side -- hence the small tolerances below.

  gradient (central differences, one-sided halves at the borders)   ImageProcessing/Differential.hpp:46-61
  polar form (2 |g|, atan2)                                          FeatureDescriptors/Orientation.cpp:45-51
  compute_orientation_histogram                                      FeatureDescriptors/Orientation.hpp:90-135
  ComputeSIFTDescriptor<4, 8>::operator(), accumulate, normalize     FeatureDescriptors/SIFT.hpp:62-145, 204-258
"""
import numpy as np

from oracle import oracle as O
from sara_b200 import synthetic as S


def int_round(x):
    """DO::Sara::int_round: int(round(x)), round half away from zero (Core/Math/UsualFunctions.hpp)."""
    return int(np.floor(abs(x) + 0.5) * np.sign(x))


def gradient_polar64(img):
    f = img.astype(np.float64)
    gx = np.empty_like(f)
    gy = np.empty_like(f)
    gx[:, 1:-1] = (f[:, 2:] - f[:, :-2]) / 2
    gx[:, 0] = (f[:, 1] - f[:, 0]) / 2
    gx[:, -1] = (f[:, -1] - f[:, -2]) / 2
    gy[1:-1, :] = (f[2:, :] - f[:-2, :]) / 2
    gy[0, :] = (f[1, :] - f[0, :]) / 2
    gy[-1, :] = (f[-1, :] - f[-2, :]) / 2
    return 2 * np.hypot(gx, gy), np.arctan2(gy, gx)


def orientation_histogram64(mag, ori, x, y, s, trunc=3.0, blur=1.5, n=36):
    h, w = mag.shape
    rx, ry = int_round(x), int_round(y)
    sigma = s * blur
    rad = int_round(sigma * trunc)
    v, u = np.mgrid[-rad:rad + 1, -rad:rad + 1]
    X, Y = rx + u, ry + v
    ok = (X >= 0) & (X < w) & (Y >= 0) & (Y < h)
    X, Y, u, v = X[ok], Y[ok], u[ok], v[ok]
    o = ori[Y, X]
    o = np.where(o < 0, o + 2 * np.pi, o)
    b = np.floor(o / (2 * np.pi) * n).astype(int) % n
    wgt = np.exp(-(u * u + v * v) / (2 * sigma * sigma)) * mag[Y, X]
    return np.bincount(b, weights=wgt, minlength=n)


def sift_descriptor64(mag, ori, x, y, s, theta, N=4, Ob=8, lam=3.0, max_bin=0.2):
    h, w = mag.shape
    l = lam * s
    r = np.sqrt(2.0) * l * (N + 1) / 2
    T = np.array([[np.cos(theta), np.sin(theta)], [-np.sin(theta), np.cos(theta)]]) / l
    rr, rx, ry = int_round(r), int_round(x), int_round(y)
    v, u = np.mgrid[-rr:rr + 1, -rr:rr + 1]
    X, Y = rx + u, ry + v
    ok = (X >= 0) & (X < w) & (Y >= 0) & (Y < h)
    X, Y, u, v = X[ok], Y[ok], u[ok], v[ok]
    px = T[0, 0] * u + T[0, 1] * v
    py = T[1, 0] * u + T[1, 1] * v
    weight = np.exp(-(px * px + py * py) / (2.0 * (N * N * 0.25)))
    m = mag[Y, X]
    o = ori[Y, X] - theta
    o = np.where(o < 0, o + 2 * np.pi, o) * Ob / (2 * np.pi)
    px = px + (N / 2.0 - 0.5)
    py = py + (N / 2.0 - 0.5)
    keep = (np.minimum(px, py) > -1.0) & (np.maximum(px, py) < N)
    px, py, o, weight, m = px[keep], py[keep], o[keep], weight[keep], m[keep]
    # std::modf: integer part truncated toward zero, fraction with the sign of the argument
    xi, yi, oi = np.trunc(px), np.trunc(py), np.trunc(o)
    xf, yf, of = px - xi, py - yi, o - oi
    xi, yi, oi = xi.astype(int), yi.astype(int), oi.astype(int)
    hist = np.zeros(N * N * Ob)
    for dy in (0, 1):
        yy = yi + dy
        wy = 1 - yf if dy == 0 else yf
        for dx in (0, 1):
            xx = xi + dx
            wx = 1 - xf if dx == 0 else xf
            inside = (yy >= 0) & (yy < N) & (xx >= 0) & (xx < N)
            for do in (0, 1):
                oo = (oi + do) % Ob
                wo = 1 - of if do == 0 else of
                idx = N * Ob * yy + xx * Ob + oo
                np.add.at(hist, idx[inside], (wy * wx * wo * weight * m)[inside])
    nrm = np.linalg.norm(hist)
    hist = hist / nrm
    hist = np.minimum(hist, max_bin)
    hist = hist / np.linalg.norm(hist)
    return np.minimum(hist * 512.0, 255.0)


def _frame():
    return S.tex(320, 240, 4242)


def test_gradient_polar_against_float64():
    img = _frame()
    g = O.gradient_polar(img)
    mag, ori = gradient_polar64(img)
    assert np.allclose(g[..., 0], mag, rtol=2e-5, atol=1e-7)
    # the angle of a (numerically) zero gradient is arbitrary
    strong = mag > 1e-4
    d = np.abs(g[..., 1][strong] - ori[strong])
    d = np.minimum(d, 2 * np.pi - d)
    assert d.max() < 1e-4


def test_orientation_histogram_against_float64():
    img = _frame()
    g = O.gradient_polar(img)
    mag, ori = g[..., 0].astype(np.float64), g[..., 1].astype(np.float64)  # same inputs for both
    rng = np.random.default_rng(5)
    for _ in range(40):
        x, y = rng.uniform(-3, 323), rng.uniform(-3, 243)  # windows that leave the image included
        s = rng.uniform(0.8, 6.0)
        got = O.orientation_histogram(g, x, y, s)
        want = orientation_histogram64(mag, ori, x, y, s)
        assert np.allclose(got, want, rtol=1e-4, atol=1e-5 * max(want.max(), 1e-6)), (x, y, s)


def test_sift_descriptor_against_float64():
    img = _frame()
    g = O.gradient_polar(img)
    mag, ori = g[..., 0].astype(np.float64), g[..., 1].astype(np.float64)
    rng = np.random.default_rng(6)
    worst = []
    for i in range(200):
        x, y = rng.uniform(-2, 322), rng.uniform(-2, 242)
        s = rng.uniform(0.6, 5.0)  # includes windows narrower than a warp and windows larger than the frame
        # ComputeDominantOrientations hands out angles in [-pi, pi) (Orientation.cpp:110-114); outside that range
        # "ori - theta + 2 pi" can stay negative and the reference's "% O" then indexes a neighbouring bin.
        theta = rng.uniform(-np.pi, np.pi) if i % 5 else 0.0  # the upright form too
        got = O.sift_descriptor(g, x, y, s, theta).astype(np.float64)
        want = sift_descriptor64(mag, ori, x, y, s, theta)
        assert got.shape == want.shape == (128,)
        worst.append(np.linalg.norm(got - want) / (1e-3 * np.linalg.norm(want) + 0.05))
    worst = np.array(worst)
    # Unit: the parity tolerance of tests/parity.py (1e-3 |d| + 0.05 on the 0-255 scale).  Measured: every one of
    # the descriptors within 4e-4 of that unit; a boundary sample flipping (float32 against float64) would show
    # as a few tenths, a wrong bin index or weight as tens to hundreds.
    assert np.mean(worst <= 0.01) >= 0.97 and worst.max() < 1.0, np.sort(worst)[-5:]


def test_unnormalised_descriptor_against_float64():
    """Without normalize(): the raw trilinear histogram, bin for bin (pins at(y, x, o) = N O y + O x + o)."""
    img = _frame()
    g = O.gradient_polar(img)
    mag, ori = g[..., 0].astype(np.float64), g[..., 1].astype(np.float64)

    x, y, s, theta = 160.3, 120.7, 2.0, 0.7
    got = O.sift_descriptor(g, x, y, s, theta, normalize=False).astype(np.float64)
    # the normalised float64 descriptor must be the normalisation of the oracle's raw histogram
    h = got / np.linalg.norm(got)
    h = np.minimum(h, 0.2)
    h = np.minimum(h / np.linalg.norm(h) * 512.0, 255.0)
    want = sift_descriptor64(mag, ori, x, y, s, theta)
    assert np.linalg.norm(h - want) <= 1e-3 * np.linalg.norm(want) + 0.05
    assert np.argmax(got) == np.argmax(want) or want[np.argmax(got)] >= 0.99 * want.max()


# ---- the Gaussian / DoG pyramid ------------------------------------------------------------------------------------
#   make_gaussian_kernel, apply_row/column_based_filter     ImageProcessing/LinearFiltering.hpp:78-149, 172-203
#   gaussian_pyramid (first octave 0), downscale            ImageProcessing/GaussianPyramid.hpp:36-123, Resize.cpp:31-83
#   difference_of_gaussians_pyramid                         ImageProcessing/GaussianPyramid.cpp:23-51
def _kernel64(sigma32, truncate=4.0):
    size = int(np.float32(2) * np.float32(truncate) * sigma32 + np.float32(1))  # the reference sizes it in float
    size = max(3, size)
    size += 1 - size % 2
    c = size // 2
    k = np.exp(-(np.arange(size, dtype=np.float64) - c) ** 2 / (2.0 * float(sigma32) ** 2))
    return k / k.sum()


def _gaussian64(img, sigma32):
    from scipy.ndimage import correlate1d

    k = _kernel64(sigma32)
    rows = correlate1d(img, k, axis=1, mode="nearest")  # x first, borders replicated
    return correlate1d(rows, k, axis=0, mode="nearest")


def _enlarge64(img, fact):
    """enlarge + interpolate (Resize.cpp:86-128, Interpolation.hpp:34-78): bilinear at dst * (src / dst), the
    neighbour past the last row / column replaced by the last one; the result is stored as float."""
    h, w = img.shape
    dh, dw = int(h * fact), int(w * fact)
    py, px = np.arange(dh) * (h / dh), np.arange(dw) * (w / dw)
    y0, x0 = np.floor(py).astype(int), np.floor(px).astype(int)
    fy, fx = (py - y0)[:, None], (px - x0)[None, :]
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
    f = img.astype(np.float64)
    out = (1 - fy) * (1 - fx) * f[np.ix_(y0, x0)] + (1 - fy) * fx * f[np.ix_(y0, x1)] + \
        fy * (1 - fx) * f[np.ix_(y1, x0)] + fy * fx * f[np.ix_(y1, x1)]
    return out.astype(np.float32).astype(np.float64)


def _pyramid64(img, first_octave=0, scale_camera=0.5, scale_initial=1.6, scales_per_octave=3 + 3, pad=1,
               max_octaves=100):
    f32 = np.float32
    I = img.astype(np.float64)
    resize = f32(2.0) ** f32(-first_octave)
    cam, init = f32(scale_camera) * resize, f32(scale_initial)
    if first_octave < 0:
        I = _enlarge64(img, float(resize))  # and NO blur up to scale_initial: GaussianPyramid.hpp:55-56
    elif cam < init:
        assert first_octave == 0
        I = _gaussian64(I, np.sqrt(init * init - cam * cam, dtype=f32))
    h, w = I.shape
    n_oct = min(int(np.log(f32(min(w, h)) / (f32(2) * f32(pad))) / np.log(f32(2))), max_octaves)
    k = np.power(f32(2), f32(1) / f32(3), dtype=f32)  # std::pow(2.f, 1.f / 3.f)
    down = int(np.floor(np.log(f32(2)) / np.log(k)))
    G = []
    for o in range(n_oct):
        if o > 0:
            src = G[o - 1][down]
            sh, sw = src.shape
            dh, dw = sh // 2, sw // 2
            ys = (np.arange(dh) * (f32(sh) / f32(dh))).astype(int)  # nearest sample at float ratio, truncated
            xs = (np.arange(dw) * (f32(sw) / f32(dw))).astype(int)
            I = src[np.ix_(ys, xs)]
        layers, sig = [I], init
        for s in range(1, scales_per_octave):
            ks = k * sig
            layers.append(_gaussian64(layers[-1], np.sqrt(ks * ks - sig * sig, dtype=f32)))
            sig = f32(sig * k)
        G.append(layers)
    return G


import pytest  # noqa: E402


@pytest.mark.parametrize("first_octave", [0, -1])
def test_gaussian_and_dog_pyramid_against_float64(first_octave):
    img = S.tex(322, 241, 99)  # odd sizes: the even sub-sampling of every octave drops a row / column
    ref = O.compute_sift_keypoints(img, O.PyramidParams(first_octave_index=first_octave), parallel=True)
    G = _pyramid64(img, first_octave)
    assert ref.octave_scaling(0) == 2.0 ** first_octave
    assert ref.num_octaves == len(G) and ref.num_scales == len(G[0]) == 6
    for o in range(len(G)):
        for s in range(6):
            got = ref.gaussian(s, o)
            assert got.shape == G[o][s].shape, (s, o)
            assert np.abs(got - G[o][s]).max() < 2e-6, (s, o)  # fp32 sums of <= 25 products of values in [0, 1]
            if s < 5:
                assert np.abs(ref.dog(s, o) - (G[o][s + 1] - G[o][s])).max() < 3e-6, (s, o)


# ---- scale-space extrema --------------------------------------------------------------------------------------------
#   LocalScaleSpaceExtremum (non-strict 3x3x3)      ImageProcessing/Extrema.hpp:28-75
#   on_edge, refine_extremum                         FeatureDetectors/RefineExtremum.cpp:24-130
#   local_scale_space_extrema                        FeatureDetectors/RefineExtremum.cpp:363-521
#   gradient / hessian of the pyramid                ImageProcessing/GaussianPyramid.hpp:184-244
# The decisions are re-taken in float64 on the ORACLE's own DoG layers (so ties stay ties); the linear algebra is
# numpy's (eigvalsh, solve) instead of the oracle's restated Eigen routines.
def _extrema64(D, scale_initial, k, thres, edge_ratio, pad, n_iter):
    """D: list of DoG layers of one octave (float64).  Returns [(xi, yi, s, x, y, scale, value, type)] in the
    reference's order (scale-major, raster)."""
    out = []
    n = len(D)
    h, w = D[0].shape
    rel = lambda s: float(np.float32(np.float32(k) ** s) * np.float32(scale_initial))

    def grad3(x, y, s):
        return np.array([(D[s][y, x + 1] - D[s][y, x - 1]) / 2, (D[s][y + 1, x] - D[s][y - 1, x]) / 2,
                         (D[s + 1][y, x] - D[s - 1][y, x]) / 2])

    def hess3(x, y, s):
        c = D[s][y, x]
        H = np.empty((3, 3))
        H[0, 0] = D[s][y, x + 1] - 2 * c + D[s][y, x - 1]
        H[1, 1] = D[s][y + 1, x] - 2 * c + D[s][y - 1, x]
        H[2, 2] = D[s + 1][y, x] - 2 * c + D[s - 1][y, x]
        H[0, 1] = H[1, 0] = (D[s][y + 1, x + 1] - D[s][y + 1, x - 1] - D[s][y - 1, x + 1] + D[s][y - 1, x - 1]) / 4
        H[0, 2] = H[2, 0] = (D[s + 1][y, x + 1] - D[s + 1][y, x - 1] - D[s - 1][y, x + 1] + D[s - 1][y, x - 1]) / 4
        H[1, 2] = H[2, 1] = (D[s + 1][y + 1, x] - D[s + 1][y - 1, x] - D[s - 1][y + 1, x] + D[s - 1][y - 1, x]) / 4
        return H

    from scipy.ndimage import maximum_filter, minimum_filter

    for s in range(1, n - 1):
        cube = np.stack([D[s - 1], D[s], D[s + 1]])
        is_max = D[s] >= maximum_filter(cube, size=3, mode="nearest")[1]
        is_min = D[s] <= minimum_filter(cube, size=3, mode="nearest")[1]
        for y in range(pad, h - pad):
            for x in np.nonzero((is_max | is_min)[y, pad:w - pad])[0] + pad:
                x = int(x)
                v = D[s][y, x]
                typ = 1 if is_max[y, x] else -1  # the maximum is tested first
                if abs(v) < float(np.float32(0.8) * np.float32(thres)):
                    continue
                hxx = D[s][y, x + 1] - 2 * v + D[s][y, x - 1]
                hyy = D[s][y + 1, x] - 2 * v + D[s][y - 1, x]
                hxy = (D[s][y + 1, x + 1] - D[s][y + 1, x - 1] - D[s][y - 1, x + 1] + D[s][y - 1, x - 1]) / 4
                if (hxx + hyy) ** 2 * edge_ratio >= (edge_ratio + 1) ** 2 * abs(hxx * hyy - hxy * hxy):
                    continue
                # refine_extremum; the map is an Image<uint8_t>: a minimum is stored as 255 and comes back as the
                # "type" 255, for which the Hessian test never passes and neither value update applies
                t = 1 if typ == 1 else 255
                cx, cy = x, y
                hvec = np.zeros(3)
                g = np.zeros(3)
                ok = True
                for _ in range(n_iter):
                    if cx < pad or cx >= w - pad or cy < pad or cy >= h - pad:
                        break
                    g = grad3(cx, cy, s)
                    H = hess3(cx, cy, s)
                    if (np.linalg.eigvalsh(H) * t).max() >= 0:
                        hvec = np.zeros(3)
                        break
                    hvec = -np.linalg.solve(H, g)
                    if np.abs(hvec[:2]).max() > 1.5:
                        ok = False  # "return false": the caller ignores it, pos and val keep their initial values
                        break
                    if np.abs(hvec[:2]).min() > 0.6:
                        cx += 1 if hvec[0] > 0 else -1
                        cy += 1 if hvec[1] > 0 else -1
                        continue
                    break
                px, py, sc, val = float(x), float(y), rel(s), v
                if ok:
                    px, py, sc = float(cx), float(cy), rel(s)
                    old = D[s][cy, cx]
                    new = old + 0.5 * g.dot(hvec)
                    if t == 1 and old <= new:
                        px, py = px + hvec[0], py + hvec[1]
                        sc *= float(np.float32(k)) ** hvec[2]
                        val = new
                if abs(val) < thres:
                    continue
                out.append((x, y, s, px, py, sc, val, typ))
    return out


def test_scale_space_extrema_against_float64():
    img = S.tex(640, 480, 7)
    pp = O.PyramidParams(first_octave_index=0)
    for pad, iters in ((1, 5), (5, 5)):
        ref = O.compute_dog_extrema(img, pp, 4.0, 0.01, 10.0, pad, iters)
        got = ref.extrema
        want = []
        for o in range(ref.num_octaves):
            D = [ref.dog(s, o).astype(np.float64) for s in range(ref.num_scales - 1)]
            want += [(o,) + e for e in _extrema64(D, 1.6, 2.0 ** (1.0 / 3.0), float(np.float32(0.01)), 10.0, pad, iters)]
        assert len(got) > 300
        key_got = [(int(e["o"]), int(e["xi"]), int(e["yi"]), int(e["s"])) for e in got]
        key_want = [(e[0], e[1], e[2], e[3]) for e in want]
        # Measured: identical sets (355 and 337 extrema, half of them minima).  float64 against float32 linear
        # algebra could decide a borderline candidate differently, hence the half per cent of slack.
        common = set(key_got) & set(key_want)
        assert len(common) >= 0.995 * max(len(key_got), len(key_want)), (len(key_got), len(key_want), len(common))
        # same order (octave, scale, raster) on the common ones
        assert [k for k in key_got if k in common] == [k for k in key_want if k in common]
        idx = {k: i for i, k in enumerate(key_want)}
        n_min = 0
        for e, k in zip(got, key_got):
            if k not in common:
                continue
            w_ = want[idx[k]]
            assert (e["extremum_type"] == 1) == (w_[8] == 1)
            n_min += e["extremum_type"] != 1
            # measured: 3e-5 px, 3e-9, 2e-7 relative
            assert abs(e["x"] - w_[4]) < 5e-4 and abs(e["y"] - w_[5]) < 5e-4, (k, e["x"], e["y"], w_)
            assert abs(e["extremum_value"] - w_[7]) < 1e-6, (k, e["extremum_value"], w_[7])
            scale = 1.0 / np.sqrt(e["shape"][0])  # OERegion(pos, scale): shape = I / scale^2
            assert abs(scale - w_[6]) < 1e-5 * w_[6], (k, scale, w_[6])
            if e["extremum_type"] != 1:  # quirk N2: minima are never refined
                assert e["x"] == e["xi"] and e["y"] == e["yi"]
        assert n_min > 100


# ---- dominant orientations, end to end ------------------------------------------------------------------------------
#   lowe_smooth_histogram, find_peaks, refine_peak     FeatureDescriptors/Orientation.hpp:136-215
#   ComputeDominantOrientations::operator()            FeatureDescriptors/Orientation.cpp:90-118
def dominant_orientations64(mag, ori, x, y, sigma, peak_ratio=0.8, n=36):
    h = orientation_histogram64(mag, ori, x, y, sigma, n=n)
    for _ in range(6):  # the in-place loop of the reference reads old neighbours only: a circular box filter
        h = (np.roll(h, 1) + h + np.roll(h, -1)) / 3.0
    peaks = [i for i in range(n) if h[i] >= peak_ratio * h.max() and h[i] > h[(i - 1) % n] and h[i] > h[(i + 1) % n]]
    out = []
    for i in peaks:
        y0, y1, y2 = h[(i - 1) % n], h[i], h[(i + 1) % n]
        p = (i + 0.5 - ((y2 - y0) / 2.0) / (y0 - 2.0 * y1 + y2)) * 2 * np.pi / n
        out.append(p - 2 * np.pi if p > np.pi else p)
    return np.array(out), h


def test_dominant_orientations_against_float64():
    img = _frame()
    g = O.gradient_polar(img)
    mag, ori = g[..., 0].astype(np.float64), g[..., 1].astype(np.float64)
    rng = np.random.default_rng(8)
    n_same, n_multi, diffs = 0, 0, []
    for _ in range(300):
        x, y = rng.uniform(-3, 323), rng.uniform(-3, 243)
        s = rng.uniform(0.8, 6.0)
        got = O.dominant_orientations(g, x, y, s)
        want, h = dominant_orientations64(mag, ori, x, y, s)
        if len(got) != len(want):
            # only a bin within rounding of the 0.8 max threshold or of a neighbour may be decided differently
            srt = np.sort(h)
            assert np.min(np.abs(h - 0.8 * h.max())) < 1e-4 * h.max() or np.min(np.diff(srt)) < 1e-6 * h.max(), (x, y, s)
            continue
        n_same += 1
        n_multi += len(got) > 1
        assert np.all((got >= -np.pi - 1e-6) & (got <= np.pi + 1e-6))
        if len(got):
            d = np.abs(got - want)
            diffs += list(np.minimum(d, 2 * np.pi - d))
    assert n_same >= 297 and n_multi >= 20, (n_same, n_multi)
    diffs = np.array(diffs)
    # measured: 412 angles, 411 within 1e-5 rad, one flat peak (tiny second difference) at 4.6e-4
    assert np.mean(diffs < 1e-4) >= 0.99 and diffs.max() < 2e-3, np.sort(diffs)[-5:]


# ---- the wiring of compute_sift_keypoints ----------------------------------------------------------------------------
#   FeatureDetectors/SIFT.cpp:27-108; ComputeDominantOrientations on the pyramid, Orientation.cpp:120-166;
#   ComputeSIFTDescriptor on the keypoint list, FeatureDescriptors/SIFT.hpp:150-200
@pytest.mark.parametrize("first_octave", [0, -1])
def test_whole_chain_wiring_against_float64(first_octave):
    """Which layer feeds the gradients (the Gaussian G(s, o) of the extremum's own (s, o)), which scale goes where
    (the orientation window uses the layer's nominal scale, the descriptor the REFINED one), one keypoint per
    dominant orientation in extremum order, and the final rescaling to image coordinates."""
    img = S.tex(480, 360, 21) if first_octave == 0 else S.tex(300, 220, 22)
    ref = O.compute_sift_keypoints(img, O.PyramidParams(first_octave_index=first_octave), parallel=True)
    assert ref.octave_scaling(0) == 2.0 ** first_octave  # positions of octave 0 are halved when the frame was doubled
    ext, kps, desc = ref.extrema, ref.keypoints, ref.descriptors
    assert len(ext) > 80 and len(kps) >= len(ext)
    k32 = np.power(np.float32(2), np.float32(1) / np.float32(3), dtype=np.float32)
    polar = {}
    j, n_desc_ok = 0, 0
    for e in ext:
        s, o = int(e["s"]), int(e["o"])
        if (s, o) not in polar:
            polar[(s, o)] = gradient_polar64(ref.gaussian(s, o))
        mag, ori = polar[(s, o)]
        nominal = float(np.float32(k32 ** np.float32(s)) * np.float32(1.6))
        thetas, _ = dominant_orientations64(mag, ori, float(e["x"]), float(e["y"]), nominal)
        refined = 1.0 / np.sqrt(float(e["shape"][0]))
        z = float(ref.octave_scaling(o))
        for th in thetas:
            kp = kps[j]
            assert (int(kp["s"]), int(kp["o"]), int(kp["xi"]), int(kp["yi"])) == (s, o, int(e["xi"]), int(e["yi"])), j
            d = abs(float(kp["orientation"]) - th)
            assert min(d, 2 * np.pi - d) < 2e-3, (j, kp["orientation"], th)
            assert abs(kp["x"] - e["x"] * z) < 1e-3 * z and abs(kp["y"] - e["y"] * z) < 1e-3 * z
            assert abs(kp["shape"][0] - e["shape"][0] / (z * z)) < 1e-5 * e["shape"][0] / (z * z)
            want = sift_descriptor64(mag, ori, float(e["x"]), float(e["y"]), refined, float(kp["orientation"]))
            err = np.linalg.norm(desc[j] - want) / (1e-3 * np.linalg.norm(want) + 0.05)
            n_desc_ok += err <= 1.0
            j += 1
    assert j == len(kps), (j, len(kps))  # same number of orientations for every extremum
    assert n_desc_ok >= 0.995 * len(kps), (n_desc_ok, len(kps))


# ---- function pyramids of the sibling detectors ------------------------------------------------------------------------
#   laplacian_pyramid (scale-normalised)        ImageProcessing/GaussianPyramid.hpp:153-178, Differential.hpp:106-135
#   det_of_hessian_pyramid                      FeatureDetectors/Hessian.hpp:35-57, Differential.hpp:191-226
@pytest.mark.parametrize("which", ["log", "doh"])
def test_function_pyramids_against_float64(which):
    img = S.tex(200, 150, 5)
    res = O.compute_function_extrema(img, which, O.PyramidParams(first_octave_index=0, scale_count_per_octave=5))
    k32 = np.power(np.float32(2), np.float32(1) / np.float32(3), dtype=np.float32)
    assert res.num_scales == 5 and res.num_octaves >= 4
    for o in range(res.num_octaves):
        for s in range(res.num_scales):
            g = np.pad(res.gaussian(s, o).astype(np.float64), 1, mode="edge")  # replicated border
            c = g[1:-1, 1:-1]
            hxx = g[1:-1, 2:] - 2 * c + g[1:-1, :-2]
            hyy = g[2:, 1:-1] - 2 * c + g[:-2, 1:-1]
            hxy = (g[2:, 2:] - g[2:, :-2] - g[:-2, 2:] + g[:-2, :-2]) / 4
            rel = float(np.float32(k32 ** np.float32(s)) * np.float32(1.6))  # scale_relative_to_octave(s)
            want = (hxx + hyy) * rel ** 2 if which == "log" else (hxx * hyy - hxy * hxy) * rel ** 4
            got = res.dog(s, o)
            assert got.shape == want.shape
            # fp32 second differences of values in [0, 1] carry ~1e-7; the scale normalisation multiplies that
            assert np.abs(got - want).max() <= 5e-5 * max(np.abs(want).max(), 1e-3), (which, s, o)


# ---- Hessian-Laplace -------------------------------------------------------------------------------------------------
#   ComputeHessianLaplaceMaxima::operator()      FeatureDetectors/Hessian.cpp:19-57
#   laplace_maxima, select_laplace_scale          FeatureDetectors/RefineExtremum.cpp:523-709
#   refine_extremum (2-D)                         FeatureDetectors/RefineExtremum.cpp:132-221
#   LocalMax (non-strict, 8 neighbours)           ImageProcessing/Extrema.hpp:28-60, 106
def _select_laplace_scale64(Gprev, x, y, rel_s, rel_prev, num_scales):
    f32 = np.float32
    radius = int(np.ceil(np.sqrt(f32(2)) * f32(4)))
    h, w = Gprev.shape
    if x - radius < 0 or x + radius >= w or y - radius < 0 or y + radius >= h:
        return None
    patch = Gprev[y - radius:y + radius + 1, x - radius:x + radius + 1]
    ratio = np.power(f32(2), f32(1) / f32(num_scales), dtype=f32)
    scales = [f32(rel_s) / np.sqrt(f32(2))]
    with np.errstate(invalid="ignore"):
        inc = np.sqrt(scales[0] * scales[0] - f32(rel_prev) * f32(rel_prev), dtype=f32)  # NaN: k < sqrt(2)
    patches = [_gaussian64(patch, inc) if inc > f32(1e-3) else patch]
    for i in range(1, num_scales + 1):
        scales.append(f32(ratio * scales[i - 1]))
        inc = np.sqrt(scales[i] * scales[i] - scales[i - 1] * scales[i - 1], dtype=f32)
        patches.append(_gaussian64(patches[i - 1], inc))
    c = radius
    logs = [(p[c, c + 1] + p[c, c - 1] + p[c + 1, c] + p[c - 1, c] - 4 * p[c, c]) * float(s) ** 2
            for p, s in zip(patches, scales)]
    for i in range(1, num_scales):
        if (logs[i] <= logs[i - 1] and logs[i] <= logs[i + 1]) or (logs[i] >= logs[i - 1] and logs[i] >= logs[i + 1]):
            fp = (logs[i + 1] - logs[i - 1]) / 2
            fs = logs[i - 1] - 2 * logs[i] + logs[i + 1]
            return float(scales[i]) * float(ratio) ** (-fp / fs)
    return None


def _refine2d64(F, x, y, pad, n_iter):
    """type = 1.  Returns (px, py, value)."""
    h, w = F.shape
    x0, y0, v0 = x, y, F[y, x]
    hv, g = np.zeros(2), np.zeros(2)
    for _ in range(n_iter):
        if x < pad or x >= w - pad or y < pad or y >= h - pad:
            break
        g = np.array([(F[y, x + 1] - F[y, x - 1]) / 2, (F[y + 1, x] - F[y - 1, x]) / 2])
        hxx = F[y, x + 1] - 2 * F[y, x] + F[y, x - 1]
        hyy = F[y + 1, x] - 2 * F[y, x] + F[y - 1, x]
        hxy = (F[y + 1, x + 1] - F[y + 1, x - 1] - F[y - 1, x + 1] + F[y - 1, x - 1]) / 4
        if hxx * hyy - hxy * hxy <= 0 or hxx + hyy >= 0:
            g = np.zeros(2)  # the offset of an earlier iteration, if any, stays
            break
        hv = -np.linalg.solve(np.array([[hxx, hxy], [hxy, hyy]]), g)
        if np.abs(hv).max() > 1.5:
            return float(x0), float(y0), v0  # "return false", ignored by the caller: p and val keep their initial values
        if np.abs(hv).min() > 0.6:
            x += 1 if hv[0] > 0 else -1
            y += 1 if hv[1] > 0 else -1
            continue
        break
    old = F[y, x]
    new = old + 0.5 * g.dot(hv)
    if old <= new:
        return x + hv[0], y + hv[1], new
    return float(x), float(y), v0


def test_hessian_laplace_against_float64():
    img = S.tex(320, 240, 13)
    pp = O.PyramidParams(first_octave_index=0, scale_count_per_octave=4)
    pad, num_scales, iters, thres = 1, 10, 5, float(np.float32(1e-5))
    ref = O.compute_hessian_laplace(img, pp, 1e-5, pad, num_scales, iters)
    got = ref.extrema
    k32 = np.power(np.float32(2), np.float32(1) / np.float32(3), dtype=np.float32)
    rel = lambda s: float(np.float32(k32 ** np.float32(s)) * np.float32(1.6))
    want = []
    for o in range(ref.num_octaves):
        G = [ref.gaussian(s, o).astype(np.float64) for s in range(ref.num_scales)]
        for s in range(1, ref.num_scales):
            F = ref.dog(s, o).astype(np.float64)  # the det-of-Hessian layer (pinned above)
            h, w = F.shape
            for y in range(pad, h - pad):
                for x in range(pad, w - pad):
                    v = F[y, x]
                    if v < thres:
                        continue
                    nb = F[y - 1:y + 2, x - 1:x + 2]
                    if (nb > v).any():
                        continue
                    scale = _select_laplace_scale64(G[s - 1], x, y, rel(s), rel(s - 1), num_scales)
                    if scale is None:
                        continue
                    px, py, val = _refine2d64(F, x, y, pad, iters)
                    want.append((o, x, y, s, px, py, scale, val))
    assert len(got) > 100, len(got)
    key_got = [(int(e["o"]), int(e["xi"]), int(e["yi"]), int(e["s"])) for e in got]
    key_want = [w_[:4] for w_ in want]
    common = set(key_got) & set(key_want)
    assert len(common) >= 0.99 * max(len(key_got), len(key_want)), (len(key_got), len(key_want), len(common))
    assert [k for k in key_got if k in common] == [k for k in key_want if k in common]
    idx = {k: i for i, k in enumerate(key_want)}
    n_ok = 0
    for e, k in zip(got, key_got):
        if k not in common:
            continue
        w_ = want[idx[k]]
        scale = 1.0 / np.sqrt(float(e["shape"][0]))
        n_ok += abs(e["x"] - w_[4]) < 1e-3 and abs(e["y"] - w_[5]) < 1e-3 and abs(scale - w_[6]) < 1e-3 * w_[6] and \
            abs(e["extremum_value"] - w_[7]) <= 1e-4 * abs(w_[7]) + 1e-9
    assert n_ok >= 0.98 * len(common), (n_ok, len(common))


# ---- Harris-Laplace: the cornerness layers (laplace_maxima itself is pinned by the Hessian-Laplace test) ---------------
#   ComputeHarrisLaplaceCorners::operator()      FeatureDetectors/Harris.cpp:165-230
def test_harris_cornerness_against_float64():
    img = S.tex(200, 150, 17)
    k = float(np.sqrt(np.float32(2.0)))
    ref = O.compute_harris_laplace(img, O.PyramidParams(0, 3, k, 1), kappa=0.04)
    assert ref.num_scales == 3 and ref.num_octaves >= 4
    kf = np.float32(k)
    for o in range(ref.num_octaves):
        for s in range(ref.num_scales):
            g = ref.gaussian(s, o).astype(np.float64)
            gx = np.empty_like(g)
            gy = np.empty_like(g)
            gx[:, 1:-1] = (g[:, 2:] - g[:, :-2]) / 2
            gx[:, 0] = (g[:, 1] - g[:, 0]) / 2
            gx[:, -1] = (g[:, -1] - g[:, -2]) / 2
            gy[1:-1, :] = (g[2:, :] - g[:-2, :]) / 2
            gy[0, :] = (g[1, :] - g[0, :]) / 2
            gy[-1, :] = (g[-1, :] - g[-2, :]) / 2
            sigma_i = np.float32(kf ** np.float32(s)) * np.float32(1.6)
            sigma_d = float(sigma_i * (np.float32(1) / np.sqrt(np.float32(2))))
            a, b, c = (_gaussian64(m, sigma_i) for m in (gx * gx, gx * gy, gy * gy))
            want = ((a * c - b * b) - 0.04 * (a + c) ** 2) * sigma_d ** 2
            got = ref.dog(s, o)
            assert got.shape == want.shape
            # products of small gradients: the layer is of the order 1e-6 .. 1e-4
            assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max() + 1e-12, (s, o, np.abs(want).max())
