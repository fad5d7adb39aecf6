"""Frame ingest: 8-bit frames -> float32 gray, the conversion the reference's
video loop runs before compute_sift_keypoints (video_sift_matching.cpp:184-200,
ImageProcessing/FastColorConversion.cpp:42-67, non-Halide branch).  There are only 2^24 RGB
triplets, so parity is checked EXHAUSTIVELY: oracle vs an independent numpy float64
restatement (CPU), and the CUDA kernel vs the oracle (GPU), bit for bit."""
import numpy as np
import pytest

from oracle import oracle as O


def _all_rgb():
    v = np.arange(256, dtype=np.uint8)
    rgb = np.empty((256, 256, 256, 3), np.uint8)
    rgb[..., 0] = v[:, None, None]
    rgb[..., 1] = v[None, :, None]
    rgb[..., 2] = v[None, None, :]
    return rgb.reshape(4096, 4096, 3)


def test_oracle_rgb8_to_gray32f_exhaustive():
    rgb = _all_rgb()
    got = O.rgb8_to_gray32f(rgb)
    f = rgb.astype(np.float64) / 255.0  # to_normalized_float_channel<uint8_t, double>
    want = ((0.2125 * f[..., 0] + 0.7154 * f[..., 1]) + 0.0721 * f[..., 2]).astype(np.float32)
    assert got.tobytes() == want.tobytes()
    assert got.min() == 0.0 and abs(float(got.max()) - 1.0) < 1e-6


def test_oracle_gray8_to_gray32f_exhaustive():
    g = np.arange(256, dtype=np.uint8).reshape(16, 16)
    got = O.gray8_to_gray32f(g)
    assert got.tobytes() == (g.astype(np.float32) / np.float32(255)).tobytes()
    assert got[0, 0] == 0.0 and got[-1, -1] == 1.0


@pytest.mark.gpu
def test_gpu_conversion_exhaustive_and_ragged():
    import sara_b200 as sb

    ctx = sb.SiftContext(4096, 4096, max_keypoints=1024)
    try:
        rgb = _all_rgb()
        assert ctx.to_gray32f(rgb).tobytes() == O.rgb8_to_gray32f(rgb).tobytes()
        rng = np.random.default_rng(9)
        for h, w in [(1, 1), (3, 5), (37, 23), (375, 500), (1080, 1921)]:  # pixel counts not multiples of 4
            a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            assert ctx.to_gray32f(a).tobytes() == O.rgb8_to_gray32f(a).tobytes()
            g = rng.integers(0, 256, (h, w), dtype=np.uint8)
            assert ctx.to_gray32f(g).tobytes() == O.gray8_to_gray32f(g).tobytes()
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [3, 1])
def test_sift_from_8bit_frames(channels):
    """sara_b200_sift_u8 == from_rgb8_to_gray32f + compute_sift_keypoints of the oracle, and
    bit-identical to the float entry point fed with the converted image."""
    import torch

    import sara_b200 as sb
    from sara_b200 import synthetic as S
    from parity import assert_extrema_identical, assert_pyramids_identical, compare_keypoints

    base = S.tex(800, 600, 77)
    if channels == 3:
        rng = np.random.default_rng(1)
        tint = np.stack([base * 0.9, base, np.clip(base * 1.1, 0, 1)], axis=2) + rng.normal(0, 0.01, (600, 800, 3))
        u8 = np.clip(np.rint(tint * 255), 0, 255).astype(np.uint8)
        gray = O.rgb8_to_gray32f(u8)
    else:
        u8 = np.clip(np.rint(base * 255), 0, 255).astype(np.uint8)
        gray = O.gray8_to_gray32f(u8)
    pp, po = sb.ImagePyramidParams(first_octave_index=0), O.PyramidParams(first_octave_index=0)
    ctx = sb.SiftContext(800, 600, max_keypoints=65536, num_slots=2)
    try:
        ref = O.compute_sift_keypoints(gray, po, parallel=True)
        kl = ctx.compute_sift_keypoints_u8(u8, pp)
        assert_pyramids_identical(ctx, ref)
        assert_extrema_identical(ctx.extrema(), ref.extrema)
        assert len(ref.keypoints) > 300
        compare_keypoints(kl.features, kl.descriptors, ref.keypoints, ref.descriptors, ref, ctx.oriented())
        kf = ctx.compute_sift_keypoints(gray, pp)
        assert kf.features.tobytes() == kl.features.tobytes() and kf.descriptors.tobytes() == kl.descriptors.tobytes()
        # device-resident 8-bit frame (a decoder surface), asynchronous form, second slot
        ctx.enqueue_u8(1, torch.from_numpy(u8).cuda(), pp)
        kd = ctx.collect(1)
        assert kd.features.tobytes() == kl.features.tobytes() and kd.descriptors.tobytes() == kl.descriptors.tobytes()
        with pytest.raises(ValueError):
            ctx.enqueue_u8(0, np.zeros((4, 4, 2), np.uint8), pp)
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(3840, 2160), (3001, 2003), (1920, 1080)])
def test_pageable_frames_are_staged_without_changing_a_bit(size):
    """Frames in ordinary (pageable) host memory of 4 MiB or more are uploaded by the multi-threaded pinned-chunk
    stager (HostStager in csrc/ctx.cu): same results as the same frame resident on the device, on repeated calls
    (chunk reuse), with and without CUDA graphs, for float and 8-bit frames, for sizes that do not divide into
    the threads' parts or the 2 MiB chunks."""
    import torch

    import sara_b200 as sb
    from sara_b200 import synthetic as S

    w, h = size
    img = S.tex(w, h, 31)
    u8 = np.clip(np.rint(np.stack([img, img[::-1], img[:, ::-1]], axis=2) * 255), 0, 255).astype(np.uint8)
    pp = sb.ImagePyramidParams(first_octave_index=0)
    ctx = sb.SiftContext(w, h, max_keypoints=131072, num_slots=2)
    try:
        want = ctx.compute_sift_keypoints(torch.from_numpy(img).cuda(), pp)
        want8 = None
        ctx.enqueue_u8(1, torch.from_numpy(u8).cuda(), pp)
        want8 = ctx.collect(1)
        assert len(want) > 100 and len(want8) > 100
        for graphs in (True, False):
            ctx.set_graphs(graphs)
            for _ in range(3):
                got = ctx.compute_sift_keypoints(img, pp)
                assert got.features.tobytes() == want.features.tobytes()
                assert got.descriptors.tobytes() == want.descriptors.tobytes()
                got8 = ctx.compute_sift_keypoints_u8(u8, pp)
                assert got8.features.tobytes() == want8.features.tobytes()
                assert got8.descriptors.tobytes() == want8.descriptors.tobytes()
    finally:
        ctx.close()
