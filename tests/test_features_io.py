"""Keypoint text format and de-duplication: Features/IO.hpp:78-134,
Features/Utilities.cpp:23-82.  CPU only: the Python mirror and the C++ header (compiled with
g++ against the header-only adapter) must write the same bytes and read each other's files."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import sara_b200 as sb
from sara_b200.api import KEYPOINT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _keys(n=7, seed=0):
    rng = np.random.default_rng(seed)
    f = np.zeros(n, KEYPOINT_DTYPE)
    f["x"], f["y"] = rng.uniform(0, 4000, n), rng.uniform(0, 2000, n)
    a = (1.0 / rng.uniform(1.5, 30, n) ** 2).astype(np.float32)
    f["shape"] = np.stack([a, np.zeros(n), np.zeros(n), a], axis=1)
    f["orientation"] = rng.uniform(-np.pi, np.pi, n)
    f["extremum_value"] = rng.normal(0, 0.05, n)
    f["type"] = 11
    f["extremum_type"] = rng.choice([-1, 1], n)
    d = rng.uniform(0, 255, (n, 128)).astype(np.float32)
    d[rng.random((n, 128)) < 0.3] = 0
    return sb.KeypointList(f, d)


def test_text_format_matches_the_reference_layout(tmp_path):
    k = _keys(1)
    k.features["x"], k.features["y"] = 12.5, 1e-7
    k.features["shape"][0] = [0.25, 0, 0, 0.0123456789]
    k.features["orientation"][0] = -1.5
    k.descriptors[0, :] = 0
    k.descriptors[0, :3] = [1, 22.5, 255]
    p = str(tmp_path / "k.txt")
    assert sb.write_keypoints(k, p)
    lines = open(p).read().split("\n")
    assert lines[0] == "1 128"
    # ostream default precision (6 significant digits); Eigen pads a vector's coefficients to one width
    assert lines[1].startswith("12.5 1e-07      0.25         0         0 0.0123457 -1.5 11 ")
    assert lines[1].split()[8:11] == ["1", "22.5", "255"]
    assert lines[1].endswith("    0")  # "0" right-aligned to the width of "22.5"


def test_round_trip_and_redundant_features(tmp_path):
    k = _keys(40, 3)
    # duplicates: same descriptor, different extremum values; near-duplicate within 1e-6
    k.descriptors[5] = k.descriptors[17]
    k.features["extremum_value"][5], k.features["extremum_value"][17] = 0.3, 0.1
    k.descriptors[8] = k.descriptors[2] + np.float32(1e-5)
    p = str(tmp_path / "k.txt")
    sb.write_keypoints(k, p)
    r = sb.read_keypoints(p)
    assert len(r) == 40 and r.descriptors.shape == (40, 128)
    assert np.allclose(r.features["x"], k.features["x"], rtol=1e-5) and np.allclose(r.descriptors, k.descriptors, rtol=1e-5)
    assert np.allclose(r.features["shape"], k.features["shape"], rtol=1e-5)
    u = sb.remove_redundant_features(k)
    assert len(u) == 38
    # sorted lexicographically by descriptor, and of the exact duplicates the larger extremum value survives
    d = u.descriptors
    for a, b in zip(d[:-1], d[1:]):
        ne = np.nonzero(a != b)[0]
        assert len(ne) and a[ne[0]] < b[ne[0]]
    kept = u.features[[bool((row == k.descriptors[17]).all()) for row in d]]
    assert len(kept) == 1 and np.float32(kept["extremum_value"][0]) == np.float32(0.3)
    assert sb.read_keypoints(str(tmp_path / "missing.txt")) is None


def test_cpp_header_agrees_with_python(tmp_path):
    src = tmp_path / "io.cpp"
    src.write_text(textwrap.dedent("""
        #include "sara_b200_io.hpp"
        using namespace sara_b200;
        int main(int argc, char** argv)
        {
          std::vector<OERegion> f; DescriptorMatrix d;
          if (!read_keypoints(f, d, argv[1])) return 1;
          if (!write_keypoints(f, d, argv[2])) return 2;
          remove_redundant_features(f, d);
          if (!write_keypoints(f, d, argv[3])) return 3;
          return read_keypoints(f, d, "/nonexistent/file") ? 4 : 0;
        }"""))
    exe = str(tmp_path / "io")
    # the adapter header declares the C ABI; nothing of it is called here, so no library is linked
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe])
    k = _keys(25, 9)
    k.descriptors[3] = k.descriptors[11]
    a, b, c = (str(tmp_path / n) for n in ("a.txt", "b.txt", "c.txt"))
    sb.write_keypoints(k, a)
    subprocess.check_call([exe, a, b, c], stderr=subprocess.DEVNULL)
    ra, rb = sb.read_keypoints(a), sb.read_keypoints(b)
    assert ra.features.tobytes() == rb.features.tobytes() and ra.descriptors.tobytes() == rb.descriptors.tobytes()
    # C++ remove_redundant_features == Python's on the values the file carries
    want = sb.remove_redundant_features(ra)
    got = sb.read_keypoints(c)
    assert len(got) == len(want) == 24
    assert np.array_equal(got.descriptors, want.descriptors)
    # and the Python writer produces the bytes the C++ writer produces
    d = str(tmp_path / "d.txt")
    sb.write_keypoints(ra, d)
    assert open(d).read() == open(b).read()
