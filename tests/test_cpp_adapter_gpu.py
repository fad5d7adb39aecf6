"""The C++ drop-in (include/sara_b200.hpp) end to end on the GPU: a small C++ program calls
sara_b200::compute_sift_keypoints with the reference's signature and defaults and must return
exactly what the Python binding of the same C ABI returns."""
import os
import subprocess

import numpy as np
import pytest

import sara_b200 as sb
from sara_b200 import synthetic as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r"""
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sara_b200.hpp"

int main(int argc, char** argv)
{
  const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
  std::vector<float> img(static_cast<std::size_t>(w) * h);
  FILE* f = std::fopen(argv[1], "rb");
  if (!f || std::fread(img.data(), sizeof(float), img.size(), f) != img.size())
    return 2;
  std::fclose(f);
  namespace sara = sara_b200;
  // same call as cpp/examples/Sara/FeatureDescriptors/sift_example.cpp:59-60, first octave 0
  const auto keys = sara::compute_sift_keypoints(sara::ImageView<float>{img.data(), w, h},
                                                 sara::ImagePyramidParams(0));
  const auto& feats = sara::features<sara::OERegion, float>(keys);
  const auto& desc = sara::descriptors<sara::OERegion, float>(keys);
  f = std::fopen(argv[4], "wb");
  const int n = static_cast<int>(feats.size());
  std::fwrite(&n, sizeof(int), 1, f);
  for (const auto& k : feats)
  {
    const float rec[6] = {k.x(), k.y(), k.shape_matrix(0, 0), k.orientation, k.extremum_value,
                          static_cast<float>(static_cast<int>(k.extremum_type))};
    std::fwrite(rec, sizeof(float), 6, f);
  }
  std::fwrite(desc.data(), sizeof(float), static_cast<std::size_t>(desc.rows()) * desc.cols(), f);
  std::fclose(f);
  // error mapping: fewer than 4 scales must throw std::runtime_error like DoG.hpp:86-89
  try
  {
    sara::compute_sift_keypoints(sara::ImageView<float>{img.data(), w, h}, sara::ImagePyramidParams(0, 3));
    return 3;
  }
  catch (const std::runtime_error&)
  {
  }
  return desc.cols() == 128 || n == 0 ? 0 : 4;
}
"""


def test_cpp_compute_sift_keypoints_matches_python(tmp_path):
    w, h = 640, 480
    img = S.tex(w, h, 123)
    raw = tmp_path / "img.f32"
    img.tofile(raw)
    src = tmp_path / "main.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path / "main"
    lib = sb.library_path()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), lib,
                           f"-Wl,-rpath,{os.path.dirname(lib)}"])
    out = tmp_path / "out.bin"
    assert subprocess.call([str(exe), str(raw), str(w), str(h), str(out)]) == 0
    blob = np.fromfile(out, dtype=np.uint8)
    n = int(blob[:4].view(np.int32)[0])
    rec = blob[4:4 + 24 * n].view(np.float32).reshape(n, 6)
    desc = blob[4 + 24 * n:].view(np.float32).reshape(n, 128)
    kl = sb.compute_sift_keypoints(img, sb.ImagePyramidParams(first_octave_index=0))
    f = sb.features(kl)
    assert n == len(f) and n > 100
    assert np.array_equal(rec[:, 0], f["x"]) and np.array_equal(rec[:, 1], f["y"])
    assert np.array_equal(rec[:, 2], f["shape"][:, 0]) and np.array_equal(rec[:, 3], f["orientation"])
    assert np.array_equal(rec[:, 4], f["extremum_value"]) and np.array_equal(rec[:, 5], f["extremum_type"].astype(np.float32))
    assert desc.tobytes() == sb.descriptors(kl).tobytes()


MATCH_PROGRAM = r"""
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sara_b200.hpp"

namespace sara = sara_b200;
using Keys = sara::KeypointList<sara::OERegion, float>;

static Keys make_keys(int n, int dim)
{
  return Keys{std::vector<sara::OERegion>(n), sara::DescriptorMatrix{n, dim}};
}

int main(int argc, char** argv)
{
  // cpp/test/Sara/FeatureMatching/test_featurematching_matching.cpp:27-57
  {
    auto keys1 = make_keys(1, 2);
    auto keys2 = make_keys(10, 2);
    for (int i = 0; i < 10; ++i)
    {
      std::get<0>(keys2)[i].coords(0) = std::get<0>(keys2)[i].coords(1) = float(i);
      std::get<1>(keys2).data()[2 * i] = std::get<1>(keys2).data()[2 * i + 1] = float(i);
    }
    constexpr auto nearest_neighbor_ratio = 0.6f;
    sara::AnnMatcher matcher{keys1, keys2, nearest_neighbor_ratio};
    auto matches = matcher.compute_matches();
    if (matches.size() != 1u)
      return 10;
    const auto& m = matches.front();
    if (&m.x() != &std::get<0>(keys1)[0] || &m.y() != &std::get<0>(keys2)[0] || m.score() != 0.f)
      return 11;
  }
  // empty key list: "the list of key-points is empty" (AnnMatcher.cpp:45-46) -> std::runtime_error
  try
  {
    auto a = make_keys(3, 128), b = make_keys(0, 128);
    sara::AnnMatcher{a, b, 0.6f}.compute_matches();
    return 12;
  }
  catch (const std::runtime_error&)
  {
  }
  // two descriptor sets from files: n1, n2, then the rows
  FILE* f = std::fopen(argv[1], "rb");
  int n[2];
  if (!f || std::fread(n, sizeof(int), 2, f) != 2)
    return 2;
  auto k1 = make_keys(n[0], 128), k2 = make_keys(n[1], 128);
  if (std::fread(std::get<1>(k1).data(), sizeof(float), size_t(n[0]) * 128, f) != size_t(n[0]) * 128 ||
      std::fread(std::get<1>(k2).data(), sizeof(float), size_t(n[1]) * 128, f) != size_t(n[1]) * 128)
    return 3;
  std::fclose(f);
  for (int i = 0; i < n[0]; ++i)
    std::get<0>(k1)[i].coords(0) = float(i);  // distinct features (Match::operator== compares features)
  for (int i = 0; i < n[1]; ++i)
    std::get<0>(k2)[i].coords(0) = float(i);
  const auto matches = sara::match(k1, k2, 0.8f);  // SfM/Helpers/KeypointMatching.cpp:19-25
  f = std::fopen(argv[2], "wb");
  for (const auto& m : matches)
  {
    const int rec[3] = {m.x_index(), m.y_index(), m.rank()};
    const float s = m.score();
    std::fwrite(rec, sizeof(int), 3, f);
    std::fwrite(&s, sizeof(float), 1, f);
  }
  std::fclose(f);
  return 0;
}
"""


def test_cpp_ann_matcher(tmp_path):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "match_flann.npz"))
    d1, d2 = gold["d1"], gold["d2"]
    inp = tmp_path / "desc.bin"
    with open(inp, "wb") as f:
        np.array([len(d1), len(d2)], np.int32).tofile(f)
        d1.tofile(f)
        d2.tofile(f)
    src = tmp_path / "match.cpp"
    src.write_text(MATCH_PROGRAM)
    exe = tmp_path / "match"
    lib = sb.library_path()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), lib,
                           f"-Wl,-rpath,{os.path.dirname(lib)}"])
    out = tmp_path / "matches.bin"
    assert subprocess.call([str(exe), str(inp), str(out)]) == 0
    rec = np.fromfile(out, dtype=np.dtype([("x", "<i4"), ("y", "<i4"), ("rank", "<i4"), ("score", "<f4")]))
    from oracle import match as M

    ref = M.ann_match(d1, d2, 0.8)
    assert len(rec) == len(ref) > 100
    assert np.array_equal(rec["x"], ref["x_index"]) and np.array_equal(rec["y"], ref["y_index"])
    assert np.array_equal(rec["rank"], ref["rank"])
    assert np.array_equal(rec["score"].view(np.uint32), ref["score"].view(np.uint32))


DETECTOR_PROGRAM = r"""
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sara_b200.hpp"

namespace sara = sara_b200;

int main(int argc, char** argv)
{
  const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
  std::vector<float> img(static_cast<std::size_t>(w) * h);
  FILE* f = std::fopen(argv[1], "rb");
  if (!f || std::fread(img.data(), sizeof(float), img.size(), f) != img.size())
    return 2;
  std::fclose(f);
  const sara::ImageView<float> I{img.data(), w, h};
  std::vector<sara::Point2i> so;
  f = std::fopen(argv[4], "wb");
  auto dump = [&](const std::vector<sara::OERegion>& e) {
    const int n = static_cast<int>(e.size());
    std::fwrite(&n, sizeof(int), 1, f);
    for (int i = 0; i < n; ++i)
    {
      const float rec[5] = {e[i].x(), e[i].y(), e[i].shape_matrix(0, 0), float(so[i](0)), float(so[i](1))};
      std::fwrite(rec, sizeof(float), 5, f);
    }
  };
  // the reference's own test setups: test_featuredetectors_{dog,log,hessian,harris}.cpp construct the functors
  // with their default arguments (first octave 0 here to keep the frame small)
  dump(sara::ComputeDoGExtrema{sara::ImagePyramidParams(0)}(I, &so));
  dump(sara::ComputeLoGExtrema{sara::ImagePyramidParams(0, 3 + 2)}(I, &so));
  dump(sara::ComputeDoHExtrema{sara::ImagePyramidParams(0, 3 + 2, std::pow(2.f, 1.f / 3.f), 2)}(I, &so));
  dump(sara::ComputeHessianLaplaceMaxima{sara::ImagePyramidParams(0, 3 + 1)}(I, &so));
  dump(sara::ComputeHarrisLaplaceCorners{sara::ImagePyramidParams(0, 2 + 1, std::sqrt(2.f), 1), 0.04f, 1e-9f}(I, &so));
  std::fclose(f);
  try
  {
    sara::ComputeDoGExtrema{sara::ImagePyramidParams(0, 3)};
    return 3;
  }
  catch (const std::runtime_error&)
  {
  }
  return 0;
}
"""


def test_cpp_detector_functors(tmp_path):
    w, h = 480, 360
    img = S.tex(w, h, 321)
    raw = tmp_path / "img.f32"
    img.tofile(raw)
    src = tmp_path / "det.cpp"
    src.write_text(DETECTOR_PROGRAM)
    exe = tmp_path / "det"
    lib = sb.library_path()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), lib,
                           f"-Wl,-rpath,{os.path.dirname(lib)}"])
    out = tmp_path / "out.bin"
    assert subprocess.call([str(exe), str(raw), str(w), str(h), str(out)]) == 0
    blob = np.fromfile(out, dtype=np.uint8)
    ctx = sb.SiftContext(w, h, device=0)
    k3 = float(np.float32(2.0) ** np.float32(1.0 / 3.0))
    expected = [
        ctx.dog_extrema(img, sb.ImagePyramidParams(0)),
        ctx.function_extrema("log", img, sb.ImagePyramidParams(0, 5)),
        ctx.function_extrema("doh", img, sb.ImagePyramidParams(0, 5, k3, 2), 1e-6, 10.0, 1, 2),
        ctx.hessian_laplace(img, sb.ImagePyramidParams(0, 4)),
        ctx.harris_laplace(img, sb.ImagePyramidParams(0, 3, float(np.sqrt(np.float32(2.0))), 1), 0.04, 1e-9),
    ]
    off = 0
    for e in expected:
        n = int(blob[off:off + 4].view(np.int32)[0])
        rec = blob[off + 4:off + 4 + 20 * n].view(np.float32).reshape(n, 5)
        off += 4 + 20 * n
        assert n == len(e) and n > 0
        assert np.array_equal(rec[:, 0], e["x"]) and np.array_equal(rec[:, 1], e["y"])
        assert np.array_equal(rec[:, 2], e["shape"][:, 0])
        assert np.array_equal(rec[:, 3], e["s"].astype(np.float32)) and np.array_equal(rec[:, 4], e["o"].astype(np.float32))
    ctx.close()
