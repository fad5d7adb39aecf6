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
