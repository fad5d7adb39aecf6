// ========================================================================== //
// sara_b200_io.hpp -- keypoint wire format and de-duplication on the adapter's
// types (header-only; works with the POD mirrors of sara_b200.hpp and, inside a
// Sara build, with the real OERegion / Tensor_<float, 2>).
//
// Mirrors
//   write_keypoints / read_keypoints   Features/IO.hpp:78-134 (text format)
//   remove_redundant_features          Features/Utilities.cpp:23-82
// The HDF5 variants (IO.hpp:139-164) need HDF5, which this image lacks; they store
// the same two arrays as the datasets `<group>/features` and `<group>/descriptors`.
//
// Text format, one keypoint per line after the "N dim" header:
//   x y <shape_matrix in memory order> orientation int(type) <descriptor>
// with the two vectors printed as Eigen's default IOFormat prints a row vector:
// every coefficient right-aligned to the width of the widest one.
// ========================================================================== //
#pragma once

#include <algorithm>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sara_b200.hpp"

namespace sara_b200 {

  namespace io_detail {

    inline std::string eigen_row(const float* v, int n)
    {
      std::vector<std::string> s(static_cast<std::size_t>(n));
      std::size_t width = 0;
      for (int i = 0; i < n; ++i)
      {
        std::ostringstream os;
        os << v[i];
        s[i] = os.str();
        width = std::max(width, s[i].size());
      }
      std::string out;
      for (int i = 0; i < n; ++i)
      {
        if (i)
          out += ' ';
        out.append(width - s[i].size(), ' ');
        out += s[i];
      }
      return out;
    }

  }  // namespace io_detail

  //! write_keypoints(features, descriptors, name), Features/IO.hpp:107-134.
  inline bool write_keypoints(const std::vector<OERegion>& features, const DescriptorMatrix& descriptors,
                              const std::string& name)
  {
    std::ofstream file{name.c_str()};
    if (!file.is_open())
    {
      std::cerr << "Can't open file" << std::endl;
      return false;
    }
    file << features.size() << " " << descriptors.cols() << std::endl;
    for (std::size_t i = 0; i < features.size(); ++i)
    {
      const OERegion& feat = features[i];
      file << feat.x() << ' ' << feat.y() << ' ';
      file << io_detail::eigen_row(feat.shape_matrix.v, 4) << ' ';
      file << feat.orientation << ' ';
      file << int(feat.type) << ' ';
      file << io_detail::eigen_row(descriptors[static_cast<int>(i)], descriptors.cols()) << std::endl;
    }
    return true;
  }

  //! read_keypoints(features, descriptors, name), Features/IO.hpp:78-105.
  inline bool read_keypoints(std::vector<OERegion>& features, DescriptorMatrix& descriptors, const std::string& name)
  {
    std::ifstream file{name.c_str()};
    if (!file.is_open())
    {
      std::cerr << "Can't open file " << name << std::endl;
      return false;
    }
    int num_features = 0, descriptor_dim = 0;
    file >> num_features >> descriptor_dim;
    features.assign(static_cast<std::size_t>(num_features), OERegion{});
    descriptors = DescriptorMatrix{num_features, descriptor_dim};
    for (int i = 0; i < num_features; ++i)
    {
      OERegion& f = features[i];
      int feature_type = 0;
      file >> f.coords(0) >> f.coords(1);
      for (int r = 0; r < 2; ++r)  // Core/EigenExtension.hpp:162-170: rows first
        for (int c = 0; c < 2; ++c)
          file >> f.shape_matrix(r, c);
      file >> f.orientation >> feature_type;
      f.type = static_cast<OERegion::Type>(feature_type);
      float* row = descriptors.data() + static_cast<std::size_t>(i) * descriptor_dim;
      for (int j = 0; j < descriptor_dim; ++j)
        file >> row[j];
    }
    return true;
  }

  //! remove_redundant_features, Features/Utilities.cpp:23-82.
  inline void remove_redundant_features(std::vector<OERegion>& features, DescriptorMatrix& descriptors)
  {
    if (features.size() != static_cast<std::size_t>(descriptors.rows()))
      throw std::runtime_error{"Fatal: the number of features and descriptors are not equal"};
    const int dim = descriptors.cols();
    auto compare_equal = [&](std::size_t i1, std::size_t i2) {
      const float *a = descriptors[static_cast<int>(i1)], *b = descriptors[static_cast<int>(i2)];
      float sq = 0.f;
      for (int k = 0; k < dim; ++k)
        sq += (a[k] - b[k]) * (a[k] - b[k]);
      return sq < 1e-6;
    };
    auto compare_less = [&](std::size_t i1, std::size_t i2) {
      const float *a = descriptors[static_cast<int>(i1)], *b = descriptors[static_cast<int>(i2)];
      if (std::lexicographical_compare(a, a + dim, b, b + dim))
        return true;
      return compare_equal(i1, i2) && features[i1].extremum_value > features[i2].extremum_value;
    };
    std::vector<std::size_t> indices(features.size());
    std::iota(indices.begin(), indices.end(), std::size_t{0});
    std::sort(indices.begin(), indices.end(), compare_less);
    indices.erase(std::unique(indices.begin(), indices.end(), compare_equal), indices.end());

    std::vector<OERegion> unique_features(indices.size());
    DescriptorMatrix unique_descriptors{static_cast<int>(indices.size()), dim};
    for (std::size_t i = 0; i < indices.size(); ++i)
    {
      unique_features[i] = features[indices[i]];
      std::copy(descriptors[static_cast<int>(indices[i])], descriptors[static_cast<int>(indices[i])] + dim,
                unique_descriptors.data() + i * dim);
    }
    features.swap(unique_features);
    descriptors = std::move(unique_descriptors);
  }

}  // namespace sara_b200
