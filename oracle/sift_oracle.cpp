// ========================================================================== //
// oracle/sift_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement (Eigen-free C++17) of the reference's non-Halide SIFT path
// `DO::Sara::compute_sift_keypoints` (oddkiva/sara @ 83492a30).  It exists so
// that the CUDA path in `sara_b200/csrc` can be checked against the reference
// algorithm; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
// `--impl reference` legs of `bench.py` may load it.  The product never does.
//
// The reference itself cannot be compiled in this image (Eigen 3.4, Boost and
// HDF5 are hard requirements and absent), so this is a "port" oracle.  It is
// pinned against the reference's own known-answer tests (tests/test_oracle_*.py
// port them one by one); the reference holds NO golden keypoints/descriptors, so
// the end-to-end output of compute_sift_keypoints is **parity unpinned** by the
// reference binary.  What stands in for it: tests/test_independent_float64_pin.py,
// a second implementation of every stage (pyramid, extrema, orientations,
// descriptors, the wiring of the chain) in float64 numpy written from the
// reference sources, not from this file, and tests/test_eigen_boundary_pin.py
// (float64 LAPACK against the restated Eigen routines).
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference/cpp/src/DO/Sara/).  Deliberately reproduced reference quirks
// are tagged N1..N8 as in SURVEY.md section 8(a).
//
// Numerics fixed by this restatement (the reference leaves them to overload
// resolution / Eigen internals):
//   * unqualified sqrt/log/exp/floor/cos/sin on floats are the float versions;
//     where the reference expression is double by the language rules
//     (std::pow(float,int), the bilinear interpolation) it is double here too;
//   * no FP contraction (build with -ffp-contract=off, x86-64 baseline: no FMA);
//   * Eigen-only numerics: Gaussian taps use expf + a sequential sum; the 3x3
//     inverse is Eigen's cofactor formula; the eigenvalue sign test of
//     SelfAdjointEigenSolver is a cyclic Jacobi iteration in fp32; 128-vector
//     norms are sequential sums.
// ========================================================================== //

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <vector>

#ifdef _OPENMP
#  include <omp.h>
#endif

namespace oracle {

  // ------------------------------------------------------------------------ //
  // Image container: x fastest, contiguous (Core/Image/Image.hpp:44-103).
  struct Image
  {
    int w = 0, h = 0;
    std::vector<float> data;
    Image() = default;
    Image(int w_, int h_)
      : w(w_)
      , h(h_)
      , data(static_cast<size_t>(w_) * h_)
    {
    }
    float& operator()(int x, int y)
    {
      return data[static_cast<size_t>(y) * w + x];
    }
    float operator()(int x, int y) const
    {
      return data[static_cast<size_t>(y) * w + x];
    }
  };

  // ImageProcessing/ImagePyramid.hpp:29-198.
  struct PyramidParams
  {
    int first_octave_index = -1;
    int scale_count_per_octave = 6;
    float scale_geometric_factor = 1.2599210498948732f;  // std::pow(2.f, 1.f/3.f)
    int image_padding_size = 1;
    float scale_camera = 0.5f;
    float scale_initial = 1.6f;
    int num_octaves_max = std::numeric_limits<int>::max();
  };

  // ImageProcessing/ImagePyramid.hpp:206-340.
  struct Pyramid
  {
    float scale_initial = 0;
    float scale_geometric_factor = 0;
    int num_octaves = 0;
    int num_scales = 0;
    std::vector<std::vector<Image>> octaves;  // [o][s]
    std::vector<float> oct_scaling;

    void reset(int no, int ns, float s0, float k)
    {
      num_octaves = no;
      num_scales = ns;
      scale_initial = s0;
      scale_geometric_factor = k;
      octaves.assign(no, std::vector<Image>(ns));
      oct_scaling.assign(no, 0.f);
    }
    Image& operator()(int s, int o)
    {
      return octaves[o][s];
    }
    const Image& operator()(int s, int o) const
    {
      return octaves[o][s];
    }
    float operator()(int x, int y, int s, int o) const
    {
      return octaves[o][s](x, y);
    }
    // ImagePyramid.hpp:316-319: std::pow(float, int) * float is a double.
    double scale_relative_to_octave(int s) const
    {
      return std::pow(static_cast<double>(scale_geometric_factor),
                      static_cast<double>(s)) *
             static_cast<double>(scale_initial);
    }
  };

  static int g_threads_pyramid = 1;  // faithful: pyramid is single-threaded.
  static int g_threads_other = 0;    // 0 = omp default.

  // ------------------------------------------------------------------------ //
  // ImageProcessing/LinearFiltering.hpp:172-203 make_gaussian_kernel.
  std::vector<float> make_gaussian_kernel(float sigma, float gauss_truncate)
  {
    int kernel_size = static_cast<int>(2 * gauss_truncate * sigma + 1);
    kernel_size = std::max(3, kernel_size);
    if (kernel_size % 2 == 0)
      ++kernel_size;
    const int c = kernel_size / 2;

    std::vector<float> kernel(kernel_size);
    const float denom = 2 * (sigma * sigma);
    for (int i = 0; i < kernel_size; ++i)
    {
      const float d = static_cast<float>(i) - static_cast<float>(c);
      kernel[i] = expf(-(d * d) / denom);
    }
    float sum = 0.f;
    for (int i = 0; i < kernel_size; ++i)
      sum += kernel[i];
    for (int i = 0; i < kernel_size; ++i)
      kernel[i] /= sum;
    return kernel;
  }

  // LinearFiltering.hpp:44-63 convolve_array: in place, left-to-right taps,
  // accumulator starts at 0, separate multiply and add.
  static void convolve_array(float* signal, const float* kernel, int signal_size,
                             int kernel_size)
  {
    for (int i = 0; i < signal_size; ++i)
    {
      float sum = 0.f;
      for (int j = 0; j < kernel_size; ++j)
        sum += signal[i + j] * kernel[j];
      signal[i] = sum;
    }
  }

  // LinearFiltering.hpp:78-107.
  void apply_row_based_filter(const Image& src, Image& dst, const float* kernel,
                              int kernel_size, int threads)
  {
    if (src.w != dst.w || src.h != dst.h)
      throw std::domain_error{
          "Source and destination image sizes are not equal!"};
    const int w = src.w, h = src.h, half_size = kernel_size / 2;
#pragma omp parallel for num_threads(threads) if (threads != 1)
    for (int y = 0; y < h; ++y)
    {
      std::vector<float> buffer(w + half_size * 2);
      for (int x = 0; x < half_size; ++x)
        buffer[x] = src(0, y);
      for (int x = 0; x < w; ++x)
        buffer[half_size + x] = src(x, y);
      for (int x = 0; x < half_size; ++x)
        buffer[w + half_size + x] = src(w - 1, y);
      convolve_array(buffer.data(), kernel, w, kernel_size);
      for (int x = 0; x < w; ++x)
        dst(x, y) = buffer[x];
    }
  }

  // LinearFiltering.hpp:120-149 (src may alias dst: each column is copied to
  // a padded buffer first).
  void apply_column_based_filter(const Image& src, Image& dst,
                                 const float* kernel, int kernel_size,
                                 int threads)
  {
    if (src.w != dst.w || src.h != dst.h)
      throw std::domain_error{
          "Source and destination image sizes are not equal!"};
    const int w = src.w, h = src.h, half_size = kernel_size / 2;
#pragma omp parallel for num_threads(threads) if (threads != 1)
    for (int x = 0; x < w; ++x)
    {
      std::vector<float> buffer(h + half_size * 2);
      for (int y = 0; y < half_size; ++y)
        buffer[y] = src(x, 0);
      for (int y = 0; y < h; ++y)
        buffer[half_size + y] = src(x, y);
      for (int y = 0; y < half_size; ++y)
        buffer[h + half_size + y] = src(x, h - 1);
      convolve_array(buffer.data(), kernel, h, kernel_size);
      for (int y = 0; y < h; ++y)
        dst(x, y) = buffer[y];
    }
  }

  // LinearFiltering.cpp:30-68 (#else branch) + LinearFiltering.hpp:446-454.
  Image gaussian(const Image& src, float sigma, float gauss_truncate = 4.f)
  {
    Image dst(src.w, src.h);
    const auto kernel = make_gaussian_kernel(sigma, gauss_truncate);
    const int th = g_threads_pyramid;
    apply_row_based_filter(src, dst, kernel.data(),
                           static_cast<int>(kernel.size()), th);
    apply_column_based_filter(dst, dst, kernel.data(),
                              static_cast<int>(kernel.size()), th);
    return dst;
  }

  // ------------------------------------------------------------------------ //
  // ImageProcessing/Interpolation.hpp:34-78, N = 2, T = float.
  // Iteration order of the 2x2 sub-array is x fastest
  // (Core/ArrayIterators/Utilities.hpp:152-171, ColMajor incrementer); the
  // accumulator starts at PixelTraits<double>::min() == 0.
  double interpolate(const Image& image, double px, double py)
  {
    const double pos[2] = {px, py};
    const int size[2] = {image.w, image.h};
    int start[2];
    double frac[2];
    for (int i = 0; i < 2; ++i)
    {
      if (pos[i] < 0 || pos[i] >= size[i])
        throw std::out_of_range{
            "Cannot interpolate: position is out of image domain"};
      double ith_int_part;
      frac[i] = std::modf(pos[i], &ith_int_part);
      start[i] = static_cast<int>(ith_int_part);
    }
    double value = 0.;
    for (int y = start[1]; y < start[1] + 2; ++y)
      for (int x = start[0]; x < start[0] + 2; ++x)
      {
        double weight = 1.;
        weight *= (x == start[0]) ? (1. - frac[0]) : frac[0];
        weight *= (y == start[1]) ? (1. - frac[1]) : frac[1];
        const int ox = x < image.w ? 0 : -1;
        const int oy = y < image.h ? 0 : -1;
        value += weight * static_cast<double>(image(x + ox, y + oy));
      }
    return value;
  }

  // ImageProcessing/Resize.cpp:86-128 (#else branch) enlarge(src, dst).
  void enlarge(const Image& src, Image& dst, int threads)
  {
    if (dst.w < src.w || dst.h < src.h)
      throw std::range_error{"The destination image must have smaller sizes "
                             "than the source image!"};
    if (std::min(dst.w, dst.h) <= 0)
      throw std::range_error{
          "The sizes of the destination image must be positive!"};
    const int wh = dst.w * dst.h;
    const double sx = static_cast<double>(src.w) / static_cast<double>(dst.w);
    const double sy = static_cast<double>(src.h) / static_cast<double>(dst.h);
#pragma omp parallel for num_threads(threads) if (threads != 1)
    for (int xy = 0; xy < wh; ++xy)
    {
      const int w = dst.w;
      const int y = xy / w;
      const int x = xy - y * w;
      dst(x, y) = static_cast<float>(interpolate(src, x * sx, y * sy));
    }
  }

  // Resize.hpp:212-216 enlarge(image, double fact) -> Resize.hpp:190-209.
  Image enlarge(const Image& image, double fact)
  {
    const int nw = static_cast<int>(static_cast<double>(image.w) * fact);
    const int nh = static_cast<int>(static_cast<double>(image.h) * fact);
    Image dst(nw, nh);
    enlarge(image, dst, g_threads_pyramid);
    return dst;
  }

  // Resize.cpp:31-61 (#else branch) scale(): nearest sample with a float ratio.
  void scale(const Image& src, Image& dst, int threads)
  {
    const float sx = static_cast<float>(src.w) / static_cast<float>(dst.w);
    const float sy = static_cast<float>(src.h) / static_cast<float>(dst.h);
    const int w = dst.w;
    const int wh = dst.w * dst.h;
#pragma omp parallel for num_threads(threads) if (threads != 1)
    for (int xy = 0; xy < wh; ++xy)
    {
      const int y = xy / w;
      const int x = xy - y * w;
      const int xi = static_cast<int>(static_cast<float>(x) * sx);
      const int yi = static_cast<int>(static_cast<float>(y) * sy);
      dst(x, y) = src(xi, yi);
    }
  }

  // Resize.cpp:64-83 downscale().
  Image downscale(const Image& src, int fact)
  {
    Image dst(src.w / fact, src.h / fact);
    scale(src, dst, g_threads_pyramid);
    return dst;
  }

  // ------------------------------------------------------------------------ //
  // ImageProcessing/GaussianPyramid.hpp:35-125 gaussian_pyramid<float>.
  Pyramid gaussian_pyramid(const Image& image, const PyramidParams& params,
                           float gauss_truncate)
  {
    const float resize_factor =
        std::pow(2.f, -static_cast<float>(params.first_octave_index));
    const float camera_sigma = params.scale_camera * resize_factor;
    const float init_sigma = params.scale_initial;

    Image I;
    if (params.first_octave_index < 0)
      I = enlarge(image, resize_factor);  // N4: no blur at all.
    else if (params.first_octave_index > 0)
    {
      if (camera_sigma < init_sigma)
      {
        const float sigma =
            std::sqrt(init_sigma * init_sigma - camera_sigma * camera_sigma);
        I = gaussian(image, sigma, gauss_truncate);
      }
      else
        I = image;
      I = downscale(I, static_cast<int>(std::round(1 / resize_factor)));
    }
    else
    {
      if (camera_sigma < init_sigma)
      {
        const float sigma =
            std::sqrt(init_sigma * init_sigma - camera_sigma * camera_sigma);
        I = gaussian(image, sigma);  // N5: default truncate 4.
      }
      else
        I = image;
    }

    const int l = std::min(I.w, I.h);
    const int b = params.image_padding_size;
    const int num_octaves =
        std::min(static_cast<int>(logf(l / (2.f * b)) / logf(2.f)),
                 params.num_octaves_max);

    const float k = params.scale_geometric_factor;
    const int num_scales = params.scale_count_per_octave;
    const int downscale_index =
        static_cast<int>(floorf(logf(2.f) / logf(k)));  // N3: 2 for k=2^(1/3).

    Pyramid G;
    G.reset(std::max(num_octaves, 0), num_scales, init_sigma, k);

    for (int o = 0; o < num_octaves; ++o)
    {
      G.oct_scaling[o] =
          (o == 0) ? 1 / resize_factor : G.oct_scaling[o - 1] * 2;

      float sigma_s_1 = init_sigma;
      if (o == 0)
        G(0, o) = std::move(I);
      else
        G(0, o) = downscale(G(downscale_index, o - 1), 2);

      for (int s = 1; s < num_scales; ++s)
      {
        const float ks = k * sigma_s_1;
        const float sigma = sqrtf(ks * ks - sigma_s_1 * sigma_s_1);
        G(s, o) = gaussian(G(s - 1, o), sigma);
        sigma_s_1 *= k;
      }
    }
    return G;
  }

  // ImageProcessing/GaussianPyramid.cpp:23-51 (#else branch).
  Pyramid difference_of_gaussians_pyramid(const Pyramid& gaussians)
  {
    Pyramid D;
    D.reset(gaussians.num_octaves, gaussians.num_scales - 1,
            gaussians.scale_initial, gaussians.scale_geometric_factor);
    for (int o = 0; o < D.num_octaves; ++o)
    {
      D.oct_scaling[o] = gaussians.oct_scaling[o];
      for (int s = 0; s < D.num_scales; ++s)
      {
        const Image& a = gaussians(s + 1, o);
        const Image& b = gaussians(s, o);
        Image d(a.w, a.h);
        const size_t n = d.data.size();
        for (size_t i = 0; i < n; ++i)
          d.data[i] = a.data[i] - b.data[i];
        D(s, o) = std::move(d);
      }
    }
    return D;
  }

  // ------------------------------------------------------------------------ //
  // ImageProcessing/Extrema.hpp:28-47 CompareWithNeighborhood3.
  template <typename Compare>
  static bool compare_with_neighborhood3(float val, int x, int y,
                                         const Image& I, bool with_center)
  {
    Compare cmp;
    for (int v = -1; v <= 1; ++v)
      for (int u = -1; u <= 1; ++u)
      {
        if (u == 0 && v == 0 && !with_center)
          continue;
        if (!cmp(val, I(x + u, y + v)))
          return false;
      }
    return true;
  }

  // Extrema.hpp:63-75 LocalScaleSpaceExtremum.
  template <typename Compare>
  static bool local_scale_space_extremum(int x, int y, int s, int o,
                                         const Pyramid& I)
  {
    const float v = I(x, y, s, o);
    return compare_with_neighborhood3<Compare>(v, x, y, I(s - 1, o), true) &&
           compare_with_neighborhood3<Compare>(v, x, y, I(s, o), false) &&
           compare_with_neighborhood3<Compare>(v, x, y, I(s + 1, o), true);
  }

  // ImageProcessing/Differential.hpp:191-226 Hessian functor, N = 2, with the
  // replicated-border variants (never hit when padding >= 1).
  static void hessian2(const Image& I, int x, int y, float H[4])
  {
    const float c = I(x, y);
    {
      const float next = x == I.w - 1 ? c : I(x + 1, y);
      const float prev = x == 0 ? c : I(x - 1, y);
      H[0] = next - 2.f * c + prev;
    }
    {
      const float next = y == I.h - 1 ? c : I(x, y + 1);
      const float prev = y == 0 ? c : I(x, y - 1);
      H[3] = next - 2.f * c + prev;
    }
    {
      const int nx = x == I.w - 1 ? 0 : 1, px = x == 0 ? 0 : -1;
      const int ny = y == I.h - 1 ? 0 : 1, py = y == 0 ? 0 : -1;
      H[1] = H[2] = (I(x + nx, y + ny) - I(x + px, y + ny) -
                     I(x + nx, y + py) + I(x + px, y + py)) /
                    4.f;
    }
  }

  // FeatureDetectors/RefineExtremum.cpp:24-30 on_edge.
  bool on_edge(const Image& I, int x, int y, float edge_ratio)
  {
    float H[4];
    hessian2(I, x, y, H);
    const float tr = H[0] + H[3];
    const float det = H[0] * H[3] - H[1] * H[2];
    const float e1 = edge_ratio + 1.f;
    return (tr * tr) * edge_ratio >= (e1 * e1) * std::abs(det);
  }

  // ImageProcessing/GaussianPyramid.hpp:184-198 gradient(I, x, y, s, o).
  static void gradient3(const Pyramid& I, int x, int y, int s, int o, float d[3])
  {
    d[0] = (I(x + 1, y, s, o) - I(x - 1, y, s, o)) / 2.f;
    d[1] = (I(x, y + 1, s, o) - I(x, y - 1, s, o)) / 2.f;
    d[2] = (I(x, y, s + 1, o) - I(x, y, s - 1, o)) / 2.f;
  }

  // GaussianPyramid.hpp:203-233 hessian(I, x, y, s, o). Row-major H[3*i+j].
  static void hessian3(const Pyramid& I, int x, int y, int s, int o, float H[9])
  {
    const float c = I(x, y, s, o);
    H[0] = I(x + 1, y, s, o) - 2.f * c + I(x - 1, y, s, o);
    H[4] = I(x, y + 1, s, o) - 2.f * c + I(x, y - 1, s, o);
    H[8] = I(x, y, s + 1, o) - 2.f * c + I(x, y, s - 1, o);
    H[1] = H[3] = (I(x + 1, y + 1, s, o) - I(x - 1, y + 1, s, o) -
                   I(x + 1, y - 1, s, o) + I(x - 1, y - 1, s, o)) /
                  4.f;
    H[2] = H[6] = (I(x + 1, y, s + 1, o) - I(x - 1, y, s + 1, o) -
                   I(x + 1, y, s - 1, o) + I(x - 1, y, s - 1, o)) /
                  4.f;
    H[5] = H[7] = (I(x, y + 1, s + 1, o) - I(x, y - 1, s + 1, o) -
                   I(x, y + 1, s - 1, o) + I(x, y - 1, s - 1, o)) /
                  4.f;
  }

  // Stand-in for Eigen::SelfAdjointEigenSolver<Matrix3f>::eigenvalues()
  // (RefineExtremum.cpp:74-81; only the sign of the largest eigenvalue is
  // used).  Cyclic Jacobi in fp32, at most 8 sweeps, +,-,*,/,sqrt only so the
  // CUDA kernel can repeat it bit for bit.
  void sym3_eigenvalues(const float H[9], float lambda[3])
  {
    float a00 = H[0], a11 = H[4], a22 = H[8];
    float a01 = H[1], a02 = H[2], a12 = H[5];
    for (int sweep = 0; sweep < 8; ++sweep)
    {
      if (a01 == 0.f && a02 == 0.f && a12 == 0.f)
        break;
      // (p, q) = (0, 1)
      if (a01 != 0.f)
      {
        const float theta = (a11 - a00) / (2.f * a01);
        const float at = std::abs(theta);
        float t = 1.f / (at + sqrtf(theta * theta + 1.f));
        if (theta < 0.f)
          t = -t;
        const float c = 1.f / sqrtf(t * t + 1.f);
        const float sn = t * c;
        const float tau = t * a01;
        a00 = a00 - tau;
        a11 = a11 + tau;
        a01 = 0.f;
        const float b02 = c * a02 - sn * a12;
        const float b12 = sn * a02 + c * a12;
        a02 = b02;
        a12 = b12;
      }
      // (p, q) = (0, 2)
      if (a02 != 0.f)
      {
        const float theta = (a22 - a00) / (2.f * a02);
        const float at = std::abs(theta);
        float t = 1.f / (at + sqrtf(theta * theta + 1.f));
        if (theta < 0.f)
          t = -t;
        const float c = 1.f / sqrtf(t * t + 1.f);
        const float sn = t * c;
        const float tau = t * a02;
        a00 = a00 - tau;
        a22 = a22 + tau;
        a02 = 0.f;
        const float b01 = c * a01 - sn * a12;
        const float b12 = sn * a01 + c * a12;
        a01 = b01;
        a12 = b12;
      }
      // (p, q) = (1, 2)
      if (a12 != 0.f)
      {
        const float theta = (a22 - a11) / (2.f * a12);
        const float at = std::abs(theta);
        float t = 1.f / (at + sqrtf(theta * theta + 1.f));
        if (theta < 0.f)
          t = -t;
        const float c = 1.f / sqrtf(t * t + 1.f);
        const float sn = t * c;
        const float tau = t * a12;
        a11 = a11 - tau;
        a22 = a22 + tau;
        a12 = 0.f;
        const float b01 = c * a01 - sn * a02;
        const float b02 = sn * a01 + c * a02;
        a01 = b01;
        a02 = b02;
      }
    }
    lambda[0] = a00;
    lambda[1] = a11;
    lambda[2] = a22;
  }

  // Eigen 3.4 Matrix3f::inverse(): cofactors times 1/det
  // (Eigen/src/LU/InverseImpl.h, compute_inverse<..., 3>); used at
  // RefineExtremum.cpp:85.  Row-major in/out.
  void inverse3(const float m[9], float r[9])
  {
    auto M = [&](int i, int j) { return m[3 * i + j]; };
    auto cof = [&](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1);
    };
    const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const float det = (c0 * M(0, 0) + c1 * M(1, 0)) + c2 * M(2, 0);
    const float invdet = 1.f / det;
    r[3 * 0 + 0] = c0 * invdet;
    r[3 * 0 + 1] = c1 * invdet;
    r[3 * 0 + 2] = c2 * invdet;
    r[3 * 1 + 0] = cof(0, 1) * invdet;
    r[3 * 1 + 1] = cof(1, 1) * invdet;
    r[3 * 1 + 2] = cof(2, 1) * invdet;
    r[3 * 2 + 0] = cof(0, 2) * invdet;
    r[3 * 2 + 1] = cof(1, 2) * invdet;
    r[3 * 2 + 2] = cof(2, 2) * invdet;
  }

  // Optional trace of every Newton iteration (tests/test_eigen_boundary_pin.py checks the
  // Eigen-boundary restatements -- eigenvalue sign test, cofactor inverse -- against an
  // independent float64 solver).  25 floats per record:
  // x, y, s, o, type, iteration, H[9], g[3], lambda[3], h[3], decision
  // (decision: 0 = "not definite": h = 0, stop; 1 = Newton step computed).
  static bool g_trace_on = false;
  static std::vector<float> g_trace;

  // FeatureDetectors/RefineExtremum.cpp:32-130 refine_extremum (3-D).
  // `type` arrives as the uint8 map value: 1 for maxima, 255 for minima (N2).
  bool refine_extremum(const Pyramid& I, int x, int y, int s, int o, int type,
                       float pos[3], float& val, int border_sz, int num_iter)
  {
    float D_prime[3] = {0.f, 0.f, 0.f};
    float D_second[9];
    float h[3] = {0.f, 0.f, 0.f};
    float lambda[3];

    pos[0] = float(x);
    pos[1] = float(y);
    pos[2] = static_cast<float>(I.scale_relative_to_octave(s));

    for (int i = 0; i < num_iter; ++i)
    {
      if (x < border_sz || x >= I(s, o).w - border_sz || y < border_sz ||
          y >= I(s, o).h - border_sz || s < 1 || s >= I.num_scales - 1)
        break;  // N8: D_prime and h keep their previous values.

      gradient3(I, x, y, s, o, D_prime);
      hessian3(I, x, y, s, o, D_second);

      sym3_eigenvalues(D_second, lambda);
      const float ft = float(type);
      const float lmax =
          std::max(std::max(lambda[0] * ft, lambda[1] * ft), lambda[2] * ft);
      auto trace = [&](float decision) {
        if (!g_trace_on)
          return;
        float rec[25] = {float(x), float(y), float(s), float(o), ft, float(i)};
        for (int k = 0; k < 9; ++k)
          rec[6 + k] = D_second[k];
        for (int k = 0; k < 3; ++k)
        {
          rec[15 + k] = D_prime[k];
          rec[18 + k] = lambda[k];
          rec[21 + k] = h[k];
        }
        rec[24] = decision;
#pragma omp critical(oracle_trace)
        g_trace.insert(g_trace.end(), rec, rec + 25);
      };
      if (lmax >= 0)
      {
        h[0] = h[1] = h[2] = 0.f;
        trace(0.f);
        break;
      }

      float inv[9];
      inverse3(D_second, inv);
      for (int r = 0; r < 3; ++r)
        h[r] = ((-inv[3 * r + 0]) * D_prime[0] + (-inv[3 * r + 1]) * D_prime[1]) +
               (-inv[3 * r + 2]) * D_prime[2];
      trace(1.f);

      if (std::max(std::abs(h[0]), std::abs(h[1])) > 1.5f)
        return false;

      if (std::min(std::abs(h[0]), std::abs(h[1])) > 0.6f)
      {
        x += h[0] > 0 ? 1 : -1;
        y += h[1] > 0 ? 1 : -1;
        continue;
      }
      break;
    }

    pos[0] = float(x);
    pos[1] = float(y);
    pos[2] = static_cast<float>(I.scale_relative_to_octave(s));
    const float oldval = I(x, y, s, o);
    const float newval =
        oldval +
        0.5f * ((D_prime[0] * h[0] + D_prime[1] * h[1]) + D_prime[2] * h[2]);

    if ((type == 1 && oldval <= newval) || (type == -1 && oldval >= newval))
    {
      pos[0] += h[0];
      pos[1] += h[1];
      pos[2] *= std::pow(I.scale_geometric_factor, h[2]);
      val = newval;
    }
    return true;
  }

  // Features/Feature.hpp:40-179 OERegion, as a POD (no Eigen).
  struct Keypoint
  {
    float x, y;             // coords
    float shape[4];         // shape_matrix, column-major (isotropic here)
    float orientation;      //
    float extremum_value;   //
    std::uint8_t type;      // OERegion::Type, stays Undefined (= 11)
    std::int8_t extremum_type;  // -1 Min, 1 Max
    std::int16_t pad_;
    std::int32_t s, o;      // scale_octave_pairs[i]
    std::int32_t xi, yi;    // raster slot the extremum was emitted at
  };

  // Features/Feature.hpp:78-82: shape_matrix = I * std::pow(scale, -2);
  // std::pow(float, int) is evaluated in double, then narrowed by Eigen.
  static Keypoint make_oeregion(float x, float y, float scale)
  {
    Keypoint k;
    std::memset(&k, 0, sizeof k);
    k.x = x;
    k.y = y;
    const float a = static_cast<float>(std::pow(static_cast<double>(scale), -2.0));
    k.shape[0] = a;
    k.shape[1] = 0.f;
    k.shape[2] = 0.f;
    k.shape[3] = a;
    k.orientation = 0.f;
    k.extremum_value = 0.f;
    k.type = 11;            // Type::Undefined
    k.extremum_type = -2;   // ExtremumType::Undefined
    return k;
  }

  // Features/Feature.cpp:28-39 OERegion::radius(0) for an isotropic shape
  // matrix a*I: JacobiSVD gives singular values (a, a), U = +-I, so the result
  // is sqrt((1/sqrt(a))^2) == 1/sqrt(a).
  static float oeregion_scale(const Keypoint& k)
  {
    const float r = 1.f / sqrtf(k.shape[0]);
    return sqrtf(r * r + 0.f * 0.f);
  }

  // FeatureDetectors/RefineExtremum.cpp:363-521 local_scale_space_extrema
  // (non-Halide branch).
  std::vector<Keypoint> local_scale_space_extrema(const Pyramid& I, int s, int o,
                                                  float extremum_thres,
                                                  float edge_ratio_thres,
                                                  int img_padding_sz,
                                                  int refine_iterations)
  {
    const int w = I(s, o).w;
    const int h = I(s, o).h;
    const int wh = w * h;
    const int th = g_threads_other;

    std::vector<std::uint8_t> map(static_cast<size_t>(wh), 0);  // N2

#pragma omp parallel for num_threads(th > 0 ? th : omp_get_max_threads())
    for (int xy = 0; xy < wh; ++xy)
    {
      const int y = xy / w;
      const int x = xy - y * w;
      const bool in_domain = img_padding_sz <= x && x < w - img_padding_sz &&
                             img_padding_sz <= y && y < h - img_padding_sz;
      if (!in_domain)
        continue;

      int type = 0;
      if (local_scale_space_extremum<std::greater_equal<float>>(x, y, s, o, I))
        type = 1;
      else if (local_scale_space_extremum<std::less_equal<float>>(x, y, s, o, I))
        type = -1;
      else
        continue;

      if (std::abs(I(x, y, s, o)) < 0.8f * extremum_thres)
        continue;
      if (on_edge(I(s, o), x, y, edge_ratio_thres))
        continue;
      map[xy] = static_cast<std::uint8_t>(type);
    }

    std::vector<float> loc(static_cast<size_t>(wh) * 3, 0.f);
    std::vector<float> value(static_cast<size_t>(wh), 0.f);
#pragma omp parallel for num_threads(th > 0 ? th : omp_get_max_threads())
    for (int xy = 0; xy < wh; ++xy)
    {
      const int y = xy / w;
      const int x = xy - y * w;
      const std::uint8_t type = map[xy];
      if (type == 0)
        continue;
      float* pos = &loc[static_cast<size_t>(xy) * 3];
      float& val = value[xy];
      val = I(x, y, s, o);
      refine_extremum(I, x, y, s, o, type, pos, val, img_padding_sz,
                      refine_iterations);
      if (std::abs(val) < extremum_thres)
        map[xy] = 0;
    }

    std::vector<Keypoint> extrema;
    extrema.reserve(10000);
    for (int xy = 0; xy < wh; ++xy)
    {
      const int y = xy / w;
      const int x = xy - y * w;
      const std::uint8_t type = map[xy];
      if (type == 0)
        continue;
      const float* pos = &loc[static_cast<size_t>(xy) * 3];
      Keypoint dog = make_oeregion(pos[0], pos[1], pos[2]);
      dog.extremum_value = value[xy];
      dog.extremum_type = type == 1 ? 1 : -1;
      dog.s = s;
      dog.o = o;
      dog.xi = x;  // N7: emitted at the original raster slot.
      dog.yi = y;
      extrema.push_back(dog);
    }
    return extrema;
  }

  // ------------------------------------------------------------------------ //
  // Polar gradient pyramid: interleaved (mag, ori) per pixel.
  struct PolarPyramid
  {
    int num_octaves = 0, num_scales = 0;
    std::vector<std::vector<int>> w, h;
    std::vector<std::vector<std::vector<float>>> data;  // [o][s][2*(y*w+x)]
  };

  // FeatureDescriptors/Orientation.cpp:24-56 (#else branch) on top of
  // ImageProcessing/Differential.hpp:46-61 (Gradient functor, one-sided /2 at
  // the borders).
  void gradient_polar_coordinates(const Image& f, std::vector<float>& out,
                                  int threads)
  {
    const int w = f.w, h = f.h;
    out.resize(static_cast<size_t>(w) * h * 2);
#pragma omp parallel for num_threads(threads) if (threads != 1)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
      {
        float gx, gy;
        if (w == 1)
          gx = (f(x, y) - f(x, y)) / 2;
        else if (x == 0)
          gx = (f(x + 1, y) - f(x, y)) / 2;
        else if (x == w - 1)
          gx = (f(x, y) - f(x - 1, y)) / 2;
        else
          gx = (f(x + 1, y) - f(x - 1, y)) / 2;
        if (h == 1)
          gy = (f(x, y) - f(x, y)) / 2;
        else if (y == 0)
          gy = (f(x, y + 1) - f(x, y)) / 2;
        else if (y == h - 1)
          gy = (f(x, y) - f(x, y - 1)) / 2;
        else
          gy = (f(x, y + 1) - f(x, y - 1)) / 2;
        const float r = 2 * sqrtf(gx * gx + gy * gy);
        const float theta = atan2f(gy, gx);
        const size_t i = (static_cast<size_t>(y) * w + x) * 2;
        out[i] = r;
        out[i + 1] = theta;
      }
  }

  // Orientation.hpp:69-86.
  PolarPyramid gradient_polar_coordinates(const Pyramid& G, int threads)
  {
    PolarPyramid P;
    P.num_octaves = G.num_octaves;
    P.num_scales = G.num_scales;
    P.w.assign(G.num_octaves, std::vector<int>(G.num_scales));
    P.h.assign(G.num_octaves, std::vector<int>(G.num_scales));
    P.data.assign(G.num_octaves,
                  std::vector<std::vector<float>>(G.num_scales));
    for (int o = 0; o < G.num_octaves; ++o)
      for (int s = 0; s < G.num_scales; ++s)
      {
        P.w[o][s] = G(s, o).w;
        P.h[o][s] = G(s, o).h;
        gradient_polar_coordinates(G(s, o), P.data[o][s], threads);
      }
    return P;
  }

  // Geometry/Tools/Utilities.hpp:28-32.
  static inline int int_round(float x)
  {
    return static_cast<int>(std::round(x));
  }

  static const float kPi = static_cast<float>(M_PI);
  static const float kTwoPi = static_cast<float>(2. * M_PI);

  // FeatureDescriptors/Orientation.hpp:91-139 compute_orientation_histogram,
  // N = 36.
  void compute_orientation_histogram(float hist[36], const float* grad, int w,
                                     int h, float x, float y, float s,
                                     float patch_truncation_factor,
                                     float blur_factor)
  {
    constexpr int N = 36;
    for (int i = 0; i < N; ++i)
      hist[i] = 0.f;
    const int rounded_x = int_round(x);
    const int rounded_y = int_round(y);
    const float sigma = s * blur_factor;
    const int patch_radius = int_round(sigma * patch_truncation_factor);
    for (int v = -patch_radius; v <= patch_radius; ++v)
      for (int u = -patch_radius; u <= patch_radius; ++u)
      {
        if (rounded_x + u < 0 || rounded_x + u >= w || rounded_y + v < 0 ||
            rounded_y + v >= h)
          continue;
        const size_t i =
            (static_cast<size_t>(rounded_y + v) * w + (rounded_x + u)) * 2;
        const float mag = grad[i];
        float ori = grad[i + 1];
        ori = ori < 0 ? ori + kTwoPi : ori;
        int bin_index = static_cast<int>(floorf(ori / kTwoPi * N));
        bin_index %= N;
        const float weight =
            expf(-(u * u + v * v) / (2.f * sigma * sigma));
        hist[bin_index] += weight * mag;
      }
  }

  // Orientation.hpp:147-165 lowe_smooth_histogram.
  void lowe_smooth_histogram(float hist[36], int num_iters)
  {
    constexpr int N = 36;
    for (int iter = 0; iter < num_iters; ++iter)
    {
      const float first = hist[0];
      float prev = hist[N - 1];
      for (int i = 0; i < N - 1; ++i)
      {
        const float val = (prev + hist[i] + hist[i + 1]) / 3.f;
        prev = hist[i];
        hist[i] = val;
      }
      hist[N - 1] = (prev + hist[N - 1] + first) / 3.f;
    }
  }

  // Orientation.hpp:176-189 find_peaks.
  int find_peaks(const float hist[36], float peak_ratio_thres, int peaks[36])
  {
    constexpr int N = 36;
    float max = hist[0];
    for (int i = 1; i < N; ++i)
      max = std::max(max, hist[i]);
    int n = 0;
    for (int i = 0; i < N; ++i)
      if (hist[i] >= peak_ratio_thres * max &&
          hist[i] > hist[(i - 1 + N) % N] && hist[i] > hist[(i + 1) % N])
        peaks[n++] = i;
    return n;
  }

  // Orientation.hpp:193-214 refine_peak.
  float refine_peak(const float hist[36], int i)
  {
    constexpr int N = 36;
    const float y0 = hist[(i - 1 + N) % N];
    const float y1 = hist[i];
    const float y2 = hist[(i + 1) % N];
    const float fprime = (y2 - y0) / 2.f;
    const float fsecond = y0 - 2.f * y1 + y2;
    const float h = -fprime / fsecond;
    return float(i) + 0.5f + h;
  }

  // Orientation.cpp:90-118 ComputeDominantOrientations::operator()(gradients,
  // x, y, sigma).
  int dominant_orientations(const float* grad, int w, int h, float x, float y,
                            float sigma, float peak_ratio_thres,
                            float patch_truncation_factor, float blur_factor,
                            float out[36])
  {
    constexpr int O = 36;
    float hist[O];
    compute_orientation_histogram(hist, grad, w, h, x, y, sigma,
                                  patch_truncation_factor, blur_factor);
    lowe_smooth_histogram(hist, 6);
    int peaks[O];
    const int n = find_peaks(hist, peak_ratio_thres, peaks);
    for (int i = 0; i < n; ++i)
    {
      float p = refine_peak(hist, peaks[i]);
      p *= kTwoPi / O;
      if (p > kPi)
        p -= 2.f * kPi;
      out[i] = p;
    }
    return n;
  }

  // FeatureDescriptors/SIFT.hpp:204-238 accumulate (N6: modf truncation).
  static void sift_accumulate(float* h, float px, float py, float ori,
                              float weight, float mag)
  {
    constexpr int N = 4, O = 8;
    float xif, yif, oriif;
    const float xfrac = std::modf(px, &xif);
    const float yfrac = std::modf(py, &yif);
    const float orifrac = std::modf(ori, &oriif);
    const int xi = int(xif);
    const int yi = int(yif);
    const int orii = int(oriif);
    for (int dy = 0; dy < 2; ++dy)
    {
      const int y = yi + dy;
      if (y < 0 || y >= N)
        continue;
      const float wy = (dy == 0) ? 1 - yfrac : yfrac;
      for (int dx = 0; dx < 2; ++dx)
      {
        const int x = xi + dx;
        if (x < 0 || x >= N)
          continue;
        const float wx = (dx == 0) ? 1 - xfrac : xfrac;
        for (int dori = 0; dori < 2; ++dori)
        {
          const int o = (orii + dori) % O;
          const float wo = (dori == 0) ? 1 - orifrac : orifrac;
          h[N * O * y + x * O + o] += wy * wx * wo * weight * mag;
        }
      }
    }
  }

  static void normalize128(float* h)
  {
    float sq = 0.f;
    for (int i = 0; i < 128; ++i)
      sq += h[i] * h[i];
    const float n = sqrtf(sq);
    for (int i = 0; i < 128; ++i)
      h[i] /= n;
  }

  // FeatureDescriptors/SIFT.hpp:62-145 ComputeSIFTDescriptor<4,8>::operator().
  void sift_descriptor(float x, float y, float s, float theta, const float* grad,
                       int w, int h, float bin_scale_unit_length,
                       float max_bin_value, bool do_normalization, float* desc)
  {
    constexpr int N = 4, O = 8, Dim = 128;
    const float lambda = bin_scale_unit_length;
    const float l = lambda * s;
    const float r = sqrtf(2.f) * l * (N + 1) / 2.f;
    const float ct = cosf(theta), st = sinf(theta);
    const float T00 = ct / l, T01 = st / l, T10 = -st / l, T11 = ct / l;

    for (int i = 0; i < Dim; ++i)
      desc[i] = 0.f;

    const int rounded_r = int_round(r);
    const int rounded_x = int_round(x);
    const int rounded_y = int_round(y);
    for (int v = -rounded_r; v <= rounded_r; ++v)
      for (int u = -rounded_r; u <= rounded_r; ++u)
      {
        float px = T00 * float(u) + T01 * float(v);
        float py = T10 * float(u) + T11 * float(v);
        if (rounded_x + u < 0 || rounded_x + u >= w || rounded_y + v < 0 ||
            rounded_y + v >= h)
          continue;
        constexpr float sigma = N * N * 0.25f;
        const float weight = expf(-(px * px + py * py) / (2.f * sigma));
        const size_t i =
            (static_cast<size_t>(rounded_y + v) * w + (rounded_x + u)) * 2;
        const float mag = grad[i];
        float ori = grad[i + 1] - theta;
        ori = ori < 0.f ? ori + 2.f * kPi : ori;
        ori *= float(O) / (2.f * kPi);
        px += N / 2.f - 0.5f;
        py += N / 2.f - 0.5f;
        if (std::min(px, py) <= -1.f || std::max(px, py) >= float(N))
          continue;
        sift_accumulate(desc, px, py, ori, weight, mag);
      }

    if (do_normalization)
    {
      // SIFT.hpp:241-252 normalize(): L2, clamp, L2; then SIFT.hpp:128.
      normalize128(desc);
      for (int i = 0; i < Dim; ++i)
        desc[i] = std::min(desc[i], max_bin_value);
      normalize128(desc);
      for (int i = 0; i < Dim; ++i)
        desc[i] = std::min(desc[i] * 512.f, 255.f);
    }
  }

  // ------------------------------------------------------------------------ //
  // Whole-pipeline result, kept alive behind an opaque handle.
  struct Result
  {
    Pyramid G, D;
    std::vector<Keypoint> extrema;    // after ComputeDoGExtrema (octave coords)
    std::vector<Keypoint> keypoints;  // final, rescaled to image coordinates
    std::vector<Keypoint> oriented;   // after orientation, octave coordinates
    std::vector<float> descriptors;   // N' x 128 row-major
    double ms_dog = 0, ms_grad = 0, ms_ori = 0, ms_desc = 0;
  };

  static double now_ms()
  {
#ifdef _OPENMP
    return omp_get_wtime() * 1e3;
#else
    return 0.;
#endif
  }

  // FeatureDetectors/DoG.cpp:23-87 ComputeDoGExtrema::operator().
  void compute_dog_extrema(const Image& image, const PyramidParams& pp,
                           float gauss_truncate, float extremum_thres,
                           float edge_ratio_thres, int img_padding_sz,
                           int extremum_refinement_iter, Result& R)
  {
    if (pp.scale_count_per_octave < 4)  // DoG.hpp:86-89
      throw std::runtime_error{
          "Error: The extraction of DoG extrema needs (1 + 3) = 4 scales per "
          "octave at the very minimum!"};
    R.G = gaussian_pyramid(image, pp, gauss_truncate);
    R.D = difference_of_gaussians_pyramid(R.G);
    R.extrema.clear();
    for (int o = 0; o < R.D.num_octaves; ++o)
      for (int s = 1; s < R.D.num_scales - 1; ++s)
      {
        auto e = local_scale_space_extrema(R.D, s, o, extremum_thres,
                                           edge_ratio_thres, img_padding_sz,
                                           extremum_refinement_iter);
        R.extrema.insert(R.extrema.end(), e.begin(), e.end());
      }
  }

  // ------------------------------------------------------------------------ //
  // Sibling detectors on the same pyramid.
  // ImageProcessing/Differential.hpp:106-135 Laplacian functor (N = 2, borders replicated) and
  // ImageProcessing/GaussianPyramid.hpp:156-178 laplacian_pyramid: every layer times
  // float(square(scale_relative_to_octave(s))) (the square is taken in double).
  Pyramid laplacian_pyramid(const Pyramid& gaussians)
  {
    Pyramid L;
    L.reset(gaussians.num_octaves, gaussians.num_scales, gaussians.scale_initial,
            gaussians.scale_geometric_factor);
    for (int o = 0; o < L.num_octaves; ++o)
    {
      L.oct_scaling[o] = gaussians.oct_scaling[o];
      for (int s = 0; s < L.num_scales; ++s)
      {
        const Image& g = gaussians(s, o);
        Image l(g.w, g.h);
        const double sr = gaussians.scale_relative_to_octave(s);
        const float norm = static_cast<float>(sr * sr);
        for (int y = 0; y < g.h; ++y)
          for (int x = 0; x < g.w; ++x)
          {
            const float c = g(x, y);
            float value = 0.f;
            if (x == 0)
              value += g(x + 1, y) + c;
            else if (x == g.w - 1)
              value += c + g(x - 1, y);
            else
              value += g(x + 1, y) + g(x - 1, y);
            if (y == 0)
              value += g(x, y + 1) + c;
            else if (y == g.h - 1)
              value += c + g(x, y - 1);
            else
              value += g(x, y + 1) + g(x, y - 1);
            float out = value - 4 * c;  // 2 * N * (*in)
            out *= norm;
            l(x, y) = out;
          }
        L(s, o) = std::move(l);
      }
    }
    return L;
  }

  // FeatureDetectors/Hessian.hpp:35-57 det_of_hessian_pyramid: Hessian functor
  // (Differential.hpp:191-226), 2 x 2 determinant, times float(quartic(scale)) (quartic in double).
  Pyramid det_of_hessian_pyramid(const Pyramid& gaussians)
  {
    Pyramid D;
    D.reset(gaussians.num_octaves, gaussians.num_scales, gaussians.scale_initial,
            gaussians.scale_geometric_factor);
    for (int o = 0; o < D.num_octaves; ++o)
    {
      D.oct_scaling[o] = gaussians.oct_scaling[o];
      for (int s = 0; s < D.num_scales; ++s)
      {
        const Image& g = gaussians(s, o);
        Image d(g.w, g.h);
        const double sr = gaussians.scale_relative_to_octave(s);
        const float norm = static_cast<float>(sr * sr * sr * sr);
        for (int y = 0; y < g.h; ++y)
          for (int x = 0; x < g.w; ++x)
          {
            float H[4];
            hessian2(g, x, y, H);
            float det = H[0] * H[3] - H[2] * H[1];  // Eigen 2 x 2 determinant
            det *= norm;
            d(x, y) = det;
          }
        D(s, o) = std::move(d);
      }
    }
    return D;
  }

  // FeatureDetectors/LoG.cpp:20-58 ComputeLoGExtrema::operator() (which = 1) and
  // FeatureDetectors/Hessian.cpp:59-98 ComputeDoHExtrema::operator() (which = 2): the function
  // pyramid has as many layers as the Gaussian one, extrema are searched on s = 1 .. N - 2.
  void compute_function_extrema(const Image& image, const PyramidParams& pp, int which,
                                float extremum_thres, float edge_ratio_thres,
                                int img_padding_sz, int extremum_refinement_iter, Result& R)
  {
    R.G = gaussian_pyramid(image, pp, 4.f);
    R.D = which == 1 ? laplacian_pyramid(R.G) : det_of_hessian_pyramid(R.G);
    R.extrema.clear();
    for (int o = 0; o < R.D.num_octaves; ++o)
      for (int s = 1; s < R.D.num_scales - 1; ++s)
      {
        auto e = local_scale_space_extrema(R.D, s, o, extremum_thres, edge_ratio_thres,
                                           img_padding_sz, extremum_refinement_iter);
        R.extrema.insert(R.extrema.end(), e.begin(), e.end());
      }
  }

  // ------------------------------------------------------------------------ //
  // Hessian-Laplace.
  // FeatureDetectors/RefineExtremum.cpp:132-221 refine_extremum, 2-D version.  Returns false when the
  // offset is too large; pos / val are only written as the reference writes them.
  static bool refine_extremum_2d(const Image& I, int x, int y, int type, float pos[2], float& val,
                                 int border_sz, int num_iter)
  {
    float g[2] = {0.f, 0.f};
    float H[4] = {0.f, 0.f, 0.f, 0.f};
    float h[2] = {0.f, 0.f};
    pos[0] = static_cast<float>(x);
    pos[1] = static_cast<float>(y);
    for (int i = 0; i < num_iter; ++i)
    {
      if (x < border_sz || x >= I.w - border_sz || y < border_sz || y >= I.h - border_sz)
        break;
      // Differential.hpp:46-61 Gradient functor (borders replicated)
      {
        const float c = I(x, y);
        g[0] = x == 0 ? (I(x + 1, y) - c) / 2 : x == I.w - 1 ? (c - I(x - 1, y)) / 2 : (I(x + 1, y) - I(x - 1, y)) / 2;
        g[1] = y == 0 ? (I(x, y + 1) - c) / 2 : y == I.h - 1 ? (c - I(x, y - 1)) / 2 : (I(x, y + 1) - I(x, y - 1)) / 2;
      }
      hessian2(I, x, y, H);
      const float det = H[0] * H[3] - H[2] * H[1];
      const float tr = H[0] + H[3];
      if (det <= 0.f || tr * type >= 0.f)
      {
        g[0] = g[1] = 0.f;
        break;
      }
      // Eigen 2 x 2 inverse: adjugate times 1 / det; h = (-inverse) * g
      const float invdet = 1.f / det;
      const float i00 = H[3] * invdet, i01 = -H[2] * invdet, i10 = -H[1] * invdet, i11 = H[0] * invdet;
      h[0] = (-i00) * g[0] + (-i01) * g[1];
      h[1] = (-i10) * g[0] + (-i11) * g[1];
      if (std::max(std::fabs(h[0]), std::fabs(h[1])) > 1.5f)
        return false;
      if (std::min(std::fabs(h[0]), std::fabs(h[1])) > 0.6f)
      {
        x += h[0] > 0 ? 1 : -1;
        y += h[1] > 0 ? 1 : -1;
        continue;
      }
      break;
    }
    pos[0] = static_cast<float>(x);
    pos[1] = static_cast<float>(y);
    const float oldval = I(x, y);
    const float newval = oldval + 0.5f * (g[0] * h[0] + g[1] * h[1]);
    if ((type == 1 && oldval <= newval) || (type == -1 && oldval >= newval))
    {
      pos[0] += h[0];
      pos[1] += h[1];
      val = newval;
    }
    return true;
  }

  // The per-layer constants of select_laplace_scale (RefineExtremum.cpp:523-657): they depend on the
  // layer s only, never on the candidate.
  struct LaplaceScales
  {
    std::vector<float> scales;      // num_scales + 1
    std::vector<float> inc_sigma;   // blur that takes patch i - 1 to patch i; [0]: from the nearest Gaussian (<= 0: none)
    float ratio = 0.f;
  };
  static LaplaceScales laplace_scales(const Pyramid& G, int s, int num_scales)
  {
    LaplaceScales L;
    L.scales.resize(num_scales + 1);
    L.inc_sigma.resize(num_scales + 1);
    L.ratio = std::pow(2.f, 1.f / num_scales);
    const double nearest_sigma = G.scale_relative_to_octave(s - 1);
    L.scales[0] = static_cast<float>(G.scale_relative_to_octave(s)) / std::sqrt(2.f);
    const double inc0 = std::sqrt(static_cast<double>(L.scales[0] * L.scales[0]) - nearest_sigma * nearest_sigma);
    L.inc_sigma[0] = inc0 > 1e-3f ? static_cast<float>(inc0) : 0.f;  // NaN compares false: no blur
    for (int i = 1; i <= num_scales; ++i)
    {
      L.scales[i] = L.ratio * L.scales[i - 1];
      L.inc_sigma[i] = std::sqrt(L.scales[i] * L.scales[i] - L.scales[i - 1] * L.scales[i - 1]);
    }
    return L;
  }

  static bool select_laplace_scale(float& scale, int x, int y, int s, int o, const Pyramid& G, int num_scales,
                                   const LaplaceScales& L)
  {
    const Image& nearest = G(s - 1, o);
    const int patch_radius = static_cast<int>(std::ceil(std::sqrt(2.f) * 4.f));
    if (x - patch_radius < 0 || x + patch_radius >= nearest.w || y - patch_radius < 0 || y + patch_radius >= nearest.h)
      return false;
    const int side = 2 * patch_radius + 1;
    Image patch(side, side);
    for (int v = 0; v < side; ++v)
      for (int u = 0; u < side; ++u)
        patch(u, v) = nearest(x - patch_radius + u, y - patch_radius + v);
    std::vector<float> LoGs(num_scales + 1);
    auto log_at_centre = [&](const Image& p, float sc) {
      const int c = patch_radius;
      float value = 0.f;
      value += p(c + 1, c) + p(c - 1, c);
      value += p(c, c + 1) + p(c, c - 1);
      const float lap = value - 4 * p(c, c);
      return lap * (sc * sc);
    };
    if (L.inc_sigma[0] > 0.f)
      patch = gaussian(patch, L.inc_sigma[0]);
    LoGs[0] = log_at_centre(patch, L.scales[0]);
    for (int i = 1; i <= num_scales; ++i)
    {
      patch = gaussian(patch, L.inc_sigma[i]);
      LoGs[i] = log_at_centre(patch, L.scales[i]);
    }
    bool is_extremum = false;
    int i = 1;
    for (; i < num_scales; ++i)
    {
      is_extremum = (LoGs[i] <= LoGs[i - 1] && LoGs[i] <= LoGs[i + 1]) || (LoGs[i] >= LoGs[i - 1] && LoGs[i] >= LoGs[i + 1]);
      if (is_extremum)
        break;
    }
    if (is_extremum)
    {
      const float fprime = (LoGs[i + 1] - LoGs[i - 1]) / 2.f;
      const float fsecond = LoGs[i - 1] - 2.f * LoGs[i] + LoGs[i + 1];
      const float hh = -fprime / fsecond;
      scale = L.scales[i] * std::pow(L.ratio, hh);
    }
    return is_extremum;
  }

  // RefineExtremum.cpp:659-709 laplace_maxima.
  static std::vector<Keypoint> laplace_maxima(const Pyramid& F, const Pyramid& G, int s, int o, float extremum_thres,
                                              int img_padding_sz, int num_scales, int refine_iterations)
  {
    std::vector<Keypoint> corners;
    const Image& I = F(s, o);
    const LaplaceScales L = laplace_scales(G, s, num_scales);
    for (int y = img_padding_sz; y < I.h - img_padding_sz; ++y)
      for (int x = img_padding_sz; x < I.w - img_padding_sz; ++x)
      {
        if (!compare_with_neighborhood3<std::greater_equal<float>>(I(x, y), x, y, I, false))
          continue;
        if (I(x, y) < extremum_thres)
          continue;
        float scale = static_cast<float>(F.scale_relative_to_octave(s));
        if (!select_laplace_scale(scale, x, y, s, o, G, num_scales, L))
          continue;
        float val = I(x, y);
        float p[2];
        refine_extremum_2d(I, x, y, 1, p, val, img_padding_sz, refine_iterations);
        Keypoint c = make_oeregion(p[0], p[1], scale);
        c.orientation = 0.f;
        c.extremum_type = 1;
        c.extremum_value = val;
        c.s = s;
        c.o = o;
        c.xi = x;
        c.yi = y;
        corners.push_back(c);
      }
    return corners;
  }

  // FeatureDetectors/Hessian.cpp:19-57 ComputeHessianLaplaceMaxima::operator(): s = 1 .. N - 1.
  void compute_hessian_laplace(const Image& image, const PyramidParams& pp, float extremum_thres,
                               int img_padding_sz, int num_scales, int extremum_refinement_iter, Result& R)
  {
    R.G = gaussian_pyramid(image, pp, 4.f);
    R.D = det_of_hessian_pyramid(R.G);
    R.extrema.clear();
    for (int o = 0; o < R.D.num_octaves; ++o)
      for (int s = 1; s < R.D.num_scales; ++s)
      {
        auto e = laplace_maxima(R.D, R.G, s, o, extremum_thres, img_padding_sz, num_scales,
                                extremum_refinement_iter);
        R.extrema.insert(R.extrema.end(), e.begin(), e.end());
      }
  }

  // FeatureDetectors/Harris.cpp:165-230 ComputeHarrisLaplaceCorners::operator(): per layer
  // Gradient (Differential.hpp:46-61) -> SecondMomentMatrix (g g^T, SecondMomentMatrix.hpp:31-56) ->
  // Gaussian(sigma_I) on every coefficient -> det - kappa * pow(trace, 2) (pow(float, int) is a double
  // expression, narrowed when stored) -> times float(sigma_D * sigma_D); then laplace_maxima, s = 1 .. N - 1.
  Pyramid harris_cornerness_from_gaussians(const Pyramid& G, float kappa)
  {
    Pyramid C;
    C.reset(G.num_octaves, G.num_scales, G.scale_initial, G.scale_geometric_factor);
    static const float scale_factor = 1 / std::sqrt(2.f);
    for (int o = 0; o < G.num_octaves; ++o)
    {
      C.oct_scaling[o] = G.oct_scaling[o];
      for (int s = 0; s < G.num_scales; ++s)
      {
        const Image& g = G(s, o);
        const float sigma_I = static_cast<float>(G.scale_relative_to_octave(s));
        const float sigma_D = sigma_I * scale_factor;
        Image mxx(g.w, g.h), mxy(g.w, g.h), myy(g.w, g.h);
        for (int y = 0; y < g.h; ++y)
          for (int x = 0; x < g.w; ++x)
          {
            const float c = g(x, y);
            const float gx = x == 0 ? (g(x + 1, y) - c) / 2 : x == g.w - 1 ? (c - g(x - 1, y)) / 2 : (g(x + 1, y) - g(x - 1, y)) / 2;
            const float gy = y == 0 ? (g(x, y + 1) - c) / 2 : y == g.h - 1 ? (c - g(x, y - 1)) / 2 : (g(x, y + 1) - g(x, y - 1)) / 2;
            mxx(x, y) = gx * gx;
            mxy(x, y) = gx * gy;
            myy(x, y) = gy * gy;
          }
        const Image sxx = gaussian(mxx, sigma_I), sxy = gaussian(mxy, sigma_I), syy = gaussian(myy, sigma_I);
        Image c(g.w, g.h);
        const float norm = static_cast<float>(sigma_D * sigma_D);
        for (size_t i = 0; i < c.data.size(); ++i)
        {
          const float det = sxx.data[i] * syy.data[i] - sxy.data[i] * sxy.data[i];
          const float tr = sxx.data[i] + syy.data[i];
          float v = static_cast<float>(det - kappa * std::pow(static_cast<double>(tr), 2));
          v *= norm;
          c.data[i] = v;
        }
        C(s, o) = std::move(c);
      }
    }
    return C;
  }

  void compute_harris_laplace(const Image& image, const PyramidParams& pp, float kappa, float extremum_thres,
                              int img_padding_sz, int num_scales, int extremum_refinement_iter, Result& R)
  {
    R.G = gaussian_pyramid(image, pp, 4.f);
    R.D = harris_cornerness_from_gaussians(R.G, kappa);
    R.extrema.clear();
    for (int o = 0; o < R.D.num_octaves; ++o)
      for (int s = 1; s < R.D.num_scales; ++s)
      {
        auto e = laplace_maxima(R.D, R.G, s, o, extremum_thres, img_padding_sz, num_scales,
                                extremum_refinement_iter);
        R.extrema.insert(R.extrema.end(), e.begin(), e.end());
      }
  }

  // FeatureDetectors/SIFT.cpp:27-108 compute_sift_keypoints.
  void compute_sift_keypoints(const Image& image, const PyramidParams& pp,
                              float gauss_truncate, float extremum_thres,
                              float edge_ratio_thres,
                              int extremum_refinement_iter, bool parallel,
                              Result& R)
  {
    double t0 = now_ms();
    // N1: the 5th constructor slot of ComputeDoGExtrema is img_padding_sz;
    // the refinement iteration count keeps its default, 5.
    compute_dog_extrema(image, pp, gauss_truncate, extremum_thres,
                        edge_ratio_thres,
                        /* img_padding_sz */ extremum_refinement_iter,
                        /* extremum_refinement_iter */ 5, R);
    double t1 = now_ms();
    R.ms_dog = t1 - t0;

    const PolarPyramid nabla_G =
        gradient_polar_coordinates(R.G, g_threads_pyramid);
    double t2 = now_ms();
    R.ms_grad = t2 - t1;

    // Orientation.cpp:133-166 (serial in the reference; threaded only in the
    // "all cores" baseline mode, order preserved).
    const size_t n = R.extrema.size();
    std::vector<int> counts(n, 0);
    std::vector<float> oris(n * 36);
    const int th_ori = g_threads_pyramid;
#pragma omp parallel for num_threads(th_ori) if (th_ori != 1) schedule(dynamic, 64)
    for (size_t i = 0; i < n; ++i)
    {
      const Keypoint& e = R.extrema[i];
      const float s = static_cast<float>(R.G.scale_relative_to_octave(e.s));
      counts[i] = dominant_orientations(
          nabla_G.data[e.o][e.s].data(), nabla_G.w[e.o][e.s],
          nabla_G.h[e.o][e.s], e.x, e.y, s, 0.8f, 3.f, 1.5f, &oris[i * 36]);
    }
    R.oriented.clear();
    for (size_t i = 0; i < n; ++i)
      for (int j = 0; j < counts[i]; ++j)
      {
        R.oriented.push_back(R.extrema[i]);
        R.oriented.back().orientation = oris[i * 36 + j];
      }
    double t3 = now_ms();
    R.ms_ori = t3 - t2;

    const int m = static_cast<int>(R.oriented.size());
    R.descriptors.assign(static_cast<size_t>(m) * 128, 0.f);
    const int th = g_threads_other;
#pragma omp parallel for num_threads(th > 0 ? th : omp_get_max_threads()) if (parallel) schedule(dynamic, 64)
    for (int i = 0; i < m; ++i)
    {
      const Keypoint& f = R.oriented[i];
      sift_descriptor(f.x, f.y, oeregion_scale(f), f.orientation,
                      nabla_G.data[f.o][f.s].data(), nabla_G.w[f.o][f.s],
                      nabla_G.h[f.o][f.s], 3.f, 0.2f, true,
                      &R.descriptors[static_cast<size_t>(i) * 128]);
    }

    // SIFT.cpp:92-98 rescale to image coordinates.
    R.keypoints = R.oriented;
    for (int i = 0; i < m; ++i)
    {
      Keypoint& f = R.keypoints[i];
      const float z = R.G.oct_scaling[f.o];
      f.x *= z;
      f.y *= z;
      const float z2 = z * z;
      for (int j = 0; j < 4; ++j)
        f.shape[j] /= z2;
    }
    double t4 = now_ms();
    R.ms_desc = t4 - t3;
  }

}  // namespace oracle


// ========================================================================== //
// C interface for ctypes (tests, smoke, bench cpu_baseline only).
// ========================================================================== //
using namespace oracle;

static Image make_image(const float* p, int w, int h)
{
  Image I(w, h);
  std::memcpy(I.data.data(), p, sizeof(float) * static_cast<size_t>(w) * h);
  return I;
}

static PyramidParams make_params(int fo, int ns, float k, int pad, float cam,
                                 float init, int omax)
{
  PyramidParams p;
  p.first_octave_index = fo;
  p.scale_count_per_octave = ns;
  p.scale_geometric_factor = k;
  p.image_padding_size = pad;
  p.scale_camera = cam;
  p.scale_initial = init;
  p.num_octaves_max = omax;
  return p;
}

static thread_local char g_err[256];

#define ORACLE_TRY try {
#define ORACLE_CATCH                                                           \
  }                                                                            \
  catch (const std::exception& e)                                              \
  {                                                                            \
    std::strncpy(g_err, e.what(), sizeof(g_err) - 1);                          \
    return -1;                                                                 \
  }

extern "C" {

const char* oracle_last_error()
{
  return g_err;
}

// mode 0: reference-faithful threading (pyramid, polar gradient, orientation
// serial; extrema passes and -- if `parallel` -- descriptors OpenMP).
// mode 1: every stage OpenMP over `threads` cores (0 = all).
void oracle_set_threading(int mode, int threads)
{
#ifdef _OPENMP
  const int all = threads > 0 ? threads : omp_get_max_threads();
#else
  const int all = 1;
#endif
  g_threads_other = all;
  g_threads_pyramid = mode == 0 ? 1 : all;
}

// Threads the OpenMP stages actually use (the explicit count of oracle_set_threading,
// which overrides OMP_NUM_THREADS; else the OpenMP default).
int oracle_num_threads()
{
#ifdef _OPENMP
  return g_threads_other > 0 ? g_threads_other : omp_get_max_threads();
#else
  return 1;
#endif
}

int oracle_make_gaussian_kernel(float sigma, float gauss_truncate, float* out,
                                int capacity)
{
  const auto k = make_gaussian_kernel(sigma, gauss_truncate);
  if (static_cast<int>(k.size()) > capacity)
    return -static_cast<int>(k.size());
  std::memcpy(out, k.data(), sizeof(float) * k.size());
  return static_cast<int>(k.size());
}

void oracle_convolve_array(float* signal, const float* kernel, int signal_size,
                           int kernel_size)
{
  convolve_array(signal, kernel, signal_size, kernel_size);
}

int oracle_row_filter(const float* src, int w, int h, const float* kernel,
                      int ksz, float* dst)
{
  ORACLE_TRY
  Image s = make_image(src, w, h), d(w, h);
  apply_row_based_filter(s, d, kernel, ksz, 1);
  std::memcpy(dst, d.data.data(), sizeof(float) * d.data.size());
  return 0;
  ORACLE_CATCH
}

int oracle_column_filter(const float* src, int w, int h, const float* kernel,
                         int ksz, float* dst)
{
  ORACLE_TRY
  Image s = make_image(src, w, h), d(w, h);
  apply_column_based_filter(s, d, kernel, ksz, 1);
  std::memcpy(dst, d.data.data(), sizeof(float) * d.data.size());
  return 0;
  ORACLE_CATCH
}

int oracle_gaussian(const float* src, int w, int h, float sigma, float truncate,
                    float* dst)
{
  ORACLE_TRY
  Image s = make_image(src, w, h);
  Image d = gaussian(s, sigma, truncate);
  std::memcpy(dst, d.data.data(), sizeof(float) * d.data.size());
  return 0;
  ORACLE_CATCH
}

int oracle_interpolate(const float* src, int w, int h, double x, double y,
                       double* out)
{
  ORACLE_TRY
  Image s = make_image(src, w, h);
  *out = interpolate(s, x, y);
  return 0;
  ORACLE_CATCH
}

int oracle_enlarge(const float* src, int w, int h, float* dst, int dw, int dh)
{
  ORACLE_TRY
  Image s = make_image(src, w, h), d(dw, dh);
  enlarge(s, d, 1);
  std::memcpy(dst, d.data.data(), sizeof(float) * d.data.size());
  return 0;
  ORACLE_CATCH
}

int oracle_downscale(const float* src, int w, int h, int fact, float* dst)
{
  ORACLE_TRY
  Image s = make_image(src, w, h);
  Image d = downscale(s, fact);
  std::memcpy(dst, d.data.data(), sizeof(float) * d.data.size());
  return 0;
  ORACLE_CATCH
}

// Extremum predicates on a 3-layer stack (layers contiguous, each w*h).
// kind: 0 = local max (>=), 1 = local min (<=), 2 = strict max, 3 = strict min.
int oracle_local_scale_space_extremum(const float* stack3, int w, int h, int x,
                                      int y, int kind)
{
  Pyramid P;
  P.reset(1, 3, 1.6f, 1.26f);
  for (int s = 0; s < 3; ++s)
    P(s, 0) = make_image(stack3 + static_cast<size_t>(s) * w * h, w, h);
  switch (kind)
  {
  case 0:
    return local_scale_space_extremum<std::greater_equal<float>>(x, y, 1, 0, P);
  case 1:
    return local_scale_space_extremum<std::less_equal<float>>(x, y, 1, 0, P);
  case 2:
    return local_scale_space_extremum<std::greater<float>>(x, y, 1, 0, P);
  default:
    return local_scale_space_extremum<std::less<float>>(x, y, 1, 0, P);
  }
}

int oracle_local_extremum(const float* img, int w, int h, int x, int y, int kind)
{
  Image I = make_image(img, w, h);
  const float v = I(x, y);
  switch (kind)
  {
  case 0:
    return compare_with_neighborhood3<std::greater_equal<float>>(v, x, y, I, false);
  case 1:
    return compare_with_neighborhood3<std::less_equal<float>>(v, x, y, I, false);
  case 2:
    return compare_with_neighborhood3<std::greater<float>>(v, x, y, I, false);
  default:
    return compare_with_neighborhood3<std::less<float>>(v, x, y, I, false);
  }
}

int oracle_on_edge(const float* img, int w, int h, int x, int y, float ratio)
{
  Image I = make_image(img, w, h);
  return on_edge(I, x, y, ratio) ? 1 : 0;
}

void oracle_hessian2(const float* img, int w, int h, int x, int y, float* H)
{
  Image I = make_image(img, w, h);
  hessian2(I, x, y, H);
}

void oracle_sym3_eigenvalues(const float* H, float* lambda)
{
  sym3_eigenvalues(H, lambda);
}

void oracle_inverse3(const float* m, float* r)
{
  inverse3(m, r);
}

void oracle_gradient_polar(const float* img, int w, int h, float* out)
{
  Image I = make_image(img, w, h);
  std::vector<float> g;
  gradient_polar_coordinates(I, g, 1);
  std::memcpy(out, g.data(), sizeof(float) * g.size());
}

void oracle_orientation_histogram(const float* grad, int w, int h, float x,
                                  float y, float s, float trunc, float blur,
                                  float* hist36)
{
  compute_orientation_histogram(hist36, grad, w, h, x, y, s, trunc, blur);
}

void oracle_lowe_smooth_histogram(float* hist36, int iters)
{
  lowe_smooth_histogram(hist36, iters);
}

int oracle_find_peaks(const float* hist36, float ratio, int* peaks)
{
  return find_peaks(hist36, ratio, peaks);
}

float oracle_refine_peak(const float* hist36, int i)
{
  return refine_peak(hist36, i);
}

int oracle_dominant_orientations(const float* grad, int w, int h, float x,
                                 float y, float sigma, float* out36)
{
  return dominant_orientations(grad, w, h, x, y, sigma, 0.8f, 3.f, 1.5f, out36);
}

void oracle_sift_descriptor(const float* grad, int w, int h, float x, float y,
                            float s, float theta, int normalize, float* desc128)
{
  sift_descriptor(x, y, s, theta, grad, w, h, 3.f, 0.2f, normalize != 0, desc128);
}

// from_rgb8_to_gray32f, non-Halide branch (ImageProcessing/FastColorConversion.cpp:42-67):
// DO::Sara::convert(ImageView<Rgb8>, ImageView<float>) = per pixel smart_convert_color
// (Core/Pixel/SmartColorConversion.hpp:237-246): each channel to double with
// to_normalized_float_channel<uint8_t, double> (Core/Pixel/ChannelConversion.hpp:41-53),
// rgb_to_gray<double> (Core/Pixel/ColorConversion.hpp:27-33), then narrowed to float
// (ChannelConversion.hpp:79-82).
void oracle_rgb8_to_gray32f(const unsigned char* src, int n_pixels, float* dst)
{
  for (int i = 0; i < n_pixels; ++i)
  {
    double rgb[3];
    for (int c = 0; c < 3; ++c)
    {
      const double float_min = 0.0, float_max = 255.0;
      const double float_range = float_max - float_min;
      rgb[c] = (static_cast<double>(src[3 * i + c]) - float_min) / float_range;
    }
    const double gray = 0.2125 * rgb[0] + 0.7154 * rgb[1] + 0.0721 * rgb[2];
    dst[i] = static_cast<float>(gray);
  }
}

// ImageView<uint8_t> -> float: convert_channel(Int, float&) = to_normalized_float_channel
// <uint8_t, float> (Core/Pixel/ChannelConversion.hpp:41-53, 95-99).
void oracle_gray8_to_gray32f(const unsigned char* src, int n_pixels, float* dst)
{
  for (int i = 0; i < n_pixels; ++i)
  {
    const float float_min = 0.f, float_max = 255.f;
    const float float_range = float_max - float_min;
    dst[i] = (static_cast<float>(src[i]) - float_min) / float_range;
  }
}

// Refinement trace (see g_trace): switch on, run oracle_dog_extrema / oracle_sift, read.
void oracle_trace_refinement(int on)
{
  g_trace_on = on != 0;
  g_trace.clear();
}
int oracle_trace_size()
{
  return static_cast<int>(g_trace.size() / 25);
}
void oracle_trace_copy(float* dst)
{
  std::memcpy(dst, g_trace.data(), g_trace.size() * sizeof(float));
}

// ---- whole pipeline ------------------------------------------------------ //
struct oracle_result;

int oracle_sift(const float* image, int w, int h, int fo, int ns, float k,
                int pad, float cam, float init, int omax, float gauss_truncate,
                float extremum_thres, float edge_ratio_thres,
                int extremum_refinement_iter, int parallel, void** out)
{
  ORACLE_TRY
  auto* R = new Result;
  try
  {
    compute_sift_keypoints(make_image(image, w, h),
                           make_params(fo, ns, k, pad, cam, init, omax),
                           gauss_truncate, extremum_thres, edge_ratio_thres,
                           extremum_refinement_iter, parallel != 0, *R);
  }
  catch (...)
  {
    delete R;
    throw;
  }
  *out = R;
  return 0;
  ORACLE_CATCH
}

// ComputeDoGExtrema with explicit padding / iteration slots (DoG.hpp:72-78);
// pyramids and extrema only.
int oracle_dog_extrema(const float* image, int w, int h, int fo, int ns, float k,
                       int pad, float cam, float init, int omax,
                       float gauss_truncate, float extremum_thres,
                       float edge_ratio_thres, int img_padding_sz,
                       int refine_iter, void** out)
{
  ORACLE_TRY
  auto* R = new Result;
  try
  {
    compute_dog_extrema(make_image(image, w, h),
                        make_params(fo, ns, k, pad, cam, init, omax),
                        gauss_truncate, extremum_thres, edge_ratio_thres,
                        img_padding_sz, refine_iter, *R);
  }
  catch (...)
  {
    delete R;
    throw;
  }
  *out = R;
  return 0;
  ORACLE_CATCH
}

// ComputeLoGExtrema (which = 1) / ComputeDoHExtrema (which = 2); the function pyramid is
// read back as layer kind 1 (in place of the DoG).
int oracle_function_extrema(const float* image, int w, int h, int fo, int ns, float k,
                            int pad, float cam, float init, int omax, int which,
                            float extremum_thres, float edge_ratio_thres,
                            int img_padding_sz, int refine_iter, void** out)
{
  ORACLE_TRY
  auto* R = new Result;
  try
  {
    compute_function_extrema(make_image(image, w, h), make_params(fo, ns, k, pad, cam, init, omax),
                             which, extremum_thres, edge_ratio_thres, img_padding_sz, refine_iter, *R);
  }
  catch (...)
  {
    delete R;
    throw;
  }
  *out = R;
  return 0;
  ORACLE_CATCH
}

// ComputeHessianLaplaceMaxima (FeatureDetectors/Hessian.hpp:84-94 defaults: ImagePyramidParams(-1, 3 + 1),
// 1e-5, padding 1, 10 scales, 5 iterations); the det-of-Hessian pyramid is layer kind 1.
int oracle_hessian_laplace(const float* image, int w, int h, int fo, int ns, float k,
                           int pad, float cam, float init, int omax, float extremum_thres,
                           int img_padding_sz, int num_scales, int refine_iter, void** out)
{
  ORACLE_TRY
  auto* R = new Result;
  try
  {
    compute_hessian_laplace(make_image(image, w, h), make_params(fo, ns, k, pad, cam, init, omax),
                            extremum_thres, img_padding_sz, num_scales, refine_iter, *R);
  }
  catch (...)
  {
    delete R;
    throw;
  }
  *out = R;
  return 0;
  ORACLE_CATCH
}

// The constants of select_laplace_scale for layer s: scales[num_scales + 1], inc_sigma[num_scales + 1]
// (0: no blur), ratio.  Shared with the GPU library's caller so that both sides use the same bits.
void oracle_laplace_scales(float k, float scale_initial, int s, int num_scales, float* scales, float* inc_sigma,
                           float* ratio)
{
  Pyramid G;
  G.reset(1, s + 1, scale_initial, k);
  const LaplaceScales L = laplace_scales(G, s, num_scales);
  for (int i = 0; i <= num_scales; ++i)
  {
    scales[i] = L.scales[i];
    inc_sigma[i] = L.inc_sigma[i];
  }
  *ratio = L.ratio;
}

// ComputeHarrisLaplaceCorners (FeatureDetectors/Harris.hpp:125-138 defaults: ImagePyramidParams(-1, 2 + 1,
// sqrt(2), 1), kappa 0.04, 1e-6, padding 1, 10 scales, 5 iterations); the cornerness pyramid is layer kind 1.
int oracle_harris_laplace(const float* image, int w, int h, int fo, int ns, float k,
                          int pad, float cam, float init, int omax, float kappa, float extremum_thres,
                          int img_padding_sz, int num_scales, int refine_iter, void** out)
{
  ORACLE_TRY
  auto* R = new Result;
  try
  {
    compute_harris_laplace(make_image(image, w, h), make_params(fo, ns, k, pad, cam, init, omax), kappa,
                           extremum_thres, img_padding_sz, num_scales, refine_iter, *R);
  }
  catch (...)
  {
    delete R;
    throw;
  }
  *out = R;
  return 0;
  ORACLE_CATCH
}

void oracle_free(void* r)
{
  delete static_cast<Result*>(r);
}

int oracle_num_octaves(void* r)
{
  return static_cast<Result*>(r)->G.num_octaves;
}
int oracle_num_scales(void* r)
{
  return static_cast<Result*>(r)->G.num_scales;
}
float oracle_octave_scaling(void* r, int o)
{
  return static_cast<Result*>(r)->G.oct_scaling[o];
}
void oracle_layer_size(void* r, int o, int* w, int* h)
{
  const Image& I = static_cast<Result*>(r)->G(0, o);
  *w = I.w;
  *h = I.h;
}
// which: 0 = Gaussian, 1 = DoG.
void oracle_copy_layer(void* r, int which, int s, int o, float* dst)
{
  auto* R = static_cast<Result*>(r);
  const Image& I = which == 0 ? R->G(s, o) : R->D(s, o);
  std::memcpy(dst, I.data.data(), sizeof(float) * I.data.size());
}
int oracle_num_extrema(void* r)
{
  return static_cast<int>(static_cast<Result*>(r)->extrema.size());
}
int oracle_num_keypoints(void* r)
{
  return static_cast<int>(static_cast<Result*>(r)->keypoints.size());
}
int oracle_keypoint_stride()
{
  return static_cast<int>(sizeof(Keypoint));
}
// which: 0 = extrema (octave coords), 1 = oriented (octave coords),
// 2 = final keypoints (image coords).
void oracle_copy_keypoints(void* r, int which, void* dst)
{
  auto* R = static_cast<Result*>(r);
  const auto& v =
      which == 0 ? R->extrema : (which == 1 ? R->oriented : R->keypoints);
  std::memcpy(dst, v.data(), sizeof(Keypoint) * v.size());
}
void oracle_copy_descriptors(void* r, float* dst)
{
  auto* R = static_cast<Result*>(r);
  std::memcpy(dst, R->descriptors.data(), sizeof(float) * R->descriptors.size());
}
void oracle_stage_ms(void* r, double* ms4)
{
  auto* R = static_cast<Result*>(r);
  ms4[0] = R->ms_dog;
  ms4[1] = R->ms_grad;
  ms4[2] = R->ms_ori;
  ms4[3] = R->ms_desc;
}

}  // extern "C"
