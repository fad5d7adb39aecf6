// ========================================================================== //
// sara_b200.hpp -- header-only C++17 adapter over the C ABI (sara_b200.h).
//
// Keeps the C++ surface of the reference path so that a Sara call site switches
// by changing a namespace:
//
//   DO::Sara::compute_sift_keypoints(image, pyramid_params, gauss_truncate,
//       extremum_thres, edge_ratio_thres, extremum_refinement_iter, parallel)
//       -> KeypointList<OERegion, float>        FeatureDetectors/SIFT.hpp:24-33
//
// Without Eigen/Sara headers on the include path the types below are Eigen-free
// POD mirrors with the same member names and memory order:
//   ImageView<float>      Core/Image/Image.hpp:44-103   (data(), width(), height())
//   ImagePyramidParams    ImageProcessing/ImagePyramid.hpp:29-198
//   OERegion              Features/Feature.hpp:40-179
//   KeypointList<F, T>    Features/KeypointList.hpp:35-36  = tuple<vector<F>, Tensor_<T, 2>>
// With <DO/Sara/Features/KeypointList.hpp> reachable (a real Sara build), define
// SARA_B200_WITH_SARA before including this header and the overload taking and
// returning the REAL Sara types is compiled as well.
//
// Errors: the C ABI never throws; this adapter rethrows the std exception the
// reference would have thrown (std::runtime_error for < 4 scales, DoG.hpp:86-89;
// std::domain_error for bad sizes, LinearFiltering.hpp:82-84).
// ========================================================================== //
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "sara_b200.h"

namespace sara_b200 {

  // ---- ImageView<float> ---------------------------------------------------- //
  template <typename T>
  class ImageView;

  template <>
  class ImageView<float>
  {
  public:
    ImageView() = default;
    ImageView(const float* data, int width, int height)
      : _data{data}
      , _w{width}
      , _h{height}
    {
    }
    const float* data() const { return _data; }
    int width() const { return _w; }
    int height() const { return _h; }
    float operator()(int x, int y) const { return _data[static_cast<std::size_t>(y) * _w + x]; }

  private:
    const float* _data = nullptr;
    int _w = 0, _h = 0;
  };

  // ---- ImagePyramidParams ---------------------------------------------------- //
  class ImagePyramidParams
  {
  public:
    ImagePyramidParams(int first_octave_index = -1, int scale_count_per_octave = 3 + 3,
                       float scale_geometric_factor = std::pow(2.f, 1.f / 3.f),
                       int image_padding_size = 1, float scale_camera = 0.5f,
                       float scale_initial = 1.6f,
                       int num_octaves_max = std::numeric_limits<int>::max())
      : _p{first_octave_index, scale_count_per_octave, scale_geometric_factor, image_padding_size,
           scale_camera, scale_initial, num_octaves_max}
    {
    }
    int first_octave_index() const { return _p.first_octave_index; }
    int scale_count_per_octave() const { return _p.scale_count_per_octave; }
    float scale_geometric_factor() const { return _p.scale_geometric_factor; }
    int image_padding_size() const { return _p.image_padding_size; }
    float scale_camera() const { return _p.scale_camera; }
    float scale_initial() const { return _p.scale_initial; }
    int num_octaves_max() const { return _p.num_octaves_max; }
    const sara_b200_pyramid_params& c_params() const { return _p; }

  private:
    sara_b200_pyramid_params _p;
  };

  // ---- OERegion (POD mirror: same member names, same order) ---------------- //
  struct Point2f
  {
    float v[2] = {0.f, 0.f};
    float operator()(int i) const { return v[i]; }
    float& operator()(int i) { return v[i]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
  };

  struct Matrix2f  // column-major like Eigen's default
  {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    float operator()(int r, int c) const { return v[c * 2 + r]; }
    float& operator()(int r, int c) { return v[c * 2 + r]; }
  };

  struct OERegion
  {
    enum class Type : std::uint8_t
    {
      Harris, HarAff, HarLap, FAST, SUSAN, DoG, LoG, DoH, MSER, HesAff, HesLap, Undefined
    };
    enum class ExtremumType : std::int8_t
    {
      Min = -1, Saddle = 0, Max = 1, Undefined = -2
    };

    float x() const { return coords(0); }
    float y() const { return coords(1); }
    const Point2f& center() const { return coords; }
    // Feature.cpp:28-39 for the isotropic shape matrices this path produces.
    float radius(float = 0.f) const { return 1.f / std::sqrt(shape_matrix(0, 0)); }
    float scale(float a = 0.f) const { return radius(a); }

    Point2f coords;
    Matrix2f shape_matrix;
    float orientation = 0.f;
    float extremum_value = 0.f;
    Type type = Type::Undefined;
    ExtremumType extremum_type = ExtremumType::Undefined;
  };

  // ---- Tensor_<float, 2> (row-major N x 128) -------------------------------- //
  class DescriptorMatrix
  {
  public:
    DescriptorMatrix() = default;
    DescriptorMatrix(int rows, int cols)
      : _rows{rows}
      , _cols{cols}
      , _data(static_cast<std::size_t>(rows) * cols)
    {
    }
    int rows() const { return _rows; }
    int cols() const { return _cols; }
    int size(int i) const { return i == 0 ? _rows : _cols; }
    float* data() { return _data.data(); }
    const float* data() const { return _data.data(); }
    const float* operator[](int r) const { return _data.data() + static_cast<std::size_t>(r) * _cols; }
    float operator()(int r, int c) const { return (*this)[r][c]; }

  private:
    int _rows = 0, _cols = 0;
    std::vector<float> _data;
  };

  template <typename F, typename T>
  using KeypointList = std::tuple<std::vector<F>, DescriptorMatrix>;

  template <typename F, typename T>
  inline const std::vector<F>& features(const KeypointList<F, T>& keys)
  {
    return std::get<0>(keys);
  }
  template <typename F, typename T>
  inline const DescriptorMatrix& descriptors(const KeypointList<F, T>& keys)
  {
    return std::get<1>(keys);
  }

  // ---- error mapping ---------------------------------------------------------- //
  [[noreturn]] inline void rethrow(int rc, const sara_b200_ctx* ctx)
  {
    const std::string msg = sara_b200_last_error(ctx);
    switch (rc)
    {
    case SARA_B200_ERR_BAD_ARG:
      throw std::domain_error{msg};
    case SARA_B200_ERR_OVERFLOW:
      throw std::length_error{msg};
    case SARA_B200_ERR_OOM:
      throw std::bad_alloc{};
    default:
      throw std::runtime_error{msg};  // TOO_FEW_SCALES (DoG.hpp:86-89), CUDA, BUSY
    }
  }

  // ---- RAII context ------------------------------------------------------------ //
  class Context
  {
  public:
    Context(int max_width, int max_height, int device = 0, int max_keypoints = 262144,
            int num_slots = 1, int min_first_octave_index = -1)
    {
      sara_b200_limits lim{max_width, max_height, max_keypoints, num_slots, min_first_octave_index};
      sara_b200_ctx* c = nullptr;
      const int rc = sara_b200_create(device, &lim, &c);
      if (rc != 0)
        rethrow(rc, nullptr);
      _ctx.reset(c);
      _max_w = max_width;
      _max_h = max_height;
      _max_kp = max_keypoints;
    }
    sara_b200_ctx* get() const { return _ctx.get(); }
    bool fits(int w, int h) const { return w <= _max_w && h <= _max_h; }

    // The drop-in body of compute_sift_keypoints, filling any keypoint type whose
    // members are named like OERegion's.
    template <typename Region, typename MakeDescriptors>
    auto compute(const float* image, int w, int h, const sara_b200_sift_args& args,
                 MakeDescriptors&& make_descriptors)
    {
      int rc = sara_b200_sift_enqueue(get(), 0, image, w, h, 0, &args, nullptr);
      if (rc != 0)
        rethrow(rc, get());
      int n = 0;
      rc = sara_b200_wait(get(), 0, &n);
      if (rc != 0)
        rethrow(rc, get());
      std::vector<sara_b200_keypoint> raw(static_cast<std::size_t>(n > 0 ? n : 1));
      auto desc = make_descriptors(n, 128);
      rc = sara_b200_collect(get(), 0, raw.data(), desc.data(), n > 0 ? n : 1, &n);
      if (rc != 0)
        rethrow(rc, get());
      std::vector<Region> regions(static_cast<std::size_t>(n));
      for (int i = 0; i < n; ++i)
      {
        const sara_b200_keypoint& k = raw[i];
        Region& r = regions[i];
        r.coords(0) = k.x;
        r.coords(1) = k.y;
        r.shape_matrix(0, 0) = k.shape[0];
        r.shape_matrix(1, 0) = k.shape[1];
        r.shape_matrix(0, 1) = k.shape[2];
        r.shape_matrix(1, 1) = k.shape[3];
        r.orientation = k.orientation;
        r.extremum_value = k.extremum_value;
        r.type = static_cast<decltype(r.type)>(k.type);
        r.extremum_type = static_cast<decltype(r.extremum_type)>(k.extremum_type);
      }
      return std::make_tuple(std::move(regions), std::move(desc));
    }

  private:
    struct Deleter
    {
      void operator()(sara_b200_ctx* c) const { sara_b200_destroy(c); }
    };
    std::unique_ptr<sara_b200_ctx, Deleter> _ctx;
    int _max_w = 0, _max_h = 0, _max_kp = 0;
  };

  // Per-thread default context (the reference function is stateless; one ctx per
  // host thread keeps that contract, see the threading note in sara_b200.h).
  inline Context& default_context(int w, int h)
  {
    // Sized for the largest width and the largest height seen so far, so that frames of
    // alternating orientation do not recreate it on every call.
    thread_local std::unique_ptr<Context> ctx;
    thread_local int max_w = 0, max_h = 0;
    if (!ctx || !ctx->fits(w, h))
    {
      max_w = std::max(max_w, w);
      max_h = std::max(max_h, h);
      ctx.reset();
      ctx = std::make_unique<Context>(max_w, max_h);
    }
    return *ctx;
  }

  inline sara_b200_sift_args make_sift_args(const sara_b200_pyramid_params& pp, float gauss_truncate,
                                            float extremum_thres, float edge_ratio_thres,
                                            int extremum_refinement_iter)
  {
    sara_b200_sift_args a;
    a.pyramid_params = pp;
    a.gauss_truncate = gauss_truncate;
    a.extremum_thres = extremum_thres;
    a.edge_ratio_thres = edge_ratio_thres;
    a.extremum_refinement_iter = extremum_refinement_iter;
    return a;
  }

  //! Same signature and defaults as DO::Sara::compute_sift_keypoints
  //! (FeatureDetectors/SIFT.hpp:24-33).  `parallel` is accepted for source
  //! compatibility; the GPU path is always parallel.
  inline auto compute_sift_keypoints(const ImageView<float>& image,
                                     const ImagePyramidParams& pyramid_params = ImagePyramidParams(),
                                     float gauss_truncate = 4.f, float extremum_thres = 0.01f,
                                     float edge_ratio_thres = 10.f, int extremum_refinement_iter = 5,
                                     bool /*parallel*/ = false) -> KeypointList<OERegion, float>
  {
    const auto args = make_sift_args(pyramid_params.c_params(), gauss_truncate, extremum_thres,
                                     edge_ratio_thres, extremum_refinement_iter);
    return default_context(image.width(), image.height())
        .compute<OERegion>(image.data(), image.width(), image.height(), args,
                           [](int n, int d) { return DescriptorMatrix{n, d}; });
  }

  // ---- detector functors: ComputeDoGExtrema (FeatureDetectors/DoG.hpp:72-165), ComputeLoGExtrema
  // (LoG.hpp:71-117), ComputeDoHExtrema / ComputeHessianLaplaceMaxima (Hessian.hpp:60-240),
  // ComputeHarrisLaplaceCorners (Harris.hpp:95-175): same constructor arguments and defaults, same call
  // operator `std::vector<OERegion> operator()(const ImageView<float>&, std::vector<Point2i>* scale_octave_pairs)`.
  struct Point2i
  {
    int v[2] = {0, 0};
    int operator()(int i) const { return v[i]; }
    int& operator()(int i) { return v[i]; }
  };

  namespace detail {
    template <typename Call>
    inline std::vector<OERegion> run_detector(const ImageView<float>& I, std::vector<Point2i>* scale_octave_pairs,
                                              Call&& call)
    {
      Context& ctx = default_context(I.width(), I.height());
      int rc = call(ctx.get());
      if (rc != 0)
        rethrow(rc, ctx.get());
      int n = 0;
      rc = sara_b200_copy_extrema(ctx.get(), 0, nullptr, 0, &n);
      std::vector<sara_b200_keypoint> raw(static_cast<std::size_t>(n > 0 ? n : 1));
      if (n > 0)
      {
        rc = sara_b200_copy_extrema(ctx.get(), 0, raw.data(), n, &n);
        if (rc != 0)
          rethrow(rc, ctx.get());
      }
      std::vector<OERegion> out(static_cast<std::size_t>(n));
      if (scale_octave_pairs)
        scale_octave_pairs->assign(static_cast<std::size_t>(n), Point2i{});
      for (int i = 0; i < n; ++i)
      {
        const sara_b200_keypoint& k = raw[i];
        OERegion& r = out[i];
        r.coords(0) = k.x;
        r.coords(1) = k.y;
        r.shape_matrix(0, 0) = k.shape[0];
        r.shape_matrix(1, 0) = k.shape[1];
        r.shape_matrix(0, 1) = k.shape[2];
        r.shape_matrix(1, 1) = k.shape[3];
        r.orientation = k.orientation;
        r.extremum_value = k.extremum_value;
        r.type = static_cast<OERegion::Type>(k.type);
        r.extremum_type = static_cast<OERegion::ExtremumType>(k.extremum_type);
        if (scale_octave_pairs)
        {
          (*scale_octave_pairs)[i](0) = k.s;
          (*scale_octave_pairs)[i](1) = k.o;
        }
      }
      return out;
    }

    inline sara_b200_dog_args dog_args(const ImagePyramidParams& pp, float gauss_truncate, float thres, float edge_ratio,
                                       int padding, int iters)
    {
      sara_b200_dog_args a;
      a.pyramid_params = pp.c_params();
      a.gauss_truncate = gauss_truncate;
      a.extremum_thres = thres;
      a.edge_ratio_thres = edge_ratio;
      a.img_padding_sz = padding;
      a.extremum_refinement_iter = iters;
      return a;
    }
  }  // namespace detail

  class ComputeDoGExtrema
  {
  public:
    ComputeDoGExtrema(const ImagePyramidParams& pyramid_params = ImagePyramidParams(), float gauss_truncate = 4.f,
                      float extremum_thres = 0.01f, float edge_ratio_thres = 10.f, int img_padding_sz = 1,
                      int extremum_refinement_iter = 5)
      : _args{detail::dog_args(pyramid_params, gauss_truncate, extremum_thres, edge_ratio_thres, img_padding_sz,
                               extremum_refinement_iter)}
    {
      if (pyramid_params.scale_count_per_octave() < 4)  // DoG.hpp:86-89
        throw std::runtime_error{"Error: The extraction of DoG extrema needs (1 + 3) = 4 scales per octave at the very "
                                 "minimum!"};
    }
    std::vector<OERegion> operator()(const ImageView<float>& I, std::vector<Point2i>* scale_octave_pairs = nullptr)
    {
      return detail::run_detector(I, scale_octave_pairs, [&](sara_b200_ctx* c) {
        return sara_b200_dog_extrema(c, 0, I.data(), I.width(), I.height(), 0, &_args);
      });
    }

  private:
    sara_b200_dog_args _args;
  };

  class ComputeLoGExtrema
  {
  public:
    ComputeLoGExtrema(const ImagePyramidParams& pyr_params = ImagePyramidParams(-1, 3 + 2), float extremum_thres = 0.01f,
                      float edge_ratio_thres = 10.f, int img_padding_sz = 1, int extremum_refinement_iter = 5)
      : _args{detail::dog_args(pyr_params, 4.f, extremum_thres, edge_ratio_thres, img_padding_sz, extremum_refinement_iter)}
    {
    }
    std::vector<OERegion> operator()(const ImageView<float>& I, std::vector<Point2i>* scale_octave_pairs = nullptr)
    {
      return detail::run_detector(I, scale_octave_pairs, [&](sara_b200_ctx* c) {
        return sara_b200_log_extrema(c, 0, I.data(), I.width(), I.height(), 0, &_args);
      });
    }

  private:
    sara_b200_dog_args _args;
  };

  class ComputeDoHExtrema
  {
  public:
    ComputeDoHExtrema(const ImagePyramidParams& pyr_params = ImagePyramidParams(-1, 3 + 2, std::pow(2.f, 1.f / 3.f), 2),
                      float extremum_thres = 1e-6f, float edge_ratio_thres = 10.f, int img_padding_sz = 1,
                      int extremum_refinement_iter = 2)
      : _args{detail::dog_args(pyr_params, 4.f, extremum_thres, edge_ratio_thres, img_padding_sz, extremum_refinement_iter)}
    {
    }
    std::vector<OERegion> operator()(const ImageView<float>& I, std::vector<Point2i>* scale_octave_pairs = nullptr)
    {
      return detail::run_detector(I, scale_octave_pairs, [&](sara_b200_ctx* c) {
        return sara_b200_doh_extrema(c, 0, I.data(), I.width(), I.height(), 0, &_args);
      });
    }

  private:
    sara_b200_dog_args _args;
  };

  class ComputeHessianLaplaceMaxima
  {
  public:
    ComputeHessianLaplaceMaxima(const ImagePyramidParams& pyr_params = ImagePyramidParams(-1, 3 + 1),
                                float extremum_thres = 1e-5f, int img_padding_sz = 1, int num_scales = 10,
                                int extremum_refinement_iter = 5)
      : _args{detail::dog_args(pyr_params, 4.f, extremum_thres, 10.f, img_padding_sz, extremum_refinement_iter)}
      , _num_scales{num_scales}
    {
    }
    std::vector<OERegion> operator()(const ImageView<float>& I, std::vector<Point2i>* scale_octave_pairs = nullptr)
    {
      return detail::run_detector(I, scale_octave_pairs, [&](sara_b200_ctx* c) {
        return sara_b200_hessian_laplace(c, 0, I.data(), I.width(), I.height(), 0, &_args, _num_scales);
      });
    }

  private:
    sara_b200_dog_args _args;
    int _num_scales;
  };

  class ComputeHarrisLaplaceCorners
  {
  public:
    ComputeHarrisLaplaceCorners(const ImagePyramidParams& pyr_params = ImagePyramidParams(-1, 2 + 1, std::sqrt(2.f), 1),
                                float kappa = 0.04f, float extremum_thres = 1e-6f, int img_padding_sz = 1,
                                int scale_count = 10, int extremum_refinement_iter = 5)
      : _args{detail::dog_args(pyr_params, 4.f, extremum_thres, 10.f, img_padding_sz, extremum_refinement_iter)}
      , _kappa{kappa}
      , _scale_count{scale_count}
    {
    }
    std::vector<OERegion> operator()(const ImageView<float>& I, std::vector<Point2i>* scale_octave_pairs = nullptr)
    {
      return detail::run_detector(I, scale_octave_pairs, [&](sara_b200_ctx* c) {
        return sara_b200_harris_laplace(c, 0, I.data(), I.width(), I.height(), 0, &_args, _kappa, _scale_count);
      });
    }

  private:
    sara_b200_dog_args _args;
    float _kappa;
    int _scale_count;
  };

  // ---- Match + AnnMatcher (Match/Match.hpp:27-177, FeatureMatching/AnnMatcher.hpp:33-84) ---- //
  class Match
  {
  public:
    enum class Direction : std::uint8_t
    {
      SourceToTarget,
      TargetToSource
    };
    Match() = default;
    Match(const OERegion* x, const OERegion* y, float score = std::numeric_limits<float>::max(),
          Direction matching_dir = Direction::SourceToTarget, int x_index = -1, int y_index = -1)
      : _x{x}, _y{y}, _x_index{x_index}, _y_index{y_index}, _score{score}, _matching_dir{matching_dir}
    {
    }
    const OERegion* x_pointer() const { return _x; }
    const OERegion* y_pointer() const { return _y; }
    const OERegion& x() const
    {
      if (_x == nullptr)
        throw std::runtime_error{"x is null"};
      return *_x;
    }
    const OERegion& y() const
    {
      if (_y == nullptr)
        throw std::runtime_error{"y is null"};
      return *_y;
    }
    const Point2f& x_pos() const { return x().center(); }
    const Point2f& y_pos() const { return y().center(); }
    int rank() const { return _rank; }
    int& rank() { return _rank; }
    float score() const { return _score; }
    Direction matching_direction() const { return _matching_dir; }
    int x_index() const { return _x_index; }
    int y_index() const { return _y_index; }

  private:
    const OERegion* _x = nullptr;
    const OERegion* _y = nullptr;
    int _x_index = -1, _y_index = -1, _rank = -1;
    float _score = std::numeric_limits<float>::max();
    Direction _matching_dir = Direction::SourceToTarget;
  };

  inline sara_b200_keypoint to_c_keypoint(const OERegion& r)
  {
    sara_b200_keypoint k{};
    k.x = r.coords(0);
    k.y = r.coords(1);
    k.shape[0] = r.shape_matrix(0, 0);
    k.shape[1] = r.shape_matrix(1, 0);
    k.shape[2] = r.shape_matrix(0, 1);
    k.shape[3] = r.shape_matrix(1, 1);
    k.orientation = r.orientation;
    k.extremum_value = r.extremum_value;
    k.type = static_cast<std::uint8_t>(r.type);
    k.extremum_type = static_cast<std::int8_t>(r.extremum_type);
    return k;
  }

  //! Same constructors and compute_matches() as DO::Sara::AnnMatcher; the search underneath is the exact
  //! one on the GPU (sara_b200_compute_matches).  Like the reference, it keeps references to the key lists
  //! and the returned matches point into their feature vectors.
  class AnnMatcher
  {
  public:
    AnnMatcher(const KeypointList<OERegion, float>& keys1, const KeypointList<OERegion, float>& keys2,
               float sift_ratio_thres = 1.2f)
      : _keys1{keys1}, _keys2{keys2}, _ratio{sift_ratio_thres}
    {
      check_sizes();
    }
    AnnMatcher(const KeypointList<OERegion, float>& keys, float sift_ratio_thres = 1.2f,
               float min_max_metric_dist_thres = 0.5f, float pixel_dist_thres = 10.f)
      : _keys1{keys}, _keys2{keys}, _ratio{sift_ratio_thres}, _metric{min_max_metric_dist_thres}
      , _pixel{pixel_dist_thres}, _self{true}
    {
      check_sizes();
    }

    std::vector<Match> compute_matches()
    {
      const std::vector<OERegion>& f1 = std::get<0>(_keys1);
      const std::vector<OERegion>& f2 = std::get<0>(_keys2);
      const DescriptorMatrix& d1 = std::get<1>(_keys1);
      const DescriptorMatrix& d2 = std::get<1>(_keys2);
      if (d1.rows() == 0 || d2.rows() == 0)
        throw std::runtime_error{"Error: the list of key-points is empty!"};  // AnnMatcher.cpp:45-46
      std::vector<sara_b200_keypoint> k1(f1.size()), k2(f2.size());
      std::transform(f1.begin(), f1.end(), k1.begin(), to_c_keypoint);
      std::transform(f2.begin(), f2.end(), k2.begin(), to_c_keypoint);
      sara_b200_match_args args;
      sara_b200_default_match_args(&args);
      args.sift_ratio_thres = _ratio;
      args.self_matching = _self ? 1 : 0;
      args.min_max_metric_dist_thres = _metric;
      args.pixel_dist_thres = _pixel;
      Context& ctx = default_context(64, 64);
      std::vector<sara_b200_match> raw(std::max<std::size_t>(1024, 4 * (f1.size() + f2.size())));
      int n = 0;
      int rc = sara_b200_compute_matches(ctx.get(), d1.data(), k1.data(), d1.rows(), d2.data(), k2.data(), d2.rows(),
                                         d1.cols(), 0, &args, raw.data(), static_cast<int>(raw.size()), &n, nullptr);
      if (rc == SARA_B200_ERR_OVERFLOW && n > static_cast<int>(raw.size()))
      {
        raw.resize(n);
        rc = sara_b200_compute_matches(ctx.get(), d1.data(), k1.data(), d1.rows(), d2.data(), k2.data(), d2.rows(),
                                       d1.cols(), 0, &args, raw.data(), n, &n, nullptr);
      }
      if (rc != 0)
        rethrow(rc, ctx.get());
      std::vector<Match> matches;
      matches.reserve(n);
      for (int i = 0; i < n; ++i)
      {
        const sara_b200_match& m = raw[i];
        Match out{&f1[m.x_index], &f2[m.y_index], m.score, static_cast<Match::Direction>(m.direction), m.x_index,
                  m.y_index};
        out.rank() = m.rank;
        matches.push_back(out);
      }
      return matches;
    }
    std::vector<Match> compute_self_matches() { return compute_matches(); }

  private:
    void check_sizes() const
    {
      // size_consistency_predicate (AnnMatcher.cpp:181-184, 198-200)
      if (static_cast<int>(std::get<0>(_keys1).size()) != std::get<1>(_keys1).rows() ||
          static_cast<int>(std::get<0>(_keys2).size()) != std::get<1>(_keys2).rows())
        throw std::runtime_error{"The list of keypoints are inconsistent in size!"};
    }
    const KeypointList<OERegion, float>& _keys1;
    const KeypointList<OERegion, float>& _keys2;
    float _ratio;
    float _metric = 0.5f, _pixel = 10.f;
    bool _self = false;
  };

  //! DO::Sara::match (SfM/Helpers/KeypointMatching.cpp:19-25)
  inline std::vector<Match> match(const KeypointList<OERegion, float>& keys1,
                                  const KeypointList<OERegion, float>& keys2, float lowe_ratio = 0.6f)
  {
    AnnMatcher matcher{keys1, keys2, lowe_ratio};
    return matcher.compute_matches();
  }

}  // namespace sara_b200


// ---- real Sara types (compiled only inside a Sara build) --------------------- //
#if defined(SARA_B200_WITH_SARA)
#  include <DO/Sara/Core/Image.hpp>
#  include <DO/Sara/Features/KeypointList.hpp>
#  include <DO/Sara/ImageProcessing/ImagePyramid.hpp>

namespace sara_b200 {

  // Expected layout of DO::Sara::OERegion with default Eigen alignment.
  static_assert(sizeof(DO::Sara::OERegion) == 48, "unexpected OERegion layout");

  inline auto compute_sift_keypoints(const DO::Sara::ImageView<float>& image,
                                     const DO::Sara::ImagePyramidParams& pp = DO::Sara::ImagePyramidParams(),
                                     float gauss_truncate = 4.f, float extremum_thres = 0.01f,
                                     float edge_ratio_thres = 10.f, int extremum_refinement_iter = 5,
                                     bool /*parallel*/ = false)
      -> DO::Sara::KeypointList<DO::Sara::OERegion, float>
  {
    const sara_b200_pyramid_params cp{pp.first_octave_index(), pp.scale_count_per_octave(),
                                      pp.scale_geometric_factor(), pp.image_padding_size(),
                                      pp.scale_camera(), pp.scale_initial(), pp.num_octaves_max()};
    const auto args =
        make_sift_args(cp, gauss_truncate, extremum_thres, edge_ratio_thres, extremum_refinement_iter);
    return default_context(image.width(), image.height())
        .compute<DO::Sara::OERegion>(image.data(), image.width(), image.height(), args, [](int n, int d) {
          return DO::Sara::Tensor_<float, 2>{n, d};
        });
  }

}  // namespace sara_b200
#endif
