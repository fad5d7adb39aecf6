// Host side of the C ABI (include/sara_b200.h): context, frame slots, pyramid
// geometry, launch sequence.  No CPU compute path exists here: every stage is a
// CUDA kernel launch, and creation fails when no device is usable.
//
// Host-side restatements (scalars only): ImagePyramidParams / gaussian_pyramid
// orchestration (ImageProcessing/GaussianPyramid.hpp:35-125), make_gaussian_kernel
// (ImageProcessing/LinearFiltering.hpp:172-203), ImagePyramid::scale_relative_to_octave
// (ImageProcessing/ImagePyramid.hpp:316-319), the argument plumbing of
// compute_sift_keypoints (FeatureDetectors/SIFT.cpp:27-108, quirk N1) and of
// ComputeDoGExtrema (FeatureDetectors/DoG.cpp:23-87).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "match.cuh"

using namespace sb;

namespace {

  thread_local char g_create_error[512] = "";

  // One side stream per octave after the first (octave o runs on side[o - 1]; beyond kNumSide they wrap).
  constexpr int kNumSide = 8;

  struct Slot
  {
    cudaStream_t own_stream = nullptr;
    // Octaves overlap: octave o + 1 only needs scale `downscale_index` of octave o, so the octaves
    // run on alternating side streams chained by events and are joined back at the end.
    cudaStream_t side[kNumSide] = {};
    cudaEvent_t ev_down[kMaxOctaves] = {};
    cudaEvent_t ev_join[kNumSide] = {};
    cudaStream_t stream = nullptr;  // stream of the frame in flight
    // pyramid arena (grown on demand)
    char* arena = nullptr;
    size_t arena_bytes = 0;
    // keypoint buffers (sized at creation)
    char* kbuf = nullptr;
    Candidate* cand = nullptr;
    Keypoint* ext_tmp = nullptr;
    Keypoint* ext = nullptr;
    Keypoint* kp_oct = nullptr;
    Keypoint* kp_out = nullptr;
    float* desc = nullptr;
    int* ori_count = nullptr;
    float* oris = nullptr;
    int* scratch = nullptr;
    Counters* counters = nullptr;
    Counters* h_counters = nullptr;  // pinned
    // geometry of the frame in flight / last frame
    PyramidDesc P{};
    float* d_img = nullptr;
    uint8_t* d_u8 = nullptr;  // staging of an 8-bit host frame
    float* d_tmp = nullptr;
    int* seg_offsets = nullptr;
    int n_segments = 0;
    int img_w = 0, img_h = 0;
    int downscale_index = 0;
    std::vector<Taps> stage_taps;  // per scale s >= 1
    bool busy = false;
    bool has_keypoints = false;  // the last frame ran the keypoint stages
    bool function_pyramid = false;  // the D stack holds a LoG / DoH pyramid (n_scales layers) instead of the DoG
    int classified_upto = 0;     // octaves already classified while the pyramid was still running
    cudaEvent_t ev[9] = {};  // [7], [8]: around the pyramid's longest launch
    double top_kernel_bytes = 0.;
    sara_b200_timings timings{};
    int pyramid_launches = 0, total_launches = 0, extra_launches = 0, extra_launches_pre = 0;
    // CUDA graphs of whole frames, keyed by everything the launch sequence depends on
    // (geometry, arguments, input kind and -- for device-resident input -- the pointer).
    struct FrameGraph
    {
      int w = 0, h = 0, u8_channels = 0, mode = 0;
      bool on_device = false, overlap = true;
      const void* dev_ptr = nullptr;
      sara_b200_sift_args args{};
      cudaGraphExec_t exec = nullptr;
      PyramidDesc P{};
      int n_segments = 0, downscale_index = 0, pyramid_launches = 0, total_launches = 0;
      float *d_img = nullptr, *d_tmp = nullptr;
      uint8_t* d_u8 = nullptr;
      int* seg_offsets = nullptr;
      unsigned long long last_use = 0;
    };
    std::vector<FrameGraph> graphs;
    unsigned long long graph_clock = 0;
  };

  }  // namespace

// Uploads from PAGEABLE host memory (what a Sara caller's ImageView<float> points at).  cudaMemcpyAsync stages such
// a copy through the driver's bounce buffer on the calling thread (about 12 GB/s: 2.7 ms for a 4K float frame, most
// of the drop-in call); here kThreads host threads each copy a slice of the frame into their own pair of pinned
// chunks and queue the DMA on their own stream, so the host-side copy runs on several cores and overlaps with the
// transfers.  Pinned and device-resident sources never come here.  Four threads measured best on a 4K float frame
// (2: 3.4 ms, 4: 2.6 ms, 6: 3.9 ms, 8: 3.2 ms for the whole call; SARA_B200_STAGER_THREADS overrides).
struct HostStager
{
  static constexpr int kThreads = 8;  // upper bound; `threads` are used
  static constexpr size_t kChunk = size_t(2) << 20;
  static constexpr size_t kMinBytes = size_t(4) << 20;
  unsigned char* pinned = nullptr;  // kThreads x 2 chunks
  cudaStream_t stream[kThreads] = {};
  cudaEvent_t chunk_done[kThreads][2] = {};
  bool chunk_used[kThreads][2] = {};
  cudaEvent_t done[kThreads] = {};
  cudaEvent_t start = nullptr;
  bool ready = false, broken = false;
  int threads = 4;
};

struct sara_b200_ctx
{
  int device = 0;
  sara_b200_limits lim{};
  int cap_kp = 0, cap_ext = 0, cap_cand = 0;
  bool profiling = false;
  int pyramid_mode = SARA_B200_PYRAMID_AUTO;
  bool octave_overlap = true;
  bool use_graphs = true;
  std::vector<Slot> slots;
  double* d_gray_lut = nullptr;   // 3 x 256 products of the rgb -> gray conversion (ingest.cu)
  float* scratch = nullptr;       // sara_b200_gaussian / sara_b200_to_gray32f work buffers (grown on demand)
  size_t scratch_bytes = 0;
  sb::match::Workspace match_ws;  // nearest-neighbour search (match.cu)
  sb::LaplaceTable* d_laplace = nullptr;  // constants of select_laplace_scale (Hessian-Laplace), uploaded per call
  unsigned char* match_io = nullptr;  // device copies of host descriptors + result buffers of the matcher
  size_t match_io_bytes = 0;
  HostStager stager;
  char err[512] = "";
};

namespace {

  int fail(sara_b200_ctx* ctx, int code, const char* fmt, ...)
  {
    char* dst = ctx ? ctx->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
  }

#define CU(ctx, call)                                                                              \
  do                                                                                               \
  {                                                                                                \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? SARA_B200_ERR_OOM : SARA_B200_ERR_CUDA,  \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);     \
  } while (0)

  bool is_pageable(const void* p)
  {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
      cudaGetLastError();
      return false;
    }
    return a.type == cudaMemoryTypeUnregistered;
  }

  // Host -> device copy of `bytes` ordered on `st`, whatever kind of host memory `src` is.
  int upload(sara_b200_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st)
  {
    HostStager& H = ctx->stager;
    if (bytes < HostStager::kMinBytes || H.broken || !is_pageable(src))
    {
      CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
      return 0;
    }
    constexpr int TMAX = HostStager::kThreads;
    constexpr size_t CH = HostStager::kChunk;
    if (!H.ready)
    {
      const char* e = getenv("SARA_B200_STAGER_THREADS");
      const int hw = static_cast<int>(std::thread::hardware_concurrency());
      H.threads = std::max(1, std::min({TMAX, e ? atoi(e) : 4, hw > 0 ? hw : 1}));
      const int T = H.threads;
      bool ok = cudaHostAlloc(reinterpret_cast<void**>(&H.pinned), T * 2 * CH, cudaHostAllocDefault) == cudaSuccess &&
                cudaEventCreateWithFlags(&H.start, cudaEventDisableTiming) == cudaSuccess;
      for (int t = 0; t < T && ok; ++t)
        ok = cudaStreamCreateWithFlags(&H.stream[t], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&H.done[t], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&H.chunk_done[t][0], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&H.chunk_done[t][1], cudaEventDisableTiming) == cudaSuccess;
      if (!ok)
      {
        cudaGetLastError();
        H.broken = true;  // (whatever was created is released by sara_b200_destroy)
        CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return 0;
      }
      H.ready = true;
    }
    // the destination may still be read by work queued on `st`
    CU(ctx, cudaEventRecord(H.start, st));
    const int T = H.threads;
    const size_t part = ((bytes + T - 1) / T + 255) & ~size_t(255);
    cudaError_t err[TMAX];
    auto work = [&](int t) {
      err[t] = cudaSetDevice(ctx->device);
      if (err[t] == cudaSuccess)
        err[t] = cudaStreamWaitEvent(H.stream[t], H.start, 0);
      const size_t lo = std::min(bytes, part * t), hi = std::min(bytes, part * (t + 1));
      int b = 0;
      for (size_t off = lo; off < hi && err[t] == cudaSuccess; off += CH, b ^= 1)
      {
        const size_t n = std::min(CH, hi - off);
        unsigned char* stage = H.pinned + (static_cast<size_t>(t) * 2 + b) * CH;
        if (H.chunk_used[t][b])
          err[t] = cudaEventSynchronize(H.chunk_done[t][b]);  // the previous DMA out of this chunk
        if (err[t] != cudaSuccess)
          break;
        memcpy(stage, static_cast<const unsigned char*>(src) + off, n);
        err[t] = cudaMemcpyAsync(static_cast<unsigned char*>(dst) + off, stage, n, cudaMemcpyHostToDevice, H.stream[t]);
        if (err[t] == cudaSuccess)
          err[t] = cudaEventRecord(H.chunk_done[t][b], H.stream[t]);
        H.chunk_used[t][b] = true;
      }
      if (err[t] == cudaSuccess)
        err[t] = cudaEventRecord(H.done[t], H.stream[t]);
    };
    std::thread th[TMAX - 1];
    for (int t = 1; t < T; ++t)
      th[t - 1] = std::thread(work, t);
    work(0);
    for (int t = 1; t < T; ++t)
      th[t - 1].join();
    for (int t = 0; t < T; ++t)
      if (err[t] != cudaSuccess)
        return fail(ctx, SARA_B200_ERR_CUDA, "staged upload failed: %s", cudaGetErrorString(err[t]));
    for (int t = 0; t < T; ++t)
      CU(ctx, cudaStreamWaitEvent(st, H.done[t], 0));
    return 0;
  }

  // The API calls run on the context's device and leave the caller's current device as it was.
  struct DeviceGuard
  {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device)
    {
      if (cudaGetDevice(&prev) != cudaSuccess)
        prev = -1;
      if (prev != device)
        err = cudaSetDevice(device);
      else
        prev = -1;  // nothing to restore
    }
    ~DeviceGuard()
    {
      if (prev >= 0)
        cudaSetDevice(prev);
    }
  };

  size_t align_up(size_t v, size_t a)
  {
    return (v + a - 1) / a * a;
  }

  // make_gaussian_kernel, LinearFiltering.hpp:172-203 (expf + sequential sum; the
  // taps are computed once on the host and handed to the kernels as bits).
  int make_taps(float sigma, float gauss_truncate, Taps* out)
  {
    int kernel_size = static_cast<int>(2 * gauss_truncate * sigma + 1);
    kernel_size = std::max(3, kernel_size);
    if (kernel_size % 2 == 0)
      ++kernel_size;
    if (kernel_size > kMaxTaps)
      return -kernel_size;
    const int c = kernel_size / 2;
    const float denom = 2 * (sigma * sigma);
    for (int i = 0; i < kernel_size; ++i)
    {
      const float d = static_cast<float>(i) - static_cast<float>(c);
      out->v[i] = expf(-(d * d) / denom);
    }
    float sum = 0.f;
    for (int i = 0; i < kernel_size; ++i)
      sum += out->v[i];
    for (int i = 0; i < kernel_size; ++i)
      out->v[i] /= sum;
    out->n = kernel_size;
    return kernel_size;
  }

  struct Geometry
  {
    int base_w = 0, base_h = 0;
    int n_octaves = 0, n_scales = 0, downscale_index = 0;
    int ow[kMaxOctaves], oh[kMaxOctaves];
    float scaling[kMaxOctaves];
    float resize_factor = 1.f;
    float pre_sigma = 0.f;  // > 0: pre-blur of the input image
    int pre_downscale = 0;  // > 0: first_octave_index > 0
  };

  // gaussian_pyramid(), GaussianPyramid.hpp:35-125: sizes, octave count, scaling.
  int plan_geometry(sara_b200_ctx* ctx, int w, int h, const sara_b200_pyramid_params& pp, Geometry* g)
  {
    if (w <= 0 || h <= 0)
      return fail(ctx, SARA_B200_ERR_BAD_ARG, "image sizes must be positive (got %dx%d)", w, h);
    if (pp.scale_count_per_octave < 1 || pp.scale_count_per_octave > kMaxScales)
      return fail(ctx, SARA_B200_ERR_BAD_ARG, "scale_count_per_octave must be in [1, %d]", kMaxScales);
    if (!(pp.scale_geometric_factor > 1.f))
      return fail(ctx, SARA_B200_ERR_BAD_ARG, "scale_geometric_factor must be > 1");
    if (pp.image_padding_size < 1)
      return fail(ctx, SARA_B200_ERR_BAD_ARG, "image_padding_size must be >= 1");
    if (pp.first_octave_index < -3 || pp.first_octave_index > 6)
      return fail(ctx, SARA_B200_ERR_BAD_ARG, "first_octave_index out of the supported range [-3, 6]");

    const float resize_factor = std::pow(2.f, -static_cast<float>(pp.first_octave_index));
    const float camera_sigma = pp.scale_camera * resize_factor;
    const float init_sigma = pp.scale_initial;
    g->resize_factor = resize_factor;
    g->pre_sigma = 0.f;
    g->pre_downscale = 0;
    if (pp.first_octave_index < 0)
    {
      // enlarge(image, fact): Resize.hpp:190-216, no blur at all (quirk N4).
      const double fact = resize_factor;
      g->base_w = static_cast<int>(static_cast<double>(w) * fact);
      g->base_h = static_cast<int>(static_cast<double>(h) * fact);
    }
    else
    {
      if (camera_sigma < init_sigma)
        g->pre_sigma = std::sqrt(init_sigma * init_sigma - camera_sigma * camera_sigma);
      g->base_w = w;
      g->base_h = h;
      if (pp.first_octave_index > 0)
      {
        g->pre_downscale = static_cast<int>(std::round(1 / resize_factor));
        g->base_w = w / g->pre_downscale;
        g->base_h = h / g->pre_downscale;
        if (g->base_w <= 0 || g->base_h <= 0)
          return fail(ctx, SARA_B200_ERR_BAD_ARG, "image too small for first_octave_index %d",
                      pp.first_octave_index);
      }
    }
    const int l = std::min(g->base_w, g->base_h);
    const int b = pp.image_padding_size;
    int n_oct = std::min(static_cast<int>(logf(l / (2.f * b)) / logf(2.f)), pp.num_octaves_max);
    n_oct = std::max(n_oct, 0);
    if (n_oct > kMaxOctaves)
      return fail(ctx, SARA_B200_ERR_BAD_ARG, "more than %d octaves", kMaxOctaves);
    g->n_octaves = n_oct;
    g->n_scales = pp.scale_count_per_octave;
    g->downscale_index = static_cast<int>(floorf(logf(2.f) / logf(pp.scale_geometric_factor)));
    if (n_oct > 1 && (g->downscale_index < 0 || g->downscale_index >= g->n_scales))
      return fail(ctx, SARA_B200_ERR_BAD_ARG,
                  "downscale index %d is outside the %d scales of an octave (GaussianPyramid.hpp:99-114)",
                  g->downscale_index, g->n_scales);
    int cw = g->base_w, ch = g->base_h;
    for (int o = 0; o < n_oct; ++o)
    {
      if (o > 0)
      {
        cw /= 2;
        ch /= 2;
      }
      if (cw <= 0 || ch <= 0)
      {
        g->n_octaves = o;
        break;
      }
      g->ow[o] = cw;
      g->oh[o] = ch;
      g->scaling[o] = o == 0 ? 1 / resize_factor : g->scaling[o - 1] * 2;
    }
    return 0;
  }

  // Lays the pyramid, extremum maps and row counters out in the slot's arena.
  // `need_f32`: the slot holds a float copy of the input (host float frames, all 8-bit frames);
  // `u8_bytes`: staging for an 8-bit HOST frame.
  int layout_slot(sara_b200_ctx* ctx, Slot& S, int w, int h, const Geometry& g,
                  const sara_b200_pyramid_params& pp, bool need_f32, size_t u8_bytes = 0)
  {
    const int n_s = g.n_scales;
    const int n_ext_layers = std::max(n_s - 3, 0);
    // Room for the sibling detectors (LoG / DoH: a function pyramid with as many layers as the Gaussian
    // one, searched on n_s - 2 of them): one more layer in the D stack, one more map layer.
    const int n_fun_layers = std::max(n_s - 1, 1);  // Hessian-Laplace searches s = 1 .. n_s - 1
    size_t bytes = 0;
    auto take = [&](size_t n) {
      const size_t off = bytes;
      bytes = align_up(bytes + n, 256);
      return off;
    };
    const size_t off_img = need_f32 ? take(sizeof(float) * w * h) : 0;
    const size_t off_u8 = u8_bytes ? take(u8_bytes) : 0;
    const size_t off_tmp = g.pre_downscale > 0 ? take(sizeof(float) * w * h) : 0;
    size_t off_G[kMaxOctaves], off_D[kMaxOctaves], off_map[kMaxOctaves];
    int pitch[kMaxOctaves];
    int n_seg = 0, n_seg_fun = 0;
    for (int o = 0; o < g.n_octaves; ++o)
    {
      pitch[o] = static_cast<int>(align_up(g.ow[o], 32));
      const size_t layer = sizeof(float) * pitch[o] * g.oh[o];
      off_G[o] = take(layer * n_s);
      off_D[o] = take(layer * n_s);
      off_map[o] = take(align_up(g.ow[o], 16) * g.oh[o] * n_fun_layers);
      n_seg += n_ext_layers * g.oh[o];
      n_seg_fun += n_fun_layers * g.oh[o];
    }
    const size_t off_rows = take(sizeof(int) * std::max(n_seg_fun, 1));
    const size_t off_segoff = take(sizeof(int) * std::max(n_seg_fun, 1));

    if (bytes > S.arena_bytes)
    {
      if (S.arena)
      {
        CU(ctx, cudaStreamSynchronize(S.stream ? S.stream : S.own_stream));
        for (auto& fg : S.graphs)  // they point into the old arena
          if (fg.exec)
            cudaGraphExecDestroy(fg.exec);
        S.graphs.clear();
        CU(ctx, cudaFree(S.arena));
        S.arena = nullptr;
        S.arena_bytes = 0;
      }
      CU(ctx, cudaMalloc(&S.arena, bytes));
      S.arena_bytes = bytes;
    }
    S.d_img = need_f32 ? reinterpret_cast<float*>(S.arena + off_img) : nullptr;
    S.d_u8 = u8_bytes ? reinterpret_cast<uint8_t*>(S.arena + off_u8) : nullptr;
    S.d_tmp = g.pre_downscale > 0 ? reinterpret_cast<float*>(S.arena + off_tmp) : nullptr;
    S.seg_offsets = reinterpret_cast<int*>(S.arena + off_segoff);
    S.n_segments = n_seg;
    S.img_w = w;
    S.img_h = h;
    S.downscale_index = g.downscale_index;

    PyramidDesc& P = S.P;
    std::memset(&P, 0, sizeof(P));
    P.n_octaves = g.n_octaves;
    P.n_scales = n_s;
    P.k = pp.scale_geometric_factor;
    for (int s = 0; s < n_s; ++s)  // ImagePyramid.hpp:316-319: pow(float, int) * float is a double
      P.scale_rel[s] = static_cast<float>(std::pow(static_cast<double>(pp.scale_geometric_factor),
                                                   static_cast<double>(s)) *
                                          static_cast<double>(pp.scale_initial));
    int seg_base = 0;
    for (int o = 0; o < g.n_octaves; ++o)
    {
      OctaveDesc& oc = P.oct[o];
      oc.G = reinterpret_cast<float*>(S.arena + off_G[o]);
      oc.D = reinterpret_cast<float*>(S.arena + off_D[o]);
      oc.map = reinterpret_cast<uint8_t*>(S.arena + off_map[o]);
      oc.map_pitch = static_cast<int>(align_up(g.ow[o], 16));
      oc.row_count = reinterpret_cast<int*>(S.arena + off_rows) + seg_base;
      oc.w = g.ow[o];
      oc.h = g.oh[o];
      oc.pitch = pitch[o];
      oc.layer_stride = pitch[o] * g.oh[o];
      oc.seg_base = seg_base;
      oc.scaling = g.scaling[o];
      seg_base += n_ext_layers * g.oh[o];
    }

    // Per-scale increments, GaussianPyramid.hpp:116-121 (default truncate 4, quirk N5).
    S.stage_taps.assign(n_s, Taps{});
    const float k = pp.scale_geometric_factor;
    float sigma_s_1 = pp.scale_initial;
    for (int s = 1; s < n_s; ++s)
    {
      const float ks = k * sigma_s_1;
      const float sigma = sqrtf(ks * ks - sigma_s_1 * sigma_s_1);
      if (make_taps(sigma, 4.f, &S.stage_taps[s]) < 0)
        return fail(ctx, SARA_B200_ERR_BAD_ARG, "Gaussian kernel of scale %d exceeds %d taps", s, kMaxTaps);
      sigma_s_1 *= k;
    }
    return 0;
  }

  // Work buffer of the stand-alone building blocks (kept until the context is destroyed).
  int grow_scratch(sara_b200_ctx* ctx, size_t bytes)
  {
    if (bytes <= ctx->scratch_bytes)
      return 0;
    if (ctx->scratch)
    {
      CU(ctx, cudaStreamSynchronize(ctx->slots[0].own_stream));
      CU(ctx, cudaFree(ctx->scratch));
      ctx->scratch = nullptr;
      ctx->scratch_bytes = 0;
    }
    CU(ctx, cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return 0;
  }

  int check_slot(sara_b200_ctx* ctx, int slot)
  {
    if (!ctx)
      return fail(nullptr, SARA_B200_ERR_BAD_ARG, "null context");
    if (slot < 0 || slot >= static_cast<int>(ctx->slots.size()))
      return fail(ctx, SARA_B200_ERR_BAD_ARG, "slot %d out of range [0, %d)", slot,
                  static_cast<int>(ctx->slots.size()));
    return 0;
  }

  // Gaussian pyramid + DoG pyramid of one frame.
  // `early`: extrema parameters when the caller goes on to the extrema stage; octave 0 is then
  // classified as soon as its own DoG layers exist, beside the smaller octaves still being built.
  // `u8_channels`: 0 for a float image; 1 (gray8) or 3 (interleaved RGB8) for an 8-bit frame,
  // which is converted on the device first (ingest.cu).
  int enqueue_pyramid(sara_b200_ctx* ctx, Slot& S, const void* image_any, int w, int h, bool on_device,
                      const sara_b200_pyramid_params& pp, float gauss_truncate, cudaStream_t st,
                      const ExtremaParams* early = nullptr, int u8_channels = 0, bool skip_upload = false)
  {
    const float* image = static_cast<const float*>(image_any);
    Geometry g;
    int rc = plan_geometry(ctx, w, h, pp, &g);
    if (rc)
      return rc;
    const size_t u8_bytes = static_cast<size_t>(w) * h * u8_channels;
    rc = layout_slot(ctx, S, w, h, g, pp, !on_device || u8_channels > 0, on_device ? 0 : u8_bytes);
    if (rc)
      return rc;
    S.stream = st;
    // a context made for several frames in flight trades a lone frame's latency for machine time
    set_march_schedule(ctx->slots.size() >= 4);
    S.pyramid_launches = 0;
    S.extra_launches = 0;
    S.extra_launches_pre = 0;
    S.classified_upto = 0;
    S.top_kernel_bytes = 0.;
    const bool prof = ctx->profiling;
    if (prof)
      CU(ctx, cudaEventRecord(S.ev[0], st));
    const float* d_src = image;
    if (u8_channels > 0)
    {
      const uint8_t* d_u8 = static_cast<const uint8_t*>(image_any);
      if (!on_device)
      {
        if (!skip_upload)
        {
          const int urc = upload(ctx, S.d_u8, image_any, u8_bytes, st);
          if (urc)
            return urc;
        }
        d_u8 = S.d_u8;
      }
      if (u8_channels == 3)
        launch_rgb8_to_gray32f(d_u8, S.d_img, w * h, ctx->d_gray_lut, st);
      else
        launch_gray8_to_gray32f(d_u8, S.d_img, w * h, st);
      ++S.extra_launches_pre;
      d_src = S.d_img;
    }
    else if (!on_device)
    {
      if (!skip_upload)
      {
        const int urc = upload(ctx, S.d_img, image, sizeof(float) * w * h, st);
        if (urc)
          return urc;
      }
      d_src = S.d_img;
    }
    if (prof)
      CU(ctx, cudaEventRecord(S.ev[1], st));
    const PyramidDesc& P = S.P;
    if (P.n_octaves > 0)
    {
      const OctaveDesc& o0 = P.oct[0];
      if (pp.first_octave_index < 0)
      {
        launch_enlarge(d_src, w, h, w, o0.G, o0.w, o0.h, o0.pitch, st);
        ++S.pyramid_launches;
      }
      else
      {
        Taps pre;
        if (g.pre_sigma > 0.f)
        {
          // fo > 0 forwards gauss_truncate, fo == 0 uses the default 4 (GaussianPyramid.hpp:61,72).
          const float trunc = pp.first_octave_index > 0 ? gauss_truncate : 4.f;
          if (make_taps(g.pre_sigma, trunc, &pre) < 0)
            return fail(ctx, SARA_B200_ERR_BAD_ARG, "pre-blur kernel exceeds %d taps", kMaxTaps);
        }
        if (g.pre_downscale > 0)
        {
          const float* full = d_src;
          if (g.pre_sigma > 0.f)
          {
            launch_gaussian_stage(d_src, w, S.d_tmp, w, nullptr, 0, w, h, pre, st);
            ++S.pyramid_launches;
            full = S.d_tmp;
          }
          launch_downscale(full, w, h, w, o0.G, o0.w, o0.h, o0.pitch, st);
        }
        else if (g.pre_sigma > 0.f)
        {
          // fo == 0: base = gaussian(image, sqrt(s0^2 - s_cam^2)), GaussianPyramid.hpp:69-73
          const int pm = ctx->pyramid_mode;
          bool fast = false;
          if ((pm == SARA_B200_PYRAMID_AUTO || pm == SARA_B200_PYRAMID_MARCH) && march_kernel_supported(pre))
            fast = launch_march(d_src, w, o0.G, nullptr, nullptr, w, h, o0.pitch, 0, 0, 0, pre, st);
          if (!fast && pm != SARA_B200_PYRAMID_GENERIC && stage_kernel_supported(pre.n))
            fast = launch_stage(d_src, w, o0.G, nullptr, nullptr, w, h, o0.pitch, 0, 0, 0, pre, st);
          if (!fast)
            launch_gaussian_stage(d_src, w, o0.G, o0.pitch, nullptr, 0, w, h, pre, st);
        }
        else
          launch_copy2d(d_src, w, o0.G, o0.pitch, w, h, st);
        ++S.pyramid_launches;
      }
      // Which kernels build an octave (sara_b200_set_pyramid_mode): the per-stage marching
      // kernel and the fused octave kernel cover the tap counts of the default schedule; any
      // other schedule runs on the generic kernel.
      const int mode = ctx->pyramid_mode;
      const bool fused = mode == SARA_B200_PYRAMID_FUSED && fused_octave_supported(S.stage_taps.data(), P.n_scales);
      // The marching scatter-form kernel is the default; the older gather-form stage kernel stays
      // selectable.  Both are "one launch per scale" and share the orchestration below.
      bool marched = (mode == SARA_B200_PYRAMID_AUTO || mode == SARA_B200_PYRAMID_MARCH) && P.n_scales > 1;
      for (int s = 1; s < P.n_scales && marched; ++s)
        marched = march_kernel_supported(S.stage_taps[s]);
      bool staged = !marched && (mode == SARA_B200_PYRAMID_AUTO || mode == SARA_B200_PYRAMID_STAGE) && P.n_scales > 1;
      for (int s = 1; s < P.n_scales && staged; ++s)
        staged = stage_kernel_supported(S.stage_taps[s].n);
      const auto launch_inc = marched ? launch_march : launch_stage;
      staged = staged || marched;
      const cudaStream_t main_st = st;
      bool side_used[kNumSide] = {};
      for (int o = 0; o < P.n_octaves; ++o)
      {
        const OctaveDesc& oc = P.oct[o];
        const OctaveDesc* next = o + 1 < P.n_octaves ? &P.oct[o + 1] : nullptr;
        if (staged && o > 0 && ctx->octave_overlap)
        {
          // Octave o starts as soon as its base exists (event recorded below, after the launch
          // that wrote it) and runs beside the remaining scales of the octaves above it.
          st = S.side[(o - 1) % kNumSide];
          side_used[(o - 1) % kNumSide] = true;
          CU(ctx, cudaStreamWaitEvent(st, S.ev_down[o - 1], 0));
        }
        // The small octaves at the end of the pyramid run in one single-CTA launch.
        static const int tail_px = [] {
          const char* e = getenv("SARA_B200_TAIL_PIXELS");
          return e ? atoi(e) : 4096;
        }();
        if (mode != SARA_B200_PYRAMID_GENERIC && o > 0 && oc.w * oc.h <= tail_px)
        {
          const int n = launch_tail_octaves(P, o, S.downscale_index, S.stage_taps.data(), st);
          if (n > 0)
          {
            S.pyramid_launches += n;
            break;
          }
        }
        if (staged)
        {
          const bool fuse_down = next != nullptr && S.downscale_index >= 1 &&
                                 downscale_is_even_sampling(oc.w, oc.h, next->w, next->h);
          int s_first = 1;
          // Octaves that cannot fill the machine: scales 1 and 2 -- all that the next octave waits for --
          // in one launch (octave_head_kernel), the remaining scales as usual.
          static const int head_px = [] {
            const char* e = getenv("SARA_B200_HEAD_PIXELS");
            return e ? atoi(e) : 600000;
          }();
          if (o > 0 && oc.w * oc.h <= head_px && P.n_scales >= 3 && S.downscale_index == 2 && (fuse_down || !next) &&
              launch_octave_head(oc, next, S.stage_taps[1], S.stage_taps[2], st))
          {
            ++S.pyramid_launches;
            if (next)
              CU(ctx, cudaEventRecord(S.ev_down[o], st));
            s_first = 3;
          }
          for (int s = s_first; s < P.n_scales; ++s)
          {
            const bool down = fuse_down && s == S.downscale_index;
            // The launch with the most taps on the largest octave is the pyramid's longest kernel.
            const bool top = prof && o == 0 && s == P.n_scales - 1;
            if (top)
            {
              CU(ctx, cudaEventRecord(S.ev[7], st));
              S.top_kernel_bytes = 12.0 * oc.w * oc.h;  // reads G(s-1), writes G(s) and D(s-1): 3 x 4 B per pixel
            }
            if (!launch_inc(oc.G + static_cast<size_t>(s - 1) * oc.layer_stride, oc.pitch,
                              oc.G + static_cast<size_t>(s) * oc.layer_stride,
                              oc.D + static_cast<size_t>(s - 1) * oc.layer_stride, down ? next->G : nullptr, oc.w, oc.h,
                              oc.pitch, down ? next->w : 0, down ? next->h : 0, down ? next->pitch : 0, S.stage_taps[s],
                              st))
              return fail(ctx, SARA_B200_ERR_CUDA, "stage kernel could not be launched (tensor map / attributes)");
            ++S.pyramid_launches;
            if (top)
              CU(ctx, cudaEventRecord(S.ev[8], st));
            if (next && !fuse_down && s == S.downscale_index)
            {
              launch_downscale(oc.G + static_cast<size_t>(s) * oc.layer_stride, oc.w, oc.h, oc.pitch, next->G, next->w,
                               next->h, next->pitch, st);
              ++S.pyramid_launches;
            }
            if (next && s == S.downscale_index)
              CU(ctx, cudaEventRecord(S.ev_down[o], st));  // the base of octave o + 1 is written
          }
          if (next && S.downscale_index == 0)
          {
            if (!fuse_down)
            {
              launch_downscale(oc.G, oc.w, oc.h, oc.pitch, next->G, next->w, next->h, next->pitch, st);
              ++S.pyramid_launches;
            }
            CU(ctx, cudaEventRecord(S.ev_down[o], st));
          }
          // (not under profiling: the stage timings must not overlap)
          if (o == 0 && early != nullptr && !prof && ctx->octave_overlap && P.n_octaves > 1 && P.n_scales >= 4 &&
              S.n_segments > 0)
          {
            // octave 0 is complete on this stream: classify it now (its DoG layers are still in L2)
            S.extra_launches = launch_classify(P, *early, S.n_segments, 0, 1, true, st);
            S.classified_upto = 1;
          }
          continue;
        }
        if (fused)
        {
          if (prof && o == 0)
          {
            CU(ctx, cudaEventRecord(S.ev[7], st));
            S.top_kernel_bytes = 48.0 * oc.w * oc.h;  // reads G(0), writes G(1..5) and D(0..4)
          }
          const int n = launch_fused_octave(oc, next, S.downscale_index, S.stage_taps.data(), P.n_scales, st);
          if (prof && o == 0)
            CU(ctx, cudaEventRecord(S.ev[8], st));
          if (n < 0)
            return fail(ctx, SARA_B200_ERR_CUDA, "fused octave kernel could not be launched (tensor map / attributes)");
          S.pyramid_launches += n;
          continue;
        }
        for (int s = 1; s < P.n_scales; ++s)
        {
          launch_gaussian_stage(oc.G + static_cast<size_t>(s - 1) * oc.layer_stride, oc.pitch,
                                oc.G + static_cast<size_t>(s) * oc.layer_stride, oc.pitch,
                                oc.D + static_cast<size_t>(s - 1) * oc.layer_stride, oc.pitch, oc.w, oc.h,
                                S.stage_taps[s], st);
          ++S.pyramid_launches;
          if (next && s == S.downscale_index)
          {
            launch_downscale(oc.G + static_cast<size_t>(s) * oc.layer_stride, oc.w, oc.h, oc.pitch,
                             next->G, next->w, next->h, next->pitch, st);
            ++S.pyramid_launches;
          }
        }
        if (next && S.downscale_index == 0)
        {
          launch_downscale(oc.G, oc.w, oc.h, oc.pitch, next->G, next->w, next->h, next->pitch, st);
          ++S.pyramid_launches;
        }
      }
      // join the side streams back into the frame's stream
      st = main_st;
      for (int i = 0; i < kNumSide; ++i)
        if (side_used[i])
        {
          CU(ctx, cudaEventRecord(S.ev_join[i], S.side[i]));
          CU(ctx, cudaStreamWaitEvent(st, S.ev_join[i], 0));
        }
    }
    if (prof)
      CU(ctx, cudaEventRecord(S.ev[2], st));
    CU(ctx, cudaGetLastError());
    S.total_launches = S.pyramid_launches + S.extra_launches + S.extra_launches_pre;
    return 0;
  }

  int enqueue_extrema(sara_b200_ctx* ctx, Slot& S, float extremum_thres, float edge_ratio, int pad,
                      int refine_iter, cudaStream_t st)
  {
    if (S.P.n_scales < 4)  // DoG.hpp:86-89
      return fail(ctx, SARA_B200_ERR_TOO_FEW_SCALES,
                  "Error: The extraction of DoG extrema needs (1 + 3) = 4 scales per octave at the very "
                  "minimum!");
    if (pad < 1)
      return fail(ctx, SARA_B200_ERR_BAD_ARG,
                  "img_padding_sz must be >= 1 (the 3x3x3 neighbourhood must stay inside the layer)");
    CU(ctx, cudaMemsetAsync(S.counters, 0, sizeof(Counters), st));
    if (S.P.n_octaves > 0 && S.n_segments > 0)
    {
      ExtremaParams ep{extremum_thres, edge_ratio, pad, refine_iter};
      S.total_launches += launch_extrema(S.P, ep, S.n_segments, S.seg_offsets, S.cand, ctx->cap_cand,
                                         S.ext_tmp, S.classified_upto, S.scratch, S.ext, ctx->cap_ext, S.counters, st);
    }
    if (ctx->profiling)
      CU(ctx, cudaEventRecord(S.ev[3], st));
    CU(ctx, cudaGetLastError());
    return 0;
  }

  int enqueue_describe(sara_b200_ctx* ctx, Slot& S, cudaStream_t st)
  {
    if (S.P.n_octaves > 0)
    {
      S.total_launches += launch_orientation(S.P, S.ext, ctx->cap_ext, S.ori_count, S.oris, S.scratch,
                                             S.kp_oct, ctx->cap_kp, S.counters, st);
      if (ctx->profiling)
        CU(ctx, cudaEventRecord(S.ev[4], st));
      S.total_launches += launch_descriptors(S.P, S.kp_oct, S.kp_out, S.desc, ctx->cap_kp, S.counters, st);
    }
    else if (ctx->profiling)
      CU(ctx, cudaEventRecord(S.ev[4], st));
    if (ctx->profiling)
      CU(ctx, cudaEventRecord(S.ev[5], st));
    CU(ctx, cudaGetLastError());
    return 0;
  }

  int finish_enqueue(sara_b200_ctx* ctx, Slot& S, cudaStream_t st)
  {
    CU(ctx, cudaMemcpyAsync(S.h_counters, S.counters, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    S.busy = true;
    return 0;
  }

  int wait_slot(sara_b200_ctx* ctx, Slot& S)
  {
    CU(ctx, cudaStreamSynchronize(S.stream ? S.stream : S.own_stream));
    if (ctx->profiling && S.busy)
    {
      auto ms = [&](int a, int b) {
        float t = 0.f;
        cudaEventElapsedTime(&t, S.ev[a], S.ev[b]);
        return t;
      };
      S.timings = sara_b200_timings{};
      S.timings.upload = ms(0, 1);
      S.timings.pyramid = ms(1, 2);
      if (S.top_kernel_bytes > 0.)
      {
        S.timings.pyramid_top_kernel = ms(7, 8);
        S.timings.pyramid_top_kernel_mbytes = static_cast<float>(S.top_kernel_bytes * 1e-6);
      }
      if (S.has_keypoints)
      {
        S.timings.extrema = ms(2, 3);
        S.timings.orientation = ms(3, 4);
        S.timings.descriptor = ms(4, 5);
        S.timings.total = ms(1, 5);
      }
      else
        S.timings.total = ms(1, 2);
      cudaGetLastError();
    }
    S.timings.pyramid_launches = S.pyramid_launches;
    S.timings.total_launches = S.total_launches;
    S.busy = false;
    return 0;
  }

  int copy_keypoints(sara_b200_ctx* ctx, Slot& S, const Keypoint* d_src, int n_true, int cap_dev,
                     sara_b200_keypoint* dst, int capacity, int* n_out)
  {
    if (n_out)
      *n_out = n_true;
    if (n_true > cap_dev)
      return fail(ctx, SARA_B200_ERR_OVERFLOW, "%d entries exceed the context capacity %d", n_true, cap_dev);
    if (n_true > capacity)
      return fail(ctx, SARA_B200_ERR_OVERFLOW, "%d entries exceed the caller's capacity %d", n_true, capacity);
    if (dst && n_true > 0)
      CU(ctx, cudaMemcpy(dst, d_src, sizeof(Keypoint) * n_true, cudaMemcpyDeviceToHost));
    return 0;
  }

}  // namespace

// ============================================================================ //
extern "C" {

int sara_b200_version(void)
{
  return SARA_B200_VERSION;
}

const char* sara_b200_last_error(const sara_b200_ctx* ctx)
{
  return ctx ? ctx->err : g_create_error;
}

void sara_b200_default_pyramid_params(sara_b200_pyramid_params* p)
{
  p->first_octave_index = -1;
  p->scale_count_per_octave = 3 + 3;
  p->scale_geometric_factor = std::pow(2.f, 1.f / 3.f);
  p->image_padding_size = 1;
  p->scale_camera = 0.5f;
  p->scale_initial = 1.6f;
  p->num_octaves_max = INT_MAX;
}

void sara_b200_default_sift_args(sara_b200_sift_args* a)
{
  sara_b200_default_pyramid_params(&a->pyramid_params);
  a->gauss_truncate = 4.f;
  a->extremum_thres = 0.01f;
  a->edge_ratio_thres = 10.f;
  a->extremum_refinement_iter = 5;
}

void sara_b200_default_dog_args(sara_b200_dog_args* a)
{
  sara_b200_default_pyramid_params(&a->pyramid_params);
  a->gauss_truncate = 4.f;
  a->extremum_thres = 0.01f;
  a->edge_ratio_thres = 10.f;
  a->img_padding_sz = 1;
  a->extremum_refinement_iter = 5;
}

int sara_b200_create(int device, const sara_b200_limits* limits, sara_b200_ctx** out)
{
  if (!out || !limits)
    return fail(nullptr, SARA_B200_ERR_BAD_ARG, "null argument");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, SARA_B200_ERR_CUDA,
                "no usable CUDA device (%s); this library has no CPU path",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= n_dev)
    return fail(nullptr, SARA_B200_ERR_BAD_ARG, "device %d out of range [0, %d)", device, n_dev);
  if (limits->max_width <= 0 || limits->max_height <= 0)
    return fail(nullptr, SARA_B200_ERR_BAD_ARG, "limits: max_width / max_height must be positive");
  DeviceGuard guard(device);
  CU(nullptr, guard.err);

  auto* ctx = new sara_b200_ctx;
  ctx->device = device;
  ctx->lim = *limits;
  if (limits->max_keypoints > (1 << 22))
  {
    delete ctx;
    return fail(nullptr, SARA_B200_ERR_BAD_ARG, "limits: max_keypoints must be <= %d", 1 << 22);
  }
  ctx->cap_kp = limits->max_keypoints > 0 ? limits->max_keypoints : 262144;
  ctx->cap_ext = ctx->cap_kp;
  ctx->cap_cand = 2 * ctx->cap_kp;
  const int n_slots = limits->num_slots > 0 ? limits->num_slots : 1;
  ctx->slots.resize(n_slots);

  for (Slot& S : ctx->slots)
  {
    cudaError_t err = cudaStreamCreateWithFlags(&S.own_stream, cudaStreamNonBlocking);
    // The smaller an octave, the shorter its launches and the longer the chain that still hangs on it
    // (octave o + 1 waits for scale 2 of octave o): the side streams get rising priorities, so that a
    // small octave's CTAs take the slots that free up before the big octaves' next launches do.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lower = more urgent
    static const int use_prio = [] {
      const char* e = getenv("SARA_B200_SIDE_PRIORITY");
      return e ? atoi(e) : 1;
    }();
    for (int i = 0; i < kNumSide && err == cudaSuccess; ++i)
    {
      const int prio = std::max(prio_hi, prio_lo - (i + 1));
      err = use_prio ? cudaStreamCreateWithPriority(&S.side[i], cudaStreamNonBlocking, prio)
                     : cudaStreamCreateWithFlags(&S.side[i], cudaStreamNonBlocking);
      if (err == cudaSuccess)
        err = cudaEventCreateWithFlags(&S.ev_join[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < kMaxOctaves && err == cudaSuccess; ++i)
      err = cudaEventCreateWithFlags(&S.ev_down[i], cudaEventDisableTiming);
    for (int i = 0; i < 9 && err == cudaSuccess; ++i)
      err = cudaEventCreate(&S.ev[i]);
    size_t bytes = 0;
    auto take = [&](size_t n) {
      const size_t off = bytes;
      bytes = align_up(bytes + n, 256);
      return off;
    };
    const size_t o_cand = take(sizeof(Candidate) * ctx->cap_cand);
    const size_t o_tmp = take(sizeof(Keypoint) * ctx->cap_cand);
    const size_t o_ext = take(sizeof(Keypoint) * ctx->cap_ext);
    const size_t o_kpo = take(sizeof(Keypoint) * ctx->cap_kp);
    const size_t o_kpf = take(sizeof(Keypoint) * ctx->cap_kp);
    const size_t o_desc = take(sizeof(float) * 128 * ctx->cap_kp);
    const size_t o_oc = take(sizeof(int) * ctx->cap_ext);
    const size_t o_or = take(sizeof(float) * kMaxOri * ctx->cap_ext);
    const size_t o_scr = take(sizeof(int) * (1024 + 2 * static_cast<size_t>(ctx->cap_cand)));
    const size_t o_cnt = take(sizeof(Counters));
    if (err == cudaSuccess)
      err = cudaMalloc(&S.kbuf, bytes);
    if (err == cudaSuccess)
      err = cudaHostAlloc(&S.h_counters, sizeof(Counters), cudaHostAllocDefault);
    if (err != cudaSuccess)
    {
      const int code = fail(nullptr, err == cudaErrorMemoryAllocation ? SARA_B200_ERR_OOM : SARA_B200_ERR_CUDA,
                            "context creation failed: %s", cudaGetErrorString(err));
      sara_b200_destroy(ctx);
      return code;
    }
    S.cand = reinterpret_cast<Candidate*>(S.kbuf + o_cand);
    S.ext_tmp = reinterpret_cast<Keypoint*>(S.kbuf + o_tmp);
    S.ext = reinterpret_cast<Keypoint*>(S.kbuf + o_ext);
    S.kp_oct = reinterpret_cast<Keypoint*>(S.kbuf + o_kpo);
    S.kp_out = reinterpret_cast<Keypoint*>(S.kbuf + o_kpf);
    S.desc = reinterpret_cast<float*>(S.kbuf + o_desc);
    S.ori_count = reinterpret_cast<int*>(S.kbuf + o_oc);
    S.oris = reinterpret_cast<float*>(S.kbuf + o_or);
    S.scratch = reinterpret_cast<int*>(S.kbuf + o_scr);
    S.counters = reinterpret_cast<Counters*>(S.kbuf + o_cnt);
    std::memset(S.h_counters, 0, sizeof(Counters));
    cudaMemset(S.counters, 0, sizeof(Counters));

    // Size the pyramid arena for the largest frame up front (default scale count).
    sara_b200_pyramid_params pp;
    sara_b200_default_pyramid_params(&pp);
    pp.first_octave_index = limits->min_first_octave_index < 0 ? limits->min_first_octave_index : 0;
    Geometry g;
    if (plan_geometry(ctx, limits->max_width, limits->max_height, pp, &g) == 0)
    {
      const int rc = layout_slot(ctx, S, limits->max_width, limits->max_height, g, pp, true,
                                 static_cast<size_t>(3) * limits->max_width * limits->max_height);
      if (rc)
      {
        std::memcpy(g_create_error, ctx->err, sizeof(g_create_error));
        sara_b200_destroy(ctx);
        return rc;
      }
    }
    S.P = PyramidDesc{};
  }
  {
    double lut[768];
    fill_rgb_to_gray_lut(lut);
    cudaError_t err = cudaMalloc(&ctx->d_gray_lut, sizeof(lut));
    if (err == cudaSuccess)
      err = cudaMemcpy(ctx->d_gray_lut, lut, sizeof(lut), cudaMemcpyHostToDevice);
    if (err != cudaSuccess)
    {
      const int code = fail(nullptr, SARA_B200_ERR_CUDA, "context creation failed: %s", cudaGetErrorString(err));
      sara_b200_destroy(ctx);
      return code;
    }
  }
  *out = ctx;
  return 0;
}

void sara_b200_destroy(sara_b200_ctx* ctx)
{
  if (!ctx)
    return;
  DeviceGuard guard(ctx->device);
  for (Slot& S : ctx->slots)
  {
    if (S.own_stream)
      cudaStreamSynchronize(S.own_stream);
    if (S.stream && S.stream != S.own_stream && S.busy)
      cudaStreamSynchronize(S.stream);
    for (auto& fg : S.graphs)
      if (fg.exec)
        cudaGraphExecDestroy(fg.exec);
    cudaFree(S.arena);
    cudaFree(S.kbuf);
    cudaFreeHost(S.h_counters);
    for (auto& ev : S.ev)
      if (ev)
        cudaEventDestroy(ev);
    for (auto& e : S.ev_down)
      if (e)
        cudaEventDestroy(e);
    for (auto& e : S.ev_join)
      if (e)
        cudaEventDestroy(e);
    for (auto& sd : S.side)
      if (sd)
      {
        cudaStreamSynchronize(sd);
        cudaStreamDestroy(sd);
      }
    if (S.own_stream)
      cudaStreamDestroy(S.own_stream);
  }
  {
    HostStager& H = ctx->stager;
    for (int t = 0; t < HostStager::kThreads; ++t)
    {
      if (H.stream[t])
      {
        cudaStreamSynchronize(H.stream[t]);
        cudaStreamDestroy(H.stream[t]);
      }
      for (cudaEvent_t e : {H.done[t], H.chunk_done[t][0], H.chunk_done[t][1]})
        if (e)
          cudaEventDestroy(e);
    }
    if (H.start)
      cudaEventDestroy(H.start);
    if (H.pinned)
      cudaFreeHost(H.pinned);
  }
  cudaFree(ctx->d_gray_lut);
  cudaFree(ctx->scratch);
  cudaFree(ctx->match_io);
  cudaFree(ctx->d_laplace);
  ctx->match_ws.release();
  delete ctx;
}

int sara_b200_host_alloc(void** ptr, uint64_t bytes)
{
  if (!ptr)
    return SARA_B200_ERR_BAD_ARG;
  cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess)
    return fail(nullptr, SARA_B200_ERR_OOM, "cudaHostAlloc(%llu) failed: %s",
                static_cast<unsigned long long>(bytes), cudaGetErrorString(e));
  return 0;
}

void sara_b200_host_free(void* ptr)
{
  if (ptr)
    cudaFreeHost(ptr);
}

int sara_b200_set_profiling(sara_b200_ctx* ctx, int on)
{
  if (!ctx)
    return SARA_B200_ERR_BAD_ARG;
  ctx->profiling = on != 0;
  return 0;
}

int sara_b200_set_pyramid_mode(sara_b200_ctx* ctx, int mode)
{
  if (!ctx)
    return SARA_B200_ERR_BAD_ARG;
  if (mode < SARA_B200_PYRAMID_AUTO || mode > SARA_B200_PYRAMID_MARCH)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "unknown pyramid mode %d", mode);
  ctx->pyramid_mode = mode;
  return 0;
}

int sara_b200_set_octave_overlap(sara_b200_ctx* ctx, int on)
{
  if (!ctx)
    return SARA_B200_ERR_BAD_ARG;
  ctx->octave_overlap = on != 0;
  return 0;
}

int sara_b200_set_graphs(sara_b200_ctx* ctx, int on)
{
  if (!ctx)
    return SARA_B200_ERR_BAD_ARG;
  ctx->use_graphs = on != 0;
  return 0;
}

int sara_b200_last_timings(sara_b200_ctx* ctx, int slot, sara_b200_timings* out)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  if (!out)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "null output");
  *out = ctx->slots[slot].timings;
  return 0;
}

static int sift_enqueue_impl(sara_b200_ctx* ctx, int slot, const void* image, int w, int h, int image_on_device,
                             const sara_b200_sift_args* args, void* stream, int u8_channels)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  if (!image || !args)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "null argument");
  if (u8_channels != 0 && u8_channels != 1 && u8_channels != 3)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "8-bit frames have 1 (gray) or 3 (interleaved RGB) channels, got %d",
                u8_channels);
  if (u8_channels > 0 && image_on_device && (reinterpret_cast<uintptr_t>(image) & 3) != 0)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "a device-resident 8-bit frame must be 4-byte aligned");
  Slot& S = ctx->slots[slot];
  if (S.busy)
    return fail(ctx, SARA_B200_ERR_BUSY, "slot %d holds an un-collected frame", slot);
  if (args->pyramid_params.scale_count_per_octave < 4)
    return fail(ctx, SARA_B200_ERR_TOO_FEW_SCALES,
                "Error: The extraction of DoG extrema needs (1 + 3) = 4 scales per octave at the very "
                "minimum!");
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : S.own_stream;
  S.function_pyramid = false;
  if (args->extremum_refinement_iter < 1)  // it becomes img_padding_sz (quirk N1); checked before any launch
    return fail(ctx, SARA_B200_ERR_BAD_ARG,
                "img_padding_sz must be >= 1 (the 3x3x3 neighbourhood must stay inside the layer)");
  const ExtremaParams early{args->extremum_thres, args->edge_ratio_thres, args->extremum_refinement_iter, 5};
  const bool on_device = image_on_device != 0;
  // Records the frame's whole launch sequence on `st` (everything but the host -> device copy
  // when `skip_upload`).  Quirk N1 (SIFT.cpp:45-51 vs DoG.hpp:72-78): extremum_refinement_iter
  // lands in the img_padding_sz slot; the iteration count keeps its default, 5.
  auto record = [&](bool skip_upload) -> int {
    int r = enqueue_pyramid(ctx, S, image, w, h, on_device, args->pyramid_params, args->gauss_truncate, st, &early,
                            u8_channels, skip_upload);
    if (r)
      return r;
    r = enqueue_extrema(ctx, S, args->extremum_thres, args->edge_ratio_thres, args->extremum_refinement_iter, 5, st);
    if (r)
      return r;
    r = enqueue_describe(ctx, S, st);
    if (r)
      return r;
    S.has_keypoints = true;
    return finish_enqueue(ctx, S, st);
  };
  if (!ctx->use_graphs || ctx->profiling)
    return record(false);

  // ---- CUDA-graph path: one graph launch per frame instead of ~45 kernel launches, side-stream
  // events and tensor-map encodes (the sequence is captured once per geometry / argument set) ----
  Slot::FrameGraph* fg = nullptr;
  for (auto& g : S.graphs)
    if (g.w == w && g.h == h && g.u8_channels == u8_channels && g.on_device == on_device &&
        g.mode == ctx->pyramid_mode && g.overlap == ctx->octave_overlap && (!on_device || g.dev_ptr == image) &&
        std::memcmp(&g.args, args, sizeof(*args)) == 0)
    {
      fg = &g;
      break;
    }
  if (!fg)
  {
    // size the arena outside the capture (layout_slot may allocate)
    Geometry geo;
    rc = plan_geometry(ctx, w, h, args->pyramid_params, &geo);
    if (rc)
      return rc;
    rc = layout_slot(ctx, S, w, h, geo, args->pyramid_params, !on_device || u8_channels > 0,
                     on_device ? 0 : static_cast<size_t>(w) * h * u8_channels);
    if (rc)
      return rc;
    if (S.graphs.size() >= 8)  // evict the least recently used
    {
      auto lru = std::min_element(S.graphs.begin(), S.graphs.end(),
                                  [](const Slot::FrameGraph& a, const Slot::FrameGraph& b) { return a.last_use < b.last_use; });
      cudaGraphExecDestroy(lru->exec);
      S.graphs.erase(lru);
    }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess)
    {
      cudaGetLastError();
      ctx->use_graphs = false;
      return record(false);
    }
    rc = record(true);
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    S.busy = false;
    if (rc != 0 || ce != cudaSuccess || graph == nullptr)
    {
      if (graph)
        cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc)
        return rc;  // a genuine argument error: the direct path would fail the same way
      ctx->use_graphs = false;
      return record(false);
    }
    Slot::FrameGraph g;
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess)
    {
      cudaGetLastError();
      ctx->use_graphs = false;
      return record(false);
    }
    g.w = w;
    g.h = h;
    g.u8_channels = u8_channels;
    g.mode = ctx->pyramid_mode;
    g.on_device = on_device;
    g.overlap = ctx->octave_overlap;
    g.dev_ptr = on_device ? image : nullptr;
    g.args = *args;
    g.P = S.P;
    g.n_segments = S.n_segments;
    g.downscale_index = S.downscale_index;
    g.pyramid_launches = S.pyramid_launches;
    g.total_launches = S.total_launches;
    g.d_img = S.d_img;
    g.d_tmp = S.d_tmp;
    g.d_u8 = S.d_u8;
    g.seg_offsets = S.seg_offsets;
    S.graphs.push_back(g);
    fg = &S.graphs.back();
  }
  // restore the slot state the accessors read (another geometry may have run in between)
  S.P = fg->P;
  S.n_segments = fg->n_segments;
  S.downscale_index = fg->downscale_index;
  S.pyramid_launches = fg->pyramid_launches;
  S.total_launches = fg->total_launches;
  S.d_img = fg->d_img;
  S.d_tmp = fg->d_tmp;
  S.d_u8 = fg->d_u8;
  S.seg_offsets = fg->seg_offsets;
  S.img_w = w;
  S.img_h = h;
  S.stream = st;
  S.classified_upto = 0;
  fg->last_use = ++S.graph_clock;
  if (!on_device)
  {
    const int urc = u8_channels > 0 ? upload(ctx, S.d_u8, image, static_cast<size_t>(w) * h * u8_channels, st)
                                    : upload(ctx, S.d_img, image, sizeof(float) * w * h, st);
    if (urc)
      return urc;
  }
  CU(ctx, cudaGraphLaunch(fg->exec, st));
  S.has_keypoints = true;
  S.busy = true;
  return 0;
}

int sara_b200_sift_enqueue(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                           int image_on_device, const sara_b200_sift_args* args, void* stream)
{
  return sift_enqueue_impl(ctx, slot, image, w, h, image_on_device, args, stream, 0);
}

int sara_b200_sift_enqueue_u8(sara_b200_ctx* ctx, int slot, const uint8_t* image, int w, int h, int channels,
                              int image_on_device, const sara_b200_sift_args* args, void* stream)
{
  if (channels != 1 && channels != 3)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "8-bit frames have 1 (gray) or 3 (interleaved RGB) channels, got %d",
                channels);
  return sift_enqueue_impl(ctx, slot, image, w, h, image_on_device, args, stream, channels);
}

int sara_b200_wait(sara_b200_ctx* ctx, int slot, int* n_out)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  Slot& S = ctx->slots[slot];
  rc = wait_slot(ctx, S);
  if (rc)
    return rc;
  const Counters c = *S.h_counters;
  if (n_out)
    *n_out = S.has_keypoints ? std::min(c.n_kp, ctx->cap_kp) : 0;
  if (c.overflow)
    return fail(ctx, SARA_B200_ERR_OVERFLOW,
                "frame exceeds the context capacity (candidates %d, extrema %d, keypoints %d; max_keypoints %d)",
                c.n_cand, c.n_ext, c.n_kp, ctx->cap_kp);
  return 0;
}

int sara_b200_collect(sara_b200_ctx* ctx, int slot, sara_b200_keypoint* keypoints, float* descriptors,
                      int capacity, int* n_out)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  Slot& S = ctx->slots[slot];
  rc = wait_slot(ctx, S);
  if (rc)
    return rc;
  const Counters c = *S.h_counters;
  if (n_out)
    *n_out = c.n_kp;
  if (c.overflow)
    return fail(ctx, SARA_B200_ERR_OVERFLOW,
                "frame exceeds the context capacity (candidates %d, extrema %d, keypoints %d; max_keypoints %d)",
                c.n_cand, c.n_ext, c.n_kp, ctx->cap_kp);
  if (c.n_kp > capacity)
    return fail(ctx, SARA_B200_ERR_OVERFLOW, "%d keypoints exceed the caller's capacity %d", c.n_kp, capacity);
  cudaStream_t st = S.stream ? S.stream : S.own_stream;
  if (c.n_kp > 0)
  {
    if (keypoints)
      CU(ctx, cudaMemcpyAsync(keypoints, S.kp_out, sizeof(Keypoint) * c.n_kp, cudaMemcpyDeviceToHost, st));
    if (descriptors)
      CU(ctx, cudaMemcpyAsync(descriptors, S.desc, sizeof(float) * 128 * c.n_kp, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
  }
  return 0;
}

int sara_b200_collect_device(sara_b200_ctx* ctx, int slot, sara_b200_keypoint* d_keypoints, float* d_descriptors,
                             int capacity, int* n_out)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  Slot& S = ctx->slots[slot];
  rc = wait_slot(ctx, S);
  if (rc)
    return rc;
  const Counters c = *S.h_counters;
  if (n_out)
    *n_out = c.n_kp;
  if (c.overflow)
    return fail(ctx, SARA_B200_ERR_OVERFLOW,
                "frame exceeds the context capacity (candidates %d, extrema %d, keypoints %d; max_keypoints %d)",
                c.n_cand, c.n_ext, c.n_kp, ctx->cap_kp);
  if (c.n_kp > capacity)
    return fail(ctx, SARA_B200_ERR_OVERFLOW, "%d keypoints exceed the caller's capacity %d", c.n_kp, capacity);
  cudaStream_t st = S.stream ? S.stream : S.own_stream;
  if (c.n_kp > 0)
  {
    if (d_keypoints)
      CU(ctx, cudaMemcpyAsync(d_keypoints, S.kp_out, sizeof(Keypoint) * c.n_kp, cudaMemcpyDeviceToDevice, st));
    if (d_descriptors)
      CU(ctx, cudaMemcpyAsync(d_descriptors, S.desc, sizeof(float) * 128 * c.n_kp, cudaMemcpyDeviceToDevice, st));
    CU(ctx, cudaStreamSynchronize(st));
  }
  return 0;
}

int sara_b200_device_results(sara_b200_ctx* ctx, int slot, const sara_b200_keypoint** keypoints,
                             const float** descriptors, int* n_out)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  Slot& S = ctx->slots[slot];
  rc = wait_slot(ctx, S);
  if (rc)
    return rc;
  if (keypoints)
    *keypoints = S.kp_out;
  if (descriptors)
    *descriptors = S.desc;
  if (n_out)
    *n_out = std::min(S.h_counters->n_kp, ctx->cap_kp);
  return S.h_counters->overflow ? fail(ctx, SARA_B200_ERR_OVERFLOW, "frame exceeds the context capacity") : 0;
}

int sara_b200_sift(sara_b200_ctx* ctx, const float* image, int w, int h, int image_on_device,
                   const sara_b200_sift_args* args, sara_b200_keypoint* keypoints, float* descriptors,
                   int capacity, int* n_out)
{
  int rc = sara_b200_sift_enqueue(ctx, 0, image, w, h, image_on_device, args, nullptr);
  if (rc)
    return rc;
  return sara_b200_collect(ctx, 0, keypoints, descriptors, capacity, n_out);
}

int sara_b200_sift_u8(sara_b200_ctx* ctx, const uint8_t* image, int w, int h, int channels, int image_on_device,
                      const sara_b200_sift_args* args, sara_b200_keypoint* keypoints, float* descriptors,
                      int capacity, int* n_out)
{
  int rc = sara_b200_sift_enqueue_u8(ctx, 0, image, w, h, channels, image_on_device, args, nullptr);
  if (rc)
    return rc;
  return sara_b200_collect(ctx, 0, keypoints, descriptors, capacity, n_out);
}

int sara_b200_dog_extrema(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                          int image_on_device, const sara_b200_dog_args* args)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  if (!image || !args)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "null argument");
  Slot& S = ctx->slots[slot];
  if (S.busy)
    return fail(ctx, SARA_B200_ERR_BUSY, "slot %d holds an un-collected frame", slot);
  if (args->pyramid_params.scale_count_per_octave < 4)
    return fail(ctx, SARA_B200_ERR_TOO_FEW_SCALES,
                "Error: The extraction of DoG extrema needs (1 + 3) = 4 scales per octave at the very "
                "minimum!");
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  cudaStream_t st = S.own_stream;
  if (args->img_padding_sz < 1)
    return fail(ctx, SARA_B200_ERR_BAD_ARG,
                "img_padding_sz must be >= 1 (the 3x3x3 neighbourhood must stay inside the layer)");
  const ExtremaParams early{args->extremum_thres, args->edge_ratio_thres, args->img_padding_sz,
                            args->extremum_refinement_iter};
  S.function_pyramid = false;
  rc = enqueue_pyramid(ctx, S, image, w, h, image_on_device != 0, args->pyramid_params,
                       args->gauss_truncate, st, &early);
  if (rc)
    return rc;
  rc = enqueue_extrema(ctx, S, args->extremum_thres, args->edge_ratio_thres, args->img_padding_sz,
                       args->extremum_refinement_iter, st);
  if (rc)
    return rc;
  S.has_keypoints = false;
  rc = finish_enqueue(ctx, S, st);
  if (rc)
    return rc;
  return wait_slot(ctx, S);
}

// ---- sibling detectors on the same pyramid ------------------------------------
// ComputeLoGExtrema (FeatureDetectors/LoG.cpp:20-58) and ComputeDoHExtrema
// (FeatureDetectors/Hessian.cpp:59-98): gaussian_pyramid (default truncation), a function pyramid
// with as many layers as the Gaussian one -- laplacian_pyramid (GaussianPyramid.hpp:156-178) or
// det_of_hessian_pyramid (Hessian.hpp:35-57) -- then local_scale_space_extrema on s = 1 .. N - 2
// with the very kernels of the DoG detector (classify, ordered compaction, refinement).
static int function_extrema(sara_b200_ctx* ctx, int slot, const float* image, int w, int h, int image_on_device,
                            int which, const sara_b200_dog_args* args)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  if (!image || !args)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "null argument");
  Slot& S = ctx->slots[slot];
  if (S.busy)
    return fail(ctx, SARA_B200_ERR_BUSY, "slot %d holds an un-collected frame", slot);
  if (args->pyramid_params.scale_count_per_octave < 3)
    return fail(ctx, SARA_B200_ERR_TOO_FEW_SCALES, "scale-space extrema need at least 3 scales per octave");
  if (args->pyramid_params.scale_count_per_octave + 1 > kMaxScales)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "too many scales per octave");
  if (args->img_padding_sz < 1)
    return fail(ctx, SARA_B200_ERR_BAD_ARG,
                "img_padding_sz must be >= 1 (the 3x3x3 neighbourhood must stay inside the layer)");
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  cudaStream_t st = S.own_stream;
  S.function_pyramid = false;
  rc = enqueue_pyramid(ctx, S, image, w, h, image_on_device != 0, args->pyramid_params, 4.f, st, nullptr);
  if (rc)
    return rc;
  // the function pyramid replaces the DoG in the D stack
  const PyramidDesc& P = S.P;
  float norm[kMaxScales];
  for (int s = 0; s < P.n_scales; ++s)
  {
    // float(square(scale_relative_to_octave(s))) / float(quartic(...)): powers taken in double
    const double sr = std::pow(static_cast<double>(args->pyramid_params.scale_geometric_factor), static_cast<double>(s)) *
                      static_cast<double>(args->pyramid_params.scale_initial);
    norm[s] = static_cast<float>(which == 1 ? sr * sr : sr * sr * sr * sr);
  }
  S.total_launches += launch_function_pyramid(P, which, norm, st);
  S.function_pyramid = true;

  // the same descriptor with one more "DoG" layer: N function layers, N - 2 of them searched
  PyramidDesc Pf = P;
  Pf.n_scales = P.n_scales + 1;
  const int n_fun = P.n_scales - 2;
  int seg_base = 0;
  int* rows0 = P.oct[0].row_count;
  for (int o = 0; o < P.n_octaves; ++o)
  {
    Pf.oct[o].row_count = rows0 + seg_base;
    Pf.oct[o].seg_base = seg_base;
    seg_base += n_fun * P.oct[o].h;
  }
  CU(ctx, cudaMemsetAsync(S.counters, 0, sizeof(Counters), st));
  if (P.n_octaves > 0 && seg_base > 0)
  {
    ExtremaParams ep{args->extremum_thres, args->edge_ratio_thres, args->img_padding_sz, args->extremum_refinement_iter};
    S.total_launches += launch_extrema(Pf, ep, seg_base, S.seg_offsets, S.cand, ctx->cap_cand, S.ext_tmp, 0, S.scratch,
                                       S.ext, ctx->cap_ext, S.counters, st);
  }
  if (ctx->profiling)
    CU(ctx, cudaEventRecord(S.ev[3], st));
  CU(ctx, cudaGetLastError());
  S.has_keypoints = false;
  rc = finish_enqueue(ctx, S, st);
  if (rc)
    return rc;
  return wait_slot(ctx, S);
}

int sara_b200_log_extrema(sara_b200_ctx* ctx, int slot, const float* image, int w, int h, int image_on_device,
                          const sara_b200_dog_args* args)
{
  return function_extrema(ctx, slot, image, w, h, image_on_device, 1, args);
}

int sara_b200_doh_extrema(sara_b200_ctx* ctx, int slot, const float* image, int w, int h, int image_on_device,
                          const sara_b200_dog_args* args)
{
  return function_extrema(ctx, slot, image, w, h, image_on_device, 2, args);
}

// ComputeHessianLaplaceMaxima::operator() (FeatureDetectors/Hessian.hpp:60-94, Hessian.cpp:19-57):
// gaussian_pyramid, det_of_hessian_pyramid, then laplace_maxima (RefineExtremum.cpp:659-709) on the layers
// s = 1 .. N - 1: spatial local maxima above the threshold, Laplace scale selection on a 13 x 13 patch
// (select_laplace_scale, RefineExtremum.cpp:523-657), 2-D sub-pixel refinement.  Only extremum_thres,
// img_padding_sz and extremum_refinement_iter of `args` are used.  Reference defaults: ImagePyramidParams(-1, 3 + 1),
// 1e-5, padding 1, 10 scales, 5 iterations.
static int laplace_detector(sara_b200_ctx* ctx, int slot, const float* image, int w, int h, int image_on_device,
                            const sara_b200_dog_args* args, int num_scales, bool harris, float kappa)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  if (!image || !args)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "null argument");
  Slot& S = ctx->slots[slot];
  if (S.busy)
    return fail(ctx, SARA_B200_ERR_BUSY, "slot %d holds an un-collected frame", slot);
  const sara_b200_pyramid_params& pp = args->pyramid_params;
  if (pp.scale_count_per_octave < 2 || pp.scale_count_per_octave + 2 > kMaxScales)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "Laplace detectors: scale_count_per_octave outside [2, %d]", kMaxScales - 2);
  if (num_scales < 2 || num_scales > kLaplaceMaxScales)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "Hessian-Laplace: num_scales outside [2, %d]", kLaplaceMaxScales);
  if (args->img_padding_sz < 1)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "img_padding_sz must be >= 1 (the 3 x 3 neighbourhood must stay inside the layer)");
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  cudaStream_t st = S.own_stream;
  S.function_pyramid = false;
  rc = enqueue_pyramid(ctx, S, image, w, h, image_on_device != 0, pp, 4.f, st, nullptr);
  if (rc)
    return rc;
  const PyramidDesc& P = S.P;
  float norm[kMaxScales];
  auto scale_rel = [&](int s) {  // ImagePyramid::scale_relative_to_octave: a double
    return std::pow(static_cast<double>(pp.scale_geometric_factor), static_cast<double>(s)) *
           static_cast<double>(pp.scale_initial);
  };
  if (!harris)
  {
    for (int s = 0; s < P.n_scales; ++s)
    {
      const double sr = scale_rel(s);
      norm[s] = static_cast<float>(sr * sr * sr * sr);
    }
    S.total_launches += launch_function_pyramid(P, 2, norm, st);
  }
  else
  {
    // ComputeHarrisLaplaceCorners (Harris.cpp:171-193): per layer Gradient -> SecondMomentMatrix ->
    // Gaussian(sigma_I) of every coefficient -> det - kappa trace^2 -> times float(sigma_D^2)
    const float scale_factor = 1 / std::sqrt(2.f);
    const size_t layer0 = align_up(sizeof(float) * P.oct[0].pitch * P.oct[0].h, 256);
    rc = grow_scratch(ctx, 6 * layer0);
    if (rc)
      return rc;
    for (int o = 0; o < P.n_octaves; ++o)
    {
      const OctaveDesc& oc = P.oct[o];
      const size_t layer = static_cast<size_t>(oc.pitch) * oc.h;
      float* m[6];
      for (int c = 0; c < 6; ++c)
        m[c] = ctx->scratch + c * (layer0 / sizeof(float));
      for (int s = 0; s < P.n_scales; ++s)
      {
        const float sigma_I = static_cast<float>(scale_rel(s));
        const float sigma_D = sigma_I * scale_factor;
        Taps taps;
        if (make_taps(sigma_I, 4.f, &taps) < 0)
          return fail(ctx, SARA_B200_ERR_BAD_ARG, "Harris-Laplace: the integration blur exceeds %d taps", kMaxTaps);
        launch_second_moment(oc.G + s * layer, oc.w, oc.h, oc.pitch, m[0], m[1], m[2], st);
        for (int c = 0; c < 3; ++c)
          launch_gaussian_stage(m[c], oc.pitch, m[3 + c], oc.pitch, nullptr, 0, oc.w, oc.h, taps, st);
        launch_cornerness(m[3], m[4], m[5], oc.w, oc.h, oc.pitch, kappa, static_cast<float>(sigma_D * sigma_D),
                          oc.D + s * layer, st);
        S.total_launches += 5;
      }
    }
  }
  S.function_pyramid = true;

  // the constants of select_laplace_scale, per layer (RefineExtremum.cpp:559-603), and the blur taps
  static thread_local LaplaceTable T;
  std::memset(&T, 0, sizeof T);
  T.num_scales = num_scales;
  T.ratio = std::pow(2.f, 1.f / num_scales);
  for (int s = 1; s < P.n_scales; ++s)
  {
    const double nearest_sigma = scale_rel(s - 1);
    T.scales[s][0] = static_cast<float>(scale_rel(s)) / std::sqrt(2.f);
    const double inc0 = std::sqrt(static_cast<double>(T.scales[s][0] * T.scales[s][0]) - nearest_sigma * nearest_sigma);
    float inc[kLaplaceMaxScales + 1];
    inc[0] = inc0 > 1e-3f ? static_cast<float>(inc0) : 0.f;  // NaN compares false: no blur
    for (int i = 1; i <= num_scales; ++i)
    {
      T.scales[s][i] = T.ratio * T.scales[s][i - 1];
      inc[i] = std::sqrt(T.scales[s][i] * T.scales[s][i] - T.scales[s][i - 1] * T.scales[s][i - 1]);
    }
    for (int i = 0; i <= num_scales; ++i)
    {
      if (!(inc[i] > 0.f))
        continue;
      Taps taps;
      if (make_taps(inc[i], 4.f, &taps) < 0 || taps.n > kLaplaceMaxTaps)
        return fail(ctx, SARA_B200_ERR_BAD_ARG, "Hessian-Laplace: a scale-selection blur needs more than %d taps",
                    kLaplaceMaxTaps);
      T.n_taps[s][i] = taps.n;
      for (int j = 0; j < taps.n; ++j)
        T.taps[s][i][j] = taps.v[j];
    }
  }
  if (!ctx->d_laplace)
    CU(ctx, cudaMalloc(&ctx->d_laplace, sizeof(LaplaceTable)));
  CU(ctx, cudaMemcpyAsync(ctx->d_laplace, &T, sizeof(LaplaceTable), cudaMemcpyHostToDevice, st));
  CU(ctx, cudaStreamSynchronize(st));  // T is reused by the next call of this thread

  // descriptor arranged for N - 1 searched layers: "n_scales - 3" of the extrema kernels = N - 1
  PyramidDesc Pf = P;
  Pf.n_scales = P.n_scales + 2;
  const int n_fun = P.n_scales - 1;
  int seg_base = 0;
  int* rows0 = P.oct[0].row_count;
  for (int o = 0; o < P.n_octaves; ++o)
  {
    Pf.oct[o].row_count = rows0 + seg_base;
    Pf.oct[o].seg_base = seg_base;
    seg_base += n_fun * P.oct[o].h;
  }
  CU(ctx, cudaMemsetAsync(S.counters, 0, sizeof(Counters), st));
  if (P.n_octaves > 0 && seg_base > 0)
    S.total_launches += launch_laplace_maxima(Pf, ctx->d_laplace, args->extremum_thres, args->img_padding_sz,
                                              args->extremum_refinement_iter, seg_base, S.seg_offsets, S.cand,
                                              ctx->cap_cand, S.ext_tmp, S.scratch, S.ext, ctx->cap_ext, S.counters, st);
  if (ctx->profiling)
    CU(ctx, cudaEventRecord(S.ev[3], st));
  CU(ctx, cudaGetLastError());
  S.has_keypoints = false;
  rc = finish_enqueue(ctx, S, st);
  if (rc)
    return rc;
  return wait_slot(ctx, S);
}

int sara_b200_hessian_laplace(sara_b200_ctx* ctx, int slot, const float* image, int w, int h, int image_on_device,
                              const sara_b200_dog_args* args, int num_scales)
{
  return laplace_detector(ctx, slot, image, w, h, image_on_device, args, num_scales, false, 0.f);
}

// ComputeHarrisLaplaceCorners::operator() (FeatureDetectors/Harris.hpp:125-138, Harris.cpp:165-230): the Harris
// cornerness of every Gaussian layer (integration scale sigma_I = the layer's scale, differentiation scale
// sigma_I / sqrt(2)), then laplace_maxima on s = 1 .. N - 1.  Reference defaults: ImagePyramidParams(-1, 2 + 1,
// sqrt(2), 1), kappa 0.04, 1e-6, padding 1, 10 scales, 5 iterations.
int sara_b200_harris_laplace(sara_b200_ctx* ctx, int slot, const float* image, int w, int h, int image_on_device,
                             const sara_b200_dog_args* args, float kappa, int num_scales)
{
  return laplace_detector(ctx, slot, image, w, h, image_on_device, args, num_scales, true, kappa);
}

int sara_b200_pyramid_enqueue(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                              int image_on_device, const sara_b200_pyramid_params* params,
                              float gauss_truncate, void* stream)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  if (!image || !params)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "null argument");
  Slot& S = ctx->slots[slot];
  if (S.busy)
    return fail(ctx, SARA_B200_ERR_BUSY, "slot %d holds an un-collected frame", slot);
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : S.own_stream;
  S.function_pyramid = false;
  rc = enqueue_pyramid(ctx, S, image, w, h, image_on_device != 0, *params, gauss_truncate, st);
  if (rc)
    return rc;
  S.has_keypoints = false;
  std::memset(S.h_counters, 0, sizeof(Counters));
  S.busy = true;
  return 0;
}

int sara_b200_num_octaves(sara_b200_ctx* ctx, int slot)
{
  return check_slot(ctx, slot) ? -1 : ctx->slots[slot].P.n_octaves;
}

int sara_b200_num_scales(sara_b200_ctx* ctx, int slot)
{
  return check_slot(ctx, slot) ? -1 : ctx->slots[slot].P.n_scales;
}

int sara_b200_layer_size(sara_b200_ctx* ctx, int slot, int octave, int* w, int* h)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  const PyramidDesc& P = ctx->slots[slot].P;
  if (octave < 0 || octave >= P.n_octaves)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "octave %d out of range [0, %d)", octave, P.n_octaves);
  if (w)
    *w = P.oct[octave].w;
  if (h)
    *h = P.oct[octave].h;
  return 0;
}

float sara_b200_octave_scaling_factor(sara_b200_ctx* ctx, int slot, int octave)
{
  if (check_slot(ctx, slot))
    return 0.f;
  const PyramidDesc& P = ctx->slots[slot].P;
  if (octave < 0 || octave >= P.n_octaves)
    return 0.f;
  return P.oct[octave].scaling;
}

int sara_b200_copy_layer(sara_b200_ctx* ctx, int slot, int which, int s, int o, float* dst)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  Slot& S = ctx->slots[slot];
  const PyramidDesc& P = S.P;
  if (!dst || o < 0 || o >= P.n_octaves || s < 0 || s >= (which == 0 || ctx->slots[slot].function_pyramid ? P.n_scales : P.n_scales - 1) ||
      (which != 0 && which != 1))
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "copy_layer: bad (which=%d, s=%d, o=%d)", which, s, o);  // std::out_of_range in ImagePyramid
  rc = wait_slot(ctx, S);
  if (rc)
    return rc;
  const OctaveDesc& oc = P.oct[o];
  const float* src = (which == 0 ? oc.G : oc.D) + static_cast<size_t>(s) * oc.layer_stride;
  CU(ctx, cudaMemcpy2D(dst, sizeof(float) * oc.w, src, sizeof(float) * oc.pitch, sizeof(float) * oc.w, oc.h,
                       cudaMemcpyDeviceToHost));
  return 0;
}

int sara_b200_copy_extrema(sara_b200_ctx* ctx, int slot, sara_b200_keypoint* dst, int capacity, int* n_out)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  Slot& S = ctx->slots[slot];
  rc = wait_slot(ctx, S);
  if (rc)
    return rc;
  return copy_keypoints(ctx, S, S.ext, S.h_counters->n_ext, ctx->cap_ext, dst, capacity, n_out);
}

int sara_b200_copy_oriented(sara_b200_ctx* ctx, int slot, sara_b200_keypoint* dst, int capacity, int* n_out)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  Slot& S = ctx->slots[slot];
  rc = wait_slot(ctx, S);
  if (rc)
    return rc;
  const int n = S.has_keypoints ? S.h_counters->n_kp : 0;
  return copy_keypoints(ctx, S, S.kp_oct, n, ctx->cap_kp, dst, capacity, n_out);
}

// ---- stage functors on caller-supplied extrema ---------------------------------------------
// ComputeDominantOrientations::operator() (FeatureDescriptors/Orientation.hpp:233-242,
// Orientation.cpp:135-166) followed -- when `descriptors` is asked for -- by
// ComputeSIFTDescriptor<4, 8>::operator() (FeatureDescriptors/SIFT.hpp:62-166) and the rescale of
// SIFT.cpp:92-98, on the Gaussian pyramid the slot holds (after sara_b200_dog_extrema,
// sara_b200_pyramid_enqueue + wait, or a whole sift call).
int sara_b200_describe_extrema(sara_b200_ctx* ctx, int slot, const sara_b200_keypoint* extrema, int n,
                               sara_b200_keypoint* oriented, sara_b200_keypoint* keypoints, float* descriptors,
                               int capacity, int* n_out)
{
  int rc = check_slot(ctx, slot);
  if (rc)
    return rc;
  Slot& S = ctx->slots[slot];
  rc = wait_slot(ctx, S);
  if (rc)
    return rc;
  if (n_out)
    *n_out = 0;
  if (n < 0 || (n > 0 && !extrema) || capacity < 0)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "describe_extrema: bad arguments");
  if (S.P.n_octaves <= 0)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "describe_extrema: the slot holds no Gaussian pyramid");
  if (n > ctx->cap_ext)
    return fail(ctx, SARA_B200_ERR_OVERFLOW, "%d extrema exceed the context capacity %d", n, ctx->cap_ext);
  for (int i = 0; i < n; ++i)  // the (s, o) pair of every extremum (scale_octave_pairs) must address a layer
    if (extrema[i].o < 0 || extrema[i].o >= S.P.n_octaves || extrema[i].s < 0 || extrema[i].s >= S.P.n_scales)
      return fail(ctx, SARA_B200_ERR_BAD_ARG, "describe_extrema: extremum %d has (s, o) = (%d, %d) outside the pyramid",
                  i, extrema[i].s, extrema[i].o);
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  cudaStream_t st = S.own_stream;
  S.stream = st;
  Counters zero{};
  zero.n_ext = n;
  CU(ctx, cudaMemcpyAsync(S.counters, &zero, sizeof(Counters), cudaMemcpyHostToDevice, st));
  if (n > 0)
    CU(ctx, cudaMemcpyAsync(S.ext, extrema, sizeof(Keypoint) * n, cudaMemcpyHostToDevice, st));
  const bool prof = ctx->profiling;
  ctx->profiling = false;  // the stage events of a whole frame do not apply here
  rc = enqueue_describe(ctx, S, st);
  ctx->profiling = prof;
  if (rc)
    return rc;
  S.has_keypoints = true;
  rc = finish_enqueue(ctx, S, st);
  if (rc)
    return rc;
  CU(ctx, cudaStreamSynchronize(st));
  S.busy = false;
  const Counters c = *S.h_counters;
  if (n_out)
    *n_out = c.n_kp;
  if (c.overflow || c.n_kp > ctx->cap_kp)
    return fail(ctx, SARA_B200_ERR_OVERFLOW, "%d oriented keypoints exceed the context capacity %d", c.n_kp, ctx->cap_kp);
  if (c.n_kp > capacity)
    return fail(ctx, SARA_B200_ERR_OVERFLOW, "%d keypoints exceed the caller's capacity %d", c.n_kp, capacity);
  if (c.n_kp > 0)
  {
    if (oriented)
      CU(ctx, cudaMemcpyAsync(oriented, S.kp_oct, sizeof(Keypoint) * c.n_kp, cudaMemcpyDeviceToHost, st));
    if (keypoints)
      CU(ctx, cudaMemcpyAsync(keypoints, S.kp_out, sizeof(Keypoint) * c.n_kp, cudaMemcpyDeviceToHost, st));
    if (descriptors)
      CU(ctx, cudaMemcpyAsync(descriptors, S.desc, sizeof(float) * 128 * c.n_kp, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
  }
  return 0;
}

int sara_b200_gaussian(sara_b200_ctx* ctx, const float* src, int w, int h, float sigma, float gauss_truncate,
                       float* dst)
{
  if (!ctx)
    return fail(nullptr, SARA_B200_ERR_BAD_ARG, "null context");
  if (!src || !dst || w <= 0 || h <= 0)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "gaussian: bad arguments");
  Taps taps;
  if (make_taps(sigma, gauss_truncate, &taps) < 0)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "Gaussian kernel exceeds %d taps", kMaxTaps);
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  const size_t bytes = sizeof(float) * w * h;
  int rc = grow_scratch(ctx, 2 * align_up(bytes, 256));
  if (rc)
    return rc;
  float* d_a = ctx->scratch;
  float* d_b = reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->scratch) + align_up(bytes, 256));
  cudaStream_t st = ctx->slots[0].own_stream;
  CU(ctx, cudaMemcpyAsync(d_a, src, bytes, cudaMemcpyHostToDevice, st));
  launch_gaussian_stage(d_a, w, d_b, w, nullptr, 0, w, h, taps, st);
  CU(ctx, cudaMemcpyAsync(dst, d_b, bytes, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  return 0;
}

int sara_b200_to_gray32f(sara_b200_ctx* ctx, const uint8_t* src, int w, int h, int channels, float* dst)
{
  if (!ctx)
    return fail(nullptr, SARA_B200_ERR_BAD_ARG, "null context");
  if (!src || !dst || w <= 0 || h <= 0 || (channels != 1 && channels != 3))
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "to_gray32f: bad arguments");
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  const size_t n = static_cast<size_t>(w) * h;
  const size_t in_bytes = align_up(n * channels, 256);
  int rc = grow_scratch(ctx, in_bytes + sizeof(float) * n);
  if (rc)
    return rc;
  uint8_t* d_in = reinterpret_cast<uint8_t*>(ctx->scratch);
  float* d_out = reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->scratch) + in_bytes);
  cudaStream_t st = ctx->slots[0].own_stream;
  CU(ctx, cudaMemcpyAsync(d_in, src, n * channels, cudaMemcpyHostToDevice, st));
  if (channels == 3)
    launch_rgb8_to_gray32f(d_in, d_out, static_cast<int>(n), ctx->d_gray_lut, st);
  else
    launch_gray8_to_gray32f(d_in, d_out, static_cast<int>(n), st);
  CU(ctx, cudaMemcpyAsync(dst, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  return 0;
}

int sara_b200_make_gaussian_kernel(float sigma, float gauss_truncate, float* taps, int capacity)
{
  Taps t;
  const int n = make_taps(sigma, gauss_truncate, &t);
  if (n < 0)
    return n;
  if (n > capacity || !taps)
    return -n;
  std::memcpy(taps, t.v, sizeof(float) * n);
  return n;
}

}  // extern "C"

// ---- descriptor matching: AnnMatcher over the device search of match.cu ------
namespace {

  int grow_match_io(sara_b200_ctx* ctx, size_t bytes)
  {
    if (bytes <= ctx->match_io_bytes)
      return 0;
    if (ctx->match_io)
    {
      CU(ctx, cudaStreamSynchronize(ctx->slots[0].own_stream));
      CU(ctx, cudaFree(ctx->match_io));
      ctx->match_io = nullptr;
      ctx->match_io_bytes = 0;
    }
    CU(ctx, cudaMalloc(&ctx->match_io, bytes));
    ctx->match_io_bytes = bytes;
    return 0;
  }

  struct KnnResult
  {
    std::vector<int32_t> idx;
    std::vector<float> dist;
  };

  // One search, device pointers in, host vectors out (k entries per query).
  int knn_to_host(sara_b200_ctx* ctx, const float* d_q, int nq, const float* d_data, int nd, int dim, int k, int mode,
                  int32_t* d_idx, float* d_dist, int32_t* h_idx, float* h_dist, sara_b200_knn_stats* stats,
                  cudaStream_t st)
  {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (stats)
    {
      CU(ctx, cudaEventCreate(&e0));
      CU(ctx, cudaEventCreate(&e1));
      CU(ctx, cudaEventRecord(e0, st));
    }
    sb::match::KnnStats ks;
    const int rc = sb::match::knn(ctx->match_ws, d_q, nq, d_data, nd, dim, k, mode, d_idx, d_dist, &ks, st, ctx->err,
                                  sizeof ctx->err);
    if (rc)
      return rc;
    if (stats)
      CU(ctx, cudaEventRecord(e1, st));
    CU(ctx, cudaMemcpyAsync(h_idx, d_idx, sizeof(int32_t) * nq * k, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(h_dist, d_dist, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    if (stats)
    {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      stats->used_tensor_cores |= ks.used_tensor_cores;
      stats->n_redone += ks.n_redone;
      stats->launches += ks.launches;
      stats->splits = ks.splits;
      stats->gpu_ms += ms;
    }
    return 0;
  }

  template <class T>
  T* carve_io(unsigned char*& cur, size_t n)
  {
    T* p = reinterpret_cast<T*>(cur);
    cur += align_up(n * sizeof(T), 256);
    return p;
  }

}  // namespace

void sara_b200_default_match_args(sara_b200_match_args* a)
{
  if (!a)
    return;
  a->sift_ratio_thres = 1.2f;  // AnnMatcher.hpp:41
  a->self_matching = 0;
  a->min_max_metric_dist_thres = 0.5f;  // AnnMatcher.hpp:45-46
  a->pixel_dist_thres = 10.f;
  a->knn_mode = SARA_B200_KNN_AUTO;
}

int sara_b200_knn(sara_b200_ctx* ctx, const float* queries, int nq, const float* data, int nd, int dim, int k,
                  int on_device, int knn_mode, int32_t* idx, float* dist, sara_b200_knn_stats* stats)
{
  if (!ctx)
    return fail(nullptr, SARA_B200_ERR_BAD_ARG, "null context");
  if (!queries || !data || !idx || !dist || nq < 0 || nd < 0 || dim < 1 || k < 1)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "knn: bad arguments");
  if (stats)
    memset(stats, 0, sizeof *stats);
  if (nq == 0)
    return 0;
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  cudaStream_t st = ctx->slots[0].own_stream;
  const size_t qb = sizeof(float) * nq * dim, db = sizeof(float) * nd * dim;
  int rc = grow_match_io(ctx, (on_device ? 0 : align_up(qb, 256) + align_up(db, 256)) + 2 * align_up(sizeof(float) * nq * k, 256));
  if (rc)
    return rc;
  unsigned char* cur = ctx->match_io;
  const float* d_q = queries;
  const float* d_d = data;
  if (!on_device)
  {
    float* q = carve_io<float>(cur, static_cast<size_t>(nq) * dim);
    float* d = carve_io<float>(cur, static_cast<size_t>(nd) * dim);
    CU(ctx, cudaMemcpyAsync(q, queries, qb, cudaMemcpyHostToDevice, st));
    if (nd)
      CU(ctx, cudaMemcpyAsync(d, data, db, cudaMemcpyHostToDevice, st));
    d_q = q;
    d_d = d;
  }
  int32_t* d_idx = carve_io<int32_t>(cur, static_cast<size_t>(nq) * k);
  float* d_dist = carve_io<float>(cur, static_cast<size_t>(nq) * k);
  return knn_to_host(ctx, d_q, nq, d_d, nd, dim, k, knn_mode, d_idx, d_dist, idx, dist, stats, st);
}

int sara_b200_compute_matches(sara_b200_ctx* ctx, const float* desc1, const sara_b200_keypoint* kp1, int n1, const float* desc2,
                    const sara_b200_keypoint* kp2, int n2, int dim, int on_device, const sara_b200_match_args* args,
                    sara_b200_match* out, int capacity, int* n_out, sara_b200_knn_stats* stats)
{
  if (!ctx)
    return fail(nullptr, SARA_B200_ERR_BAD_ARG, "null context");
  if (!args || !n_out || capacity < 0 || (capacity > 0 && !out) || n1 < 0 || n2 < 0 || dim < 1)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "match: bad arguments");
  *n_out = 0;
  if (stats)
    memset(stats, 0, sizeof *stats);
  if (n1 == 0 || n2 == 0 || !desc1 || !desc2)
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "Error: the list of key-points is empty!");  // AnnMatcher.cpp:45-46
  const bool self = args->self_matching != 0;
  if (self && (!kp1 || !kp2))
    return fail(ctx, SARA_B200_ERR_BAD_ARG, "match: self matching needs the features (KeyProximity)");
  DeviceGuard guard(ctx->device);
  CU(ctx, guard.err);
  cudaStream_t st = ctx->slots[0].own_stream;

  const float sq_thres = args->sift_ratio_thres * args->sift_ratio_thres;
  const float sq_metric = args->min_max_metric_dist_thres * args->min_max_metric_dist_thres;
  const float sq_pixel = args->pixel_dist_thres * args->pixel_dist_thres;
  const int K = 3;

  // device staging: descriptors (when they come from the host), k-NN outputs, radii / counts / offsets
  const size_t b1 = sizeof(float) * n1 * dim, b2 = sizeof(float) * n2 * dim;
  const int nmax = std::max(n1, n2);
  size_t need = (on_device ? 0 : align_up(b1, 256) + align_up(b2, 256)) + 2 * align_up(sizeof(float) * nmax * K, 256) +
                3 * align_up(sizeof(float) * nmax, 256);
  int rc = grow_match_io(ctx, need);
  if (rc)
    return rc;
  unsigned char* cur = ctx->match_io;
  const float* d1 = desc1;
  const float* d2 = desc2;
  if (!on_device)
  {
    float* p1 = carve_io<float>(cur, static_cast<size_t>(n1) * dim);
    float* p2 = carve_io<float>(cur, static_cast<size_t>(n2) * dim);
    CU(ctx, cudaMemcpyAsync(p1, desc1, b1, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(p2, desc2, b2, cudaMemcpyHostToDevice, st));
    d1 = p1;
    d2 = p2;
  }
  int32_t* d_idx = carve_io<int32_t>(cur, static_cast<size_t>(nmax) * K);
  float* d_dist = carve_io<float>(cur, static_cast<size_t>(nmax) * K);
  float* d_radius = carve_io<float>(cur, nmax);
  int* d_count = carve_io<int>(cur, nmax);
  int* d_off = carve_io<int>(cur, nmax);

  struct Side
  {
    const float* d;
    const sara_b200_keypoint* kp;
    int n;
  };
  const Side sides[2] = {{d1, kp1, n1}, {d2, kp2, n2}};
  std::vector<sara_b200_match> matches;
  matches.reserve(100000);

  auto squared_ref_distance = [](const float* M, float ax, float ay, float bx, float by) {  // Metric.hpp:46-49
    const float dx = bx - ax, dy = by - ay;
    const float mx = M[0] * dx + M[2] * dy, my = M[1] * dx + M[3] * dy;
    return dx * mx + dy * my;
  };
  auto is_redundant = [&](const sara_b200_keypoint& f1, const sara_b200_keypoint& f2) {  // KeyProximity.cpp:17-30
    const float sd1 = squared_ref_distance(f1.shape, f1.x, f1.y, f2.x, f2.y);
    const float sd2 = squared_ref_distance(f2.shape, f1.x, f1.y, f2.x, f2.y);
    const float dx = f1.x - f2.x, dy = f1.y - f2.y;
    return dx * dx + dy * dy < sq_pixel || sd1 < sq_metric || sd2 < sq_metric;
  };

  // both directions of AnnMatcher.cpp:242-254; `one` queries, `two` is the indexed set
  for (int dir = 0; dir < 2; ++dir)
  {
    const Side& one = sides[dir];
    const Side& two = sides[1 - dir];
    auto push = [&](int i1, int i2, float score, int rank) {
      sara_b200_match m;
      m.x_index = dir == 0 ? i1 : i2;
      m.y_index = dir == 0 ? i2 : i1;
      m.rank = rank;
      m.score = score;
      m.direction = dir;
      matches.push_back(m);
    };
    if (two.n == 1 && !self)  // boundary case 2 (AnnMatcher.cpp:88-103)
    {
      if (1.f < sq_thres)
        for (int i1 = 0; i1 < one.n; ++i1)
          push(i1, 0, 1.f, 1);
      continue;
    }
    std::vector<int32_t> idx(static_cast<size_t>(one.n) * K);
    std::vector<float> dist(static_cast<size_t>(one.n) * K);
    rc = knn_to_host(ctx, one.d, one.n, two.d, two.n, dim, K, args->knn_mode, d_idx, d_dist, idx.data(), dist.data(),
                     stats, st);
    if (rc)
      return rc;
    if (two.n == 2 && self)  // boundary case 3 (AnnMatcher.cpp:106-125)
    {
      if (1.f < sq_thres)
        for (int i1 = 0; i1 < one.n; ++i1)
          push(i1, idx[static_cast<size_t>(i1) * K + 1], 1.f, 1);
      continue;
    }
    const int top1 = self ? 1 : 0;

    // adaptive radius search (AnnMatcher.cpp:139-146): every neighbour with dist < d(top1) * ratio^2
    std::vector<int> r_off, r_idx;
    std::vector<float> r_dist;
    if (sq_thres > 1.f)
    {
      std::vector<float> radius(one.n);
      for (int i1 = 0; i1 < one.n; ++i1)
        radius[i1] = dist[static_cast<size_t>(i1) * K + top1] * sq_thres;
      CU(ctx, cudaMemcpyAsync(d_radius, radius.data(), sizeof(float) * one.n, cudaMemcpyHostToDevice, st));
      rc = sb::match::radius_pass(one.d, one.n, two.d, two.n, dim, d_radius, d_count, nullptr, nullptr, nullptr, st,
                                  ctx->err, sizeof ctx->err);
      if (rc)
        return rc;
      std::vector<int> count(one.n);
      CU(ctx, cudaMemcpyAsync(count.data(), d_count, sizeof(int) * one.n, cudaMemcpyDeviceToHost, st));
      CU(ctx, cudaStreamSynchronize(st));
      r_off.assign(one.n + 1, 0);
      size_t total = 0;
      for (int i1 = 0; i1 < one.n; ++i1)
      {
        r_off[i1] = static_cast<int>(total);
        total += count[i1];
        if (total > (size_t(1) << 27))
          return fail(ctx, SARA_B200_ERR_OVERFLOW, "match: more than 2^27 neighbours inside the adaptive radii");
      }
      r_off[one.n] = static_cast<int>(total);
      r_idx.resize(total);
      r_dist.resize(total);
      if (total)
      {
        int* d_ridx = nullptr;
        float* d_rdist = nullptr;
        CU(ctx, cudaMalloc(&d_ridx, sizeof(int) * total));
        if (cudaMalloc(&d_rdist, sizeof(float) * total) != cudaSuccess)
        {
          cudaFree(d_ridx);
          return fail(ctx, SARA_B200_ERR_OOM, "match: cudaMalloc of the radius results failed");
        }
        cudaMemcpyAsync(d_off, r_off.data(), sizeof(int) * one.n, cudaMemcpyHostToDevice, st);
        rc = sb::match::radius_pass(one.d, one.n, two.d, two.n, dim, d_radius, d_count, d_off, d_ridx, d_rdist, st,
                                    ctx->err, sizeof ctx->err);
        cudaMemcpyAsync(r_idx.data(), d_ridx, sizeof(int) * total, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(r_dist.data(), d_rdist, sizeof(float) * total, cudaMemcpyDeviceToHost, st);
        const cudaError_t e = cudaStreamSynchronize(st);
        cudaFree(d_ridx);
        cudaFree(d_rdist);
        if (rc)
          return rc;
        if (e != cudaSuccess)
          return fail(ctx, SARA_B200_ERR_CUDA, "match: radius search failed: %s", cudaGetErrorString(e));
      }
    }

    std::vector<std::pair<float, int>> seg;
    for (int i1 = 0; i1 < one.n; ++i1)
    {
      const int32_t* I = idx.data() + static_cast<size_t>(i1) * K;
      const float* D = dist.data() + static_cast<size_t>(i1) * K;
      const float top1_score = D[top1 + 1] > 0.f ? D[top1] / D[top1 + 1] : 0.f;
      if (!(sq_thres > 1.f))
      {
        // K = 1: the loop of AnnMatcher.cpp:149-170 runs for rank = top1 < 1 only, i.e. never when
        // self matching (top1 = 1): self matches exist only on the radius branch.
        if (top1 != 0 || top1_score > sq_thres)
          continue;
        push(i1, I[0], top1_score, 1);
        continue;
      }
      // RadiusResultSet::copy(sorted): by (dist, index)
      seg.clear();
      for (int e = r_off[i1]; e < r_off[i1 + 1]; ++e)
        seg.emplace_back(r_dist[e], r_idx[e]);
      std::sort(seg.begin(), seg.end());
      const int Kn = static_cast<int>(seg.size());
      for (int rank = top1; rank < Kn; ++rank)
      {
        float score = 0.f;
        if (rank == top1)
          score = top1_score;
        else if (seg[top1].first)
          score = seg[rank].first / seg[top1].first;
        if (score > sq_thres)
          break;
        const int i2 = seg[rank].second;
        if (self && is_redundant(one.kp[i1], two.kp[i2]))
          continue;
        push(i1, i2, score, top1 == 0 ? rank + 1 : rank);
      }
    }
  }

  // AnnMatcher.cpp:256-276: lexicographic sort, unique, sort by score (ties keep the lexicographic order)
  std::sort(matches.begin(), matches.end(), [](const sara_b200_match& a, const sara_b200_match& b) {
    if (a.x_index != b.x_index)
      return a.x_index < b.x_index;
    if (a.y_index != b.y_index)
      return a.y_index < b.y_index;
    return a.score < b.score;
  });
  auto same_feature = [](const sara_b200_keypoint& a, const sara_b200_keypoint& b) {  // Feature.hpp:140-146
    return a.x == b.x && a.y == b.y && a.shape[0] == b.shape[0] && a.shape[1] == b.shape[1] && a.shape[2] == b.shape[2] &&
           a.shape[3] == b.shape[3] && a.orientation == b.orientation && a.type == b.type;
  };
  auto equal = [&](const sara_b200_match& a, const sara_b200_match& b) {  // Match.hpp:159-162
    if (kp1 && kp2)
      return same_feature(kp1[a.x_index], kp1[b.x_index]) && same_feature(kp2[a.y_index], kp2[b.y_index]);
    return a.x_index == b.x_index && a.y_index == b.y_index;
  };
  matches.resize(std::unique(matches.begin(), matches.end(), equal) - matches.begin());
  std::stable_sort(matches.begin(), matches.end(),
                   [](const sara_b200_match& a, const sara_b200_match& b) { return a.score < b.score; });

  *n_out = static_cast<int>(matches.size());
  const int n_copy = std::min<int>(*n_out, capacity);
  if (n_copy)
    memcpy(out, matches.data(), sizeof(sara_b200_match) * n_copy);
  if (*n_out > capacity)
    return fail(ctx, SARA_B200_ERR_OVERFLOW, "match: %d matches, capacity %d", *n_out, capacity);
  return 0;
}
