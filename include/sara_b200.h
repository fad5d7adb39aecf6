/* ========================================================================== *
 * sara_b200.h -- C ABI of the B200-native SIFT path.
 *
 * Drop-in boundary for ONE path of oddkiva/sara: per-frame SIFT
 *   Gaussian pyramid -> DoG -> scale-space extrema (+refinement, edge
 *   rejection) -> dominant orientations -> 128-D descriptors.
 * The reference has no FFI for this path; its boundary is the C++ signature
 *   DO::Sara::compute_sift_keypoints      cpp/src/DO/Sara/FeatureDetectors/SIFT.hpp:24-33
 * (bound to Python at python/oddkiva/sara/pybind11/FeatureDetectors.cpp:116-124)
 * and the stage functor
 *   DO::Sara::ComputeDoGExtrema           cpp/src/DO/Sara/FeatureDetectors/DoG.hpp:72-165.
 * Every entry point below names the reference interface it replaces.  Plain
 * pointers and sizes only; nothing here throws; all functions return 0 on
 * success or a negative sara_b200_status and leave a message retrievable with
 * sara_b200_last_error().
 *
 * Threading contract: one ctx per (host thread, GPU).  Calls on one ctx are
 * serialised by the caller; different ctxs are independent.  A ctx owns
 * `num_slots` independent frame slots (device arena + stream each) so that
 * several frames can be in flight: enqueue on slot i, collect slot i later.
 * A ctx with four or more slots cuts the small pyramid layers into fewer,
 * taller segments (less machine time per frame, a slightly longer lone frame);
 * the results are bit-identical.
 *
 * There is NO CPU fallback: if no CUDA device is usable sara_b200_create fails
 * with SARA_B200_ERR_CUDA.
 * ========================================================================== */
#ifndef SARA_B200_H
#define SARA_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#  define SARA_B200_API __attribute__((visibility("default")))
#else
#  define SARA_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SARA_B200_VERSION 100 /* 0.1.0 */

typedef enum sara_b200_status
{
  SARA_B200_OK = 0,
  SARA_B200_ERR_BAD_ARG = -1,         /* std::domain_error / std::range_error in the reference */
  SARA_B200_ERR_TOO_FEW_SCALES = -2,  /* std::runtime_error, DoG.hpp:86-89 */
  SARA_B200_ERR_CUDA = -3,            /* SHAKTI_SAFE_CUDA_CALL would throw */
  SARA_B200_ERR_OOM = -4,
  SARA_B200_ERR_OVERFLOW = -5,        /* more keypoints than the ctx / caller buffers hold */
  SARA_B200_ERR_BUSY = -6             /* slot has an un-collected frame */
} sara_b200_status;

/* ImagePyramidParams, cpp/src/DO/Sara/ImageProcessing/ImagePyramid.hpp:33-49
 * (same fields, same order, same defaults via sara_b200_default_pyramid_params). */
typedef struct sara_b200_pyramid_params
{
  int32_t first_octave_index;     /* -1 */
  int32_t scale_count_per_octave; /* 3 + 3 */
  float scale_geometric_factor;   /* 2^(1/3) */
  int32_t image_padding_size;     /* 1 */
  float scale_camera;             /* 0.5 */
  float scale_initial;            /* 1.6 */
  int32_t num_octaves_max;        /* INT_MAX */
} sara_b200_pyramid_params;

/* One keypoint.  The first 34 bytes carry the fields of DO::Sara::OERegion
 * (cpp/src/DO/Sara/Features/Feature.hpp:40-179) in declaration order; (s, o)
 * is the entry of `scale_octave_pairs` (DoG.cpp:70-82); (xi, yi) is the raster
 * slot of the DoG layer the extremum was emitted at (RefineExtremum.cpp:495-515). */
typedef struct sara_b200_keypoint
{
  float x, y;            /* OERegion::coords */
  float shape[4];        /* OERegion::shape_matrix, column-major (= I / sigma^2) */
  float orientation;     /* OERegion::orientation */
  float extremum_value;  /* OERegion::extremum_value */
  uint8_t type;          /* OERegion::Type, stays Undefined (= 11) as in the reference */
  int8_t extremum_type;  /* OERegion::ExtremumType: -1 Min, 1 Max */
  int16_t reserved;
  int32_t s, o;
  int32_t xi, yi;
} sara_b200_keypoint; /* 52 bytes */

typedef struct sara_b200_ctx sara_b200_ctx;

/* Limits a ctx is sized for; everything is allocated once at creation. */
typedef struct sara_b200_limits
{
  int32_t max_width;       /* of the INPUT image */
  int32_t max_height;
  int32_t max_keypoints;   /* per frame, <= 0 -> 262144; at most 4194304 */
  int32_t num_slots;       /* frames in flight, <= 0 -> 1 */
  int32_t min_first_octave_index; /* most negative first_octave_index to size for (0 or -1) */
} sara_b200_limits;

/* SIFT arguments: exactly the scalar arguments of compute_sift_keypoints
 * (SIFT.hpp:26-32).  `parallel` has no meaning on the GPU and is omitted. */
typedef struct sara_b200_sift_args
{
  sara_b200_pyramid_params pyramid_params;
  float gauss_truncate;             /* 4.f  */
  float extremum_thres;             /* 0.01f */
  float edge_ratio_thres;           /* 10.f */
  int32_t extremum_refinement_iter; /* 5; NB SIFT.cpp:45-51 passes it as img_padding_sz */
} sara_b200_sift_args;

/* Arguments of the ComputeDoGExtrema constructor, DoG.hpp:72-78. */
typedef struct sara_b200_dog_args
{
  sara_b200_pyramid_params pyramid_params;
  float gauss_truncate;
  float extremum_thres;
  float edge_ratio_thres;
  int32_t img_padding_sz;
  int32_t extremum_refinement_iter;
} sara_b200_dog_args;

/* Stage timings of the last frame run with profiling on, in ms (CUDA events). */
typedef struct sara_b200_timings
{
  float upload;      /* host -> device copy of the frame */
  float pyramid;     /* Gaussian pyramid + DoG (all octaves) */
  float extrema;     /* classify + compaction + refinement */
  float orientation; /* dominant orientations + expansion */
  float descriptor;  /* 128-D descriptors */
  float total;       /* first kernel to last kernel */
  int32_t pyramid_launches; /* kernels launched by the pyramid stage */
  int32_t total_launches;   /* kernels launched for the frame */
  float pyramid_top_kernel;        /* the pyramid's longest launch (most taps, octave 0), ms; 0 if not measured */
  float pyramid_top_kernel_mbytes; /* its algorithmic HBM traffic in MB */
} sara_b200_timings;

SARA_B200_API int sara_b200_version(void);
SARA_B200_API const char* sara_b200_last_error(const sara_b200_ctx* ctx); /* ctx may be NULL: creation errors */

SARA_B200_API void sara_b200_default_pyramid_params(sara_b200_pyramid_params* p); /* ImagePyramid.hpp:36-42 */
SARA_B200_API void sara_b200_default_sift_args(sara_b200_sift_args* a);           /* SIFT.hpp:26-32 */
SARA_B200_API void sara_b200_default_dog_args(sara_b200_dog_args* a);             /* DoG.hpp:72-78 */

SARA_B200_API int sara_b200_create(int device, const sara_b200_limits* limits, sara_b200_ctx** out);
SARA_B200_API void sara_b200_destroy(sara_b200_ctx* ctx);

/* Pinned host memory helpers (frames and results move at PCIe speed only from
 * pinned memory). */
SARA_B200_API int sara_b200_host_alloc(void** ptr, uint64_t bytes);
SARA_B200_API void sara_b200_host_free(void* ptr);

/* Which kernels build the Gaussian / DoG pyramid.  All modes produce the same bits; the
 * choice exists for benchmarking and parity tests.  AUTO picks the fastest measured one. */
typedef enum sara_b200_pyramid_mode
{
  SARA_B200_PYRAMID_AUTO = 0,
  SARA_B200_PYRAMID_GENERIC = 1, /* one launch per scale, any tap count */
  SARA_B200_PYRAMID_STAGE = 2,   /* TMA-staged marching kernel, one launch per scale (default schedule) */
  SARA_B200_PYRAMID_FUSED = 3,   /* TMA-staged fused octave kernel, one launch per octave (default schedule) */
  SARA_B200_PYRAMID_MARCH = 4    /* TMA-staged scatter-form marching kernel, one launch per scale (default schedule) */
} sara_b200_pyramid_mode;
SARA_B200_API int sara_b200_set_pyramid_mode(sara_b200_ctx* ctx, int mode);
/* Octave o + 1 only needs one scale of octave o, so by default the octaves of a frame overlap
 * on side streams.  Turning this off serialises them (used to time single launches alone). */
SARA_B200_API int sara_b200_set_octave_overlap(sara_b200_ctx* ctx, int on);

/* sara_b200_sift_enqueue[_u8] replays a CUDA graph of the frame's launch sequence (captured
 * the first time a geometry / argument set / device pointer is seen on a slot; on by
 * default, off while profiling).  Turning it off issues every launch directly. */
SARA_B200_API int sara_b200_set_graphs(sara_b200_ctx* ctx, int on);

/* Record CUDA events around the stages (sara_b200_last_timings). */
SARA_B200_API int sara_b200_set_profiling(sara_b200_ctx* ctx, int on);
SARA_B200_API int sara_b200_last_timings(sara_b200_ctx* ctx, int slot, sara_b200_timings* out);

/* ---- compute_sift_keypoints (SIFT.hpp:24-33) ------------------------------
 * Synchronous form: image in (host pointer, or device pointer when
 * `image_on_device`), keypoints + descriptors out into caller-owned HOST
 * buffers of `capacity` entries (descriptors: capacity x 128 floats, row-major,
 * as Tensor_<float, 2>, KeypointList.hpp:35-36).  *n_out receives the number of
 * keypoints; if it exceeds `capacity` nothing is copied and OVERFLOW is
 * returned with *n_out set, so a caller can retry with larger buffers via
 * sara_b200_collect.  image: w x h float32, contiguous, x fastest
 * (ImageView<float>, Core/Image/Image.hpp:44-103). */
SARA_B200_API int sara_b200_sift(sara_b200_ctx* ctx, const float* image, int w, int h, int image_on_device,
                   const sara_b200_sift_args* args, sara_b200_keypoint* keypoints,
                   float* descriptors, int capacity, int* n_out);

/* ---- frame ingest: from_rgb8_to_gray32f + compute_sift_keypoints --------------
 * The video loop of the reference (cpp/examples/Sara/FeatureMatching/
 * video_sift_matching.cpp:184-200) converts every decoded RGB8 frame with
 * from_rgb8_to_gray32f (ImageProcessing/FastColorConversion.cpp:42-67) before
 * compute_sift_keypoints.  These entry points take the 8-bit frame itself --
 * `channels` = 3: interleaved RGB8 (ImageView<Rgb8>), 1: gray8 -- convert it on the
 * device with the reference's arithmetic (bit-identical float image) and run the same
 * chain, so a frame crosses PCIe as 3 or 1 bytes per pixel instead of 4.  A device-resident
 * 8-bit frame (e.g. a decoder surface) must be 4-byte aligned. */
SARA_B200_API int sara_b200_sift_u8(sara_b200_ctx* ctx, const uint8_t* image, int w, int h, int channels,
                      int image_on_device, const sara_b200_sift_args* args, sara_b200_keypoint* keypoints,
                      float* descriptors, int capacity, int* n_out);
SARA_B200_API int sara_b200_sift_enqueue_u8(sara_b200_ctx* ctx, int slot, const uint8_t* image, int w, int h,
                              int channels, int image_on_device, const sara_b200_sift_args* args, void* stream);
/* The conversion alone (unit parity): host in, host out (w * h floats). */
SARA_B200_API int sara_b200_to_gray32f(sara_b200_ctx* ctx, const uint8_t* src, int w, int h, int channels,
                         float* dst);

/* Asynchronous form: enqueue the whole frame on `slot`'s stream (or on
 * `stream`, a cudaStream_t passed as void*, when non-NULL) and return at once;
 * collect later.  A host `image` must stay valid until the matching collect. */
SARA_B200_API int sara_b200_sift_enqueue(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                           int image_on_device, const sara_b200_sift_args* args, void* stream);
/* Waits for the slot, copies results to host buffers (either may be NULL to
 * skip that copy), frees the slot. */
SARA_B200_API int sara_b200_collect(sara_b200_ctx* ctx, int slot, sara_b200_keypoint* keypoints,
                      float* descriptors, int capacity, int* n_out);
/* Same as sara_b200_collect, but into caller-owned DEVICE buffers (device-to-device copies):
 * for consumers that stay on the GPU -- a matcher, or the NCCL gather of a multi-GPU run
 * (sara_b200/parallel.py).  Frees the slot. */
SARA_B200_API int sara_b200_collect_device(sara_b200_ctx* ctx, int slot, sara_b200_keypoint* d_keypoints,
                             float* d_descriptors, int capacity, int* n_out);
/* Device-resident results of a finished slot (for GPU consumers such as a
 * matcher): pointers stay valid until the slot is enqueued again. */
SARA_B200_API int sara_b200_device_results(sara_b200_ctx* ctx, int slot, const sara_b200_keypoint** keypoints,
                             const float** descriptors, int* n_out);
/* Waits for the slot's GPU work and reads back only the keypoint count (clamped to the
 * context capacity; OVERFLOW is returned when the frame exceeded it). */
SARA_B200_API int sara_b200_wait(sara_b200_ctx* ctx, int slot, int* n_out);

/* ---- ComputeDoGExtrema::operator() (DoG.hpp:116-131, DoG.cpp:23-87) --------
 * Pyramid + DoG + extrema only, explicit padding / iteration arguments.
 * Results stay in the slot; read them with the accessors below. */
SARA_B200_API int sara_b200_dog_extrema(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                          int image_on_device, const sara_b200_dog_args* args);

/* ---- sibling detectors on the same pyramid ------------------------------------
 * ComputeLoGExtrema::operator() (FeatureDetectors/LoG.hpp:71-99, LoG.cpp:20-58) and
 * ComputeDoHExtrema::operator() (FeatureDetectors/Hessian.hpp:195-224, Hessian.cpp:59-98):
 * gaussian_pyramid, then laplacian_pyramid (GaussianPyramid.hpp:156-178) resp.
 * det_of_hessian_pyramid (Hessian.hpp:35-57) -- as many layers as the Gaussian pyramid --
 * and local_scale_space_extrema on s = 1 .. N - 2.  `gauss_truncate` of the args is not used
 * (both functors call gaussian_pyramid with its default).  Results stay in the slot: the
 * function pyramid is read with sara_b200_copy_layer(which = 1, s < N), the extrema with
 * sara_b200_copy_extrema.  Reference defaults: LoG ImagePyramidParams(-1, 3 + 2), thres 0.01,
 * edge ratio 10, padding 1, 5 iterations; DoH ImagePyramidParams(-1, 3 + 2, 2^(1/3), 2),
 * thres 1e-6, edge ratio 10, padding 1, 2 iterations. */
SARA_B200_API int sara_b200_log_extrema(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                          int image_on_device, const sara_b200_dog_args* args);
SARA_B200_API int sara_b200_doh_extrema(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                          int image_on_device, const sara_b200_dog_args* args);

/* ComputeHessianLaplaceMaxima::operator() (FeatureDetectors/Hessian.hpp:60-94, Hessian.cpp:19-57):
 * det-of-Hessian pyramid, then laplace_maxima (RefineExtremum.cpp:659-709) on s = 1 .. N - 1 -- spatial local
 * maxima >= extremum_thres, Laplace scale selection on a 13 x 13 patch over `num_scales` (<= 16) blur steps
 * (select_laplace_scale, RefineExtremum.cpp:523-657), 2-D refinement (RefineExtremum.cpp:132-221).  Of `args`
 * only pyramid_params, extremum_thres, img_padding_sz and extremum_refinement_iter are used.  Reference
 * defaults: ImagePyramidParams(-1, 3 + 1), 1e-5, padding 1, 10 scales, 5 iterations.  Results as above
 * (sara_b200_copy_extrema; the function pyramid through sara_b200_copy_layer(which = 1)). */
SARA_B200_API int sara_b200_hessian_laplace(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                              int image_on_device, const sara_b200_dog_args* args, int num_scales);

/* ComputeHarrisLaplaceCorners::operator() (FeatureDetectors/Harris.hpp:125-138, Harris.cpp:165-230): Harris
 * cornerness of every Gaussian layer (Gradient -> second-moment matrix -> Gaussian(sigma_I) -> det - kappa
 * trace^2, times sigma_D^2 with sigma_D = sigma_I / sqrt(2)), then laplace_maxima as above.  Reference defaults:
 * ImagePyramidParams(-1, 2 + 1, sqrt(2), 1), kappa 0.04, 1e-6, padding 1, 10 scales, 5 iterations. */
SARA_B200_API int sara_b200_harris_laplace(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                             int image_on_device, const sara_b200_dog_args* args, float kappa, int num_scales);

/* ---- gaussian_pyramid + difference_of_gaussians_pyramid only --------------
 * (GaussianPyramid.hpp:35-125, GaussianPyramid.cpp:23-51): the "fused pyramid
 * + DoG" benchmark configuration.  Asynchronous on the slot's stream (or
 * `stream`); finish with sara_b200_wait. */
SARA_B200_API int sara_b200_pyramid_enqueue(sara_b200_ctx* ctx, int slot, const float* image, int w, int h,
                              int image_on_device, const sara_b200_pyramid_params* params,
                              float gauss_truncate, void* stream);

/* ---- stage accessors: ComputeDoGExtrema::gaussians(), diff_of_gaussians(),
 * extrema(s, o) (DoG.hpp:133-165) -------------------------------------------- */
SARA_B200_API int sara_b200_num_octaves(sara_b200_ctx* ctx, int slot);
SARA_B200_API int sara_b200_num_scales(sara_b200_ctx* ctx, int slot);  /* Gaussian layers per octave */
SARA_B200_API int sara_b200_layer_size(sara_b200_ctx* ctx, int slot, int octave, int* w, int* h);
SARA_B200_API float sara_b200_octave_scaling_factor(sara_b200_ctx* ctx, int slot, int octave);
/* which: 0 = Gaussian G(s, o), 1 = DoG D(s, o).  dst: w*h floats on the host. */
SARA_B200_API int sara_b200_copy_layer(sara_b200_ctx* ctx, int slot, int which, int s, int o, float* dst);
/* Extrema before orientation assignment, octave coordinates, reference order
 * (octave-major, scale-minor, raster). */
SARA_B200_API int sara_b200_copy_extrema(sara_b200_ctx* ctx, int slot, sara_b200_keypoint* dst, int capacity,
                           int* n_out);
/* Oriented keypoints in octave coordinates (before the rescale of SIFT.cpp:92-98). */
SARA_B200_API int sara_b200_copy_oriented(sara_b200_ctx* ctx, int slot, sara_b200_keypoint* dst, int capacity,
                            int* n_out);

/* ---- stage functors on caller-supplied extrema ---------------------------------
 * ComputeDominantOrientations::operator() (FeatureDescriptors/Orientation.hpp:233-242,
 * Orientation.cpp:135-166) and ComputeSIFTDescriptor<4, 8>::operator()
 * (FeatureDescriptors/SIFT.hpp:62-166) + the rescale of SIFT.cpp:92-98, run on `n` extrema
 * the caller supplies (host array; x, y, shape, extremum fields and the (s, o) pair of
 * `scale_octave_pairs`, octave coordinates) against the Gaussian pyramid the slot holds.
 * Outputs (host, each may be NULL): `oriented` -- one copy of the extremum per dominant
 * orientation, octave coordinates, extrema without a peak dropped, order kept;
 * `keypoints` / `descriptors` -- the same rows rescaled to image coordinates and their
 * 128-float descriptors.  *n_out = number of oriented keypoints. */
SARA_B200_API int sara_b200_describe_extrema(sara_b200_ctx* ctx, int slot, const sara_b200_keypoint* extrema, int n,
                               sara_b200_keypoint* oriented, sara_b200_keypoint* keypoints,
                               float* descriptors, int capacity, int* n_out);

/* ---- building blocks exposed for unit parity (same semantics as the
 * reference functions named) --------------------------------------------------
 * gaussian(): LinearFiltering.hpp:446-454 / LinearFiltering.cpp:30-68;
 * host in, host out. */
SARA_B200_API int sara_b200_gaussian(sara_b200_ctx* ctx, const float* src, int w, int h, float sigma,
                       float gauss_truncate, float* dst);
/* make_gaussian_kernel(): LinearFiltering.hpp:172-203.  Returns the tap count
 * (negative of the needed capacity if `capacity` is too small). */
SARA_B200_API int sara_b200_make_gaussian_kernel(float sigma, float gauss_truncate, float* taps, int capacity);

/* ---- descriptor matching: AnnMatcher (FeatureMatching/AnnMatcher.hpp:33-84,
 * AnnMatcher.cpp:57-282) -------------------------------------------------------
 * The reference builds two FLANN KD-tree forests (third-party/flann,
 * KDTreeIndexParams{8}) and asks each descriptor for its 3 nearest neighbours in the
 * other set -- an approximate search.  Here the search is EXACT (the answer of FLANN's
 * LinearIndex, distances with the bits of flann::L2<float>, algorithms/dist.h:151-178):
 * candidates come from a tcgen05 tensor-core pass over all pairs (dim == 128), are
 * re-ranked in fp32 and certified; see sara_b200/csrc/match.cu. */
typedef struct sara_b200_match
{
  int32_t x_index, y_index; /* Match::x_index(), y_index(): rows of keys1 / keys2 (Match/Match.hpp:100-113) */
  int32_t rank;             /* Match::rank() */
  float score;              /* Match::score(): ratio of SQUARED distances (AnnMatcher.cpp:135-158) */
  int32_t direction;        /* Match::Direction: 0 SourceToTarget, 1 TargetToSource */
} sara_b200_match;          /* 20 bytes */

typedef struct sara_b200_match_args
{
  float sift_ratio_thres;          /* 1.2f (AnnMatcher.hpp:41); squared inside, as the reference does */
  int32_t self_matching;           /* 0: AnnMatcher(keys1, keys2, ratio); 1: AnnMatcher(keys, ratio, metric, pixel) */
  float min_max_metric_dist_thres; /* 0.5f, KeyProximity (FeatureMatching/KeyProximity.hpp:33) */
  float pixel_dist_thres;          /* 10.f */
  int32_t knn_mode;                /* SARA_B200_KNN_AUTO */
} sara_b200_match_args;

enum
{
  SARA_B200_KNN_AUTO = 0,   /* tensor cores when dim == 128 and the problem is not tiny */
  SARA_B200_KNN_SCALAR = 1, /* exact fp32 CUDA-core search only */
  SARA_B200_KNN_TENSOR = 2  /* tcgen05 candidates + exact re-ranking (dim must be 128) */
};

typedef struct sara_b200_knn_stats
{
  int32_t used_tensor_cores;
  int32_t n_redone;  /* queries the certificate handed to the exact scalar kernel */
  int32_t launches;
  int32_t splits;
  float gpu_ms;      /* CUDA events around the search, copies excluded */
} sara_b200_knn_stats;

SARA_B200_API void sara_b200_default_match_args(sara_b200_match_args* a);
/* tree.knnSearch(query, indices, dists, k, SearchParams()) for every row of `queries` against
 * `data` (both n x dim row-major floats; host pointers, or device pointers when `on_device`),
 * k <= 8, dim <= 256.  idx / dist: nq x k on the host, ascending distance, equal distances by
 * ascending index (KNNSimpleResultSet, util/result_set.h:151-171); unused entries are
 * (-1, FLT_MAX).  stats may be NULL. */
SARA_B200_API int sara_b200_knn(sara_b200_ctx* ctx, const float* queries, int nq, const float* data, int nd, int dim,
                                int k, int on_device, int knn_mode, int32_t* idx, float* dist,
                                sara_b200_knn_stats* stats);
/* AnnMatcher::compute_matches().  desc1 / desc2: n x dim descriptors (KeypointList's
 * Tensor_<float, 2>); kp1 / kp2: the matching features, needed for self matching (KeyProximity)
 * and for Match::operator== on features, NULL otherwise (matches are then equal when their index
 * pairs are).  For self matching pass the same arrays twice.  Matches are written to `out`
 * (capacity entries) ordered by score; *n_out is the full count, OVERFLOW is returned when it
 * exceeds `capacity`.  Empty key lists give BAD_ARG ("the list of key-points is empty",
 * AnnMatcher.cpp:45-46). */
SARA_B200_API int sara_b200_compute_matches(sara_b200_ctx* ctx, const float* desc1, const sara_b200_keypoint* kp1, int n1,
                                  const float* desc2, const sara_b200_keypoint* kp2, int n2, int dim, int on_device,
                                  const sara_b200_match_args* args, sara_b200_match* out, int capacity, int* n_out,
                                  sara_b200_knn_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* SARA_B200_H */
