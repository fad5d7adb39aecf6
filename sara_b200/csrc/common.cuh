// Shared declarations of the B200-native SIFT path (device descriptors, limits,
// error helpers).  Product code: never includes or links anything from oracle/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sara_b200.h"

namespace sb {

  constexpr int kMaxOctaves = 16;
  constexpr int kMaxScales = 12;  // Gaussian layers per octave
  constexpr int kMaxTaps = 129;   // generic separable filter: radius <= 64
  constexpr int kMaxOri = 18;     // strict local maxima of a 36-bin ring

  // One octave of the pyramid in HBM.  All layers of an octave share (w, h,
  // pitch); Gaussian layer s is G + s * layer_stride, DoG layer s is
  // D + s * layer_stride (floats).  pitch is a multiple of 32 floats so that
  // every row starts on a 128-byte line (TMA needs 16-byte row strides).
  struct OctaveDesc
  {
    float* G;
    float* D;
    uint8_t* map;   // extremum map, (n_scales - 3) layers of h rows of map_pitch bytes (ImageProcessing/Extrema.hpp)
    int map_pitch;  // multiple of 16 >= w
    int* row_count; // (n_scales - 3) * h ints: candidates per raster row
    int w, h, pitch;
    int layer_stride;
    int seg_base;   // index of this octave's first row segment in the global segment list
    float scaling;  // ImagePyramid::octave_scaling_factor(o)
  };

  struct PyramidDesc
  {
    int n_octaves;
    int n_scales;               // Gaussian layers per octave
    float scale_rel[kMaxScales];  // float(scale_relative_to_octave(s)), ImagePyramid.hpp:316-319
    float k;                    // scale_geometric_factor
    OctaveDesc oct[kMaxOctaves];
  };

  struct Taps
  {
    int n;
    float v[kMaxTaps];
  };

  // Candidate extremum produced by the classify + compaction passes.
  struct Candidate
  {
    int x, y;
    int so;    // (o << 8) | s
    int type;  // uint8 map value: 1 for maxima, 255 for minima (reference quirk N2)
  };

  // Device-side counters of one frame slot.
  struct Counters
  {
    int n_cand;
    int n_ext;
    int n_kp;
    int overflow;  // bit 0: candidates, bit 1: extrema, bit 2: keypoints
    int ori_next;  // work queues of the orientation / descriptor kernels (next keypoint to take)
    int desc_next;
    int pad[2];
  };

  typedef sara_b200_keypoint Keypoint;
  static_assert(sizeof(Keypoint) == 52, "keypoint record layout is part of the ABI");

  // ---- kernel launchers (defined in the .cu files) --------------------------
  void launch_gaussian_stage(const float* src, int src_pitch, float* dst, int dst_pitch,
                             float* dog, int dog_pitch, int w, int h, const Taps& taps,
                             cudaStream_t st);
  void launch_downscale(const float* src, int sw, int sh, int spitch, float* dst, int dw, int dh,
                        int dpitch, cudaStream_t st);
  void launch_enlarge(const float* src, int sw, int sh, int spitch, float* dst, int dw, int dh,
                      int dpitch, cudaStream_t st);
  void launch_copy2d(const float* src, int spitch, float* dst, int dpitch, int w, int h,
                     cudaStream_t st);

  // Fused octave kernel for the default SIFT schedule (pyramid_fused.cu).
  bool fused_octave_supported(const Taps* taps, int n_scales);
  // Returns the number of kernels launched, or -1 if the launch could not be set up.
  int launch_fused_octave(const OctaveDesc& oct, const OctaveDesc* next, int downscale_index,
                          const Taps* taps, int n_scales, cudaStream_t st);

  // Single-stage marching kernel (pyramid_stage.cu): G(s-1) -> G(s), D(s-1), optionally the
  // base of the next octave; for the tap counts of the default schedule.
  bool stage_kernel_supported(int n_taps);
  bool launch_stage(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h,
                    int pitch, int nw, int nh, int npitch, const Taps& taps, cudaStream_t st);
  // Marching scatter-form kernel (pyramid_march.cu): same contract as launch_stage, fewer fp32
  // instructions (symmetric taps share their products) and no shared-memory window loads.
  bool march_kernel_supported(const Taps& taps);
  // Segment height of the layers that cannot fill the machine, for the launches that follow on this host
  // thread: `throughput` = several frames in flight (fewer, taller segments: less warm-up work), otherwise
  // the shortest chain for a lone frame.
  void set_march_schedule(bool throughput);
  bool launch_march(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h,
                    int pitch, int nw, int nh, int npitch, const Taps& taps, cudaStream_t st);
  bool downscale_is_even_sampling(int sw, int sh, int dw, int dh);
  // All octaves from `first_octave` on in one single-CTA launch (pyramid.cu); 0 if not applicable.
  // Scales 1 and 2 of a small octave in one launch (pyramid.cu); false if the schedule is not the default one.
  bool launch_octave_head(const OctaveDesc& oc, const OctaveDesc* next, const Taps& t1, const Taps& t2, cudaStream_t st);
  int launch_tail_octaves(const PyramidDesc& P, int first_octave, int downscale_index, const Taps* taps,
                          cudaStream_t st);

  // Frame ingest (ingest.cu): 8-bit frames -> float32 gray on the device.
  void fill_rgb_to_gray_lut(double* lut768);
  void launch_rgb8_to_gray32f(const uint8_t* src, float* dst, int n_pixels, const double* d_lut, cudaStream_t st);
  void launch_gray8_to_gray32f(const uint8_t* src, float* dst, int n_pixels, cudaStream_t st);

  // Sibling detectors (extrema.cu): fills the D stack of every octave with n_scales function layers,
  // which = 1: sigma^2 Laplacian, 2: sigma^4 det Hessian; norm[s] is the scale normalisation of layer s.
  int launch_function_pyramid(const PyramidDesc& P, int which, const float* norm, cudaStream_t st);

  // Per-layer constants of select_laplace_scale (RefineExtremum.cpp:523-657), made on the host: the scales of
  // the num_scales + 1 patches and the taps of the blur that leads to each of them (n_taps 0: no blur).
  constexpr int kLaplaceMaxScales = 16;
  constexpr int kLaplaceMaxTaps = 65;
  struct LaplaceTable
  {
    int num_scales;
    float ratio;
    float scales[kMaxScales][kLaplaceMaxScales + 1];
    int n_taps[kMaxScales][kLaplaceMaxScales + 1];
    float taps[kMaxScales][kLaplaceMaxScales + 1][kLaplaceMaxTaps];
  };
  int launch_laplace_maxima(const PyramidDesc& Pf, const LaplaceTable* d_table, float thres, int pad, int refine_iter,
                            int n_segments, int* seg_offsets, Candidate* cand, int cap_cand, Keypoint* ext_tmp,
                            int* scratch, Keypoint* ext, int cap_ext, Counters* counters, cudaStream_t st);

  // Harris cornerness (extrema.cu): g g^T of the Gradient functor, and det - kappa trace^2 (in double, as
  // pow(float, int) makes it in the reference) times the scale normalisation.
  void launch_second_moment(const float* G, int w, int h, int pitch, float* mxx, float* mxy, float* myy, cudaStream_t st);
  void launch_cornerness(const float* sxx, const float* sxy, const float* syy, int w, int h, int pitch, float kappa,
                         float norm, float* dst, cudaStream_t st);

  struct ExtremaParams
  {
    float extremum_thres;
    float edge_ratio;
    int pad;
    int refine_iter;
  };
  int launch_classify(const PyramidDesc& P, const ExtremaParams& ep, int n_segments, int o_lo, int o_hi,
                      bool zero_counts, cudaStream_t st);
  int launch_extrema(const PyramidDesc& P, const ExtremaParams& ep, int n_segments, int* seg_offsets,
                     Candidate* cand, int cap_cand, Keypoint* ext_tmp, int classified_upto, int* bsums,
                     Keypoint* ext, int cap_ext, Counters* counters, cudaStream_t st);
  int launch_orientation(const PyramidDesc& P, const Keypoint* ext, int cap_ext, int* ori_count,
                         float* oris, int* bsums, Keypoint* kp_oct, int cap_kp, Counters* counters,
                         cudaStream_t st);
  int launch_descriptors(const PyramidDesc& P, const Keypoint* kp_oct, Keypoint* kp_out, float* desc,
                         int cap_kp, Counters* counters, cudaStream_t st);

}  // namespace sb
