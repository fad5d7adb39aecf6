// Generic pyramid kernels: one separable Gaussian stage (row pass, column pass,
// optional DoG epilogue) for ANY tap count, nearest-neighbour downscale and the
// double-precision bilinear enlarge.  These serve every parameter set; the
// default SIFT schedule runs on the fused octave kernel in pyramid_fused.cu.
//
// Bit-exactness contract (checked against the oracle with memcmp): this file is
// compiled with -fmad=false; every tap is a separate multiply and add, taps are
// accumulated left to right starting from 0.f, exactly as
// DO::Sara::convolve_array (ImageProcessing/LinearFiltering.hpp:44-63).
#include "common.cuh"

namespace sb {

  namespace {

    constexpr int TW = 64;  // output tile width
    constexpr int TH = 32;  // output tile height
    constexpr int NT = 256;

    // apply_row_based_filter followed by apply_column_based_filter
    // (LinearFiltering.hpp:78-149) on one TW x TH tile.  Borders: both passes
    // replicate the border sample of their INPUT (LinearFiltering.hpp:95-100,
    // 137-142); loading the source with clamped coordinates reproduces both,
    // because the row-filtered value of a replicated row is the row-filtered
    // value of the border row.
    __global__ void __launch_bounds__(NT)
        gaussian_stage_kernel(const float* __restrict__ src, int src_pitch, float* __restrict__ dst,
                              int dst_pitch, float* __restrict__ dog, int dog_pitch, int w, int h,
                              const __grid_constant__ Taps taps)
    {
      extern __shared__ float smem[];
      const int K = taps.n;
      const int c = K / 2;
      const int SW = TW + 2 * c;  // staged source width
      const int SH = TH + 2 * c;  // staged source height
      float* s_src = smem;             // SH x SW
      float* s_row = smem + SH * SW;   // SH x TW
      __shared__ float s_taps[kMaxTaps];

      const int tid = threadIdx.x;
      const int x0 = blockIdx.x * TW;
      const int y0 = blockIdx.y * TH;

      for (int i = tid; i < K; i += NT)
        s_taps[i] = taps.v[i];

      for (int i = tid; i < SW * SH; i += NT)
      {
        const int yy = i / SW;
        const int xx = i - yy * SW;
        const int gx = min(max(x0 + xx - c, 0), w - 1);
        const int gy = min(max(y0 + yy - c, 0), h - 1);
        s_src[i] = __ldg(src + static_cast<size_t>(gy) * src_pitch + gx);
      }
      __syncthreads();

      // Row pass: SH rows x TW columns.
      for (int i = tid; i < SH * TW; i += NT)
      {
        const int yy = i / TW;
        const int xx = i - yy * TW;
        const float* p = s_src + yy * SW + xx;
        float sum = 0.f;
        for (int j = 0; j < K; ++j)
          sum = __fadd_rn(sum, __fmul_rn(p[j], s_taps[j]));
        s_row[i] = sum;
      }
      __syncthreads();

      // Column pass + DoG epilogue (GaussianPyramid.cpp:23-51: D = G(s+1) - G(s)).
      for (int i = tid; i < TH * TW; i += NT)
      {
        const int yy = i / TW;
        const int xx = i - yy * TW;
        const int gx = x0 + xx;
        const int gy = y0 + yy;
        if (gx >= w || gy >= h)
          continue;
        const float* p = s_row + yy * TW + xx;
        float sum = 0.f;
        for (int j = 0; j < K; ++j)
          sum = __fadd_rn(sum, __fmul_rn(p[j * TW], s_taps[j]));
        dst[static_cast<size_t>(gy) * dst_pitch + gx] = sum;
        if (dog != nullptr)
          dog[static_cast<size_t>(gy) * dog_pitch + gx] =
              __fsub_rn(sum, s_src[(yy + c) * SW + xx + c]);
      }
    }

    // scale(), ImageProcessing/Resize.cpp:31-61: nearest sample at
    // (int(x * sx), int(y * sy)) with float ratios.
    __global__ void downscale_kernel(const float* __restrict__ src, int sw, int sh, int spitch,
                                     float* __restrict__ dst, int dw, int dh, int dpitch)
    {
      const int x = blockIdx.x * blockDim.x + threadIdx.x;
      const int y = blockIdx.y * blockDim.y + threadIdx.y;
      if (x >= dw || y >= dh)
        return;
      const float sx = __fdiv_rn(static_cast<float>(sw), static_cast<float>(dw));
      const float sy = __fdiv_rn(static_cast<float>(sh), static_cast<float>(dh));
      const int xi = static_cast<int>(__fmul_rn(static_cast<float>(x), sx));
      const int yi = static_cast<int>(__fmul_rn(static_cast<float>(y), sy));
      dst[static_cast<size_t>(y) * dpitch + x] = __ldg(src + static_cast<size_t>(yi) * spitch + xi);
    }

    // enlarge(), Resize.cpp:86-128 + interpolate(), Interpolation.hpp:34-78:
    // bilinear in double, x-fastest tap order, accumulator starting at 0.0,
    // far taps clamped (offset -1).
    __global__ void enlarge_kernel(const float* __restrict__ src, int sw, int sh, int spitch,
                                   float* __restrict__ dst, int dw, int dh, int dpitch)
    {
      const int x = blockIdx.x * blockDim.x + threadIdx.x;
      const int y = blockIdx.y * blockDim.y + threadIdx.y;
      if (x >= dw || y >= dh)
        return;
      const double sx = __ddiv_rn(static_cast<double>(sw), static_cast<double>(dw));
      const double sy = __ddiv_rn(static_cast<double>(sh), static_cast<double>(dh));
      const double px = __dmul_rn(static_cast<double>(x), sx);
      const double py = __dmul_rn(static_cast<double>(y), sy);
      const double ipx = trunc(px), ipy = trunc(py);
      const double fx = __dsub_rn(px, ipx), fy = __dsub_rn(py, ipy);
      const int x0 = static_cast<int>(ipx), y0 = static_cast<int>(ipy);
      double value = 0.;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
        {
          double weight = 1.;
          weight = __dmul_rn(weight, dx == 0 ? __dsub_rn(1., fx) : fx);
          weight = __dmul_rn(weight, dy == 0 ? __dsub_rn(1., fy) : fy);
          const int xs = (x0 + dx < sw) ? x0 + dx : x0 + dx - 1;
          const int ys = (y0 + dy < sh) ? y0 + dy : y0 + dy - 1;
          const double v = static_cast<double>(__ldg(src + static_cast<size_t>(ys) * spitch + xs));
          value = __dadd_rn(value, __dmul_rn(weight, v));
        }
      dst[static_cast<size_t>(y) * dpitch + x] = static_cast<float>(value);
    }

    __global__ void copy2d_kernel(const float* __restrict__ src, int spitch, float* __restrict__ dst,
                                  int dpitch, int w, int h)
    {
      const int x = blockIdx.x * blockDim.x + threadIdx.x;
      const int y = blockIdx.y * blockDim.y + threadIdx.y;
      if (x < w && y < h)
        dst[static_cast<size_t>(y) * dpitch + x] = src[static_cast<size_t>(y) * spitch + x];
    }

    // ---- tail kernel: all the small octaves of a pyramid in ONE launch --------------------
    // Octaves of a few thousand pixels cannot fill the machine and cost a launch per scale;
    // a single CTA keeps the whole octave in shared memory and walks through every scale
    // of every remaining octave: row pass A -> B, column pass B -> C (+ G, D to HBM), the
    // next octave's base is sub-sampled from scale `down` on the way.  Same arithmetic as
    // gaussian_stage_kernel (separate multiply and add, left to right from 0).
    constexpr int TAIL_NT = 1024;
    constexpr int TAIL_MAX_PIXELS = 4096;  // larger octaves are faster on the multi-CTA stage kernel
    constexpr int TAIL_MAX_TAPS = 32;
    constexpr int TAIL_MAX_SCALES = 8;

    struct TailParams
    {
      int first_octave, down;
      int n_taps[TAIL_MAX_SCALES];
      float taps[TAIL_MAX_SCALES][TAIL_MAX_TAPS];
    };

    // One scale of the tail kernel.  KT > 0: compile-time tap count (unrolled), KT == 0: runtime.
    // Pixels whose window stays inside the image skip the border clamps.
    template <int KT>
    __device__ __forceinline__ void tail_stage(const float* A, float* B, float* C, float* N, const float* k, int w, int h,
                                               int pitch, float* Gs, float* Ds, float* Gn, int nw, int nh, int npitch,
                                               int k_runtime = 0)
    {
      const int K = KT > 0 ? KT : k_runtime;
      const int c = K / 2;
      const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
      float kr[KT > 0 ? KT : 1];  // taps in registers (the parameter bank is slow to index)
#pragma unroll
      for (int j = 0; j < (KT > 0 ? KT : 1); ++j)
        kr[j] = k[j];
      for (int y = ty; y < h; y += TAIL_NT / 32)
       for (int x = tx; x < w; x += 32)
      {
        const int i = y * w + x;
        const float* row = A + y * w;
        float sum = 0.f;
        if (x >= c && x + c < w)
        {
          const float* p = row + x - c;
#pragma unroll
          for (int j = 0; j < (KT > 0 ? KT : 1); ++j)
            if (KT > 0)
              sum = __fadd_rn(sum, __fmul_rn(p[j], kr[j]));
          if (KT == 0)
            for (int j = 0; j < K; ++j)
              sum = __fadd_rn(sum, __fmul_rn(p[j], k[j]));
        }
        else if (KT > 0)
        {
          // border pixels: clamped loads, still unrolled so that they are all in flight together
          float v[KT > 0 ? KT : 1];
#pragma unroll
          for (int j = 0; j < (KT > 0 ? KT : 1); ++j)
            v[j] = row[min(max(x - c + j, 0), w - 1)];
#pragma unroll
          for (int j = 0; j < (KT > 0 ? KT : 1); ++j)
            sum = __fadd_rn(sum, __fmul_rn(v[j], kr[j]));
        }
        else
        {
#pragma unroll 1
          for (int j = 0; j < K; ++j)
            sum = __fadd_rn(sum, __fmul_rn(row[min(max(x - c + j, 0), w - 1)], k[j]));
        }
        B[i] = sum;
      }
      __syncthreads();
      for (int y = ty; y < h; y += TAIL_NT / 32)
       for (int x = tx; x < w; x += 32)
      {
        const int i = y * w + x;
        float sum = 0.f;
        if (y >= c && y + c < h)
        {
          const float* p = B + (y - c) * w + x;
#pragma unroll
          for (int j = 0; j < (KT > 0 ? KT : 1); ++j)
            if (KT > 0)
              sum = __fadd_rn(sum, __fmul_rn(p[j * w], kr[j]));
          if (KT == 0)
            for (int j = 0; j < K; ++j)
              sum = __fadd_rn(sum, __fmul_rn(p[j * w], k[j]));
        }
        else if (KT > 0)
        {
          float v[KT > 0 ? KT : 1];
#pragma unroll
          for (int j = 0; j < (KT > 0 ? KT : 1); ++j)
            v[j] = B[min(max(y - c + j, 0), h - 1) * w + x];
#pragma unroll
          for (int j = 0; j < (KT > 0 ? KT : 1); ++j)
            sum = __fadd_rn(sum, __fmul_rn(v[j], kr[j]));
        }
        else
        {
#pragma unroll 1
          for (int j = 0; j < K; ++j)
            sum = __fadd_rn(sum, __fmul_rn(B[min(max(y - c + j, 0), h - 1) * w + x], k[j]));
        }
        C[i] = sum;
        const size_t g = static_cast<size_t>(y) * pitch + x;
        Gs[g] = sum;
        Ds[g] = __fsub_rn(sum, A[i]);
        if (Gn != nullptr && ((x | y) & 1) == 0 && (x >> 1) < nw && (y >> 1) < nh)
        {
          N[(y >> 1) * nw + (x >> 1)] = sum;
          Gn[static_cast<size_t>(y >> 1) * npitch + (x >> 1)] = sum;
        }
      }
    }

    __global__ void __launch_bounds__(TAIL_NT, 1)
        tail_octaves_kernel(const __grid_constant__ PyramidDesc P, const __grid_constant__ TailParams tp, int n_pixels)
    {
      extern __shared__ float t_sm[];
      float* A = t_sm;                  // G(s-1)
      float* B = A + n_pixels;          // row-filtered
      float* C = B + n_pixels;          // G(s)
      float* N = C + n_pixels;          // base of the next octave
      const int tid = threadIdx.x;
      {
        const OctaveDesc& oc = P.oct[tp.first_octave];
        for (int i = tid; i < oc.w * oc.h; i += TAIL_NT)
          A[i] = oc.G[static_cast<size_t>(i / oc.w) * oc.pitch + (i % oc.w)];
      }
      __syncthreads();
      for (int o = tp.first_octave; o < P.n_octaves; ++o)
      {
        const OctaveDesc& oc = P.oct[o];
        const int w = oc.w, h = oc.h;
        const bool has_next = o + 1 < P.n_octaves;
        const int nw = has_next ? P.oct[o + 1].w : 0, nh = has_next ? P.oct[o + 1].h : 0;
        for (int s = 1; s < P.n_scales; ++s)
        {
          const int K = tp.n_taps[s];
          float* Gs = oc.G + static_cast<size_t>(s) * oc.layer_stride;
          float* Ds = oc.D + static_cast<size_t>(s - 1) * oc.layer_stride;
          float* Gn = (has_next && s == tp.down) ? P.oct[o + 1].G : nullptr;
          const int npitch = has_next ? P.oct[o + 1].pitch : 0;
          switch (K)
          {
          case 11: tail_stage<11>(A, B, C, N, tp.taps[s], w, h, oc.pitch, Gs, Ds, Gn, nw, nh, npitch); break;
          case 13: tail_stage<13>(A, B, C, N, tp.taps[s], w, h, oc.pitch, Gs, Ds, Gn, nw, nh, npitch); break;
          case 17: tail_stage<17>(A, B, C, N, tp.taps[s], w, h, oc.pitch, Gs, Ds, Gn, nw, nh, npitch); break;
          case 21: tail_stage<21>(A, B, C, N, tp.taps[s], w, h, oc.pitch, Gs, Ds, Gn, nw, nh, npitch); break;
          case 25: tail_stage<25>(A, B, C, N, tp.taps[s], w, h, oc.pitch, Gs, Ds, Gn, nw, nh, npitch); break;
          default: tail_stage<0>(A, B, C, N, tp.taps[s], w, h, oc.pitch, Gs, Ds, Gn, nw, nh, npitch, K); break;
          }
          __syncthreads();
          float* t = A;
          A = C;
          C = t;
        }
        if (has_next)
        {
          for (int i = tid; i < nw * nh; i += TAIL_NT)
            A[i] = N[i];
          __syncthreads();
        }
      }
    }

  }  // namespace

  void launch_gaussian_stage(const float* src, int src_pitch, float* dst, int dst_pitch, float* dog,
                             int dog_pitch, int w, int h, const Taps& taps, cudaStream_t st)
  {
    const int c = taps.n / 2;
    const size_t smem = sizeof(float) * (static_cast<size_t>(TH + 2 * c) * (TW + 2 * c) +
                                         static_cast<size_t>(TH + 2 * c) * TW);
    if (smem > 48 * 1024)  // per-device attribute, cheap to repeat
      cudaFuncSetAttribute(gaussian_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem));
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH);
    gaussian_stage_kernel<<<grid, NT, smem, st>>>(src, src_pitch, dst, dst_pitch, dog, dog_pitch, w, h,
                                                  taps);
  }

  void launch_downscale(const float* src, int sw, int sh, int spitch, float* dst, int dw, int dh,
                        int dpitch, cudaStream_t st)
  {
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    downscale_kernel<<<grid, block, 0, st>>>(src, sw, sh, spitch, dst, dw, dh, dpitch);
  }

  void launch_enlarge(const float* src, int sw, int sh, int spitch, float* dst, int dw, int dh,
                      int dpitch, cudaStream_t st)
  {
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    enlarge_kernel<<<grid, block, 0, st>>>(src, sw, sh, spitch, dst, dw, dh, dpitch);
  }

  void launch_copy2d(const float* src, int spitch, float* dst, int dpitch, int w, int h,
                     cudaStream_t st)
  {
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    copy2d_kernel<<<grid, block, 0, st>>>(src, spitch, dst, dpitch, w, h);
  }

  // Runs octaves [first_octave, n_octaves) in one launch when they fit (see tail_octaves_kernel).
  // Returns 0 if the configuration is not covered (the caller then uses the per-scale path).
  int launch_tail_octaves(const PyramidDesc& P, int first_octave, int downscale_index, const Taps* taps,
                          cudaStream_t st)
  {
    if (first_octave >= P.n_octaves || P.n_scales > TAIL_MAX_SCALES || downscale_index < 1 ||
        downscale_index >= P.n_scales)
      return 0;
    const OctaveDesc& f = P.oct[first_octave];
    const int n_pixels = f.w * f.h;
    if (n_pixels > TAIL_MAX_PIXELS)
      return 0;
    TailParams tp{};
    tp.first_octave = first_octave;
    tp.down = downscale_index;
    for (int s = 1; s < P.n_scales; ++s)
    {
      if (taps[s].n > TAIL_MAX_TAPS)
        return 0;
      tp.n_taps[s] = taps[s].n;
      for (int j = 0; j < taps[s].n; ++j)
        tp.taps[s][j] = taps[s].v[j];
    }
    for (int o = first_octave; o + 1 < P.n_octaves; ++o)
      if (!downscale_is_even_sampling(P.oct[o].w, P.oct[o].h, P.oct[o + 1].w, P.oct[o + 1].h))
        return 0;
    const size_t smem = sizeof(float) * (3 * static_cast<size_t>(n_pixels) + n_pixels / 4 + 64);
    if (cudaFuncSetAttribute(tail_octaves_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess)
      return 0;
    tail_octaves_kernel<<<1, TAIL_NT, smem, st>>>(P, tp, n_pixels);
    return 1;
  }

}  // namespace sb
