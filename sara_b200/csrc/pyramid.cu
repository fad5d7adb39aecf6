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

}  // namespace sb
