// Frame ingest: 8-bit frames (what a video decoder delivers) -> the float32 image the
// pyramid starts from, on the device, so that a frame crosses PCIe as 1 or 3 bytes per pixel
// instead of 4.
//
// Restates the behaviour of
//   from_rgb8_to_gray32f (non-Halide branch)  ImageProcessing/FastColorConversion.cpp:42-67
//     = DO::Sara::convert(ImageView<Rgb8>, ImageView<float>): per pixel smart_convert_color
//       (Core/Pixel/SmartColorConversion.hpp:237-246): channels to double with
//       to_normalized_float_channel (Core/Pixel/ChannelConversion.hpp:41-53: v / 255.0),
//       rgb_to_gray in double (Core/Pixel/ColorConversion.hpp:27-33:
//       0.2125 R + 0.7154 G + 0.0721 B, left to right), narrowed to float;
//   ImageView<uint8_t>::convert<float>()       Core/Pixel/ChannelConversion.hpp:95-99
//     = float(v) / 255.f.
// The three products of the colour conversion only depend on the 8-bit channel value, so they
// come from 3 x 256 double tables computed on the host with the reference's operations; the
// kernel does the two double additions and the narrowing.  Bit-exact by construction.
#include "common.cuh"

namespace sb {

  namespace {

    __global__ void __launch_bounds__(256)
        rgb8_to_gray32f_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int n_pixels,
                               const double* __restrict__ lut)
    {
      __shared__ double s_lut[768];
      for (int i = threadIdx.x; i < 768; i += blockDim.x)
        s_lut[i] = lut[i];
      __syncthreads();
      // 4 pixels (12 bytes = three 32-bit words) per thread and step
      const int n4 = n_pixels >> 2;
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
      for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x)
      {
        const uint32_t a = __ldg(s32 + 3 * q), b = __ldg(s32 + 3 * q + 1), c = __ldg(s32 + 3 * q + 2);
        const uint32_t px[4][3] = {{a & 255u, (a >> 8) & 255u, (a >> 16) & 255u},
                                   {a >> 24, b & 255u, (b >> 8) & 255u},
                                   {(b >> 16) & 255u, b >> 24, c & 255u},
                                   {(c >> 8) & 255u, (c >> 16) & 255u, c >> 24}};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          o[k] = static_cast<float>(
              __dadd_rn(__dadd_rn(s_lut[px[k][0]], s_lut[256 + px[k][1]]), s_lut[512 + px[k][2]]));
        reinterpret_cast<float4*>(dst)[q] = make_float4(o[0], o[1], o[2], o[3]);
      }
      // tail (n_pixels not a multiple of 4)
      if (blockIdx.x == 0 && threadIdx.x < (n_pixels & 3))
      {
        const int i = (n4 << 2) + threadIdx.x;
        dst[i] = static_cast<float>(__dadd_rn(__dadd_rn(s_lut[src[3 * i]], s_lut[256 + src[3 * i + 1]]),
                                              s_lut[512 + src[3 * i + 2]]));
      }
    }

    __global__ void __launch_bounds__(256)
        gray8_to_gray32f_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int n_pixels)
    {
      const int n4 = n_pixels >> 2;
      const uchar4* s4 = reinterpret_cast<const uchar4*>(src);
      for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x)
      {
        const uchar4 v = __ldg(s4 + q);
        reinterpret_cast<float4*>(dst)[q] =
            make_float4(__fdiv_rn(static_cast<float>(v.x), 255.f), __fdiv_rn(static_cast<float>(v.y), 255.f),
                        __fdiv_rn(static_cast<float>(v.z), 255.f), __fdiv_rn(static_cast<float>(v.w), 255.f));
      }
      if (blockIdx.x == 0 && threadIdx.x < (n_pixels & 3))
      {
        const int i = (n4 << 2) + threadIdx.x;
        dst[i] = __fdiv_rn(static_cast<float>(src[i]), 255.f);
      }
    }

  }  // namespace

  // Host side of the tables: lut[c * 256 + v] = coeff_c * (double(v) / 255.0).
  void fill_rgb_to_gray_lut(double* lut768)
  {
    const double coeff[3] = {0.2125, 0.7154, 0.0721};
    for (int c = 0; c < 3; ++c)
      for (int v = 0; v < 256; ++v)
      {
        const volatile double channel = (static_cast<double>(v) - 0.0) / 255.0;
        const volatile double prod = coeff[c] * channel;
        lut768[c * 256 + v] = prod;
      }
  }

  // src and dst on the device; src 4-byte aligned, dst 16-byte aligned.
  void launch_rgb8_to_gray32f(const uint8_t* src, float* dst, int n_pixels, const double* d_lut, cudaStream_t st)
  {
    const int blocks = min(148 * 8, (n_pixels / 4 + 255) / 256 + 1);
    rgb8_to_gray32f_kernel<<<blocks, 256, 0, st>>>(src, dst, n_pixels, d_lut);
  }

  void launch_gray8_to_gray32f(const uint8_t* src, float* dst, int n_pixels, cudaStream_t st)
  {
    const int blocks = min(148 * 8, (n_pixels / 4 + 255) / 256 + 1);
    gray8_to_gray32f_kernel<<<blocks, 256, 0, st>>>(src, dst, n_pixels);
  }

}  // namespace sb
