// Generic pyramid kernels: one separable Gaussian stage (row pass, column pass,
// optional DoG epilogue) for ANY tap count, nearest-neighbour downscale and the
// double-precision bilinear enlarge.  These serve every parameter set; the
// default SIFT schedule runs on march_kernel (pyramid_march.cu) for the large
// octaves, octave_head_kernel / tail_octaves_kernel (below) for the small ones.
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
    constexpr int TAIL_MAX_PIXELS = 12288;  // shared-memory limit; the caller decides from which octave on (ctx.cu)
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
      float kr[KT > 0 ? KT : 1];  // taps in registers (the parameter bank is slow to index)
#pragma unroll
      for (int j = 0; j < (KT > 0 ? KT : 1); ++j)
        kr[j] = k[j];
      // flat pixel index: a 60 x 33 octave takes 2 rounds of the 1024 threads, not 2 x 2
      for (int i = threadIdx.x; i < w * h; i += TAIL_NT)
      {
        const int y = i / w, x = i - y * w;
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
      for (int i = threadIdx.x; i < w * h; i += TAIL_NT)
      {
        const int y = i / w, x = i - y * w;
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

    // ---- scales 1 and 2 of a small octave in ONE launch ("head" of the octave) --------------------
    // Octave o + 1 starts from scale 2 of octave o, so from the second octave on the pyramid is a
    // chain of short, strictly dependent launches: two per octave before the next one can start.
    // For the octaves that cannot fill the machine anyway this kernel does both in one launch:
    // a CTA owns a 64 x 32 tile and recomputes the 11 / 6 pixel halo of the intermediate images
    // in shared memory (G0 tile -> row pass -> G1 -> row pass -> G2), writes G(1), D(0), G(2),
    // D(1) and the sub-sampled base of the next octave.  Arithmetic contract as everywhere:
    // acc = RN(acc + RN(v * k[j])), j ascending; every pass replicates the border OF THE IMAGE
    // (LinearFiltering.hpp:95-100, 137-142): intermediates exist only at image positions and are
    // addressed through clamped coordinates.
    constexpr int HD_NT = 1024;

    template <int K1, int K2, int TW, int TH>
    struct HeadGeom
    {
      static constexpr int c1 = K1 / 2, c2 = K2 / 2;
      static constexpr int w_g1 = TW + 2 * c2, h_g1 = TH + 2 * c2;   // G1 region
      static constexpr int w_r1 = w_g1, h_r1 = h_g1 + 2 * c1;         // row-filtered G0
      static constexpr int w_g0 = w_g1 + 2 * c1, h_g0 = h_r1;         // G0 region
      static constexpr int w_r2 = TW, h_r2 = h_g1;                    // row-filtered G1
      static constexpr int n_g0 = w_g0 * h_g0, n_r1 = w_r1 * h_r1, n_g1 = w_g1 * h_g1, n_r2 = w_r2 * h_r2;
      static constexpr int smem_floats = n_g0 + n_r1 + n_g1 + n_r2;
    };

    struct HeadParams
    {
      const float* G0;
      float *G1, *G2, *D0, *D1, *nextG;
      int w, h, pitch, nw, nh, npitch;
      float k1[16], k2[16];
    };

    // EDGE = false: the whole G0 region of the tile lies inside the image, no coordinate is clamped.
    template <int K1, int K2, int TW, int TH, bool EDGE>
    __device__ __forceinline__ void head_tile(const HeadParams& p, float* hs, int x0, int y0)
    {
      using Gm = HeadGeom<K1, K2, TW, TH>;
      constexpr int c1 = Gm::c1, c2 = Gm::c2;
      float* g0 = hs;
      float* r1 = g0 + Gm::n_g0;
      float* g1 = r1 + Gm::n_r1;
      float* r2 = g1 + Gm::n_g1;
      const int tid = threadIdx.x;
      const int w = p.w, h = p.h;
      float k1[K1], k2[K2];
#pragma unroll
      for (int j = 0; j < K1; ++j)
        k1[j] = p.k1[j];
#pragma unroll
      for (int j = 0; j < K2; ++j)
        k2[j] = p.k2[j];
      auto cx = [&](int x) { return EDGE ? min(max(x, 0), w - 1) : x; };
      auto cy = [&](int y) { return EDGE ? min(max(y, 0), h - 1) : y; };

      // G0 tile: origin (x0 - c2 - c1, y0 - c2 - c1), clamped reads = the replicated border of the image
      const int gx0 = x0 - c2 - c1, gy0 = y0 - c2 - c1;
      for (int i = tid; i < Gm::n_g0; i += HD_NT)
      {
        const int ty = i / Gm::w_g0, tx = i - ty * Gm::w_g0;
        g0[i] = p.G0[static_cast<size_t>(cy(gy0 + ty)) * p.pitch + cx(gx0 + tx)];
      }
      __syncthreads();
      // R1 (row pass of scale 1): columns x0 - c2 .., rows y0 - c2 - c1 ..  A tile position holds the value AT
      // THE CLAMPED image position; for an in-image x the window reads clamp(x - c1 + j).
      const int rx0 = x0 - c2, ry0 = gy0;
      for (int i = tid; i < Gm::n_r1; i += HD_NT)
      {
        const int ty = i / Gm::w_r1, tx = i - ty * Gm::w_r1;
        const int x = cx(rx0 + tx);
        const float* row = g0 + ty * Gm::w_g0;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < K1; ++j)
          acc = __fadd_rn(acc, __fmul_rn(row[cx(x - c1 + j) - gx0], k1[j]));
        r1[i] = acc;
      }
      __syncthreads();
      // G1 (column pass of scale 1): region origin (x0 - c2, y0 - c2)
      const int ax0 = x0 - c2, ay0 = y0 - c2;
      for (int i = tid; i < Gm::n_g1; i += HD_NT)
      {
        const int ty = i / Gm::w_g1, tx = i - ty * Gm::w_g1;
        const int y = cy(ay0 + ty);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < K1; ++j)
          acc = __fadd_rn(acc, __fmul_rn(r1[(cy(y - c1 + j) - ry0) * Gm::w_r1 + tx], k1[j]));
        g1[i] = acc;
        const int X = ax0 + tx, Y = ay0 + ty;
        if (X >= x0 && X < min(x0 + TW, w) && Y >= y0 && Y < min(y0 + TH, h))
        {
          const size_t g = static_cast<size_t>(Y) * p.pitch + X;
          p.G1[g] = acc;
          p.D0[g] = __fsub_rn(acc, g0[(Y - gy0) * Gm::w_g0 + (X - gx0)]);
        }
      }
      __syncthreads();
      // R2 (row pass of scale 2): columns x0 .., rows y0 - c2 ..
      for (int i = tid; i < Gm::n_r2; i += HD_NT)
      {
        const int ty = i / Gm::w_r2, tx = i - ty * Gm::w_r2;
        const int x = cx(x0 + tx);
        const float* row = g1 + ty * Gm::w_g1;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < K2; ++j)
          acc = __fadd_rn(acc, __fmul_rn(row[cx(x - c2 + j) - ax0], k2[j]));
        r2[i] = acc;
      }
      __syncthreads();
      // G2 (column pass of scale 2), D1, next octave
      for (int i = tid; i < TW * TH; i += HD_NT)
      {
        const int ty = i / TW, tx = i - ty * TW;
        const int X = x0 + tx, Y = y0 + ty;
        if (X >= w || Y >= h)
          continue;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < K2; ++j)
          acc = __fadd_rn(acc, __fmul_rn(r2[(cy(Y - c2 + j) - ay0) * Gm::w_r2 + tx], k2[j]));
        const size_t g = static_cast<size_t>(Y) * p.pitch + X;
        p.G2[g] = acc;
        p.D1[g] = __fsub_rn(acc, g1[(Y - ay0) * Gm::w_g1 + (X - ax0)]);
        if (p.nextG != nullptr && ((X | Y) & 1) == 0 && (X >> 1) < p.nw && (Y >> 1) < p.nh)
          p.nextG[static_cast<size_t>(Y >> 1) * p.npitch + (X >> 1)] = acc;
      }
    }

    template <int K1, int K2, int TW, int TH>
    __global__ void __launch_bounds__(HD_NT)
        octave_head_kernel(const __grid_constant__ HeadParams p)
    {
      using Gm = HeadGeom<K1, K2, TW, TH>;
      extern __shared__ float hs[];
      const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
      const int gx0 = x0 - Gm::c2 - Gm::c1, gy0 = y0 - Gm::c2 - Gm::c1;
      const bool inside = gx0 >= 0 && gy0 >= 0 && gx0 + Gm::w_g0 <= p.w && gy0 + Gm::h_g0 <= p.h;  // block-uniform
      if (inside)
        head_tile<K1, K2, TW, TH, false>(p, hs, x0, y0);
      else
        head_tile<K1, K2, TW, TH, true>(p, hs, x0, y0);
    }

    template <int TW, int TH>
    bool launch_head_tiles(const HeadParams& p, cudaStream_t st)
    {
      using Gm = HeadGeom<11, 13, TW, TH>;
      const int smem = Gm::smem_floats * sizeof(float);
      if (cudaFuncSetAttribute(octave_head_kernel<11, 13, TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
          cudaSuccess)
        return false;
      const dim3 grid((p.w + TW - 1) / TW, (p.h + TH - 1) / TH);
      octave_head_kernel<11, 13, TW, TH><<<grid, HD_NT, smem, st>>>(p);
      return true;
    }

  }  // namespace

  // Scales 1 and 2 (11 and 13 taps: the default schedule) of one octave in one launch, with D(0), D(1) and --
  // when `next` is given -- the base of the next octave (even sampling).  False if not applicable.
  bool launch_octave_head(const OctaveDesc& oc, const OctaveDesc* next, const Taps& t1, const Taps& t2, cudaStream_t st)
  {
    if (t1.n != 11 || t2.n != 13)
      return false;
    HeadParams p{};
    p.G0 = oc.G;
    p.G1 = oc.G + oc.layer_stride;
    p.G2 = oc.G + 2 * static_cast<size_t>(oc.layer_stride);
    p.D0 = oc.D;
    p.D1 = oc.D + oc.layer_stride;
    p.nextG = next ? next->G : nullptr;
    p.w = oc.w;
    p.h = oc.h;
    p.pitch = oc.pitch;
    p.nw = next ? next->w : 0;
    p.nh = next ? next->h : 0;
    p.npitch = next ? next->pitch : 0;
    for (int j = 0; j < 11; ++j)
      p.k1[j] = t1.v[j];
    for (int j = 0; j < 13; ++j)
      p.k2[j] = t2.v[j];
    // 64 x 32 tiles (1.9x halo work) when they give the machine enough CTAs, 32 x 16 tiles (3.9x halo work, a
    // quarter of the latency per CTA) for the tiny octaves
    const int big_tiles = ((oc.w + 63) / 64) * ((oc.h + 31) / 32);
    return big_tiles >= 100 ? launch_head_tiles<64, 32>(p, st) : launch_head_tiles<32, 16>(p, st);
  }

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
