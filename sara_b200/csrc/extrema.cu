// Scale-space extrema of the DoG pyramid: classification, ORDERED compaction
// (the list order must equal the reference's raster order), sub-pixel
// refinement and the final threshold rejection.
//
// Restates (not copies) the behaviour of
//   local_scale_space_extrema   FeatureDetectors/RefineExtremum.cpp:363-521
//   refine_extremum (3-D)       FeatureDetectors/RefineExtremum.cpp:32-130
//   on_edge                     FeatureDetectors/RefineExtremum.cpp:24-30
//   LocalScaleSpaceExtremum     ImageProcessing/Extrema.hpp:28-75
// including the reference quirks N2 (uint8 map: minima typed 255, never
// refined), N7 (emitted at the original raster slot) and N8 (D', h carried
// over when a later iteration leaves the domain).  Compiled with -fmad=false;
// every decision is taken on bit-identical fp32 values, so the integer outputs
// (x, y, s, o, type) are exact.
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"

namespace sb {

  namespace {

    __device__ __forceinline__ float ld(const float* p, int pitch, int x, int y)
    {
      return __ldg(p + static_cast<size_t>(y) * pitch + x);
    }

    // ---- pass 1: classify every pixel of D(s, o), s = 1 .. n_scales - 3 ------
    // All octaves and all scales in ONE streaming pass.  A warp owns a column block of 120
    // pixels (32 lanes x 4 pixels, the first and last lane are halo lanes) and a segment of
    // rows, and marches down: per step it loads one row of each of the MIDDLE DoG layers
    // 1 .. n-2 (16-byte loads, issued one row ahead), keeps three rows of each in registers and
    // decides the non-strict 3x3x3 test of LocalScaleSpaceExtremum
    // (ImageProcessing/Extrema.hpp:28-75) from separable maxima / minima: v >= its neighbours
    // <=> v equals the maximum over them and itself.  The scales next to the outer layers
    // (s = 1 and s = n-2) are decided on the streamed layers first; the few survivors read the
    // 3x3 of the outer layer, the threshold and the edge test (RefineExtremum.cpp:407-437) in a
    // separate, rarely executed routine.  Nothing is staged in shared memory and no two warps
    // talk to each other.
    constexpr int CLS_W = 120;     // pixels a warp classifies per row
    constexpr int CLS_SEG = 8;     // rows per warp: short segments = several waves of warps at 5 blocks per SM (measured: 32
                                   // rows 54.6 us, 16 rows 48.4, 8 rows 44.6 for octave 0 of a 4K frame; 2 halo rows per segment)
    constexpr int CLS_MAXL = 5;    // DoG layers of the default schedule (register-resident path)

    struct ClassifyTiles
    {
      int base[kMaxOctaves + 1];  // first warp of every octave
      int n_cb[kMaxOctaves];      // column blocks of the octave
      int seg;                    // rows a warp classifies
    };

    __device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
    __device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }

    // on_edge(): Hessian functor, ImageProcessing/Differential.hpp:191-226, RefineExtremum.cpp:24-30.
    __device__ __forceinline__ bool on_edge_at(const float* __restrict__ Dc, int pitch, int x, int y, float v,
                                               float edge_ratio)
    {
      const float c2 = __fmul_rn(2.f, v);
      const float hxx = __fadd_rn(__fsub_rn(ld(Dc, pitch, x + 1, y), c2), ld(Dc, pitch, x - 1, y));
      const float hyy = __fadd_rn(__fsub_rn(ld(Dc, pitch, x, y + 1), c2), ld(Dc, pitch, x, y - 1));
      const float hxy = __fdiv_rn(
          __fadd_rn(__fsub_rn(__fsub_rn(ld(Dc, pitch, x + 1, y + 1), ld(Dc, pitch, x - 1, y + 1)),
                              ld(Dc, pitch, x + 1, y - 1)),
                    ld(Dc, pitch, x - 1, y - 1)),
          4.f);
      const float tr = __fadd_rn(hxx, hyy);
      const float det = __fsub_rn(__fmul_rn(hxx, hyy), __fmul_rn(hxy, hxy));
      const float e1 = __fadd_rn(edge_ratio, 1.f);
      return __fmul_rn(__fmul_rn(tr, tr), edge_ratio) >= __fmul_rn(__fmul_rn(e1, e1), fabsf(det));
    }

    // The rare path: a pixel that is an extremum of the streamed layers and passes the threshold.
    // `outer` (or nullptr): the layer whose 3x3 neighbourhood still has to be compared.  Returns the
    // uint8 map value (1 maximum, 255 minimum: the reference stores -1 in an Image<uint8_t>, quirk
    // N2; the maximum is tested first, RefineExtremum.cpp:419-426).
    __device__ __noinline__ int finish_candidate(const float* __restrict__ Dc, const float* __restrict__ outer,
                                                 int pitch, int x, int y, float v, int flags, float edge_ratio)
    {
      bool is_max = (flags & 1) != 0, is_min = (flags & 2) != 0;
      if (outer != nullptr)
      {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx)
          {
            const float a = ld(outer, pitch, x + dx, y + dy);
            is_max = is_max && v >= a;
            is_min = is_min && v <= a;
          }
      }
      if (!(is_max || is_min) || on_edge_at(Dc, pitch, x, y, v, edge_ratio))
        return 0;
      return is_max ? 1 : 255;
    }

    template <int NL>  // DoG layers per octave (n_scales - 1); layers 1 .. NL - 2 are streamed and classified
    __global__ void __launch_bounds__(128, 5)
        classify_sweep_kernel(const __grid_constant__ PyramidDesc P, const __grid_constant__ ClassifyTiles Tl,
                              const ExtremaParams ep)
    {
      constexpr int NM = NL - 2;  // streamed (middle) layers; r[m] is DoG layer m + 1
      static_assert(NM >= 2, "each classified layer has at most one outer neighbour");
      const int lane = threadIdx.x & 31;
      const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
      if (gw >= Tl.base[P.n_octaves])
        return;
      int o = 0;
      while (o + 1 < P.n_octaves && gw >= Tl.base[o + 1])
        ++o;
      const OctaveDesc& oct = P.oct[o];
      const int local = gw - Tl.base[o];
      const int sg = local / Tl.n_cb[o], cb = local - sg * Tl.n_cb[o];
      const int w = oct.w, h = oct.h, pitch = oct.pitch;
      const int x0 = CLS_W * cb - 4 + 4 * lane;  // first of the lane's four pixels
      const int ya = sg * Tl.seg, yb = min(ya + Tl.seg, h);
      const bool ld_on = x0 >= 0 && x0 < w;      // the 16-byte load stays inside the padded row
      const bool out_on = lane >= 1 && lane <= 30 && x0 < w;
      const float thr = __fmul_rn(0.8f, ep.extremum_thres);
      const size_t ls = oct.layer_stride;
      const float* const D1 = oct.D + ls + x0;   // layer 1 at the lane's first pixel

      float r[NM][4][4];  // [middle layer][row slot][pixel]: three rows in use, the fourth being loaded
      auto load_row = [&](int slot, int y) {
        const bool on = ld_on && y >= 0 && y < h;
#pragma unroll
        for (int m = 0; m < NM; ++m)
        {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (on)
            v = __ldg(reinterpret_cast<const float4*>(D1 + m * ls + static_cast<size_t>(y) * pitch));
          r[m][slot][0] = v.x;
          r[m][slot][1] = v.y;
          r[m][slot][2] = v.z;
          r[m][slot][3] = v.w;
        }
      };
      // classify row y; its rows y - 1, y, y + 1 are in slots (mid + 3) % 4, mid, (mid + 1) % 4
      auto classify_row = [&](int mid, int y) {
        const int up = (mid + 3) & 3, dn = (mid + 1) & 3;
        float hmx[NM][4], hmn[NM][4];
#pragma unroll
        for (int m = 0; m < NM; ++m)
        {
          float cmx[6], cmn[6];
#pragma unroll
          for (int c = 0; c < 4; ++c)
          {
            cmx[c + 1] = max3(r[m][up][c], r[m][mid][c], r[m][dn][c]);
            cmn[c + 1] = min3(r[m][up][c], r[m][mid][c], r[m][dn][c]);
          }
          cmx[0] = __shfl_up_sync(0xffffffffu, cmx[4], 1);
          cmn[0] = __shfl_up_sync(0xffffffffu, cmn[4], 1);
          cmx[5] = __shfl_down_sync(0xffffffffu, cmx[1], 1);
          cmn[5] = __shfl_down_sync(0xffffffffu, cmn[1], 1);
#pragma unroll
          for (int c = 0; c < 4; ++c)
          {
            hmx[m][c] = max3(cmx[c], cmx[c + 1], cmx[c + 2]);
            hmn[m][c] = min3(cmn[c], cmn[c + 1], cmn[c + 2]);
          }
        }
        const bool row_ok = ep.pad <= y && y < h - ep.pad && out_on;
        // bit 4 m + c: pixel c of scale s = m + 1 is a maximum / minimum of the streamed layers and
        // passes the cheap rejections (all the rejections are ANDed in the reference)
        unsigned mx_mask = 0u, mn_mask = 0u;
#pragma unroll
        for (int m = 0; m < NM; ++m)
#pragma unroll
          for (int c = 0; c < 4; ++c)
          {
            const float v = r[m][mid][c];
            float M = hmx[m][c], mn = hmn[m][c];
            if (m > 0)
            {
              M = fmaxf(M, hmx[m - 1][c]);
              mn = fminf(mn, hmn[m - 1][c]);
            }
            if (m + 1 < NM)
            {
              M = fmaxf(M, hmx[m + 1][c]);
              mn = fminf(mn, hmn[m + 1][c]);
            }
            const int x = x0 + c;
            const bool live = row_ok && ep.pad <= x && x < w - ep.pad && !(fabsf(v) < thr);
            if (live && v >= M)
              mx_mask |= 1u << (4 * m + c);
            if (live && v <= mn)
              mn_mask |= 1u << (4 * m + c);
          }
        if (out_on)
        {
#pragma unroll
          for (int m = 0; m < NM; ++m)
            *reinterpret_cast<uchar4*>(oct.map + (static_cast<size_t>(m) * h + y) * oct.map_pitch + x0) =
                make_uchar4(0, 0, 0, 0);
        }
        // the few survivors, one per lane and round
        unsigned pend = mx_mask | mn_mask;
        while (__any_sync(0xffffffffu, pend != 0u))
        {
          if (pend != 0u)
          {
            const int b = __ffs(pend) - 1;
            pend &= pend - 1u;
            const int m = b >> 2, x = x0 + (b & 3);
            const float* Dc = oct.D + (m + 1) * ls;
            const float* outer = m == 0 ? oct.D : (m == NM - 1 ? oct.D + (NL - 1) * ls : nullptr);
            const int flags = ((mx_mask >> b) & 1u) | (((mn_mask >> b) & 1u) << 1);
            const int t = finish_candidate(Dc, outer, pitch, x, y, ld(Dc, pitch, x, y), flags, ep.edge_ratio);
            if (t != 0)
            {
              oct.map[(static_cast<size_t>(m) * h + y) * oct.map_pitch + x] = static_cast<uint8_t>(t);
              atomicAdd(oct.row_count + m * h + y, 1);
            }
          }
        }
      };

      // the loads of row y + 2 are in flight while row y is classified
      load_row(0, ya - 1);
      load_row(1, ya);
      load_row(2, ya + 1);
      for (int y = ya; y < yb; y += 4)
      {
        load_row(3, y + 2);
        classify_row(1, y);
        if (y + 1 >= yb)
          break;
        load_row(0, y + 3);
        classify_row(2, y + 1);
        if (y + 2 >= yb)
          break;
        load_row(1, y + 4);
        classify_row(3, y + 2);
        if (y + 3 >= yb)
          break;
        load_row(2, y + 5);
        classify_row(0, y + 3);
      }
    }

    // Any other number of layers: one thread per pixel, straight from global memory.
    __global__ void __launch_bounds__(256)
        classify_generic_kernel(const __grid_constant__ PyramidDesc P, int o, const ExtremaParams ep)
    {
      const OctaveDesc& oct = P.oct[o];
      const int w = oct.w, h = oct.h, pitch = oct.pitch;
      const int n_s = P.n_scales - 3;
      const float thr = __fmul_rn(0.8f, ep.extremum_thres);
      const long long n = static_cast<long long>(w) * h * n_s;
      for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
           i += static_cast<long long>(gridDim.x) * blockDim.x)
      {
        const int x = static_cast<int>(i % w);
        const int y = static_cast<int>((i / w) % h);
        const int s = static_cast<int>(i / (static_cast<long long>(w) * h)) + 1;
        uint8_t t = 0;
        if (ep.pad <= x && x < w - ep.pad && ep.pad <= y && y < h - ep.pad)
        {
          const float* Dc = oct.D + static_cast<size_t>(s) * oct.layer_stride;
          const float v = ld(Dc, pitch, x, y);
          bool is_max = true, is_min = true;
          for (int ds = -1; ds <= 1; ++ds)
            for (int dy = -1; dy <= 1; ++dy)
              for (int dx = -1; dx <= 1; ++dx)
              {
                const float a = ld(Dc + ds * static_cast<long long>(oct.layer_stride), pitch, x + dx, y + dy);
                is_max = is_max && v >= a;
                is_min = is_min && v <= a;
              }
          if ((is_max || is_min) && !(fabsf(v) < thr) && !on_edge_at(Dc, pitch, x, y, v, ep.edge_ratio))
          {
            t = is_max ? 1 : 255;
            atomicAdd(oct.row_count + (s - 1) * h + y, 1);
          }
        }
        oct.map[(static_cast<size_t>(s - 1) * h + y) * oct.map_pitch + x] = t;
      }
    }

    // ---- ordered compaction: one warp per raster row -------------------------
    __global__ void __launch_bounds__(256)
        compact_rows_kernel(const __grid_constant__ PyramidDesc P, int n_segments,
                            const int* __restrict__ seg_off, const int* __restrict__ seg_chunk_off,
                            Candidate* __restrict__ cand, int cap_cand)
    {
      const int lane = threadIdx.x & 31;
      const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
      for (int seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seg < n_segments;
           seg += warps_per_grid)
      {
        int o = 0;
        while (o + 1 < P.n_octaves && seg >= P.oct[o + 1].seg_base)
          ++o;
        const OctaveDesc& oc = P.oct[o];
        const int r = seg - oc.seg_base;
        const int s = r / oc.h + 1;
        const int y = r - (s - 1) * oc.h;
        if (oc.row_count[r] == 0)
          continue;
        int out = seg_off[seg] + seg_chunk_off[seg >> 10];
        const uint8_t* row = oc.map + (static_cast<size_t>(s - 1) * oc.h + y) * oc.map_pitch;
        // The map is almost empty: scan it 4 bytes per lane (128 pixels per warp step) and
        // only expand the words that hold a candidate.  Order inside a row = x order.
        const int w = oc.w;
        const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(row) & 3u);  // row start inside its 4-byte word
        const uint32_t* words = reinterpret_cast<const uint32_t*>(row - mis);
        const int n_words = (mis + w + 3) >> 2;
        for (int w00 = 0; w00 < n_words; w00 += 128)
        {
          uint32_t vv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)  // four 128-pixel steps requested together
          {
            const int wi = w00 + 32 * k + lane;
            vv[k] = wi < n_words ? __ldg(words + wi) : 0u;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
          {
            const int wi = w00 + 32 * k + lane;
            uint32_t v = vv[k];
            // mask bytes outside [row, row + w)
            const int x_first = 4 * wi - mis;
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (x_first + b < 0 || x_first + b >= w)
                v &= ~(0xffu << (8 * b));
            const int mine =
                (v & 0xffu ? 1 : 0) + (v & 0xff00u ? 1 : 0) + (v & 0xff0000u ? 1 : 0) + (v >> 24 ? 1 : 0);
            if (__ballot_sync(0xffffffffu, mine != 0) == 0u)
              continue;
            // exclusive prefix of `mine` across the warp
            int inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
              const int t = __shfl_up_sync(0xffffffffu, inc, d);
              if (lane >= d)
                inc += t;
            }
            int pos = out + inc - mine;
#pragma unroll
            for (int b = 0; b < 4; ++b)
            {
              const uint32_t t = (v >> (8 * b)) & 0xffu;
              if (t != 0)
              {
                if (pos < cap_cand)
                  cand[pos] = Candidate{x_first + b, y, (o << 8) | s, static_cast<int>(t)};
                ++pos;
              }
            }
            out += __shfl_sync(0xffffffffu, inc, 31);
          }
        }
      }
    }

    // ---- refinement ------------------------------------------------------------
    struct Sampler
    {
      const float* base;
      int pitch, stride;
      __device__ __forceinline__ float operator()(int x, int y, int s) const
      {
        return __ldg(base + static_cast<size_t>(s) * stride + static_cast<size_t>(y) * pitch + x);
      }
    };

    // Sign pattern of the eigenvalues of a symmetric 3x3: cyclic Jacobi in fp32
    // with +,-,*,/,sqrt only (the operations the CPU check repeats bit for bit);
    // stands in for Eigen::SelfAdjointEigenSolver (RefineExtremum.cpp:74-81).
    __device__ void sym3_eigenvalues(const float H[9], float lambda[3])
    {
      float a00 = H[0], a11 = H[4], a22 = H[8];
      float a01 = H[1], a02 = H[2], a12 = H[5];
#pragma unroll 1
      for (int sweep = 0; sweep < 8; ++sweep)
      {
        if (a01 == 0.f && a02 == 0.f && a12 == 0.f)
          break;
        if (a01 != 0.f)
        {
          const float theta = __fdiv_rn(__fsub_rn(a11, a00), __fmul_rn(2.f, a01));
          float t = __fdiv_rn(1.f, __fadd_rn(fabsf(theta), __fsqrt_rn(__fadd_rn(__fmul_rn(theta, theta), 1.f))));
          if (theta < 0.f)
            t = -t;
          const float c = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), 1.f)));
          const float sn = __fmul_rn(t, c);
          const float tau = __fmul_rn(t, a01);
          a00 = __fsub_rn(a00, tau);
          a11 = __fadd_rn(a11, tau);
          a01 = 0.f;
          const float b02 = __fsub_rn(__fmul_rn(c, a02), __fmul_rn(sn, a12));
          const float b12 = __fadd_rn(__fmul_rn(sn, a02), __fmul_rn(c, a12));
          a02 = b02;
          a12 = b12;
        }
        if (a02 != 0.f)
        {
          const float theta = __fdiv_rn(__fsub_rn(a22, a00), __fmul_rn(2.f, a02));
          float t = __fdiv_rn(1.f, __fadd_rn(fabsf(theta), __fsqrt_rn(__fadd_rn(__fmul_rn(theta, theta), 1.f))));
          if (theta < 0.f)
            t = -t;
          const float c = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), 1.f)));
          const float sn = __fmul_rn(t, c);
          const float tau = __fmul_rn(t, a02);
          a00 = __fsub_rn(a00, tau);
          a22 = __fadd_rn(a22, tau);
          a02 = 0.f;
          const float b01 = __fsub_rn(__fmul_rn(c, a01), __fmul_rn(sn, a12));
          const float b12 = __fadd_rn(__fmul_rn(sn, a01), __fmul_rn(c, a12));
          a01 = b01;
          a12 = b12;
        }
        if (a12 != 0.f)
        {
          const float theta = __fdiv_rn(__fsub_rn(a22, a11), __fmul_rn(2.f, a12));
          float t = __fdiv_rn(1.f, __fadd_rn(fabsf(theta), __fsqrt_rn(__fadd_rn(__fmul_rn(theta, theta), 1.f))));
          if (theta < 0.f)
            t = -t;
          const float c = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), 1.f)));
          const float sn = __fmul_rn(t, c);
          const float tau = __fmul_rn(t, a12);
          a11 = __fsub_rn(a11, tau);
          a22 = __fadd_rn(a22, tau);
          a12 = 0.f;
          const float b01 = __fsub_rn(__fmul_rn(c, a01), __fmul_rn(sn, a02));
          const float b02 = __fadd_rn(__fmul_rn(sn, a01), __fmul_rn(c, a02));
          a01 = b01;
          a02 = b02;
        }
      }
      lambda[0] = a00;
      lambda[1] = a11;
      lambda[2] = a22;
    }

    // Matrix3f::inverse(): cofactors times 1/det (Eigen 3.4 size-3 path), used
    // at RefineExtremum.cpp:85.  Row-major.
    __device__ void inverse3(const float m[9], float r[9])
    {
#define M_(i, j) m[3 * (i) + (j)]
#define COF_(i, j)                                                                                 \
  __fsub_rn(__fmul_rn(M_(((i) + 1) % 3, ((j) + 1) % 3), M_(((i) + 2) % 3, ((j) + 2) % 3)),         \
            __fmul_rn(M_(((i) + 1) % 3, ((j) + 2) % 3), M_(((i) + 2) % 3, ((j) + 1) % 3)))
      const float c0 = COF_(0, 0), c1 = COF_(1, 0), c2 = COF_(2, 0);
      const float det = __fadd_rn(__fadd_rn(__fmul_rn(c0, M_(0, 0)), __fmul_rn(c1, M_(1, 0))),
                                  __fmul_rn(c2, M_(2, 0)));
      const float invdet = __fdiv_rn(1.f, det);
      r[0] = __fmul_rn(c0, invdet);
      r[1] = __fmul_rn(c1, invdet);
      r[2] = __fmul_rn(c2, invdet);
      r[3] = __fmul_rn(COF_(0, 1), invdet);
      r[4] = __fmul_rn(COF_(1, 1), invdet);
      r[5] = __fmul_rn(COF_(2, 1), invdet);
      r[6] = __fmul_rn(COF_(0, 2), invdet);
      r[7] = __fmul_rn(COF_(1, 2), invdet);
      r[8] = __fmul_rn(COF_(2, 2), invdet);
#undef COF_
#undef M_
    }

    __global__ void __launch_bounds__(128)
        refine_kernel(const __grid_constant__ PyramidDesc P, const ExtremaParams ep,
                      const Candidate* __restrict__ cand, const Counters* __restrict__ counters,
                      int cap_cand, Keypoint* __restrict__ ext_tmp, int* __restrict__ keep)
    {
      const int n = min(counters->n_cand, cap_cand);
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      {
        const Candidate cd = cand[i];
        const int o = cd.so >> 8, s = cd.so & 0xff;
        const OctaveDesc& oc = P.oct[o];
        const Sampler I{oc.D, oc.pitch, oc.layer_stride};
        const int n_dog = P.n_scales - 1;
        const int type = cd.type;  // 1 or 255
        int x = cd.x, y = cd.y;

        float Dp[3] = {0.f, 0.f, 0.f};
        float Dpp[9];
        float hh[3] = {0.f, 0.f, 0.f};
        float lambda[3];
        float val = I(x, y, s);
        bool refined_ok = true;

#pragma unroll 1
        for (int it = 0; it < ep.refine_iter; ++it)
        {
          if (x < ep.pad || x >= oc.w - ep.pad || y < ep.pad || y >= oc.h - ep.pad || s < 1 ||
              s >= n_dog - 1)
            break;
          // gradient / hessian of the pyramid, GaussianPyramid.hpp:184-233.
          Dp[0] = __fdiv_rn(__fsub_rn(I(x + 1, y, s), I(x - 1, y, s)), 2.f);
          Dp[1] = __fdiv_rn(__fsub_rn(I(x, y + 1, s), I(x, y - 1, s)), 2.f);
          Dp[2] = __fdiv_rn(__fsub_rn(I(x, y, s + 1), I(x, y, s - 1)), 2.f);
          const float c2 = __fmul_rn(2.f, I(x, y, s));
          Dpp[0] = __fadd_rn(__fsub_rn(I(x + 1, y, s), c2), I(x - 1, y, s));
          Dpp[4] = __fadd_rn(__fsub_rn(I(x, y + 1, s), c2), I(x, y - 1, s));
          Dpp[8] = __fadd_rn(__fsub_rn(I(x, y, s + 1), c2), I(x, y, s - 1));
          Dpp[1] = Dpp[3] = __fdiv_rn(
              __fadd_rn(__fsub_rn(__fsub_rn(I(x + 1, y + 1, s), I(x - 1, y + 1, s)), I(x + 1, y - 1, s)),
                        I(x - 1, y - 1, s)),
              4.f);
          Dpp[2] = Dpp[6] = __fdiv_rn(
              __fadd_rn(__fsub_rn(__fsub_rn(I(x + 1, y, s + 1), I(x - 1, y, s + 1)), I(x + 1, y, s - 1)),
                        I(x - 1, y, s - 1)),
              4.f);
          Dpp[5] = Dpp[7] = __fdiv_rn(
              __fadd_rn(__fsub_rn(__fsub_rn(I(x, y + 1, s + 1), I(x, y - 1, s + 1)), I(x, y + 1, s - 1)),
                        I(x, y - 1, s - 1)),
              4.f);

          sym3_eigenvalues(Dpp, lambda);
          const float ft = static_cast<float>(type);
          const float lmax =
              fmaxf(fmaxf(__fmul_rn(lambda[0], ft), __fmul_rn(lambda[1], ft)), __fmul_rn(lambda[2], ft));
          if (lmax >= 0.f)
          {
            hh[0] = hh[1] = hh[2] = 0.f;
            break;
          }
          float inv[9];
          inverse3(Dpp, inv);
#pragma unroll
          for (int r = 0; r < 3; ++r)
            hh[r] = __fadd_rn(__fadd_rn(__fmul_rn(-inv[3 * r + 0], Dp[0]), __fmul_rn(-inv[3 * r + 1], Dp[1])),
                              __fmul_rn(-inv[3 * r + 2], Dp[2]));
          if (fmaxf(fabsf(hh[0]), fabsf(hh[1])) > 1.5f)
          {
            refined_ok = false;  // `return false`: position and value stay as initialised
            break;
          }
          if (fminf(fabsf(hh[0]), fabsf(hh[1])) > 0.6f)
          {
            x += hh[0] > 0.f ? 1 : -1;
            y += hh[1] > 0.f ? 1 : -1;
            continue;
          }
          break;
        }

        float px, py, pz;
        if (!refined_ok)
        {
          // RefineExtremum.cpp:41-43: pos initialised to the integer slot before
          // the loop and never updated on the early return.
          px = static_cast<float>(cd.x);
          py = static_cast<float>(cd.y);
          pz = P.scale_rel[s];
        }
        else
        {
          px = static_cast<float>(x);
          py = static_cast<float>(y);
          pz = P.scale_rel[s];
          const float oldval = I(x, y, s);
          const float newval = __fadd_rn(
              oldval,
              __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fmul_rn(Dp[0], hh[0]), __fmul_rn(Dp[1], hh[1])),
                                        __fmul_rn(Dp[2], hh[2]))));
          if ((type == 1 && oldval <= newval) || (type == -1 && oldval >= newval))
          {
            px = __fadd_rn(px, hh[0]);
            py = __fadd_rn(py, hh[1]);
            // std::pow(float, float): evaluated in double and rounded, which
            // agrees with a correctly rounded powf.
            pz = __fmul_rn(pz, static_cast<float>(pow(static_cast<double>(P.k), static_cast<double>(hh[2]))));
            val = newval;
          }
        }

        Keypoint kp;
        kp.x = px;
        kp.y = py;
        // OERegion(coords, scale): shape = I * std::pow(scale, -2) (Feature.hpp:78-82),
        // a double expression narrowed to float; scale^2 is exact in double.
        const double pz2 = static_cast<double>(pz) * static_cast<double>(pz);
        const float a = static_cast<float>(1.0 / pz2);
        kp.shape[0] = a;
        kp.shape[1] = 0.f;
        kp.shape[2] = 0.f;
        kp.shape[3] = a;
        kp.orientation = 0.f;
        kp.extremum_value = val;
        kp.type = 11;
        kp.extremum_type = type == 1 ? 1 : -1;
        kp.reserved = 0;
        kp.s = s;
        kp.o = o;
        kp.xi = cd.x;
        kp.yi = cd.y;
        ext_tmp[i] = kp;
        keep[i] = fabsf(val) < ep.extremum_thres ? 0 : 1;
      }
    }

    __global__ void __launch_bounds__(256)
        emit_kept_kernel(const Keypoint* __restrict__ ext_tmp, const int* __restrict__ keep,
                         const int* __restrict__ off, const int* __restrict__ chunk_off,
                         const Counters* __restrict__ counters, int cap_cand, Keypoint* __restrict__ ext,
                         int cap_ext)
    {
      const int n = min(counters->n_cand, cap_cand);
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (keep[i])
        {
          const int pos = off[i] + chunk_off[i >> 10];
          if (pos < cap_ext)
            ext[pos] = ext_tmp[i];
        }
    }

    // ---- sibling detectors: function pyramids from the Gaussian pyramid -------------------
    // which = 1: Laplacian functor (ImageProcessing/Differential.hpp:106-135, borders replicated)
    //            times float(sigma^2)                  (GaussianPyramid.hpp:156-178)
    // which = 2: Hessian functor (Differential.hpp:191-226) -> 2 x 2 determinant times float(sigma^4)
    //            (FeatureDetectors/Hessian.hpp:35-57)
    struct FunctionNorm
    {
      float v[kMaxScales];
    };
    template <int WHICH>
    __global__ void __launch_bounds__(256)
        function_layers_kernel(const __grid_constant__ PyramidDesc P, int o, const FunctionNorm norm)
    {
      const OctaveDesc& oct = P.oct[o];
      const int w = oct.w, h = oct.h, pitch = oct.pitch;
      const long long n = static_cast<long long>(w) * h * P.n_scales;
      for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
           i += static_cast<long long>(gridDim.x) * blockDim.x)
      {
        const int x = static_cast<int>(i % w);
        const int y = static_cast<int>((i / w) % h);
        const int s = static_cast<int>(i / (static_cast<long long>(w) * h));
        const float* G = oct.G + static_cast<size_t>(s) * oct.layer_stride;
        const float c = ld(G, pitch, x, y);
        float out;
        if (WHICH == 1)
        {
          float value = 0.f;
          if (x == 0)
            value = __fadd_rn(value, __fadd_rn(ld(G, pitch, x + 1, y), c));
          else if (x == w - 1)
            value = __fadd_rn(value, __fadd_rn(c, ld(G, pitch, x - 1, y)));
          else
            value = __fadd_rn(value, __fadd_rn(ld(G, pitch, x + 1, y), ld(G, pitch, x - 1, y)));
          if (y == 0)
            value = __fadd_rn(value, __fadd_rn(ld(G, pitch, x, y + 1), c));
          else if (y == h - 1)
            value = __fadd_rn(value, __fadd_rn(c, ld(G, pitch, x, y - 1)));
          else
            value = __fadd_rn(value, __fadd_rn(ld(G, pitch, x, y + 1), ld(G, pitch, x, y - 1)));
          out = __fsub_rn(value, __fmul_rn(4.f, c));
        }
        else
        {
          const float xn = x == w - 1 ? c : ld(G, pitch, x + 1, y), xp = x == 0 ? c : ld(G, pitch, x - 1, y);
          const float yn = y == h - 1 ? c : ld(G, pitch, x, y + 1), yp = y == 0 ? c : ld(G, pitch, x, y - 1);
          const float hxx = __fadd_rn(__fsub_rn(xn, __fmul_rn(2.f, c)), xp);
          const float hyy = __fadd_rn(__fsub_rn(yn, __fmul_rn(2.f, c)), yp);
          const int nx = x == w - 1 ? 0 : 1, px = x == 0 ? 0 : -1;
          const int ny = y == h - 1 ? 0 : 1, py = y == 0 ? 0 : -1;
          const float hxy = __fdiv_rn(
              __fadd_rn(__fsub_rn(__fsub_rn(ld(G, pitch, x + nx, y + ny), ld(G, pitch, x + px, y + ny)),
                                  ld(G, pitch, x + nx, y + py)),
                        ld(G, pitch, x + px, y + py)),
              4.f);
          out = __fsub_rn(__fmul_rn(hxx, hyy), __fmul_rn(hxy, hxy));
        }
        oct.D[static_cast<size_t>(s) * oct.layer_stride + static_cast<size_t>(y) * pitch + x] = __fmul_rn(out, norm.v[s]);
      }
    }

    // ---- Hessian-Laplace: laplace_maxima (FeatureDetectors/RefineExtremum.cpp:659-709) -----------
    // pass 1: spatial local maxima of the function layers s = 1 .. N - 1 (LocalMax: v >= its eight
    // neighbours, ImageProcessing/Extrema.hpp:50-60, 105-108) that reach the threshold.
    __global__ void __launch_bounds__(256)
        local_max_kernel(const __grid_constant__ PyramidDesc P, int o, float thres, int pad)
    {
      const OctaveDesc& oct = P.oct[o];
      const int w = oct.w, h = oct.h, pitch = oct.pitch;
      const int n_s = P.n_scales - 3;  // layers searched (the caller's descriptor is arranged for this)
      const long long n = static_cast<long long>(w) * h * n_s;
      for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
           i += static_cast<long long>(gridDim.x) * blockDim.x)
      {
        const int x = static_cast<int>(i % w);
        const int y = static_cast<int>((i / w) % h);
        const int s = static_cast<int>(i / (static_cast<long long>(w) * h)) + 1;
        uint8_t t = 0;
        if (pad <= x && x < w - pad && pad <= y && y < h - pad)
        {
          const float* Dc = oct.D + static_cast<size_t>(s) * oct.layer_stride;
          const float v = ld(Dc, pitch, x, y);
          bool is_max = !(v < thres);
          for (int dy = -1; dy <= 1 && is_max; ++dy)
            for (int dx = -1; dx <= 1; ++dx)
              if ((dx | dy) != 0)
                is_max = is_max && v >= ld(Dc, pitch, x + dx, y + dy);
          if (is_max)
          {
            t = 1;
            atomicAdd(oct.row_count + (s - 1) * h + y, 1);
          }
        }
        oct.map[(static_cast<size_t>(s - 1) * h + y) * oct.map_pitch + x] = t;
      }
    }

    // pass 2: a warp per candidate.  select_laplace_scale (RefineExtremum.cpp:523-657): the 13 x 13 patch of
    // G(s - 1, o) around the candidate is blurred num_scales times (row pass, column pass, borders of the PATCH
    // replicated: gaussian() of LinearFiltering.hpp, taps made on the host per layer), the scale-normalised
    // Laplacian at its centre is followed along the scales, and the first local extremum of that profile
    // fixes the scale.  Then the 2-D refine_extremum (RefineExtremum.cpp:132-221) on the function layer.
    constexpr int LP_R = 6, LP_SIDE = 13, LP_N = LP_SIDE * LP_SIDE;
    constexpr int LP_WARPS = 4;

    __global__ void __launch_bounds__(LP_WARPS * 32)
        laplace_refine_kernel(const __grid_constant__ PyramidDesc P, const LaplaceTable* __restrict__ T, int pad,
                              int refine_iter, const Candidate* __restrict__ cand,
                              const Counters* __restrict__ counters, int cap_cand, Keypoint* __restrict__ ext_tmp,
                              int* __restrict__ keep)
    {
      __shared__ float s_a[LP_WARPS][LP_N + 3], s_b[LP_WARPS][LP_N + 3];
      __shared__ float s_log[LP_WARPS][kLaplaceMaxScales + 1];
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      float* A = s_a[wid];
      float* B = s_b[wid];
      float* LoG = s_log[wid];
      const int n = min(counters->n_cand, cap_cand);
      const int ns = T->num_scales;
      for (int i = blockIdx.x * LP_WARPS + wid; i < n; i += gridDim.x * LP_WARPS)
      {
        const Candidate cd = cand[i];
        const int o = cd.so >> 8, s = cd.so & 0xff;
        const OctaveDesc& oc = P.oct[o];
        const int x = cd.x, y = cd.y;
        const bool inside = x - LP_R >= 0 && x + LP_R < oc.w && y - LP_R >= 0 && y + LP_R < oc.h;  // warp-uniform
        bool found = false;
        float scale = 0.f;
        if (inside)
        {
          const float* G = oc.G + static_cast<size_t>(s - 1) * oc.layer_stride;
          for (int p = lane; p < LP_N; p += 32)
          {
            const int v = p / LP_SIDE, u = p - v * LP_SIDE;
            A[p] = G[static_cast<size_t>(y - LP_R + v) * oc.pitch + (x - LP_R + u)];
          }
          __syncwarp();
          for (int k = 0; k <= ns; ++k)
          {
            const int K = T->n_taps[s][k];
            if (K > 0)
            {
              const float* taps = T->taps[s][k];
              const int c = K / 2;
              for (int p = lane; p < LP_N; p += 32)  // row pass A -> B
              {
                const int v = p / LP_SIDE, u = p - v * LP_SIDE;
                float acc = 0.f;
                for (int j = 0; j < K; ++j)
                  acc = __fadd_rn(acc, __fmul_rn(A[v * LP_SIDE + min(max(u - c + j, 0), LP_SIDE - 1)], taps[j]));
                B[p] = acc;
              }
              __syncwarp();
              for (int p = lane; p < LP_N; p += 32)  // column pass B -> A
              {
                const int v = p / LP_SIDE, u = p - v * LP_SIDE;
                float acc = 0.f;
                for (int j = 0; j < K; ++j)
                  acc = __fadd_rn(acc, __fmul_rn(B[min(max(v - c + j, 0), LP_SIDE - 1) * LP_SIDE + u], taps[j]));
                A[p] = acc;
              }
              __syncwarp();
            }
            if (lane == 0)
            {
              const int ctr = LP_R * LP_SIDE + LP_R;
              float value = __fadd_rn(0.f, __fadd_rn(A[ctr + 1], A[ctr - 1]));
              value = __fadd_rn(value, __fadd_rn(A[ctr + LP_SIDE], A[ctr - LP_SIDE]));
              const float lap = __fsub_rn(value, __fmul_rn(4.f, A[ctr]));
              const float sc = T->scales[s][k];
              LoG[k] = __fmul_rn(lap, __fmul_rn(sc, sc));
            }
            __syncwarp();
          }
          // first local extremum of the profile (every lane: the result is warp-uniform)
          int k = 1;
          for (; k < ns; ++k)
          {
            found = (LoG[k] <= LoG[k - 1] && LoG[k] <= LoG[k + 1]) || (LoG[k] >= LoG[k - 1] && LoG[k] >= LoG[k + 1]);
            if (found)
              break;
          }
          if (found)
          {
            const float fprime = __fdiv_rn(__fsub_rn(LoG[k + 1], LoG[k - 1]), 2.f);
            const float fsecond = __fadd_rn(__fsub_rn(LoG[k - 1], __fmul_rn(2.f, LoG[k])), LoG[k + 1]);
            const float hh = __fdiv_rn(-fprime, fsecond);
            // std::pow(float, float): evaluated in double and rounded (agrees with a correctly rounded powf)
            scale = __fmul_rn(T->scales[s][k], static_cast<float>(pow(static_cast<double>(T->ratio), static_cast<double>(hh))));
          }
          __syncwarp();
        }
        if (lane != 0)
          continue;
        keep[i] = found ? 1 : 0;
        if (!found)
          continue;
        // refine_extremum, 2-D, type = 1, on the function layer
        const float* F = oc.D + static_cast<size_t>(s) * oc.layer_stride;
        const int pitch = oc.pitch;
        int rx = x, ry = y;
        float g0 = 0.f, g1 = 0.f, h0 = 0.f, h1 = 0.f;
        float val = ld(F, pitch, x, y);
        float px = static_cast<float>(x), py = static_cast<float>(y);
        bool ok = true;
#pragma unroll 1
        for (int it = 0; it < refine_iter; ++it)
        {
          if (rx < pad || rx >= oc.w - pad || ry < pad || ry >= oc.h - pad)
            break;
          const float c = ld(F, pitch, rx, ry);
          const float xn = ld(F, pitch, rx + 1, ry), xp = ld(F, pitch, rx - 1, ry);
          const float yn = ld(F, pitch, rx, ry + 1), yp = ld(F, pitch, rx, ry - 1);
          g0 = __fdiv_rn(__fsub_rn(xn, xp), 2.f);
          g1 = __fdiv_rn(__fsub_rn(yn, yp), 2.f);
          const float hxx = __fadd_rn(__fsub_rn(xn, __fmul_rn(2.f, c)), xp);
          const float hyy = __fadd_rn(__fsub_rn(yn, __fmul_rn(2.f, c)), yp);
          const float hxy = __fdiv_rn(
              __fadd_rn(__fsub_rn(__fsub_rn(ld(F, pitch, rx + 1, ry + 1), ld(F, pitch, rx - 1, ry + 1)),
                                  ld(F, pitch, rx + 1, ry - 1)),
                        ld(F, pitch, rx - 1, ry - 1)),
              4.f);
          const float det = __fsub_rn(__fmul_rn(hxx, hyy), __fmul_rn(hxy, hxy));
          const float tr = __fadd_rn(hxx, hyy);
          if (det <= 0.f || tr >= 0.f)  // type == 1
          {
            g0 = g1 = 0.f;
            break;
          }
          const float invdet = __fdiv_rn(1.f, det);
          const float i00 = __fmul_rn(hyy, invdet), i01 = __fmul_rn(-hxy, invdet);
          const float i10 = __fmul_rn(-hxy, invdet), i11 = __fmul_rn(hxx, invdet);
          h0 = __fadd_rn(__fmul_rn(-i00, g0), __fmul_rn(-i01, g1));
          h1 = __fadd_rn(__fmul_rn(-i10, g0), __fmul_rn(-i11, g1));
          if (fmaxf(fabsf(h0), fabsf(h1)) > 1.5f)
          {
            ok = false;
            break;
          }
          if (fminf(fabsf(h0), fabsf(h1)) > 0.6f)
          {
            rx += h0 > 0.f ? 1 : -1;
            ry += h1 > 0.f ? 1 : -1;
            continue;
          }
          break;
        }
        if (ok)
        {
          px = static_cast<float>(rx);
          py = static_cast<float>(ry);
          const float oldval = ld(F, pitch, rx, ry);
          const float newval = __fadd_rn(oldval, __fmul_rn(0.5f, __fadd_rn(__fmul_rn(g0, h0), __fmul_rn(g1, h1))));
          if (oldval <= newval)
          {
            px = __fadd_rn(px, h0);
            py = __fadd_rn(py, h1);
            val = newval;
          }
        }
        Keypoint kp;
        kp.x = px;
        kp.y = py;
        const double sc2 = static_cast<double>(scale) * static_cast<double>(scale);
        const float a = static_cast<float>(1.0 / sc2);
        kp.shape[0] = a;
        kp.shape[1] = 0.f;
        kp.shape[2] = 0.f;
        kp.shape[3] = a;
        kp.orientation = 0.f;
        kp.extremum_value = val;
        kp.type = 11;
        kp.extremum_type = 1;
        kp.reserved = 0;
        kp.s = s;
        kp.o = o;
        kp.xi = x;
        kp.yi = y;
        ext_tmp[i] = kp;
      }
    }

    // ---- Harris cornerness (FeatureDetectors/Harris.cpp:171-193) -----------------------------------
    __global__ void __launch_bounds__(256)
        second_moment_kernel(const float* __restrict__ G, int w, int h, int pitch, float* __restrict__ mxx,
                             float* __restrict__ mxy, float* __restrict__ myy)
    {
      const long long n = static_cast<long long>(w) * h;
      for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
           i += static_cast<long long>(gridDim.x) * blockDim.x)
      {
        const int x = static_cast<int>(i % w), y = static_cast<int>(i / w);
        const float c = ld(G, pitch, x, y);
        // Gradient functor, Differential.hpp:46-61 (one-sided differences on the border, all divided by 2)
        const float gx = x == 0       ? __fdiv_rn(__fsub_rn(ld(G, pitch, x + 1, y), c), 2.f)
                         : x == w - 1 ? __fdiv_rn(__fsub_rn(c, ld(G, pitch, x - 1, y)), 2.f)
                                      : __fdiv_rn(__fsub_rn(ld(G, pitch, x + 1, y), ld(G, pitch, x - 1, y)), 2.f);
        const float gy = y == 0       ? __fdiv_rn(__fsub_rn(ld(G, pitch, x, y + 1), c), 2.f)
                         : y == h - 1 ? __fdiv_rn(__fsub_rn(c, ld(G, pitch, x, y - 1)), 2.f)
                                      : __fdiv_rn(__fsub_rn(ld(G, pitch, x, y + 1), ld(G, pitch, x, y - 1)), 2.f);
        const size_t g = static_cast<size_t>(y) * pitch + x;
        mxx[g] = __fmul_rn(gx, gx);
        mxy[g] = __fmul_rn(gx, gy);
        myy[g] = __fmul_rn(gy, gy);
      }
    }

    __global__ void __launch_bounds__(256)
        cornerness_kernel(const float* __restrict__ sxx, const float* __restrict__ sxy, const float* __restrict__ syy,
                          int w, int h, int pitch, float kappa, float norm, float* __restrict__ dst)
    {
      const long long n = static_cast<long long>(w) * h;
      for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
           i += static_cast<long long>(gridDim.x) * blockDim.x)
      {
        const int x = static_cast<int>(i % w), y = static_cast<int>(i / w);
        const size_t g = static_cast<size_t>(y) * pitch + x;
        const float a = sxx[g], b = sxy[g], d = syy[g];
        const float det = __fsub_rn(__fmul_rn(a, d), __fmul_rn(b, b));
        const float tr = __fadd_rn(a, d);
        // det - kappa * pow(trace, 2): pow(float, int) is a double, so the whole expression is (Harris.cpp:185-188)
        const double t2 = static_cast<double>(tr) * static_cast<double>(tr);
        const float v = static_cast<float>(__dsub_rn(static_cast<double>(det), __dmul_rn(static_cast<double>(kappa), t2)));
        dst[g] = __fmul_rn(v, norm);
      }
    }

  }  // namespace

  void launch_second_moment(const float* G, int w, int h, int pitch, float* mxx, float* mxy, float* myy, cudaStream_t st)
  {
    const long long n = static_cast<long long>(w) * h;
    const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 16)));
    second_moment_kernel<<<blocks, 256, 0, st>>>(G, w, h, pitch, mxx, mxy, myy);
  }

  void launch_cornerness(const float* sxx, const float* sxy, const float* syy, int w, int h, int pitch, float kappa,
                         float norm, float* dst, cudaStream_t st)
  {
    const long long n = static_cast<long long>(w) * h;
    const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 16)));
    cornerness_kernel<<<blocks, 256, 0, st>>>(sxx, sxy, syy, w, h, pitch, kappa, norm, dst);
  }

  // laplace_maxima over every octave and the layers s = 1 .. N - 1 of the function pyramid in the D stack.
  // `Pf`: the descriptor arranged for N - 1 searched layers (n_scales = N + 2, row counters / segment bases).
  int launch_laplace_maxima(const PyramidDesc& Pf, const LaplaceTable* d_table, float thres, int pad, int refine_iter,
                            int n_segments, int* seg_offsets, Candidate* cand, int cap_cand, Keypoint* ext_tmp,
                            int* scratch, Keypoint* ext, int cap_ext, Counters* counters, cudaStream_t st)
  {
    int launches = 0;
    cudaMemsetAsync(Pf.oct[0].row_count, 0, sizeof(int) * n_segments, st);
    for (int o = 0; o < Pf.n_octaves; ++o)
    {
      const long long n = static_cast<long long>(Pf.oct[o].w) * Pf.oct[o].h * (Pf.n_scales - 3);
      const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 16)));
      local_max_kernel<<<blocks, 256, 0, st>>>(Pf, o, thres, pad);
      ++launches;
    }
    int* chunk_off = scratch;
    int* keep = scratch + 1024;
    int* keep_off = keep + cap_cand;
    launches += exclusive_scan(Pf.oct[0].row_count, seg_offsets, chunk_off, n_segments, nullptr, 0, &counters->n_cand,
                               cap_cand, &counters->overflow, 1, st);
    compact_rows_kernel<<<std::max(1, std::min((n_segments + 7) / 8, 148 * 16)), 256, 0, st>>>(Pf, n_segments, seg_offsets,
                                                                                              chunk_off, cand, cap_cand);
    laplace_refine_kernel<<<148 * 4, LP_WARPS * 32, 0, st>>>(Pf, d_table, pad, refine_iter, cand, counters, cap_cand,
                                                             ext_tmp, keep);
    launches += 2;
    launches += exclusive_scan(keep, keep_off, chunk_off, 0, &counters->n_cand, cap_cand, &counters->n_ext, cap_ext,
                               &counters->overflow, 2, st);
    emit_kept_kernel<<<296, 256, 0, st>>>(ext_tmp, keep, keep_off, chunk_off, counters, cap_cand, ext, cap_ext);
    return launches + 1;
  }

  int launch_function_pyramid(const PyramidDesc& P, int which, const float* norm, cudaStream_t st)
  {
    FunctionNorm fn{};
    for (int s = 0; s < P.n_scales; ++s)
      fn.v[s] = norm[s];
    int launches = 0;
    for (int o = 0; o < P.n_octaves; ++o)
    {
      const long long n = static_cast<long long>(P.oct[o].w) * P.oct[o].h * P.n_scales;
      const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 16)));
      if (which == 1)
        function_layers_kernel<1><<<blocks, 256, 0, st>>>(P, o, fn);
      else
        function_layers_kernel<2><<<blocks, 256, 0, st>>>(P, o, fn);
      ++launches;
    }
    return launches;
  }

  // Classifies octaves [o_lo, o_hi).  The row counters of ALL octaves are zeroed when
  // `zero_counts` is set (they are contiguous, starting at octave 0's).
  int launch_classify(const PyramidDesc& P, const ExtremaParams& ep, int n_segments, int o_lo, int o_hi,
                      bool zero_counts, cudaStream_t st)
  {
    if (zero_counts)
      cudaMemsetAsync(P.oct[0].row_count, 0, sizeof(int) * n_segments, st);
    const int n_layers = P.n_scales - 1;
    if (n_layers != CLS_MAXL)
    {
      int launches = 0;
      for (int o = o_lo; o < o_hi && o < P.n_octaves; ++o)
      {
        const long long n = static_cast<long long>(P.oct[o].w) * P.oct[o].h * (P.n_scales - 3);
        const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 16));
        classify_generic_kernel<<<std::max(blocks, 1), 256, 0, st>>>(P, o, ep);
        ++launches;
      }
      return launches;
    }
    ClassifyTiles T{};
    // Short segments for every launch: octave 0 then runs as several waves of warps at 5 blocks per SM
    // (better latency hiding than one wave of 32-row warps), the smaller octaves alone are a latency-bound
    // launch (a warp walks its rows one after the other).
    static const int seg_small = [] {
      const char* e = getenv("SARA_B200_CLS_SEG_SMALL");
      return e ? std::max(4, atoi(e)) : 8;
    }();
    static const int seg_big = [] {
      const char* e = getenv("SARA_B200_CLS_SEG");
      return e ? std::max(4, atoi(e)) : CLS_SEG;
    }();
    T.seg = o_lo == 0 ? seg_big : seg_small;
    for (int o = 0; o < P.n_octaves; ++o)
    {
      const bool in = o >= o_lo && o < o_hi;
      T.n_cb[o] = (P.oct[o].w + CLS_W - 1) / CLS_W;
      const int n_sg = (P.oct[o].h + T.seg - 1) / T.seg;
      T.base[o + 1] = T.base[o] + (in ? T.n_cb[o] * n_sg : 0);
    }
    if (T.base[P.n_octaves] == 0)
      return 0;
    classify_sweep_kernel<CLS_MAXL><<<(T.base[P.n_octaves] + 3) / 4, 128, 0, st>>>(P, T, ep);
    return 1;
  }

  // `classified_upto`: octaves below it were already classified (and the counters zeroed) by an
  // earlier launch_classify on the same stream.
  int launch_extrema(const PyramidDesc& P, const ExtremaParams& ep, int n_segments, int* seg_offsets,
                     Candidate* cand, int cap_cand, Keypoint* ext_tmp, int classified_upto, int* scratch,
                     Keypoint* ext, int cap_ext, Counters* counters, cudaStream_t st)
  {
    int launches = launch_classify(P, ep, n_segments, classified_upto, P.n_octaves, classified_upto == 0, st);
    // scratch layout: [chunk offsets (1024)] [keep flags cap_cand] [flag offsets cap_cand]
    int* chunk_off = scratch;
    int* keep = scratch + 1024;
    int* keep_off = keep + cap_cand;

    launches += exclusive_scan(P.oct[0].row_count, seg_offsets, chunk_off, n_segments, nullptr, 0,
                               &counters->n_cand, cap_cand, &counters->overflow, 1, st);
    // one warp per raster row of every classified layer
    compact_rows_kernel<<<std::max(1, std::min((n_segments + 7) / 8, 148 * 16)), 256, 0, st>>>(P, n_segments, seg_offsets,
                                                                                              chunk_off, cand, cap_cand);
    refine_kernel<<<592, 128, 0, st>>>(P, ep, cand, counters, cap_cand, ext_tmp, keep);
    launches += 2;
    launches += exclusive_scan(keep, keep_off, chunk_off, 0, &counters->n_cand, cap_cand,
                               &counters->n_ext, cap_ext, &counters->overflow, 2, st);
    emit_kept_kernel<<<296, 256, 0, st>>>(ext_tmp, keep, keep_off, chunk_off, counters, cap_cand, ext,
                                          cap_ext);
    ++launches;
    return launches;
  }

}  // namespace sb
