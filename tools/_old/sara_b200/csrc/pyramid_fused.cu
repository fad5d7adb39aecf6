// Fused octave kernel: ONE launch turns the octave base G(0, o) into G(1..5, o),
// D(0..4, o) and the base of the next octave, for the default SIFT schedule
// (6 scales, k = 2^(1/3): Gaussian increments of 11, 13, 17, 21, 25 taps).
//
// What it restates (behaviour, not code):
//   gaussian_pyramid               ImageProcessing/GaussianPyramid.hpp:106-122
//   apply_row/column_based_filter  ImageProcessing/LinearFiltering.hpp:78-149
//   difference_of_gaussians_pyramid ImageProcessing/GaussianPyramid.cpp:23-51
//   downscale(G(2, o), 2)          ImageProcessing/Resize.cpp:31-83
//
// Design ("fused octave kernel"):
//  * A CTA owns a strip of TX output columns x a segment of rows and MARCHES down
//    the rows 8 at a time.  All five cascade stages advance together, each one
//    lagging the previous by its kernel radius, so the cascade halo (41 px) is
//    recomputed only at the strip/segment edges instead of around every tile.
//  * The octave base is staged by TMA (cp.async.bulk.tensor.2d, two boxes per
//    8-row block, mbarrier completion), double buffered two blocks ahead.
//  * Per stage: row pass (x-convolution) from an 8-row block into a ring of
//    row-filtered rows, column pass (y-convolution) out of that ring.  Both are
//    register tiled (8 outputs per thread along the convolution axis) and use
//    packed f32x2 arithmetic over the OTHER axis, so sliding windows stay aligned.
//  * Arithmetic contract: every tap is RN(acc + RN(b * k)), left to right from
//    +0, exactly DO::Sara::convolve_array.  ptxas contracts mul.f32x2 + add.f32x2
//    into FFMA2 even with --fmad=false (profiles/microbench), so the addition is
//    issued as fma.rn.f32x2(acc, ONE, p) with ONE = 1.0f read from the kernel
//    parameters: acc * 1 + p is a single rounding of the exact sum, i.e. RN(acc + p).
//  * Borders replicate at EVERY stage (LinearFiltering.hpp:95-100, 137-142): rows
//    by clamping the ring row a column-pass window reads, columns by overwriting
//    the out-of-image part of each stage's block with the border column (warp
//    shuffle broadcast).
//  * D(s-1) = G(s) - G(s-1) is emitted by stage s; G(s-1) is re-read through L2
//    (this CTA wrote it a few steps earlier).
#include <cuda.h>
#include <cstdio>

#include "common.cuh"
#include "fp32x2_tma.cuh"

namespace sb {

  namespace fused {

    constexpr int NT = 640;  // threads per CTA (20 warps)
    constexpr int HALO = 41;  // 5 + 6 + 8 + 10 + 12
    constexpr int TMA_SKEW = 3;  // (HALO + TMA_SKEW) % 4 == 0
    constexpr int ROWTAB = 40;   // ring-row table entries per stage

    __host__ __device__ constexpr int K_(int s) { return s == 1 ? 11 : s == 2 ? 13 : s == 3 ? 17 : s == 4 ? 21 : 25; }
    __host__ __device__ constexpr int C_(int s) { return K_(s) / 2; }
    // cumulative radius of stages 1..s (closed form: a recursive constexpr called with a loop
    // variable is NOT folded by nvcc and becomes a real recursive device call)
    __host__ __device__ constexpr int CS_(int s) { return s <= 0 ? 0 : s == 1 ? 5 : s == 2 ? 11 : s == 3 ? 19 : s == 4 ? 29 : 41; }
    // halo still needed after stage s (s = 0: the input)
    __host__ __device__ constexpr int H_(int s) { return HALO - CS_(s); }
    // 8-row blocks kept in the ring of stage s: the windows of a column pass span
    // rows [a - 2c, a + 7] of the block that starts at row a.
    __host__ __device__ constexpr int NB_(int s) { return (2 * C_(s) + 7) / 8 + 1; }
    __host__ __device__ constexpr int round_to(int v, int mod, int rem) { return v + ((rem - v % mod) + mod) % mod; }

    template <int TX>
    struct Cfg
    {
      __host__ __device__ static constexpr int W(int s) { return TX + 2 * H_(s); }
      __host__ __device__ static constexpr int NCH4(int s) { return (W(s) + 3) / 4; }  // 4-column chunks of a row pass
      // Ring pitch (floats).  = 8 (mod 16): the two float4 stores of a row-pass item are conflict free.
      __host__ __device__ static constexpr int PR(int s) { return round_to(4 * NCH4(s), 16, 8); }
      // Block of G(s) rows handed to stage s + 1, row-pair interleaved: [4][P][2] floats.
      // The row pass of stage s + 1 reads positions up to 4 NCH4 + K + 7; P = 2 (mod 4) makes its
      // 16-byte window loads conflict free.
      __host__ __device__ static constexpr int P(int s)
      {
        int need = 4 * NCH4(s + 1) + K_(s + 1) + 8;
        need = need < W(s) ? W(s) : need;
        return round_to(need, 4, 2);
      }
      // TMA box width.  The x coordinate of a box must be 16-byte aligned (a misaligned one
      // raises "illegal instruction", profiles/microbench/tma_probe.cu), so the two boxes of a
      // block start at x0 - 44 instead of x0 - 41 and carry TMA_SKEW extra columns.
      __host__ __device__ static constexpr int BOXW() { return ((W(0) + TMA_SKEW + 1) / 2 + 3) & ~3; }
      // shared memory map (float offsets)
      __host__ __device__ static constexpr int inraw_floats() { return 2 * 2 * 8 * BOXW(); }
      __host__ __device__ static constexpr int off_out(int s)
      {
        int o = inraw_floats();
        for (int i = 0; i < s; ++i)
          o += 8 * P(i);
        return o;
      }
      __host__ __device__ static constexpr int off_ring(int s)
      {
        int o = off_out(5);
        for (int i = 1; i < s; ++i)
          o += NB_(i) * 8 * PR(i);
        return o;
      }
      __host__ __device__ static constexpr int off_rowtab() { return off_ring(6); }
      __host__ __device__ static constexpr int total_floats() { return off_rowtab() + 5 * ROWTAB; }
      __host__ __device__ static constexpr int smem_bytes() { return total_floats() * 4 + 64; }
      // warp items
      __host__ __device__ static constexpr int NWA(int s) { return (NCH4(s) + 7) / 8; }          // row pass: 4 row pairs x 8 chunks
      __host__ __device__ static constexpr int NWB(int s) { return 2 * ((W(s) / 2 + 31) / 32); }  // column pass: 32 column pairs x 4 rows
    };

    // One cascade stage as the kernel sees it (uniform values, read from the constant bank).
    struct StageDesc
    {
      int K, c, depth;      // taps, radius, ring rows (8 * NB)
      int W, HS, HP;        // width of the stage's region; halo after / before the stage
      int CSP, CSS;         // cumulative radius before / after the stage
      int off_in, P_in;     // block of G(s-1): float offset, pitch (x positions)
      int off_ring, PR;     // ring of row-filtered rows
      int off_out, P_out;   // block of G(s) for the next stage (unused for s = 5)
      int NCH4, NWA, NWB;
    };

    struct Params
    {
      float* G;      // G(0, o); layer s at G + s * layer_stride
      float* D;      // D(0, o)
      float* nextG;  // G(0, o + 1) or nullptr
      int w, h, pitch, layer_stride;
      int nw, nh, npitch;
      int hy;  // rows per segment
      float one;
      StageDesc sd[6];  // [1..5]
      float taps[6][28];  // [1..5][K]
    };

    struct Ctl
    {
      int x0, y0, y1, Y;
    };

    // One block of NTAP consecutive taps on a register window: acc[q] += w[q + jj] * k[jb + jj].
    template <int NTAP>
    __device__ __forceinline__ void tap_block(u64 (&acc)[4], const u64 (&w)[12], const float* __restrict__ taps, int jb,
                                              u64 one)
    {
#pragma unroll
      for (int jj = 0; jj < NTAP; ++jj)
      {
        const float k = taps[jb + jj];
        const u64 kk = pack2(k, k);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          acc[q] = add2(acc[q], mul2(w[q + jj], kk), one);
      }
    }

    // ---- row pass of stage s: block of G(s-1) (interleaved row pairs) -> ring of stage s ----
    // Warp item = 4 row pairs x 8 chunks of 4 columns; a lane filters 4 columns of 2 rows
    // (the two rows ride in the two halves of the f32x2 registers).
    __device__ __forceinline__ void row_item(float* sm, const Params& prm, const Ctl& ctl, int s, int t, int wi, int lane)
    {
      const StageDesc& sd = prm.sd[s];
      // quarter-warps hold 4 chunks x 2 row pairs: conflict-free 16-byte loads and stores
      const int rp = ((lane >> 2) & 1) | ((lane >> 3) & 2);
      const int ch = wi * 8 + ((lane & 3) | ((lane >> 1) & 4));
      if (ch >= sd.NCH4)
        return;
      const int u = t - (s - 1);
      const int a = ctl.Y + 8 * u - sd.CSP;
      const int ya = a + 2 * rp;
      const int lo = max(0, ctl.y0 - sd.HP), hi = min(prm.h, ctl.y1 + sd.HP);
      if (ya + 1 < lo || ya >= hi)
        return;
      const int K = sd.K;
      const float* taps = prm.taps[s];
      const ulonglong2* in = reinterpret_cast<const ulonglong2*>(sm + sd.off_in + (rp * sd.P_in + 4 * ch) * 2);
      u64 w[12];
#pragma unroll
      for (int m = 0; m < 6; ++m)
      {
        const ulonglong2 v = in[m];
        w[2 * m] = v.x;
        w[2 * m + 1] = v.y;
      }
      const u64 one = pack2(prm.one, prm.one);
      u64 acc[4] = {0ull, 0ull, 0ull, 0ull};
      int jb = 0;
#pragma unroll 1
      while (jb + 8 <= K)
      {
        tap_block<8>(acc, w, taps, jb, one);
        jb += 8;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          w[q] = w[q + 8];
        if (K - jb > 1)
        {
#pragma unroll
          for (int m = 0; m < 4; ++m)
          {
            const ulonglong2 v = in[jb / 2 + 2 + m];  // positions jb + 4 + 2m, jb + 5 + 2m
            w[4 + 2 * m] = v.x;
            w[5 + 2 * m] = v.y;
          }
        }
      }
      const int rem = K - jb;  // 1, 3 or 5
      tap_block<1>(acc, w, taps, jb, one);
      if (rem >= 3)
      {
        const u64(&w1)[12] = w;  // taps jb + 1, jb + 2
        const float k1 = taps[jb + 1], k2 = taps[jb + 2];
        const u64 kk1 = pack2(k1, k1), kk2 = pack2(k2, k2);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          acc[q] = add2(acc[q], mul2(w1[q + 1], kk1), one);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          acc[q] = add2(acc[q], mul2(w1[q + 2], kk2), one);
        if (rem >= 5)
        {
          const float k3 = taps[jb + 3], k4 = taps[jb + 4];
          const u64 kk3 = pack2(k3, k3), kk4 = pack2(k4, k4);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            acc[q] = add2(acc[q], mul2(w1[q + 3], kk3), one);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            acc[q] = add2(acc[q], mul2(w1[q + 4], kk4), one);
        }
      }
      const int slot = (u + 48) % (sd.depth >> 3);  // 48: multiple of both ring sizes (3 and 4 blocks)
      float* out = sm + sd.off_ring + (slot * 8 + 2 * rp) * sd.PR + 4 * ch;
      *reinterpret_cast<float4*>(out) = make_float4(lo2(acc[0]), lo2(acc[1]), lo2(acc[2]), lo2(acc[3]));
      *reinterpret_cast<float4*>(out + sd.PR) = make_float4(hi2(acc[0]), hi2(acc[1]), hi2(acc[2]), hi2(acc[3]));
    }

    // ---- column pass of stage s --------------------------------------------------------
    // Warp item = 32 column pairs x 4 output rows [b + 4 half, +4); a lane owns two adjacent
    // columns (the halves of the f32x2 registers).  Window rows come through the ring-row table
    // (clamped to the image and wrapped around the ring), so borders need no special path.
    __device__ __forceinline__ void col_item(float* sm, const Params& prm, const Ctl& ctl, int s, int t, int wi, int lane)
    {
      const StageDesc& sd = prm.sd[s];
      const int half = wi & 1, grp = wi >> 1;
      const int K = sd.K;
      const int u = t - (s - 1);
      const int b = ctl.Y + 8 * u - sd.CSS + 4 * half;  // first output row of this item
      const int i = 2 * (grp * 32 + lane);                 // local column of the low lane
      const int xs = ctl.x0 - sd.HS;                       // absolute x of local column 0 (even)
      const int x = xs + i;
      const int w_img = prm.w;
      const bool active = i < sd.W && x >= 0 && x < w_img;
      const bool central = active && i >= sd.HS && i < sd.W - sd.HS;
      const bool pair_ok = x + 1 < w_img;

      // G(s-1) for the D epilogue: issued first, consumed last.
      float2 prev[4];
      if (central)
      {
        const float* gp = prm.G + static_cast<size_t>(s - 1) * prm.layer_stride;
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
          const int y = b + r;
          prev[r] = make_float2(0.f, 0.f);
          if (y >= ctl.y0 && y < ctl.y1)
          {
            const float* q = gp + static_cast<size_t>(y) * prm.pitch + x;
            if (pair_ok)
              prev[r] = __ldcg(reinterpret_cast<const float2*>(q));
            else
              prev[r].x = __ldcg(q);
          }
        }
      }

      u64 acc[4] = {0ull, 0ull, 0ull, 0ull};
      if (active)
      {
        const float* taps = prm.taps[s];
        const float* ring = sm + sd.off_ring + i;
        const int* rowtab = reinterpret_cast<const int*>(sm + prm.sd[0].off_ring) + (s - 1) * ROWTAB + 4 * half;
        const u64 one = pack2(prm.one, prm.one);
        u64 w[12];
#pragma unroll
        for (int n = 0; n < 12; ++n)
          w[n] = *reinterpret_cast<const u64*>(ring + rowtab[n]);
        int jb = 0;
#pragma unroll 1
        while (jb + 8 <= K)
        {
          tap_block<8>(acc, w, taps, jb, one);
          jb += 8;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            w[q] = w[q + 8];
          if (K - jb > 1)
          {
#pragma unroll
            for (int n = 4; n < 12; ++n)
              w[n] = *reinterpret_cast<const u64*>(ring + rowtab[jb + n]);
          }
        }
        const int rem = K - jb;
        tap_block<1>(acc, w, taps, jb, one);
        if (rem >= 3)
        {
          const float k1 = taps[jb + 1], k2 = taps[jb + 2];
          const u64 kk1 = pack2(k1, k1), kk2 = pack2(k2, k2);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            acc[q] = add2(acc[q], mul2(w[q + 1], kk1), one);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            acc[q] = add2(acc[q], mul2(w[q + 2], kk2), one);
          if (rem >= 5)
          {
            const float k3 = taps[jb + 3], k4 = taps[jb + 4];
            const u64 kk3 = pack2(k3, k3), kk4 = pack2(k4, k4);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              acc[q] = add2(acc[q], mul2(w[q + 3], kk3), one);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              acc[q] = add2(acc[q], mul2(w[q + 4], kk4), one);
          }
        }
      }

      // ---- block of G(s) for the next stage (row-pair interleaved) ----
      if (s < 5)
      {
        float* out = sm + sd.off_out;
        const int P_out = sd.P_out;
        if (active)
        {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            *reinterpret_cast<float4*>(out + ((2 * half + e) * P_out + i) * 2) =
                make_float4(lo2(acc[2 * e]), lo2(acc[2 * e + 1]), hi2(acc[2 * e]), hi2(acc[2 * e + 1]));
        }
        // Replicate the border columns over the out-of-image part of the block
        // (LinearFiltering.hpp:95-100 at the next stage): warp-shuffle broadcast.
        const int iL = -xs;              // local column of x = 0
        const int iR = w_img - xs;       // local column of x = w (first one outside)
        const bool left = iL > 0 && (iL >> 6) == grp;                   // this warp item holds x = 0
        const int xr = (w_img - 1) & ~1;                                // low lane of the pair holding x = w - 1
        const bool right = iR < P_out && xr - xs >= 0 && ((xr - xs) >> 6) == grp;
        if (left || right)  // warp-uniform
        {
          __syncwarp();
          if (left)
          {
            const int src = (iL >> 1) & 31;
            u64 v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e)
              v[e] = pack2(__shfl_sync(0xffffffffu, lo2(acc[2 * e]), src),
                           __shfl_sync(0xffffffffu, lo2(acc[2 * e + 1]), src));
            for (int ii = lane; ii < iL; ii += 32)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                *reinterpret_cast<u64*>(out + ((2 * half + e) * P_out + ii) * 2) = v[e];
          }
          if (right)
          {
            const int src = ((xr - xs) >> 1) & 31;
            const bool odd = (w_img & 1) != 0;  // x = w - 1 is the low lane of its pair
            u64 v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
              const float e0 = odd ? lo2(acc[2 * e]) : hi2(acc[2 * e]);
              const float e1 = odd ? lo2(acc[2 * e + 1]) : hi2(acc[2 * e + 1]);
              v[e] = pack2(__shfl_sync(0xffffffffu, e0, src), __shfl_sync(0xffffffffu, e1, src));
            }
            for (int ii = iR + lane; ii < P_out; ii += 32)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                *reinterpret_cast<u64*>(out + ((2 * half + e) * P_out + ii) * 2) = v[e];
          }
        }
      }

      // ---- global results: G(s), D(s-1), base of the next octave ----
      if (central)
      {
        float* gs = prm.G + static_cast<size_t>(s) * prm.layer_stride;
        float* ds = prm.D + static_cast<size_t>(s - 1) * prm.layer_stride;
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
          const int y = b + r;
          if (y < ctl.y0 || y >= ctl.y1)
            continue;
          const size_t o = static_cast<size_t>(y) * prm.pitch + x;
          const float g0 = lo2(acc[r]), g1 = hi2(acc[r]);
          if (pair_ok)
          {
            *reinterpret_cast<float2*>(gs + o) = make_float2(g0, g1);
            *reinterpret_cast<float2*>(ds + o) = make_float2(__fsub_rn(g0, prev[r].x), __fsub_rn(g1, prev[r].y));
          }
          else
          {
            gs[o] = g0;
            ds[o] = __fsub_rn(g0, prev[r].x);
          }
          if (s == 2 && prm.nextG != nullptr && (y & 1) == 0)
          {
            const int xx = x >> 1, yy = y >> 1;
            if (xx < prm.nw && yy < prm.nh)
              prm.nextG[static_cast<size_t>(yy) * prm.npitch + xx] = g0;
          }
        }
      }
    }

    // ---- staging: TMA-written block (plain rows) -> interleaved row pairs, borders replicated ----
    template <int TX>
    __device__ __forceinline__ void convert_item(float* sm, const Params& prm, const Ctl& ctl, int buf, int rp, int lane)
    {
      using C = Cfg<TX>;
      constexpr int BW = C::BOXW(), OFF0 = C::off_out(0), P0 = C::P(0), W0 = C::W(0);
      const float* raw = sm + buf * (2 * 8 * BW);
      float* out = sm + OFF0 + rp * P0 * 2;
      const int xs = ctl.x0 - HALO;
      for (int i = lane; i < P0; i += 32)
      {
        int ic = min(i, W0 - 1);
        const int xc = min(max(xs + ic, 0), prm.w - 1);
        ic = min(max(xc - xs, 0), W0 - 1);
        const int jc = ic + TMA_SKEW;  // column inside the two TMA boxes
        const int half = jc >= BW ? 1 : 0;
        const float* p = raw + half * (8 * BW) + (2 * rp) * BW + (jc - half * BW);
        *reinterpret_cast<float2*>(out + 2 * i) = make_float2(p[0], p[BW]);
      }
    }

    // Number of warp items of stage s's row / column pass at step t (0 when the stage's
    // 8-row block lies outside the rows this CTA needs).
    __device__ __forceinline__ int row_items(const Params& prm, const Ctl& ctl, int s, int t)
    {
      const StageDesc& sd = prm.sd[s];
      const int u = t - (s - 1);
      const int a = ctl.Y + 8 * u - sd.CSP;
      const int lo = max(0, ctl.y0 - sd.HP), hi = min(prm.h, ctl.y1 + sd.HP);
      return (u >= 0 && a + 7 >= lo && a < hi) ? sd.NWA : 0;
    }
    __device__ __forceinline__ int col_items(const Params& prm, const Ctl& ctl, int s, int t)
    {
      const StageDesc& sd = prm.sd[s];
      const int u = t - (s - 1);
      const int b = ctl.Y + 8 * u - sd.CSS;
      const int lo = max(0, ctl.y0 - sd.HS), hi = min(prm.h, ctl.y1 + sd.HS);
      return (u >= 0 && b + 7 >= lo && b < hi) ? sd.NWB : 0;
    }

    template <int TX>
    __global__ void __launch_bounds__(NT, 1)
        fused_octave_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params prm)
    {
      using C = Cfg<TX>;
      extern __shared__ __align__(1024) unsigned char smem_raw[];
      float* sm = reinterpret_cast<float*>(smem_raw);
      constexpr int TOTAL = C::total_floats(), OFF_TAB = C::off_rowtab();
      int* rowtab = reinterpret_cast<int*>(sm + OFF_TAB);
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + TOTAL);  // 2 mbarriers
      int* ctr = reinterpret_cast<int*>(bars + 2);                                    // 4 phase counters

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      Ctl ctl;
      ctl.x0 = blockIdx.x * TX;
      ctl.y0 = blockIdx.y * prm.hy;
      ctl.y1 = min(ctl.y0 + prm.hy, prm.h);
      ctl.Y = ctl.y0 - HALO;
      const int h = prm.h;
      const int T = 5 + (ctl.y1 - ctl.y0 + 2 * HALO - 1) / 8;
      const int in_lo = max(0, ctl.y0 - HALO), in_hi = min(h, ctl.y1 + HALO);
      constexpr int BW = C::BOXW();
      constexpr unsigned kBlockBytes = 2u * 8u * BW * 4u;

      if (tid == 0)
      {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncthreads();

      auto block_needed = [&](int ublk) {
        const int r0 = ctl.Y + 8 * ublk;
        return r0 + 7 >= in_lo && r0 < in_hi;
      };
      auto issue = [&](int ublk) {  // thread 0 only
        float* dst = sm + (ublk & 1) * (2 * 8 * BW);
        mbar_expect_tx(&bars[ublk & 1], kBlockBytes);
        tma_load_2d(dst, &tmap, ctl.x0 - HALO - TMA_SKEW, ctl.Y + 8 * ublk, &bars[ublk & 1]);
        tma_load_2d(dst + 8 * BW, &tmap, ctl.x0 - HALO - TMA_SKEW + BW, ctl.Y + 8 * ublk, &bars[ublk & 1]);
      };
      // Every thread tracks how often each buffer was armed: the wait parity.
      unsigned uses0 = 0, uses1 = 0;
      auto note_issue = [&](int ublk) {
        if (ublk & 1)
          ++uses1;
        else
          ++uses0;
      };
      auto wait_block = [&](int ublk) {
        const unsigned n = (ublk & 1) ? uses1 : uses0;  // uses so far, including this block's
        mbar_wait(&bars[ublk & 1], (n - 1) & 1);
      };

      // prologue: blocks 0 and 1 in flight, block 0 staged.
      for (int ub = 0; ub < 2; ++ub)
        if (block_needed(ub))
        {
          if (tid == 0)
            issue(ub);
          note_issue(ub);
        }
      if (block_needed(0))
      {
        wait_block(0);
        if (warp < 4)
          convert_item<TX>(sm, prm, ctl, 0, warp, lane);
      }
      __syncthreads();

      int phase = 0;
      for (int t = 0; t < T; ++t)
      {
        if (block_needed(t + 2))
        {
          if (tid == 0)
            issue(t + 2);
          note_issue(t + 2);
        }

        // ---------------- phase A: row passes of all stages ----------------
        {
          // Ring-row table of this step's column passes (read after the barrier below):
          // entry m of stage s = float offset of the ring row that holds image row
          // clamp(b_s - c_s + m), b_s = first output row of stage s at this step.
          if (tid < 5 * ROWTAB)
          {
            const int s = tid / ROWTAB + 1, m = tid - (s - 1) * ROWTAB;
            const StageDesc& sd = prm.sd[s];
            const int u = t - (s - 1);
            const int b = ctl.Y + 8 * u - sd.CSS;
            const int yc = min(max(b - sd.c + m, 0), h - 1);
            const int origin = ctl.Y - sd.CSP - 8 * 48;  // row of ring slot 0, 48 blocks up (row_item: slot = (u + 48) % NB)
            const int rr = yc - origin;                   // > 0
            rowtab[tid] = (rr % sd.depth) * sd.PR;
          }
          int cnt[6];
#pragma unroll
          for (int s = 5; s >= 1; --s)
            cnt[s] = row_items(prm, ctl, s, t);
          const int total = cnt[5] + cnt[4] + cnt[3] + cnt[2] + cnt[1];
          if (tid == 0)
            ctr[(phase + 2) & 3] = 0;
          int* my = &ctr[phase & 3];
          while (true)
          {
            int id = 0;
            if (lane == 0)
              id = atomicAdd(my, 1);
            id = __shfl_sync(0xffffffffu, id, 0);
            if (id >= total)
              break;
            int s = 5;
#pragma unroll
            for (int e = 5; e >= 2; --e)
              if (s == e && id >= cnt[e])
              {
                id -= cnt[e];
                s = e - 1;
              }
            row_item(sm, prm, ctl, s, t, id, lane);
          }
          ++phase;
        }
        __syncthreads();

        // ---------------- phase B: column passes + staging of the next input block ----------------
        {
          int cnt[6];
#pragma unroll
          for (int s = 5; s >= 1; --s)
            cnt[s] = col_items(prm, ctl, s, t);
          const int total = cnt[5] + cnt[4] + cnt[3] + cnt[2] + cnt[1];
          const bool stage_next = block_needed(t + 1);
          const int n_conv = stage_next ? 4 : 0;
          if (tid == 0)
            ctr[(phase + 2) & 3] = 0;
          int* my = &ctr[phase & 3];
          while (true)
          {
            int id = 0;
            if (lane == 0)
              id = atomicAdd(my, 1);
            id = __shfl_sync(0xffffffffu, id, 0);
            if (id >= total + n_conv)
              break;
            if (id >= total)
            {
              wait_block(t + 1);
              convert_item<TX>(sm, prm, ctl, (t + 1) & 1, id - total, lane);
              continue;
            }
            int s = 5;
#pragma unroll
            for (int e = 5; e >= 2; --e)
              if (s == e && id >= cnt[e])
              {
                id -= cnt[e];
                s = e - 1;
              }
            col_item(sm, prm, ctl, s, t, id, lane);
          }
          ++phase;
        }
        __syncthreads();
      }
    }

    // ---- host side ------------------------------------------------------------------------
    EncodeTiledFn encode_fn()
    {
      static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
          p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
      }();
      return fn;
    }

    template <int TX>
    bool launch(const OctaveDesc& oc, const OctaveDesc* next, const Taps* taps, cudaStream_t st)
    {
      using C = Cfg<TX>;
      static_assert(C::smem_bytes() <= 232448, "fused octave kernel exceeds 227 KB of shared memory");
      static_assert(TX % 4 == 0 && (HALO + TMA_SKEW) % 4 == 0, "TMA box x coordinates must be 16-byte aligned");
      static bool configured = false;
      if (!configured)
      {
        if (cudaFuncSetAttribute(fused_octave_kernel<TX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 C::smem_bytes()) != cudaSuccess)
          return false;
        configured = true;
      }
      EncodeTiledFn enc = encode_fn();
      if (!enc)
        return false;
      CUtensorMap tmap;
      const cuuint64_t dims[2] = {static_cast<cuuint64_t>(oc.w), static_cast<cuuint64_t>(oc.h)};
      const cuuint64_t strides[1] = {static_cast<cuuint64_t>(oc.pitch) * sizeof(float)};
      const cuuint32_t box[2] = {static_cast<cuuint32_t>(C::BOXW()), 8u};
      const cuuint32_t estr[2] = {1u, 1u};
      if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, oc.G, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
          CUDA_SUCCESS)
        return false;

      Params prm{};
      prm.G = oc.G;
      prm.D = oc.D;
      prm.nextG = next ? next->G : nullptr;
      prm.w = oc.w;
      prm.h = oc.h;
      prm.pitch = oc.pitch;
      prm.layer_stride = oc.layer_stride;
      prm.nw = next ? next->w : 0;
      prm.nh = next ? next->h : 0;
      prm.npitch = next ? next->pitch : 0;
      prm.one = 1.f;
      for (int s = 1; s <= 5; ++s)
      {
        for (int j = 0; j < K_(s); ++j)
          prm.taps[s][j] = taps[s].v[j];
        StageDesc& sd = prm.sd[s];
        sd.K = K_(s);
        sd.c = C_(s);
        sd.depth = 8 * NB_(s);
        sd.W = C::W(s);
        sd.HS = H_(s);
        sd.HP = H_(s - 1);
        sd.CSP = CS_(s - 1);
        sd.CSS = CS_(s);
        sd.off_in = C::off_out(s - 1);
        sd.P_in = C::P(s - 1);
        sd.off_ring = C::off_ring(s);
        sd.PR = C::PR(s);
        sd.off_out = s < 5 ? C::off_out(s) : 0;
        sd.P_out = s < 5 ? C::P(s) : 0;
        sd.NCH4 = C::NCH4(s);
        sd.NWA = C::NWA(s);
        sd.NWB = C::NWB(s);
      }
      prm.sd[0].off_ring = C::off_rowtab();  // sd[0] carries the ring-row table offset

      const int n_strips = (oc.w + TX - 1) / TX;
      int n_segs = 148 / n_strips;
      n_segs = n_segs < 1 ? 1 : n_segs;
      const int max_segs = (oc.h + 31) / 32;  // at least 32 rows per segment
      n_segs = n_segs > max_segs ? max_segs : n_segs;
      int hy = (oc.h + n_segs - 1) / n_segs;
      hy = (hy + 7) & ~7;
      n_segs = (oc.h + hy - 1) / hy;
      prm.hy = hy;
      dim3 grid(n_strips, n_segs);
      fused_octave_kernel<TX><<<grid, NT, C::smem_bytes(), st>>>(tmap, prm);
      return true;
    }

  }  // namespace fused

  // The fused kernel is specialised for the default schedule: 6 Gaussian layers
  // per octave with 11, 13, 17, 21, 25 taps (scale_initial 1.6, k = 2^(1/3)).
  bool fused_octave_supported(const Taps* taps, int n_scales)
  {
    if (n_scales != 6)
      return false;
    for (int s = 1; s <= 5; ++s)
      if (taps[s].n != fused::K_(s))
        return false;
    return fused::encode_fn() != nullptr;
  }

  // downscale(): dst(x, y) = src(int(x * sx), int(y * sy)) with float ratios
  // (Resize.cpp:31-61).  The fused kernel emits src(2x, 2y); this checks that the
  // two agree for the sizes at hand (they do for every size met so far).
  bool downscale_is_even_sampling(int sw, int sh, int dw, int dh)
  {
    const float sx = static_cast<float>(sw) / static_cast<float>(dw);
    const float sy = static_cast<float>(sh) / static_cast<float>(dh);
    for (int x = 0; x < dw; ++x)
      if (static_cast<int>(static_cast<float>(x) * sx) != 2 * x)
        return false;
    for (int y = 0; y < dh; ++y)
      if (static_cast<int>(static_cast<float>(y) * sy) != 2 * y)
        return false;
    return true;
  }

  int launch_fused_octave(const OctaveDesc& oct, const OctaveDesc* next, int downscale_index, const Taps* taps,
                          int n_scales, cudaStream_t st)
  {
    (void) n_scales;
    int launches = 0;
    const bool fuse_down =
        next != nullptr && downscale_index == 2 && downscale_is_even_sampling(oct.w, oct.h, next->w, next->h);
    const OctaveDesc* nx = fuse_down ? next : nullptr;
    bool ok;
    if (oct.w > 1024)
      ok = fused::launch<240>(oct, nx, taps, st);
    else
      ok = fused::launch<128>(oct, nx, taps, st);
    if (!ok)
      return -1;
    ++launches;
    if (next != nullptr && !fuse_down)
    {
      launch_downscale(oct.G + static_cast<size_t>(downscale_index) * oct.layer_stride, oct.w, oct.h, oct.pitch,
                       next->G, next->w, next->h, next->pitch, st);
      ++launches;
    }
    return launches;
  }

}  // namespace sb
