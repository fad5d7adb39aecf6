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

namespace sb {

  namespace fused {

    typedef unsigned long long u64;

    constexpr int NT = 384;  // threads per CTA
    constexpr int NWARP = NT / 32;
    constexpr int HALO = 41;  // 5 + 6 + 8 + 10 + 12
    constexpr int TMA_SKEW = 3;  // (HALO + TMA_SKEW) % 4 == 0

    __host__ __device__ constexpr int K_(int s) { return s == 1 ? 11 : s == 2 ? 13 : s == 3 ? 17 : s == 4 ? 21 : 25; }
    __host__ __device__ constexpr int C_(int s) { return K_(s) / 2; }
    // cumulative radius of stages 1..s
    __host__ __device__ constexpr int CS_(int s) { return s <= 0 ? 0 : CS_(s - 1) + C_(s); }
    // halo still needed after stage s (s = 0: the input)
    __host__ __device__ constexpr int H_(int s) { return HALO - CS_(s); }
    // 8-row blocks kept in the ring of stage s: the window of a column pass spans
    // rows [a - 2c, a + 7] of the block that starts at row a.
    __host__ __device__ constexpr int NB_(int s) { return (2 * C_(s) + 7) / 8 + 1; }

    template <int TX>
    struct Cfg
    {
      __host__ __device__ static constexpr int W(int s) { return TX + 2 * H_(s); }
      __host__ __device__ static constexpr int NCH(int s) { return (W(s) + 7) / 8; }  // 8-column chunks of a row pass
      __host__ __device__ static constexpr int PR(int s) { return 8 * NCH(s); }       // ring pitch (floats)
      // Block of G(s) rows handed to stage s + 1, row-pair interleaved: [4][P][2].
      // P = 2 (mod 16) makes the float4 window loads of a row pass conflict free.
      __host__ __device__ static constexpr int P(int s)
      {
        int need = 8 * NCH(s + 1) + K_(s + 1);
        need = need < W(s) ? W(s) : need;
        need = (need + 1) & ~1;
        return need + ((2 - need % 16) + 16) % 16;
      }
      // TMA box width.  The x coordinate of a box must be 16-byte aligned (a misaligned one
      // raises "illegal instruction", profiles/microbench/tma_probe.cu), so the two boxes of a
      // block start at x0 - 44 instead of x0 - 41 and carry TMA_SKEW extra columns.
      __host__ __device__ static constexpr int BOXW() { return ((W(0) + TMA_SKEW + 1) / 2 + 3) & ~3; }
      // shared memory map (float offsets)
      __host__ __device__ static constexpr int inraw_floats() { return 2 * 2 * 8 * BOXW(); }
      __host__ __device__ static constexpr int off_out(int s) { return s == 0 ? inraw_floats() : off_out(s - 1) + 8 * P(s - 1); }
      __host__ __device__ static constexpr int off_ring(int s) { return s == 1 ? off_out(4) + 8 * P(4) : off_ring(s - 1) + NB_(s - 1) * 8 * PR(s - 1); }
      __host__ __device__ static constexpr int total_floats() { return off_ring(5) + NB_(5) * 8 * PR(5); }
      __host__ __device__ static constexpr int smem_bytes() { return total_floats() * 4 + 64; }
      // warp items
      __host__ __device__ static constexpr int NWA(int s) { return (NCH(s) + 7) / 8; }   // row pass: 4 row pairs x 8 chunks
      __host__ __device__ static constexpr int NWB(int s) { return (W(s) / 2 + 31) / 32; }  // column pass: 32 column pairs
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
      float taps[5][25];
    };

    struct Ctl
    {
      int x0, y0, y1, Y;
    };

    // ---- packed fp32x2 arithmetic -----------------------------------------------
    __device__ __forceinline__ u64 pack2(float lo, float hi)
    {
      u64 r;
      asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
      return r;
    }
    __device__ __forceinline__ float lo2(u64 v) { return __uint_as_float(static_cast<unsigned>(v)); }
    __device__ __forceinline__ float hi2(u64 v) { return __uint_as_float(static_cast<unsigned>(v >> 32)); }
    __device__ __forceinline__ u64 mul2(u64 a, u64 b)
    {
      u64 r;
      asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
      return r;
    }
    // RN(acc + p) as acc * ONE + p (see the header comment).
    __device__ __forceinline__ u64 add2(u64 acc, u64 p, u64 one)
    {
      u64 r;
      asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(acc), "l"(one), "l"(p));
      return r;
    }

    // ---- TMA / mbarrier ------------------------------------------------------------
    __device__ __forceinline__ unsigned smem_u32(const void* p)
    {
      return static_cast<unsigned>(__cvta_generic_to_shared(p));
    }
    __device__ __forceinline__ void mbar_init(void* bar, int count)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    }
    __device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes)
    {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                   : "memory");
    }
    __device__ __forceinline__ void mbar_wait(void* bar, unsigned parity)
    {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "WAIT_%=:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra DONE_%=;\n"
          "bra WAIT_%=;\n"
          "DONE_%=:\n"
          "}\n" ::"r"(smem_u32(bar)),
          "r"(parity)
          : "memory");
    }
    __device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, void* bar)
    {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
          ::"r"(smem_u32(dst)),
          "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
          : "memory");
    }

    // ---- row pass of stage S: block of G(S-1) (interleaved row pairs) -> ring of stage S ----
    template <int TX, int S>
    __device__ __forceinline__ void row_item(float* sm, const Params& prm, const Ctl& ctl, int t, int wi, int lane)
    {
      using C = Cfg<TX>;
      constexpr int K = K_(S);
      const int rp = lane & 3, ch = wi * 8 + (lane >> 2);
      if (ch >= C::NCH(S))
        return;
      const int u = t - (S - 1);
      const int a = ctl.Y + 8 * u - CS_(S - 1);
      const int ya = a + 2 * rp;
      const int lo = max(0, ctl.y0 - H_(S - 1)), hi = min(prm.h, ctl.y1 + H_(S - 1));
      if (ya + 1 < lo || ya >= hi)
        return;
      const float* in = sm + C::off_out(S - 1) + (rp * C::P(S - 1) + 8 * ch) * 2;
      constexpr int NV = (K + 7 + 1) / 2;  // 16-byte loads: two x positions (both rows) each
      u64 win[2 * NV];
#pragma unroll
      for (int m = 0; m < NV; ++m)
      {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(in + 4 * m);
        win[2 * m] = v.x;
        win[2 * m + 1] = v.y;
      }
      const u64 one = pack2(prm.one, prm.one);
      u64 acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        acc[q] = 0ull;
#pragma unroll
      for (int j = 0; j < K; ++j)
      {
        const u64 kk = pack2(prm.taps[S - 1][j], prm.taps[S - 1][j]);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          acc[q] = add2(acc[q], mul2(win[q + j], kk), one);
      }
      const int slot = (u + 8 * NB_(S)) % NB_(S);
      float* out = sm + C::off_ring(S) + (slot * 8 + 2 * rp) * C::PR(S) + 8 * ch;
      *reinterpret_cast<float4*>(out) = make_float4(lo2(acc[0]), lo2(acc[1]), lo2(acc[2]), lo2(acc[3]));
      *reinterpret_cast<float4*>(out + 4) = make_float4(lo2(acc[4]), lo2(acc[5]), lo2(acc[6]), lo2(acc[7]));
      *reinterpret_cast<float4*>(out + C::PR(S)) = make_float4(hi2(acc[0]), hi2(acc[1]), hi2(acc[2]), hi2(acc[3]));
      *reinterpret_cast<float4*>(out + C::PR(S) + 4) = make_float4(hi2(acc[4]), hi2(acc[5]), hi2(acc[6]), hi2(acc[7]));
    }

    // ---- column pass of stage S --------------------------------------------------------
    // One lane = two adjacent columns (packed), 8 output rows [b, b + 8).
    template <int TX, int S>
    __device__ __forceinline__ void col_item(float* sm, const Params& prm, const Ctl& ctl, int t, int wi, int lane)
    {
      using C = Cfg<TX>;
      constexpr int K = K_(S), c = C_(S), NB = NB_(S), PR = C::PR(S);
      const int u = t - (S - 1);
      const int a = ctl.Y + 8 * u - CS_(S - 1);
      const int b = a - c;
      const int i = 2 * (wi * 32 + lane);     // local column of the low lane
      const int xs = ctl.x0 - H_(S);          // absolute x of local column 0 (even)
      const int x = xs + i;
      const int w = prm.w, h = prm.h;
      const bool active = i < C::W(S) && x >= 0 && x < w;
      const bool central = active && i >= H_(S) && i < H_(S) + TX;
      const bool pair_ok = x + 1 < w;

      // G(S-1) for the D epilogue: issued first, consumed last.
      float2 prev[8];
      if (central)
      {
        const float* gp = prm.G + static_cast<size_t>(S - 1) * prm.layer_stride;
#pragma unroll
        for (int r = 0; r < 8; ++r)
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

      u64 acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        acc[r] = 0ull;
      const float* ring = sm + C::off_ring(S) + i;
      const u64 one = pack2(prm.one, prm.one);
      const bool fast = b - c >= 0 && b + 7 + c <= h - 1;  // no row of the window needs clamping
      if (active)
      {
        if (fast)
        {
          u64 win[K + 7];
          int base[NB];
#pragma unroll
          for (int e = 0; e < NB; ++e)  // base[e]: ring block holding rows of block u - (NB - 1) + e
            base[e] = ((u - (NB - 1) + e + 8 * NB) % NB) * 8 * PR;
#pragma unroll
          for (int q = 0; q < K + 7; ++q)
          {
            const int rel = q - 2 * c;                       // row relative to the current block start
            const int blk = (rel + 8 * NB) / 8 - NB;         // floor(rel / 8), in [-(NB - 1), 0]
            const int rin = rel - 8 * blk;
            win[q] = *reinterpret_cast<const u64*>(ring + base[blk + NB - 1] + rin * PR);
          }
#pragma unroll
          for (int j = 0; j < K; ++j)
          {
            const u64 kk = pack2(prm.taps[S - 1][j], prm.taps[S - 1][j]);
#pragma unroll
            for (int r = 0; r < 8; ++r)
              acc[r] = add2(acc[r], mul2(win[r + j], kk), one);
          }
        }
        else
        {
          // Border steps: rows of the window are clamped to the image
          // (LinearFiltering.hpp:137-142).  Same arithmetic, compact code.
          const int origin = ctl.Y - CS_(S - 1) - 64 * NB;  // row of ring block 0, shifted to keep the modulo positive
#pragma unroll 1
          for (int j = 0; j < K; ++j)
          {
            const u64 kk = pack2(prm.taps[S - 1][j], prm.taps[S - 1][j]);
#pragma unroll
            for (int r = 0; r < 8; ++r)
            {
              const int yq = min(max(b + r - c + j, 0), h - 1);
              const int rr = (yq - origin) % (8 * NB);
              const u64 v = *reinterpret_cast<const u64*>(ring + rr * PR);
              acc[r] = add2(acc[r], mul2(v, kk), one);
            }
          }
        }
      }

      // ---- block of G(S) for the next stage (row-pair interleaved) ----
      if (S < 5)
      {
        float* out = sm + C::off_out(S);
        if (active)
        {
#pragma unroll
          for (int rp = 0; rp < 4; ++rp)
            *reinterpret_cast<float4*>(out + (rp * C::P(S) + i) * 2) =
                make_float4(lo2(acc[2 * rp]), lo2(acc[2 * rp + 1]), hi2(acc[2 * rp]), hi2(acc[2 * rp + 1]));
        }
        // Replicate the border columns over the out-of-image part of the block.
        const int iL = -xs;              // local column of x = 0
        const int iR = w - xs;           // local column of x = w (first one outside)
        const bool left = iL > 0 && (iL >> 6) == wi;                       // this warp item holds x = 0
        const int xr = (w - 1) & ~1;                                       // low lane of the pair holding x = w - 1
        const bool right = iR < C::P(S) && xr - xs >= 0 && ((xr - xs) >> 6) == wi;
        if (left || right)  // warp-uniform
        {
          __syncwarp();
          if (left)
          {
            const int src = (iL >> 1) & 31;
            u64 v[4];
#pragma unroll
            for (int rp = 0; rp < 4; ++rp)
              v[rp] = pack2(__shfl_sync(0xffffffffu, lo2(acc[2 * rp]), src),
                            __shfl_sync(0xffffffffu, lo2(acc[2 * rp + 1]), src));
            for (int ii = lane; ii < iL; ii += 32)
#pragma unroll
              for (int rp = 0; rp < 4; ++rp)
                *reinterpret_cast<u64*>(out + (rp * C::P(S) + ii) * 2) = v[rp];
          }
          if (right)
          {
            const int src = ((xr - xs) >> 1) & 31;
            const bool odd = (w & 1) != 0;  // x = w - 1 is the low lane of its pair
            u64 v[4];
#pragma unroll
            for (int rp = 0; rp < 4; ++rp)
            {
              const float e0 = odd ? lo2(acc[2 * rp]) : hi2(acc[2 * rp]);
              const float e1 = odd ? lo2(acc[2 * rp + 1]) : hi2(acc[2 * rp + 1]);
              v[rp] = pack2(__shfl_sync(0xffffffffu, e0, src), __shfl_sync(0xffffffffu, e1, src));
            }
            for (int ii = iR + lane; ii < C::P(S); ii += 32)
#pragma unroll
              for (int rp = 0; rp < 4; ++rp)
                *reinterpret_cast<u64*>(out + (rp * C::P(S) + ii) * 2) = v[rp];
          }
        }
      }

      // ---- global results: G(S), D(S-1), base of the next octave ----
      if (central)
      {
        float* gs = prm.G + static_cast<size_t>(S) * prm.layer_stride;
        float* ds = prm.D + static_cast<size_t>(S - 1) * prm.layer_stride;
#pragma unroll
        for (int r = 0; r < 8; ++r)
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
          if (S == 2 && prm.nextG != nullptr && (y & 1) == 0)
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
      constexpr int BW = C::BOXW();
      const float* raw = sm + buf * (2 * 8 * BW);
      float* out = sm + C::off_out(0) + rp * C::P(0) * 2;
      const int xs = ctl.x0 - HALO;
      for (int i = lane; i < C::P(0); i += 32)
      {
        int ic = min(i, C::W(0) - 1);
        const int xc = min(max(xs + ic, 0), prm.w - 1);
        ic = min(max(xc - xs, 0), C::W(0) - 1);
        const int jc = ic + TMA_SKEW;  // column inside the two TMA boxes
        const int half = jc >= BW ? 1 : 0;
        const float* p = raw + half * (8 * BW) + (2 * rp) * BW + (jc - half * BW);
        *reinterpret_cast<float2*>(out + 2 * i) = make_float2(p[0], p[BW]);
      }
    }

    template <int TX>
    __global__ void __launch_bounds__(NT, 1)
        fused_octave_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params prm)
    {
      using C = Cfg<TX>;
      extern __shared__ __align__(1024) unsigned char smem_raw[];
      float* sm = reinterpret_cast<float*>(smem_raw);
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + C::total_floats());  // 2 mbarriers
      int* ctr = reinterpret_cast<int*>(bars + 2);                                                // 4 phase counters

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
          int cnt[6];
          int total = 0;
#pragma unroll
          for (int s = 5; s >= 1; --s)
          {
            const int u = t - (s - 1);
            const int a = ctl.Y + 8 * u - CS_(s - 1);
            const int lo = max(0, ctl.y0 - H_(s - 1)), hi = min(h, ctl.y1 + H_(s - 1));
            const bool on = u >= 0 && a + 7 >= lo && a < hi;
            cnt[s] = on ? C::NWA(s) : 0;
            total += cnt[s];
          }
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
            if (id < cnt[5])
              row_item<TX, 5>(sm, prm, ctl, t, id, lane);
            else if ((id -= cnt[5]) < cnt[4])
              row_item<TX, 4>(sm, prm, ctl, t, id, lane);
            else if ((id -= cnt[4]) < cnt[3])
              row_item<TX, 3>(sm, prm, ctl, t, id, lane);
            else if ((id -= cnt[3]) < cnt[2])
              row_item<TX, 2>(sm, prm, ctl, t, id, lane);
            else
              row_item<TX, 1>(sm, prm, ctl, t, id - cnt[2], lane);
          }
          ++phase;
        }
        __syncthreads();

        // ---------------- phase B: column passes + staging of the next input block ----------------
        {
          int cnt[6];
          int total = 0;
#pragma unroll
          for (int s = 5; s >= 1; --s)
          {
            const int u = t - (s - 1);
            const int b = ctl.Y + 8 * u - CS_(s);
            const int lo = max(0, ctl.y0 - H_(s)), hi = min(h, ctl.y1 + H_(s));
            const bool on = u >= 0 && b + 7 >= lo && b < hi;
            cnt[s] = on ? C::NWB(s) : 0;
            total += cnt[s];
          }
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
            if (id < cnt[5])
              col_item<TX, 5>(sm, prm, ctl, t, id, lane);
            else if ((id -= cnt[5]) < cnt[4])
              col_item<TX, 4>(sm, prm, ctl, t, id, lane);
            else if ((id -= cnt[4]) < cnt[3])
              col_item<TX, 3>(sm, prm, ctl, t, id, lane);
            else if ((id -= cnt[3]) < cnt[2])
              col_item<TX, 2>(sm, prm, ctl, t, id, lane);
            else if ((id -= cnt[2]) < cnt[1])
              col_item<TX, 1>(sm, prm, ctl, t, id, lane);
            else
            {
              wait_block(t + 1);
              convert_item<TX>(sm, prm, ctl, (t + 1) & 1, id - cnt[1], lane);
            }
          }
          ++phase;
        }
        __syncthreads();
      }
    }

    // ---- host side ------------------------------------------------------------------------
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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
        for (int j = 0; j < K_(s); ++j)
          prm.taps[s - 1][j] = taps[s].v[j];

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
  static bool downscale_is_even_sampling(int sw, int sh, int dw, int dh)
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
