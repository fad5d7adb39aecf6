// Single-stage marching kernel: one separable Gaussian increment G(s-1) -> G(s) of one
// octave, with the DoG layer D(s-1) = G(s) - G(s-1) and (for the down-sampled scale)
// the base of the next octave emitted from the same pass.
//
// What it restates (behaviour, not code):
//   apply_row/column_based_filter   ImageProcessing/LinearFiltering.hpp:78-149
//   gaussian_pyramid (one scale)    ImageProcessing/GaussianPyramid.hpp:116-121
//   difference_of_gaussians_pyramid ImageProcessing/GaussianPyramid.cpp:23-51
//   downscale(G(2, o), 2)           ImageProcessing/Resize.cpp:31-83
//
// Why a per-stage kernel next to the fused octave kernel: with the reference's
// arithmetic (separate multiply and add, no FMA) the cascade is bound by the fp32
// pipe, not by HBM (profiles/microbench/r01_fp32_issue_b200.txt), so what matters is
// keeping that pipe busy.  Here every thread of a launch runs the SAME tap count:
// the work is statically balanced (no scheduler, no recomputed cascade halo), the
// code is a few hundred instructions, and a CTA needs ~75 KB of shared memory, so
// three CTAs share an SM and cover each other's barriers.  The price is that G(s-1)
// is read back (through L2) by the next launch.
//
//  * CTA = strip of 128 output columns x a segment of rows, marching 16 rows a step.
//  * TMA (cp.async.bulk.tensor.2d + mbarrier) stages 16-row blocks of G(s-1), two
//    blocks ahead, into a ring of four buffers (two of them are still needed for D).
//  * Row pass: a thread filters 4 columns of 2 rows (the rows ride in the halves of
//    f32x2 registers); its input is a row-pair interleaved copy of the block with the
//    image border replicated.  Column pass: a thread filters 4 rows of 2 adjacent
//    columns out of a ring of row-filtered rows; window rows come through a per-step
//    table that clamps them to the image (border replication) and wraps the ring.
//  * Arithmetic: RN(acc + RN(b * k)) per tap, left to right from +0, as
//    DO::Sara::convolve_array; the add is fma.rn.f32x2(acc, ONE, p) (see pyramid_fused.cu).
#include "common.cuh"
#include "fp32x2_tma.cuh"

namespace sb {

  namespace stage {

    using namespace fused;

    constexpr int TX = 128;  // output columns per strip
    constexpr int R = 16;    // rows per step
    constexpr int NT = 256;  // threads per CTA

    __host__ __device__ constexpr int round_to(int v, int mod, int rem) { return v + ((rem - v % mod) + mod) % mod; }

    template <int K>
    struct SC
    {
      static constexpr int c = K / 2;
      static constexpr int skew = (4 - c % 4) % 4;         // TMA x coordinates must be 16-byte aligned
      static constexpr int LEAD = c + skew;                // column of x0 inside a staged block
      static constexpr int BW = (TX + 2 * c + skew + 3) & ~3;  // TMA box width
      static constexpr int NPOS = TX + 2 * c;              // x positions a row pass reads: x0 - c + i
      static constexpr int PI = round_to(TX + K + 1, 4, 2);   // pitch of the interleaved block (x positions)
      static constexpr int PRR = 136;                      // ring pitch, = 8 (mod 16)
      static constexpr int NBR = (2 * c + R - 1) / R + 1;  // 16-row blocks in the ring
      static constexpr int NTAB = K + R;                   // ring-row table entries (K + R - 1 used)
      // shared memory map (floats)
      static constexpr int off_raw = 0;                    // [4][R][BW]
      static constexpr int off_ini = 4 * R * BW;           // [R / 2][PI][2]
      static constexpr int off_ring = off_ini + (R / 2) * PI * 2;  // [NBR * R][PRR]
      static constexpr int off_tab = off_ring + NBR * R * PRR;     // [NTAB] ints
      static constexpr int total = off_tab + ((NTAB + 3) & ~3);
      static constexpr int smem_bytes = total * 4 + 64;    // + 4 mbarriers
      static_assert(BW <= 256, "TMA box dimension limit");
      static_assert(LEAD % 4 == 0 && TX % 4 == 0, "aligned TMA coordinates");
    };

    struct Params
    {
      float* out;         // G(s, o)
      float* dog;         // D(s-1, o) or nullptr
      float* nextG;       // G(0, o + 1) or nullptr
      int w, h, pitch;
      int nw, nh, npitch;
      int hy;             // rows per segment
      float one;
      float taps[28];
    };

    template <int K>
    __global__ void __launch_bounds__(NT, 3)
        stage_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params prm)
    {
      using S = SC<K>;
      constexpr int c = S::c, BW = S::BW, PI = S::PI, PRR = S::PRR, NBR = S::NBR;
      extern __shared__ __align__(1024) unsigned char smem_raw[];
      float* sm = reinterpret_cast<float*>(smem_raw);
      float* raw = sm + S::off_raw;
      float* ini = sm + S::off_ini;
      float* ring = sm + S::off_ring;
      int* rowtab = reinterpret_cast<int*>(sm + S::off_tab);
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + S::total);

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      const int w = prm.w, h = prm.h;
      const int x0 = blockIdx.x * TX;
      const int y0 = blockIdx.y * prm.hy;
      const int y1 = min(y0 + prm.hy, h);
      const int Y = y0 - c;  // first row of input block 0
      const int T = (y1 - y0 + 2 * c + R - 1) / R;
      const int in_hi = min(h, y1 + c);
      constexpr unsigned kBlockBytes = R * BW * 4u;

      if (tid == 0)
      {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncthreads();

      // Input block u = rows [Y + 16 u, +16); needed while it starts above the last row read.
      auto block_needed = [&](int u) { return u < T && Y + R * u < in_hi; };
      auto issue = [&](int u) {  // thread 0
        mbar_expect_tx(&bars[u & 3], kBlockBytes);
        tma_load_2d(raw + (u & 3) * (R * BW), &tmap, x0 - S::LEAD, Y + R * u, &bars[u & 3]);
      };
      // Row-pair interleaved copy of block u with the left/right image border replicated
      // (LinearFiltering.hpp:95-100): position i <-> x = x0 - c + i.
      // The clamped source column of each position is the same at every step.
      constexpr int NCV = (PI + 31) / 32;
      int cv_src[NCV];
#pragma unroll
      for (int k = 0; k < NCV; ++k)
      {
        const int i = lane + 32 * k;
        const int xc = min(max(x0 - c + min(i, S::NPOS - 1), 0), w - 1);
        cv_src[k] = min(max(xc - (x0 - S::LEAD), 0), BW - 1);
      }
      auto convert = [&](int u) {
        mbar_wait(&bars[u & 3], (u >> 2) & 1);
        const float* src = raw + (u & 3) * (R * BW) + (2 * warp) * BW;  // warp = row pair
        float* dst = ini + warp * PI * 2 + 2 * lane;
#pragma unroll
        for (int k = 0; k < NCV; ++k)
          if (lane + 32 * k < PI)
            *reinterpret_cast<float2*>(dst + 64 * k) = make_float2(src[cv_src[k]], src[BW + cv_src[k]]);
      };

      if (tid == 0)
      {
        if (block_needed(0))
          issue(0);
        if (block_needed(1))
          issue(1);
      }
      if (block_needed(0))
        convert(0);
      __syncthreads();

      // fixed roles
      const int r_ch = (warp & 3) * 8 + ((lane & 3) | ((lane >> 1) & 4));       // 4-column chunk, 0..31
      const int r_rp = (warp >> 2) * 4 + (((lane >> 2) & 1) | ((lane >> 3) & 2));  // row pair, 0..7
      const int c_i = 2 * (tid & 63);  // low column of the pair
      const int c_q = tid >> 6;        // which 4 rows of the 16
      const int x = x0 + c_i;
      const bool x_ok = x < w, pair_ok = x + 1 < w;
      const u64 one = pack2(prm.one, prm.one);

      for (int t = 0; t < T; ++t)
      {
        if (tid == 0 && block_needed(t + 2))
          issue(t + 2);
        const int a = Y + R * t;  // first row of this step's input block

        // ---------------- row pass: block t -> ring ----------------
        if (tid < S::NTAB)
        {
          // ring row (float offset) of image row clamp(a - 2c + m): window rows of this step's column pass
          const int yc = min(max(a - 2 * c + tid, 0), h - 1);
          const int rel = max(yc - Y, 0);
          rowtab[tid] = (((rel >> 4) % NBR) * R + (rel & 15)) * PRR;
        }
        {
          const ulonglong2* in = reinterpret_cast<const ulonglong2*>(ini + (r_rp * PI + 4 * r_ch) * 2);
          constexpr int NV = (K + 3 + 1) / 2;
          u64 win[2 * NV];
#pragma unroll
          for (int m = 0; m < NV; ++m)
          {
            const ulonglong2 v = in[m];
            win[2 * m] = v.x;
            win[2 * m + 1] = v.y;
          }
          u64 acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
          for (int j = 0; j < K; ++j)
          {
            const u64 kk = pack2(prm.taps[j], prm.taps[j]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              acc[q] = add2(acc[q], mul2(win[q + j], kk), one);
          }
          float* o = ring + ((t % NBR) * R + 2 * r_rp) * PRR + 4 * r_ch;
          *reinterpret_cast<float4*>(o) = make_float4(lo2(acc[0]), lo2(acc[1]), lo2(acc[2]), lo2(acc[3]));
          *reinterpret_cast<float4*>(o + PRR) = make_float4(hi2(acc[0]), hi2(acc[1]), hi2(acc[2]), hi2(acc[3]));
        }
        __syncthreads();

        // ---------------- column pass: ring -> G(s), D(s-1), next octave ----------------
        {
          const int yb = a - c + 4 * c_q;  // first output row of this thread
          const float* col = ring + c_i;
          const int* tab = rowtab + 4 * c_q;
          u64 win[K + 3];
          // ring-row offsets: 16-byte broadcast loads (tab is 16-byte aligned: 4 * c_q entries in)
          int toff[(K + 3 + 3) & ~3];
#pragma unroll
          for (int n4 = 0; n4 < (K + 3 + 3) / 4; ++n4)
          {
            const int4 tv = *reinterpret_cast<const int4*>(tab + 4 * n4);
            toff[4 * n4] = tv.x;
            toff[4 * n4 + 1] = tv.y;
            toff[4 * n4 + 2] = tv.z;
            toff[4 * n4 + 3] = tv.w;
          }
#pragma unroll
          for (int n = 0; n < K + 3; ++n)
            win[n] = *reinterpret_cast<const u64*>(col + toff[n]);
          u64 acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
          for (int j = 0; j < K; ++j)
          {
            const u64 kk = pack2(prm.taps[j], prm.taps[j]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              acc[q] = add2(acc[q], mul2(win[q + j], kk), one);
          }
          if (x_ok && yb + 3 >= y0 && yb < y1)
          {
            // G(s-1) at (x, y) is still in the staging ring (blocks t - 1 and t): rel = y - Y
            const int rel0 = yb - Y;
            float* po = prm.out + static_cast<size_t>(yb) * prm.pitch + x;
            float* pd = prm.dog != nullptr ? prm.dog + static_cast<size_t>(yb) * prm.pitch + x : nullptr;
            if (pair_ok && yb >= y0 && yb + 3 < y1)
            {
              // common case: four full rows, two columns each
#pragma unroll
              for (int r = 0; r < 4; ++r)
              {
                const int rel = rel0 + r;
                const float2 pv = *reinterpret_cast<const float2*>(raw + ((rel >> 4) & 3) * (R * BW) + (rel & 15) * BW +
                                                                   S::LEAD + c_i);
                const float g0 = lo2(acc[r]), g1 = hi2(acc[r]);
                *reinterpret_cast<float2*>(po + static_cast<size_t>(r) * prm.pitch) = make_float2(g0, g1);
                if (pd != nullptr)
                  *reinterpret_cast<float2*>(pd + static_cast<size_t>(r) * prm.pitch) =
                      make_float2(__fsub_rn(g0, pv.x), __fsub_rn(g1, pv.y));
              }
              if (prm.nextG != nullptr)
              {
                // downscale(G(s), 2): even rows and columns (x is even)
                const int xx = x >> 1;
                if (xx < prm.nw)
#pragma unroll
                  for (int r = 0; r < 4; ++r)
                    if (((yb + r) & 1) == 0 && ((yb + r) >> 1) < prm.nh)
                      prm.nextG[static_cast<size_t>((yb + r) >> 1) * prm.npitch + xx] = lo2(acc[r]);
              }
            }
            else
            {
#pragma unroll
              for (int r = 0; r < 4; ++r)
              {
                const int y = yb + r;
                if (y < y0 || y >= y1)
                  continue;
                const size_t o = static_cast<size_t>(r) * prm.pitch;
                const float g0 = lo2(acc[r]), g1 = hi2(acc[r]);
                const int rel = rel0 + r;
                const float2 pv = *reinterpret_cast<const float2*>(raw + ((rel >> 4) & 3) * (R * BW) + (rel & 15) * BW +
                                                                   S::LEAD + c_i);
                if (pair_ok)
                {
                  *reinterpret_cast<float2*>(po + o) = make_float2(g0, g1);
                  if (pd != nullptr)
                    *reinterpret_cast<float2*>(pd + o) = make_float2(__fsub_rn(g0, pv.x), __fsub_rn(g1, pv.y));
                }
                else
                {
                  po[o] = g0;
                  if (pd != nullptr)
                    pd[o] = __fsub_rn(g0, pv.x);
                }
                if (prm.nextG != nullptr && (y & 1) == 0)
                {
                  const int xx = x >> 1, yy = y >> 1;
                  if (xx < prm.nw && yy < prm.nh)
                    prm.nextG[static_cast<size_t>(yy) * prm.npitch + xx] = g0;
                }
              }
            }
          }
        }
        if (block_needed(t + 1))
          convert(t + 1);
        __syncthreads();
      }
    }

    template <int K>
    bool launch(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h, int pitch, int nw,
                int nh, int npitch, const Taps& taps, cudaStream_t st)
    {
      using S = SC<K>;
      static_assert(3 * (S::smem_bytes + 1024) <= 233472, "three CTAs must fit one SM");
      if (cudaFuncSetAttribute(stage_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::smem_bytes) !=
          cudaSuccess)
        return false;
      EncodeTiledFn enc = encode_fn();
      if (!enc)
        return false;
      CUtensorMap tmap;
      const cuuint64_t dims[2] = {static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h)};
      const cuuint64_t strides[1] = {static_cast<cuuint64_t>(src_pitch) * sizeof(float)};
      const cuuint32_t box[2] = {static_cast<cuuint32_t>(S::BW), static_cast<cuuint32_t>(R)};
      const cuuint32_t estr[2] = {1u, 1u};
      if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(src), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
      Params prm{};
      prm.out = dst;
      prm.dog = dog;
      prm.nextG = nextG;
      prm.w = w;
      prm.h = h;
      prm.pitch = pitch;
      prm.nw = nw;
      prm.nh = nh;
      prm.npitch = npitch;
      prm.one = 1.f;
      for (int j = 0; j < K; ++j)
        prm.taps[j] = taps.v[j];
      // Fill the machine: 148 SMs x 3 resident CTAs; segments are multiples of 16 rows.
      const int n_strips = (w + TX - 1) / TX;
      int n_segs = (148 * 3) / n_strips;
      n_segs = n_segs < 1 ? 1 : n_segs;
      int hy = (h + n_segs - 1) / n_segs;
      hy = (hy + R - 1) / R * R;
      n_segs = (h + hy - 1) / hy;
      prm.hy = hy;
      stage_kernel<K><<<dim3(n_strips, n_segs), NT, S::smem_bytes, st>>>(tmap, prm);
      return true;
    }

  }  // namespace stage

  bool stage_kernel_supported(int n_taps)
  {
    return (n_taps == 11 || n_taps == 13 || n_taps == 17 || n_taps == 21 || n_taps == 25) &&
           fused::encode_fn() != nullptr;
  }

  // One Gaussian increment with the optional DoG / next-octave outputs.  `pitch` is shared by
  // dst and dog; src (16-byte aligned, src_pitch a multiple of 4 floats: TMA) has its own.
  // Returns false if the kernel could not be set up.
  bool launch_stage(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h, int pitch,
                    int nw, int nh, int npitch, const Taps& taps, cudaStream_t st)
  {
    if ((src_pitch & 3) != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0)
      return false;
    switch (taps.n)
    {
    case 11: return stage::launch<11>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 13: return stage::launch<13>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 17: return stage::launch<17>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 21: return stage::launch<21>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 25: return stage::launch<25>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    default: return false;
    }
  }

}  // namespace sb
