// Marching scatter-form kernel: one separable Gaussian increment G(s-1) -> G(s) of one
// octave, with D(s-1) = G(s) - G(s-1) and (for the down-sampled scale) the base of the
// next octave emitted from the same pass.  Default pyramid path.
//
// What it restates (behaviour, not code):
//   apply_row/column_based_filter   ImageProcessing/LinearFiltering.hpp:78-149
//   convolve_array                  ImageProcessing/LinearFiltering.hpp:44-63
//   gaussian_pyramid (one scale)    ImageProcessing/GaussianPyramid.hpp:116-121
//   difference_of_gaussians_pyramid ImageProcessing/GaussianPyramid.cpp:23-51
//   downscale(G(2, o), 2)           ImageProcessing/Resize.cpp:31-83
//
// The arithmetic contract is the reference's: every output is
//   acc = 0; for j = 0 .. K-1: acc = RN(acc + RN(in[x - c + j] * k[j]))
// left to right, multiply and add rounded separately.  The cascade is therefore bound by
// the fp32 pipe (2 K lane-operations per pixel and pass), and this kernel is organised
// around issuing as few fp32 instructions as the contract allows and nothing else:
//
//  * SCATTER form.  A thread keeps the partial sums of the outputs "in flight" in
//    registers and marches along the filter direction; every new input value v is
//    multiplied by the taps and added to the 2c + 1 sums it belongs to.  The inputs of a
//    given output arrive in increasing j, so the order of the additions is the
//    reference's.  Each input is read from shared memory ONCE (the gather form re-reads it
//    K times), which removes the window loads that used to cost as much time as the
//    arithmetic.
//  * SYMMETRIC TAPS.  make_gaussian_kernel gives k[j] == k[K-1-j] bit for bit, so
//    RN(v * k[j]) is the same number for both: c + 1 multiplies per input instead of K
//    (-24 % fp32 instructions).  The additions are untouched.
//  * Packed f32x2 (FMUL2 / FFMA2): the two halves of a register pair are two independent
//    streams (two columns 32 apart), the taps are scalar (uniform-register) operands.
//    RN(acc + p) is issued as fma.rn.f32x2(acc, ONE, p), see fp32x2_tma.cuh.
//
//  CTA = strip of TX = 64 NW output columns x a segment of rows, marching down in blocks
//  of 32 rows.  Per block:
//   1. TMA (cp.async.bulk.tensor.2d + mbarrier, double buffered, two blocks ahead) stages
//      32 rows x (128 + 2c) columns per region of 128 output columns; out-of-image columns
//      of the first / last strip are overwritten with the replicated border pixel
//      (LinearFiltering.hpp:95-100).
//   2. Row pass: lane = row of the block, a warp owns 64 output columns as two packed runs
//      of 32; fully unrolled, so only the products and sums that exist are issued
//      (triangular ramp-up / ramp-down at the run ends).  Output -> F (row-filtered block).
//   3. Column pass: a thread owns two columns (32 apart) for the whole segment; 2c sums in
//      flight shift down one slot per row; rows outside the image replicate the border row
//      (LinearFiltering.hpp:137-142) by feeding F's first / last row again.  Epilogue:
//      G(s), D(s-1) (G(s-1) re-read through L2, prefetched four rows ahead), next octave.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "fp32x2_tma.cuh"

namespace sb {

  namespace march {

    using namespace fused;

    constexpr int R = 32;   // rows per block = lanes of a row-pass warp
    constexpr int T = 32;   // outputs of one packed run in the row pass
    constexpr int WC = 64;  // output columns per warp
    constexpr int PD = 4;   // prefetch distance (rows) of G(s-1) in the column pass

    // Smallest box width >= need that keeps 16-byte shared-memory loads of 8 consecutive rows
    // conflict free (width = 4 mod 8 floats) and is a legal TMA box (multiple of 4 floats).
    __host__ __device__ constexpr int box_width(int need)
    {
      int b = (need + 3) & ~3;
      while (b % 8 != 4)
        b += 4;
      return b;
    }

    template <int K, int NW>
    struct MC
    {
      static constexpr int c = K / 2;
      static constexpr int skew = (4 - c % 4) % 4;  // TMA x coordinates must be 16-byte aligned
      static constexpr int LEAD = c + skew;         // staged column of the region's first output column
      static constexpr int TX = NW * WC;
      static constexpr int NREG = (NW + 1) / 2;     // TMA regions (128 output columns each; 64 when NW == 1)
      static constexpr int RW = NW >= 2 ? 128 : 64;
      static constexpr int BW = box_width(RW + 2 * c + skew);
      static constexpr int PF = TX + 1;             // odd pitch: lanes = rows store conflict free
      static constexpr int NIN = T + 2 * c;         // inputs of one run
      static constexpr int NCH = (skew + NIN + 3) / 4;
      static constexpr int region_floats = R * BW;
      static constexpr int raw_floats = NREG * region_floats;  // one buffer
      static constexpr int off_F = 2 * raw_floats;
      static constexpr int off_bar = off_F + ((R * PF + 3) & ~3);
      static constexpr int smem_bytes = off_bar * 4 + 32;
      static_assert(BW <= 256, "TMA box dimension limit");
      static_assert(LEAD % 4 == 0, "aligned TMA coordinates");
      static_assert((NW >= 2 ? 64 : 0) + 32 + 4 * NCH <= BW, "row-pass chunks stay inside the staged region");
      static_assert(2 * c <= R - PD, "ramp-up fits the first block");
    };

    struct Params
    {
      const float* src;  // G(s-1, o)
      float* out;        // G(s, o)
      float* dog;        // D(s-1, o) or nullptr
      float* nextG;      // G(0, o + 1) or nullptr
      int w, h, pitch, src_pitch;
      int nw, nh, npitch;
      int hy;            // rows per segment
      float one, neg_one;
      float taps[16];    // the c + 1 distinct taps k[0 .. c]
    };

    template <int K, int NW>
    __global__ void __launch_bounds__(NW * 32, NW == 4 ? 2 : 4)
        march_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params prm)
    {
      using S = MC<K, NW>;
      constexpr int c = S::c, BW = S::BW, PF = S::PF, skew = S::skew;
      extern __shared__ __align__(1024) unsigned char smem_raw[];
      float* sm = reinterpret_cast<float*>(smem_raw);
      float* F = sm + S::off_F;
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + S::off_bar);

      const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
      const int w = prm.w, h = prm.h;
      const int x0 = blockIdx.x * S::TX;
      const int y0 = blockIdx.y * prm.hy;
      const int y1 = min(y0 + prm.hy, h);
      const int r_begin = y0 - c, r_end = y1 + c;  // virtual rows fed to the column pass
      const int NB = (min(r_end, h) - 1 - r_begin) / R + 1;  // blocks that hold a real row
      const int n_reg_on = (S::NREG == 2 && x0 + 128 >= w) ? 1 : S::NREG;
      const bool edge = x0 == 0 || x0 + S::TX + c > w;  // some staged column lies outside the image

      if (tid == 0)
      {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncthreads();

      auto issue = [&](int u) {  // thread 0
        void* bar = &bars[u & 1];
        mbar_expect_tx(bar, static_cast<unsigned>(n_reg_on * S::region_floats * 4));
        float* dst = sm + (u & 1) * S::raw_floats;
        for (int g = 0; g < n_reg_on; ++g)
          tma_load_2d(dst + g * S::region_floats, &tmap, x0 - S::LEAD + 128 * g, r_begin + R * u, bar);
      };
      if (tid == 0)
      {
        issue(0);
        if (NB > 1)
          issue(1);
      }

      const u64 one = pack2(prm.one, prm.one);
      const u64 neg_one = pack2(prm.neg_one, prm.neg_one);
      const bool warp_on = x0 + WC * wq < w;  // warp-uniform: this warp's columns exist

      // ---- column-pass state (lives across blocks) ----
      const int fc = WC * wq + lane;         // F column of the low half; the high half is 32 further
      const int xa = x0 + fc;                // image column of the low half
      const bool lo_ok = xa < w, hi_ok = xa + 32 < w;
      u64 A[2 * c];
#pragma unroll
      for (int s = 0; s < 2 * c; ++s)
        A[s] = 0ull;
      const float* gsrc = prm.src + xa;
      auto load_prev = [&](int y) {  // G(s-1)(xa, y), G(s-1)(xa + 32, y); y clamped (unused when outside)
        const size_t o = static_cast<size_t>(min(max(y, 0), h - 1)) * prm.src_pitch;
        const float a = lo_ok ? __ldg(gsrc + o) : 0.f;
        const float b = hi_ok ? __ldg(gsrc + o + 32) : 0.f;
        return pack2(a, b);
      };
      u64 gq[PD];
#pragma unroll
      for (int q = 0; q < PD; ++q)
        gq[q] = warp_on ? load_prev(r_begin - c + q) : 0ull;

      for (int u = 0; u < NB; ++u)
      {
        float* raw = sm + (u & 1) * S::raw_floats;
        mbar_wait(&bars[u & 1], (u >> 1) & 1);
        const int rb = r_begin + R * u;  // first row of this block

        if (edge)  // CTA-uniform
        {
          // replicate the border pixel into the staged columns that lie outside the image
          if (tid < S::NREG * R)
          {
            const int g = tid / R, row = tid - g * R;
            float* rr = raw + g * S::region_floats + row * BW;
            const int xs = x0 - S::LEAD + 128 * g;  // image column of staged column 0
            const int p0 = -xs;                     // staged column of image column 0
            if (xs < 0 && p0 < BW)
            {
              const float v = rr[p0];
              for (int p = 0; p < p0; ++p)
                rr[p] = v;
            }
            const int pw = w - 1 - xs;  // staged column of image column w - 1
            if (pw >= 0 && pw < BW - 1)
            {
              const float v = rr[pw];
              const int pe = min(pw + c, BW - 1);
              for (int p = pw + 1; p <= pe; ++p)
                rr[p] = v;
            }
          }
          __syncthreads();
        }

        // ---------------- row pass: raw block -> F ----------------
        if (warp_on)
        {
          const float* rlo = raw + (wq >> 1) * S::region_floats + lane * BW + WC * (wq & 1);
          float* fo = F + lane * PF + WC * wq;
          u64 acc[T];
#pragma unroll
          for (int ch = 0; ch < S::NCH; ++ch)
          {
            const float4 a4 = *reinterpret_cast<const float4*>(rlo + 4 * ch);
            const float4 b4 = *reinterpret_cast<const float4*>(rlo + 32 + 4 * ch);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
            {
              const int i = 4 * ch + e - skew;  // input index of the run: image column (run start) - c + i
              if (i < 0 || i >= S::NIN)
                continue;
              const u64 v = pack2_once(av[e], bv[e]);
              const int jlo = i - T + 1 > 0 ? i - T + 1 : 0;
              const int jhi = i < K - 1 ? i : K - 1;
              u64 p[c + 1];
#pragma unroll
              for (int m = 0; m <= c; ++m)
              {
                const bool need = (m >= jlo && m <= jhi) || (K - 1 - m >= jlo && K - 1 - m <= jhi);
                if (need)
                  p[m] = mul2(v, pack2(prm.taps[m], prm.taps[m]));
              }
#pragma unroll
              for (int j = jhi; j >= jlo; --j)
              {
                const int t = i - j;
                const int m = j <= c ? j : K - 1 - j;
                acc[t] = add2(j == 0 ? 0ull : acc[t], p[m], one);
              }
              if (i >= 2 * c)
              {
                const int t = i - 2 * c;
                fo[t] = lo2(acc[t]);
                fo[t + 32] = hi2(acc[t]);
              }
            }
          }
        }
        __syncthreads();  // F complete, raw buffer free
        if (tid == 0 && u + 2 < NB)
        {
          if (edge)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(u + 2);
        }

        // ---------------- column pass: F -> G(s), D(s-1), next octave ----------------
        if (warp_on)
        {
          const int n_steps = u == NB - 1 ? ((r_end - rb + PD - 1) / PD) * PD : R;
          for (int k0 = 0; k0 < n_steps; k0 += PD)
          {
#pragma unroll
            for (int q = 0; q < PD; ++q)
            {
              const int r = rb + k0 + q;
              const int fr = min(max(r, 0), h - 1) - rb;
              const float* fp = F + fr * PF + fc;
              const u64 v = pack2_once(fp[0], fp[32]);
              u64 p[c + 1];
#pragma unroll
              for (int m = 0; m <= c; ++m)
                p[m] = mul2(v, pack2(prm.taps[m], prm.taps[m]));
              const u64 E = add2(A[0], p[0], one);
#pragma unroll
              for (int s = 1; s < 2 * c; ++s)
                A[s - 1] = add2(A[s], p[s <= c ? s : 2 * c - s], one);
              A[2 * c - 1] = add2(0ull, p[0], one);

              const int y = r - c;
              if (y >= y0 && y < y1)
              {
                const size_t o = static_cast<size_t>(y) * prm.pitch + xa;
                if (lo_ok)
                  prm.out[o] = lo2(E);
                if (hi_ok)
                  prm.out[o + 32] = hi2(E);
                if (prm.dog != nullptr)
                {
                  const u64 d = add2(gq[q], E, neg_one);  // RN(E - G(s-1)): gq * (-1) + E
                  if (lo_ok)
                    prm.dog[o] = lo2(d);
                  if (hi_ok)
                    prm.dog[o + 32] = hi2(d);
                }
                if (prm.nextG != nullptr && ((y | lane) & 1) == 0)
                {
                  // downscale(G(s), 2): even rows and columns (xa is even when the lane is)
                  const int yy = y >> 1, xx = xa >> 1;
                  if (yy < prm.nh)
                  {
                    float* pn = prm.nextG + static_cast<size_t>(yy) * prm.npitch;
                    if (lo_ok && xx < prm.nw)
                      pn[xx] = lo2(E);
                    if (hi_ok && xx + 16 < prm.nw)
                      pn[xx + 16] = hi2(E);
                  }
                }
              }
              gq[q] = load_prev(y + PD);
            }
          }
        }
        __syncthreads();  // F free
      }
    }

    template <int K, int NW>
    bool launch(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h, int pitch, int nw,
                int nh, int npitch, const Taps& taps, cudaStream_t st)
    {
      using S = MC<K, NW>;
      constexpr int per_sm = NW == 4 ? 2 : 4;
      static_assert(per_sm * (S::smem_bytes + 1024) <= 233472, "resident CTAs must fit one SM");
      static bool configured = false;
      if (!configured)
      {
        if (cudaFuncSetAttribute(march_kernel<K, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::smem_bytes) !=
            cudaSuccess)
          return false;
        configured = true;
      }
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
      prm.src = src;
      prm.out = dst;
      prm.dog = dog;
      prm.nextG = nextG;
      prm.w = w;
      prm.h = h;
      prm.pitch = pitch;
      prm.src_pitch = src_pitch;
      prm.nw = nw;
      prm.nh = nh;
      prm.npitch = npitch;
      prm.one = 1.f;
      prm.neg_one = -1.f;
      for (int j = 0; j <= S::c; ++j)
        prm.taps[j] = taps.v[j];

      // Segments: fill the machine (148 SMs x resident CTAs) while keeping segments tall --
      // every segment pays 2c warm-up rows and whole 32-row blocks in the row pass.
      const int n_strips = (w + S::TX - 1) / S::TX;
      const int slots = 148 * per_sm;
      int best_segs = 1;
      double best_cost = 1e30;
      const int max_segs = (h + 15) / 16;
      for (int n = 1; n <= max_segs && n_strips * n <= 2 * slots; ++n)
      {
        const int hy = (h + n - 1) / n;
        if ((h + hy - 1) / hy != n)
          continue;
        const int blocks = (hy + 2 * S::c + R - 1) / R;
        const double cta = blocks * R * 1.15 + (hy + 2 * S::c);      // row pass (T = 32 ramps) + column pass
        const int ctas = n_strips * n;
        const int per = (ctas + 147) / 148;                            // CTAs on the busiest SM
        // one CTA alone on an SM leaves its barrier bubbles uncovered
        const double cost = cta * per * (per == 1 ? 1.25 : 1.0) * (per > per_sm ? 1.5 : 1.0);
        if (cost < best_cost)
        {
          best_cost = cost;
          best_segs = n;
        }
      }
      static const int force = [] {
        const char* e = getenv("SARA_B200_MARCH_SEGS");
        return e ? atoi(e) : 0;
      }();
      int n_segs = force > 0 ? force : best_segs;
      int hy = (h + n_segs - 1) / n_segs;
      n_segs = (h + hy - 1) / hy;
      prm.hy = hy;
      march_kernel<K, NW><<<dim3(n_strips, n_segs), NW * 32, S::smem_bytes, st>>>(tmap, prm);
      return true;
    }

    template <int K>
    bool launch_k(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h, int pitch,
                  int nw, int nh, int npitch, const Taps& taps, cudaStream_t st)
    {
      if (w > 640)
        return launch<K, 4>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
      if (w > 64)
        return launch<K, 2>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
      return launch<K, 1>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    }

  }  // namespace march

  bool march_kernel_supported(const Taps& taps)
  {
    const int n = taps.n;
    if (!(n == 11 || n == 13 || n == 17 || n == 21 || n == 25) || fused::encode_fn() == nullptr)
      return false;
    for (int j = 0; j < n / 2; ++j)  // product sharing needs bit-symmetric taps
      if (memcmp(&taps.v[j], &taps.v[n - 1 - j], sizeof(float)) != 0)
        return false;
    return true;
  }

  // Same contract as launch_stage (pyramid_stage.cu).
  bool launch_march(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h, int pitch,
                    int nw, int nh, int npitch, const Taps& taps, cudaStream_t st)
  {
    if ((src_pitch & 3) != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0)
      return false;
    switch (taps.n)
    {
    case 11: return march::launch_k<11>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 13: return march::launch_k<13>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 17: return march::launch_k<17>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 21: return march::launch_k<21>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 25: return march::launch_k<25>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    default: return false;
    }
  }

}  // namespace sb
