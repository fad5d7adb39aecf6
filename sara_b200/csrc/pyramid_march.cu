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
//  CTA (128 threads) = strip of 248 output columns x a segment of rows, marching down in
//  blocks of 32 rows.  Per block:
//   1. TMA (cp.async.bulk.tensor.2d + mbarrier, double buffered, two blocks ahead) stages
//      32 rows x (124 + 2c) columns for each half of the strip; out-of-image columns of the
//      first / last strip are overwritten with the replicated border pixel
//      (LinearFiltering.hpp:95-100).
//   2. Row pass: a warp owns 8 rows, a lane one row and one of its four runs of 31 output
//      columns in BOTH half strips (the two halves of its f32x2 registers); fully unrolled,
//      so only the products and sums that exist are issued (triangular ramp-up / ramp-down
//      at the run ends).  Output -> F (row-filtered block).
//   3. Column pass: a thread owns two adjacent columns for the whole segment; 2c sums in
//      flight shift down one slot per row; rows outside the image replicate the border row
//      (LinearFiltering.hpp:137-142) by feeding F's first / last row again.  Epilogue:
//      G(s), D(s-1) (G(s-1) re-read through L2, prefetched eight rows ahead), next octave;
//      8-byte stores.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "fp32x2_tma.cuh"

namespace sb {

  namespace march {

    using namespace fused;

    constexpr int R = 32;   // rows per block: 4 warps x 8 rows in the row pass
    constexpr int T = 31;   // outputs of one run of the row pass (odd: see the bank note below)
    constexpr int HW = 4 * T;   // columns of one half strip (4 runs)
    constexpr int TX = 2 * HW;  // 248 output columns per strip
    constexpr int NP = TX / 2;  // column pairs of the column pass (124 of the 128 threads)
    constexpr int NT = 128;

    // Smallest width >= need with width = 4 (mod 8) floats: a legal TMA box (multiple of 4
    // floats) whose rows r = 0..7 start in eight different groups of four banks.
    __host__ __device__ constexpr int pad_4mod8(int need)
    {
      int b = (need + 3) & ~3;
      while (b % 8 != 4)
        b += 4;
      return b;
    }

    // Shared-memory bank note.  In the row pass a lane is (row r8 = lane & 7 of its warp's
    // eight rows, run g = lane >> 3 of the row's four runs); at step i it reads word
    // row * BW + skew + T g + i (+ the same in the second region).  With BW = 4 (mod 8) the
    // eight rows fall into eight different groups of four banks, and with T odd the four
    // runs take the four banks of a group: 32 lanes, 32 banks, for scalar 4-byte loads that
    // land directly in the two halves of an f32x2 register pair (no packing moves).  The
    // stores into F (pitch PF = 4 (mod 8)) are conflict free for the same reason.
    // NBUF: staging buffers.  2 = TMA runs two blocks ahead, two CTAs per SM; 1 = the next block
    // is requested as soon as the row pass has consumed the current one (it lands under the
    // column pass), three CTAs per SM.
    template <int K, int NBUF>
    struct MC
    {
      static constexpr int c = K / 2;
      static constexpr int skew = (4 - c % 4) % 4;  // TMA x coordinates must be 16-byte aligned
      static constexpr int LEAD = c + skew;         // staged column of a region's first output column
      static constexpr int BW = pad_4mod8(HW + 2 * c + skew);
      static constexpr int PF = pad_4mod8(TX);
      static constexpr int NIN = T + 2 * c;         // inputs of one run
      static constexpr int region_floats = R * BW;
      static constexpr int raw_floats = 2 * region_floats;  // one buffer: two regions (half strips)
      static constexpr int PER_SM = NBUF == 2 ? 2 : 3;
      static constexpr int off_F = NBUF * raw_floats;
      // G(s-1) for the DoG epilogue comes back from L2 through a cp.async ring in shared memory
      // (8 bytes per thread and row); it takes whatever room two resident CTAs leave, and the
      // prefetch distance is its depth minus one.
      static constexpr int off_ring = off_F + R * PF;
      static constexpr int budget_floats = (233472 / PER_SM - 1024 - 64) / 4;
      static constexpr int RING = (budget_floats - off_ring) / (2 * NT) < 12 ? (budget_floats - off_ring) / (2 * NT) : 12;
      static constexpr int PD = RING - 1;
      static constexpr int off_bar = off_ring + RING * 2 * NT;
      static constexpr int smem_bytes = off_bar * 4 + 32;
      static_assert(RING >= 5, "prefetch ring too short");
      static_assert(BW <= 256, "TMA box dimension limit");
      static_assert(LEAD % 4 == 0 && HW % 4 == 0, "aligned TMA coordinates");
      static_assert(T % 2 == 1 && BW % 8 == 4 && PF % 8 == 4, "bank-conflict-free layout");
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

    template <int K, int NBUF>
    __global__ void __launch_bounds__(NT, NBUF == 2 ? 2 : 3)
        march_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params prm)
    {
      using S = MC<K, NBUF>;
      constexpr int c = S::c, BW = S::BW, PF = S::PF, skew = S::skew;
      extern __shared__ __align__(1024) unsigned char smem_raw[];
      float* sm = reinterpret_cast<float*>(smem_raw);
      float* F = sm + S::off_F;
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + S::off_bar);

      const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
      const int w = prm.w, h = prm.h;
      const int x0 = blockIdx.x * TX;
      const int y0 = blockIdx.y * prm.hy;
      const int y1 = min(y0 + prm.hy, h);
      const int r_begin = y0 - c, r_end = y1 + c;  // virtual rows fed to the column pass
      const int NB = (min(r_end, h) - 1 - r_begin) / R + 1;  // blocks that hold a real row
      const int n_reg_on = x0 + HW >= w ? 1 : 2;
      const bool edge = x0 == 0 || x0 + TX + c > w;  // some staged column lies outside the image

      if (tid == 0)
      {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncthreads();

      auto issue = [&](int u) {  // thread 0
        void* bar = &bars[u % NBUF];
        mbar_expect_tx(bar, static_cast<unsigned>(n_reg_on * S::region_floats * 4));
        float* dst = sm + (u % NBUF) * S::raw_floats;
        for (int g = 0; g < n_reg_on; ++g)
          tma_load_2d(dst + g * S::region_floats, &tmap, x0 - S::LEAD + HW * g, r_begin + R * u, bar);
      };
      if (tid == 0)
      {
        issue(0);
        if (NBUF == 2 && NB > 1)
          issue(1);
      }

      const u64 one = pack2(prm.one, prm.one);
      const u64 neg_one = pack2(prm.neg_one, prm.neg_one);

      // ---- row-pass roles ----
      const int r_row = 8 * wq + (lane & 7);  // row of the block
      const int r_run = lane >> 3;            // run of the row (both halves)
      const int r_in = r_row * BW + skew + T * r_run;  // first input word inside a region
      float* const r_out = F + r_row * PF + T * r_run;

      // ---- column-pass roles and state (lives across blocks) ----
      const int xa = x0 + 2 * tid;  // image column of the low half; the high half is the next column
      const bool col_on = tid < NP && xa < w;
      const bool warp_on = x0 + 64 * wq < w && 32 * wq < NP;  // warp-uniform
      u64 A[2 * c];
#pragma unroll
      for (int s = 0; s < 2 * c; ++s)
        A[s] = 0ull;
      const long long pitch_b = static_cast<long long>(prm.pitch) * 4;
      const long long spitch_b = static_cast<long long>(prm.src_pitch) * 4;
      const long long dog_delta = reinterpret_cast<const char*>(prm.dog) - reinterpret_cast<const char*>(prm.out);
      // output pointer of the step about to run (output row y = r - c)
      char* po = reinterpret_cast<char*>(prm.out) + (static_cast<long long>(r_begin - c) * prm.pitch + xa) * 4;
      // G(s-1) prefetch: row yw goes to ring slot slot_w; the step about to run reads slot_r
      constexpr int RING = S::RING, PD = S::PD;
      u64* const ring = reinterpret_cast<u64*>(sm + S::off_ring) + tid;
      const bool dog_on = col_on && prm.dog != nullptr;
      const char* pgw =
          reinterpret_cast<const char*>(prm.src) + (static_cast<long long>(r_begin - c) * prm.src_pitch + xa) * 4;
      int yw = r_begin - c, slot_w = 0, slot_r = 0;
      auto prefetch = [&]() {
        if (dog_on && yw >= 0 && yw < h)
          cp_async8(ring + slot_w * NT, pgw);
        cp_async_commit();
        pgw += spitch_b;
        ++yw;
        slot_w = slot_w + 1 == RING ? 0 : slot_w + 1;
      };
#pragma unroll
      for (int q = 0; q < PD; ++q)
        prefetch();

      for (int u = 0; u < NB; ++u)
      {
        float* raw = sm + (u % NBUF) * S::raw_floats;
        mbar_wait(&bars[u % NBUF], (u / NBUF) & 1);
        const int rb = r_begin + R * u;  // first row of this block

        if (edge)  // CTA-uniform
        {
          // replicate the border pixel into the staged columns that lie outside the image
          if (tid < 2 * R)
          {
            const int g = tid / R, row = tid - g * R;
            float* rr = raw + g * S::region_floats + row * BW;
            const int xs = x0 - S::LEAD + HW * g;  // image column of staged column 0
            const int p0 = -xs;                    // staged column of image column 0
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
        // Inputs are consumed in groups of RG; a group's loads are issued one group ahead and
        // a __syncwarp() closes every group.  The fence is a scheduling device: without it
        // ptxas hoists all the loads of this straight-line code to the top and then, short of
        // registers, re-orders the arithmetic output by output -- serial chains of dependent
        // FFMA2 -- instead of input by input (2c + 1 independent FFMA2 per input).
        {
          const float* in0 = raw + r_in;
          const float* in1 = in0 + S::region_floats;
          constexpr int RG = 4, NG = (S::NIN + RG - 1) / RG;
          u64 acc[T];
          u64 vin[NG * RG];
#pragma unroll
          for (int e = 0; e < RG; ++e)
            vin[e] = pack2(in0[e], in1[e]);
#pragma unroll
          for (int g = 0; g < NG; ++g)
          {
#pragma unroll
            for (int e = 0; e < RG; ++e)  // next group's inputs
              if ((g + 1) * RG + e < S::NIN)
                vin[(g + 1) * RG + e] = pack2(in0[(g + 1) * RG + e], in1[(g + 1) * RG + e]);
#pragma unroll
            for (int e = 0; e < RG; ++e)
            {
              const int i = g * RG + e;  // input i <-> image column (run start) - c + i
              if (i >= S::NIN)
                continue;
              const u64 v = vin[i];
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
                r_out[t] = lo2(acc[t]);
                r_out[t + HW] = hi2(acc[t]);
              }
            }
            __syncwarp();
          }
        }
        __syncthreads();  // F complete, raw buffer free
        if (tid == 0 && u + NBUF < NB)
        {
          if (edge)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(u + NBUF);
        }

        // ---------------- column pass: F -> G(s), D(s-1), next octave ----------------
        if (warp_on)
        {
          const int n_steps = u == NB - 1 ? r_end - rb : R;
          // F row of step k: the block row of image row clamp(rb + k) (border rows replicate)
          const int k_lo = max(-rb, 0), k_hi = min(h - 1 - rb, R - 1);
          // steps whose output row y = rb + k - c lies in [y0, y1)
          const int ks_lo = y0 + c - rb;
          const float* fcol = F + 2 * tid;
#pragma unroll 4
          for (int k = 0; k < n_steps; ++k)
          {
            prefetch();  // G(s-1) of the output row PD steps ahead
            const int fr = min(max(k, k_lo), k_hi);
            const u64 v = *reinterpret_cast<const u64*>(fcol + fr * PF);
            u64 p[c + 1];
#pragma unroll
            for (int m = 0; m <= c; ++m)
              p[m] = mul2(v, pack2(prm.taps[m], prm.taps[m]));
            const u64 E = add2(A[0], p[0], one);
#pragma unroll
            for (int s = 1; s < 2 * c; ++s)
              A[s - 1] = add2(A[s], p[s <= c ? s : 2 * c - s], one);
            A[2 * c - 1] = add2(0ull, p[0], one);

            if (k >= ks_lo)  // uniform; k < ks_hi holds for every step of the loop
            {
              cp_async_wait<PD>();  // the copy issued PD steps ago (this row) has landed
              if (col_on)
              {
                // 8-byte stores: when w is odd the high half of the last pair falls into the row padding
                *reinterpret_cast<u64*>(po) = E;
                if (prm.dog != nullptr)
                  *reinterpret_cast<u64*>(po + dog_delta) = add2(ring[slot_r * NT], E, neg_one);  // RN(E - G(s-1))
                const int y = rb + k - c;
                if (prm.nextG != nullptr && (y & 1) == 0)
                {
                  // downscale(G(s), 2): even rows and columns (xa is even)
                  const int yy = y >> 1, xx = xa >> 1;
                  if (yy < prm.nh && xx < prm.nw)
                    prm.nextG[static_cast<size_t>(yy) * prm.npitch + xx] = lo2(E);
                }
              }
            }
            po += pitch_b;
            slot_r = slot_r + 1 == RING ? 0 : slot_r + 1;
          }
        }
        __syncthreads();  // F free
      }
    }

    template <int K, int NBUF>
    bool launch(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h, int pitch, int nw,
                int nh, int npitch, const Taps& taps, cudaStream_t st)
    {
      using S = MC<K, NBUF>;
      constexpr int per_sm = S::PER_SM;
      static_assert(per_sm * (S::smem_bytes + 1024) <= 233472, "resident CTAs must fit one SM");
      static bool configured = false;
      if (!configured)
      {
        if (cudaFuncSetAttribute(march_kernel<K, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::smem_bytes) !=
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
      const int n_strips = (w + TX - 1) / TX;
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
        const double cta = blocks * R * 1.15 + (hy + 2 * S::c);      // row pass (ramps of the runs) + column pass
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
      march_kernel<K, NBUF><<<dim3(n_strips, n_segs), NT, S::smem_bytes, st>>>(tmap, prm);
      return true;
    }

    template <int K>
    bool launch_k(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h, int pitch,
                  int nw, int nh, int npitch, const Taps& taps, cudaStream_t st)
    {
      static const int nbuf = [] {
        const char* e = getenv("SARA_B200_MARCH_NBUF");
        return e ? atoi(e) : 2;
      }();
      if (nbuf == 1)
        return launch<K, 1>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
      return launch<K, 2>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
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
