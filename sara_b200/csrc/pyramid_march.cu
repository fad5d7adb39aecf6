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
#include <algorithm>
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
    constexpr int NTC = 128;    // threads of one warp group (row warps / column warps)
    constexpr int NT = 2 * NTC;

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
    template <int K>
    struct MC
    {
      static constexpr int c = K / 2;
      static constexpr int skew = (4 - c % 4) % 4;  // TMA x coordinates must be 16-byte aligned
      static constexpr int LEAD = c + skew;         // staged column of a region's first output column
      static constexpr int BW = pad_4mod8(HW + 2 * c + skew);
      static constexpr int PF = pad_4mod8(TX);
      static constexpr int NIN = T + 2 * c;         // inputs of one run
      static constexpr int region_floats = R * BW;
      static constexpr int raw_floats = 2 * region_floats;  // the staging buffer: two regions (half strips)
      static constexpr int PER_SM = 2;
      static constexpr int off_F = raw_floats;      // two F buffers (row warps fill one while column warps drain the other)
      // G(s-1) for the DoG epilogue comes back from L2 through a cp.async ring in shared memory
      // (8 bytes per column thread and row); it takes whatever room two resident CTAs leave,
      // and the prefetch distance is its depth minus one.
      static constexpr int off_ring = off_F + 2 * R * PF;
      static constexpr int budget_floats = (233472 / PER_SM - 1024 - 128) / 4;
      static constexpr int RING = 8;
      static constexpr int PD = RING - 1;
      static constexpr int off_bar = off_ring + RING * 2 * NTC;
      static constexpr int smem_bytes = off_bar * 4 + 64;
      static_assert(off_bar + 16 <= budget_floats, "two CTAs per SM");
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

    __device__ __forceinline__ void mbar_arrive(void* bar)
    {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    __device__ __forceinline__ void row_group_sync()  // the 128 threads of the row warps
    {
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }

    // Warp-specialised: warps 0-3 run the row pass of block u + 1 while warps 4-7 run the column
    // pass of block u (hand-over through two F buffers and mbarriers), so that four warps per
    // scheduler are resident with two CTAs per SM -- enough to cover the shared-memory, barrier
    // and instruction-fetch bubbles that two warps per scheduler left exposed (profiles/README.md).
    // DOG: also emit D(s-1) = G(s) - G(s-1); NEXT: also emit the base of the next octave.
    template <int K, bool DOG, bool NEXT>
    __global__ void __launch_bounds__(NT, 2)
        march_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params prm)
    {
      using S = MC<K>;
      constexpr int c = S::c, BW = S::BW, PF = S::PF, skew = S::skew;
      extern __shared__ __align__(1024) unsigned char smem_raw[];
      float* sm = reinterpret_cast<float*>(smem_raw);
      float* raw = sm;
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + S::off_bar);
      unsigned long long* bar_tma = bars;          // staging buffer filled (TMA transaction bytes)
      unsigned long long* bar_full = bars + 1;     // [2] F buffer written by the row warps
      unsigned long long* bar_empty = bars + 3;    // [2] F buffer drained by the column warps

      const int tid = threadIdx.x, lane = tid & 31;
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
        mbar_init(bar_tma, 1);
        mbar_init(&bar_full[0], NTC);
        mbar_init(&bar_full[1], NTC);
        mbar_init(&bar_empty[0], NTC);
        mbar_init(&bar_empty[1], NTC);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncthreads();

      const u64 one = pack2(prm.one, prm.one);

      if (tid < NTC)
      {
        // =========================== row warps: staging buffer -> F ===========================
        const int wq = tid >> 5;
        auto issue = [&](int u) {  // thread 0
          mbar_expect_tx(bar_tma, static_cast<unsigned>(n_reg_on * S::region_floats * 4));
          for (int g = 0; g < n_reg_on; ++g)
            tma_load_2d(raw + g * S::region_floats, &tmap, x0 - S::LEAD + HW * g, r_begin + R * u, bar_tma);
          if (u + 1 < NB)  // pull the block after this one into L2: its TMA load will wait on the column warps only
            for (int g = 0; g < n_reg_on; ++g)
              tma_prefetch_2d(&tmap, x0 - S::LEAD + HW * g, r_begin + R * (u + 1));
        };
        if (tid == 0)
          issue(0);
        const int r_row = 8 * wq + (lane & 7);  // row of the block
        const int r_run = lane >> 3;            // run of the row (both halves)
        const float* const in0 = raw + r_row * BW + skew + T * r_run;
        const float* const in1 = in0 + S::region_floats;

        for (int u = 0; u < NB; ++u)
        {
          mbar_wait(bar_tma, u & 1);
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
            row_group_sync();
          }
          if (u >= 2)
            mbar_wait(&bar_empty[u & 1], ((u >> 1) - 1) & 1);  // the column warps are done with this F buffer
          float* const r_out = sm + S::off_F + (u & 1) * (R * PF) + r_row * PF + T * r_run;
          // this warp's eight rows: skip the arithmetic when none of them is a row the column pass reads
          // (above / below the image, or past the end of the segment) -- warp-uniform
          const int wr0 = r_begin + R * u + 8 * wq;
          const bool rows_used = wr0 + 8 > 0 && wr0 < min(r_end, h);

          // Inputs are consumed in groups of RG; a group's loads are issued one group ahead and
          // a __syncwarp() closes every group.  The fence is a scheduling device: without it
          // ptxas hoists all the loads of this straight-line code to the top and then, short of
          // registers, re-orders the arithmetic output by output -- serial chains of dependent
          // FFMA2 -- instead of input by input (2c + 1 independent FFMA2 per input).
          if (rows_used)
          {
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
          mbar_arrive(&bar_full[u & 1]);  // release: this thread's part of F is written
          row_group_sync();               // every row warp has consumed the staging buffer
          if (tid == 0 && u + 1 < NB)
          {
            if (edge)
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(u + 1);
          }
        }
      }
      else
      {
        // ================== column warps: F -> G(s), D(s-1), next octave ===================
        // Steps run in groups of RING = 8: the ring slot of a step is its index in the group
        // (every block has a multiple of 8 steps), so all ring addresses are immediates.  Blocks
        // that lie inside the image and inside the segment's store range take a lean path with
        // no row clamping and no store predicate.
        // Threads beyond the strip's last column pair (the last 4 of the 128, and the tail of a
        // ragged last strip) shadow the last valid pair: same inputs, same bits, same addresses --
        // a benign duplicate store instead of a divergent branch around every global access.
        const int wq = (tid - NTC) >> 5;
        const int n_pairs = min(NP, (w - x0 + 1) >> 1);  // >= 1
        const int ct = min(tid - NTC, n_pairs - 1);
        const u64 neg_one = pack2(prm.neg_one, prm.neg_one);
        const int xa = x0 + 2 * ct;  // image column of the low half; the high half is the next column
        constexpr bool col_on = true;
        const bool warp_on = 32 * wq < n_pairs;  // warp-uniform
        u64 A[2 * c];
#pragma unroll
        for (int s = 0; s < 2 * c; ++s)
          A[s] = 0ull;
        const long long pitch_b = static_cast<long long>(prm.pitch) * 4;
        const long long spitch_b = static_cast<long long>(prm.src_pitch) * 4;
        // output pointers of the step about to run (output row y = r - c)
        char* po = reinterpret_cast<char*>(prm.out) + (static_cast<long long>(r_begin - c) * prm.pitch + xa) * 4;
        char* pd = DOG ? reinterpret_cast<char*>(prm.dog) + (static_cast<long long>(r_begin - c) * prm.pitch + xa) * 4
                       : nullptr;
        // G(s-1) prefetch ring: the copy for output row y + PD is issued at the step of row y
        constexpr int RING = S::RING, PD = S::PD;
        static_assert(RING == 8 && R % RING == 0, "static ring slots");
        u64* const ring = reinterpret_cast<u64*>(sm + S::off_ring) + (tid - NTC);
        const char* pgw =
            reinterpret_cast<const char*>(prm.src) + (static_cast<long long>(r_begin - c) * prm.src_pitch + xa) * 4;
        int yw = r_begin - c;  // row the next copy fetches
        const bool next_on = NEXT && (xa >> 1) < prm.nw;
        float* const pnext = prm.nextG + (xa >> 1);
        if (DOG && warp_on)
        {
#pragma unroll
          for (int q = 0; q < PD; ++q)
          {
            if (col_on && yw >= 0 && yw < h)
              cp_async8(ring + q * NTC, pgw);
            cp_async_commit();
            pgw += spitch_b;
            ++yw;
          }
        }

        // one step: F row `fp` in, the finished output row out of the in-flight sums
        auto advance = [&](const float* fp) {
          const u64 v = *reinterpret_cast<const u64*>(fp);
          u64 p[c + 1];
#pragma unroll
          for (int m = 0; m <= c; ++m)
            p[m] = mul2(v, pack2(prm.taps[m], prm.taps[m]));
          const u64 E = add2(A[0], p[0], one);
#pragma unroll
          for (int s = 1; s < 2 * c; ++s)
            A[s - 1] = add2(A[s], p[s <= c ? s : 2 * c - s], one);
          A[2 * c - 1] = add2(0ull, p[0], one);
          return E;
        };

        for (int u = 0; u < NB; ++u)
        {
          const int rb = r_begin + R * u;  // first row of this block
          mbar_wait(&bar_full[u & 1], (u >> 1) & 1);
          if (warp_on)
          {
            const float* fcol = sm + S::off_F + (u & 1) * (R * PF) + 2 * ct;
            // steps whose output row y = rb + k - c lies in [y0, y1): k in [ks_lo, ks_hi)
            const int ks_lo = y0 + c - rb, ks_hi = y1 + c - rb;
            // (the last block is never lean: it may have to run on past its 32 rows, replicating the border row)
            const bool lean = u != NB - 1 && rb >= 0 && rb + R <= h && ks_lo <= 0 && ks_hi >= R && yw + R <= h && yw >= 0;
            if (lean)
            {
              // 8-byte stores: when w is odd the high half of the last pair falls into the row padding
#pragma unroll 1
              for (int k0 = 0; k0 < R; k0 += RING)
              {
#pragma unroll
                for (int q = 0; q < RING; ++q)
                {
                  if (DOG)
                  {
                    if (col_on)
                      cp_async8(ring + ((q + PD) % RING) * NTC, pgw);
                    cp_async_commit();
                    pgw += spitch_b;
                  }
                  const u64 E = advance(fcol + (k0 + q) * PF);
                  if (DOG)
                    cp_async_wait<PD>();  // the copy issued PD steps ago (this row) has landed
                  if (col_on)
                  {
                    *reinterpret_cast<u64*>(po) = E;
                    if (DOG)
                      *reinterpret_cast<u64*>(pd) = add2(ring[q * NTC], E, neg_one);  // RN(E - G(s-1))
                    if (NEXT && next_on && (q & 1) == ((rb - c) & 1))  // even output row: downscale(G(s), 2)
                    {
                      const int yy = (rb + k0 + q - c) >> 1;
                      if (yy < prm.nh)
                        pnext[static_cast<size_t>(yy) * prm.npitch] = lo2(E);
                    }
                  }
                  po += pitch_b;
                  if (DOG)
                    pd += pitch_b;
                }
              }
              yw += R;
            }
            else
            {
              const int n_steps = u == NB - 1 ? ((r_end - rb + RING - 1) / RING) * RING : R;
              // F row of step k: the block row of image row clamp(rb + k) (border rows replicate)
              const int k_lo = max(-rb, 0), k_hi = min(h - 1 - rb, R - 1);
#pragma unroll 1
              for (int k0 = 0; k0 < n_steps; k0 += RING)
              {
#pragma unroll
                for (int q = 0; q < RING; ++q)
                {
                  const int k = k0 + q;
                  if (DOG)
                  {
                    if (col_on && yw >= 0 && yw < h)
                      cp_async8(ring + ((q + PD) % RING) * NTC, pgw);
                    cp_async_commit();
                    pgw += spitch_b;
                    ++yw;
                  }
                  const u64 E = advance(fcol + min(max(k, k_lo), k_hi) * PF);
                  if (DOG)
                    cp_async_wait<PD>();
                  if (k >= ks_lo && k < ks_hi && col_on)
                  {
                    *reinterpret_cast<u64*>(po) = E;
                    if (DOG)
                      *reinterpret_cast<u64*>(pd) = add2(ring[q * NTC], E, neg_one);
                    const int y = rb + k - c;
                    if (NEXT && next_on && (y & 1) == 0 && (y >> 1) < prm.nh)
                      pnext[static_cast<size_t>(y >> 1) * prm.npitch] = lo2(E);
                  }
                  po += pitch_b;
                  if (DOG)
                    pd += pitch_b;
                }
              }
            }
          }
          mbar_arrive(&bar_empty[u & 1]);  // this thread no longer reads the F buffer
        }
      }
    }

    thread_local int t_short_blocks = 4;

    template <int K>
    bool launch(const float* src, int src_pitch, float* dst, float* dog, float* nextG, int w, int h, int pitch, int nw,
                int nh, int npitch, const Taps& taps, cudaStream_t st)
    {
      using S = MC<K>;
      constexpr int per_sm = S::PER_SM;
      static_assert(per_sm * (S::smem_bytes + 1024) <= 233472, "resident CTAs must fit one SM");
      static bool configured = false;
      if (!configured)
      {
        if (cudaFuncSetAttribute(march_kernel<K, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 S::smem_bytes) != cudaSuccess ||
            cudaFuncSetAttribute(march_kernel<K, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 S::smem_bytes) != cudaSuccess ||
            cudaFuncSetAttribute(march_kernel<K, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 S::smem_bytes) != cudaSuccess)
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

      // Segments.  Every segment pays 2c warm-up rows and works in 32-row blocks.  A large layer
      // is cut into as many CTAs as the machine holds at once (148 SMs x 2 resident CTAs) and no
      // more -- a second wave would only add warm-up rows.  A small layer cannot fill the machine;
      // it gets segments of a whole number of blocks (warm-up included): 4 blocks when a frame runs
      // alone (measured best for its latency), 6 when several frames are in flight (a quarter less
      // machine time for octave 1 of a 4K frame than the two-block segments of earlier versions).
      const int n_strips = (w + TX - 1) / TX;
      const int slots = 148 * per_sm;
      // Short-segment rule, in 32-row blocks through the pipeline (set_march_schedule): taller segments
      // waste fewer warm-up rows (machine time), shorter ones finish a lone launch sooner (latency).
      static const int env_blocks = [] {
        const char* e = getenv("SARA_B200_MARCH_BLOCKS");
        return e && atoi(e) > 0 ? atoi(e) : 0;
      }();
      const int short_blocks = env_blocks > 0 ? env_blocks : t_short_blocks;
      const int hy_short = short_blocks * R - 2 * S::c;
      int best_segs = std::max(1, std::min((h + hy_short - 1) / hy_short, slots / n_strips));
      // a layer that would fill most of the machine anyway gets exactly one CTA per slot
      if (best_segs * n_strips * 100 >= slots * 60)
        best_segs = std::max(1, slots / n_strips);
      static const int force = [] {
        const char* e = getenv("SARA_B200_MARCH_SEGS");
        return e ? atoi(e) : 0;
      }();
      int n_segs = force > 0 ? force : best_segs;
      int hy = (h + n_segs - 1) / n_segs;
      n_segs = (h + hy - 1) / hy;
      prm.hy = hy;
      if (nextG != nullptr && dog == nullptr)
        return false;  // not a combination the pyramid produces
      const dim3 grid(n_strips, n_segs);
      if (nextG != nullptr)
        march_kernel<K, true, true><<<grid, NT, S::smem_bytes, st>>>(tmap, prm);
      else if (dog != nullptr)
        march_kernel<K, true, false><<<grid, NT, S::smem_bytes, st>>>(tmap, prm);
      else
        march_kernel<K, false, false><<<grid, NT, S::smem_bytes, st>>>(tmap, prm);
      return true;
    }

  }  // namespace march

  void set_march_schedule(bool throughput) { march::t_short_blocks = throughput ? 6 : 4; }

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
    case 11: return march::launch<11>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 13: return march::launch<13>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 17: return march::launch<17>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 21: return march::launch<21>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    case 25: return march::launch<25>(src, src_pitch, dst, dog, nextG, w, h, pitch, nw, nh, npitch, taps, st);
    default: return false;
    }
  }

}  // namespace sb
