// Descriptor matching on the device: the nearest-neighbour search
// underneath AnnMatcher::compute_matches (FeatureMatching/AnnMatcher.cpp:219-282).
//
// The reference asks FLANN (its vendored third-party/flann) for the 3 nearest neighbours of
// every descriptor in the other image, squared L2 distance, through a randomised KD-tree forest
// -- an approximate search.  This file computes the EXACT neighbours (what FLANN's own
// LinearIndex returns), with distances that carry the bits of flann::L2<float>
// (algorithms/dist.h:151-178: groups of four squared differences), in two steps:
//
//  1. CANDIDATES on the tensor cores (dim = 128).  ||a - b||^2 = |a|^2 + |b|^2 - 2 a.b, and a.b
//     over all pairs is a dense N1 x N2 x 128 contraction -- the one GEMM of the SIFT path.  Each
//     fp32 descriptor is split into two bf16 numbers (x = hi + lo, relative residual 2^-17), and
//     a.b ~ hi.hi + lo.hi + hi.lo is accumulated in fp32 by tcgen05.mma (kind::f16, M 128 x N 128
//     x K 16, operands staged by TMA into 128-byte-swizzled shared memory, accumulators in TMEM,
//     double buffered).  Four epilogue warps read the accumulators back with tcgen05.ld -- a
//     thread owns one query row -- and keep that row's 8 best keys |b|^2 - 2 a.b in registers.
//     Nothing of the N1 x N2 matrix ever reaches memory.
//  2. EXACT re-ranking.  A warp per query recomputes the FLANN distance of its <= 32 candidates
//     in fp32 (separate multiply and add, the library is compiled with -fmad=false), orders them
//     by (distance, index) -- FLANN's result set keeps the first of two equal distances -- and
//     certifies the result: if the worst key a split kept, minus a bound on the bf16 / fp32
//     error, is not above the k-th exact distance, a dropped point could belong to the answer,
//     and the query is re-done by the exact scalar kernel below.
//
// The scalar kernel (any dimension <= 256, thread per query, data tiles broadcast from shared
// memory) is also the whole search for dim != 128, the radius search of the `ratio > 1` branch
// (AnnMatcher.cpp:141-146), and the fallback for uncertified queries.  There is no CPU path.
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"
#include "fp32x2_tma.cuh"
#include "match.cuh"

namespace sb {
  namespace match {

    using namespace fused;

    constexpr int KC = 8;        // candidates kept per (query, split)
    constexpr int MAX_SPLITS = 4;  // KC * MAX_SPLITS <= 32: one candidate per lane in the re-ranking
    constexpr int QB = 128;      // queries per block (both kernels)

    // ------------------------------------------------------------------------------------------
    // flann::L2<float>::operator() (dist.h:151-178) for one pair of rows, float4 loads when aligned
    __device__ __forceinline__ float l2_flann_rows(const float* __restrict__ a, const float* __restrict__ b, int dim)
    {
      float result = 0.f;
      int i = 0;
      if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0)
      {
        for (; i + 3 < dim; i += 4)
        {
          const float4 x = __ldg(reinterpret_cast<const float4*>(a + i));
          const float4 y = __ldg(reinterpret_cast<const float4*>(b + i));
          const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
          result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
      }
      else
      {
        for (; i + 3 < dim; i += 4)
        {
          const float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
          result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
      }
      for (; i < dim; ++i)
      {
        const float d0 = a[i] - b[i];
        result += d0 * d0;
      }
      return result;
    }

    // Sorted insertion into a thread-private list (ascending keys, static indexing only).  A new key
    // equal to a stored one goes BEHIND it: points are visited in increasing index order, so equal
    // distances keep the lower index first (KNNSimpleResultSet::addPoint, util/result_set.h:151-171).
    template <int K>
    __device__ __forceinline__ void insert_sorted(float (&kd)[K], int (&ki)[K], float key, int id)
    {
#pragma unroll
      for (int i = 0; i < K; ++i)
      {
        if (key < kd[i])
        {
          const float tk = kd[i];
          const int ti = ki[i];
          kd[i] = key;
          ki[i] = id;
          key = tk;
          id = ti;
        }
      }
    }

    // ------------------------------------------------------------------------------------------
    // Exact scalar search.  Block = 128 queries (thread per query; the query vectors sit transposed
    // in shared memory, conflict free), grid.y = splits of the data range; data rows are staged in
    // tiles of TJ rows and read as broadcasts.
    //   MODE 0: top-KC of the split -> cand_key / cand_idx[(q * splits + split) * KC + i]
    //   MODE 1: count the points with dist < radius[q]      -> atomicAdd(count[q])
    //   MODE 2: write them at fill_off[q] + (running index) -> (out_idx, out_dist), unsorted
    constexpr int TJ = 32;

    template <int MODE>
    __global__ void __launch_bounds__(QB)
        knn_exact_kernel(const float* __restrict__ queries, const int* __restrict__ qlist, int nq,
                         const float* __restrict__ data, int nd, int dim, int splits, float* __restrict__ cand_key,
                         int* __restrict__ cand_idx, const float* __restrict__ radius, int* __restrict__ count,
                         const int* __restrict__ fill_off, int* __restrict__ fill_cursor, int* __restrict__ out_idx,
                         float* __restrict__ out_dist)
    {
      extern __shared__ float smem[];
      float* qs = smem;                  // [dim][QB]
      float* tile = smem + dim * QB;     // [TJ][dim]
      const int t = threadIdx.x;
      const int qslot = blockIdx.x * QB + t;
      const bool live = qslot < nq;
      const int q = live ? (qlist ? qlist[qslot] : qslot) : 0;

      // transposed query block: coalesced reads of 128 rows, one row after the other
      for (int r = 0; r < QB; ++r)
      {
        const int qs_slot = blockIdx.x * QB + r;
        if (qs_slot >= nq)
          break;
        const int qr = qlist ? qlist[qs_slot] : qs_slot;
        for (int d = t; d < dim; d += QB)
          qs[d * QB + r] = queries[static_cast<size_t>(qr) * dim + d];
      }

      const int per = (nd + splits - 1) / splits;
      const int j0 = blockIdx.y * per, j1 = min(nd, j0 + per);

      float kd[KC];
      int ki[KC];
#pragma unroll
      for (int i = 0; i < KC; ++i)
      {
        kd[i] = FLT_MAX;
        ki[i] = -1;
      }
      const float rad = (MODE != 0 && live) ? radius[q] : 0.f;
      int n_in = 0;

      for (int jb = j0; jb < j1; jb += TJ)
      {
        const int rows = min(TJ, j1 - jb);
        __syncthreads();
        for (int e = t; e < rows * dim; e += QB)
          tile[e] = data[static_cast<size_t>(jb) * dim + e];
        __syncthreads();
        if (!live)
          continue;
        for (int r = 0; r < rows; ++r)
        {
          const float* b = tile + r * dim;
          float result = 0.f;
          int i = 0;
          for (; i + 3 < dim; i += 4)
          {
            const float d0 = b[i] - qs[i * QB + t], d1 = b[i + 1] - qs[(i + 1) * QB + t];
            const float d2 = b[i + 2] - qs[(i + 2) * QB + t], d3 = b[i + 3] - qs[(i + 3) * QB + t];
            result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          }
          for (; i < dim; ++i)
          {
            const float d0 = b[i] - qs[i * QB + t];
            result += d0 * d0;
          }
          if (MODE == 0)
          {
            if (result < kd[KC - 1])
              insert_sorted<KC>(kd, ki, result, jb + r);
          }
          else if (result < rad)
          {
            if (MODE == 2)
            {
              const int at = fill_off[q] + atomicAdd(fill_cursor + q, 1);
              out_idx[at] = jb + r;
              out_dist[at] = result;
            }
            ++n_in;
          }
        }
      }
      if (!live)
        return;
      if (MODE == 0)
      {
        const size_t base = (static_cast<size_t>(q) * splits + blockIdx.y) * KC;
#pragma unroll
        for (int i = 0; i < KC; ++i)
        {
          cand_key[base + i] = kd[i];
          cand_idx[base + i] = ki[i];
        }
      }
      else if (MODE == 1 && n_in)
        atomicAdd(count + q, n_in);
    }

    // ------------------------------------------------------------------------------------------
    // fp32 -> (hi, lo) bf16 split, row layout [hi(128) | lo(128)], rows padded to a multiple of
    // 128 with zeros; norm[r] = sum x^2 (padded rows: +inf, so that they are never candidates).
    __global__ void __launch_bounds__(256)
        split_bf16_kernel(const float* __restrict__ src, int n, int n_pad, __nv_bfloat16* __restrict__ dst,
                          float* __restrict__ norm)
    {
      const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
      if (warp >= n_pad)
        return;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (warp < n)
        v = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(warp) * 128) + lane);
      const float x[4] = {v.x, v.y, v.z, v.w};
      __nv_bfloat16 hi[4], lo[4];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        hi[i] = __float2bfloat16_rn(x[i]);
        lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(hi[i]));
        s += x[i] * x[i];
      }
      __nv_bfloat16* row = dst + static_cast<size_t>(warp) * 256;
      *reinterpret_cast<uint2*>(row + 4 * lane) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(row + 128 + 4 * lane) = *reinterpret_cast<const uint2*>(lo);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, d);
      if (lane == 0)
        norm[warp] = warp < n ? s : __int_as_float(0x7f800000);
    }

    // ------------------------------------------------------------------------------------------
    // tcgen05 candidate kernel.
    constexpr int BLK_BYTES = 128 * 128;          // one 64-column block of a tile: 128 rows x 128 bytes
    constexpr int TILE_BYTES = 4 * BLK_BYTES;     // 128 rows x 256 bf16 (hi | lo)
    constexpr int B_STAGES = 2;
    constexpr int MMA_SMEM = (1 + B_STAGES) * TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    constexpr int TMEM_COLS = 256;                // two accumulator stages of 128 columns

    __device__ __forceinline__ void mbar_arrive_plain(void* bar)
    {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    __device__ __forceinline__ void tc_fence_before()
    {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __device__ __forceinline__ void tc_fence_after()
    {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    __device__ __forceinline__ void tc_commit(void* bar)  // arrives on `bar` when all prior MMAs of this thread are done
    {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                   : "memory");
    }
    // K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1.
    __device__ __forceinline__ uint64_t umma_desc(unsigned smem_addr)
    {
      return static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4) | (static_cast<uint64_t>(1) << 16) |
             (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
             (static_cast<uint64_t>(2) << 61);
    }
    // D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M 128, N 128, K 16
    __device__ __forceinline__ void umma_bf16(unsigned tmem_d, uint64_t a_desc, uint64_t b_desc, unsigned idesc,
                                              unsigned accumulate)
    {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
          "}\n" ::"r"(tmem_d),
          "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
          : "memory");
    }
    __device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32])
    {
      unsigned r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i)
        v[i] = __uint_as_float(r[i]);
    }

    // grid = (query tiles, splits), 256 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
    // allocator, warps 4-7 epilogue (warp w reads TMEM lanes 32 (w % 4) ..).
    __global__ void __launch_bounds__(256, 1)
        knn_mma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                       const float* __restrict__ norm_d, int n_tiles_d, int splits, float* __restrict__ cand_key,
                       int* __restrict__ cand_idx)
    {
      extern __shared__ unsigned char smem_raw[];
      unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
      unsigned char* sA = smem;
      unsigned char* sB = smem + TILE_BYTES;
      uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + B_STAGES) * TILE_BYTES);
      uint64_t* a_full = bars;            // 1
      uint64_t* b_full = bars + 1;        // B_STAGES
      uint64_t* b_empty = bars + 3;       // B_STAGES
      uint64_t* acc_full = bars + 5;      // 2
      uint64_t* acc_empty = bars + 7;     // 2
      unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 9);

      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int per = (n_tiles_d + splits - 1) / splits;
      const int t0 = blockIdx.y * per, t1 = min(n_tiles_d, t0 + per);
      const int n_my = max(0, t1 - t0);

      if (threadIdx.x == 0)
      {
        mbar_init(a_full, 1);
        for (int s = 0; s < B_STAGES; ++s)
        {
          mbar_init(b_full + s, 1);
          mbar_init(b_empty + s, 1);
        }
        for (int s = 0; s < 2; ++s)
        {
          mbar_init(acc_full + s, 1);
          mbar_init(acc_empty + s, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      if (warp == 2)
      {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      const unsigned tmem_base = *tmem_slot;

      if (warp == 0 && lane == 0)
      {
        // ---- TMA producer ----
        mbar_expect_tx(a_full, TILE_BYTES);
        for (int b = 0; b < 4; ++b)
          tma_load_2d(sA + b * BLK_BYTES, &map_q, 64 * b, blockIdx.x * 128, a_full);
        for (int i = 0; i < n_my; ++i)
        {
          const int s = i % B_STAGES;
          const unsigned ph = (i / B_STAGES) & 1;
          mbar_wait(b_empty + s, ph ^ 1);
          mbar_expect_tx(b_full + s, TILE_BYTES);
          for (int b = 0; b < 4; ++b)
            tma_load_2d(sB + s * TILE_BYTES + b * BLK_BYTES, &map_d, 64 * b, (t0 + i) * 128, b_full + s);
        }
      }
      else if (warp == 1 && lane == 0)
      {
        // ---- MMA issuer ----
        // instruction descriptor: D fp32 (bit 4), A and B bf16 (bits 7, 10), both K-major, N >> 3 at
        // bit 17, M >> 4 at bit 24
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const unsigned a_addr = smem_u32(sA);
        mbar_wait(a_full, 0);
        for (int i = 0; i < n_my; ++i)
        {
          const int s = i % B_STAGES;
          const unsigned ph = (i / B_STAGES) & 1;
          const int as = i & 1;
          const unsigned aph = (i >> 1) & 1;
          mbar_wait(acc_empty + as, aph ^ 1);
          mbar_wait(b_full + s, ph);
          tc_fence_after();
          const unsigned b_addr = smem_u32(sB + s * TILE_BYTES);
          const unsigned d_tmem = tmem_base + as * 128;
          // hi.hi (blocks 0,1 x 0,1), lo.hi (2,3 x 0,1), hi.lo (0,1 x 2,3)
          const int ablk[6] = {0, 1, 2, 3, 0, 1};
          const int bblk[6] = {0, 1, 0, 1, 2, 3};
          unsigned acc = 0;
#pragma unroll
          for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
            {
              umma_bf16(d_tmem, umma_desc(a_addr + ablk[p] * BLK_BYTES + kk * 32),
                        umma_desc(b_addr + bblk[p] * BLK_BYTES + kk * 32), idesc, acc);
              acc = 1;
            }
          tc_commit(b_empty + s);   // the stage may be refilled once these MMAs have read it
          tc_commit(acc_full + as); // and the accumulator is complete
        }
      }
      else if (warp >= 4)
      {
        // ---- epilogue: thread = one query row ----
        const int row = threadIdx.x - 128;  // TMEM lane
        const int ew = warp - 4;
        float kd[KC];
        int ki[KC];
#pragma unroll
        for (int i = 0; i < KC; ++i)
        {
          kd[i] = __int_as_float(0x7f800000);
          ki[i] = -1;
        }
        for (int i = 0; i < n_my; ++i)
        {
          const int as = i & 1;
          const unsigned aph = (i >> 1) & 1;
          mbar_wait(acc_full + as, aph);
          tc_fence_after();
          const int jbase = (t0 + i) * 128;
#pragma unroll 1
          for (int c = 0; c < 4; ++c)
          {
            float v[32];
            __syncwarp();
            tmem_ld32(tmem_base + (static_cast<unsigned>(32 * ew) << 16) + as * 128 + 32 * c, v);
            const float4* nb4 = reinterpret_cast<const float4*>(norm_d + jbase + 32 * c);
#pragma unroll
            for (int g = 0; g < 8; ++g)
            {
              const float4 nb = __ldg(nb4 + g);
              const float k0 = nb.x - 2.f * v[4 * g], k1 = nb.y - 2.f * v[4 * g + 1];
              const float k2 = nb.z - 2.f * v[4 * g + 2], k3 = nb.w - 2.f * v[4 * g + 3];
              if (k0 < kd[KC - 1])
                insert_sorted<KC>(kd, ki, k0, jbase + 32 * c + 4 * g);
              if (k1 < kd[KC - 1])
                insert_sorted<KC>(kd, ki, k1, jbase + 32 * c + 4 * g + 1);
              if (k2 < kd[KC - 1])
                insert_sorted<KC>(kd, ki, k2, jbase + 32 * c + 4 * g + 2);
              if (k3 < kd[KC - 1])
                insert_sorted<KC>(kd, ki, k3, jbase + 32 * c + 4 * g + 3);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0)
            mbar_arrive_plain(acc_empty + as);
        }
        const size_t base = ((static_cast<size_t>(blockIdx.x) * 128 + row) * splits + blockIdx.y) * KC;
#pragma unroll
        for (int i = 0; i < KC; ++i)
        {
          cand_key[base + i] = kd[i];
          cand_idx[base + i] = ki[i];
        }
      }

      tc_fence_before();
      __syncthreads();
      if (warp == 2)
      {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
      }
    }

    // ------------------------------------------------------------------------------------------
    // Re-ranking: a warp per query, one candidate per lane (splits * KC <= 32).
    //   approx != 0: keys are |b|^2 - 2 a.b from the tensor cores: recompute the exact distance of
    //                every candidate and certify; uncertified queries are appended to `redo`.
    //   approx == 0: keys are exact distances already.
    // Output: k entries per query, ascending (distance, index); unused entries (-1, FLT_MAX).
    __global__ void __launch_bounds__(256)
        rerank_kernel(const float* __restrict__ queries, const int* __restrict__ qlist, int nq,
                      const float* __restrict__ data, int nd, int dim, int splits, int k, int approx,
                      const float* __restrict__ norm_q, const float* __restrict__ cand_key,
                      const int* __restrict__ cand_idx, int* __restrict__ out_idx, float* __restrict__ out_dist,
                      int* __restrict__ redo, int* __restrict__ n_redo)
    {
      const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
      if (w >= nq)
        return;
      const int q = qlist ? qlist[w] : w;
      const int C = splits * KC;
      float key = FLT_MAX;
      int idx = -1;
      if (lane < C)
      {
        key = cand_key[static_cast<size_t>(q) * C + lane];
        idx = cand_idx[static_cast<size_t>(q) * C + lane];
      }
      float d = FLT_MAX;
      if (idx >= 0)
        d = approx ? l2_flann_rows(data + static_cast<size_t>(idx) * dim, queries + static_cast<size_t>(q) * dim, dim) : key;
      // rank among the candidates by (d, idx)
      int rank = 0;
      for (int l = 0; l < 32; ++l)
      {
        const float dl = __shfl_sync(0xffffffffu, d, l);
        const int il = __shfl_sync(0xffffffffu, idx, l);
        if (il >= 0 && idx >= 0 && (dl < d || (dl == d && il < idx)))
          ++rank;
      }
      const unsigned valid = __ballot_sync(0xffffffffu, idx >= 0);
      const int n_valid = __popc(valid);
      if (idx >= 0 && rank < k)
      {
        out_idx[static_cast<size_t>(q) * k + rank] = idx;
        out_dist[static_cast<size_t>(q) * k + rank] = d;
      }
      if (lane >= n_valid && lane < k)
      {
        out_idx[static_cast<size_t>(q) * k + lane] = -1;
        out_dist[static_cast<size_t>(q) * k + lane] = FLT_MAX;
      }
      if (approx)
      {
        // the k-th exact distance among the candidates (or the last valid one)
        const int kth = min(k, n_valid) - 1;
        const unsigned who = __ballot_sync(0xffffffffu, idx >= 0 && rank == kth);
        const float dk = who ? __shfl_sync(0xffffffffu, d, __ffs(who) - 1) : FLT_MAX;
        // the smallest key any split may have dropped: its KC-th kept key (inf when the split kept everything)
        float thr = __int_as_float(0x7f800000);
        if (lane < C && (lane % KC) == KC - 1 && idx >= 0)
          thr = key;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          thr = fminf(thr, __shfl_xor_sync(0xffffffffu, thr, o));
        const float na = norm_q[q];
        const float thr_dist = na + thr;
        const float eps = 1e-3f * (fabsf(na) + fabsf(thr_dist)) + 1.f;
        const bool kept_all = !(thr < __int_as_float(0x7f800000));  // no split dropped anything
        const bool certain = n_valid >= min(k, nd) && (kept_all || thr_dist - eps > dk);
        if (!certain && lane == 0)
          redo[atomicAdd(n_redo, 1)] = q;
      }
    }

    // ------------------------------------------------------------------------------------------
    // host side
    namespace {
      int grow(void** p, size_t* have, size_t need)
      {
        if (need <= *have)
          return 0;
        if (*p)
          cudaFree(*p);
        *p = nullptr;
        *have = 0;
        if (cudaMalloc(p, need) != cudaSuccess)
          return -1;
        *have = need;
        return 0;
      }

      template <class T>
      T* carve(unsigned char*& cur, size_t n)
      {
        T* p = reinterpret_cast<T*>(cur);
        cur += (n * sizeof(T) + 255) / 256 * 256;
        return p;
      }

      bool encode_map(CUtensorMap* m, const __nv_bfloat16* base, int rows)
      {
        EncodeTiledFn enc = encode_fn();
        if (!enc)
          return false;
        const cuuint64_t dims[2] = {256, static_cast<cuuint64_t>(rows)};
        const cuuint64_t strides[1] = {512};
        const cuuint32_t box[2] = {64, 128};
        const cuuint32_t estr[2] = {1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }

      size_t exact_smem(int dim)
      {
        return static_cast<size_t>(dim) * (QB + TJ) * sizeof(float);
      }
    }  // namespace

    void Workspace::release()
    {
      if (buf)
        cudaFree(buf);
      buf = nullptr;
      bytes = 0;
    }

    bool mma_path_available()
    {
      return encode_fn() != nullptr;
    }

    // k nearest neighbours (k <= KC) of every query row among the data rows; everything on `st`.
    // d_idx / d_dist: nq x k, device.  Returns 0, or a negative sara_b200 status with `err` filled.
    int knn(Workspace& ws, const float* d_q, int nq, const float* d_data, int nd, int dim, int k, int mode,
            int* d_idx, float* d_dist, KnnStats* stats, cudaStream_t st, char* err, size_t errlen)
    {
      if (stats)
        *stats = KnnStats{};
      if (k < 1 || k > KC)
      {
        snprintf(err, errlen, "knn: k = %d outside [1, %d]", k, KC);
        return SARA_B200_ERR_BAD_ARG;
      }
      if (dim < 1 || dim > 256)
      {
        snprintf(err, errlen, "knn: descriptor dimension %d outside [1, 256]", dim);
        return SARA_B200_ERR_BAD_ARG;
      }
      if (nq == 0)
        return 0;
      const bool aligned = ((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_data)) & 15) == 0;
      bool use_mma = mode != SARA_B200_KNN_SCALAR && dim == 128 && nd >= 1 && aligned && mma_path_available();
      if (mode == SARA_B200_KNN_TENSOR && !use_mma)
      {
        snprintf(err, errlen, "knn: the tensor-core path needs dim == 128 and 16-byte aligned descriptors");
        return SARA_B200_ERR_BAD_ARG;
      }
      if (mode == SARA_B200_KNN_AUTO && static_cast<double>(nq) * nd < 256.0 * 256.0)
        use_mma = false;  // tiny problems: the scalar kernel alone is one launch

      const int nq_pad = (nq + 127) / 128 * 128, nd_pad = (nd + 127) / 128 * 128;
      const int q_tiles = nq_pad / 128, d_tiles = nd_pad / 128;
      int splits = 1;
      if (use_mma)
        while (splits < MAX_SPLITS && q_tiles * splits * 2 <= 148 && d_tiles >= 2 * splits * 2)
          splits *= 2;
      else
        while (splits < MAX_SPLITS && ((nq + QB - 1) / QB) * splits * 2 <= 2 * 148 && nd >= 2 * splits * 4 * TJ)
          splits *= 2;

      // workspace
      size_t need = 0;
      auto add = [&](size_t b) { need += (b + 255) / 256 * 256; };
      add(static_cast<size_t>(nq_pad) * MAX_SPLITS * KC * 4);  // cand_key
      add(static_cast<size_t>(nq_pad) * MAX_SPLITS * KC * 4);  // cand_idx
      add(static_cast<size_t>(nq) * 4);                          // redo list
      add(256);                                                  // n_redo
      if (use_mma)
      {
        add(static_cast<size_t>(nq_pad) * 512);
        add(static_cast<size_t>(nd_pad) * 512);
        add(static_cast<size_t>(nq_pad) * 4);
        add(static_cast<size_t>(nd_pad) * 4);
      }
      if (need > ws.bytes)
      {
        cudaStreamSynchronize(st);
        if (grow(reinterpret_cast<void**>(&ws.buf), &ws.bytes, need) != 0)
        {
          snprintf(err, errlen, "knn: cudaMalloc of %zu workspace bytes failed", need);
          return SARA_B200_ERR_OOM;
        }
      }
      unsigned char* cur = ws.buf;
      float* cand_key = carve<float>(cur, static_cast<size_t>(nq_pad) * MAX_SPLITS * KC);
      int* cand_idx = carve<int>(cur, static_cast<size_t>(nq_pad) * MAX_SPLITS * KC);
      int* redo = carve<int>(cur, nq);
      int* n_redo = carve<int>(cur, 64);

      const size_t ex_smem = exact_smem(dim);
      cudaFuncSetAttribute(knn_exact_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (use_mma)
        cudaFuncSetAttribute(knn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MMA_SMEM);

      int launches = 0;
      if (use_mma)
      {
        __nv_bfloat16* qb = carve<__nv_bfloat16>(cur, static_cast<size_t>(nq_pad) * 256);
        __nv_bfloat16* db = carve<__nv_bfloat16>(cur, static_cast<size_t>(nd_pad) * 256);
        float* norm_q = carve<float>(cur, nq_pad);
        float* norm_d = carve<float>(cur, nd_pad);
        CUtensorMap map_q, map_d;
        if (!encode_map(&map_q, qb, nq_pad) || !encode_map(&map_d, db, nd_pad))
        {
          snprintf(err, errlen, "knn: cuTensorMapEncodeTiled failed");
          return SARA_B200_ERR_CUDA;
        }
        cudaMemsetAsync(n_redo, 0, 4, st);
        split_bf16_kernel<<<(nq_pad * 32 + 255) / 256, 256, 0, st>>>(d_q, nq, nq_pad, qb, norm_q);
        split_bf16_kernel<<<(nd_pad * 32 + 255) / 256, 256, 0, st>>>(d_data, nd, nd_pad, db, norm_d);
        knn_mma_kernel<<<dim3(q_tiles, splits), 256, MMA_SMEM, st>>>(map_q, map_d, norm_d, d_tiles, splits, cand_key,
                                                                     cand_idx);
        rerank_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(d_q, nullptr, nq, d_data, nd, dim, splits, k, 1, norm_q,
                                                             cand_key, cand_idx, d_idx, d_dist, redo, n_redo);
        launches += 4;
        int h_redo = 0;
        if (cudaMemcpyAsync(&h_redo, n_redo, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
        {
          snprintf(err, errlen, "knn: tensor-core candidate pass failed: %s", cudaGetErrorString(cudaGetLastError()));
          return SARA_B200_ERR_CUDA;
        }
        if (h_redo > 0)
        {
          // uncertified queries: exact scalar search over the whole data set
          int s2 = 1;
          while (s2 < MAX_SPLITS && ((h_redo + QB - 1) / QB) * s2 * 2 <= 2 * 148 && nd >= 2 * s2 * 4 * TJ)
            s2 *= 2;
          knn_exact_kernel<0><<<dim3((h_redo + QB - 1) / QB, s2), QB, ex_smem, st>>>(
              d_q, redo, h_redo, d_data, nd, dim, s2, cand_key, cand_idx, nullptr, nullptr, nullptr, nullptr, nullptr,
              nullptr);
          rerank_kernel<<<(h_redo * 32 + 255) / 256, 256, 0, st>>>(d_q, redo, h_redo, d_data, nd, dim, s2, k, 0, nullptr,
                                                                 cand_key, cand_idx, d_idx, d_dist, nullptr, nullptr);
          launches += 2;
        }
        if (stats)
        {
          stats->used_tensor_cores = 1;
          stats->n_redone = h_redo;
        }
      }
      else
      {
        knn_exact_kernel<0><<<dim3((nq + QB - 1) / QB, splits), QB, ex_smem, st>>>(
            d_q, nullptr, nq, d_data, nd, dim, splits, cand_key, cand_idx, nullptr, nullptr, nullptr, nullptr, nullptr,
            nullptr);
        rerank_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(d_q, nullptr, nq, d_data, nd, dim, splits, k, 0, nullptr,
                                                             cand_key, cand_idx, d_idx, d_dist, nullptr, nullptr);
        launches += 2;
      }
      if (stats)
      {
        stats->launches = launches;
        stats->splits = splits;
      }
      const cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess)
      {
        snprintf(err, errlen, "knn: kernel launch failed: %s", cudaGetErrorString(e));
        return SARA_B200_ERR_CUDA;
      }
      return 0;
    }

    // RadiusResultSet (util/result_set.h:477-510): per query, the points with dist < radius[q].
    // Pass 1 (d_out_idx == nullptr): counts into d_count (zeroed here).  Pass 2: fills
    // (d_out_idx, d_out_dist) at d_off[q] + running position; the caller sorts each segment by
    // (dist, index) as the result set's copy() does.
    int radius_pass(const float* d_q, int nq, const float* d_data, int nd, int dim, const float* d_radius, int* d_count,
                    const int* d_off, int* d_out_idx, float* d_out_dist, cudaStream_t st, char* err, size_t errlen)
    {
      if (nq == 0 || nd == 0)
        return 0;
      if (dim < 1 || dim > 256)
      {
        snprintf(err, errlen, "radius search: descriptor dimension %d outside [1, 256]", dim);
        return SARA_B200_ERR_BAD_ARG;
      }
      cudaFuncSetAttribute(knn_exact_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(knn_exact_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      int splits = 1;
      while (splits < 16 && ((nq + QB - 1) / QB) * splits * 2 <= 2 * 148 && nd >= 2 * splits * 4 * TJ)
        splits *= 2;
      const dim3 grid((nq + QB - 1) / QB, splits);
      cudaMemsetAsync(d_count, 0, sizeof(int) * nq, st);
      if (!d_out_idx)
        knn_exact_kernel<1><<<grid, QB, exact_smem(dim), st>>>(d_q, nullptr, nq, d_data, nd, dim, splits, nullptr, nullptr,
                                                               d_radius, d_count, nullptr, nullptr, nullptr, nullptr);
      else
        knn_exact_kernel<2><<<grid, QB, exact_smem(dim), st>>>(d_q, nullptr, nq, d_data, nd, dim, splits, nullptr, nullptr,
                                                               d_radius, nullptr, d_off, d_count, d_out_idx, d_out_dist);
      const cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess)
      {
        snprintf(err, errlen, "radius search: kernel launch failed: %s", cudaGetErrorString(e));
        return SARA_B200_ERR_CUDA;
      }
      return 0;
    }

  }  // namespace match
}  // namespace sb
