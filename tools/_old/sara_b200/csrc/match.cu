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
//     thread owns one query row -- and reduce every chunk of 32 columns to its two smallest keys
//     |b|^2 - 2 a.b (min / max only, the column index rides in the low mantissa bits): 1/16 of the
//     N1 x N2 matrix reaches memory, as one float2 per (chunk, query).  A warp per query then picks
//     the 8 smallest of its row and the floor of everything that was dropped.
//  2. EXACT re-ranking.  A warp per query recomputes the FLANN distance of its 8 candidates in
//     fp32 (separate multiply and add, the library is compiled with -fmad=false), orders them by
//     (distance, index) -- FLANN's result set keeps the first of two equal distances -- and
//     certifies the result: if the floor of the dropped keys, minus a bound on the bf16 / fp32
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
    // 128 with zeros; norm[r] = sum x^2 (padded rows: 1e30, never among the best).
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
        norm[warp] = warp < n ? s : 1e30f;
    }

    // ------------------------------------------------------------------------------------------
    // tcgen05 candidate kernel.
    constexpr int BLK_BYTES = 128 * 128;          // one 64-column block of a tile: 128 rows x 128 bytes
    constexpr int TILE_BYTES = 4 * BLK_BYTES;     // 128 rows x 256 bf16 (hi | lo)
    constexpr int B_STAGES = 2;
    constexpr int MMA_SMEM = (1 + B_STAGES) * TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    constexpr int TMEM_COLS = 256;                // two accumulator stages of 128 columns
    constexpr float kBig = 1e30f;                 // |b|^2 of a padding row, and "no key": finite, so that tagging cannot make a NaN

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
    // (bits(key) & mask) | tag as a single LOP3 (truth table 0xEA = (a & b) | c)
    __device__ __forceinline__ float tag_key(float key, unsigned mask, int tag)
    {
      unsigned r;
      asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(__float_as_uint(key)), "r"(mask), "r"(static_cast<unsigned>(tag)));
      return __uint_as_float(r);
    }
    // (m1 <= m2 <= m3) <- the three smallest of {m1, m2, m3, x}; min / max only
    __device__ __forceinline__ void insert3(float& m1, float& m2, float& m3, float x)
    {
      const float t1 = fmaxf(m1, x);
      m1 = fminf(m1, x);
      const float t2 = fmaxf(m2, t1);
      m2 = fminf(m2, t1);
      m3 = fminf(m3, t2);
    }
    __device__ __forceinline__ void tmem_ld_wait()
    {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    __device__ __forceinline__ void tmem_ld32_issue(unsigned taddr, unsigned (&r)[32])
    {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
    }

    // grid = (query tiles, splits), 384 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
    // allocator, warps 4-11 epilogue (warp w reads TMEM lanes 32 (w % 4) ..).
    __global__ void __launch_bounds__(384, 1)
        knn_mma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                       const float* __restrict__ norm_d, int n_tiles_d, int splits, int nq_pad, float2* __restrict__ pairs,
                       float* __restrict__ floor_split)
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
          mbar_init(acc_empty + s, 8);
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
        // ---- epilogue: thread = one query row, no data-dependent branch ----
        // A chunk of 32 accumulator columns becomes 32 keys |b|^2 - 2 a.b whose five low mantissa
        // bits are replaced by the column's position in the chunk (a perturbation of <= 31 ulp of
        // an approximate number); the three smallest keys of the chunk are tracked with min / max
        // only.  The two smallest are stored as one float2, pairs[chunk][query], for select_kernel;
        // the third is a lower bound of everything the chunk dropped and only its running minimum
        // is kept (floor_split[split][query]).
        // Eight epilogue warps: warps 4-7 take the chunks 0, 1 of a tile, warps 8-11 the chunks 2, 3
        // (a warp reads the TMEM lanes 32 (warp % 4) ..: both groups cover the 128 rows).
        const int ew = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = 32 * ew + lane;  // TMEM lane
        const size_t qrow = static_cast<size_t>(blockIdx.x) * 128 + row;
        float floor3 = kBig;  // the smallest THIRD key of any chunk: everything a chunk dropped is >= it
        unsigned tag_mask;    // ~31 in a register the optimiser cannot see through: (bits & mask) | tag is ONE lop3
        asm volatile("mov.u32 %0, 0xffffffe0;" : "=r"(tag_mask));
        for (int i = 0; i < n_my; ++i)
        {
          const int as = i & 1;
          const unsigned aph = (i >> 1) & 1;
          mbar_wait(acc_full + as, aph);
          tc_fence_after();
          const int jbase = (t0 + i) * 128;
#pragma unroll 1
          for (int c = 2 * half; c < 2 * half + 2; ++c)
          {
            unsigned r[32];
            __syncwarp();
            tmem_ld32_issue(tmem_base + (static_cast<unsigned>(32 * ew) << 16) + as * 128 + 32 * c, r);
            float4 nb[8];  // |b|^2 of the 32 columns: the same addresses for every lane, in flight under the TMEM load
            const float4* nb4 = reinterpret_cast<const float4*>(norm_d + jbase + 32 * c);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              nb[g] = __ldg(nb4 + g);
            tmem_ld_wait();
            // two independent (smallest, second, third) chains over the even / odd columns
            float a1 = kBig, a2 = kBig, a3 = kBig, b1 = kBig, b2 = kBig, b3 = kBig;
#pragma unroll
            for (int g = 0; g < 8; ++g)
            {
              const float n4[4] = {nb[g].x, nb[g].y, nb[g].z, nb[g].w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
              {
                const float key = __fmaf_rn(-2.f, __uint_as_float(r[4 * g + e]), n4[e]);
                const float x = tag_key(key, tag_mask, 4 * g + e);
                if (e & 1)
                  insert3(b1, b2, b3, x);
                else
                  insert3(a1, a2, a3, x);
              }
            }
            insert3(a1, a2, a3, b1);
            insert3(a1, a2, a3, b2);
            insert3(a1, a2, a3, b3);
            pairs[(static_cast<size_t>(jbase / 32 + c)) * nq_pad + qrow] = make_float2(a1, a2);
            floor3 = fminf(floor3, a3);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0)
            mbar_arrive_plain(acc_empty + as);
        }
        floor_split[(static_cast<size_t>(blockIdx.y) * 2 + half) * nq_pad + qrow] = floor3;
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
    // Exact search for a FEW queries (the uncertified ones): a block per query, a thread per data
    // row (strided), the query broadcast from shared memory; thread-private sorted lists, merged
    // by KC rounds of a block-wide minimum over (distance, index).
    // grid.y slices of the data range share a query; every block writes its KC best as exact
    // candidates, cand[(q * slices + slice) * KC + i], for rerank_kernel to merge.
    __global__ void __launch_bounds__(1024)
        knn_few_kernel(const float* __restrict__ queries, const int* __restrict__ qlist, const float* __restrict__ data,
                       int nd, int dim, float* __restrict__ cand_key, int* __restrict__ cand_idx)
    {
      extern __shared__ float smem[];
      float* qv = smem;  // dim floats
      __shared__ float r_d[32];
      __shared__ int r_i[32];
      const int q = qlist[blockIdx.x];
      for (int d = threadIdx.x; d < dim; d += blockDim.x)
        qv[d] = queries[static_cast<size_t>(q) * dim + d];
      __syncthreads();
      float kd[KC];
      int ki[KC];
#pragma unroll
      for (int i = 0; i < KC; ++i)
      {
        kd[i] = FLT_MAX;
        ki[i] = -1;
      }
      const bool vec = (dim & 3) == 0 && (reinterpret_cast<uintptr_t>(data) & 15) == 0;
      const int slices = gridDim.y;
      const int per = (nd + slices - 1) / slices;
      const int j_end = min(nd, (static_cast<int>(blockIdx.y) + 1) * per);
      for (int j = blockIdx.y * per + threadIdx.x; j < j_end; j += blockDim.x)
      {
        const float* b = data + static_cast<size_t>(j) * dim;
        float result = 0.f;
        int i = 0;
        if (vec)
        {
          // eight 16-byte loads in flight per step; the additions keep FLANN's order
          for (; i + 31 < dim; i += 32)
          {
            float4 x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
              x[u] = __ldg(reinterpret_cast<const float4*>(b + i) + u);
#pragma unroll
            for (int u = 0; u < 8; ++u)
            {
              const float d0 = x[u].x - qv[i + 4 * u], d1 = x[u].y - qv[i + 4 * u + 1];
              const float d2 = x[u].z - qv[i + 4 * u + 2], d3 = x[u].w - qv[i + 4 * u + 3];
              result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            }
          }
          for (; i + 3 < dim; i += 4)
          {
            const float4 x = __ldg(reinterpret_cast<const float4*>(b + i));
            const float d0 = x.x - qv[i], d1 = x.y - qv[i + 1], d2 = x.z - qv[i + 2], d3 = x.w - qv[i + 3];
            result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          }
        }
        else
          for (; i + 3 < dim; i += 4)
          {
            const float d0 = b[i] - qv[i], d1 = b[i + 1] - qv[i + 1], d2 = b[i + 2] - qv[i + 2], d3 = b[i + 3] - qv[i + 3];
            result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          }
        for (; i < dim; ++i)
        {
          const float d0 = b[i] - qv[i];
          result += d0 * d0;
        }
        if (result < kd[KC - 1])
          insert_sorted<KC>(kd, ki, result, j);  // j ascends within a thread: equal distances keep the lower index first
      }
      // k rounds: the block-wide smallest (distance, index) among the list heads; its owner pops it
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      for (int round = 0; round < KC; ++round)
      {
        float bd = kd[0];
        int bi = ki[0] >= 0 ? ki[0] : INT_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          const float od = __shfl_xor_sync(0xffffffffu, bd, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (od < bd || (od == bd && oi < bi))
          {
            bd = od;
            bi = oi;
          }
        }
        if (lane == 0)
        {
          r_d[warp] = bd;
          r_i[warp] = bi;
        }
        __syncthreads();
        float gd = r_d[0];
        int gi = r_i[0];
        for (int w2 = 1; w2 < static_cast<int>(blockDim.x >> 5); ++w2)
          if (r_d[w2] < gd || (r_d[w2] == gd && r_i[w2] < gi))
          {
            gd = r_d[w2];
            gi = r_i[w2];
          }
        __syncthreads();
        if (threadIdx.x == 0)
        {
          const size_t at = (static_cast<size_t>(q) * slices + blockIdx.y) * KC + round;
          cand_idx[at] = gi == INT_MAX ? -1 : gi;
          cand_key[at] = gi == INT_MAX ? FLT_MAX : gd;
        }
        if (ki[0] == gi && gi != INT_MAX)  // pop
        {
#pragma unroll
          for (int i = 0; i + 1 < KC; ++i)
          {
            kd[i] = kd[i + 1];
            ki[i] = ki[i + 1];
          }
          kd[KC - 1] = FLT_MAX;
          ki[KC - 1] = -1;
        }
      }
    }

    // ------------------------------------------------------------------------------------------
    // Selection: the KC smallest tagged keys of every query row.  Block = 32 queries x 8 parts: a
    // lane is a query (the float2 reads of a warp are 256 contiguous bytes), warp p scans the chunks
    // c = p (mod 8) into a thread-private sorted list; the lists meet in shared memory and warp 0
    // merges the eight lists of its queries.  Output: cand_key / cand_idx[q * KC + i] (data index
    // decoded from the chunk number and the five tag bits) and floor_key[q] = the smallest key that
    // was dropped anywhere: min(third keys of the chunks, the KC-th key taken here).
    constexpr int SEL_PARTS = 8;
    __global__ void __launch_bounds__(32 * SEL_PARTS)
        select_kernel(const float2* __restrict__ pairs, const float* __restrict__ floor_split, int splits, int nq,
                      int nq_pad, int nd, int n_chunks, float* __restrict__ cand_key, int* __restrict__ cand_idx,
                      float* __restrict__ floor_key)
    {
      __shared__ float s_kd[SEL_PARTS][KC][32];
      __shared__ int s_ki[SEL_PARTS][KC][32];
      const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
      const int q = blockIdx.x * 32 + lane;  // < nq_pad: the padded rows hold finite garbage
      float kd[KC];
      int ki[KC];
#pragma unroll
      for (int i = 0; i < KC; ++i)
      {
        kd[i] = kBig;
        ki[i] = -1;
      }
      for (int c0 = part; c0 < n_chunks; c0 += 4 * SEL_PARTS)
      {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          const int c = c0 + u * SEL_PARTS;
          v[u] = c < n_chunks ? __ldg(pairs + static_cast<size_t>(c) * nq_pad + q) : make_float2(kBig, kBig);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          const int c = c0 + u * SEL_PARTS;
          if (v[u].x < kd[KC - 1])
          {
            insert_sorted<KC>(kd, ki, v[u].x, c * 32 + (__float_as_int(v[u].x) & 31));
            if (v[u].y < kd[KC - 1])
              insert_sorted<KC>(kd, ki, v[u].y, c * 32 + (__float_as_int(v[u].y) & 31));
          }
        }
      }
#pragma unroll
      for (int i = 0; i < KC; ++i)
      {
        s_kd[part][i][lane] = kd[i];
        s_ki[part][i][lane] = ki[i];
      }
      __syncthreads();
      if (part != 0)
        return;
      for (int p = 1; p < SEL_PARTS; ++p)
#pragma unroll 1
        for (int i = 0; i < KC; ++i)
        {
          const float x = s_kd[p][i][lane];
          if (!(x < kd[KC - 1]))
            break;  // the lists are sorted
          insert_sorted<KC>(kd, ki, x, s_ki[p][i][lane]);
        }
      if (q >= nq)
        return;
      float fl = kd[KC - 1];
      for (int sp = 0; sp < splits; ++sp)
        fl = fminf(fl, floor_split[static_cast<size_t>(sp) * nq_pad + q]);
      floor_key[q] = fl;
#pragma unroll
      for (int i = 0; i < KC; ++i)
      {
        const bool real = kd[i] < kBig * 0.5f && ki[i] >= 0 && ki[i] < nd;
        cand_key[static_cast<size_t>(q) * KC + i] = real ? kd[i] : kBig;
        cand_idx[static_cast<size_t>(q) * KC + i] = real ? ki[i] : -1;
      }
    }

    // ------------------------------------------------------------------------------------------
    // Re-ranking: a warp per query, one candidate per lane (splits * KC <= 32).
    //   approx != 0: the KC candidates of select_kernel (splits = 1): recompute the exact distance of
    //                every candidate and certify against floor_key; uncertified queries are appended to `redo`.
    //   approx == 0: keys are exact distances already.
    // Output: k entries per query, ascending (distance, index); unused entries (-1, FLT_MAX).
    __global__ void __launch_bounds__(256)
        rerank_kernel(const float* __restrict__ queries, const int* __restrict__ qlist, int nq,
                      const float* __restrict__ data, int nd, int dim, int splits, int k, int approx,
                      const float* __restrict__ norm_q, const float* __restrict__ floor_key,
                      const float* __restrict__ cand_key, const int* __restrict__ cand_idx, int* __restrict__ out_idx,
                      float* __restrict__ out_dist, int* __restrict__ redo, int* __restrict__ n_redo)
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
        // everything the candidate pass dropped has a key >= floor_key[q] (select_kernel); a floor of
        // kBig means nothing real was dropped
        const float thr = floor_key[q];
        const float na = norm_q[q];
        const float thr_dist = na + thr;
        const float eps = 2.5e-4f * (fabsf(na) + fabsf(thr_dist)) + 1.f;  // >> the bf16-split / fp32 error, ~3e-5 (|a|^2 + |b|^2)
        const bool kept_all = !(thr < kBig * 0.5f);
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
      int mma_splits = 1;
      if (use_mma)
      {
        // one CTA per SM; CTAs of a query tile write disjoint chunks, so any split count works:
        // pick the one that minimises (waves) x (tiles per CTA)
        long best = LONG_MAX;
        for (int sp = 1; sp <= 16 && sp <= d_tiles; ++sp)
        {
          const long cost = static_cast<long>((q_tiles * sp + 147) / 148) * ((d_tiles + sp - 1) / sp);
          if (cost < best)
          {
            best = cost;
            mma_splits = sp;
          }
        }
      }
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
        add(static_cast<size_t>(nq_pad) * 4);                      // floor keys
        add(static_cast<size_t>(nq_pad) * 32 * 4);                 // third-key floors per (split, epilogue half)
        add(static_cast<size_t>(nd_pad / 32) * nq_pad * 8);        // two best keys per (chunk, query)
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
        float* floor_key = carve<float>(cur, nq_pad);
        float* floor_split = carve<float>(cur, static_cast<size_t>(nq_pad) * 32);
        float2* pairs = carve<float2>(cur, static_cast<size_t>(nd_pad / 32) * nq_pad);
        CUtensorMap map_q, map_d;
        if (!encode_map(&map_q, qb, nq_pad) || !encode_map(&map_d, db, nd_pad))
        {
          snprintf(err, errlen, "knn: cuTensorMapEncodeTiled failed");
          return SARA_B200_ERR_CUDA;
        }
        cudaMemsetAsync(n_redo, 0, 4, st);
        split_bf16_kernel<<<(nq_pad * 32 + 255) / 256, 256, 0, st>>>(d_q, nq, nq_pad, qb, norm_q);
        split_bf16_kernel<<<(nd_pad * 32 + 255) / 256, 256, 0, st>>>(d_data, nd, nd_pad, db, norm_d);
        knn_mma_kernel<<<dim3(q_tiles, mma_splits), 384, MMA_SMEM, st>>>(map_q, map_d, norm_d, d_tiles, mma_splits, nq_pad,
                                                                         pairs, floor_split);
        select_kernel<<<nq_pad / 32, 32 * SEL_PARTS, 0, st>>>(pairs, floor_split, 2 * mma_splits, nq, nq_pad, nd, nd_pad / 32,
                                                              cand_key, cand_idx, floor_key);
        rerank_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(d_q, nullptr, nq, d_data, nd, dim, 1, k, 1, norm_q, floor_key,
                                                             cand_key, cand_idx, d_idx, d_dist, redo, n_redo);
        launches += 5;
        splits = mma_splits;
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
          if (h_redo <= 4096)
          {
            const int slices = nd >= 4096 ? MAX_SPLITS : 1;
            knn_few_kernel<<<dim3(h_redo, slices), 1024, dim * sizeof(float), st>>>(d_q, redo, d_data, nd, dim, cand_key,
                                                                                  cand_idx);
            rerank_kernel<<<(h_redo * 32 + 255) / 256, 256, 0, st>>>(d_q, redo, h_redo, d_data, nd, dim, slices, k, 0, nullptr,
                                                                   nullptr, cand_key, cand_idx, d_idx, d_dist, nullptr,
                                                                   nullptr);
            launches += 2;
          }
          else
          {
            int s2 = 1;
            while (s2 < MAX_SPLITS && ((h_redo + QB - 1) / QB) * s2 * 2 <= 2 * 148 && nd >= 2 * s2 * 4 * TJ)
              s2 *= 2;
            knn_exact_kernel<0><<<dim3((h_redo + QB - 1) / QB, s2), QB, ex_smem, st>>>(
                d_q, redo, h_redo, d_data, nd, dim, s2, cand_key, cand_idx, nullptr, nullptr, nullptr, nullptr, nullptr,
                nullptr);
            rerank_kernel<<<(h_redo * 32 + 255) / 256, 256, 0, st>>>(d_q, redo, h_redo, d_data, nd, dim, s2, k, 0, nullptr,
                                                                   nullptr, cand_key, cand_idx, d_idx, d_dist, nullptr,
                                                                   nullptr);
            launches += 2;
          }
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
        rerank_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(d_q, nullptr, nq, d_data, nd, dim, splits, k, 0, nullptr, nullptr,
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
