// Device-wide exclusive scan of an int32 array whose length may live in device
// memory (data-dependent counts never come back to the host mid-frame).
// Two launches: per-chunk scans (1024 items per chunk), then one block scans the
// chunk totals.  The global offset of item i is out_off[i] + chunk_off[i >> 10].
// Integer adds only => deterministic, order-preserving compaction.
#pragma once
#include "common.cuh"

namespace sb {

  namespace scan_detail {

    // Exclusive scan of one value per thread across a 1024-thread block.
    __device__ __forceinline__ int block_exclusive_scan_1024(int v, int* total)
    {
      __shared__ int warp_sums[32];
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d)
          inc += t;
      }
      if (lane == 31)
        warp_sums[wid] = inc;
      __syncthreads();
      if (wid == 0)
      {
        int ws = warp_sums[lane];
        int winc = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
          const int t = __shfl_up_sync(0xffffffffu, winc, d);
          if (lane >= d)
            winc += t;
        }
        warp_sums[lane] = winc - ws;  // exclusive
        if (lane == 31)
          *total = winc;
      }
      __syncthreads();
      const int r = warp_sums[wid] + inc - v;
      __syncthreads();
      return r;
    }

    static __global__ void __launch_bounds__(1024)
        scan_chunks_kernel(const int* __restrict__ vals, int* __restrict__ out_off,
                           int* __restrict__ chunk_tot, int n_static, const int* __restrict__ n_ptr,
                           int n_cap)
    {
      __shared__ int total;
      const int n = n_ptr ? min(*n_ptr, n_cap) : n_static;
      const int n_chunks = (n + 1023) >> 10;
      for (int c = blockIdx.x; c < n_chunks; c += gridDim.x)
      {
        const int i = (c << 10) + threadIdx.x;
        const int v = i < n ? vals[i] : 0;
        const int e = block_exclusive_scan_1024(v, &total);
        if (i < n)
          out_off[i] = e;
        if (threadIdx.x == 0)
          chunk_tot[c] = total;
        __syncthreads();
      }
    }

    static __global__ void __launch_bounds__(1024)
        scan_totals_kernel(int* __restrict__ chunk_tot, int n_static, const int* __restrict__ n_ptr,
                           int n_cap, int* __restrict__ total_out, int total_cap,
                           int* __restrict__ overflow, int overflow_bit)
    {
      __shared__ int total;
      const int n = n_ptr ? min(*n_ptr, n_cap) : n_static;
      const int n_chunks = (n + 1023) >> 10;  // <= 1024
      const int v = static_cast<int>(threadIdx.x) < n_chunks ? chunk_tot[threadIdx.x] : 0;
      const int e = block_exclusive_scan_1024(v, &total);
      if (static_cast<int>(threadIdx.x) < n_chunks)
        chunk_tot[threadIdx.x] = e;
      if (threadIdx.x == 0)
      {
        if (total > total_cap)
          atomicOr(overflow, overflow_bit);
        *total_out = total;  // the true count; consumers clamp to their capacity
      }
    }

    // Small arrays (the usual case: row counters, keep flags, orientation counts): one block,
    // one launch.  The array is consumed in tiles of 16384 = 4 sub-tiles of 4096; thread t owns
    // items 4t .. 4t + 3 of every sub-tile (16-byte coalesced loads, all four issued together), a
    // running carry links sub-tiles and tiles.  Final offsets go to out_off and the chunk offsets
    // are zero, so consumers keep using out_off[i] + chunk_off[i >> 10].  vals and out_off must be
    // 16-byte aligned (they are: cudaMalloc'd arrays at 256-byte offsets).
    static __global__ void __launch_bounds__(1024)
        scan_single_kernel(const int* __restrict__ vals, int* __restrict__ out_off, int* __restrict__ chunk_off,
                           int n_static, const int* __restrict__ n_ptr, int n_cap, int* __restrict__ total_out,
                           int total_cap, int* __restrict__ overflow, int overflow_bit)
    {
      __shared__ int total;
      const int n = n_ptr ? min(*n_ptr, n_cap) : n_static;
      const bool vec = ((reinterpret_cast<uintptr_t>(vals) | reinterpret_cast<uintptr_t>(out_off)) & 15) == 0;
      int carry = 0;
      for (int base = 0; base < n; base += 16384)
      {
        int4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          const int lo = base + 4096 * j + 4 * static_cast<int>(threadIdx.x);
          if (vec && lo + 3 < n)
            v[j] = *reinterpret_cast<const int4*>(vals + lo);
          else
          {
            v[j].x = lo < n ? vals[lo] : 0;
            v[j].y = lo + 1 < n ? vals[lo + 1] : 0;
            v[j].z = lo + 2 < n ? vals[lo + 2] : 0;
            v[j].w = lo + 3 < n ? vals[lo + 3] : 0;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          const int lo = base + 4096 * j + 4 * static_cast<int>(threadIdx.x);
          if (base + 4096 * j >= n)  // block-uniform
            break;
          const int sum = v[j].x + v[j].y + v[j].z + v[j].w;
          const int run = carry + block_exclusive_scan_1024(sum, &total);
          const int4 o = make_int4(run, run + v[j].x, run + v[j].x + v[j].y, run + v[j].x + v[j].y + v[j].z);
          if (vec && lo + 3 < n)
            *reinterpret_cast<int4*>(out_off + lo) = o;
          else
          {
            if (lo < n)
              out_off[lo] = o.x;
            if (lo + 1 < n)
              out_off[lo + 1] = o.y;
            if (lo + 2 < n)
              out_off[lo + 2] = o.z;
            if (lo + 3 < n)
              out_off[lo + 3] = o.w;
          }
          carry += total;
          __syncthreads();
        }
      }
      for (int c = threadIdx.x; c <= (n >> 10); c += 1024)
        chunk_off[c] = 0;
      if (threadIdx.x == 0)
      {
        if (carry > total_cap)
          atomicOr(overflow, overflow_bit);
        *total_out = carry;  // the true count; consumers clamp to their capacity
      }
    }

  }  // namespace scan_detail

  // Returns the number of kernels launched.
  static inline int exclusive_scan(const int* vals, int* out_off, int* chunk_off, int n_static,
                                   const int* n_ptr, int n_cap, int* total_out, int total_cap,
                                   int* overflow, int overflow_bit, cudaStream_t st)
  {
    const int max_n = n_ptr ? n_cap : n_static;
    if (max_n <= (1 << 20) && vals != out_off)
    {
      scan_detail::scan_single_kernel<<<1, 1024, 0, st>>>(vals, out_off, chunk_off, n_static, n_ptr, n_cap, total_out,
                                                          total_cap, overflow, overflow_bit);
      return 1;
    }
    int grid = (max_n + 1023) >> 10;
    if (grid < 1)
      grid = 1;
    if (grid > 1024)
      grid = 1024;
    scan_detail::scan_chunks_kernel<<<grid, 1024, 0, st>>>(vals, out_off, chunk_off, n_static, n_ptr,
                                                           n_cap);
    scan_detail::scan_totals_kernel<<<1, 1024, 0, st>>>(chunk_off, n_static, n_ptr, n_cap, total_out,
                                                        total_cap, overflow, overflow_bit);
    return 2;
  }

}  // namespace sb
