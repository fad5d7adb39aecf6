// Device helpers shared by the pyramid kernels: packed fp32x2 arithmetic with the
// reference's rounding (separate multiply and add), TMA tile loads and mbarriers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace sb {
  namespace fused {

    typedef unsigned long long u64;

    // ---- packed fp32x2 arithmetic -----------------------------------------------
    __device__ __forceinline__ u64 pack2(float lo, float hi)
    {
      u64 r;
      asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
      return r;
    }
    // Same, but never re-materialised: ptxas otherwise repeats the two register moves in front
    // of every packed instruction that reads the pair.
    __device__ __forceinline__ u64 pack2_once(float lo, float hi)
    {
      u64 r;
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
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


    __device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y)
    {
      asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(
                       reinterpret_cast<unsigned long long>(map)),
                   "r"(x), "r"(y)
                   : "memory");
    }

    // ---- cp.async (LDGSTS): 8-byte global -> shared copies, completion by commit groups -----
    __device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
    {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
    }
    __device__ __forceinline__ void cp_async_commit()
    {
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    template <int N>
    __device__ __forceinline__ void cp_async_wait()
    {
      asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
    }

    // Driver entry point of cuTensorMapEncodeTiled (no link-time dependency on libcuda).
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeTiledFn encode_fn();

  }  // namespace fused
}  // namespace sb
