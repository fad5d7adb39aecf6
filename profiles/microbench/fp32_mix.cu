// Can packed f32x2 arithmetic (fmaheavy pipe) and scalar fp32 arithmetic (fmalite pipe) run side
// by side?  P packed accumulator chains + SC scalar chains per thread, un-fused multiply and add.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32_mix.bin fp32_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
struct Taps { float v[32]; };
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int P, int SC>
__global__ void __launch_bounds__(256) k_mix(float* out, const __grid_constant__ Taps t, int iters)
{
  u64 pa[P > 0 ? P : 1], pb[P > 0 ? P : 1];
  float sa[SC > 0 ? SC : 1], sb[SC > 0 ? SC : 1];
#pragma unroll
  for (int i = 0; i < P; ++i) { pa[i] = 0ull; pb[i] = pack2(0.5f + threadIdx.x * 1e-3f + i, 0.25f + i); }
#pragma unroll
  for (int i = 0; i < SC; ++i) { sa[i] = 0.f; sb[i] = 0.75f + threadIdx.x * 1e-3f + i; }
  const u64 one = pack2(t.v[31], t.v[31]);
  for (int it = 0; it < iters; ++it)
  {
#pragma unroll
    for (int j = 0; j < 25; ++j)
    {
      const u64 kk = pack2(t.v[j], t.v[j]);
#pragma unroll
      for (int i = 0; i < (P > SC ? P : SC); ++i)
      {
        if (i < P) pa[i] = fma2(pa[i], one, mul2(pb[(i + j) % (P > 0 ? P : 1)], kk));
        if (i < SC) sa[i] = __fadd_rn(sa[i], __fmul_rn(sb[(i + j) % (SC > 0 ? SC : 1)], t.v[j]));
      }
    }
#pragma unroll
    for (int i = 0; i < P; ++i) { pb[i] = pa[i]; pa[i] = 0ull; }
#pragma unroll
    for (int i = 0; i < SC; ++i) { sb[i] = sa[i]; sa[i] = 0.f; }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < P; ++i) s += __uint_as_float((unsigned) pb[i]) + __uint_as_float((unsigned) (pb[i] >> 32));
#pragma unroll
  for (int i = 0; i < SC; ++i) s += sb[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int P, int SC>
void run(float* out, const Taps& t, int clk_khz)
{
  const int iters = 2000, grid = 148 * 2;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_mix<P, SC><<<grid, 256>>>(out, t, iters); cudaDeviceSynchronize();
  cudaEventRecord(a); k_mix<P, SC><<<grid, 256>>>(out, t, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = 2.0 * 25 * (2 * P + SC) * iters * (double) grid * 256;
  printf("packed chains %d + scalar chains %d: %7.3f ms  %6.1f fp32 lane-ops/clk/SM  (%.1f Tops/s)\n", P, SC, ms,
         ops / (ms * 1e-3) / (clk_khz * 1e3) / 148, ops / (ms * 1e-3) / 1e12);
}

int main()
{
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  Taps t; for (int i = 0; i < 32; ++i) t.v[i] = (i + 1) / 325.f; t.v[31] = 1.f;
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  run<4, 0>(out, t, clk_khz);
  run<8, 0>(out, t, clk_khz);
  run<0, 8>(out, t, clk_khz);
  run<4, 4>(out, t, clk_khz);
  run<4, 2>(out, t, clk_khz);
  run<6, 2>(out, t, clk_khz);
  run<4, 8>(out, t, clk_khz);
  run<6, 6>(out, t, clk_khz);
  run<8, 4>(out, t, clk_khz);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
