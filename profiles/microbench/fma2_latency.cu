// Dependent-issue distance of packed fp32 (B200, sm_100a): a warp runs W independent
// "multiply -> add" pairs per step; the add consumes the product issued D packed instructions
// earlier.  Reports packed instructions per clock and SMSP for 1, 2 and 4 warps per SMSP.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fma2_latency fma2_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
constexpr int N = 16;  // accumulators
// D = number of packed instructions between a multiply and the add that uses its product.
template <int D>
__global__ void __launch_bounds__(128) k(const float* in, float* out, float s_param, int iters)
{
  const u64 sp = pack2(s_param, s_param);
  u64 acc[N], v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { acc[i] = pack2(in[i], in[i + 1]); v[i] = pack2(in[i + 2], in[i + 3]); }
  for (int it = 0; it < iters; ++it)
  {
    // groups of D multiplies followed by the D adds that use them: each add sits D instructions after its multiply
    u64 p[N];
#pragma unroll
    for (int g = 0; g < N; g += D)
    {
#pragma unroll
      for (int i = g; i < g + D && i < N; ++i) p[i] = mul2(acc[(i + N / 2) % N], sp);  // operand written >= N/2 pairs ago
#pragma unroll
      for (int i = g; i < g + D && i < N; ++i) acc[i] = fma2(acc[i], sp, p[i]);
    }
  }
  u64 r = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) r ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)r ^ (unsigned)(r >> 32));
}
template <int D> void run(const float* in, float* out)
{
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("distance %2d:", D);
  for (int wps : {1, 2, 4})
  {
    const int iters = 3000, blocks = 148 * wps, threads = 128;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<D><<<blocks, threads>>>(in, out, 1.0f, 10);
    cudaEventRecord(a);
    k<D><<<blocks, threads>>>(in, out, 1.0f, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double inst = double(blocks) * 4 * iters * 2.0 * N;
    printf("  %d warp/SMSP %.3f instr/clk", wps, inst / (ms * 1e-3) / (clk * 1e3) / (148 * 4));
  }
  printf("\n");
}
int main()
{
  float *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 4 * 128 * 4);
  float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-3f; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  run<1>(in, out); run<2>(in, out); run<4>(in, out); run<8>(in, out); run<16>(in, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
