// Does the operand form of FFMA2 / FMUL2 change their issue rate?  (B200, sm_100a)
//   variant 0: scalar multiplier from a UNIFORM register (kernel parameter)      FFMA2 R, R.pair, UR.F32, R.pair
//   variant 1: scalar multiplier from a per-thread VECTOR register               FFMA2 R, R.pair, R.F32,  R.pair
//   variant 2: multiplier is a full 64-bit register pair                         FFMA2 R, R.pair, R.pair, R.pair
// Each thread runs NACC independent accumulator chains; 8 warps per SM x 148 blocks x 4 = plenty.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fma2_operands fma2_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
constexpr int NACC = 12;
template <int V>
__global__ void __launch_bounds__(256) k(const float* in, float* out, float s_param, int iters)
{
  float s = s_param;
  u64 sp = pack2(s, s);
  if (V == 1) { s = in[threadIdx.x & 31]; sp = pack2(s, s); }
  if (V == 2) { sp = pack2(in[threadIdx.x & 31], in[(threadIdx.x + 1) & 31]); }
  u64 acc[NACC], p[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = pack2(in[i], in[i + 1]); p[i] = pack2(in[i + 2], in[i + 3]); }
  for (int it = 0; it < iters; ++it)
  {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma2(acc[i], sp, p[i]);
#pragma unroll
    for (int i = 0; i < NACC; ++i) p[i] = mul2(acc[(i + 1) % NACC], sp);
  }
  u64 r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r ^= acc[i] ^ p[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)r ^ (unsigned)(r >> 32));
}
template <int V> void run(const char* name, const float* in, float* out, int warps_per_sm)
{
  const int iters = 4000, blocks = 148 * (warps_per_sm / 8 > 0 ? warps_per_sm / 8 : 1), threads = warps_per_sm >= 8 ? 256 : 32 * warps_per_sm;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<V><<<blocks, threads>>>(in, out, 1.0f, 10);
  cudaEventRecord(a);
  k<V><<<blocks, threads>>>(in, out, 1.0f, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double inst = double(blocks) * (threads / 32) * iters * 2.0 * NACC;   // packed warp-instructions
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s warps/SM %2d: %.3f ms, %.3f packed instr / clk / SMSP (at %d MHz nominal), %.1f lane-ops/clk/SM\n", name, warps_per_sm, ms,
         inst / (ms * 1e-3) / (clk * 1e3) / (148 * 4), clk / 1000, inst * 64 / (ms * 1e-3) / (clk * 1e3) / 148);
}
int main()
{
  float *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 8 * 256 * 4);
  float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-3f; h[0] = 1.0f; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  for (int w : {4, 8, 16, 32}) { run<0>("UR scalar multiplier", in, out, w); run<1>("R scalar multiplier", in, out, w); run<2>("R pair multiplier", in, out, w); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
