// Micro-benchmark that decides the pyramid kernel design: how many separate
// fp32 multiplies + adds (no FMA: the reference arithmetic) one B200 SM retires
// per clock, scalar vs packed f32x2, with the tap as a constant-bank operand,
// and whether ptxas keeps the packed forms un-contracted (bit comparison).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32_issue.bin fp32_issue.cu
//
// Every variant evaluates the same function: P pairs of values; one round is
//   acc[2i+h] = sum_j b[2((i+j)%P)+h] * t[j]  (left to right, from 0),  b <- acc.
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { float v[32]; };
constexpr int K = 25;

__device__ __forceinline__ unsigned long long pack(float lo, float hi)
{
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo_of(unsigned long long v) { return __uint_as_float((unsigned) v); }
__device__ __forceinline__ float hi_of(unsigned long long v) { return __uint_as_float((unsigned) (v >> 32)); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// MODE 6: packed mul, then the add as fma2(acc, ONE, p) with ONE = (1.f, 1.f) read from the
// kernel parameters: acc * 1 + p rounds once, i.e. equals RN(acc + p), and ptxas cannot
// contract it with the multiply because it does not know ONE.
// MODE 0: scalar FMUL + FADD.  1: mul.rn.f32x2 + add.rn.f32x2.  2: scalar FMUL, packed add.
// 3: packed mul, scalar FADD.  4: fused fma (NOT the reference arithmetic; speed reference).
// 5: packed mul + packed add with the product laundered through a volatile asm.
template <int P, int MODE>
__global__ void __launch_bounds__(256) k_conv(float* out, const __grid_constant__ Taps t, int iters)
{
  float b[2 * P], acc[2 * P];
#pragma unroll
  for (int i = 0; i < 2 * P; ++i) b[i] = 0.5f + threadIdx.x * 0.001f + i * 0.01f;
  for (int it = 0; it < iters; ++it)
  {
#pragma unroll
    for (int i = 0; i < 2 * P; ++i) acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < P; ++i)
      {
        const int s = 2 * ((i + j) % P);
        if (MODE == 0)
        {
          acc[2 * i] = __fadd_rn(acc[2 * i], __fmul_rn(b[s], t.v[j]));
          acc[2 * i + 1] = __fadd_rn(acc[2 * i + 1], __fmul_rn(b[s + 1], t.v[j]));
        }
        else if (MODE == 4)
        {
          acc[2 * i] = __fmaf_rn(b[s], t.v[j], acc[2 * i]);
          acc[2 * i + 1] = __fmaf_rn(b[s + 1], t.v[j], acc[2 * i + 1]);
        }
        else
        {
          unsigned long long p;
          if (MODE == 2)
            p = pack(__fmul_rn(b[s], t.v[j]), __fmul_rn(b[s + 1], t.v[j]));
          else
            p = mul2(pack(b[s], b[s + 1]), pack(t.v[j], t.v[j]));
          if (MODE == 5)
            asm volatile("" : "+l"(p));
          if (MODE == 6)
          {
            const unsigned long long a = fma2(pack(acc[2 * i], acc[2 * i + 1]), pack(t.v[31], t.v[31]), p);
            acc[2 * i] = lo_of(a);
            acc[2 * i + 1] = hi_of(a);
          }
          else if (MODE == 3)
          {
            acc[2 * i] = __fadd_rn(acc[2 * i], lo_of(p));
            acc[2 * i + 1] = __fadd_rn(acc[2 * i + 1], hi_of(p));
          }
          else
          {
            const unsigned long long a = add2(pack(acc[2 * i], acc[2 * i + 1]), p);
            acc[2 * i] = lo_of(a);
            acc[2 * i + 1] = hi_of(a);
          }
        }
      }
#pragma unroll
    for (int i = 0; i < 2 * P; ++i) b[i] = acc[i];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * P; ++i) s = __fadd_rn(s, acc[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// smem-fed variant: the shape of the real column pass.  Each thread loads a
// window of KK + R - 1 values (stride-1 across lanes) and produces R outputs.
// PACKED: outputs r, r+1 share one f32x2 accumulator.
template <int KK, int R, bool PACKED>
__global__ void __launch_bounds__(256) k_colpass(float* out, const __grid_constant__ Taps t, int iters)
{
  __shared__ float ring[(KK + R - 1) * 256];
  for (int i = threadIdx.x; i < (KK + R - 1) * 256; i += 256) ring[i] = 0.5f + (i % 977) * 1e-4f;
  __syncthreads();
  float tot = 0.f;
  for (int it = 0; it < iters; ++it)
  {
    float win[KK + R - 1];
#pragma unroll
    for (int i = 0; i < KK + R - 1; ++i) win[i] = ring[((i + it) % (KK + R - 1)) * 256 + threadIdx.x];
    if (!PACKED)
    {
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
      for (int j = 0; j < KK; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r)
          acc[r] = __fadd_rn(acc[r], __fmul_rn(win[r + j], t.v[j]));
#pragma unroll
      for (int r = 0; r < R; ++r) tot = __fadd_rn(tot, acc[r]);
    }
    else
    {
      unsigned long long acc[R / 2];
#pragma unroll
      for (int r = 0; r < R / 2; ++r) acc[r] = pack(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < KK; ++j)
#pragma unroll
        for (int r = 0; r < R / 2; ++r)
        {
          unsigned long long p = mul2(pack(win[2 * r + j], win[2 * r + 1 + j]), pack(t.v[j], t.v[j]));
          asm volatile("" : "+l"(p));
          acc[r] = fma2(acc[r], pack(t.v[31], t.v[31]), p);
        }
#pragma unroll
      for (int r = 0; r < R / 2; ++r) tot = __fadd_rn(__fadd_rn(tot, lo_of(acc[r])), hi_of(acc[r]));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

template <typename F>
float time_it(F f)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

static float* g_out;
static Taps g_t;
static float h_ref[148 * 256], h_cmp[148 * 256];

template <int P, int MODE>
void run(const char* name, int bps, int clk_khz)
{
  const int iters = 2000, grid = 148 * bps;
  const float ms = time_it([&] { k_conv<P, MODE><<<grid, 256>>>(g_out, g_t, iters); });
  const double ops = 2.0 * K * 2 * P * iters * grid * 256;
  k_conv<P, 0><<<148, 256>>>(g_out, g_t, 7); cudaMemcpy(h_ref, g_out, sizeof(h_ref), cudaMemcpyDeviceToHost);
  k_conv<P, MODE><<<148, 256>>>(g_out, g_t, 7); cudaMemcpy(h_cmp, g_out, sizeof(h_cmp), cudaMemcpyDeviceToHost);
  int diff = 0;
  for (int i = 0; i < 148 * 256; ++i) diff += h_ref[i] != h_cmp[i];
  printf("%-34s P=%d blocks/SM=%d %8.3f ms %6.1f fp32-ops/clk/SM %6.2f Tops/s  bit-mismatch vs scalar %d/%d\n", name, P,
         bps, ms, ops / (ms * 1e-3) / (clk_khz * 1e3) / 148, ops / (ms * 1e-3) / 1e12, diff, 148 * 256);
}

int main()
{
  cudaMalloc(&g_out, 148 * 8 * 256 * 4);
  for (int i = 0; i < 32; ++i) g_t.v[i] = (i + 1) / 325.f;
  g_t.v[31] = 1.f;
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("nominal SM clock %d kHz\n", clk_khz);
  for (int bps = 1; bps <= 2; bps *= 2)
  {
    run<4, 0>("scalar FMUL+FADD", bps, clk_khz);
    run<8, 0>("scalar FMUL+FADD", bps, clk_khz);
    run<4, 1>("mul.rn.f32x2 + add.rn.f32x2", bps, clk_khz);
    run<8, 1>("mul.rn.f32x2 + add.rn.f32x2", bps, clk_khz);
    run<4, 5>("mul2 + add2, laundered product", bps, clk_khz);
    run<8, 5>("mul2 + add2, laundered product", bps, clk_khz);
    run<4, 6>("mul2 + fma2(acc, ONE, p)", bps, clk_khz);
    run<8, 6>("mul2 + fma2(acc, ONE, p)", bps, clk_khz);
    run<8, 2>("scalar FMUL + add2", bps, clk_khz);
    run<8, 3>("mul2 + scalar FADD", bps, clk_khz);
    run<8, 4>("fma (speed reference)", bps, clk_khz);
    const int iters = 2000, grid = 148 * bps;
    auto report = [&](const char* name, float ms, double ops_per_thread) {
      const double ops = ops_per_thread * grid * 256;
      printf("%-34s blocks/SM=%d %8.3f ms %6.1f fp32-ops/clk/SM %6.2f Tops/s\n", name, bps, ms,
             ops / (ms * 1e-3) / (clk_khz * 1e3) / 148, ops / (ms * 1e-3) / 1e12);
    };
    report("colpass K=25 R=8 scalar (smem)", time_it([&] { k_colpass<25, 8, false><<<grid, 256>>>(g_out, g_t, iters); }), 2.0 * 25 * 8 * iters);
    report("colpass K=25 R=8 packed (smem)", time_it([&] { k_colpass<25, 8, true><<<grid, 256>>>(g_out, g_t, iters); }), 2.0 * 25 * 8 * iters);
    report("colpass K=11 R=8 scalar (smem)", time_it([&] { k_colpass<11, 8, false><<<grid, 256>>>(g_out, g_t, iters); }), 2.0 * 11 * 8 * iters);
    report("colpass K=11 R=8 packed (smem)", time_it([&] { k_colpass<11, 8, true><<<grid, 256>>>(g_out, g_t, iters); }), 2.0 * 11 * 8 * iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
