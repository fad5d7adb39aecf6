// Probe: which TMA tile-mode coordinates are legal on sm_100a (negative / fully out-of-bounds boxes)?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
constexpr int BW = 108;
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, float* out)
{
  extern __shared__ __align__(1024) unsigned char raw[];
  float* sm = (float*) raw;
  unsigned long long* bar = (unsigned long long*) (sm + 8 * BW);
  if (threadIdx.x == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(8 * BW * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sm)), "l"((unsigned long long) &tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)) : "memory");
  for (int i = threadIdx.x; i < 8 * BW; i += blockDim.x) out[i] = sm[i];
}
int main()
{
  const int w = 640, h = 480, pitch = 640;
  float* d; cudaMalloc(&d, pitch * h * 4);
  std::vector<float> host(pitch * h);
  for (int i = 0; i < pitch * h; ++i) host[i] = 1.f + (i % pitch) + 1000.f * (i / pitch);
  cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 8 * BW * 4);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn) p;
  CUtensorMap tmap;
  cuuint64_t dims[2] = {w, h}, strides[1] = {pitch * 4};
  cuuint32_t box[2] = {BW, 8}, estr[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int) r);
  const int tests[][2] = {{0, 0}, {88, 7}, {88, -1}, {-44, 5}, {-44, -3}, {600, 470}, {700, 100}, {88, 479}, {88, 480}, {88, 500}, {88, -8}, {88, -20}, {-108, 3}, {-200, 3}, {90, 3}};
  for (auto& t : tests)
  {
    probe<<<1, 128, 8 * BW * 4 + 64>>>(tmap, t[0], t[1], out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> res(8 * BW);
    cudaMemcpy(res.data(), out, res.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < 8; ++yy) for (int xx = 0; xx < BW; ++xx)
    {
      const int gx = t[0] + xx, gy = t[1] + yy;
      const float exp = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? host[gy * pitch + gx] : 0.f;
      bad += res[yy * BW + xx] != exp;
    }
    printf("coords (%d, %d): %s, mismatches %d\n", t[0], t[1], cudaGetErrorString(e), bad);
    if (e != cudaSuccess) break;
  }
  return 0;
}
