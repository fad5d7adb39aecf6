// Device-side nearest-neighbour search of the matching row (match.cu).  Product code.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace sb {
  namespace match {

    struct Workspace  // grown on demand, owned by the context
    {
      unsigned char* buf = nullptr;
      size_t bytes = 0;
      void release();
    };

    struct KnnStats
    {
      int used_tensor_cores = 0;
      int n_redone = 0;   // queries the certificate sent to the exact scalar kernel
      int launches = 0;
      int splits = 0;
    };

    bool mma_path_available();
    int knn(Workspace& ws, const float* d_q, int nq, const float* d_data, int nd, int dim, int k, int mode, int* d_idx,
            float* d_dist, KnnStats* stats, cudaStream_t st, char* err, size_t errlen);
    int radius_pass(const float* d_q, int nq, const float* d_data, int nd, int dim, const float* d_radius, int* d_count,
                    const int* d_off, int* d_out_idx, float* d_out_dist, cudaStream_t st, char* err, size_t errlen);

  }  // namespace match
}  // namespace sb
