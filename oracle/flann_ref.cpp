// Thin C shim over the reference's VENDORED FLANN (third-party/flann/src/cpp/flann, header only),
// compiled from the sources where they lie under /root/reference into oracle/_ref/libflann_ref.so
// (oracle/Makefile, target `ref`).  TEST INFRASTRUCTURE ONLY: it lets the tests run the real
// library AnnMatcher.cpp:228-237 calls -- flann::Index<flann::L2<float>> with KDTreeIndexParams{8}
// and default SearchParams -- and FLANN's own exact LinearIndex, against the restated search of
// oracle/match_oracle.cpp.  No reference source is copied: this file only calls FLANN's public API.
#include <flann/flann.hpp>

#include <vector>

namespace {
  struct Handle
  {
    flann::Matrix<float> data;
    flann::Index<flann::L2<float>>* index;
    int dim;
  };
}

extern "C" {

// kind 0: LinearIndexParams (exact); kind 1: KDTreeIndexParams{8} (what AnnMatcher builds).
void* flannref_build(const float* data, int n, int dim, int kind)
{
  auto* h = new Handle;
  h->data = flann::Matrix<float>(const_cast<float*>(data), n, dim);
  h->dim = dim;
  if (kind == 0)
    h->index = new flann::Index<flann::L2<float>>(h->data, flann::LinearIndexParams());
  else
    h->index = new flann::Index<flann::L2<float>>(h->data, flann::KDTreeIndexParams(8));
  h->index->buildIndex();
  return h;
}

void flannref_free(void* handle)
{
  auto* h = static_cast<Handle*>(handle);
  delete h->index;
  delete h;
}

// tree.knnSearch(query, indices, dists, k, SearchParams()) as in AnnMatcher.cpp:104,131
void flannref_knn(void* handle, const float* query, int k, int* idx, float* dist)
{
  auto* h = static_cast<Handle*>(handle);
  flann::Matrix<float> q(const_cast<float*>(query), 1, h->dim);
  flann::Matrix<int> I(idx, 1, k);
  flann::Matrix<float> D(dist, 1, k);
  flann::SearchParams params;
  h->index->knnSearch(q, I, D, k, params);
}

// tree.radiusSearch(query, indices, dists, radius, SearchParams()) as in AnnMatcher.cpp:145
int flannref_radius(void* handle, const float* query, float radius, int* idx, float* dist, int cap)
{
  auto* h = static_cast<Handle*>(handle);
  flann::Matrix<float> q(const_cast<float*>(query), 1, h->dim);
  flann::Matrix<int> I(idx, 1, cap);
  flann::Matrix<float> D(dist, 1, cap);
  flann::SearchParams params;
  return h->index->radiusSearch(q, I, D, radius, params);
}

// many queries at once (tests / timing)
void flannref_knn_batch(void* handle, const float* queries, int nq, int k, int* idx, float* dist)
{
  for (int i = 0; i < nq; ++i)
    flannref_knn(handle, queries + static_cast<size_t>(i) * static_cast<Handle*>(handle)->dim, k, idx + static_cast<size_t>(i) * k,
                 dist + static_cast<size_t>(i) * k);
}

}  // extern "C"
