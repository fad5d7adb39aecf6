// CPU oracle of the descriptor-matching row.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under sara_b200/ includes, links or executes this file;
// it is loaded by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
// --impl reference legs as the checker.
//
// What it restates (paths relative to /root/reference/cpp):
//   AnnMatcher::compute_matches     src/DO/Sara/FeatureMatching/AnnMatcher.cpp:219-282
//   append_nearest_neighbors        src/DO/Sara/FeatureMatching/AnnMatcher.cpp:57-171
//   KeyProximity::operator()        src/DO/Sara/FeatureMatching/KeyProximity.cpp:17-30
//   SquaredRefDistance::operator()  src/DO/Sara/Geometry/Tools/Metric.hpp:46-49
//   Match::operator==, OERegion::operator==   src/DO/Sara/Match/Match.hpp:159-162,
//                                             src/DO/Sara/Features/Feature.hpp:140-146
// and, from the reference's vendored third-party FLANN (third-party/flann/src/cpp/flann):
//   L2<float>::operator()           algorithms/dist.h:151-178     (groups of four, see l2_flann)
//   KNNSimpleResultSet::addPoint    util/result_set.h:151-171     (ties: first come first)
//   RadiusResultSet                 util/result_set.h:477-510     (dist < radius, sorted)
//   LinearIndex::findNeighbors      algorithms/linear_index.h     (scan in index order)
//
// The reference searches with a randomised KD-tree forest (KDTreeIndexParams{8}, default
// SearchParams: 32 checks), which is APPROXIMATE and seed dependent.  The restated search is the
// exact one the forest approximates (FLANN's own LinearIndex): every distance the forest reports
// is computed by the same L2 functor, so whenever the forest finds the true neighbours its
// output is bit-identical to this one.  oracle/_ref/libflann_ref.so (built from the vendored
// FLANN sources by oracle/Makefile when /root/reference is present) provides the real
// LinearIndex and the real KD-tree forest through the `search` callbacks below;
// tests/test_match_oracle.py pins the restated search against the former bit for bit and
// measures the recall of the latter.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

extern "C" {

struct OMKeypoint  // == sara_b200_keypoint == the SIFT oracle's Keypoint (52 bytes)
{
  float x, y;
  float shape[4];  // column-major
  float orientation;
  float extremum_value;
  std::uint8_t type;
  std::int8_t extremum_type;
  std::int16_t pad_;
  std::int32_t s, o, xi, yi;
};

struct OMatch
{
  std::int32_t x_index, y_index;  // Match::x_index(), y_index(): keys1 / keys2 indices
  std::int32_t rank;              // Match::rank()
  float score;                    // Match::score(): ratio of SQUARED distances
  std::int32_t direction;         // Match::Direction: 0 SourceToTarget, 1 TargetToSource
};

// k-NN / radius search callbacks (nullptr: the restated linear search).
typedef void (*om_knn_fn)(void* index, const float* query, int k, int* idx, float* dist);
typedef int (*om_radius_fn)(void* index, const float* query, float radius, int* idx, float* dist, int cap);

// flann::L2<float> (dist.h:151-178), worst_dist = -1: four squared differences are summed left
// to right and THEN added to the running result.
float oracle_l2_flann(const float* a, const float* b, int size)
{
  float result = 0.f;
  int i = 0;
  for (; i + 3 < size; i += 4)
  {
    const float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
    result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (; i < size; ++i)
  {
    const float d0 = a[i] - b[i];
    result += d0 * d0;
  }
  return result;
}

// LinearIndex::findNeighbors into a KNNSimpleResultSet of capacity k: idx / dist get k entries,
// unused ones are (-1, FLT_MAX) as the result set initialises them.
void oracle_knn_linear(const float* data, int n, int dim, const float* queries, int nq, int k, int* idx, float* dist)
{
#pragma omp parallel for schedule(dynamic, 16)
  for (int q = 0; q < nq; ++q)
  {
    int* I = idx + static_cast<size_t>(q) * k;
    float* D = dist + static_cast<size_t>(q) * k;
    for (int j = 0; j < k; ++j)
    {
      I[j] = -1;
      D[j] = std::numeric_limits<float>::max();
    }
    int count = 0;
    float worst = std::numeric_limits<float>::max();
    for (int i = 0; i < n; ++i)
    {
      const float d = oracle_l2_flann(data + static_cast<size_t>(i) * dim, queries + static_cast<size_t>(q) * dim, dim);
      if (d >= worst)
        continue;
      if (count < k)
        ++count;
      int j = count - 1;
      for (; j > 0; --j)
      {
        if (D[j - 1] > d)
        {
          D[j] = D[j - 1];
          I[j] = I[j - 1];
        }
        else
          break;
      }
      D[j] = d;
      I[j] = i;
      worst = D[k - 1];
    }
  }
}

// RadiusResultSet + copy(sorted = true): every point with dist < radius, ordered by (dist, index).
int oracle_radius_linear(const float* data, int n, int dim, const float* query, float radius, int* idx, float* dist,
                         int cap)
{
  std::vector<std::pair<float, int>> found;
  for (int i = 0; i < n; ++i)
  {
    const float d = oracle_l2_flann(data + static_cast<size_t>(i) * dim, query, dim);
    if (d < radius)
      found.emplace_back(d, i);
  }
  std::sort(found.begin(), found.end());
  const int m = std::min<int>(static_cast<int>(found.size()), cap);
  for (int i = 0; i < m; ++i)
  {
    dist[i] = found[i].first;
    idx[i] = found[i].second;
  }
  return static_cast<int>(found.size());
}

}  // extern "C"

namespace {

  struct Side
  {
    const float* desc;
    const OMKeypoint* feat;  // may be nullptr (index equality, no proximity test)
    int n;
    void* index;  // handle for the callbacks
  };

  // Metric.hpp:46-49 with Eigen's evaluation order for 2x2 float: M * d first, then the dot.
  float squared_ref_distance(const float* M /*col-major*/, float ax, float ay, float bx, float by)
  {
    const float dx = bx - ax, dy = by - ay;
    const float mx = M[0] * dx + M[2] * dy;
    const float my = M[1] * dx + M[3] * dy;
    return dx * mx + dy * my;
  }

  // KeyProximity.cpp:17-30
  bool is_redundant(const OMKeypoint& f1, const OMKeypoint& f2, float sq_metric, float sq_pixel)
  {
    const float sd1 = squared_ref_distance(f1.shape, f1.x, f1.y, f2.x, f2.y);
    const float sd2 = squared_ref_distance(f2.shape, f1.x, f1.y, f2.x, f2.y);
    const float dx = f1.x - f2.x, dy = f1.y - f2.y;
    const float pix = dx * dx + dy * dy;
    return pix < sq_pixel || sd1 < sq_metric || sd2 < sq_metric;
  }

  bool same_feature(const OMKeypoint& a, const OMKeypoint& b)  // Feature.hpp:140-146
  {
    return a.x == b.x && a.y == b.y && a.shape[0] == b.shape[0] && a.shape[1] == b.shape[1] &&
           a.shape[2] == b.shape[2] && a.shape[3] == b.shape[3] && a.orientation == b.orientation && a.type == b.type;
  }

  struct Search
  {
    om_knn_fn knn;
    om_radius_fn radius;
    int dim;
  };

  void knn3(const Search& S, const Side& tree, const float* query, int k, int* idx, float* dist)
  {
    if (S.knn)
      S.knn(tree.index, query, k, idx, dist);
    else
      oracle_knn_linear(tree.desc, tree.n, S.dim, query, 1, k, idx, dist);
  }

  // AnnMatcher.cpp:57-171.  `one` are the querying keys, `two` the indexed ones.
  void append_nearest_neighbors(int i1, const Side& one, const Side& two, std::vector<OMatch>& matches, const Search& S,
                                float squared_ratio_thres, int dir, bool self_matching, float sq_metric, float sq_pixel,
                                std::vector<int>& vec_indices, std::vector<float>& vec_dists)
  {
    const float* query = one.desc + static_cast<size_t>(i1) * S.dim;
    auto push = [&](int i2, float score, int rank) {
      OMatch m;
      m.x_index = i1;
      m.y_index = i2;
      if (dir == 1)
        std::swap(m.x_index, m.y_index);
      m.rank = rank;
      m.score = score;
      m.direction = dir;
      matches.push_back(m);
    };

    if (two.n == 0)  // boundary case 1
      return;
    if (two.n == 1 && !self_matching)  // boundary case 2
    {
      if (1.f < squared_ratio_thres)
        push(0, 1.f, 1);
      return;
    }
    int* indices = vec_indices.data();
    float* dists = vec_dists.data();
    if (two.n == 2 && self_matching)  // boundary case 3
    {
      knn3(S, two, query, 2, indices, dists);
      if (1.f < squared_ratio_thres)
        push(indices[1], 1.f, 1);
      return;
    }

    knn3(S, two, query, 3, indices, dists);
    const int top1_index = self_matching ? 1 : 0;
    const float top1_score = dists[top1_index + 1] > 0.f ? dists[top1_index] / dists[top1_index + 1] : 0.f;
    int K = 1;
    if (squared_ratio_thres > 1.f)
    {
      const float radius = dists[top1_index] * squared_ratio_thres;
      const int cap = static_cast<int>(vec_indices.size());
      K = S.radius ? S.radius(two.index, query, radius, indices, dists, cap)
                   : oracle_radius_linear(two.desc, two.n, S.dim, query, radius, indices, dists, cap);
      K = std::min(K, cap);
    }
    for (int rank = top1_index; rank < K; ++rank)
    {
      float score = 0.f;
      if (rank == top1_index)
        score = top1_score;
      else if (dists[top1_index])
        score = dists[rank] / dists[top1_index];
      if (score > squared_ratio_thres)
        break;
      const int i2 = indices[rank];
      if (self_matching && one.feat && is_redundant(one.feat[i1], two.feat[i2], sq_metric, sq_pixel))
        continue;
      push(i2, score, top1_index == 0 ? rank + 1 : rank);
    }
  }

}  // namespace

extern "C" {

// AnnMatcher{keys1, keys2, sift_ratio_thres}.compute_matches()  (self_matching = 0), or
// AnnMatcher{keys, sift_ratio_thres, min_max_metric_dist_thres, pixel_dist_thres}.compute_matches()
// (self_matching = 1: pass the same arrays twice).  Returns the number of matches (all of them are
// counted, `cap` of them are written).  The final order is by score; std::sort leaves the order
// of equal scores unspecified, here (and in the GPU library) ties keep the (x_index, y_index)
// order of the preceding lexicographic sort.
int oracle_ann_match(const float* desc1, const OMKeypoint* feat1, int n1, const float* desc2, const OMKeypoint* feat2,
                     int n2, int dim, float sift_ratio_thres, int self_matching, float min_max_metric_dist_thres,
                     float pixel_dist_thres, om_knn_fn knn, om_radius_fn radius, void* index1, void* index2,
                     OMatch* out, int cap)
{
  if (n1 == 0 || n2 == 0)  // create_flann_matrix throws "the list of key-points is empty"
    return -1;
  const float squared_ratio_thres = sift_ratio_thres * sift_ratio_thres;
  const float sq_metric = min_max_metric_dist_thres * min_max_metric_dist_thres;
  const float sq_pixel = pixel_dist_thres * pixel_dist_thres;
  const Side one{desc1, feat1, n1, index1}, two{desc2, feat2, n2, index2};
  const Search S{knn, radius, dim};
  const size_t max_neighbors = static_cast<size_t>(std::max(std::max(n1, n2), 3));
  std::vector<int> vec_indices(max_neighbors);
  std::vector<float> vec_dists(max_neighbors);

  std::vector<OMatch> matches;
  matches.reserve(100000);
  for (int i1 = 0; i1 < n1; ++i1)
    append_nearest_neighbors(i1, one, two, matches, S, squared_ratio_thres, 0, self_matching != 0, sq_metric, sq_pixel,
                             vec_indices, vec_dists);
  for (int i2 = 0; i2 < n2; ++i2)
    append_nearest_neighbors(i2, two, one, matches, S, squared_ratio_thres, 1, self_matching != 0, sq_metric, sq_pixel,
                             vec_indices, vec_dists);

  std::sort(matches.begin(), matches.end(), [](const OMatch& a, const OMatch& b) {
    if (a.x_index != b.x_index)
      return a.x_index < b.x_index;
    if (a.y_index != b.y_index)
      return a.y_index < b.y_index;
    return a.score < b.score;
  });
  auto equal = [&](const OMatch& a, const OMatch& b) {
    if (feat1 && feat2)
      return same_feature(feat1[a.x_index], feat1[b.x_index]) && same_feature(feat2[a.y_index], feat2[b.y_index]);
    return a.x_index == b.x_index && a.y_index == b.y_index;
  };
  matches.resize(std::unique(matches.begin(), matches.end(), equal) - matches.begin());
  std::stable_sort(matches.begin(), matches.end(), [](const OMatch& a, const OMatch& b) { return a.score < b.score; });

  const int n = static_cast<int>(matches.size());
  for (int i = 0; i < std::min(n, cap); ++i)
    out[i] = matches[i];
  return n;
}

}  // extern "C"
