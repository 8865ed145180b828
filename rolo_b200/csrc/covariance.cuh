// covariance.cuh — launch interface of the exact-kNN covariance estimation and of the Morton-ordered spatial index.
#pragma once
#include "common.cuh"

namespace rolo {

struct CovScratch {           // single-cloud index (scan-to-map maps, fitness score, k > 32 fallback); device pointers sized for n points
  unsigned long long* codes_in;   // [n] unsorted codes (scratch)
  unsigned long long* codes_out;  // [n] sorted 39-bit Morton codes
  int* idx_in;                    // [n]
  int* idx_out;                   // [n]
  float4* sorted;                 // [n] xyz + original index (int bits in w)
  void* cub_temp;
  size_t cub_temp_bytes;
  int* bbox;                      // [6] ordered-int min xyz / max xyz
  int* bucket_start;              // [cov_bucket_table_entries()] first sorted position per top-bits Morton bucket
};

size_t cov_cub_temp_bytes(int n, int n_clouds = 1);
size_t cov_bucket_table_entries();

// ---- multi-cloud covariance estimation (the product path of calculate_covariances, k <= 32) ------------------------------
// A GROUP of clouds (both clouds of many scan pairs) is indexed and searched by one chain of launches: the clouds are
// concatenated in the scratch arrays (cloud c occupies [off, off + n)), sorted by ONE radix sort over (cloud id | Morton code),
// and every kernel runs on a (blocks, cloud) grid.
struct KnnCloud {
  const float4* pts;          // [n] points in the caller's order (w ignored)
  float* cov6;                // out: tile-planar regularised covariances in the caller's order (may be null)
  int* knn_out;               // out, optional: [n*k] the k nearest neighbours (original numbering, UNORDERED set)
  int n;
  int off;                    // first element of this cloud in the concatenated scratch arrays
  int chunk_off;              // number of 32-query chunks in the clouds before this one
  int shard_rank, shard_world; // queries = the sorted positions of the blocks (KNN_SHARD_BLOCK points) with index = rank mod world; world <= 1: all.
                              // A target-sharded registration deals the Morton order out in blocks, round-robin over the ranks (the expensive
                              // queries cluster in space, so contiguous slices would be badly balanced); the index always covers the whole cloud
};

struct KnnScratch {           // device pointers; `total` = sum of n over the group
  unsigned long long* codes_a;    // [total] unsorted keys; after the index build: the sorted 39-bit codes (cloud bits stripped)
  unsigned long long* codes_b;    // [total] sort output (cloud id << 39 | code)
  int* idx_a;                     // [total]
  int* idx_b;                     // [total]
  float4* sorted;                 // [total] Morton-ordered xyz + original index (int bits in w)
  void* cub_temp;
  size_t cub_temp_bytes;
  int* bbox;                      // [n_clouds * 8]
  int* bucket_start;              // [n_clouds * cov_bucket_table_entries()]
  int* fail;                      // [2 + 2 * total]: fail[0] = number of queries handed to the fallback kernel, fail[1] = chunk ticket, then (cloud, sorted position) pairs
  KnnCloud* clouds;               // [n_clouds] device copy of the descriptors
};

struct KnnGroup {
  const KnnCloud* clouds;     // host array [n_clouds]
  int n_clouds;
  int total;                  // sum of n
  int max_n;
  int k;                      // neighbours incl. self (k_correspondences_), 1 <= k <= 32
  int method;                 // RG_REG_*
  KnnScratch s;
  unsigned long long* stats = nullptr;   // optional device counters [8] (development: ROLO_KNN_STATS=1)
  cudaEvent_t ev_knn0 = nullptr, ev_knn1 = nullptr;   // optional: recorded around the search kernels (chunk + fallback) for stage timing
};
constexpr int KNN_MAX_CLOUDS = 4096;
constexpr int KNN_SHARD_BLOCK = 2048;   // points per dealt block of a sharded query set (64 chunks of 32)
// number of 32-query chunks rank `rank` of `world` owns in a cloud of n points
inline int knn_owned_chunks(int n, int rank, int world) {
  const int chunks = (n + 31) / 32;
  if (world <= 1) return chunks;
  const int per = KNN_SHARD_BLOCK / 32, blocks = (chunks + per - 1) / per;
  int own = 0;
  for (int b = rank; b < blocks; b += world) own += (b == blocks - 1) ? chunks - b * per : per;
  return own;
}
__host__ __device__ inline bool knn_owns(int sorted_pos, int rank, int world) { return world <= 1 || (sorted_pos / KNN_SHARD_BLOCK) % world == rank; }

// index build + chunk-cooperative exact kNN + fused covariance/regulariser epilogue + fallback pass. `own_launches` receives
// the number of this library's kernels that were launched (cub's are not counted).
cudaError_t launch_covariances_group(const KnnGroup& g, cudaStream_t stream, int* own_launches);

// ---- generic single-cloud path (32 < k <= 64): thread-per-query kNN into `knn`, then covariance from the lists -----------
struct CovLaunch {
  const float4* pts;          // [n] (w ignored)
  int n;
  int k;
  int method;                 // RG_REG_*
  float* cov6;                // out, tile-planar (may be null when only knn is wanted)
  int* knn;                   // [n*k] neighbour indices (original numbering) sorted by (distance, index)
  CovScratch s;
};
cudaError_t launch_covariances_generic(const CovLaunch& L, cudaStream_t stream);

// exact nearest indexed point (original numbering) + float squared distance of every query (optionally moved by the device-resident
// 3x4 row-major float transform T12_dev first); `s` must hold the index launch_sort_cloud(…, bucket_table = true) built over n points
cudaError_t launch_nearest(const float4* queries, int nq, const float* T12_dev, const CovScratch& s, int n, int* out_idx, float* out_d2, cudaStream_t stream);
// spatial index only (bbox, Morton codes, sort, Morton-ordered copy, optional bucket table)
cudaError_t launch_sort_cloud(const float4* pts, int n, const CovScratch& s, bool bucket_table, cudaStream_t stream);

}  // namespace rolo
