// covariance.cuh — launch interface of the exact-kNN covariance estimation.
#pragma once
#include "common.cuh"

namespace rolo {

struct CovScratch {           // all device pointers; sized for n points (see cov_scratch_bytes)
  unsigned long long* codes_in;   // [n] unsorted codes; after the sort: packed cell coordinates of the sorted points
  unsigned long long* codes_out;  // [n]
  int* idx_in;                    // [n]
  int* idx_out;                   // [n]
  float4* sorted;                 // [n] xyz + original index (int bits in w)
  void* cub_temp;
  size_t cub_temp_bytes;
  int* bbox;                      // [6] ordered-int min xyz / max xyz
  int* bucket_start;              // [cov_bucket_table_entries()] first sorted position per top-18-bit Morton bucket
};

size_t cov_cub_temp_bytes(int n);
size_t cov_bucket_table_entries();

struct CovLaunch {
  const float4* pts;          // [n] (w ignored)
  int n;
  int k;                      // neighbours incl. self (k_correspondences_)
  int method;                 // RG_REG_*
  float* cov6;                // out, tile-planar (may be null when only knn_out is wanted)
  int* knn;                   // [n*k] neighbour indices (original numbering) sorted by (distance, index): scratch + test output
  CovScratch s;
};

cudaError_t launch_covariances(const CovLaunch& L, cudaStream_t stream);
// spatial index only (bbox, Morton codes, sort, Morton-ordered copy, optional bucket table)
// exact nearest indexed point (original numbering) + float squared distance of every query (optionally moved by the device-resident
// 3x4 row-major float transform T12_dev first); `s` must hold the index launch_sort_cloud(…, bucket_table = true) built over n points
cudaError_t launch_nearest(const float4* queries, int nq, const float* T12_dev, const CovScratch& s, int n, int* out_idx, float* out_d2, cudaStream_t stream);
cudaError_t launch_sort_cloud(const float4* pts, int n, const CovScratch& s, bool bucket_table, cudaStream_t stream);

}  // namespace rolo
