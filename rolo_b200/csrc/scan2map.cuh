// scan2map.cuh — launch interface of the scan-to-map correspondence + normal-equation kernel (backMapping.cpp:719-1006).
#pragma once
#include "common.cuh"

namespace rolo {

constexpr int S2M_RES = 32;        // doubles per result: 21 AtA (upper triangle, row-major) + 6 AtB + 1 count + pad
constexpr int S2M_FIT_THREADS_PER_BLOCK = 128;   // partial[] needs ceil(n_scan_total / 128) x S2M_RES doubles

struct S2MMap {                    // one Morton-indexed map cloud (covariance.cuh: launch_sort_cloud)
  const float4* pts;               // original order
  const float4* sorted;            // Morton order, w = original index
  const unsigned long long* codes; // sorted codes
  const int* bucket_start;
  const int* bbox;
  int n;
};

struct S2MLaunch {
  S2MMap map[2];                   // 0 = laserCloudCornerFromMapDS, 1 = laserCloudSurfFromMapDS
  const float4* scan[2];           // laserCloudCornerLastDS, laserCloudSurfLastDS
  int n_scan[2];
  float t[12];                     // transPointAssociateToMap, rows (host: s2m::get_transformation)
  float sc[6];                     // srx crx sry cry srz crz (LMOptimization, backMapping.cpp:949-954; host sinf/cosf)
  float4* coeff[2];                // coeffSel per scan point (x y z intensity), zero when not selected
  unsigned char* flag[2];          // laserCloudOri*Flag
  int* nn;                         // [(n_scan[0] + n_scan[1]) x 5] neighbour indices (first = -1: rejected)
  double* partial;                 // [ceil(n_scan_total / 128)][S2M_RES]
  unsigned* counter;               // ticket, zero before the first launch (the kernel re-arms it)
  double* result;                  // [S2M_RES]
};

cudaError_t launch_scan2map(const S2MLaunch& L, cudaStream_t stream);

}  // namespace rolo
