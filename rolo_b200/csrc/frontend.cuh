// frontend.cuh — launch interface of the range-image projection and LOAM feature-extraction kernels.
#pragma once
#include "common.cuh"

namespace rolo {

constexpr int FE_MAX_CORNERS_PER_RING = 120;   // 6 sectors x 20 (featureExtraction.cpp:189)
constexpr int FE_SECTOR_CAP = 2048;            // sortable sector length (horizon <= 12288)
constexpr int FE_RING_CAP = 2048;              // surface candidates per ring (horizon <= 2048)

struct ProjDeskew {             // deskewPoint (imageProjection.cpp:368-396), rotation only; enabled = deskewEnabled && odomAvailable
  int enabled, time_offset;     // byte offset of the point's float relative time
  float incre_rpy[3];           // odomIncreRoll / Pitch / Yaw
  float scan_period;
  double odom_time_diff;
};

struct ProjLaunch {
  const unsigned char* raw;   // device copy of the raw points
  int n_points, stride, ring_offset;
  int n_scan, horizon, downsample;
  float min_range, max_range;
  ProjDeskew deskew;
  int* owner;                 // [n_scan*horizon] scratch: smallest point index claiming the cell
  float* range_image;         // [n_scan*horizon]
  float4* full;               // [n_scan*horizon] xyz + intensity
  int* row_count;             // [n_scan]
  float4* extracted;          // [n_scan*horizon]
  int* col_ind;               // [n_scan*horizon]
  float* point_range;         // [n_scan*horizon]
  int* start_ring;            // [n_scan]
  int* end_ring;              // [n_scan]
  int* n_extracted;           // [1]
};
cudaError_t launch_projection(const ProjLaunch& L, cudaStream_t stream);

struct FeatLaunch {
  const float4* extracted; const int* col_ind; const float* point_range; const int* start_ring; const int* end_ring;
  const int* n_dev;           // device count of extracted points (may be null: use n)
  int n;                      // host-known count (upper bound when n_dev != null)
  int n_scan, horizon;
  float edge_thr, surf_thr, leaf;
  float* curvature; int* picked; int* label;     // [n]
  float4* corner_ring; int* corner_cnt;          // [n_scan*120], [n_scan]
  float4* surf_ring; int* surf_cnt;              // [n_scan*horizon], [n_scan]
  float4* corner_out; int* n_corner;             // [n_scan*120], [1]
  float4* surf_out; int* n_surf;                 // [n_scan*horizon], [1]
  int* error_flag;                               // [1] set when a capacity limit is exceeded
};
cudaError_t launch_features(const FeatLaunch& L, cudaStream_t stream);

}  // namespace rolo
