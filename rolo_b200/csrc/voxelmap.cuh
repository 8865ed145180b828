// voxelmap.cuh — launch interface of the voxel-map build.
#pragma once
#include "common.cuh"

namespace rolo {

struct VoxelBuild {
  const float4* pts;       // target points
  const float* cov6;       // target covariances (tile-planar)
  int n;
  GridParams grid;
  VoxSlot* table;          // [mask+1]
  unsigned mask;
  int* slot_of_point;      // scratch [n]
  int* counters;           // [2]: voxel count, error flag
  double* sums;            // scratch [n][10]
  VoxPayload* vox;         // out [<= n]
};

cudaError_t launch_voxel_build(const VoxelBuild& b, cudaStream_t stream);                                 // 4 kernels
// every target cloud of a batch by the same 4 kernels on a (blocks, build) grid; dev_builds = device scratch [n_builds]
cudaError_t launch_voxel_build_batch(const VoxelBuild* host_builds, VoxelBuild* dev_builds, int n_builds, cudaStream_t stream);

}  // namespace rolo
