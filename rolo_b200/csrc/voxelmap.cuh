// voxelmap.cuh — launch interface of the voxel-map build.
#pragma once
#include "common.cuh"
#include "covariance.cuh"

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
  // optional restriction to one rank's share of a Morton-ordered copy of the target (xyz + original index in w): only the points at the
  // sorted positions the rank owns (covariance.cuh: knn_owns) are inserted / accumulated — a rank owns the points whose covariances it computed
  const float4* order = nullptr;
  int shard_rank = 0, shard_world = 1;
};

cudaError_t launch_voxel_build(const VoxelBuild& b, cudaStream_t stream, bool finalize = true);           // 4 kernels (3 without the finalize pass)
// Deterministic build: points are radix-sorted by packed voxel key (stable: index order inside a voxel), one thread per
// voxel adds its points in index order — the summation order of the reference's serial loop — and voxel ids are the ranks of the keys.
// Bit-reproducible from run to run (the default build sums with fp64 atomics in arrival order). Scratch: keys_a/keys_b [n] u64, idx_a/idx_b [n]
// int, flags [n + 1] int, cub temp (cov_cub_temp_bytes(n) is enough). Whole-cloud builds only.
struct VoxelDetScratch { unsigned long long *keys_a, *keys_b; int *idx_a, *idx_b, *flags; void* cub_temp; size_t cub_temp_bytes; };
size_t voxel_det_temp_bytes(int n);
cudaError_t launch_voxel_build_deterministic(const VoxelBuild& b, const VoxelDetScratch& s, cudaStream_t stream);
// compact per-voxel key list of a built map: keys[id] = packed key (rows of b.sums are already compact by id)
cudaError_t launch_voxel_keys(const VoxelBuild& b, unsigned long long* keys, cudaStream_t stream);
// rebuild the map from a list of (key, 10 sums) entries gathered over all ranks: entries with equal keys add up; then finalize
cudaError_t launch_voxel_import(const VoxelBuild& b, const unsigned long long* keys, const double* sums_in, int n_entries, cudaStream_t stream);
// every target cloud of a batch by the same 4 kernels on a (blocks, build) grid; dev_builds = device scratch [n_builds]
cudaError_t launch_voxel_build_batch(const VoxelBuild* host_builds, VoxelBuild* dev_builds, int n_builds, cudaStream_t stream);

}  // namespace rolo
