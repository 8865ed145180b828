// vmp_voxel.hpp — kept so `#include <rot_gicp/gicp/vmp_voxel.hpp>` keeps resolving (reference header of the same path).
//
// In the B200 build the voxel map itself (the reference's VmfVoxelMap / AdditiveVmfVoxel: hashed accumulation of point and covariance
// sums per cell, vmp_voxel.hpp:93-108,157-243) lives on the device behind the C-ABI: rg_create_voxelmap builds it, rg_get_voxels reads it
// back (keys, means, covariances, counts), the fused linearize kernel probes it. What callers can still name here are the pure
// functions of the header: the neighbour offset lists and the two cell-coordinate rules, with the reference's semantics —
//   neighbor_offsets(method)           vmp_voxel.hpp:13-47 (same order: the ABI's correspondence arrays are indexed by it)
//   uniform_voxel_coord(x, res)        floor(x / res - 0.5) per axis                        vmp_voxel.hpp:199-201
//   polar_voxel_coord(x, res)          floor([atan2(y,x)+pi, acos(z/r), r] ./ res)          vmp_voxel.hpp:208-211
#ifndef ROLO_B200_VMP_VOXEL_HPP
#define ROLO_B200_VMP_VOXEL_HPP

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <vector>
#include <Eigen/Core>
#include <rot_gicp/gicp/gicp_settings.hpp>

namespace fast_gicp {

inline std::vector<Eigen::Vector3i, Eigen::aligned_allocator<Eigen::Vector3i>> neighbor_offsets(NeighborSearchMethod search_method) {
  std::vector<Eigen::Vector3i, Eigen::aligned_allocator<Eigen::Vector3i>> offsets;
  if (search_method == NeighborSearchMethod::DIRECT1) {
    offsets.push_back(Eigen::Vector3i(0, 0, 0));
  } else if (search_method == NeighborSearchMethod::DIRECT7) {
    const int d[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (const auto& o : d) offsets.push_back(Eigen::Vector3i(o[0], o[1], o[2]));
  } else if (search_method == NeighborSearchMethod::DIRECT27) {
    for (int i = -1; i <= 1; i++) for (int j = -1; j <= 1; j++) for (int k = -1; k <= 1; k++) offsets.push_back(Eigen::Vector3i(i, j, k));
  } else {
    std::cerr << "unsupported neighbor search method" << std::endl;   // the reference abort()s here too (vmp_voxel.hpp:16-18)
    abort();
  }
  return offsets;
}

inline Eigen::Vector3i uniform_voxel_coord(const Eigen::Vector4d& x, double resolution) {
  return Eigen::Vector3i((int)std::floor(x[0] / resolution - 0.5), (int)std::floor(x[1] / resolution - 0.5), (int)std::floor(x[2] / resolution - 0.5));
}

inline Eigen::Vector3i polar_voxel_coord(const Eigen::Vector4d& x, const Eigen::Vector3d& polar_resolution) {
  const double r = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const double theta = std::atan2(x[1], x[0]) + M_PI, phi = std::acos(x[2] / r);
  return Eigen::Vector3i((int)std::floor(theta / polar_resolution[0]), (int)std::floor(phi / polar_resolution[1]), (int)std::floor(r / polar_resolution[2]));
}

}  // namespace fast_gicp

#endif  // ROLO_B200_VMP_VOXEL_HPP
