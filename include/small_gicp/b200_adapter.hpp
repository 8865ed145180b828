// small_gicp/b200_adapter.hpp — a small_gicp-SHAPED surface over the B200 voxelized-GICP core (include/rolo_gicp.h).
//
// The reference's back end includes six small_gicp headers (src/backMapping.cpp:5-10) and carries, commented out, the VGICP prior-patch
// registration they were meant for (backMapping.cpp:1993-2028):
//     small_gicp::PointCloud target_points / source_points;   resize(n); point(i) << x, y, z, 1.0;
//     small_gicp::estimate_covariances_omp(cloud, 10, numberOfCores);
//     small_gicp::GaussianVoxelMap target_voxelmap(leaf);     target_voxelmap.insert(target_points);
//     small_gicp::Registration<small_gicp::GICPFactor, small_gicp::ParallelReductionOMP> vgicp;
//     vgicp.reduction.num_threads = ...; vgicp.rejector.max_dist_sq = 1.0; vgicp.optimizer.max_iterations = 50;
//     small_gicp::RegistrationResult r = vgicp.align(target_voxelmap, source_points, target_voxelmap, Eigen::Isometry3d::Identity());
//     r.T_target_source, r.error
// small_gicp itself (koide3/small_gicp) is NOT in the reference tree and no version is pinned (CMakeLists.txt:70-74): there is no
// numerical parity target. These headers give that code something to compile against and run it on the GPU core, with this mapping:
//     estimate_covariances_omp(cloud, k, threads)  -> rg_calculate_covariances with k neighbours, PLANE regulariser (eigenvalues 1e-3, 1, 1:
//                                                     the same regulariser small_gicp's covariance estimation applies); threads ignored
//     GaussianVoxelMap(leaf) + insert(points)      -> UNIFORM voxel map of that resolution over the inserted points (rg_set_resolution);
//                                                     cell = floor(x / leaf - 0.5) as in rot_gicp, one voxel per query (DIRECT1)
//     Registration<GICPFactor, *>::align           -> 6-DoF Levenberg-Marquardt of the rot_gicp cost (sum sqrt(n) e^T (C_B + R C_A R^T)^-1 e)
//                                                     from `init`; optimizer.max_iterations -> max iterations; rejector.max_dist_sq is
//                                                     accepted and ignored (a source point without a voxel in its cell has no correspondence)
//     RegistrationResult                           -> T_target_source, converged, iterations, num_inliers (correspondences), H, b, error
#ifndef ROLO_B200_SMALL_GICP_ADAPTER_HPP
#define ROLO_B200_SMALL_GICP_ADAPTER_HPP

#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>

#include <rolo_gicp.h>

namespace small_gicp {

struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud>;
  using ConstPtr = std::shared_ptr<const PointCloud>;
  std::vector<Eigen::Vector4d, Eigen::aligned_allocator<Eigen::Vector4d>> points;
  std::vector<Eigen::Vector4d, Eigen::aligned_allocator<Eigen::Vector4d>> normals;
  std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> covs;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(size_t n) { points.resize(n); normals.resize(n); covs.resize(n); }
  Eigen::Vector4d& point(size_t i) { return points[i]; }
  const Eigen::Vector4d& point(size_t i) const { return points[i]; }
  Eigen::Matrix4d& cov(size_t i) { return covs[i]; }
  const Eigen::Matrix4d& cov(size_t i) const { return covs[i]; }
};

namespace detail {
inline void check(int rc, rg_ctx* c, const char* what) {
  if (rc == RG_OK || rc == RG_ERR_NOT_CONVERGED) return;
  const std::string msg = std::string("small_gicp(b200)::") + what + ": " + rg_last_error(c);
  if (rc == RG_ERR_INVALID) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}
inline std::vector<float> xyzw(const PointCloud& c) {
  std::vector<float> p(c.size() * 4);
  for (size_t i = 0; i < c.size(); i++) { p[i * 4] = (float)c.points[i][0]; p[i * 4 + 1] = (float)c.points[i][1]; p[i * 4 + 2] = (float)c.points[i][2]; p[i * 4 + 3] = 1.f; }
  return p;
}
inline std::vector<double> row_major(const PointCloud& c) {
  std::vector<double> m(c.size() * 16);
  for (size_t k = 0; k < c.size(); k++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[k * 16 + i * 4 + j] = c.covs[k](i, j);
  return m;
}
struct Ctx {
  rg_ctx* c = nullptr;
  Ctx() { check(rg_create(-1, &c), nullptr, "rg_create"); }
  ~Ctx() { rg_destroy(c); }
  Ctx(const Ctx&) = delete; Ctx& operator=(const Ctx&) = delete;
};
}  // namespace detail

// covariances of every point from its k nearest neighbours (itself included), regularised to eigenvalues (1e-3, 1, 1)
inline void estimate_covariances_omp(PointCloud& cloud, int num_neighbors = 20, int /*num_threads*/ = 4) {
  detail::Ctx x;
  const std::vector<float> p = detail::xyzw(cloud);
  detail::check(rg_set_correspondence_randomness(x.c, num_neighbors), x.c, "estimate_covariances");
  detail::check(rg_set_regularization_method(x.c, RG_REG_PLANE), x.c, "estimate_covariances");
  detail::check(rg_set_input_source(x.c, p.data(), cloud.size(), 16), x.c, "estimate_covariances");
  detail::check(rg_calculate_covariances(x.c, RG_SOURCE), x.c, "estimate_covariances");
  std::vector<double> m(cloud.size() * 16);
  if (rg_get_covariances(x.c, RG_SOURCE, m.data()) != cloud.size()) throw std::runtime_error("small_gicp(b200)::estimate_covariances: read-back failed");
  cloud.covs.resize(cloud.size());
  for (size_t k = 0; k < cloud.size(); k++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) cloud.covs[k](i, j) = m[k * 16 + i * 4 + j];
}
inline void estimate_covariances(PointCloud& cloud, int num_neighbors = 20) { estimate_covariances_omp(cloud, num_neighbors, 1); }

// the target side of a VGICP registration: leaf size + the inserted points with their covariances (the voxel map itself is built on the device)
struct GaussianVoxelMap {
  using Ptr = std::shared_ptr<GaussianVoxelMap>;
  using ConstPtr = std::shared_ptr<const GaussianVoxelMap>;
  explicit GaussianVoxelMap(double leaf_size) : leaf_size(leaf_size) {}
  void insert(const PointCloud& points) {
    cloud.points.insert(cloud.points.end(), points.points.begin(), points.points.end());
    cloud.covs.insert(cloud.covs.end(), points.covs.begin(), points.covs.end());
  }
  size_t size() const { return cloud.size(); }
  double leaf_size;
  PointCloud cloud;
};

struct GICPFactor {};
struct ParallelReductionOMP { int num_threads = 4; };
struct SerialReduction {};
struct DistanceRejector { double max_dist_sq = 1.0; };
struct LevenbergMarquardtOptimizer { bool verbose = false; int max_iterations = 20; int max_inner_iterations = 10; double init_lambda = 1e-3; double lambda_factor = 10.0; };
struct TerminationCriteria { double translation_eps = 1e-3; double rotation_eps = 0.1 * M_PI / 180.0; };

struct RegistrationResult {
  Eigen::Isometry3d T_target_source = Eigen::Isometry3d::Identity();
  bool converged = false;
  size_t iterations = 0, num_inliers = 0;
  Eigen::Matrix<double, 6, 6> H;
  Eigen::Matrix<double, 6, 1> b;
  double error = 0.0;
};

template <typename PointFactor = GICPFactor, typename Reduction = ParallelReductionOMP>
struct Registration {
  TerminationCriteria criteria;
  DistanceRejector rejector;
  PointFactor point_factor;
  Reduction reduction;
  LevenbergMarquardtOptimizer optimizer;

  // `target_tree` is the same voxel map in the reference's call (backMapping.cpp:2024-2025)
  RegistrationResult align(const GaussianVoxelMap& target, const PointCloud& source, const GaussianVoxelMap& /*target_tree*/,
                           const Eigen::Isometry3d& init_T = Eigen::Isometry3d::Identity()) const {
    detail::Ctx x;
    rg_ctx* c = x.c;
    const std::vector<float> tp = detail::xyzw(target.cloud), sp = detail::xyzw(source);
    detail::check(rg_set_resolution(c, target.leaf_size), c, "align");
    detail::check(rg_set_neighbor_search_method(c, RG_DIRECT1), c, "align");
    detail::check(rg_set_optimizer_type(c, RG_OPT_LM), c, "align");
    detail::check(rg_set_maximum_iterations(c, optimizer.max_iterations), c, "align");
    detail::check(rg_set_rotation_epsilon(c, criteria.rotation_eps), c, "align");
    detail::check(rg_set_transformation_epsilon(c, criteria.translation_eps), c, "align");
    detail::check(rg_set_input_target(c, tp.data(), target.cloud.size(), 16), c, "align");
    detail::check(rg_set_input_source(c, sp.data(), source.size(), 16), c, "align");
    if (target.cloud.covs.size() == target.cloud.size() && !target.cloud.covs.empty()) {
      const std::vector<double> m = detail::row_major(target.cloud);
      detail::check(rg_set_covariances(c, RG_TARGET, m.data(), target.cloud.size()), c, "align");
    }
    if (source.covs.size() == source.size() && !source.covs.empty()) {
      const std::vector<double> m = detail::row_major(source);
      detail::check(rg_set_covariances(c, RG_SOURCE, m.data(), source.size()), c, "align");
    }
    float guess[16], out[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) guess[i * 4 + j] = (float)init_T.matrix()(i, j);
    detail::check(rg_align(c, guess, out), c, "align");
    RegistrationResult r;
    double T[16], H[36];
    rg_get_final_pose(c, T); rg_get_final_hessian(c, H);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.T_target_source.matrix()(i, j) = T[i * 4 + j];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) r.H(i, j) = H[i * 6 + j];
    r.converged = rg_has_converged(c) != 0;
    r.iterations = (size_t)rg_nr_iterations(c) + 1;
    r.num_inliers = rg_num_correspondences(c);
    double br[6], cost = 0.0;
    detail::check(rg_linearize(c, T, H, br, &cost), c, "align");            // error and gradient at the result
    for (int i = 0; i < 6; i++) r.b(i) = br[i];
    r.error = cost;
    return r;
  }
};

}  // namespace small_gicp

#endif  // ROLO_B200_SMALL_GICP_ADAPTER_HPP
