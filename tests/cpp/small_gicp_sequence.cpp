// small_gicp_sequence.cpp — the reference's commented VGICP prior-patch registration (src/backMapping.cpp:1993-2028), statement for
// statement, compiled against include/small_gicp (the B200 adapter) and run on the GPU core. Built by tests/test_gpu_small_gicp.py
// against oracle/ref_shim (no Eigen in the image). Prints the result pose, error and inlier count.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <small_gicp/ann/gaussian_voxelmap.hpp>
#include <small_gicp/factors/gicp_factor.hpp>
#include <small_gicp/points/point_cloud.hpp>
#include <small_gicp/registration/reduction_omp.hpp>
#include <small_gicp/registration/registration.hpp>
#include <small_gicp/util/normal_estimation_omp.hpp>

static std::vector<float> read_bin(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END); long bytes = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  std::vector<float> buf(bytes / 4);
  if (std::fread(buf.data(), 1, bytes, f) != (size_t)bytes) std::exit(2);
  std::fclose(f);
  return buf;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::vector<float> target_patch = read_bin(argv[1]), source_patch = read_bin(argv[2]);
  const double mappingSurfLeafSize = std::atof(argv[3]);
  const int numberOfCores = 4;

  auto target_points = std::make_shared<small_gicp::PointCloud>();
  target_points->resize(target_patch.size() / 4);
  for (size_t i = 0; i < target_points->size(); ++i) target_points->point(i) << target_patch[i * 4], target_patch[i * 4 + 1], target_patch[i * 4 + 2], 1.0;
  auto source_points = std::make_shared<small_gicp::PointCloud>();
  source_points->resize(source_patch.size() / 4);
  for (size_t i = 0; i < source_points->size(); ++i) source_points->point(i) << source_patch[i * 4], source_patch[i * 4 + 1], source_patch[i * 4 + 2], 1.0;

  small_gicp::estimate_covariances_omp(*target_points, 10, numberOfCores);
  small_gicp::estimate_covariances_omp(*source_points, 10, numberOfCores);

  auto target_voxelmap = std::make_shared<small_gicp::GaussianVoxelMap>(static_cast<double>(mappingSurfLeafSize));
  target_voxelmap->insert(*target_points);

  small_gicp::Registration<small_gicp::GICPFactor, small_gicp::ParallelReductionOMP> vgicp;
  vgicp.reduction.num_threads = numberOfCores;
  vgicp.rejector.max_dist_sq = 1.0;
  vgicp.optimizer.max_iterations = 50;

  small_gicp::RegistrationResult vgicp_result = vgicp.align(*target_voxelmap, *source_points, *target_voxelmap, Eigen::Isometry3d::Identity());

  Eigen::Matrix4d prior_patch_transform = vgicp_result.T_target_source.matrix();
  double prior_fitness_score = vgicp_result.error;
  std::printf("T=");
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) std::printf(" %.12e", prior_patch_transform(i, j));
  std::printf("\nerror=%.12e converged=%d iterations=%zu inliers=%zu cov00=%.9e\n", prior_fitness_score, (int)vgicp_result.converged, vgicp_result.iterations,
              vgicp_result.num_inliers, source_points->cov(0)(0, 0));
  return 0;
}
