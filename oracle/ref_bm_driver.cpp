// ref_bm_driver.cpp — C API over the REFERENCE's own scan-to-map member functions (class mapOptimization, src/backMapping.cpp):
//   pointAssociateToMap, trans2Affine3f, scan2MapOptimization, updatePointAssociateToMap, cornerOptimization, surfOptimization,
//   combineOptimizationCoeffs, LMOptimization
// for pinning oracle/scan2map_oracle.cpp. backMapping.cpp cannot be compiled as a file here (roscpp, tf, GTSAM, small_gicp, PCL I/O),
// so oracle/extract_ref_members.py takes the text of exactly those member functions from the reference file where it lies, at build time,
// and splices it into this file at the marker below on its way into the compiler (a pipe: the reference text is never written to disk),
// inside a class that declares the members they use — with the reference's own include/rolo/utility.h in front (`using namespace std`, PointType, the overload sets the float /
// double promotions depend on). The functions run unmodified. Stand-ins: oracle/ref_shim_bm (kd-tree, cv:: operations, Eigen QR over
// oracle/s2m_numerics.hpp), oracle/ref_shim_ros (ROS, PCL containers, cv::Mat), oracle/ref_shim (Eigen). Test infrastructure; built only
// where the reference checkout exists (`make -C oracle ref`), output in oracle/_ref/.
#include "rolo_ros_shim.h"
#include <pcl/point_cloud.h>
#include <pcl/common/transforms.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl_conversions/pcl_conversions.h>
#include <pcl/kdtree/kdtree_flann.h>
#include "cv_ops.hpp"
#include "eigen_colpivqr.hpp"

// Eigen::Matrix<S, R, C, Options> with the 4th parameter utility.h's parameter helpers name (never executed here)
namespace Eigen {
template <typename S, int R, int C> struct RowMajorMat {};
template <typename S, int R, int C> struct Map<const RowMajorMat<S, R, C>> {
  const S* p;
  explicit Map(const S* q) : p(q) {}
  operator Matrix<S, R, C>() const { Matrix<S, R, C> m; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) m(i, j) = p[i * C + j]; return m; }
};
template <typename S, int R, int C, int O = 0> using MatrixO = typename std::conditional<O == 0, Matrix<S, R, C>, RowMajorMat<S, R, C>>::type;
}  // namespace Eigen

#define Matrix MatrixO
#include "rolo/utility.h"

namespace pcl {
inline float rad2deg(float alpha) { return alpha * 57.29578f; }      // pcl/common/impl/angles.hpp
inline double rad2deg(double alpha) { return alpha * 57.29578; }
}  // namespace pcl

// the members of class mapOptimization the extracted functions touch (backMapping.cpp:75-190), nothing else
class ScanToMapMembers {
 public:
  pcl::PointCloud<PointType>::Ptr cloudKeyPoses3D{new pcl::PointCloud<PointType>()};
  pcl::PointCloud<PointType>::Ptr laserCloudCornerLastDS{new pcl::PointCloud<PointType>()}, laserCloudSurfLastDS{new pcl::PointCloud<PointType>()};
  pcl::PointCloud<PointType>::Ptr laserCloudCornerFromMapDS{new pcl::PointCloud<PointType>()}, laserCloudSurfFromMapDS{new pcl::PointCloud<PointType>()};
  pcl::PointCloud<PointType>::Ptr laserCloudOri{new pcl::PointCloud<PointType>()}, coeffSel{new pcl::PointCloud<PointType>()};
  pcl::KdTreeFLANN<PointType>::Ptr kdtreeCornerFromMap{new pcl::KdTreeFLANN<PointType>()}, kdtreeSurfFromMap{new pcl::KdTreeFLANN<PointType>()};
  std::vector<PointType> laserCloudOriCornerVec, coeffSelCornerVec, laserCloudOriSurfVec, coeffSelSurfVec;
  std::vector<bool> laserCloudOriCornerFlag, laserCloudOriSurfFlag;
  int laserCloudCornerLastDSNum = 0, laserCloudSurfLastDSNum = 0;
  int edgeFeatureMinValidNum = 10, surfFeatureMinValidNum = 100, numberOfCores = 4;
  float transformTobeMapped[6] = {0, 0, 0, 0, 0, 0};
  Eigen::Affine3f transPointAssociateToMap;
  bool isDegenerate = false;
  cv::Mat matP{6, 6, CV_32F, cv::Scalar::all(0)};
  int lm_calls = 0, lm_stop_after = 1 << 30;
  bool lm_converged = false;
  void transformUpdate() {}     // IMU roll / pitch fusion and the z clamp (backMapping.cpp:1060-1068): outside the path under test

ROLO_REFERENCE_MEMBER_FUNCTIONS_GO_HERE   // replaced by extract_ref_members.py in the pipe to the compiler
};
#undef Matrix

namespace {
void fill(pcl::PointCloud<PointType>& c, const float* p, size_t n, size_t stride) {
  c.points.resize(n);
  for (size_t i = 0; i < n; i++) { c.points[i].x = p[i * stride]; c.points[i].y = p[i * stride + 1]; c.points[i].z = p[i * stride + 2]; c.points[i].intensity = stride > 3 ? p[i * stride + 3] : 0.f; }
}
}  // namespace

extern "C" {

void* rbm_create() { auto* m = new ScanToMapMembers(); m->cloudKeyPoses3D->points.resize(1); return m; }
void rbm_destroy(void* h) { delete (ScanToMapMembers*)h; }
void rbm_set_num_threads(void* h, int n) { ((ScanToMapMembers*)h)->numberOfCores = n; }
void rbm_set_map(void* h, const float* corner, size_t n_corner, const float* surf, size_t n_surf, size_t stride_floats) {
  auto& m = *(ScanToMapMembers*)h;
  fill(*m.laserCloudCornerFromMapDS, corner, n_corner, stride_floats); fill(*m.laserCloudSurfFromMapDS, surf, n_surf, stride_floats);
  m.kdtreeCornerFromMap->setInputCloud(m.laserCloudCornerFromMapDS); m.kdtreeSurfFromMap->setInputCloud(m.laserCloudSurfFromMapDS);
}
/* what downsampleCurrentScan() (:657-678) and the constructor's resize calls (:250-256) leave in the members */
void rbm_set_scan(void* h, const float* corner, size_t n_corner, const float* surf, size_t n_surf, size_t stride_floats) {
  auto& m = *(ScanToMapMembers*)h;
  fill(*m.laserCloudCornerLastDS, corner, n_corner, stride_floats); fill(*m.laserCloudSurfLastDS, surf, n_surf, stride_floats);
  m.laserCloudCornerLastDSNum = (int)n_corner; m.laserCloudSurfLastDSNum = (int)n_surf;
  m.laserCloudOriCornerVec.assign(n_corner, PointType()); m.coeffSelCornerVec.assign(n_corner, PointType()); m.laserCloudOriCornerFlag.assign(n_corner, false);
  m.laserCloudOriSurfVec.assign(n_surf, PointType()); m.coeffSelSurfVec.assign(n_surf, PointType()); m.laserCloudOriSurfFlag.assign(n_surf, false);
}

/* cornerOptimization() + surfOptimization() at pose tf (roll pitch yaw x y z), then the flags / coefficients they left, then
 * combineOptimizationCoeffs() and the first part of LMOptimization reproduced through its own products: AtA, AtB of the selected rows are
 * returned by running LMOptimization on a copy of the pose and reading nothing else — so this entry only reports selection + coefficients;
 * the normal equations are compared through rbm_optimize's pose trajectory. Returns laserCloudSelNum. */
int rbm_associate(void* h, const float tf[6], float* coeff_corner, uint8_t* flag_corner, float* coeff_surf, uint8_t* flag_surf) {
  auto& m = *(ScanToMapMembers*)h;
  for (int i = 0; i < 6; i++) m.transformTobeMapped[i] = tf[i];
  m.laserCloudOri->clear(); m.coeffSel->clear();
  m.cornerOptimization(); m.surfOptimization();
  for (int i = 0; i < m.laserCloudCornerLastDSNum; i++) {
    const bool f = m.laserCloudOriCornerFlag[i];
    if (flag_corner) flag_corner[i] = f ? 1 : 0;
    if (coeff_corner) { const PointType& c = m.coeffSelCornerVec[i]; coeff_corner[i * 4] = f ? c.x : 0.f; coeff_corner[i * 4 + 1] = f ? c.y : 0.f; coeff_corner[i * 4 + 2] = f ? c.z : 0.f; coeff_corner[i * 4 + 3] = f ? c.intensity : 0.f; }
  }
  for (int i = 0; i < m.laserCloudSurfLastDSNum; i++) {
    const bool f = m.laserCloudOriSurfFlag[i];
    if (flag_surf) flag_surf[i] = f ? 1 : 0;
    if (coeff_surf) { const PointType& c = m.coeffSelSurfVec[i]; coeff_surf[i * 4] = f ? c.x : 0.f; coeff_surf[i * 4 + 1] = f ? c.y : 0.f; coeff_surf[i * 4 + 2] = f ? c.z : 0.f; coeff_surf[i * 4 + 3] = f ? c.intensity : 0.f; }
  }
  m.combineOptimizationCoeffs();
  return (int)m.laserCloudOri->size();
}

/* one Gauss-Newton step of the reference at pose tf: cornerOptimization, surfOptimization, combineOptimizationCoeffs, LMOptimization(iter).
 * tf inout; returns LMOptimization's result (1 = converged); *degenerate = isDegenerate; *selected = laserCloudSelNum */
int rbm_step(void* h, float tf[6], int iter_count, int* degenerate, int* selected) {
  auto& m = *(ScanToMapMembers*)h;
  for (int i = 0; i < 6; i++) m.transformTobeMapped[i] = tf[i];
  m.laserCloudOri->clear(); m.coeffSel->clear();
  m.cornerOptimization(); m.surfOptimization(); m.combineOptimizationCoeffs();
  if (selected) *selected = (int)m.laserCloudOri->size();
  const bool conv = m.LMOptimization(iter_count);
  for (int i = 0; i < 6; i++) tf[i] = m.transformTobeMapped[i];
  if (degenerate) *degenerate = m.isDegenerate ? 1 : 0;
  return conv ? 1 : 0;
}

/* scan2MapOptimization() as the reference runs it (30 iterations at most, :681-710). tf inout. Returns 0, or -1 when the scan has too
 * few features (the ROS_WARN branch). */
int rbm_scan2map(void* h, float tf[6], int edge_min, int surf_min, int* degenerate) {
  auto& m = *(ScanToMapMembers*)h;
  m.edgeFeatureMinValidNum = edge_min; m.surfFeatureMinValidNum = surf_min;
  for (int i = 0; i < 6; i++) m.transformTobeMapped[i] = tf[i];
  m.isDegenerate = false;
  const bool enough = m.laserCloudCornerLastDSNum > edge_min && m.laserCloudSurfLastDSNum > surf_min;
  m.scan2MapOptimization();
  for (int i = 0; i < 6; i++) tf[i] = m.transformTobeMapped[i];
  if (degenerate) *degenerate = m.isDegenerate ? 1 : 0;
  return enough ? 0 : -1;
}
}
