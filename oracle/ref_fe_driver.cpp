// ref_fe_driver.cpp — C API over the REFERENCE's own range-image node sources, compiled from where they lie:
//   /root/reference/src/imageProjection.cpp   (class ImageProjection: projectPointCloud :399-474, deskewPoint :368-396, cloudExtraction :477-505)
//   /root/reference/src/featureExtraction.cpp (class FeatureExtraction: calculateSmoothness :87-110, markOccludedPoints :112-150,
//                                              extractFeatures :153-266)
// against oracle/ref_shim_ros (ROS / tf / OpenCV / PCL stand-ins) + oracle/ref_shim (Eigen stand-in), for pinning
// oracle/frontend_oracle.cpp. Test infrastructure; built only where the reference checkout exists (`make -C oracle ref`), output
// in oracle/_ref/. The member functions under test run unmodified; this file only fills their inputs (the members the ROS
// callbacks would fill) and reads their outputs.
#include "rolo_ros_shim.h"
#include <pcl/point_cloud.h>
#include <pcl/common/transforms.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl_conversions/pcl_conversions.h>

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
#define private public
#define protected public
#define main ref_image_projection_main
#include "src/imageProjection.cpp"
#undef main
#define main ref_feature_extraction_main
#include "src/featureExtraction.cpp"
#undef main
#undef private
#undef protected
#undef Matrix

namespace {
void set_param(const char* k, const std::string& v) { ros::param_store()[k] = v; }
void common_params(int n_scan, int horizon) {
  set_param("rolo/sensor", "velodyne");
  set_param("rolo/N_SCAN", std::to_string(n_scan));
  set_param("rolo/Horizon_SCAN", std::to_string(horizon));
}
std::string fstr(float v) { char b[64]; snprintf(b, sizeof b, "%.9g", (double)v); return b; }
}  // namespace

extern "C" {

// returns the number of extracted points. Raw points: x, y, z float at byte offsets 0, 4, 8; ring uint16 at ring_offset_bytes; relative
// time float at time_offset_bytes (deskew only). `deskew` != 0 reproduces deskewEnabled && odomAvailable with the given odometry increment.
int rfe_image_projection(const void* raw, size_t n_points, int stride_bytes, int ring_offset_bytes, int n_scan, int horizon_scan,
                         int downsample_rate, float lidar_min_range, float lidar_max_range, int deskew, int time_offset_bytes,
                         const float* incre_rpy, float scan_period, double odom_time_diff, float* range_image,
                         float* extracted_xyzi, int* point_col_ind, float* point_range, int* start_ring_index, int* end_ring_index) {
  common_params(n_scan, horizon_scan);
  set_param("rolo/downsampleRate", std::to_string(downsample_rate));
  set_param("rolo/lidarMinRange", fstr(lidar_min_range));
  set_param("rolo/lidarMaxRange", fstr(lidar_max_range));
  set_param("rolo/deskewEnabled", deskew ? "true" : "false");
  ImageProjection ip;                                   // constructor: allocateMemory() + resetParameters()
  const unsigned char* base = (const unsigned char*)raw;
  ip.laserCloudIn->points.resize(n_points);
  ip.deskewCloud->points.resize(n_points);
  for (size_t i = 0; i < n_points; i++) {               // what cachePointCloud() / deskewCloudInfo() leave in the members
    const float* p = (const float*)(base + i * stride_bytes);
    uint16_t ring; std::memcpy(&ring, base + i * stride_bytes + ring_offset_bytes, sizeof(ring));
    auto& d = ip.laserCloudIn->points[i];
    d.x = p[0]; d.y = p[1]; d.z = p[2]; d.intensity = 0.f; d.ring = ring; d.time = 0.f;
    if (deskew) {
      float rel; std::memcpy(&rel, base + i * stride_bytes + time_offset_bytes, sizeof(rel));
      d.time = rel;
      ip.deskewCloud->points[i].intensity = std::fabs(rel);     // imageProjection.cpp:358-359 (point cloud with its own time stamps)
    }
  }
  if (deskew) {
    ip.odomAvailable = true; ip.scanPeriod = scan_period; ip.odomTimeDiff = odom_time_diff;
    ip.odomIncreRoll = incre_rpy[0]; ip.odomIncrePitch = incre_rpy[1]; ip.odomIncreYaw = incre_rpy[2];
    ip.odomIncreX = ip.odomIncreY = ip.odomIncreZ = 0.f;
  } else {
    ip.odomAvailable = false;
  }
  ip.projectPointCloud();
  ip.cloudExtraction();
  const size_t cells = (size_t)n_scan * horizon_scan;
  for (size_t c = 0; c < cells; c++) range_image[c] = ip.rangeMat.at<float>((int)(c / horizon_scan), (int)(c % horizon_scan));
  const int n = (int)ip.extractedCloud->points.size();
  for (int i = 0; i < n; i++) {
    const auto& p = ip.extractedCloud->points[i];
    extracted_xyzi[i * 4 + 0] = p.x; extracted_xyzi[i * 4 + 1] = p.y; extracted_xyzi[i * 4 + 2] = p.z; extracted_xyzi[i * 4 + 3] = p.intensity;
    point_col_ind[i] = ip.cloudInfoStamp.pointColInd[i];
    point_range[i] = ip.cloudInfoStamp.pointRange[i];
  }
  for (int r = 0; r < n_scan; r++) { start_ring_index[r] = ip.cloudInfoStamp.startRingIndex[r]; end_ring_index[r] = ip.cloudInfoStamp.endRingIndex[r]; }
  return n;
}

// Work arrays: the reference allocates them once with new[] and never clears them (quirk Q8); here they start at zero with
// cloudSmoothness[i].ind = i, the decision frontend_oracle.cpp documents, so the comparison is about the algorithm and not about heap garbage.
int rfe_feature_extraction(const float* extracted_xyzi, const int* point_col_ind, const float* point_range, const int* start_ring_index,
                           const int* end_ring_index, int n, int n_scan, int horizon_scan, float edge_threshold, float surf_threshold,
                           float surf_leaf_size, float* curvature, int* neighbor_picked, int* label, float* corner_xyzi, int* n_corner,
                           float* surface_xyzi, int* n_surface) {
  common_params(n_scan, horizon_scan);
  set_param("rolo/edgeThreshold", fstr(edge_threshold));
  set_param("rolo/surfThreshold", fstr(surf_threshold));
  set_param("rolo/odometrySurfLeafSize", fstr(surf_leaf_size));
  FeatureExtraction fe;
  const size_t cells = (size_t)n_scan * horizon_scan;
  // guard band: the reference's neighbour marking can touch indices -5 .. n+5 (Q8); the arrays are re-seated with room on both sides
  const int G = 16;
  std::vector<float> curv(cells + 2 * G, 0.f); std::vector<int> picked(cells + 2 * G, 0), lab(cells + 2 * G, 0);
  delete[] fe.cloudCurvature; delete[] fe.cloudNeighborPicked; delete[] fe.cloudLabel;
  fe.cloudCurvature = curv.data() + G; fe.cloudNeighborPicked = picked.data() + G; fe.cloudLabel = lab.data() + G;
  for (size_t i = 0; i < fe.cloudSmoothness.size(); i++) { fe.cloudSmoothness[i].value = 0.f; fe.cloudSmoothness[i].ind = i; }
  fe.extractedCloud->points.resize(n);
  for (int i = 0; i < n; i++) {
    auto& p = fe.extractedCloud->points[i];
    p.x = extracted_xyzi[i * 4 + 0]; p.y = extracted_xyzi[i * 4 + 1]; p.z = extracted_xyzi[i * 4 + 2]; p.intensity = extracted_xyzi[i * 4 + 3];
  }
  // pointColInd is read at ind +- 5 around every candidate: pad it like the message arrays the node receives (N_SCAN * Horizon_SCAN entries)
  fe.cloudInfo.pointColInd.assign(cells + G, 0); fe.cloudInfo.pointRange.assign(cells + G, 0.f);
  for (int i = 0; i < n; i++) { fe.cloudInfo.pointColInd[i] = point_col_ind[i]; fe.cloudInfo.pointRange[i] = point_range[i]; }
  fe.cloudInfo.startRingIndex.assign(start_ring_index, start_ring_index + n_scan);
  fe.cloudInfo.endRingIndex.assign(end_ring_index, end_ring_index + n_scan);
  fe.calculateSmoothness();
  fe.markOccludedPoints();
  fe.extractFeatures();
  for (int i = 0; i < n; i++) { curvature[i] = fe.cloudCurvature[i]; neighbor_picked[i] = fe.cloudNeighborPicked[i]; label[i] = fe.cloudLabel[i]; }
  *n_corner = (int)fe.cornerCloud->points.size(); *n_surface = (int)fe.surfaceCloud->points.size();
  for (int i = 0; i < *n_corner; i++) { const auto& p = fe.cornerCloud->points[i]; corner_xyzi[i * 4] = p.x; corner_xyzi[i * 4 + 1] = p.y; corner_xyzi[i * 4 + 2] = p.z; corner_xyzi[i * 4 + 3] = p.intensity; }
  for (int i = 0; i < *n_surface; i++) { const auto& p = fe.surfaceCloud->points[i]; surface_xyzi[i * 4] = p.x; surface_xyzi[i * 4 + 1] = p.y; surface_xyzi[i * 4 + 2] = p.z; surface_xyzi[i * 4 + 3] = p.intensity; }
  fe.cloudCurvature = nullptr; fe.cloudNeighborPicked = nullptr; fe.cloudLabel = nullptr;   // owned by the vectors above
  return 0;
}

}  // extern "C"
