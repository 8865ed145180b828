// ref_shim_ros/pcl/filters/voxel_grid.h — pcl::VoxelGrid<PointT> restated from PCL 1.10 (filters/impl/voxel_grid.hpp: getMinMax3D,
// applyFilter with downsample_all_data_ = true, min_points_per_voxel_ = 0; common/impl/accumulators.hpp: AccumulatorXYZ keeps an
// Eigen::Vector3f sum, AccumulatorIntensity a float sum, both divided by the point count as float). The real PCL sorts the
// (leaf index, point index) records with std::sort on the leaf index ALONE, so the summation order inside a leaf is whatever that
// unstable sort leaves; the same std::sort call is used here. Third-party code absent from the reference tree: test infrastructure only.
#ifndef ROLO_REF_SHIM_ROS_PCL_VOXEL_GRID
#define ROLO_REF_SHIM_ROS_PCL_VOXEL_GRID
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT> class VoxelGrid {
  struct cloud_point_index_idx {
    unsigned int idx, cloud_point_index;
    cloud_point_index_idx(unsigned int idx_, unsigned int cpi) : idx(idx_), cloud_point_index(cpi) {}
    bool operator<(const cloud_point_index_idx& p) const { return idx < p.idx; }
  };
  typename PointCloud<PointT>::ConstPtr input_;
  float leaf_[3] = {0, 0, 0}, inv_[3] = {0, 0, 0};
 public:
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; inv_[0] = 1.0f / lx; inv_[1] = 1.0f / ly; inv_[2] = 1.0f / lz; }
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
  void filter(PointCloud<PointT>& output) {
    output.points.clear(); output.height = 1; output.is_dense = true; output.width = 0;
    if (!input_ || input_->points.empty()) return;
    const auto& in = input_->points;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const auto& p : in) {                                                  // getMinMax3D (dense cloud)
      mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
      mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    int min_b[3], max_b[3], div_b[3];
    for (int d = 0; d < 3; d++) {
      min_b[d] = static_cast<int>(std::floor(mn[d] * inv_[d]));
      max_b[d] = static_cast<int>(std::floor(mx[d] * inv_[d]));
      div_b[d] = max_b[d] - min_b[d] + 1;
    }
    const int divb_mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    std::vector<cloud_point_index_idx> index_vector;
    index_vector.reserve(in.size());
    for (unsigned i = 0; i < (unsigned)in.size(); i++) {
      const int ijk0 = static_cast<int>(std::floor(in[i].x * inv_[0]) - static_cast<float>(min_b[0]));
      const int ijk1 = static_cast<int>(std::floor(in[i].y * inv_[1]) - static_cast<float>(min_b[1]));
      const int ijk2 = static_cast<int>(std::floor(in[i].z * inv_[2]) - static_cast<float>(min_b[2]));
      const int idx = ijk0 * divb_mul[0] + ijk1 * divb_mul[1] + ijk2 * divb_mul[2];
      index_vector.emplace_back(static_cast<unsigned int>(idx), i);
    }
    std::sort(index_vector.begin(), index_vector.end(), std::less<cloud_point_index_idx>());
    size_t first = 0;
    while (first < index_vector.size()) {
      size_t last = first + 1;
      while (last < index_vector.size() && index_vector[last].idx == index_vector[first].idx) ++last;
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      for (size_t li = first; li < last; ++li) {
        const PointT& p = in[index_vector[li].cloud_point_index];
        sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
      }
      const float n = static_cast<float>(last - first);
      PointT o;
      o.x = sx / n; o.y = sy / n; o.z = sz / n; o.intensity = si / n;
      output.points.push_back(o);
      first = last;
    }
    output.width = (uint32_t)output.points.size();
  }
};
}  // namespace pcl
#endif
