// ref_shim_ros/pcl/point_cloud.h — pcl::PointCloud subset (vector of points + width/height/is_dense/header). Test infrastructure only.
#ifndef ROLO_REF_SHIM_ROS_PCL_POINT_CLOUD
#define ROLO_REF_SHIM_ROS_PCL_POINT_CLOUD
#include <cstdint>
#include <string>
#include <vector>
#include <pcl/point_types.h>
namespace pcl {
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };
template <typename PointT> class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  const PointT& at(size_t i) const { return points.at(i); }
  PointT& at(size_t i) { return points.at(i); }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointT& operator[](size_t i) { return points[i]; }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
  void clear() { points.clear(); width = 0; height = 0; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  PointCloud& operator+=(const PointCloud& o) {            // pcl::PointCloud::operator+= : append
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = (uint32_t)points.size(); height = 1;
    if (!o.is_dense) is_dense = false;
    return *this;
  }
  const PointT& back() const { return points.back(); }
  PointT& back() { return points.back(); }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
};
}  // namespace pcl
#endif
