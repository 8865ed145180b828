#ifndef ROLO_REF_SHIM_PCL_POINT_CLOUD
#define ROLO_REF_SHIM_PCL_POINT_CLOUD
#include <vector>
#include <pcl/point_types.h>
namespace pcl {
template <typename PointT> class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  unsigned width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  const PointT& at(size_t i) const { return points.at(i); }
  PointT& at(size_t i) { return points.at(i); }
  const PointT& operator[](size_t i) const { return points[i]; }
  void resize(size_t n) { points.resize(n); width = (unsigned)n; height = 1; }
  void clear() { points.clear(); width = 0; }
};
}  // namespace pcl
#endif
