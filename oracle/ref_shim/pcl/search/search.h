#ifndef ROLO_REF_SHIM_PCL_SEARCH
#define ROLO_REF_SHIM_PCL_SEARCH
#include <algorithm>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl { namespace search {
template <typename PointT> class Search {
 public:
  using PointCloudConstPtr = typename PointCloud<PointT>::ConstPtr;
  virtual ~Search() {}
  virtual void setInputCloud(const PointCloudConstPtr& cloud) = 0;
  virtual PointCloudConstPtr getInputCloud() const = 0;
  virtual int nearestKSearch(const PointT& p, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const = 0;
};
}}  // namespace pcl::search
#endif
