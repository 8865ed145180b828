// ref_shim_bm/pcl/kdtree/kdtree_flann.h — pcl::KdTreeFLANN<PointT> subset (setInputCloud, nearestKSearch) over the oracle's exact kd-tree
// (oracle/kdtree.hpp: float squared distances accumulated x, y, z; k smallest under (distance, index), ascending — what FLANN's exact
// single-tree search returns up to the order of equal distances). Test infrastructure only.
#ifndef ROLO_REF_SHIM_BM_KDTREE_FLANN
#define ROLO_REF_SHIM_BM_KDTREE_FLANN
#include <memory>
#include <vector>
#include <pcl/point_cloud.h>
#include "kdtree.hpp"
namespace pcl {
template <typename PointT> class KdTreeFLANN {
 public:
  using Ptr = std::shared_ptr<KdTreeFLANN<PointT>>;
  void setInputCloud(const typename PointCloud<PointT>::Ptr& cloud) {
    cloud_ = cloud;
    xyz_.resize(cloud->points.size() * 4);
    for (size_t i = 0; i < cloud->points.size(); i++) { xyz_[i * 4] = cloud->points[i].x; xyz_[i * 4 + 1] = cloud->points[i].y; xyz_[i * 4 + 2] = cloud->points[i].z; xyz_[i * 4 + 3] = 0.f; }
    tree_.build(xyz_.data(), (int)cloud->points.size());
  }
  int nearestKSearch(const PointT& p, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    k_indices.resize(k); k_sqr_distances.resize(k);
    const float q[3] = {p.x, p.y, p.z};
    const int found = tree_.knn_dist(q, k, k_indices.data(), k_sqr_distances.data());
    k_indices.resize(found); k_sqr_distances.resize(found);
    return found;
  }
 private:
  typename PointCloud<PointT>::Ptr cloud_;
  std::vector<float> xyz_;
  KdTree tree_;
};
}  // namespace pcl
#endif
