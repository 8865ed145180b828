#ifndef ROLO_REF_SHIM_PCL_TRANSFORMS
#define ROLO_REF_SHIM_PCL_TRANSFORMS
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT> void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  if (&in != &out) out = in;
  for (size_t i = 0; i < in.size(); i++) {
    const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    out.points[i].x = T(0,0) * x + T(0,1) * y + T(0,2) * z + T(0,3);
    out.points[i].y = T(1,0) * x + T(1,1) * y + T(1,2) * z + T(1,3);
    out.points[i].z = T(2,0) * x + T(2,1) * y + T(2,2) * z + T(2,3);
  }
}
template <typename PointT> void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Affine3f& T) { transformPointCloud(in, out, T.matrix()); }
}  // namespace pcl
#endif
