// ref_shim_ros/pcl/common/transforms.h — pcl::getTransformation / getTranslationAndEulerAngles restated from PCL common/impl/eigen.hpp
// (float overloads). Test infrastructure only.
#ifndef ROLO_REF_SHIM_ROS_PCL_TRANSFORMS
#define ROLO_REF_SHIM_ROS_PCL_TRANSFORMS
#include <cmath>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
namespace pcl {
inline Eigen::Affine3f getTransformation(float x, float y, float z, float roll, float pitch, float yaw) {
  Eigen::Affine3f t;
  float A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch), E = std::cos(roll), F = std::sin(roll), DE = D * E, DF = D * F;
  t(0, 0) = A * C; t(0, 1) = A * DF - B * E; t(0, 2) = B * F + A * DE; t(0, 3) = x;
  t(1, 0) = B * C; t(1, 1) = A * E + B * DF; t(1, 2) = B * DE - A * F; t(1, 3) = y;
  t(2, 0) = -D;    t(2, 1) = C * F;          t(2, 2) = C * E;          t(2, 3) = z;
  t(3, 0) = 0;     t(3, 1) = 0;              t(3, 2) = 0;              t(3, 3) = 1;
  return t;
}
inline void getTranslationAndEulerAngles(const Eigen::Affine3f& t, float& x, float& y, float& z, float& roll, float& pitch, float& yaw) {
  x = t(0, 3); y = t(1, 3); z = t(2, 3);
  roll = std::atan2(t(2, 1), t(2, 2)); pitch = std::asin(-t(2, 0)); yaw = std::atan2(t(1, 0), t(0, 0));
}
}  // namespace pcl
#endif
