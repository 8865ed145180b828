// ref_shim/pcl/point_types.h — minimal pcl::PointXYZ / PointXYZI / PointNormal (16-byte xyz1 header). Test infrastructure only.
#ifndef ROLO_REF_SHIM_PCL_POINT_TYPES
#define ROLO_REF_SHIM_PCL_POINT_TYPES
#include <memory>
#include <omp.h>
#include <Eigen/Core>
#include <pcl/pcl_config.h>
namespace pcl {
template <typename T> using shared_ptr = std::shared_ptr<T>;
struct PointXYZ {
  union { float data[4]; struct { float x, y, z; }; };
  PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
  PointXYZ(float a, float b, float c) : data{a, b, c, 1.f} {}
  Eigen::Vector4f getVector4fMap() const { return Eigen::Vector4f(data[0], data[1], data[2], data[3]); }
};
struct PointXYZI {
  union { float data[4]; struct { float x, y, z; }; };
  union { struct { float intensity; }; float data_c[4]; };
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
  Eigen::Vector4f getVector4fMap() const { return Eigen::Vector4f(data[0], data[1], data[2], data[3]); }
};
struct PointNormal {
  union { float data[4]; struct { float x, y, z; }; };
  union { float data_n[4]; struct { float normal_x, normal_y, normal_z; }; };
  union { struct { float curvature; }; float data_c[4]; };
  PointNormal() : data{0.f, 0.f, 0.f, 1.f}, data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
  Eigen::Vector4f getVector4fMap() const { return Eigen::Vector4f(data[0], data[1], data[2], data[3]); }
};
}  // namespace pcl
#endif
