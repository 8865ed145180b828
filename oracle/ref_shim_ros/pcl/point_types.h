// ref_shim_ros/pcl/point_types.h — pcl::PointXYZI and the PCL point-struct macros the node sources use. Test infrastructure only.
#ifndef ROLO_REF_SHIM_ROS_PCL_POINT_TYPES
#define ROLO_REF_SHIM_ROS_PCL_POINT_TYPES
#include <memory>
#include <Eigen/Core>
#define PCL_ADD_POINT4D union { float data[4]; struct { float x; float y; float z; }; };
#define PCL_ADD_INTENSITY union { struct { float intensity; }; float data_c[4]; }
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fseq)
namespace pcl {
template <typename T> using shared_ptr = std::shared_ptr<T>;
struct PointXYZ {
  PCL_ADD_POINT4D
  PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
};
struct PointXYZI {
  PCL_ADD_POINT4D
  PCL_ADD_INTENSITY;
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
}  // namespace pcl
#endif
