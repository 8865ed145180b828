// ref_shim_ros/pcl_conversions — toROSMsg / fromROSMsg / moveFromROSMsg carry the typed cloud by value inside the message stand-in
// (the wire format is the subject of oracle/cloud_info_oracle.cpp, not of this build). Test infrastructure only.
#ifndef ROLO_REF_SHIM_ROS_PCL_CONVERSIONS
#define ROLO_REF_SHIM_ROS_PCL_CONVERSIONS
#include "rolo_ros_shim.h"
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT> void toROSMsg(const PointCloud<PointT>& c, sensor_msgs::PointCloud2& m) {
  m.payload = std::make_shared<PointCloud<PointT>>(c); m.width = (uint32_t)c.size(); m.height = 1; m.is_dense = c.is_dense;
}
template <typename PointT> void fromROSMsg(const sensor_msgs::PointCloud2& m, PointCloud<PointT>& c) {
  if (m.payload) c = *std::static_pointer_cast<PointCloud<PointT>>(m.payload); else c.clear();
}
template <typename PointT> void moveFromROSMsg(sensor_msgs::PointCloud2& m, PointCloud<PointT>& c) { fromROSMsg(m, c); }
namespace console { enum VERBOSITY_LEVEL { L_ALWAYS, L_ERROR, L_WARN, L_INFO, L_DEBUG, L_VERBOSE }; inline void setVerbosityLevel(VERBOSITY_LEVEL) {} }
}  // namespace pcl
#endif
