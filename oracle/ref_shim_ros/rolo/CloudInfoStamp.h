// ref_shim_ros/rolo/CloudInfoStamp.h — the struct genmsg would generate from /root/reference/msg/CloudInfoStamp.msg:1-28
// (same field names and types). Test infrastructure only.
#ifndef ROLO_REF_SHIM_ROS_CLOUD_INFO_STAMP
#define ROLO_REF_SHIM_ROS_CLOUD_INFO_STAMP
#include "rolo_ros_shim.h"
namespace rolo {
struct CloudInfoStamp {
  std_msgs::Header header;
  std::vector<int32_t> startRingIndex, endRingIndex, pointColInd;
  std::vector<float> pointRange;
  float startOrientation = 0, endOrientation = 0, orientationDiff = 0;
  float initialGuessX = 0, initialGuessY = 0, initialGuessZ = 0, initialGuessRoll = 0, initialGuessPitch = 0, initialGuessYaw = 0;
  std::vector<float> covariance;
  uint8_t odomAvailable = 0;
  sensor_msgs::PointCloud2 cloud_projected, extracted_corner, extracted_surface, extracted_normal, extracted_ground;
};
using CloudInfoStampConstPtr = std::shared_ptr<const CloudInfoStamp>;
using CloudInfoStampPtr = std::shared_ptr<CloudInfoStamp>;
}  // namespace rolo
#endif
