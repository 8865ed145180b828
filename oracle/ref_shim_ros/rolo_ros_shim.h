// rolo_ros_shim.h — MINIMAL stand-ins for the ROS 1 / tf / OpenCV / PCL names that the reference's node sources
//   /root/reference/src/imageProjection.cpp, /root/reference/src/featureExtraction.cpp (+ include/rolo/utility.h)
// use, so that those files compile FROM WHERE THEY LIE (nothing is copied) into oracle/_ref/librolo_ref_fe.so in an image that has
// none of ROS, PCL, OpenCV or Eigen. TEST INFRASTRUCTURE ONLY (oracle/): the member functions under test
// (projectPointCloud, deskewPoint, cloudExtraction, calculateSmoothness, markOccludedPoints, extractFeatures) run unmodified; everything
// here is plumbing (parameter server, publishers, message structs) except three third-party numerics restated from their
// published sources: pcl::VoxelGrid<PointT>::applyFilter (PCL 1.10 filters/impl/voxel_grid.hpp), pcl::getTransformation
// (PCL common/impl/eigen.hpp) and cv::Mat element access.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <unistd.h>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>

// ------------------------------------------------------------------------------------------------ Eigen extras
namespace Eigen {
enum { ColMajor = 0, RowMajor = 1 };
using Vector2d = Matrix<double, 2, 1>; using Vector2f = Matrix<float, 2, 1>;
// `v.col(0) << a, b, c, ...;` on a column view (imageProjection.cpp:382)
template <typename S> struct BlockCommaInit {
  S* p; int k;
  template <typename T> BlockCommaInit& operator,(T v) { p[k++] = (S)v; return *this; }
};
template <typename S, int R, int C, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
inline BlockCommaInit<S> operator<<(Block<S, R, C> b, T v) { b.p[0] = (S)v; return BlockCommaInit<S>{b.p, 1}; }
template <typename M> struct Map;                                 // only what utility.h's parameter helpers name
template <typename S, int R, int C> struct Map<const Matrix<S, R, C>> {
  const S* p;
  explicit Map(const S* q) : p(q) {}
  operator Matrix<S, R, C>() const { Matrix<S, R, C> m; for (int i = 0; i < R * C; i++) m.data()[i] = p[i]; return m; }
};
template <typename S, int R, int C, int O> struct MatrixRM {};    // row-major tag type for Map<const Matrix<double,3,3,RowMajor>>
}  // namespace Eigen

// ------------------------------------------------------------------------------------------------ ROS
namespace ros {
struct Time {
  double t = 0.0;
  Time() {}
  explicit Time(double s) : t(s) {}
  double toSec() const { return t; }
  static Time now() { return Time(0.0); }
};
struct Duration { double d = 0; explicit Duration(double s = 0) : d(s) {} double toSec() const { return d; } };
struct TransportHints { TransportHints& tcpNoDelay() { return *this; } };
struct Publisher {
  int getNumSubscribers() const { return 0; }
  template <typename M> void publish(const M&) const {}
};
struct Subscriber {};
// the parameter server: a process-global string -> string map the driver fills before constructing a node object
inline std::map<std::string, std::string>& param_store() { static std::map<std::string, std::string> m; return m; }
template <typename T> inline bool param_parse(const std::string& s, T& out) { std::istringstream is(s); is >> out; return !is.fail(); }
template <> inline bool param_parse<std::string>(const std::string& s, std::string& out) { out = s; return true; }
template <> inline bool param_parse<bool>(const std::string& s, bool& out) { out = (s == "1" || s == "true"); return true; }
template <typename T> inline bool param_parse(const std::string&, std::vector<T>&) { return false; }
struct NodeHandle {
  template <typename T, typename D> bool param(const std::string& name, T& var, const D& def) const {
    auto it = param_store().find(name);
    if (it != param_store().end() && param_parse(it->second, var)) return true;
    var = (T)def;
    return false;
  }
  template <typename M, typename C> Subscriber subscribe(const std::string&, int, void (C::*)(const std::shared_ptr<const M>&), C*, const TransportHints& = TransportHints()) { return Subscriber(); }
  template <typename M> Publisher advertise(const std::string&, int) { return Publisher(); }
};
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
inline void shutdown() {}
inline bool ok() { return true; }
struct MultiThreadedSpinner { explicit MultiThreadedSpinner(int = 0) {} void spin() {} };
struct Rate { explicit Rate(double) {} void sleep() {} };
}  // namespace ros
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)

// ------------------------------------------------------------------------------------------------ messages
namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
struct Float64MultiArray { std::vector<double> data; };
}
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {0}; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
}
namespace sensor_msgs {
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 1; };
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false; uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
  std::shared_ptr<void> payload;                                   // the typed cloud carried by value (toROSMsg / fromROSMsg)
};
using PointCloud2ConstPtr = std::shared_ptr<const PointCloud2>;
struct Imu {}; struct NavSatFix {}; struct Image {};
using ImagePtr = std::shared_ptr<Image>;
}
namespace nav_msgs {
struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; };
using OdometryConstPtr = std::shared_ptr<const Odometry>;
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; };
}
namespace visualization_msgs { struct Marker {}; struct MarkerArray {}; }
namespace jsk_recognition_msgs { struct BoundingBoxArray {}; }

// ------------------------------------------------------------------------------------------------ tf
namespace tf {
struct Quaternion { double x_ = 0, y_ = 0, z_ = 0, w_ = 1; };
inline void quaternionMsgToTF(const geometry_msgs::Quaternion& m, Quaternion& q) { q.x_ = m.x; q.y_ = m.y; q.z_ = m.z; q.w_ = m.w; }
struct Matrix3x3 {
  double m[3][3];
  explicit Matrix3x3(const Quaternion& q) {
    const double d = q.x_ * q.x_ + q.y_ * q.y_ + q.z_ * q.z_ + q.w_ * q.w_, s = 2.0 / d;
    const double xs = q.x_ * s, ys = q.y_ * s, zs = q.z_ * s, wx = q.w_ * xs, wy = q.w_ * ys, wz = q.w_ * zs;
    const double xx = q.x_ * xs, xy = q.x_ * ys, xz = q.x_ * zs, yy = q.y_ * ys, yz = q.y_ * zs, zz = q.z_ * zs;
    m[0][0] = 1.0 - (yy + zz); m[0][1] = xy - wz; m[0][2] = xz + wy;
    m[1][0] = xy + wz; m[1][1] = 1.0 - (xx + zz); m[1][2] = yz - wx;
    m[2][0] = xz - wy; m[2][1] = yz + wx; m[2][2] = 1.0 - (xx + yy);
  }
  void getRPY(double& roll, double& pitch, double& yaw) const {   // not on the path under test (odometry handler only)
    pitch = std::asin(-m[2][0]); roll = std::atan2(m[2][1], m[2][2]); yaw = std::atan2(m[1][0], m[0][0]);
  }
};
}  // namespace tf

// ------------------------------------------------------------------------------------------------ OpenCV: the one matrix the path uses
#define CV_32F 5
#define CV_16UC1 2
namespace cv {
struct Scalar { double v[4]; static Scalar all(double x) { Scalar s; s.v[0] = s.v[1] = s.v[2] = s.v[3] = x; return s; } };
struct Mat {
  int rows = 0, cols = 0;
  std::vector<float> d;
  Mat() {}
  Mat(int r, int c, int /*type: CV_32F*/, const Scalar& s) : rows(r), cols(c), d((size_t)r * c, (float)s.v[0]) {}
  template <typename T> T& at(int r, int c) { static_assert(std::is_same<T, float>::value, "CV_32F only"); return d[(size_t)r * cols + c]; }
  template <typename T> const T& at(int r, int c) const { return d[(size_t)r * cols + c]; }
  void copyTo(Mat& o) const { o = *this; }              // deep copy; this stand-in has value semantics throughout
  Mat inv() const;                                       // defined in ref_shim_bm/cv_ops.hpp (only the scan-to-map driver needs it)
};
}  // namespace cv
