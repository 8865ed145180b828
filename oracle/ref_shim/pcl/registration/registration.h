// ref_shim/pcl/registration/registration.h — the members of pcl::Registration (PCL >= 1.10) that rot_gicp touches.
// align() follows PCL: initCompute bookkeeping, converged_ = false, final/transformation_ = Identity, then the
// virtual computeTransformation(output, guess). Test infrastructure only.
#ifndef ROLO_REF_SHIM_PCL_REGISTRATION
#define ROLO_REF_SHIM_PCL_REGISTRATION
#include <string>
#include <Eigen/Geometry>
#include <pcl/common/transforms.h>
#include <pcl/point_cloud.h>
#include <pcl/search/kdtree.h>
namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float> class Registration {
 public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using PointCloudSource = PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  Registration() { final_transformation_.setIdentity(); }
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  Matrix4 getFinalTransformation() { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  using KdTreePtr = std::shared_ptr<pcl::search::KdTree<PointTarget>>;
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false) { tree_ = tree; force_no_recompute_ = force_no_recompute; }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    output = *input_;
    converged_ = false;
    final_transformation_ = Matrix4::Identity();
    computeTransformation(output, guess);
  }
 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_ = 0, max_iterations_ = 10;
  Matrix4 final_transformation_;
  double transformation_epsilon_ = 0.0;
  bool converged_ = false;
  KdTreePtr tree_;
  bool force_no_recompute_ = false;
};
}  // namespace pcl
#endif
