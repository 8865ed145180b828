// drop_in_sequence.cpp — compiles the product's C++ drop-in class (include/rot_gicp/gicp/rot_vgicp.hpp) and drives it
// with the exact statement sequence of LidarOdometry::scanRegeistration (src/lidarOdometry.cpp:448-501 of the reference).
// Built by tests/test_gpu_cpp_dropin.py against oracle/ref_shim (the image has no PCL/Eigen; a ROLO checkout would use its
// own) and linked with librolo_gicp_b200.so. Prints the rotation estimate and the translation result.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <pcl/common/transforms.h>
#include <rot_gicp/gicp/rot_vgicp.hpp>
#include <rot_gicp/so3/so3.hpp>

// A USER subclass of the base class, as the reference's layout allows (lsq_registration.hpp:48-106): point-to-point cost against a
// fixed correspondence (point i of the source to point i of the target), J = [skew(p') | -I] like rot_vgicp_impl.hpp:265-267.
// It only implements the protected virtuals; align() and the whole Gauss-Newton / LM logic come from fast_gicp::LsqRegistration.
template <typename PointT>
class PointToPoint : public fast_gicp::LsqRegistration<PointT, PointT> {
 public:
  using Base = fast_gicp::LsqRegistration<PointT, PointT>;
  using PointCloudTargetConstPtr = typename Base::PointCloudTargetConstPtr;
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override { tgt_ = cloud; }
 protected:
  PointCloudTargetConstPtr tgt_;
  double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) override {
    double cost = 0.0;
    if (H && b) { H->setZero(); b->setZero(); }
    const Eigen::Matrix4d T = trans.matrix();
    for (size_t i = 0; i < this->input_->size(); i++) {
      const auto& s = this->input_->points[i]; const auto& t = tgt_->points[i];
      const double px = T(0,0)*s.x + T(0,1)*s.y + T(0,2)*s.z + T(0,3), py = T(1,0)*s.x + T(1,1)*s.y + T(1,2)*s.z + T(1,3), pz = T(2,0)*s.x + T(2,1)*s.y + T(2,2)*s.z + T(2,3);
      const double e[3] = {t.x - px, t.y - py, t.z - pz};
      cost += e[0]*e[0] + e[1]*e[1] + e[2]*e[2];
      if (H && b) {
        const double J[3][6] = {{0, -pz, py, -1, 0, 0}, {pz, 0, -px, 0, -1, 0}, {-py, px, 0, 0, 0, -1}};
        for (int r = 0; r < 3; r++) for (int a = 0; a < 6; a++) { (*b)(a) += J[r][a] * e[r]; for (int c = 0; c < 6; c++) (*H)(a, c) += J[r][a] * J[r][c]; }
      }
    }
    return cost;
  }
  double compute_error(const Eigen::Isometry3d& trans) override { return linearize(trans, nullptr, nullptr); }
  double compute_t_error(const Eigen::Vector3d&, const Eigen::Vector3d&, const Eigen::Vector3d&, const double&, const double&) override { return 0.0; }
};

using PointType = pcl::PointXYZI;

static pcl::PointCloud<PointType>::Ptr read_cloud(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END); long bytes = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  std::vector<float> buf(bytes / 4);
  if (std::fread(buf.data(), 1, bytes, f) != (size_t)bytes) std::exit(2);
  std::fclose(f);
  pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>);
  c->resize(buf.size() / 4);
  for (size_t i = 0; i < c->size(); i++) { c->points[i].x = buf[i * 4]; c->points[i].y = buf[i * 4 + 1]; c->points[i].z = buf[i * 4 + 2]; c->points[i].intensity = 1.f; }
  return c;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  pcl::PointCloud<PointType>::Ptr featureLast = read_cloud(argv[1]);
  pcl::PointCloud<PointType>::Ptr featureOld = read_cloud(argv[2]);
  Eigen::Affine3f transformation_interpolated = Eigen::Affine3f::Identity();
  Eigen::Vector3d TranslationOld(0.01, 0.0, 0.0);
  float CT_lambda = 0.3f;

  // ---- lidarOdometry.cpp:453-500, statement for statement ----
  pcl::PointCloud<PointType>::Ptr feature_propagated(new pcl::PointCloud<PointType>);
  pcl::PointCloud<PointType>::Ptr aligned(new pcl::PointCloud<PointType>);
  pcl::transformPointCloud(*featureOld, *feature_propagated, transformation_interpolated);
  fast_gicp::RotVGICP<PointType, PointType> rot_vgicp;
  rot_vgicp.setPolarResolution(0.175, 0.175, 2.0);
  rot_vgicp.setNumThreads(8);
  rot_vgicp.clearTarget();
  rot_vgicp.clearSource();
  rot_vgicp.setInputTarget(featureLast);
  rot_vgicp.setInputSource(feature_propagated);
  rot_vgicp.align(*aligned);
  Eigen::Matrix4f trans = rot_vgicp.getFinalTransformation();
  Eigen::Vector3d Translation(trans(0, 3), trans(1, 3), trans(2, 3));
  aligned->clear();
  Eigen::Vector3d Reg_translation = Eigen::Vector3d::Zero();
  rot_vgicp.computeTranslation(*aligned, Reg_translation, Translation, TranslationOld, 0.1, 0.1, CT_lambda);
  // -------------------------------------------------------------
  std::printf("converged=%d R0=[% .6f % .6f % .6f] t=[% .6f % .6f % .6f] aligned=%zu first=[% .5f % .5f % .5f]\n", (int)rot_vgicp.hasConverged(),
              trans(0, 0), trans(0, 1), trans(0, 2), Reg_translation(0), Reg_translation(1), Reg_translation(2), aligned->size(),
              aligned->points[0].x, aligned->points[0].y, aligned->points[0].z);
  // loop-closure style acceptance value (backMapping.cpp:2351): pcl::Registration::getFitnessScore()
  std::printf("fitness=%.12e\n", rot_vgicp.getFitnessScore());
  // the aliasing check of rot_vgicp_impl.hpp:150-152 must throw std::invalid_argument
  try { pcl::PointCloud<PointType>& alias = const_cast<pcl::PointCloud<PointType>&>(*feature_propagated); rot_vgicp.align(alias); std::printf("alias=NOTHROW\n"); }
  catch (const std::invalid_argument&) { std::printf("alias=invalid_argument\n"); }
  // evaluateCost + covariances round trip
  Eigen::Matrix<double, 6, 6> H; Eigen::Matrix<double, 6, 1> b;
  double c = rot_vgicp.evaluateCost(Eigen::Matrix4f::Identity(), &H, &b);
  std::printf("cost=%.6f H00=%.6f covs=%zu\n", c, H(0, 0), rot_vgicp.getSourceCovariances().size());
  // the base class on its own: a user subclass recovers a known rigid motion through the inherited align() (no GPU involved)
  {
    Eigen::Matrix<double, 6, 1> a; a << 0.03, -0.02, 0.05, 0.2, -0.1, 0.05;
    const Eigen::Matrix4f Tgt = fast_gicp::se3_exp(a).matrix().cast<float>();
    pcl::PointCloud<PointType>::Ptr src(new pcl::PointCloud<PointType>), dst(new pcl::PointCloud<PointType>), out(new pcl::PointCloud<PointType>);
    src->resize(200);
    for (size_t i = 0; i < src->size(); i++) { src->points[i].x = std::sin(0.37 * i) * 5.0f; src->points[i].y = std::cos(0.11 * i) * 4.0f; src->points[i].z = std::sin(0.05 * i + 1.0) * 2.0f; }
    pcl::transformPointCloud(*src, *dst, Tgt);
    PointToPoint<PointType> p2p;
    p2p.setOptimizerType(fast_gicp::LSQ_OPTIMIZER_TYPE::LevenbergMarquardt);
    p2p.setInputTarget(dst); p2p.setInputSource(src);
    p2p.align(*out);
    const Eigen::Matrix4f Tr = p2p.getFinalTransformation();
    float err = 0.f;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) err = std::max(err, std::abs(Tr(i, j) - Tgt(i, j)));
    std::printf("subclass converged=%d err=%.3e\n", (int)p2p.hasConverged(), err);
  }
  return 0;
}
