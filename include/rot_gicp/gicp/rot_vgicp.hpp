// rot_vgicp.hpp — drop-in fast_gicp::RotVGICP<PointSource, PointTarget> over the B200 C-ABI.
//
// Same header path, namespace, class name and public members as the reference
// (include/rot_gicp/gicp/rot_vgicp.hpp:24-104, lsq_registration.hpp:48-62), so
// src/lidarOdometry.cpp:460-494 compiles and runs unchanged:
//     fast_gicp::RotVGICP<PointType, PointType> rot_vgicp;
//     rot_vgicp.setPolarResolution(0.175, 0.175, 2.0); rot_vgicp.setNumThreads(omp_get_max_threads());
//     rot_vgicp.clearTarget(); rot_vgicp.clearSource();
//     rot_vgicp.setInputTarget(featureLast); rot_vgicp.setInputSource(feature_propagated);
//     rot_vgicp.align(*aligned); rot_vgicp.getFinalTransformation();
//     rot_vgicp.computeTranslation(*aligned, Reg_translation, Translation, TranslationOld, 0.1, 0.1, CT_lambda);
// Everything numeric happens in librolo_gicp_b200.so (link it instead of the reference's librot_gicp.so).
// This header is header-only glue; it needs the caller's own PCL (>= 1.10) and Eigen (>= 3.3.7), exactly
// like the reference header. It cannot be compiled in the build image of this repository (no PCL/Eigen
// there); the same C-ABI calls are exercised by examples/lidar_odometry_sequence.c and by tests/.
//
// Error mapping: RG_ERR_INVALID -> std::invalid_argument, other failures ->
// std::runtime_error, RG_ERR_NOT_CONVERGED -> "lm not converged!!" was already printed, pose still written.
#ifndef ROLO_B200_ROT_VGICP_HPP
#define ROLO_B200_ROT_VGICP_HPP

#include <memory>
#include <stdexcept>
#include <string>
#include <limits>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>

#include <rolo_gicp.h>
#include <rot_gicp/gicp/gicp_settings.hpp>

namespace fast_gicp {

template <typename PointSource, typename PointTarget>
class RotVGICP : public pcl::Registration<PointSource, PointTarget, float> {
public:
  using Scalar = float;
  using Base = pcl::Registration<PointSource, PointTarget, Scalar>;
  using Matrix4 = typename Base::Matrix4;
  using PointCloudSource = typename Base::PointCloudSource;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename Base::PointCloudTarget;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using CovVector = std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>;
  using Ptr = pcl::shared_ptr<RotVGICP<PointSource, PointTarget>>;
  using ConstPtr = pcl::shared_ptr<const RotVGICP<PointSource, PointTarget>>;

protected:
  using Base::reg_name_;
  using Base::input_;
  using Base::target_;
  using Base::nr_iterations_;
  using Base::max_iterations_;
  using Base::final_transformation_;
  using Base::transformation_epsilon_;
  using Base::converged_;

public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  RotVGICP() {
    reg_name_ = "RotVGICP";
    check(rg_create(-1, &ctx_), "rg_create");
    max_iterations_ = 64;              // lsq_registration_impl.hpp:11-13
    transformation_epsilon_ = 5e-4;
    final_hessian_.setIdentity();
  }
  ~RotVGICP() override { rg_destroy(ctx_); }
  RotVGICP(const RotVGICP&) = delete;
  RotVGICP& operator=(const RotVGICP&) = delete;

  // ---- RotVGICP setters (rot_vgicp.hpp:75-87) ----
  void setResolution(double resolution) { check(rg_set_resolution(ctx_, resolution), "setResolution"); }
  void setPolarResolution(double theta_res, double phi_res, double r_res) { check(rg_set_polar_resolution(ctx_, theta_res, phi_res, r_res), "setPolarResolution"); }
  void setVoxelAccumulationMode(VoxelAccumulationMode mode) { check(rg_set_voxel_accumulation_mode(ctx_, static_cast<int>(mode)), "setVoxelAccumulationMode"); }
  void setNeighborSearchMethod(NeighborSearchMethod method) { check(rg_set_neighbor_search_method(ctx_, static_cast<int>(method)), "setNeighborSearchMethod"); }
  void setNumThreads(int n) { rg_set_num_threads(ctx_, n); }
  void setCorrespondenceRandomness(int k) { check(rg_set_correspondence_randomness(ctx_, k), "setCorrespondenceRandomness"); }
  void setRegularizationMethod(RegularizationMethod method) { check(rg_set_regularization_method(ctx_, static_cast<int>(method)), "setRegularizationMethod"); }

  // ---- LsqRegistration setters (lsq_registration.hpp:51-62) ----
  void setRotationEpsilon(double eps) { check(rg_set_rotation_epsilon(ctx_, eps), "setRotationEpsilon"); }
  void setInitialLambdaFactor(double f) { check(rg_set_initial_lambda_factor(ctx_, f), "setInitialLambdaFactor"); }
  void setDebugPrint(bool on) { check(rg_set_debug_print(ctx_, on ? 1 : 0), "setDebugPrint"); }
  void setOptimizerType(LSQ_OPTIMIZER_TYPE t) { check(rg_set_optimizer_type(ctx_, static_cast<int>(t)), "setOptimizerType"); }
  const Eigen::Matrix<double, 6, 6>& getFinalHessian() const { return final_hessian_; }

  // pcl::Registration::getFitnessScore(max_range) (used by the loop-closure check, backMapping.cpp:2351,2428): mean squared
  // distance source -> nearest target under final_transformation_, on the device (rg_get_fitness_score)
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    float Tr[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Tr[i * 4 + j] = final_transformation_(i, j);
    double score = 0.0;
    check(rg_get_fitness_score(ctx_, Tr, max_range, &score), "getFitnessScore");
    return score;
  }

  double evaluateCost(const Eigen::Matrix4f& relative_pose, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) {
    double T[16], Hr[36], br[6], cost = 0.0;
    to_row_major(relative_pose.cast<double>(), T);
    const bool want = H && b;
    check(rg_linearize(ctx_, T, want ? Hr : nullptr, want ? br : nullptr, &cost), "evaluateCost");
    if (want) { for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) (*H)(i, j) = Hr[i * 6 + j]; (*b)(i) = br[i]; } }
    return cost;
  }

  // ---- clouds (rot_vgicp_impl.hpp:99-143) ----
  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    if (input_ == cloud) return;
    Base::setInputSource(cloud);
    check(rg_set_input_source(ctx_, reinterpret_cast<const float*>(cloud->points.data()), cloud->size(), sizeof(PointSource)), "setInputSource");
    source_covs_.clear();
  }
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    if (target_ == cloud) return;
    Base::setInputTarget(cloud);
    check(rg_set_input_target(ctx_, reinterpret_cast<const float*>(cloud->points.data()), cloud->size(), sizeof(PointTarget)), "setInputTarget");
    target_covs_.clear();
  }
  virtual void clearSource() { input_.reset(); source_covs_.clear(); rg_clear_source(ctx_); }
  virtual void clearTarget() { target_.reset(); target_covs_.clear(); rg_clear_target(ctx_); }
  virtual void swapSourceAndTarget() {
    input_.swap(target_); source_covs_.swap(target_covs_);
    check(rg_swap_source_and_target(ctx_), "swapSourceAndTarget");
  }
  virtual void setSourceCovariances(const CovVector& covs) { source_covs_ = covs; push_covs(RG_SOURCE, covs); }
  virtual void setTargetCovariances(const CovVector& covs) { target_covs_ = covs; push_covs(RG_TARGET, covs); }
  const CovVector& getSourceCovariances() const { pull_covs(RG_SOURCE, source_covs_); return source_covs_; }
  const CovVector& getTargetCovariances() const { pull_covs(RG_TARGET, target_covs_); return target_covs_; }

  // ---- translation stage (rot_vgicp.hpp:102-104) ----
  virtual void computeTranslation(PointCloudSource& output, Eigen::Vector3d& trans, const Eigen::Vector3d& init_guess, const Eigen::Vector3d& last_t0,
                                  const double interval_tn, const double interval_tn_1, const float ct_lambda) {
    sync_settings();
    double t[3] = {trans[0], trans[1], trans[2]};
    const int rc = rg_compute_translation(ctx_, t, init_guess.data(), last_t0.data(), interval_tn, interval_tn_1, ct_lambda);
    if (rc != RG_OK && rc != RG_ERR_NOT_CONVERGED) fail(rc, "computeTranslation");
    trans = Eigen::Vector3d(t[0], t[1], t[2]);
    nr_iterations_ = rg_nr_iterations(ctx_);
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    T.block<3, 1>(0, 3) = trans.cast<float>();
    write_output(output, T);                                  // lsq_registration_impl.hpp:75-78
  }

protected:
  // pcl::Registration::align() calls this (rot_vgicp_impl.hpp:146-160 -> lsq_registration_impl.hpp:152-179)
  void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {
    if (output.points.data() == input_->points.data() || output.points.data() == target_->points.data())
      throw std::invalid_argument("RotVGICP: destination cloud cannot be identical to source or target");
    sync_settings();
    float g[16], out[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) g[i * 4 + j] = guess(i, j);
    const int rc = rg_align(ctx_, g, out);
    if (rc != RG_OK && rc != RG_ERR_NOT_CONVERGED) fail(rc, "align");
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) final_transformation_(i, j) = out[i * 4 + j];
    converged_ = rg_has_converged(ctx_) != 0;
    nr_iterations_ = rg_nr_iterations(ctx_);
    double H[36];
    rg_get_final_hessian(ctx_, H);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) final_hessian_(i, j) = H[i * 6 + j];
    write_output(output, final_transformation_);              // lsq_registration_impl.hpp:178
  }

private:
  static void to_row_major(const Eigen::Matrix4d& M, double* T) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T[i * 4 + j] = M(i, j); }
  void sync_settings() {   // pcl::Registration setters write base-class members; forward them before every solve
    rg_set_maximum_iterations(ctx_, max_iterations_);
    rg_set_transformation_epsilon(ctx_, transformation_epsilon_);
  }
  void write_output(PointCloudSource& output, const Eigen::Matrix4f& T) {
    output = *input_;                                          // header, width/height, non-xyz fields
    float Tr[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Tr[i * 4 + j] = T(i, j);
    check(rg_transform_source(ctx_, Tr, reinterpret_cast<float*>(output.points.data()), sizeof(PointSource)), "transformPointCloud");
  }
  void push_covs(int which, const CovVector& covs) {
    std::vector<double> rm(covs.size() * 16);                  // Eigen is column-major, the ABI row-major
    for (size_t k = 0; k < covs.size(); k++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) rm[k * 16 + i * 4 + j] = covs[k](i, j);
    check(rg_set_covariances(ctx_, which, rm.data(), covs.size()), "setCovariances");
  }
  void pull_covs(int which, CovVector& covs) const {
    const size_t n = rg_get_covariances(ctx_, which, nullptr);
    if (n == 0 || covs.size() == n) return;
    std::vector<double> rm(n * 16);
    rg_get_covariances(ctx_, which, rm.data());
    covs.resize(n);
    for (size_t k = 0; k < n; k++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) covs[k](i, j) = rm[k * 16 + i * 4 + j];
  }
  [[noreturn]] void fail(int rc, const char* what) const {
    const std::string msg = std::string("RotVGICP::") + what + ": " + rg_last_error(ctx_);
    if (rc == RG_ERR_INVALID) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }
  void check(int rc, const char* what) const { if (rc != RG_OK) fail(rc, what); }

  rg_ctx* ctx_ = nullptr;
  mutable CovVector source_covs_, target_covs_;
  Eigen::Matrix<double, 6, 6> final_hessian_;
};

}  // namespace fast_gicp

#endif  // ROLO_B200_ROT_VGICP_HPP
