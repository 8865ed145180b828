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
// Everything numeric happens in librolo_gicp_b200.so (installed under the reference's library name as well: librot_gicp.so).
// Class layout as in the reference: RotVGICP derives from fast_gicp::LsqRegistration (lsq_registration.hpp), which derives from
// pcl::Registration; the protected virtuals linearize / so3_linearize / compute_error / t3_linearize / compute_t_error are overridden
// here with one fused CUDA launch each, and LsqRegistration::computeTransformation drives them through the optimiser state machine.
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
#include <rot_gicp/gicp/lsq_registration.hpp>

namespace fast_gicp {

template <typename PointSource, typename PointTarget>
class RotVGICP : public LsqRegistration<PointSource, PointTarget> {
public:
  using Scalar = float;
  using Base = LsqRegistration<PointSource, PointTarget>;
  using PclBase = pcl::Registration<PointSource, PointTarget, Scalar>;
  using Matrix4 = typename PclBase::Matrix4;
  using PointCloudSource = typename PclBase::PointCloudSource;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename PclBase::PointCloudTarget;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using CovVector = std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>;
  using Ptr = pcl::shared_ptr<RotVGICP<PointSource, PointTarget>>;
  using ConstPtr = pcl::shared_ptr<const RotVGICP<PointSource, PointTarget>>;

protected:
  using PclBase::reg_name_;
  using PclBase::input_;
  using PclBase::target_;
  using PclBase::nr_iterations_;
  using PclBase::max_iterations_;
  using PclBase::final_transformation_;
  using PclBase::transformation_epsilon_;
  using PclBase::converged_;
  using Base::final_hessian_;

public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  RotVGICP() {
    reg_name_ = "RotVGICP";
    check(rg_create(-1, &ctx_), "rg_create");
    // pcl::Registration::initCompute() would rebuild its own kd-tree over the target on every align(); nothing here reads it
    // (correspondences come from the voxel map, getFitnessScore from the device index), so the rebuild is switched off
    this->setSearchMethodTarget(this->tree_, true);
  }
  ~RotVGICP() override { rg_destroy(ctx_); }

  // ---- RotVGICP setters (rot_vgicp.hpp:75-87) ----
  void setResolution(double resolution) { check(rg_set_resolution(ctx_, resolution), "setResolution"); }
  void setPolarResolution(double theta_res, double phi_res, double r_res) { check(rg_set_polar_resolution(ctx_, theta_res, phi_res, r_res), "setPolarResolution"); }
  void setVoxelAccumulationMode(VoxelAccumulationMode mode) { check(rg_set_voxel_accumulation_mode(ctx_, static_cast<int>(mode)), "setVoxelAccumulationMode"); }
  void setNeighborSearchMethod(NeighborSearchMethod method) { check(rg_set_neighbor_search_method(ctx_, static_cast<int>(method)), "setNeighborSearchMethod"); }
  void setNumThreads(int n) { rg_set_num_threads(ctx_, n); }
  void setCorrespondenceRandomness(int k) { check(rg_set_correspondence_randomness(ctx_, k), "setCorrespondenceRandomness"); }
  void setRegularizationMethod(RegularizationMethod method) { check(rg_set_regularization_method(ctx_, static_cast<int>(method)), "setRegularizationMethod"); }

  // pcl::Registration::getFitnessScore(max_range) (used by the loop-closure check, backMapping.cpp:2351,2428): mean squared
  // distance source -> nearest target under final_transformation_, on the device (rg_get_fitness_score)
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    float Tr[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Tr[i * 4 + j] = final_transformation_(i, j);
    double score = 0.0;
    check(rg_get_fitness_score(ctx_, Tr, max_range, &score), "getFitnessScore");
    return score;
  }

  // ---- clouds (rot_vgicp_impl.hpp:99-143) ----
  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    if (input_ == cloud) return;
    PclBase::setInputSource(cloud);
    check(rg_set_input_source(ctx_, reinterpret_cast<const float*>(cloud->points.data()), cloud->size(), sizeof(PointSource)), "setInputSource");
    source_covs_.clear();
  }
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    if (target_ == cloud) return;
    PclBase::setInputTarget(cloud);
    check(rg_set_input_target(ctx_, reinterpret_cast<const float*>(cloud->points.data()), cloud->size(), sizeof(PointTarget)), "setInputTarget");
    target_covs_.clear();
  }
  void clearSource() override { input_.reset(); source_covs_.clear(); rg_clear_source(ctx_); }
  void clearTarget() override { target_.reset(); target_covs_.clear(); rg_clear_target(ctx_); }
  void swapSourceAndTarget() override {
    input_.swap(target_); source_covs_.swap(target_covs_);
    check(rg_swap_source_and_target(ctx_), "swapSourceAndTarget");
  }
  virtual void setSourceCovariances(const CovVector& covs) { source_covs_ = covs; push_covs(RG_SOURCE, covs); }
  virtual void setTargetCovariances(const CovVector& covs) { target_covs_ = covs; push_covs(RG_TARGET, covs); }
  const CovVector& getSourceCovariances() const { pull_covs(RG_SOURCE, source_covs_); return source_covs_; }
  const CovVector& getTargetCovariances() const { pull_covs(RG_TARGET, target_covs_); return target_covs_; }

  // ---- translation stage (rot_vgicp.hpp:102-104; rot_vgicp_impl.hpp:162-169) ----
  void computeTranslation(PointCloudSource& output, Eigen::Vector3d& trans, const Eigen::Vector3d& init_guess, const Eigen::Vector3d& last_t0,
                          const double interval_tn, const double interval_tn_1, const float ct_lambda) override {
    check(rg_set_ct_lambda(ctx_, ct_lambda), "computeTranslation");                       // lambda_ = ct_lambda
    Base::computeTranslation(output, trans, init_guess, last_t0, interval_tn, interval_tn_1, ct_lambda);
  }

protected:
  // pcl::Registration::align() calls this (rot_vgicp_impl.hpp:146-160), then the optimiser loop of the base class
  void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {
    check(rg_reset_voxelmap(ctx_), "computeTransformation");                              // voxelmap_.reset()
    if (output.points.data() == input_->points.data() || output.points.data() == target_->points.data())
      throw std::invalid_argument("RotVGICP: destination cloud cannot be identical to source or target");
    Base::computeTransformation(output, guess);    // covariances are estimated by the first evaluation if they are missing (:153-158)
  }

  // one fused CUDA launch each (rot_vgicp_impl.hpp:224-290, 292-388, 390-417, 498-607, 609-658)
  double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) override {
    double T[16], Hr[36], br[6], cost = 0.0;
    to_row_major(trans.matrix(), T);
    const bool want = H && b;
    check(rg_linearize(ctx_, T, want ? Hr : nullptr, want ? br : nullptr, &cost), "linearize");
    if (want) for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) (*H)(i, j) = Hr[i * 6 + j]; (*b)(i) = br[i]; }
    return cost;
  }
  double so3_linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 3, 3>* H = nullptr, Eigen::Matrix<double, 3, 1>* b = nullptr) override {
    double T[16], Hr[9], br[3], cost = 0.0;
    to_row_major(trans.matrix(), T);
    const bool want = H && b;
    check(rg_so3_linearize(ctx_, T, want ? Hr : nullptr, want ? br : nullptr, &cost), "so3_linearize");
    if (want) for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) (*H)(i, j) = Hr[i * 3 + j]; (*b)(i) = br[i]; }
    return cost;
  }
  double compute_error(const Eigen::Isometry3d& trans) override {
    double T[16], cost = 0.0;
    to_row_major(trans.matrix(), T);
    check(rg_compute_error(ctx_, T, &cost), "compute_error");
    return cost;
  }
  double t3_linearize(const Eigen::Vector3d& trans, const Eigen::Vector3d& init_guess, const Eigen::Vector3d& last_t0, const double interval_tn,
                      const double interval_tn_1, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) override {
    double Hr[36], br[6], cost = 0.0;
    const double t[3] = {trans[0], trans[1], trans[2]}, g[3] = {init_guess[0], init_guess[1], init_guess[2]}, l[3] = {last_t0[0], last_t0[1], last_t0[2]};
    check(rg_t3_linearize(ctx_, t, g, l, interval_tn, interval_tn_1, Hr, br, &cost), "t3_linearize");
    if (H && b) for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) (*H)(i, j) = Hr[i * 6 + j]; (*b)(i) = br[i]; }
    return cost;
  }
  double compute_t_error(const Eigen::Vector3d& trans, const Eigen::Vector3d& init_guess, const Eigen::Vector3d& last_t0, const double& interval_tn,
                         const double& interval_tn_1) override {
    double cost = 0.0;
    const double t[3] = {trans[0], trans[1], trans[2]}, g[3] = {init_guess[0], init_guess[1], init_guess[2]}, l[3] = {last_t0[0], last_t0[1], last_t0[2]};
    check(rg_compute_t_error(ctx_, t, g, l, interval_tn, interval_tn_1, &cost), "compute_t_error");
    return cost;
  }

private:
  static void to_row_major(const Eigen::Matrix4d& M, double* T) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T[i * 4 + j] = M(i, j); }
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
};

}  // namespace fast_gicp

#endif  // ROLO_B200_ROT_VGICP_HPP
