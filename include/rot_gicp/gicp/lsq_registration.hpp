// lsq_registration.hpp — drop-in fast_gicp::LsqRegistration<PointSource, PointTarget> over the B200 C-ABI.
//
// Same header path, namespace, class name, public members, protected members and protected VIRTUALS as the reference
// (include/rot_gicp/gicp/lsq_registration.hpp:14-106), so code that derives from LsqRegistration or names it keeps compiling:
// a subclass implements linearize / compute_error / compute_t_error (and optionally so3_linearize / t3_linearize) and inherits
// align() / computeTranslation(). The Gauss-Newton / LM / SO3-LM / translation-LM step logic itself (LDLT solves, damping, rho
// test, convergence, the reference's quirks: lsq_registration_impl.hpp:55-335) lives in librolo_gicp_b200.so as a resumable state
// machine (rg_lsq_*, include/rolo_gicp.h); this class feeds it from the virtuals. fast_gicp::RotVGICP overrides the virtuals with
// one fused CUDA launch each.
//
// Not provided: the reference's protected step helpers (step_optimize, step_gn, step_lm, rot_step_lm, step_t_optimize) — they
// are the inside of the state machine, never called from outside lsq_registration_impl.hpp.
#ifndef ROLO_B200_LSQ_REGISTRATION_HPP
#define ROLO_B200_LSQ_REGISTRATION_HPP

#include <Eigen/Core>
#include <Eigen/Geometry>

#include <pcl/point_types.h>
#include <pcl/point_cloud.h>
#include <pcl/common/transforms.h>
#include <pcl/registration/registration.h>

#include <rolo_gicp.h>
#include <rot_gicp/gicp/gicp_settings.hpp>

namespace fast_gicp {

template <typename PointSource, typename PointTarget>
class LsqRegistration : public pcl::Registration<PointSource, PointTarget, float> {
public:
  using Scalar = float;
  using Matrix4 = typename pcl::Registration<PointSource, PointTarget, Scalar>::Matrix4;

  using PointCloudSource = typename pcl::Registration<PointSource, PointTarget, Scalar>::PointCloudSource;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;

  using PointCloudTarget = typename pcl::Registration<PointSource, PointTarget, Scalar>::PointCloudTarget;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;

  using Ptr = pcl::shared_ptr<LsqRegistration<PointSource, PointTarget>>;
  using ConstPtr = pcl::shared_ptr<const LsqRegistration<PointSource, PointTarget>>;

protected:
  using pcl::Registration<PointSource, PointTarget, Scalar>::reg_name_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::input_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::nr_iterations_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::max_iterations_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::final_transformation_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::transformation_epsilon_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::converged_;

public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  LsqRegistration() {                                        // lsq_registration_impl.hpp:9-24
    this->reg_name_ = "LsqRegistration";
    max_iterations_ = 64;
    rotation_epsilon_ = 2e-3;
    transformation_epsilon_ = 5e-4;
    lsq_optimizer_type_ = LSQ_OPTIMIZER_TYPE::SO3_LevenbergMarquardt;
    lm_debug_print_ = false;
    lm_max_iterations_ = 10;
    lm_init_lambda_factor_ = 1e-9;
    lm_lambda_ = -1.0;
    final_hessian_.setIdentity();
    so3_final_hessian_.setIdentity();
    lsq_ = rg_lsq_create();
  }
  virtual ~LsqRegistration() { rg_lsq_destroy(lsq_); }
  LsqRegistration(const LsqRegistration&) = delete;
  LsqRegistration& operator=(const LsqRegistration&) = delete;

  void setRotationEpsilon(double eps) { rotation_epsilon_ = eps; }
  void setInitialLambdaFactor(double init_lambda_factor) { lm_init_lambda_factor_ = init_lambda_factor; }
  void setDebugPrint(bool lm_debug_print) { lm_debug_print_ = lm_debug_print; }
  const Eigen::Matrix<double, 6, 6>& getFinalHessian() const { return final_hessian_; }

  double evaluateCost(const Eigen::Matrix4f& relative_pose, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) {
    return this->linearize(Eigen::Isometry3f(relative_pose).template cast<double>(), H, b);   // lsq_registration_impl.hpp:50-52
  }

  virtual void swapSourceAndTarget() {}
  virtual void clearSource() {}
  virtual void clearTarget() {}
  void setOptimizerType(LSQ_OPTIMIZER_TYPE optimizier_type_) { lsq_optimizer_type_ = optimizier_type_; }

protected:
  // lsq_registration_impl.hpp:55-80: translation-only LM on t, then output = input moved by t
  virtual void computeTranslation(PointCloudSource& output, Eigen::Vector3d& trans, const Eigen::Vector3d& init_guess, const Eigen::Vector3d& last_t0,
                                  const double interval_tn, const double interval_tn_1, const float /*ct_lambda*/) {
    configure();
    const double t_in[3] = {trans[0], trans[1], trans[2]};
    rg_lsq_begin_translation(lsq_, t_in);
    drive(&init_guess, &last_t0, interval_tn, interval_tn_1);
    double t3[3]; int it = 0, conv = 0, failed = 0;
    rg_lsq_state(lsq_, nullptr, t3, &it, &conv, &failed, nullptr);
    trans = Eigen::Vector3d(t3[0], t3[1], t3[2]);
    nr_iterations_ = it; converged_ = conv != 0;
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    T(0, 3) = (float)t3[0]; T(1, 3) = (float)t3[1]; T(2, 3) = (float)t3[2];
    pcl::transformPointCloud(*input_, output, T);
  }

  // lsq_registration_impl.hpp:152-179
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {
    configure();
    double g[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) g[i * 4 + j] = (double)guess(i, j);
    rg_lsq_begin_align(lsq_, g);
    drive(nullptr, nullptr, 0.0, 0.0);
    double x0[16], H[36]; int it = 0, conv = 0, failed = 0;
    rg_lsq_state(lsq_, x0, nullptr, &it, &conv, &failed, H);
    nr_iterations_ = it; converged_ = conv != 0;
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) final_hessian_(i, j) = H[i * 6 + j];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) final_transformation_(i, j) = (float)x0[i * 4 + j];
    pcl::transformPointCloud(*input_, output, final_transformation_);
  }

  bool is_converged(const Eigen::Isometry3d& delta) const {          // lsq_registration_impl.hpp:182-191
    const Eigen::Matrix4d D = delta.matrix();
    double r = 0.0, t = 0.0;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) r = std::max(r, std::abs(D(i, j) - (i == j ? 1.0 : 0.0)));
      t = std::max(t, std::abs(D(i, 3)));
    }
    return std::max(1.0 / rotation_epsilon_ * r, 1.0 / transformation_epsilon_ * t) < 1;
  }
  bool is_rot_converged(const Eigen::Isometry3d& delta) const {      // lsq_registration_impl.hpp:328-335
    const Eigen::Matrix4d D = delta.matrix();
    double r = 0.0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r = std::max(r, std::abs(D(i, j) - (i == j ? 1.0 : 0.0)));
    return 1.0 / rotation_epsilon_ * r < 1;
  }
  bool is_t_converged(const Eigen::Vector3d& delta) const {          // lsq_registration_impl.hpp:142-148
    return 1.0 / transformation_epsilon_ * std::max(std::abs(delta[0]), std::max(std::abs(delta[1]), std::abs(delta[2]))) < 1;
  }

  virtual double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) = 0;
  virtual double compute_error(const Eigen::Isometry3d& trans) = 0;
  virtual double compute_t_error(const Eigen::Vector3d& trans, const Eigen::Vector3d& init_guess, const Eigen::Vector3d& last_t0,
                                 const double& interval_tn, const double& interval_tn_1) = 0;
  virtual double so3_linearize(const Eigen::Isometry3d& /*trans*/, Eigen::Matrix<double, 3, 3>* /*H*/ = nullptr, Eigen::Matrix<double, 3, 1>* /*b*/ = nullptr) { return 0.0; }
  virtual double t3_linearize(const Eigen::Vector3d& /*trans*/, const Eigen::Vector3d& /*init_guess*/, const Eigen::Vector3d& /*last_t0*/,
                              const double /*interval_tn*/, const double /*interval_tn_1*/,
                              Eigen::Matrix<double, 6, 6>* /*H*/, Eigen::Matrix<double, 6, 1>* /*b*/) { return 0.0; }

private:
  void configure() {
    rg_lsq_configure(lsq_, static_cast<int>(lsq_optimizer_type_), max_iterations_, rotation_epsilon_, transformation_epsilon_, lm_init_lambda_factor_, lm_debug_print_ ? 1 : 0);
  }
  // serve the optimiser's evaluation requests from the virtuals until it is done
  void drive(const Eigen::Vector3d* init_guess, const Eigen::Vector3d* last_t0, double interval_tn, double interval_tn_1) {
    double pose[16], t3[3];
    for (;;) {
      const int kind = rg_lsq_want(lsq_, pose, t3);
      if (kind == RG_REQ_NONE) break;
      Eigen::Isometry3d X = Eigen::Isometry3d::Identity();
      if (kind == RG_REQ_LIN6 || kind == RG_REQ_LIN3 || kind == RG_REQ_ERROR)
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) X.matrix()(i, j) = pose[i * 4 + j];
      if (kind == RG_REQ_LIN6) {
        Eigen::Matrix<double, 6, 6> H; Eigen::Matrix<double, 6, 1> b;
        const double cost = this->linearize(X, &H, &b);
        double Hr[36], br[6];
        for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) Hr[i * 6 + j] = H(i, j); br[i] = b(i); }
        rg_lsq_feed(lsq_, Hr, br, cost);
      } else if (kind == RG_REQ_LIN3) {
        Eigen::Matrix<double, 3, 3> H; Eigen::Matrix<double, 3, 1> b;
        const double cost = this->so3_linearize(X, &H, &b);
        double Hr[9], br[3];
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Hr[i * 3 + j] = H(i, j); br[i] = b(i); }
        so3_final_hessian_ = H;
        rg_lsq_feed(lsq_, Hr, br, cost);
      } else if (kind == RG_REQ_ERROR) {
        rg_lsq_feed(lsq_, nullptr, nullptr, this->compute_error(X));
      } else if (kind == RG_REQ_T3_LIN) {
        Eigen::Matrix<double, 6, 6> H; Eigen::Matrix<double, 6, 1> b;
        const double cost = this->t3_linearize(Eigen::Vector3d(t3[0], t3[1], t3[2]), *init_guess, *last_t0, interval_tn, interval_tn_1, &H, &b);
        double Hr[36], br[6];
        for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) Hr[i * 6 + j] = H(i, j); br[i] = b(i); }
        rg_lsq_feed(lsq_, Hr, br, cost);
      } else {   // RG_REQ_T_ERROR
        rg_lsq_feed(lsq_, nullptr, nullptr, this->compute_t_error(Eigen::Vector3d(t3[0], t3[1], t3[2]), *init_guess, *last_t0, interval_tn, interval_tn_1));
      }
    }
  }

protected:
  double rotation_epsilon_;

  LSQ_OPTIMIZER_TYPE lsq_optimizer_type_;
  int lm_max_iterations_;
  double lm_init_lambda_factor_;
  double lm_lambda_;
  bool lm_debug_print_;

  Eigen::Matrix<double, 6, 6> final_hessian_;
  Eigen::Matrix<double, 3, 3> so3_final_hessian_;

private:
  rg_lsq* lsq_ = nullptr;
};
}  // namespace fast_gicp

#endif  // ROLO_B200_LSQ_REGISTRATION_HPP
