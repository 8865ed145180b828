// so3.hpp — drop-in for the reference header of the same path (include/rot_gicp/so3/so3.hpp:21-104): skewd, so3_exp, se3_exp with the
// same names, argument meaning and numerical branches (quaternion exponential with a Taylor branch below theta^2 = 1e-10; SE(3)
// exponential of a rotation-first twist through the V matrix). The optimiser inside librolo_gicp_b200.so uses its own row-major
// copy of these (rolo_b200/csrc/lsq_host.hpp, checked against scipy in tests/test_oracle_cpu.py); this header is for callers that name them.
#ifndef ROLO_B200_SO3_HPP
#define ROLO_B200_SO3_HPP

#include <cmath>
#include <Eigen/Core>
#include <Eigen/Geometry>

namespace fast_gicp {

inline Eigen::Matrix3d skewd(const Eigen::Vector3d& x) {
  Eigen::Matrix3d S;
  S(0, 0) = 0.0;   S(0, 1) = -x[2]; S(0, 2) = x[1];
  S(1, 0) = x[2];  S(1, 1) = 0.0;   S(1, 2) = -x[0];
  S(2, 0) = -x[1]; S(2, 1) = x[0];  S(2, 2) = 0.0;
  return S;
}

// exp: so(3) -> unit quaternion
inline Eigen::Quaterniond so3_exp(const Eigen::Vector3d& omega) {
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = std::sqrt(theta_sq), half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  return Eigen::Quaterniond(real_factor, imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2]);
}

// exp: se(3) twist (omega, tau), rotation first -> rigid transform
inline Eigen::Isometry3d se3_exp(const Eigen::Matrix<double, 6, 1>& a) {
  const Eigen::Vector3d omega(a[0], a[1], a[2]);
  const double theta = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  const Eigen::Matrix3d R = so3_exp(omega).toRotationMatrix();
  const Eigen::Matrix3d Omega = skewd(omega);
  Eigen::Matrix3d V;
  if (theta < 1e-10) {
    V = R;
  } else {
    const double theta_sq = theta * theta;
    V = Eigen::Matrix3d::Identity() + (1.0 - std::cos(theta)) / theta_sq * Omega + (theta - std::sin(theta)) / (theta_sq * theta) * (Omega * Omega);
  }
  Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
  T.linear() = R;
  T.translation() = V * Eigen::Vector3d(a[3], a[4], a[5]);
  return T;
}

}  // namespace fast_gicp

#endif  // ROLO_B200_SO3_HPP
