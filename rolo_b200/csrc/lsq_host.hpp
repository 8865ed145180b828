// lsq_host.hpp — host side of the optimiser: SO(3)/SE(3) exponentials, 6x6/3x3 LDLT and the
// Gauss-Newton / Levenberg-Marquardt step logic of LsqRegistration, written as a resumable state
// machine so that one GPU launch can serve many scan pairs in lock step (the 6x6 solve stays on
// the host by design; each evaluation request is one fused-kernel launch).
//
// Behaviour follows include/rot_gicp/gicp/impl/lsq_registration_impl.hpp (reference checkout):
//   computeTransformation :152-179   step_gn :208-222   step_lm :225-270   rot_step_lm :273-324
//   is_converged :182-191  is_rot_converged :328-335  computeTranslation :55-80  step_t_optimize :84-139
//   is_t_converged :142-148
// and include/rot_gicp/so3/so3.hpp (so3_exp :58-77, se3_exp :80-104). Matrices are row-major.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>

namespace rolo {
namespace host {

inline void mat4_identity(double* T) { for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0; }
inline void mat4_mul(const double* A, const double* B, double* C) {
  double r[16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
    double s = 0.0;
    for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j];
    r[i * 4 + j] = s;
  }
  std::memcpy(C, r, sizeof(r));
}

// quaternion (w,x,y,z) of exp(omega); small-angle Taylor branch below theta^2 = 1e-10
inline void so3_exp_quat(const double w[3], double q[4]) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double im, re;
  if (th2 < 1e-10) {
    const double th4 = th2 * th2;
    im = 0.5 - 1.0 / 48.0 * th2 + 1.0 / 3840.0 * th4;
    re = 1.0 - 1.0 / 8.0 * th2 + 1.0 / 384.0 * th4;
  } else {
    const double th = std::sqrt(th2), half = 0.5 * th;
    im = std::sin(half) / th;
    re = std::cos(half);
  }
  q[0] = re; q[1] = im * w[0]; q[2] = im * w[1]; q[3] = im * w[2];
}
// rotation matrix of a (not necessarily unit) quaternion, Eigen's toRotationMatrix formula
inline void quat_to_R(const double q[4], double R[9]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
inline void so3_exp_mat4(const double w[3], double T[16]) {
  double q[4], R[9];
  so3_exp_quat(w, q); quat_to_R(q, R);
  mat4_identity(T);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j];
}
// rotation-first twist a = (omega, tau)
inline void se3_exp_mat4(const double a[6], double T[16]) {
  const double th = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  double q[4], R[9];
  so3_exp_quat(a, q); quat_to_R(q, R);
  const double O[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
  double O2[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += O[i * 3 + k] * O[k * 3 + j]; O2[i * 3 + j] = s; }
  double V[9];
  if (th < 1e-10) {
    std::memcpy(V, R, sizeof(V));
  } else {
    const double th2 = th * th;
    const double c1 = (1.0 - std::cos(th)) / th2, c2 = (th - std::sin(th)) / (th2 * th);
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
  }
  mat4_identity(T);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j];
    T[i * 4 + 3] = V[i * 3 + 0] * a[3] + V[i * 3 + 1] * a[4] + V[i * 3 + 2] * a[5];
  }
}

// LDL^T with symmetric pivoting on the largest |diagonal| (what Eigen::LDLT does), n <= 6
inline void ldlt_solve(const double* Ain, const double* rhs, int n, double* x) {
  double A[36], y[6]; int perm[6];
  for (int i = 0; i < n * n; i++) A[i] = Ain[i];
  for (int i = 0; i < n; i++) perm[i] = i;
  const double tiny = std::numeric_limits<double>::min();
  for (int k = 0; k < n; k++) {
    int piv = k; double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; i++) if (std::fabs(A[i * n + i]) > best) { best = std::fabs(A[i * n + i]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(A[k * n + j], A[piv * n + j]);
      for (int i = 0; i < n; i++) std::swap(A[i * n + k], A[i * n + piv]);
      std::swap(perm[k], perm[piv]);
    }
    const double dk = A[k * n + k];
    if (std::fabs(dk) < tiny) { for (int i = k + 1; i < n; i++) A[i * n + k] = 0; continue; }
    for (int i = k + 1; i < n; i++) A[i * n + k] /= dk;
    for (int i = k + 1; i < n; i++) for (int j = k + 1; j <= i; j++) { A[i * n + j] -= A[i * n + k] * dk * A[j * n + k]; A[j * n + i] = A[i * n + j]; }
  }
  for (int i = 0; i < n; i++) y[i] = rhs[perm[i]];
  for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) y[i] -= A[i * n + j] * y[j];
  for (int i = 0; i < n; i++) { const double d = A[i * n + i]; y[i] = (std::fabs(d) > tiny) ? y[i] / d : 0.0; }
  for (int i = n - 1; i >= 0; i--) for (int j = i + 1; j < n; j++) y[i] -= A[j * n + i] * y[j];
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
}

enum OptType { OPT_GAUSS_NEWTON = 0, OPT_LM = 1, OPT_SO3_LM = 2 };
enum ReqKind { REQ_NONE = 0, REQ_LIN6, REQ_LIN3, REQ_ERROR, REQ_T3_LIN, REQ_T_ERROR };

struct LsqConfig {
  int optimizer = OPT_SO3_LM;          // lsq.hpp:16
  int max_iterations = 64;             // lsq.hpp:11
  double rotation_epsilon = 2e-3;      // lsq.hpp:12
  double transformation_epsilon = 5e-4;// lsq.hpp:13
  int lm_max_iterations = 10;          // lsq.hpp:18
  double lm_init_lambda_factor = 1e-9; // lsq.hpp:19
  bool debug_print = false;            // lsq.hpp:17
};

struct Request {
  int kind = REQ_NONE;
  double pose[16];      // REQ_LIN6 / REQ_LIN3 / REQ_ERROR
  double t[3];          // REQ_T3_LIN / REQ_T_ERROR
};

// One registration's optimiser state. Usage: begin_*(), then loop { if (!want(req)) break; evaluate; feed(H,b,cost) }.
class LsqDriver {
 public:
  LsqConfig cfg;
  double x0[16];
  double final_hessian[36];
  double so3_final_hessian[9];
  double lm_lambda = -1.0;
  int nr_iterations = 0;
  bool converged = false;
  bool failed = false;      // "lm not converged!!" break
  double t0[3] = {0, 0, 0};
  bool t_converged = false;

  LsqDriver() {
    mat4_identity(x0);
    for (int i = 0; i < 36; i++) final_hessian[i] = (i % 7 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 9; i++) so3_final_hessian[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }

  // LsqRegistration::computeTransformation prologue (lsq.hpp:153-156)
  void begin_align(const double guess[16]) {
    std::memcpy(x0, guess, sizeof(x0));
    x0[12] = x0[13] = x0[14] = 0.0; x0[15] = 1.0;
    lm_lambda = -1.0; converged = false; failed = false;
    iter_ = 0; stage_ = STAGE_ALIGN; phase_ = PH_LINEARIZE;
    done_ = !(iter_ < cfg.max_iterations && !converged);
    nr_iterations = 0;
  }
  // LsqRegistration::computeTranslation prologue (lsq.hpp:59-61)
  void begin_translation(const double trans[3]) {
    t0[0] = trans[0]; t0[1] = trans[1]; t0[2] = trans[2];
    lm_lambda = -1.0; t_converged = false; failed = false;
    iter_ = 0; stage_ = STAGE_TRANSLATION; phase_ = PH_LINEARIZE;
    done_ = !(iter_ < cfg.max_iterations && !t_converged);
  }

  bool want(Request& r) {
    if (done_) return false;
    if (stage_ == STAGE_ALIGN) {
      nr_iterations = iter_;
      if (phase_ == PH_LINEARIZE) {
        r.kind = (cfg.optimizer == OPT_SO3_LM) ? REQ_LIN3 : REQ_LIN6;
        std::memcpy(r.pose, x0, sizeof(x0));
      } else {
        r.kind = REQ_ERROR;
        std::memcpy(r.pose, xi_, sizeof(xi_));
      }
    } else {
      nr_iterations = iter_;
      if (phase_ == PH_LINEARIZE) { r.kind = REQ_T3_LIN; std::memcpy(r.t, t0, sizeof(t0)); }
      else { r.kind = REQ_T_ERROR; std::memcpy(r.t, ti_, sizeof(ti_)); }
    }
    return true;
  }

  // H: ndof x ndof row-major, b: ndof (ignored for error requests)
  void feed(const double* H, const double* b, double cost) {
    if (stage_ == STAGE_ALIGN) feed_align(H, b, cost); else feed_translation(H, b, cost);
  }

 private:
  enum { STAGE_ALIGN, STAGE_TRANSLATION };
  enum { PH_LINEARIZE, PH_TRIAL };
  int stage_ = STAGE_ALIGN, phase_ = PH_LINEARIZE, iter_ = 0, trial_ = 0;
  bool done_ = true;
  double H_[36], b_[6], y0_ = 0, nu_ = 2.0, d_[6], delta_[16], xi_[16], ti_[3], dt_[3];

  bool is_converged(const double* delta) const {        // lsq.hpp:182-191
    double rmax = 0, tmax = 0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
      rmax = std::max(rmax, 1.0 / cfg.rotation_epsilon * std::fabs(delta[i * 4 + j] - (i == j ? 1.0 : 0.0)));
    for (int i = 0; i < 3; i++) tmax = std::max(tmax, 1.0 / cfg.transformation_epsilon * std::fabs(delta[i * 4 + 3]));
    return std::max(rmax, tmax) < 1;
  }
  bool is_rot_converged(const double* delta) const {    // lsq.hpp:328-335
    double rmax = 0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
      rmax = std::max(rmax, 1.0 / cfg.rotation_epsilon * std::fabs(delta[i * 4 + j] - (i == j ? 1.0 : 0.0)));
    return rmax < 1;
  }
  bool is_t_converged(const double* d) const {          // lsq.hpp:142-148
    double tmax = 0;
    for (int i = 0; i < 3; i++) tmax = std::max(tmax, 1.0 / cfg.transformation_epsilon * std::fabs(d[i]));
    return tmax < 1;
  }

  int ndof() const { return cfg.optimizer == OPT_SO3_LM ? 3 : 6; }

  void end_outer_align(bool ok) {
    if (!ok) { std::cerr << "lm not converged!!" << std::endl; failed = true; done_ = true; return; }   // lsq.hpp:168-171
    converged = is_converged(delta_);                                                                 // lsq.hpp:173
    iter_++;
    phase_ = PH_LINEARIZE;
    done_ = !(iter_ < cfg.max_iterations && !converged);
  }

  void prepare_trial_align() {
    const int n = ndof();
    double A[36], nb[6];
    for (int i = 0; i < n * n; i++) A[i] = H_[i];
    for (int i = 0; i < n; i++) { A[i * n + i] += lm_lambda; nb[i] = -b_[i]; }
    ldlt_solve(A, nb, n, d_);
    if (n == 6) se3_exp_mat4(d_, delta_); else so3_exp_mat4(d_, delta_);     // lsq.hpp:239 / 291-293
    mat4_mul(delta_, x0, xi_);
    phase_ = PH_TRIAL;
  }

  void feed_align(const double* H, const double* b, double cost) {
    const int n = ndof();
    if (phase_ == PH_LINEARIZE) {
      std::memcpy(H_, H, sizeof(double) * n * n);
      std::memcpy(b_, b, sizeof(double) * n);
      y0_ = cost;
      if (cfg.optimizer == OPT_GAUSS_NEWTON) {                                // lsq.hpp:208-222
        double nb[6];
        for (int i = 0; i < 6; i++) nb[i] = -b_[i];
        ldlt_solve(H_, nb, 6, d_);
        se3_exp_mat4(d_, delta_);
        mat4_mul(delta_, x0, x0);
        std::memcpy(final_hessian, H_, sizeof(final_hessian));
        end_outer_align(true);
        return;
      }
      if (lm_lambda < 0.0) {                                                  // lsq.hpp:229-231 / 281-283
        double mx = 0;
        for (int i = 0; i < n; i++) mx = std::max(mx, std::fabs(H_[i * n + i]));
        lm_lambda = cfg.lm_init_lambda_factor * mx;
      }
      nu_ = 2.0; trial_ = 0;
      if (cfg.lm_max_iterations <= 0) { end_outer_align(false); return; }
      prepare_trial_align();
      return;
    }
    // trial evaluation came back
    const double yi = cost;
    double den = 0;
    for (int i = 0; i < n; i++) den += d_[i] * (lm_lambda * d_[i] - b_[i]);
    const double rho = (y0_ - yi) / den;                                      // lsq.hpp:243 / 297
    if (cfg.debug_print) {
      if (trial_ == 0) std::printf("--- LM optimization ---\n%5s %15s %15s %15s %15s %15s %5s\n", "i", "y0", "yi", "rho", "lambda", "|delta|", "dec");
      double dn = 0; for (int i = 0; i < n; i++) dn += d_[i] * d_[i];
      std::printf("%5d %15g %15g %15g %15g %15g %5c\n", trial_, y0_, yi, rho, lm_lambda, std::sqrt(dn), rho > 0.0 ? 'x' : ' ');
    }
    if (rho < 0) {
      const bool conv = (n == 6) ? is_converged(delta_) : is_rot_converged(delta_);
      if (conv) { end_outer_align(true); return; }
      lm_lambda = nu_ * lm_lambda; nu_ = 2 * nu_;
      trial_++;
      if (trial_ >= cfg.lm_max_iterations) { end_outer_align(false); return; }
      prepare_trial_align();
      return;
    }
    std::memcpy(x0, xi_, sizeof(x0));
    lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
    if (n == 6) std::memcpy(final_hessian, H_, sizeof(final_hessian)); else std::memcpy(so3_final_hessian, H_, sizeof(so3_final_hessian));
    end_outer_align(true);
  }

  void end_outer_translation(bool ok) {
    if (!ok) { std::cerr << "lm not converged!!" << std::endl; failed = true; done_ = true; return; }   // lsq.hpp:66-69
    t_converged = is_t_converged(dt_);
    iter_++;
    phase_ = PH_LINEARIZE;
    done_ = !(iter_ < cfg.max_iterations && !t_converged);
  }
  void prepare_trial_translation() {
    double A[36], nb[6], dT[16];
    for (int i = 0; i < 36; i++) A[i] = H_[i];
    for (int i = 0; i < 6; i++) { A[i * 6 + i] += lm_lambda; nb[i] = -b_[i]; }
    ldlt_solve(A, nb, 6, d_);
    se3_exp_mat4(d_, dT);                                                     // lsq.hpp:105-106
    dt_[0] = dT[3]; dt_[1] = dT[7]; dt_[2] = dT[11];
    for (int i = 0; i < 3; i++) ti_[i] = dt_[i] + t0[i];
    phase_ = PH_TRIAL;
  }
  void feed_translation(const double* H, const double* b, double cost) {
    if (phase_ == PH_LINEARIZE) {
      std::memcpy(H_, H, sizeof(double) * 36);
      std::memcpy(b_, b, sizeof(double) * 6);
      y0_ = cost;
      if (lm_lambda < 0.0) {
        double mx = 0;
        for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H_[i * 6 + i]));
        lm_lambda = cfg.lm_init_lambda_factor * mx;
      }
      nu_ = 2.0; trial_ = 0;
      if (cfg.lm_max_iterations <= 0) { end_outer_translation(false); return; }
      prepare_trial_translation();
      return;
    }
    const double yi = cost;
    double den = 0;
    for (int i = 0; i < 6; i++) den += d_[i] * (lm_lambda * d_[i] - b_[i]);
    const double rho = (y0_ - yi) / den;                                      // lsq.hpp:112
    if (rho < 0) {
      if (is_t_converged(dt_)) { end_outer_translation(true); return; }
      lm_lambda = nu_ * lm_lambda; nu_ = 2 * nu_;
      trial_++;
      if (trial_ >= cfg.lm_max_iterations) { end_outer_translation(false); return; }
      prepare_trial_translation();
      return;
    }
    t0[0] = ti_[0]; t0[1] = ti_[1]; t0[2] = ti_[2];
    lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
    end_outer_translation(true);
  }
};

}  // namespace host
}  // namespace rolo
