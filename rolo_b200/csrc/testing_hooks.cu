// testing_hooks.cu — exports host-side logic for the CPU-only tests (include/rolo_gicp_testing.h). No CUDA calls.
#include "../../include/rolo_gicp.h"
#include "../../include/rolo_gicp_testing.h"
#include "linearize.cuh"
#include "lsq_host.hpp"
#include "s2m_math.cuh"

using namespace rolo;

// ---- the optimiser alone as a public C-ABI (include/rolo_gicp.h: rg_lsq_*): fast_gicp::LsqRegistration drives it ----
struct rg_lsq { host::LsqDriver d; };
struct rgt_lsq { rg_lsq q; };

extern "C" {

rg_lsq* rg_lsq_create(void) { return new rg_lsq; }
void rg_lsq_destroy(rg_lsq* p) { delete p; }
void rg_lsq_configure(rg_lsq* p, int optimizer, int max_iterations, double rotation_epsilon, double transformation_epsilon,
                      double lm_init_lambda_factor, int debug_print) {
  if (!p) return;
  if (optimizer < host::OPT_GAUSS_NEWTON || optimizer > host::OPT_SO3_LM) optimizer = host::OPT_LM;   // lsq.hpp:204 default branch
  p->d.cfg.optimizer = optimizer; p->d.cfg.max_iterations = max_iterations;
  p->d.cfg.rotation_epsilon = rotation_epsilon; p->d.cfg.transformation_epsilon = transformation_epsilon;
  p->d.cfg.lm_init_lambda_factor = lm_init_lambda_factor; p->d.cfg.debug_print = debug_print != 0;
}
void rg_lsq_begin_align(rg_lsq* p, const double guess[16]) { p->d.begin_align(guess); }
void rg_lsq_begin_translation(rg_lsq* p, const double trans[3]) { p->d.begin_translation(trans); }
int rg_lsq_want(rg_lsq* p, double pose16[16], double t3[3]) {
  host::Request r;
  if (!p->d.want(r)) return host::REQ_NONE;
  if (r.kind == host::REQ_T3_LIN || r.kind == host::REQ_T_ERROR) { for (int i = 0; i < 3; i++) t3[i] = r.t[i]; }
  else { for (int i = 0; i < 16; i++) pose16[i] = r.pose[i]; }
  return r.kind;
}
void rg_lsq_feed(rg_lsq* p, const double* H, const double* b, double cost) { p->d.feed(H, b, cost); }
void rg_lsq_state(const rg_lsq* p, double pose16[16], double t3[3], int* nr_iterations, int* converged, int* failed, double final_hessian36[36]) {
  if (pose16) for (int i = 0; i < 16; i++) pose16[i] = p->d.x0[i];
  if (t3) for (int i = 0; i < 3; i++) t3[i] = p->d.t0[i];
  if (nr_iterations) *nr_iterations = p->d.nr_iterations;
  if (converged) *converged = p->d.converged ? 1 : 0;
  if (failed) *failed = p->d.failed ? 1 : 0;
  if (final_hessian36) for (int i = 0; i < 36; i++) final_hessian36[i] = p->d.final_hessian[i];
}

rgt_lsq* rgt_lsq_create(int optimizer, int max_iterations, double rotation_epsilon, double transformation_epsilon) {
  rgt_lsq* p = new rgt_lsq;
  p->q.d.cfg.optimizer = optimizer; p->q.d.cfg.max_iterations = max_iterations;
  p->q.d.cfg.rotation_epsilon = rotation_epsilon; p->q.d.cfg.transformation_epsilon = transformation_epsilon;
  return p;
}
void rgt_lsq_destroy(rgt_lsq* p) { delete p; }
void rgt_lsq_begin_align(rgt_lsq* p, const double guess[16]) { rg_lsq_begin_align(&p->q, guess); }
void rgt_lsq_begin_translation(rgt_lsq* p, const double trans[3]) { rg_lsq_begin_translation(&p->q, trans); }
int rgt_lsq_want(rgt_lsq* p, double pose16[16], double t3[3]) { return rg_lsq_want(&p->q, pose16, t3); }
void rgt_lsq_feed(rgt_lsq* p, const double* H, const double* b, double cost) { rg_lsq_feed(&p->q, H, b, cost); }
void rgt_lsq_state(const rgt_lsq* p, double pose16[16], double t3[3], int* nr_iterations, int* converged, int* failed, double final_hessian36[36]) {
  rg_lsq_state(&p->q, pose16, t3, nr_iterations, converged, failed, final_hessian36);
}
int rgt_build_schedule(const int* n_points, int n_pairs, int grid, int* sched) {
  std::vector<int> s; int total = 0;
  build_schedule(n_points, n_pairs, grid, s, total);
  for (size_t i = 0; i < s.size(); i++) sched[i] = s[i];
  return total;
}
void rgt_se3_exp(const double a[6], double T[16]) { host::se3_exp_mat4(a, T); }
void rgt_ldlt_solve(const double* A, const double* rhs, int n, double* x) { host::ldlt_solve(A, rhs, n, x); }

void rgt_s2m_eigen(int n, const float* A, float* W, float* V) {
  float a[36];
  for (int i = 0; i < n * n; i++) a[i] = A[i];
  if (n == 3) s2m::jacobi_eigen<3>(a, W, V); else s2m::jacobi_eigen<6>(a, W, V);
}
void rgt_s2m_qr_solve(int rows, int cols, const float* A, const float* b, float* x) {
  float a[36], bb[6];
  for (int i = 0; i < rows * cols; i++) a[i] = A[i];
  for (int i = 0; i < rows; i++) bb[i] = b[i];
  if (rows == 5 && cols == 3) s2m::qr_solve<5, 3, true>(a, bb, x); else s2m::qr_solve<6, 6, false>(a, bb, x);
}
int rgt_s2m_invert6(const float* A, float* inv) { return s2m::invert_lu<6>(A, inv) ? 1 : 0; }
void rgt_s2m_get_transformation(const float tf[6], float t12[12]) { s2m::get_transformation(tf, t12); }

}  // extern "C"
