// ref_driver.cpp — C API around the REFERENCE's own rot_gicp sources, compiled from where they lie
// (-I /root/reference/include; nothing is copied) against oracle/ref_shim (minimal Eigen/PCL/Boost stand-ins, because
// the image has none of them). Output: oracle/_ref/librolo_ref.so — test infrastructure, used only to validate the
// oracle restatement (tests/test_oracle_vs_ref_cpu.py) and, optionally, as a CPU baseline.
// The class below only exposes protected members of fast_gicp::RotVGICP; all logic is the reference's.
#include <rot_gicp/gicp/rot_vgicp.hpp>
#include <rot_gicp/gicp/impl/rot_vgicp_impl.hpp>
#include <rot_gicp/gicp/impl/lsq_registration_impl.hpp>

#include <cstring>

using PointT = pcl::PointXYZI;

struct RefReg : public fast_gicp::RotVGICP<PointT, PointT> {
  using Base = fast_gicp::RotVGICP<PointT, PointT>;
  pcl::PointCloud<PointT>::Ptr src{new pcl::PointCloud<PointT>}, tgt{new pcl::PointCloud<PointT>};
  pcl::PointCloud<PointT> aligned;
  double lin(const Eigen::Isometry3d& T, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) { return this->linearize(T, H, b); }
  double so3lin(const Eigen::Isometry3d& T, Eigen::Matrix<double, 3, 3>* H, Eigen::Matrix<double, 3, 1>* b) { return this->so3_linearize(T, H, b); }
  double err(const Eigen::Isometry3d& T) { return this->compute_error(T); }
  double t3lin(const Eigen::Vector3d& t, const Eigen::Vector3d& g, const Eigen::Vector3d& l, double a, double c, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) {
    return this->t3_linearize(t, g, l, a, c, H, b);
  }
  double terr(const Eigen::Vector3d& t, const Eigen::Vector3d& g, const Eigen::Vector3d& l, double a, double c) { return this->compute_t_error(t, g, l, a, c); }
  int calc_cov(int which) {
    return which == 0 ? this->template calculate_covariances<PointT>(this->input_, *this->search_source_, this->source_covs_)
                      : this->template calculate_covariances<PointT>(this->target_, *this->search_target_, this->target_covs_);
  }
  void set_lambda(float l) { this->lambda_ = l; }
  int iters() const { return this->nr_iterations_; }
  size_t ncorr() const { return this->voxel_correspondences_.size(); }
};

static Eigen::Isometry3d to_iso(const double* T) {
  Eigen::Isometry3d x = Eigen::Isometry3d::Identity();
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) x(i, j) = T[i * 4 + j];
  return x;
}
static void load(pcl::PointCloud<PointT>& c, const float* pts, size_t n, size_t stride) {
  c.resize(n);
  for (size_t i = 0; i < n; i++) { c.points[i].x = pts[i * stride]; c.points[i].y = pts[i * stride + 1]; c.points[i].z = pts[i * stride + 2]; c.points[i].data[3] = 1.0f; }
}

extern "C" {
RefReg* rr_create() { return new RefReg; }
void rr_destroy(RefReg* r) { delete r; }
void rr_set_num_threads(RefReg* r, int n) { r->setNumThreads(n); }
void rr_set_resolution(RefReg* r, double v) { r->setResolution(v); }
void rr_set_polar_resolution(RefReg* r, double a, double b, double c) { r->setPolarResolution(a, b, c); }
void rr_set_neighbor_search_method(RefReg* r, int m) { r->setNeighborSearchMethod(static_cast<fast_gicp::NeighborSearchMethod>(m)); }
void rr_set_regularization_method(RefReg* r, int m) { r->setRegularizationMethod(static_cast<fast_gicp::RegularizationMethod>(m)); }
void rr_set_correspondence_randomness(RefReg* r, int k) { r->setCorrespondenceRandomness(k); }
void rr_set_optimizer_type(RefReg* r, int t) { r->setOptimizerType(static_cast<fast_gicp::LSQ_OPTIMIZER_TYPE>(t)); }
void rr_set_max_iterations(RefReg* r, int n) { r->setMaximumIterations(n); }
void rr_set_rotation_epsilon(RefReg* r, double e) { r->setRotationEpsilon(e); }
void rr_set_transformation_epsilon(RefReg* r, double e) { r->setTransformationEpsilon(e); }
void rr_set_input_source(RefReg* r, const float* p, size_t n, size_t stride) {
  r->src.reset(new pcl::PointCloud<PointT>); load(*r->src, p, n, stride); r->setInputSource(r->src);
}
void rr_set_input_target(RefReg* r, const float* p, size_t n, size_t stride) {
  r->tgt.reset(new pcl::PointCloud<PointT>); load(*r->tgt, p, n, stride); r->setInputTarget(r->tgt);
}
int rr_calculate_covariances(RefReg* r, int which) { return r->calc_cov(which) ? 0 : 1; }
size_t rr_get_covariances(RefReg* r, int which, double* out) {
  const auto& v = which == 0 ? r->getSourceCovariances() : r->getTargetCovariances();
  if (out) for (size_t k = 0; k < v.size(); k++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[k * 16 + i * 4 + j] = v[k](i, j);
  return v.size();
}
void rr_set_covariances(RefReg* r, int which, const double* in, size_t n) {
  std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> v(n);
  for (size_t k = 0; k < n; k++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) v[k](i, j) = in[k * 16 + i * 4 + j];
  if (which == 0) r->setSourceCovariances(v); else r->setTargetCovariances(v);
}
double rr_linearize(RefReg* r, const double* T, double* H, double* b) {
  Eigen::Matrix<double, 6, 6> Hm; Eigen::Matrix<double, 6, 1> bm;
  double c = r->lin(to_iso(T), H ? &Hm : nullptr, b ? &bm : nullptr);
  if (H && b) { for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) H[i * 6 + j] = Hm(i, j); b[i] = bm(i); } }
  return c;
}
double rr_so3_linearize(RefReg* r, const double* T, double* H, double* b) {
  Eigen::Matrix<double, 3, 3> Hm; Eigen::Matrix<double, 3, 1> bm;
  double c = r->so3lin(to_iso(T), &Hm, &bm);
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) H[i * 3 + j] = Hm(i, j); b[i] = bm(i); }
  return c;
}
double rr_compute_error(RefReg* r, const double* T) { return r->err(to_iso(T)); }
double rr_t3_linearize(RefReg* r, const double* t, const double* g, const double* l, double a, double c, double* H, double* b) {
  Eigen::Matrix<double, 6, 6> Hm; Eigen::Matrix<double, 6, 1> bm;
  double v = r->t3lin(Eigen::Vector3d(t[0], t[1], t[2]), Eigen::Vector3d(g[0], g[1], g[2]), Eigen::Vector3d(l[0], l[1], l[2]), a, c, &Hm, &bm);
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) H[i * 6 + j] = Hm(i, j); b[i] = bm(i); }
  return v;
}
double rr_compute_t_error(RefReg* r, const double* t, const double* g, const double* l, double a, double c) {
  return r->terr(Eigen::Vector3d(t[0], t[1], t[2]), Eigen::Vector3d(g[0], g[1], g[2]), Eigen::Vector3d(l[0], l[1], l[2]), a, c);
}
void rr_set_ct_lambda(RefReg* r, float l) { r->set_lambda(l); }
size_t rr_num_correspondences(RefReg* r) { return r->ncorr(); }
int rr_align(RefReg* r, const float* guess, float* out) {
  Eigen::Matrix4f g = Eigen::Matrix4f::Identity();
  if (guess) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) g(i, j) = guess[i * 4 + j];
  r->align(r->aligned, g);
  Eigen::Matrix4f T = r->getFinalTransformation();
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[i * 4 + j] = T(i, j);
  return 0;
}
int rr_compute_translation(RefReg* r, double* trans, const double* g, const double* l, double a, double c, float lam) {
  Eigen::Vector3d t(trans[0], trans[1], trans[2]);
  r->computeTranslation(r->aligned, t, Eigen::Vector3d(g[0], g[1], g[2]), Eigen::Vector3d(l[0], l[1], l[2]), a, c, lam);
  trans[0] = t(0); trans[1] = t(1); trans[2] = t(2);
  return 0;
}
int rr_has_converged(RefReg* r) { return r->hasConverged() ? 1 : 0; }
int rr_nr_iterations(RefReg* r) { return r->iters(); }
void rr_get_final_hessian(RefReg* r, double* H) { const auto& Hm = r->getFinalHessian(); for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H[i * 6 + j] = Hm(i, j); }
}
