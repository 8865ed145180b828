// rolo_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// Structure-faithful CPU/OpenMP restatement of the reference rot_gicp hot path
// (sdwyc/ROLO @ 3920bb7). Every function cites the reference file:line it follows
// (paths relative to /root/reference). "impl.hpp" = include/rot_gicp/gicp/impl/rot_vgicp_impl.hpp,
// "lsq.hpp" = include/rot_gicp/gicp/impl/lsq_registration_impl.hpp,
// "voxel.hpp" = include/rot_gicp/gicp/vmp_voxel.hpp, "so3.hpp" = include/rot_gicp/so3/so3.hpp.
//
// PINNING: the reference ships no golden vectors for this path; this restatement is checked against the
// reference's own sources compiled here (oracle/_ref, tests/test_oracle_vs_ref_cpu.py, agreement 1e-11).
// Third-party arithmetic absent from the tree stays restated from the published algorithms: Eigen 3.3+
// JacobiSVD (two-sided Jacobi), Eigen LDLT (diagonal pivoting), Eigen fixed-size 4x4 inverse (cofactors),
// PCL/FLANN exact kNN with float L2.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
// load this library. The product (rolo_b200/) never does.
//
// Data structures deliberately mirror the reference (AoS double 4x4 covariances,
// unordered_map<Vector3i, shared_ptr<Voxel>>, per-correspondence 4x4 inverse stored in a
// vector, per-thread H/b accumulators, schedule(guided,8)) so it is also a fair timed baseline.

#include "rolo_oracle.h"

#include <omp.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------
// tiny fixed-size linear algebra (row-major)
// ---------------------------------------------------------------------------------------
struct V4 { double v[4]; };
struct M4 { double m[16]; double& operator()(int r, int c) { return m[r * 4 + c]; } double operator()(int r, int c) const { return m[r * 4 + c]; } };

inline M4 m4_zero() { M4 a; std::memset(a.m, 0, sizeof(a.m)); return a; }
inline M4 m4_identity() { M4 a = m4_zero(); a(0,0)=a(1,1)=a(2,2)=a(3,3)=1.0; return a; }
inline M4 m4_mul(const M4& a, const M4& b) {
  M4 c;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
    double s = 0.0;
    for (int k = 0; k < 4; k++) s += a(i,k) * b(k,j);
    c(i,j) = s;
  }
  return c;
}
inline M4 m4_transpose(const M4& a) { M4 t; for (int i=0;i<4;i++) for (int j=0;j<4;j++) t(i,j)=a(j,i); return t; }
inline V4 m4_mulv(const M4& a, const V4& x) {
  V4 y;
  for (int i = 0; i < 4; i++) y.v[i] = a(i,0)*x.v[0] + a(i,1)*x.v[1] + a(i,2)*x.v[2] + a(i,3)*x.v[3];
  return y;
}
inline double quad4(const V4& e, const M4& M) {  // e^T M e
  double s = 0.0;
  for (int i = 0; i < 4; i++) { double r = 0.0; for (int j = 0; j < 4; j++) r += M(i,j) * e.v[j]; s += e.v[i] * r; }
  return s;
}

// General 4x4 inverse by cofactors (Eigen's fixed-size 4x4 path, Matrix4d::inverse(), call site impl.hpp:218).
M4 m4_inverse(const M4& A) {
  const double* m = A.m;
  double inv[16];
  inv[0]  =  m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10];
  inv[4]  = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10];
  inv[8]  =  m[4]*m[9]*m[15]  - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9];
  inv[12] = -m[4]*m[9]*m[14]  + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9];
  inv[1]  = -m[1]*m[10]*m[15] + m[1]*m[11]*m[14] + m[9]*m[2]*m[15] - m[9]*m[3]*m[14] - m[13]*m[2]*m[11] + m[13]*m[3]*m[10];
  inv[5]  =  m[0]*m[10]*m[15] - m[0]*m[11]*m[14] - m[8]*m[2]*m[15] + m[8]*m[3]*m[14] + m[12]*m[2]*m[11] - m[12]*m[3]*m[10];
  inv[9]  = -m[0]*m[9]*m[15]  + m[0]*m[11]*m[13] + m[8]*m[1]*m[15] - m[8]*m[3]*m[13] - m[12]*m[1]*m[11] + m[12]*m[3]*m[9];
  inv[13] =  m[0]*m[9]*m[14]  - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9];
  inv[2]  =  m[1]*m[6]*m[15]  - m[1]*m[7]*m[14]  - m[5]*m[2]*m[15] + m[5]*m[3]*m[14] + m[13]*m[2]*m[7]  - m[13]*m[3]*m[6];
  inv[6]  = -m[0]*m[6]*m[15]  + m[0]*m[7]*m[14]  + m[4]*m[2]*m[15] - m[4]*m[3]*m[14] - m[12]*m[2]*m[7]  + m[12]*m[3]*m[6];
  inv[10] =  m[0]*m[5]*m[15]  - m[0]*m[7]*m[13]  - m[4]*m[1]*m[15] + m[4]*m[3]*m[13] + m[12]*m[1]*m[7]  - m[12]*m[3]*m[5];
  inv[14] = -m[0]*m[5]*m[14]  + m[0]*m[6]*m[13]  + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6]  + m[12]*m[2]*m[5];
  inv[3]  = -m[1]*m[6]*m[11]  + m[1]*m[7]*m[10]  + m[5]*m[2]*m[11] - m[5]*m[3]*m[10] - m[9]*m[2]*m[7]   + m[9]*m[3]*m[6];
  inv[7]  =  m[0]*m[6]*m[11]  - m[0]*m[7]*m[10]  - m[4]*m[2]*m[11] + m[4]*m[3]*m[10] + m[8]*m[2]*m[7]   - m[8]*m[3]*m[6];
  inv[11] = -m[0]*m[5]*m[11]  + m[0]*m[7]*m[9]   + m[4]*m[1]*m[11] - m[4]*m[3]*m[9]  - m[8]*m[1]*m[7]   + m[8]*m[3]*m[5];
  inv[15] =  m[0]*m[5]*m[10]  - m[0]*m[6]*m[9]   - m[4]*m[1]*m[10] + m[4]*m[2]*m[9]  + m[8]*m[1]*m[6]   - m[8]*m[2]*m[5];
  double det = m[0]*inv[0] + m[1]*inv[4] + m[2]*inv[8] + m[3]*inv[12];
  double idet = 1.0 / det;
  M4 R;
  for (int i = 0; i < 16; i++) R.m[i] = inv[i] * idet;
  return R;
}

void inv3(const double A[9], double out[9]) {
  double c00 = A[4]*A[8]-A[5]*A[7], c01 = A[5]*A[6]-A[3]*A[8], c02 = A[3]*A[7]-A[4]*A[6];
  double det = A[0]*c00 + A[1]*c01 + A[2]*c02;
  double id = 1.0/det;
  out[0]=c00*id; out[1]=(A[2]*A[7]-A[1]*A[8])*id; out[2]=(A[1]*A[5]-A[2]*A[4])*id;
  out[3]=c01*id; out[4]=(A[0]*A[8]-A[2]*A[6])*id; out[5]=(A[2]*A[3]-A[0]*A[5])*id;
  out[6]=c02*id; out[7]=(A[1]*A[6]-A[0]*A[7])*id; out[8]=(A[0]*A[4]-A[1]*A[3])*id;
}

// ---------------------------------------------------------------------------------------
// so3.hpp
// ---------------------------------------------------------------------------------------
// so3.hpp:21-31
void skewd(const double x[3], double S[9]) {
  S[0]=0; S[1]=-x[2]; S[2]=x[1];
  S[3]=x[2]; S[4]=0; S[5]=-x[0];
  S[6]=-x[1]; S[7]=x[0]; S[8]=0;
}
// so3.hpp:58-77 — quaternion (w,x,y,z), NOT normalised (Eigen ctor does not normalise)
void so3_exp(const double omega[3], double q[4]) {
  double theta_sq = omega[0]*omega[0] + omega[1]*omega[1] + omega[2]*omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    double theta = std::sqrt(theta_sq);
    double half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  q[0] = real_factor; q[1] = imag_factor*omega[0]; q[2] = imag_factor*omega[1]; q[3] = imag_factor*omega[2];
}
// Eigen::Quaterniond::toRotationMatrix (used at so3.hpp:100, lsq.hpp:293)
void quat_to_rot(const double q[4], double R[9]) {
  const double w=q[0], x=q[1], y=q[2], z=q[3];
  const double tx=2*x, ty=2*y, tz=2*z;
  const double twx=tx*w, twy=ty*w, twz=tz*w, txx=tx*x, txy=ty*x, txz=tz*x, tyy=ty*y, tyz=tz*y, tzz=tz*z;
  R[0]=1-(tyy+tzz); R[1]=txy-twz;     R[2]=txz+twy;
  R[3]=txy+twz;     R[4]=1-(txx+tzz); R[5]=tyz-twx;
  R[6]=txz-twy;     R[7]=tyz+twx;     R[8]=1-(txx+tyy);
}
// so3.hpp:80-104 — rotation-first twist, returns 4x4
M4 se3_exp(const double a[6]) {
  const double* omega = a;
  double theta = std::sqrt(omega[0]*omega[0] + omega[1]*omega[1] + omega[2]*omega[2]);
  double q[4]; so3_exp(omega, q);
  double Rq[9]; quat_to_rot(q, Rq);
  double Om[9]; skewd(omega, Om);
  double Om2[9];
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) { double s=0; for (int k=0;k<3;k++) s+=Om[i*3+k]*Om[k*3+j]; Om2[i*3+j]=s; }
  double V[9];
  if (theta < 1e-10) {
    std::memcpy(V, Rq, sizeof(V));  // so3.matrix()
  } else {
    const double theta_sq = theta*theta;
    const double c1 = (1.0 - std::cos(theta)) / theta_sq;
    const double c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i=0;i<9;i++) V[i] = ((i%4==0)?1.0:0.0) + c1*Om[i] + c2*Om2[i];
  }
  M4 T = m4_identity();
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) T(i,j)=Rq[i*3+j];
  for (int i=0;i<3;i++) T(i,3) = V[i*3+0]*a[3] + V[i*3+1]*a[4] + V[i*3+2]*a[5];
  return T;
}

// ---------------------------------------------------------------------------------------
// Eigen::JacobiSVD<Matrix3d> restated (two-sided Jacobi, Eigen/src/SVD/JacobiSVD.h), call site impl.hpp:468
// A = U diag(S) V^T, S sorted descending.
// ---------------------------------------------------------------------------------------
struct Rot { double c, s; };
inline bool make_jacobi(double x, double y, double z, Rot& r) {
  double deno = 2.0 * std::fabs(y);
  if (deno < std::numeric_limits<double>::min()) { r.c = 1; r.s = 0; return false; }
  double tau = (x - z) / deno;
  double w = std::sqrt(tau*tau + 1.0);
  double t = (tau > 0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
  double sign_t = t > 0 ? 1.0 : -1.0;
  double n = 1.0 / std::sqrt(t*t + 1.0);
  r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  r.c = n;
  return true;
}
inline void rows_rot(double* W, int n, int p, int q, Rot j) {   // applyOnTheLeft(p,q,j)
  for (int i = 0; i < n; i++) { double x = W[p*n+i], y = W[q*n+i]; W[p*n+i] = j.c*x + j.s*y; W[q*n+i] = -j.s*x + j.c*y; }
}
inline void cols_rot(double* W, int n, int p, int q, Rot j) {   // applyOnTheRight(p,q,j)
  for (int i = 0; i < n; i++) { double x = W[i*n+p], y = W[i*n+q]; W[i*n+p] = j.c*x - j.s*y; W[i*n+q] = j.s*x + j.c*y; }
}
void svd3(const double A[9], double U[9], double S[3], double V[9]) {
  const int n = 3;
  double scale = 0.0;
  for (int i = 0; i < 9; i++) scale = std::max(scale, std::fabs(A[i]));
  if (!(scale > 0.0)) scale = 1.0;
  double W[9];
  for (int i = 0; i < 9; i++) { W[i] = A[i] / scale; U[i] = V[i] = (i%4==0) ? 1.0 : 0.0; }
  const double precision = 2.0 * std::numeric_limits<double>::epsilon();
  const double considerAsZero = std::numeric_limits<double>::min();
  double maxDiag = std::max(std::fabs(W[0]), std::max(std::fabs(W[4]), std::fabs(W[8])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 100) {
    finished = true;
    for (int p = 1; p < n; p++) for (int q = 0; q < p; q++) {
      double threshold = std::max(considerAsZero, precision * maxDiag);
      if (std::fabs(W[p*n+q]) > threshold || std::fabs(W[q*n+p]) > threshold) {
        finished = false;
        // real_2x2_jacobi_svd
        double m00 = W[p*n+p], m01 = W[p*n+q], m10 = W[q*n+p], m11 = W[q*n+q];
        Rot rot1;
        double t = m00 + m11, d = m10 - m01;
        if (std::fabs(d) < std::numeric_limits<double>::min()) { rot1.s = 0; rot1.c = 1; }
        else { double u = t / d; double tmp = std::sqrt(1.0 + u*u); rot1.s = 1.0 / tmp; rot1.c = u / tmp; }
        // m.applyOnTheLeft(0,1,rot1)
        double a00 = rot1.c*m00 + rot1.s*m10, a01 = rot1.c*m01 + rot1.s*m11;
        double a11 = -rot1.s*m01 + rot1.c*m11;
        Rot jr; make_jacobi(a00, a01, a11, jr);
        // j_left = rot1 * j_right.transpose();  transpose = (c, -s)
        Rot jl; jl.c = rot1.c*jr.c - rot1.s*(-jr.s); jl.s = rot1.c*(-jr.s) + rot1.s*jr.c;
        rows_rot(W, n, p, q, jl);
        Rot jlt = { jl.c, -jl.s };
        cols_rot(U, n, p, q, jlt);
        cols_rot(W, n, p, q, jr);
        cols_rot(V, n, p, q, jr);
        maxDiag = std::max(maxDiag, std::max(std::fabs(W[p*n+p]), std::fabs(W[q*n+q])));
      }
    }
  }
  for (int i = 0; i < n; i++) {
    double a = std::fabs(W[i*n+i]);
    S[i] = a;
    if (a != 0.0 && W[i*n+i] < 0) for (int r = 0; r < n; r++) U[r*n+i] = -U[r*n+i];
  }
  for (int i = 0; i < n; i++) S[i] *= scale;
  for (int i = 0; i < n; i++) {
    int pos = i; double mx = S[i];
    for (int k = i+1; k < n; k++) if (S[k] > mx) { mx = S[k]; pos = k; }
    if (mx == 0.0) break;
    if (pos != i) {
      std::swap(S[i], S[pos]);
      for (int r = 0; r < n; r++) { std::swap(U[r*n+i], U[r*n+pos]); std::swap(V[r*n+i], V[r*n+pos]); }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Eigen::LDLT restated (diagonal pivoting). Call sites lsq.hpp:102,213,236,288. n <= 6.
// ---------------------------------------------------------------------------------------
int ldlt_solve(const double* Ain, const double* rhs, int n, double* x) {
  double A[36]; int perm[6]; double bb[6];
  for (int i = 0; i < n*n; i++) A[i] = Ain[i];
  for (int i = 0; i < n; i++) perm[i] = i;
  // in-place LDL^T on the lower triangle with symmetric pivoting on the largest |diagonal|
  for (int k = 0; k < n; k++) {
    int piv = k; double best = std::fabs(A[k*n+k]);
    for (int i = k+1; i < n; i++) if (std::fabs(A[i*n+i]) > best) { best = std::fabs(A[i*n+i]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(A[k*n+j], A[piv*n+j]);
      for (int i = 0; i < n; i++) std::swap(A[i*n+k], A[i*n+piv]);
      std::swap(perm[k], perm[piv]);
    }
    double dk = A[k*n+k];
    if (std::fabs(dk) < std::numeric_limits<double>::min()) { for (int i = k+1; i < n; i++) A[i*n+k] = 0; continue; }
    for (int i = k+1; i < n; i++) A[i*n+k] /= dk;
    for (int i = k+1; i < n; i++) for (int j = k+1; j <= i; j++) { A[i*n+j] -= A[i*n+k] * dk * A[j*n+k]; A[j*n+i] = A[i*n+j]; }
  }
  for (int i = 0; i < n; i++) bb[i] = rhs[perm[i]];
  for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) bb[i] -= A[i*n+j] * bb[j];      // L y = P b
  for (int i = 0; i < n; i++) { double d = A[i*n+i]; bb[i] = (std::fabs(d) > std::numeric_limits<double>::min()) ? bb[i] / d : 0.0; }
  for (int i = n-1; i >= 0; i--) for (int j = i+1; j < n; j++) bb[i] -= A[j*n+i] * bb[j]; // L^T z = y
  for (int i = 0; i < n; i++) x[perm[i]] = bb[i];
  return 0;
}

#include "kdtree.hpp"

// ---------------------------------------------------------------------------------------
// voxel.hpp
// ---------------------------------------------------------------------------------------
struct Key3 { int x, y, z; bool operator==(const Key3& o) const { return x==o.x && y==o.y && z==o.z; } };
// voxel.hpp:49-58 (boost::hash_combine, pre-1.81 mixing; only affects map iteration order)
struct Key3Hash {
  size_t operator()(const Key3& k) const {
    size_t seed = 0;
    auto comb = [&seed](int v) { seed ^= std::hash<int>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2); };
    comb(k.x); comb(k.y); comb(k.z);
    return seed;
  }
};
// voxel.hpp:60-108 (VmfVoxel + AdditiveVmfVoxel)
struct Voxel {
  int num_points = 0;
  V4 mean_dir{{0,0,0,0}};
  V4 dir_reg{{0,0,0,0}};
  double r_bar = 0, kappa = 0;
  M4 cov = m4_zero();
  void append(const V4& p, const M4& c) {            // voxel.hpp:93-99
    num_points++;
    for (int i = 0; i < 4; i++) mean_dir.v[i] += p.v[i];
    double nrm = std::sqrt(mean_dir.v[0]*mean_dir.v[0] + mean_dir.v[1]*mean_dir.v[1] + mean_dir.v[2]*mean_dir.v[2] + mean_dir.v[3]*mean_dir.v[3]);
    r_bar = nrm / num_points;
    for (int i = 0; i < 16; i++) cov.m[i] += c.m[i];
  }
  void finalize() {                                   // voxel.hpp:101-108
    for (int i = 0; i < 4; i++) mean_dir.v[i] /= num_points;
    double nrm = std::sqrt(mean_dir.v[0]*mean_dir.v[0] + mean_dir.v[1]*mean_dir.v[1] + mean_dir.v[2]*mean_dir.v[2] + mean_dir.v[3]*mean_dir.v[3]);
    for (int i = 0; i < 4; i++) dir_reg.v[i] = mean_dir.v[i] / nrm;
    if (r_bar < 1e-8) kappa = 0.0;
    else if (r_bar < 0.6) kappa = 3.0 * r_bar * (1.0 + 0.6 * std::pow(r_bar, 2) + 99.0 / 175.0 * std::pow(r_bar, 4));
    else kappa = r_bar * (3.0 - std::pow(r_bar, 2)) / (1.0 - std::pow(r_bar, 2));
    for (int i = 0; i < 16; i++) cov.m[i] /= num_points;
  }
};
using VoxelPtr = std::shared_ptr<Voxel>;

// voxel.hpp:13-47
std::vector<Key3> neighbor_offsets(int method) {
  switch (method) {
    default:
      std::cerr << "unsupported neighbor search method" << std::endl;
      abort();
    case RO_DIRECT1: return { {0,0,0} };
    case RO_DIRECT7: return { {0,0,0}, {1,0,0}, {-1,0,0}, {0,1,0}, {0,-1,0}, {0,0,1}, {0,0,-1} };
    case RO_DIRECT27: break;
  }
  std::vector<Key3> o;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) o.push_back({i-1, j-1, k-1});
  return o;
}

struct VoxelMap {
  double res; double pres[3]; int type;
  std::unordered_map<Key3, VoxelPtr, Key3Hash> voxels;
  // voxel.hpp:199-201
  Key3 voxel_coord(const V4& x) const {
    return { (int)std::floor(x.v[0]/res - 0.5), (int)std::floor(x.v[1]/res - 0.5), (int)std::floor(x.v[2]/res - 0.5) };
  }
  // voxel.hpp:208-211
  Key3 polar_coord(const V4& x) const {
    double r = std::sqrt(x.v[0]*x.v[0] + x.v[1]*x.v[1] + x.v[2]*x.v[2]);
    return { (int)std::floor((std::atan2(x.v[1], x.v[0]) + M_PI) / pres[0]),
             (int)std::floor(std::acos(x.v[2] / r) / pres[1]),
             (int)std::floor(r / pres[2]) };
  }
  Key3 coord(const V4& x) const { return type == RO_VOXEL_POLAR ? polar_coord(x) : voxel_coord(x); }
  // voxel.hpp:167-197 (serial)
  void create(const std::vector<float>& cloud, const std::vector<M4>& covs) {
    voxels.clear();
    size_t n = cloud.size() / 4;
    for (size_t i = 0; i < n; i++) {
      V4 p{{(double)cloud[i*4+0], (double)cloud[i*4+1], (double)cloud[i*4+2], (double)cloud[i*4+3]}};
      Key3 c = coord(p);
      auto found = voxels.find(c);
      if (found == voxels.end()) found = voxels.insert(found, std::make_pair(c, std::shared_ptr<Voxel>(new Voxel)));
      found->second->append(p, covs[i]);
    }
    for (auto& v : voxels) v.second->finalize();
  }
  // voxel.hpp:226-233 (returns a shared_ptr copy: atomic refcount traffic is part of the baseline)
  VoxelPtr lookup(const Key3& c) const {
    auto f = voxels.find(c);
    if (f == voxels.end()) return nullptr;
    return f->second;
  }
};

struct TraceRec { double d[64]; };

}  // namespace

// ---------------------------------------------------------------------------------------
// the registration object: RotVGICP + LsqRegistration + the pcl::Registration members it uses
// ---------------------------------------------------------------------------------------
struct ro_ctx {
  // impl.hpp:20-39
  int num_threads = omp_get_max_threads();
  int k_correspondences = 20;
  int regularization_method = RO_REG_PLANE;
  double voxel_resolution = 1.0;
  float lambda_ = 1.0f;
  int search_method = RO_DIRECT1;
  int voxel_type = RO_VOXEL_POLAR;
  double polar_resolution[3] = {1.0, 0.0, 0.0};   // Vector3d::Identity() (Q4)
  // lsq.hpp:9-24
  int max_iterations = 64;
  double rotation_epsilon = 2e-3;
  double transformation_epsilon = 5e-4;
  int optimizer = RO_OPT_SO3_LM;
  int lm_max_iterations = 10;
  double lm_init_lambda_factor = 1e-9;
  double lm_lambda = -1.0;
  double final_hessian[36];
  double so3_final_hessian[9];
  int nr_iterations = 0;
  bool converged = false;
  M4 final_pose_d = m4_identity();

  std::vector<float> source, target;  // xyz1
  std::vector<M4> source_covs, target_covs;
  KdTree kd_source, kd_target;
  bool kd_source_ok = false, kd_target_ok = false;
  std::unique_ptr<VoxelMap> voxelmap;
  std::vector<std::pair<int, VoxelPtr>> corr;
  std::vector<Key3> corr_keys_dbg;
  std::vector<M4> mahalanobis;
  std::vector<TraceRec> trace;
  bool tracing = false;

  ro_ctx() {
    for (int i = 0; i < 36; i++) final_hessian[i] = (i % 7 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 9; i++) so3_final_hessian[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }

  size_t nsrc() const { return source.size() / 4; }
  size_t ntgt() const { return target.size() / 4; }

  void record(int kind, double cost, const M4* pose, const double* t3, const double* H, int hn, const double* b, int bn) {
    if (!tracing) return;
    TraceRec r; std::memset(r.d, 0, sizeof(r.d));
    r.d[0] = kind; r.d[1] = cost;
    if (pose) for (int i = 0; i < 16; i++) r.d[2+i] = pose->m[i];
    if (t3) { M4 I = m4_identity(); I(0,3)=t3[0]; I(1,3)=t3[1]; I(2,3)=t3[2]; for (int i = 0; i < 16; i++) r.d[2+i] = I.m[i]; }
    if (H) for (int i = 0; i < hn; i++) for (int j = 0; j < hn; j++) r.d[18 + i*6 + j] = H[i*hn+j];
    if (b) for (int i = 0; i < bn; i++) r.d[54+i] = b[i];
    r.d[60] = lm_lambda;
    trace.push_back(r);
  }

  // impl.hpp:421-496
  int calculate_covariances(const std::vector<float>& cloud, KdTree& tree, bool& tree_ok, std::vector<M4>& covs) {
    const int n = (int)(cloud.size() / 4);
    const int k = k_correspondences;
    if (n < k) { std::cerr << "oracle: cloud smaller than k (reference reads uninitialised columns here, Q5)" << std::endl; return -1; }
    if (!tree_ok) { tree.build(cloud.data(), n); tree_ok = true; }
    covs.resize(n);
    int bad = 0;
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8) reduction(+ : bad)
    for (int i = 0; i < n; i++) {
      std::vector<int> k_indices(k);
      int found = tree.knn(&cloud[i*4], k, k_indices.data());
      if (found < k) { bad++; continue; }
      std::vector<double> nb(4 * k);
      for (int j = 0; j < k; j++) for (int d = 0; d < 4; d++) nb[d*k + j] = (double)cloud[k_indices[j]*4 + d];
      for (int d = 0; d < 4; d++) {                       // neighbors.colwise() -= rowwise().mean()
        double s = 0; for (int j = 0; j < k; j++) s += nb[d*k+j];
        double m = s / k;
        for (int j = 0; j < k; j++) nb[d*k+j] -= m;
      }
      M4 cov;
      for (int a = 0; a < 4; a++) for (int b2 = 0; b2 < 4; b2++) {
        double s = 0; for (int j = 0; j < k; j++) s += nb[a*k+j] * nb[b2*k+j];
        cov(a,b2) = s / k;
      }
      if (regularization_method == RO_REG_NONE) {
        covs[i] = cov;
      } else if (regularization_method == RO_REG_FROBENIUS) {
        double lambda = 1e-3;
        double C[9], Ci[9];
        for (int a = 0; a < 3; a++) for (int b2 = 0; b2 < 3; b2++) C[a*3+b2] = cov(a,b2) + (a==b2 ? lambda : 0.0);
        inv3(C, Ci);
        double fro = 0; for (int t = 0; t < 9; t++) fro += Ci[t]*Ci[t];
        fro = std::sqrt(fro);
        double Cn[9], out[9];
        for (int t = 0; t < 9; t++) Cn[t] = Ci[t] / fro;
        inv3(Cn, out);
        covs[i] = m4_zero();
        for (int a = 0; a < 3; a++) for (int b2 = 0; b2 < 3; b2++) covs[i](a,b2) = out[a*3+b2];
      } else {
        double A[9], U[9], S[3], V[9];
        for (int a = 0; a < 3; a++) for (int b2 = 0; b2 < 3; b2++) A[a*3+b2] = cov(a,b2);
        svd3(A, U, S, V);
        double values[3];
        switch (regularization_method) {
          default:
            std::cerr << "here must not be reached" << std::endl;
            abort();
          case RO_REG_PLANE: values[0] = 1; values[1] = 1; values[2] = 1e-3; break;
          case RO_REG_MIN_EIG: for (int t = 0; t < 3; t++) values[t] = std::max(S[t], 1e-3); break;
          case RO_REG_NORMALIZED_MIN_EIG: {
            double mx = std::max(S[0], std::max(S[1], S[2]));
            for (int t = 0; t < 3; t++) values[t] = std::max(S[t] / mx, 1e-3);
          } break;
          case RO_REG_PLANE_S: {
            double sum = S[0] + S[1] + S[2];
            for (int t = 0; t < 3; t++) values[t] = S[t] / sum;
            values[2] = 1e-3;
          } break;
        }
        covs[i] = m4_zero();
        for (int a = 0; a < 3; a++) for (int b2 = 0; b2 < 3; b2++) {
          double s = 0; for (int t = 0; t < 3; t++) s += U[a*3+t] * values[t] * V[b2*3+t];
          covs[i](a,b2) = s;
        }
      }
    }
    return bad ? -2 : 0;
  }

  void ensure_voxelmap() {   // impl.hpp:227-230 / 295-298
    if (!voxelmap) {
      voxelmap.reset(new VoxelMap);
      voxelmap->res = voxel_resolution;
      for (int i = 0; i < 3; i++) voxelmap->pres[i] = polar_resolution[i];
      voxelmap->type = voxel_type;
      voxelmap->create(target, target_covs);
    }
  }

  // impl.hpp:172-222
  void update_correspondences(const M4& trans) {
    corr.clear();
    corr_keys_dbg.clear();
    auto offsets = neighbor_offsets(search_method);
    std::vector<std::vector<std::pair<int, VoxelPtr>>> corrs(num_threads);
    std::vector<std::vector<Key3>> ckeys(num_threads);
    for (auto& c : corrs) c.reserve((nsrc() * offsets.size()) / num_threads);
    const int n = (int)nsrc();
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      V4 mean_A{{(double)source[i*4+0], (double)source[i*4+1], (double)source[i*4+2], (double)source[i*4+3]}};
      V4 transed = m4_mulv(trans, mean_A);
      Key3 coord = voxelmap->coord(transed);
      for (const auto& off : offsets) {
        Key3 c{coord.x+off.x, coord.y+off.y, coord.z+off.z};
        auto voxel = voxelmap->lookup(c);
        if (voxel != nullptr) {
          corrs[omp_get_thread_num()].push_back(std::make_pair(i, voxel));
          ckeys[omp_get_thread_num()].push_back(c);
        }
      }
    }
    corr.reserve(nsrc() * offsets.size());
    for (size_t t = 0; t < corrs.size(); t++) {
      corr.insert(corr.end(), corrs[t].begin(), corrs[t].end());
      corr_keys_dbg.insert(corr_keys_dbg.end(), ckeys[t].begin(), ckeys[t].end());
    }
    mahalanobis.resize(corr.size());
    const M4 Tt = m4_transpose(trans);
    const int m = (int)corr.size();
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
    for (int i = 0; i < m; i++) {
      const auto& c = corr[i];
      const M4& cov_A = source_covs[c.first];
      const M4& cov_B = c.second->cov;
      M4 RCR = m4_mul(m4_mul(trans, cov_A), Tt);
      for (int t = 0; t < 16; t++) RCR.m[t] += cov_B.m[t];
      RCR(3,3) = 1.0;
      mahalanobis[i] = m4_inverse(RCR);
      mahalanobis[i](3,3) = 0.0;
    }
  }

  // impl.hpp:224-290 (ndof=6) and impl.hpp:292-388 (ndof=3)
  double linearize_n(const M4& trans, int ndof, double* H, double* b) {
    ensure_voxelmap();
    update_correspondences(trans);
    double sum_errors = 0.0;
    const int hn = ndof * ndof;
    std::vector<double> Hs((size_t)num_threads * 36, 0.0), bs((size_t)num_threads * 6, 0.0);
    const int m = (int)corr.size();
#pragma omp parallel for num_threads(num_threads) reduction(+ : sum_errors) schedule(guided, 8)
    for (int i = 0; i < m; i++) {
      const auto& c = corr[i];
      auto target_voxel = c.second;
      V4 mean_A{{(double)source[c.first*4+0], (double)source[c.first*4+1], (double)source[c.first*4+2], (double)source[c.first*4+3]}};
      const V4& mean_B = c.second->mean_dir;
      V4 transed = m4_mulv(trans, mean_A);
      V4 error; for (int t = 0; t < 4; t++) error.v[t] = mean_B.v[t] - transed.v[t];
      double w = std::sqrt((double)target_voxel->num_points);
      const M4& Mh = mahalanobis[i];
      sum_errors += w * quad4(error, Mh);
      if (H == nullptr || b == nullptr) continue;
      // J (4 x ndof): [skew(p') | -I] (impl.hpp:265-267) or skew(p') (impl.hpp:345-347)
      double J[4*6]; std::memset(J, 0, sizeof(J));
      double S[9]; skewd(transed.v, S);
      for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) J[r*6+cc] = S[r*3+cc];
      if (ndof == 6) for (int r = 0; r < 3; r++) J[r*6+3+r] = -1.0;
      // MJ = M * J ; Hi = w J^T M J ; bi = w J^T M e
      double MJ[4*6];
      for (int r = 0; r < 4; r++) for (int cc = 0; cc < ndof; cc++) { double s = 0; for (int t = 0; t < 4; t++) s += Mh(r,t) * J[t*6+cc]; MJ[r*6+cc] = s; }
      double Me[4];
      for (int r = 0; r < 4; r++) { double s = 0; for (int t = 0; t < 4; t++) s += Mh(r,t) * error.v[t]; Me[r] = s; }
      double* Ht = &Hs[(size_t)omp_get_thread_num() * 36];
      double* bt = &bs[(size_t)omp_get_thread_num() * 6];
      for (int r = 0; r < ndof; r++) {
        for (int cc = 0; cc < ndof; cc++) { double s = 0; for (int t = 0; t < 4; t++) s += J[t*6+r] * MJ[t*6+cc]; Ht[r*ndof+cc] += w * s; }
        double s = 0; for (int t = 0; t < 4; t++) s += J[t*6+r] * Me[t];
        bt[r] += w * s;
      }
    }
    if (H && b) {
      for (int t = 0; t < hn; t++) H[t] = 0; for (int t = 0; t < ndof; t++) b[t] = 0;
      for (int th = 0; th < num_threads; th++) {
        for (int t = 0; t < hn; t++) H[t] += Hs[(size_t)th*36 + t];
        for (int t = 0; t < ndof; t++) b[t] += bs[(size_t)th*6 + t];
      }
    }
    record(ndof == 6 ? 0 : 1, sum_errors, &trans, nullptr, H, ndof, b, ndof);
    return sum_errors;
  }

  // impl.hpp:390-417
  double compute_error(const M4& trans) {
    double sum_errors = 0.0;
    const int m = (int)corr.size();
#pragma omp parallel for num_threads(num_threads) reduction(+ : sum_errors)
    for (int i = 0; i < m; i++) {
      const auto& c = corr[i];
      auto target_voxel = c.second;
      V4 mean_A{{(double)source[c.first*4+0], (double)source[c.first*4+1], (double)source[c.first*4+2], (double)source[c.first*4+3]}};
      const V4& mean_B = c.second->mean_dir;
      V4 transed = m4_mulv(trans, mean_A);
      V4 error; for (int t = 0; t < 4; t++) error.v[t] = mean_B.v[t] - transed.v[t];
      double w = std::sqrt((double)target_voxel->num_points);
      sum_errors += w * quad4(error, mahalanobis[i]);
    }
    record(2, sum_errors, &trans, nullptr, nullptr, 0, nullptr, 0);
    return sum_errors;
  }

  // shared per-correspondence terms of impl.hpp:523-556 / 621-654.
  // Q3: `last_transform.matrix().col(3).head<3>() = last_t0` (impl.hpp:540, 638) addresses column 3 of a
  // 4x1 vector: an out-of-bounds write that never lands in last_transform. Restated as a no-op, so
  // last_transform = Zero in t3_linearize (impl.hpp:539) and Vector4d::Identity() = (1,0,0,0) in
  // compute_t_error (impl.hpp:637).
  double t3_terms(int i, const double t[3], const double g[3], const double last_tr[4], double dtn, double dtn1,
                  size_t pt_size, V4* transed_out, V4* error_out, V4* ct_out) const {
    const auto& c = corr[i];
    V4 mean_A{{(double)source[c.first*4+0], (double)source[c.first*4+1], (double)source[c.first*4+2], (double)source[c.first*4+3]}};
    const V4& mean_B = c.second->mean_dir;
    V4 transed{{mean_A.v[0] + t[0]*mean_A.v[3], mean_A.v[1] + t[1]*mean_A.v[3], mean_A.v[2] + t[2]*mean_A.v[3], mean_A.v[3]}};
    // propagation_transform.inverse() * mean_A (pure translation g)
    V4 begin{{mean_A.v[0] - g[0]*mean_A.v[3], mean_A.v[1] - g[1]*mean_A.v[3], mean_A.v[2] - g[2]*mean_A.v[3], mean_A.v[3]}};
    V4 error, ct;
    for (int k = 0; k < 4; k++) {
      error.v[k] = mean_B.v[k] - transed.v[k];
      ct.v[k] = (begin.v[k] - transed.v[k]) / dtn - last_tr[k] / dtn1;
    }
    double w = std::sqrt((double)c.second->num_points);
    const M4& Mh = mahalanobis[i];
    // `lambda_/pt_size` is float / size_t: a FLOAT division in the reference (impl.hpp:555,653), promoted afterwards
    const double lam = (double)(lambda_ / (float)pt_size);
    double val = w * (quad4(error, Mh) + lam * quad4(ct, Mh));
    if (transed_out) *transed_out = transed;
    if (error_out) *error_out = error;
    if (ct_out) *ct_out = ct;
    return val;
  }

  // impl.hpp:498-607
  double t3_linearize(const double t[3], const double g[3], const double /*last_t0*/[3], double dtn, double dtn1, double* H, double* b) {
    double sum_errors = 0.0;
    std::vector<double> Hs((size_t)num_threads * 36, 0.0), bs((size_t)num_threads * 6, 0.0);
    const size_t pt_size = corr.size();
    const int m = (int)corr.size();
    const double last_tr[4] = {0, 0, 0, 0};
#pragma omp parallel for num_threads(num_threads) reduction(+ : sum_errors) schedule(guided, 8)
    for (int i = 0; i < m; i++) {
      V4 transed, error, ct;
      sum_errors += t3_terms(i, t, g, last_tr, dtn, dtn1, pt_size, &transed, &error, &ct);
      if (H == nullptr || b == nullptr) continue;
      const M4& Mh = mahalanobis[i];
      double w = std::sqrt((double)corr[i].second->num_points);
      double J1[4*6]; std::memset(J1, 0, sizeof(J1));
      double S[9]; skewd(transed.v, S);
      for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) J1[r*6+cc] = S[r*3+cc];
      for (int r = 0; r < 3; r++) J1[r*6+3+r] = -1.0;
      double J2[4*6]; for (int k = 0; k < 24; k++) J2[k] = 1.0 / dtn * J1[k];     // impl.hpp:579
      const double lam = (double)(lambda_ / (float)pt_size);   // float division, impl.hpp:582,584
      double MJ1[24], MJ2[24], Me[4], Mct[4];
      for (int r = 0; r < 4; r++) for (int cc = 0; cc < 6; cc++) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 4; k++) { s1 += Mh(r,k) * J1[k*6+cc]; s2 += Mh(r,k) * J2[k*6+cc]; }
        MJ1[r*6+cc] = s1; MJ2[r*6+cc] = s2;
      }
      for (int r = 0; r < 4; r++) { double s1 = 0, s2 = 0; for (int k = 0; k < 4; k++) { s1 += Mh(r,k)*error.v[k]; s2 += Mh(r,k)*ct.v[k]; } Me[r] = s1; Mct[r] = s2; }
      double* Ht = &Hs[(size_t)omp_get_thread_num() * 36];
      double* bt = &bs[(size_t)omp_get_thread_num() * 6];
      for (int r = 0; r < 6; r++) {
        for (int cc = 0; cc < 6; cc++) {
          double s1 = 0, s2 = 0;
          for (int k = 0; k < 4; k++) { s1 += J1[k*6+r] * MJ1[k*6+cc]; s2 += J2[k*6+r] * MJ2[k*6+cc]; }
          Ht[r*6+cc] += w * (s1 + lam * s2);     // impl.hpp:582
        }
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 4; k++) { s1 += J1[k*6+r] * Me[k]; s2 += J2[k*6+r] * Mct[k]; }
        bt[r] += w * (s1 + lam * s2);            // impl.hpp:584
      }
    }
    if (H && b) {
      for (int k = 0; k < 36; k++) H[k] = 0; for (int k = 0; k < 6; k++) b[k] = 0;
      for (int th = 0; th < num_threads; th++) {
        for (int k = 0; k < 36; k++) H[k] += Hs[(size_t)th*36 + k];
        for (int k = 0; k < 6; k++) b[k] += bs[(size_t)th*6 + k];
      }
    }
    record(3, sum_errors, nullptr, t, H, 6, b, 6);
    return sum_errors;
  }

  // impl.hpp:609-658
  double compute_t_error(const double t[3], const double g[3], const double /*last_t0*/[3], double dtn, double dtn1) {
    double sum_errors = 0.0;
    const size_t pt_size = corr.size();
    const int m = (int)corr.size();
    const double last_tr[4] = {1, 0, 0, 0};   // Vector4d::Identity(), impl.hpp:637 (Q3)
#pragma omp parallel for num_threads(num_threads) reduction(+ : sum_errors)
    for (int i = 0; i < m; i++) sum_errors += t3_terms(i, t, g, last_tr, dtn, dtn1, pt_size, nullptr, nullptr, nullptr);
    record(4, sum_errors, nullptr, t, nullptr, 0, nullptr, 0);
    return sum_errors;
  }

  // lsq.hpp:182-191
  bool is_converged(const M4& delta) const {
    double rmax = 0, tmax = 0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rmax = std::max(rmax, 1.0 / rotation_epsilon * std::fabs(delta(i,j) - (i==j ? 1.0 : 0.0)));
    for (int i = 0; i < 3; i++) tmax = std::max(tmax, 1.0 / transformation_epsilon * std::fabs(delta(i,3)));
    return std::max(rmax, tmax) < 1;
  }
  // lsq.hpp:328-335
  bool is_rot_converged(const M4& delta) const {
    double rmax = 0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rmax = std::max(rmax, 1.0 / rotation_epsilon * std::fabs(delta(i,j) - (i==j ? 1.0 : 0.0)));
    return rmax < 1;
  }
  // lsq.hpp:142-148
  bool is_t_converged(const double d[3]) const {
    double tmax = 0;
    for (int i = 0; i < 3; i++) tmax = std::max(tmax, 1.0 / transformation_epsilon * std::fabs(d[i]));
    return tmax < 1;
  }

  // lsq.hpp:208-222
  bool step_gn(M4& x0, M4& delta) {
    double H[36], b[6], nb[6], d[6];
    linearize_n(x0, 6, H, b);
    for (int i = 0; i < 6; i++) nb[i] = -b[i];
    ldlt_solve(H, nb, 6, d);
    delta = se3_exp(d);
    x0 = m4_mul(delta, x0);
    std::memcpy(final_hessian, H, sizeof(H));
    return true;
  }
  // lsq.hpp:225-270
  bool step_lm(M4& x0, M4& delta) {
    double H[36], b[6];
    double y0 = linearize_n(x0, 6, H, b);
    if (lm_lambda < 0.0) { double mx = 0; for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i*6+i])); lm_lambda = lm_init_lambda_factor * mx; }
    double nu = 2.0;
    for (int i = 0; i < lm_max_iterations; i++) {
      double A[36], nb[6], d[6];
      for (int k = 0; k < 36; k++) A[k] = H[k] + ((k % 7 == 0) ? lm_lambda : 0.0);
      for (int k = 0; k < 6; k++) nb[k] = -b[k];
      ldlt_solve(A, nb, 6, d);
      delta = se3_exp(d);
      M4 xi = m4_mul(delta, x0);
      double yi = compute_error(xi);
      double den = 0; for (int k = 0; k < 6; k++) den += d[k] * (lm_lambda * d[k] - b[k]);
      double rho = (y0 - yi) / den;
      if (rho < 0) {
        if (is_converged(delta)) return true;
        lm_lambda = nu * lm_lambda; nu = 2 * nu;
        continue;
      }
      x0 = xi;
      lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      std::memcpy(final_hessian, H, sizeof(H));
      return true;
    }
    return false;
  }
  // lsq.hpp:273-324
  bool rot_step_lm(M4& x0, M4& delta) {
    double H[9], b[3];
    double y0 = linearize_n(x0, 3, H, b);
    if (lm_lambda < 0.0) { double mx = 0; for (int i = 0; i < 3; i++) mx = std::max(mx, std::fabs(H[i*3+i])); lm_lambda = lm_init_lambda_factor * mx; }
    double nu = 2.0;
    for (int i = 0; i < lm_max_iterations; i++) {
      double A[9], nb[3], d[3];
      for (int k = 0; k < 9; k++) A[k] = H[k] + ((k % 4 == 0) ? lm_lambda : 0.0);
      for (int k = 0; k < 3; k++) nb[k] = -b[k];
      ldlt_solve(A, nb, 3, d);
      double q[4], R[9]; so3_exp(d, q); quat_to_rot(q, R);
      delta = m4_identity();
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) delta(r,c) = R[r*3+c];
      M4 xi = m4_mul(delta, x0);
      double yi = compute_error(xi);
      double den = 0; for (int k = 0; k < 3; k++) den += d[k] * (lm_lambda * d[k] - b[k]);
      double rho = (y0 - yi) / den;
      if (rho < 0) {
        if (is_rot_converged(delta)) return true;
        lm_lambda = nu * lm_lambda; nu = 2 * nu;
        continue;
      }
      x0 = xi;
      lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      std::memcpy(so3_final_hessian, H, sizeof(H));
      return true;
    }
    return false;
  }
  // lsq.hpp:194-205
  bool step_optimize(M4& x0, M4& delta) {
    switch (optimizer) {
      case RO_OPT_LM: return step_lm(x0, delta);
      case RO_OPT_GAUSS_NEWTON: return step_gn(x0, delta);
      case RO_OPT_SO3_LM: return rot_step_lm(x0, delta);
    }
    return step_lm(x0, delta);
  }

  // impl.hpp:146-160 + lsq.hpp:152-179 (+ pcl::Registration::align bookkeeping)
  int align(const float guess[16], float out_T[16]) {
    trace.clear(); tracing = true;
    voxelmap.reset();
    if (source_covs.size() != nsrc()) { int rc = calculate_covariances(source, kd_source, kd_source_ok, source_covs); if (rc) return 100 - rc; }
    if (target_covs.size() != ntgt()) { int rc = calculate_covariances(target, kd_target, kd_target_ok, target_covs); if (rc) return 100 - rc; }
    M4 x0; for (int i = 0; i < 16; i++) x0.m[i] = (double)guess[i];
    x0(3,0) = x0(3,1) = x0(3,2) = 0; x0(3,3) = 1;   // Isometry3d
    lm_lambda = -1.0;
    converged = false;
    int rc = 0;
    for (int i = 0; i < max_iterations && !converged; i++) {
      nr_iterations = i;
      M4 delta;
      if (!step_optimize(x0, delta)) { std::cerr << "lm not converged!!" << std::endl; rc = 1; break; }
      converged = (optimizer == RO_OPT_SO3_LM) ? is_converged(delta) : is_converged(delta);   // lsq.hpp:173 uses is_converged for all types
    }
    final_pose_d = x0;
    for (int i = 0; i < 16; i++) out_T[i] = (float)x0.m[i];
    tracing = false;
    return rc;
  }

  // lsq.hpp:84-139
  bool step_t_optimize(double x0[3], double delta[3], const double g[3], const double last_t0[3], double dtn, double dtn1) {
    double H[36], b[6];
    double y0 = t3_linearize(x0, g, last_t0, dtn, dtn1, H, b);
    if (lm_lambda < 0.0) { double mx = 0; for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i*6+i])); lm_lambda = lm_init_lambda_factor * mx; }
    double nu = 2.0;
    for (int i = 0; i < lm_max_iterations; i++) {
      double A[36], nb[6], d[6];
      for (int k = 0; k < 36; k++) A[k] = H[k] + ((k % 7 == 0) ? lm_lambda : 0.0);
      for (int k = 0; k < 6; k++) nb[k] = -b[k];
      ldlt_solve(A, nb, 6, d);
      M4 dT = se3_exp(d);
      delta[0] = dT(0,3); delta[1] = dT(1,3); delta[2] = dT(2,3);
      double xi[3] = {delta[0] + x0[0], delta[1] + x0[1], delta[2] + x0[2]};
      double yi = compute_t_error(xi, g, last_t0, dtn, dtn1);
      double den = 0; for (int k = 0; k < 6; k++) den += d[k] * (lm_lambda * d[k] - b[k]);
      double rho = (y0 - yi) / den;
      if (rho < 0) {
        if (is_t_converged(delta)) return true;
        lm_lambda = nu * lm_lambda; nu = 2 * nu;
        continue;
      }
      x0[0] = xi[0]; x0[1] = xi[1]; x0[2] = xi[2];
      lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      return true;
    }
    return false;
  }

  // impl.hpp:162-169 + lsq.hpp:55-80
  int compute_translation(double trans[3], const double g[3], const double last_t0[3], double dtn, double dtn1, float ct_lambda) {
    trace.clear(); tracing = true;
    lambda_ = ct_lambda;
    double t0[3] = {trans[0], trans[1], trans[2]};
    lm_lambda = -1.0;
    bool t_converged = false;
    int rc = 0;
    for (int i = 0; i < max_iterations && !t_converged; i++) {
      nr_iterations = i;
      double delta_t[3];
      if (!step_t_optimize(t0, delta_t, g, last_t0, dtn, dtn1)) { std::cerr << "lm not converged!!" << std::endl; rc = 1; break; }
      t_converged = is_t_converged(delta_t);
    }
    trans[0] = t0[0]; trans[1] = t0[1]; trans[2] = t0[2];
    tracing = false;
    return rc;
  }
};

// ---------------------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------------------
extern "C" {

ro_ctx* ro_create(void) { return new ro_ctx; }
void ro_destroy(ro_ctx* c) { delete c; }

void ro_set_num_threads(ro_ctx* c, int n) { c->num_threads = (n == 0) ? omp_get_max_threads() : n; }   // impl.hpp:79-87
int  ro_get_num_threads(ro_ctx* c) { return c->num_threads; }
void ro_set_resolution(ro_ctx* c, double r) { c->voxel_resolution = r; c->voxel_type = RO_VOXEL_UNIFORM; c->voxelmap.reset(); }     // impl.hpp:45-49
void ro_set_polar_resolution(ro_ctx* c, double th, double ph, double r) {                                                          // impl.hpp:52-56
  c->polar_resolution[0] = th; c->polar_resolution[1] = ph; c->polar_resolution[2] = r; c->voxel_type = RO_VOXEL_POLAR; c->voxelmap.reset();
}
void ro_set_neighbor_search_method(ro_ctx* c, int m) { c->search_method = m; }
void ro_set_regularization_method(ro_ctx* c, int m) { c->regularization_method = m; }
void ro_set_correspondence_randomness(ro_ctx* c, int k) { c->k_correspondences = k; }
void ro_set_optimizer_type(ro_ctx* c, int t) { c->optimizer = t; }
void ro_set_max_iterations(ro_ctx* c, int n) { c->max_iterations = n; }
void ro_set_rotation_epsilon(ro_ctx* c, double e) { c->rotation_epsilon = e; }
void ro_set_transformation_epsilon(ro_ctx* c, double e) { c->transformation_epsilon = e; }
void ro_set_initial_lambda_factor(ro_ctx* c, double f) { c->lm_init_lambda_factor = f; }
void ro_set_ct_lambda(ro_ctx* c, float l) { c->lambda_ = l; }

static void load_cloud(std::vector<float>& dst, const float* pts, size_t n, size_t stride) {
  dst.resize(n * 4);
  for (size_t i = 0; i < n; i++) { dst[i*4+0] = pts[i*stride+0]; dst[i*4+1] = pts[i*stride+1]; dst[i*4+2] = pts[i*stride+2]; dst[i*4+3] = 1.0f; }
}
void ro_set_input_source(ro_ctx* c, const float* pts, size_t n, size_t stride) {   // impl.hpp:111-120
  load_cloud(c->source, pts, n, stride); c->kd_source_ok = false; c->source_covs.clear();
}
void ro_set_input_target(ro_ctx* c, const float* pts, size_t n, size_t stride) {   // impl.hpp:132-143
  load_cloud(c->target, pts, n, stride); c->kd_target_ok = false; c->target_covs.clear(); c->voxelmap.reset();
}
void ro_set_covariances(ro_ctx* c, int which, const double* cov, size_t n) {       // impl.hpp:122-130
  auto& v = which == 0 ? c->source_covs : c->target_covs;
  v.resize(n);
  std::memcpy(v.data(), cov, n * sizeof(M4));
}
size_t ro_get_covariances(ro_ctx* c, int which, double* cov) {
  auto& v = which == 0 ? c->source_covs : c->target_covs;
  if (cov) std::memcpy(cov, v.data(), v.size() * sizeof(M4));
  return v.size();
}
int ro_calculate_covariances(ro_ctx* c, int which) {
  return which == 0 ? c->calculate_covariances(c->source, c->kd_source, c->kd_source_ok, c->source_covs)
                    : c->calculate_covariances(c->target, c->kd_target, c->kd_target_ok, c->target_covs);
}
void ro_knn(ro_ctx* c, int which, int k, int* out) {
  auto& cloud = which == 0 ? c->source : c->target;
  KdTree tree; tree.build(cloud.data(), (int)(cloud.size() / 4));
  const int n = (int)(cloud.size() / 4);
#pragma omp parallel for num_threads(c->num_threads) schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    int f = tree.knn(&cloud[i*4], k, out + (size_t)i * k);
    for (int j = f; j < k; j++) out[(size_t)i*k + j] = -1;
  }
}

void ro_create_voxelmap(ro_ctx* c) { c->voxelmap.reset(); c->ensure_voxelmap(); }
void ro_reset_voxelmap(ro_ctx* c) { c->voxelmap.reset(); }
size_t ro_voxel_count(ro_ctx* c) { return c->voxelmap ? c->voxelmap->voxels.size() : 0; }
void ro_get_voxels(ro_ctx* c, int* keys, double* mean, double* cov, int* npts) {
  if (!c->voxelmap) return;
  std::vector<std::pair<Key3, Voxel*>> v;
  for (auto& kv : c->voxelmap->voxels) v.push_back({kv.first, kv.second.get()});
  std::sort(v.begin(), v.end(), [](const std::pair<Key3, Voxel*>& a, const std::pair<Key3, Voxel*>& b) {
    if (a.first.x != b.first.x) return a.first.x < b.first.x;
    if (a.first.y != b.first.y) return a.first.y < b.first.y;
    return a.first.z < b.first.z; });
  for (size_t i = 0; i < v.size(); i++) {
    keys[i*3+0] = v[i].first.x; keys[i*3+1] = v[i].first.y; keys[i*3+2] = v[i].first.z;
    std::memcpy(mean + i*4, v[i].second->mean_dir.v, 4 * sizeof(double));
    std::memcpy(cov + i*16, v[i].second->cov.m, 16 * sizeof(double));
    npts[i] = v[i].second->num_points;
  }
}

static M4 to_m4(const double T[16]) { M4 m; std::memcpy(m.m, T, sizeof(m.m)); return m; }
double ro_linearize(ro_ctx* c, const double T[16], double* H, double* b) { return c->linearize_n(to_m4(T), 6, H, b); }
double ro_so3_linearize(ro_ctx* c, const double T[16], double* H, double* b) { return c->linearize_n(to_m4(T), 3, H, b); }
double ro_compute_error(ro_ctx* c, const double T[16]) { return c->compute_error(to_m4(T)); }
double ro_t3_linearize(ro_ctx* c, const double t[3], const double g[3], const double l[3], double dtn, double dtn1, double* H, double* b) {
  return c->t3_linearize(t, g, l, dtn, dtn1, H, b);
}
double ro_compute_t_error(ro_ctx* c, const double t[3], const double g[3], const double l[3], double dtn, double dtn1) {
  return c->compute_t_error(t, g, l, dtn, dtn1);
}
size_t ro_num_correspondences(ro_ctx* c) { return c->corr.size(); }
void ro_get_correspondences(ro_ctx* c, int* src_idx, int* vox_key) {
  for (size_t i = 0; i < c->corr.size(); i++) {
    src_idx[i] = c->corr[i].first;
    vox_key[i*3+0] = c->corr_keys_dbg[i].x; vox_key[i*3+1] = c->corr_keys_dbg[i].y; vox_key[i*3+2] = c->corr_keys_dbg[i].z;
  }
}

int ro_align(ro_ctx* c, const float guess[16], float out_T[16]) { return c->align(guess, out_T); }
int ro_compute_translation(ro_ctx* c, double trans[3], const double g[3], const double l[3], double dtn, double dtn1, float lam) {
  return c->compute_translation(trans, g, l, dtn, dtn1, lam);
}
/* pcl::Registration::getFitnessScore(max_range) (PCL registration.hpp; not in /root/reference): every source point is moved by
 * T (float, ((a x + b y) + c z) + d), its nearest target point is searched (float squared distance, ties by index), and the
 * squared distances <= max_range are averaged in double, in point order. idx / d2 (optional) receive the per-point result. */
double ro_get_fitness_score(ro_ctx* c, const float T[16], double max_range, int* idx, float* d2) {
  const int n = (int)(c->source.size() / 4);
  KdTree tree; tree.build(c->target.data(), (int)(c->target.size() / 4));
  std::vector<int> id(n, -1); std::vector<float> dd(n, 0.f);
#pragma omp parallel for num_threads(c->num_threads) schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    const float x = c->source[i*4], y = c->source[i*4+1], z = c->source[i*4+2];
    const float q[3] = {((T[0]*x + T[1]*y) + T[2]*z) + T[3], ((T[4]*x + T[5]*y) + T[6]*z) + T[7], ((T[8]*x + T[9]*y) + T[10]*z) + T[11]};
    int j; float d;
    if (tree.knn_dist(q, 1, &j, &d) == 1) { id[i] = j; dd[i] = d; } else { dd[i] = INFINITY; }
  }
  double fitness = 0.0; int nr = 0;
  for (int i = 0; i < n; i++) if ((double)dd[i] <= max_range) { fitness += (double)dd[i]; nr++; }
  if (idx) std::memcpy(idx, id.data(), sizeof(int) * n);
  if (d2) std::memcpy(d2, dd.data(), sizeof(float) * n);
  return nr > 0 ? fitness / nr : std::numeric_limits<double>::max();
}
int ro_has_converged(ro_ctx* c) { return c->converged ? 1 : 0; }
int ro_nr_iterations(ro_ctx* c) { return c->nr_iterations; }
void ro_get_final_hessian(ro_ctx* c, double* H) { std::memcpy(H, c->final_hessian, 36 * sizeof(double)); }
void ro_get_final_pose_d(ro_ctx* c, double T[16]) { std::memcpy(T, c->final_pose_d.m, 16 * sizeof(double)); }

size_t ro_trace_size(ro_ctx* c) { return c->trace.size(); }
void ro_get_trace(ro_ctx* c, double* out) { std::memcpy(out, c->trace.data(), c->trace.size() * sizeof(TraceRec)); }

void ro_so3_exp(const double w[3], double q[4]) { so3_exp(w, q); }
void ro_se3_exp(const double a[6], double T[16]) { M4 m = se3_exp(a); std::memcpy(T, m.m, sizeof(m.m)); }
void ro_svd3(const double A[9], double U[9], double S[3], double V[9]) { svd3(A, U, S, V); }
int  ro_ldlt_solve(const double* A, const double* rhs, int n, double* x) { return ldlt_solve(A, rhs, n, x); }

}  // extern "C"
