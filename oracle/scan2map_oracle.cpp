/*
 * scan2map_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code) of backMapping's scan-to-map optimisation
 * (SURVEY.md §8 f3). Follows src/backMapping.cpp of the reference (sdwyc/ROLO @ 3920bb7):
 *   scan2MapOptimization      :681-710   (30 iterations; transformUpdate()'s IMU fusion is outside the path)
 *   updatePointAssociateToMap :713-716,  trans2Affine3f :339-342,  pointAssociateToMap :293-299
 *   cornerOptimization        :719-824   (5-NN, PCA line test, point-to-line coefficient)
 *   surfOptimization          :827-908   (5-NN, plane fit, point-to-plane coefficient)
 *   combineOptimizationCoeffs :911-935   (corners first, then surfaces, index order)
 *   LMOptimization            :937-1054  (Jacobian rows, AtA/AtB, QR solve, degeneracy projection, convergence test)
 * All arithmetic is float like the reference's (PointType = pcl::PointXYZI, cv::Mat CV_32F); AtA/AtB are accumulated in
 * double and rounded to float (cv::gemm's working type for CV_32F is double).
 *
 * PINNING: PINNED to the reference's own code. backMapping.cpp cannot be compiled as a file here (roscpp, tf, GTSAM, small_gicp, PCL I/O
 * are absent), but the member functions above can: oracle/extract_ref_members.py takes their text from the reference file where it lies
 * at build time and splices it into oracle/ref_bm_driver.cpp in the pipe to the compiler (unmodified, with the reference's include/rolo/utility.h in front):
 * oracle/_ref/librolo_ref_bm.so, and tests/test_scan2map_vs_ref_cpu.py requires this restatement to be BIT-EQUAL to them: selection
 * flags and coefficients per point, every Gauss-Newton pass, the degeneracy projection, the convergence test, the whole
 * scan2MapOptimization trajectory. Both sides share the restated third-party numerics (oracle/s2m_numerics.hpp) and the exact kd-tree;
 * those are pinned where the real library exists in the image: cv::eigen (3x3 and 6x6) and cv::Mat::inv are BIT-EQUAL to OpenCV 4.13
 * (python cv2; golden vectors tests/golden/opencv_s2m_golden.npz from tools/make_opencv_golden.py, plus a live check);
 * cv::solve(DECOMP_QR) agrees to rounding only (that build dispatches to LAPACK). Restated from their published algorithms without a
 * real library to compare with: pcl::getTransformation's closed form (checked against scipy), exact kNN under (float distance, index)
 * order, Eigen::colPivHouseholderQr (checked against numpy lstsq), cv::gemm's double accumulation. Compiled with -ffp-contract=off.
 */
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "kdtree.hpp"
#include "s2m_numerics.hpp"

namespace {
using namespace s2m_num;

// pcl::getTransformation(x, y, z, roll, pitch, yaw) (pcl/common/impl/eigen.hpp), rows of the 3x4 float matrix
void get_transformation(const float tf[6], float t[12]) {
  const float roll = tf[0], pitch = tf[1], yaw = tf[2];
  const float A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch), E = std::cos(roll), F = std::sin(roll);
  const float DE = D * E, DF = D * F;
  t[0] = A * C; t[1] = A * DF - B * E; t[2] = B * F + A * DE; t[3] = tf[3];
  t[4] = B * C; t[5] = A * E + B * DF; t[6] = B * DE - A * F; t[7] = tf[4];
  t[8] = -D;    t[9] = C * F;          t[10] = C * E;         t[11] = tf[5];
}

struct S2M {
  std::vector<float> map_corner, map_surf, scan_corner, scan_surf;   // xyzi, stride 4
  KdTree tree_corner, tree_surf;
  int threads = 0;
  // per-iteration products (combineOptimizationCoeffs inputs)
  std::vector<float> coeff_corner, coeff_surf;                       // 4 per scan point
  std::vector<uint8_t> flag_corner, flag_surf;
  bool degenerate = false;
  float matP[36];
};

inline void associate(const float t[12], const float* pi, float po[3]) {
  po[0] = t[0] * pi[0] + t[1] * pi[1] + t[2] * pi[2] + t[3];
  po[1] = t[4] * pi[0] + t[5] * pi[1] + t[6] * pi[2] + t[7];
  po[2] = t[8] * pi[0] + t[9] * pi[1] + t[10] * pi[2] + t[11];
}

// backMapping.cpp:719-824
void corner_optimization(S2M& s, const float t[12]) {
  const int n = (int)s.scan_corner.size() / 4;
  s.coeff_corner.assign((size_t)n * 4, 0.f); s.flag_corner.assign(n, 0);
  const float* M = s.map_corner.data();
  #pragma omp parallel for num_threads(s.threads > 0 ? s.threads : omp_get_max_threads())
  for (int i = 0; i < n; i++) {
    const float* pointOri = &s.scan_corner[i * 4];
    float sel[3]; associate(t, pointOri, sel);
    int ind[5]; float dis[5];
    if (s.tree_corner.knn_dist(sel, 5, ind, dis) < 5) continue;
    if (!(dis[4] < 1.0f)) continue;
    float cx = 0, cy = 0, cz = 0;
    for (int j = 0; j < 5; j++) { cx += M[ind[j] * 4]; cy += M[ind[j] * 4 + 1]; cz += M[ind[j] * 4 + 2]; }
    cx /= 5; cy /= 5; cz /= 5;
    float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
    for (int j = 0; j < 5; j++) {
      const float ax = M[ind[j] * 4] - cx, ay = M[ind[j] * 4 + 1] - cy, az = M[ind[j] * 4 + 2] - cz;
      a11 += ax * ax; a12 += ax * ay; a13 += ax * az; a22 += ay * ay; a23 += ay * az; a33 += az * az;
    }
    a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
    float A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33}, D[3], V[9];
    jacobi_eigen(3, A, D, V);
    if (!(D[0] > 3 * D[1])) continue;
    const float x0 = sel[0], y0 = sel[1], z0 = sel[2];
    // `cx + 0.1 * v`: the literal is a double, so the sum is formed in double and rounded once (:777-783)
    const float x1 = (float)(cx + 0.1 * V[0]), y1 = (float)(cy + 0.1 * V[1]), z1 = (float)(cz + 0.1 * V[2]);
    const float x2 = (float)(cx - 0.1 * V[0]), y2 = (float)(cy - 0.1 * V[1]), z2 = (float)(cz - 0.1 * V[2]);
    const float m11 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    const float m22 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    const float m33 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    const float a012 = std::sqrt(m11 * m11 + m22 * m22 + m33 * m33);
    const float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    const float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
    const float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
    const float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
    const float ld2 = a012 / l12;
    const float sc = (float)(1 - 0.9 * std::fabs(ld2));      // double literal (:804)
    if (sc > 0.1) {                                           // compared as double (:814)
      s.coeff_corner[i * 4] = sc * la; s.coeff_corner[i * 4 + 1] = sc * lb; s.coeff_corner[i * 4 + 2] = sc * lc; s.coeff_corner[i * 4 + 3] = sc * ld2;
      s.flag_corner[i] = 1;
    }
  }
}

// backMapping.cpp:827-908
void surf_optimization(S2M& s, const float t[12]) {
  const int n = (int)s.scan_surf.size() / 4;
  s.coeff_surf.assign((size_t)n * 4, 0.f); s.flag_surf.assign(n, 0);
  const float* M = s.map_surf.data();
  #pragma omp parallel for num_threads(s.threads > 0 ? s.threads : omp_get_max_threads())
  for (int i = 0; i < n; i++) {
    const float* pointOri = &s.scan_surf[i * 4];
    float sel[3]; associate(t, pointOri, sel);
    int ind[5]; float dis[5];
    if (s.tree_surf.knn_dist(sel, 5, ind, dis) < 5) continue;
    if (!(dis[4] < 1.0f)) continue;
    float A0[15], B0[5] = {-1, -1, -1, -1, -1}, X0[3];
    for (int j = 0; j < 5; j++) { A0[j * 3] = M[ind[j] * 4]; A0[j * 3 + 1] = M[ind[j] * 4 + 1]; A0[j * 3 + 2] = M[ind[j] * 4 + 2]; }
    qr_solve(5, 3, A0, B0, X0, true);
    float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
    const float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
    bool planeValid = true;
    for (int j = 0; j < 5; j++)
      if (std::fabs(pa * M[ind[j] * 4] + pb * M[ind[j] * 4 + 1] + pc * M[ind[j] * 4 + 2] + pd) > 0.2) { planeValid = false; break; }   // vs the double 0.2 (:879)
    if (!planeValid) continue;
    const float pd2 = pa * sel[0] + pb * sel[1] + pc * sel[2] + pd;
    const float sc = (float)(1 - 0.9 * std::fabs(pd2) / std::sqrt(std::sqrt(pointOri[0] * pointOri[0] + pointOri[1] * pointOri[1] + pointOri[2] * pointOri[2])));   // :891
    if (sc > 0.1) {
      s.coeff_surf[i * 4] = sc * pa; s.coeff_surf[i * 4 + 1] = sc * pb; s.coeff_surf[i * 4 + 2] = sc * pc; s.coeff_surf[i * 4 + 3] = sc * pd2;
      s.flag_surf[i] = 1;
    }
  }
}

// backMapping.cpp:937-1006: rows of matA / matB for the selected points, AtA / AtB (double accumulation, float result)
int build_normal_equations(const S2M& s, const float tf[6], float AtA[36], float AtB[6]) {
  const float srx = std::sin(tf[1]), crx = std::cos(tf[1]), sry = std::sin(tf[2]), cry = std::cos(tf[2]), srz = std::sin(tf[0]), crz = std::cos(tf[0]);
  double H[36] = {0}, g[6] = {0};
  int count = 0;
  for (int pass = 0; pass < 2; pass++) {
    const std::vector<float>& pts = pass == 0 ? s.scan_corner : s.scan_surf;
    const std::vector<float>& cf = pass == 0 ? s.coeff_corner : s.coeff_surf;
    const std::vector<uint8_t>& fl = pass == 0 ? s.flag_corner : s.flag_surf;
    for (size_t i = 0; i < fl.size(); i++) {
      if (!fl[i]) continue;
      count++;
      const float px = pts[i * 4 + 1], py = pts[i * 4 + 2], pz = pts[i * 4];          // lidar -> camera
      const float cx = cf[i * 4 + 1], cy = cf[i * 4 + 2], cz = cf[i * 4], ci = cf[i * 4 + 3];
      const float arx = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * cx + (-srx * srz * px - crz * srx * py - crx * pz) * cy +
                        (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * cz;
      const float ary = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * cx +
                        ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * cz;
      const float arz = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * cx + (crx * crz * px - crx * srz * py) * cy +
                        ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * cz;
      const float row[6] = {arz, arx, ary, cz, cx, cy};
      const float bi = -ci;
      for (int a = 0; a < 6; a++) { for (int b = 0; b < 6; b++) H[a * 6 + b] += (double)row[a] * (double)row[b]; g[a] += (double)row[a] * (double)bi; }
    }
  }
  for (int a = 0; a < 36; a++) AtA[a] = (float)H[a];
  for (int a = 0; a < 6; a++) AtB[a] = (float)g[a];
  return count;
}

}  // namespace

extern "C" {

void* ro_s2m_create() { return new S2M(); }
void ro_s2m_destroy(void* h) { delete (S2M*)h; }
void ro_s2m_set_num_threads(void* h, int n) { ((S2M*)h)->threads = n; }

static void copy_cloud(std::vector<float>& dst, const float* p, size_t n, size_t stride_floats) {
  dst.resize(n * 4);
  for (size_t i = 0; i < n; i++) { dst[i * 4] = p[i * stride_floats]; dst[i * 4 + 1] = p[i * stride_floats + 1]; dst[i * 4 + 2] = p[i * stride_floats + 2]; dst[i * 4 + 3] = 0.f; }
}
/* kdtreeCornerFromMap->setInputCloud(laserCloudCornerFromMapDS) / kdtreeSurfFromMap->... (backMapping.cpp:689-690) */
void ro_s2m_set_map(void* h, const float* corner, size_t n_corner, const float* surf, size_t n_surf, size_t stride_floats) {
  S2M& s = *(S2M*)h;
  copy_cloud(s.map_corner, corner, n_corner, stride_floats); copy_cloud(s.map_surf, surf, n_surf, stride_floats);
  s.tree_corner.build(s.map_corner.data(), (int)n_corner); s.tree_surf.build(s.map_surf.data(), (int)n_surf);
}
/* laserCloudCornerLastDS / laserCloudSurfLastDS */
void ro_s2m_set_scan(void* h, const float* corner, size_t n_corner, const float* surf, size_t n_surf, size_t stride_floats) {
  S2M& s = *(S2M*)h;
  copy_cloud(s.scan_corner, corner, n_corner, stride_floats); copy_cloud(s.scan_surf, surf, n_surf, stride_floats);
}

/* one pass of cornerOptimization + surfOptimization + combineOptimizationCoeffs + the matA/matB assembly at pose tf
 * (roll pitch yaw x y z): AtA 6x6, AtB 6, returns laserCloudSelNum. coeff_* (4 floats per scan point) / flag_* optional. */
int ro_s2m_linearize(void* h, const float tf[6], float AtA[36], float AtB[6], float* coeff_corner, uint8_t* flag_corner, float* coeff_surf,
                     uint8_t* flag_surf) {
  S2M& s = *(S2M*)h;
  float t[12]; get_transformation(tf, t);
  corner_optimization(s, t); surf_optimization(s, t);
  const int n = build_normal_equations(s, tf, AtA, AtB);
  if (coeff_corner) std::memcpy(coeff_corner, s.coeff_corner.data(), s.coeff_corner.size() * 4);
  if (flag_corner) std::memcpy(flag_corner, s.flag_corner.data(), s.flag_corner.size());
  if (coeff_surf) std::memcpy(coeff_surf, s.coeff_surf.data(), s.coeff_surf.size() * 4);
  if (flag_surf) std::memcpy(flag_surf, s.flag_surf.data(), s.flag_surf.size());
  return n;
}

/* scan2MapOptimization (backMapping.cpp:681-710) without transformUpdate(). tf inout. Returns the number of LMOptimization
 * calls made; *converged = LMOptimization returned true; *degenerate = isDegenerate. edge_min / surf_min are
 * edgeFeatureMinValidNum / surfFeatureMinValidNum (returns -1 when the scan has too few features, :687,705-707). */
int ro_s2m_optimize(void* h, float tf[6], int max_iter, int edge_min, int surf_min, int* converged, int* degenerate) {
  S2M& s = *(S2M*)h;
  *converged = 0; *degenerate = 0;
  if (!((int)s.scan_corner.size() / 4 > edge_min && (int)s.scan_surf.size() / 4 > surf_min)) return -1;
  s.degenerate = false;
  int it = 0;
  for (int iterCount = 0; iterCount < max_iter; iterCount++) {
    float t[12]; get_transformation(tf, t);
    corner_optimization(s, t); surf_optimization(s, t);
    float AtA[36], AtB[6], X[6];
    it++;
    const int sel = build_normal_equations(s, tf, AtA, AtB);
    if (sel < 50) continue;                                            // :956-958
    { float A[36], b[6]; std::memcpy(A, AtA, sizeof(A)); std::memcpy(b, AtB, sizeof(b)); qr_solve(6, 6, A, b, X, false); }
    if (iterCount == 0) {                                              // :1008-1029
      float A[36], E[6], V[36], V2[36];
      std::memcpy(A, AtA, sizeof(A));
      jacobi_eigen(6, A, E, V);
      std::memcpy(V2, V, sizeof(V2));
      s.degenerate = false;
      for (int i = 5; i >= 0; i--) {
        if (E[i] < 100.f) { for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0; s.degenerate = true; } else break;
      }
      float Vi[36];
      if (!invert_lu(6, V, Vi)) std::memset(Vi, 0, sizeof(Vi));       // cv::Mat::inv returns zeros for a singular matrix
      for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {          // cv::gemm: double working type
        double acc = 0.0; for (int k = 0; k < 6; k++) acc += (double)Vi[a * 6 + k] * (double)V2[k * 6 + b]; s.matP[a * 6 + b] = (float)acc; }
    }
    if (s.degenerate) {
      float X2[6]; std::memcpy(X2, X, sizeof(X2));
      for (int a = 0; a < 6; a++) { double acc = 0.0; for (int k = 0; k < 6; k++) acc += (double)s.matP[a * 6 + k] * (double)X2[k]; X[a] = (float)acc; }
    }
    for (int a = 0; a < 6; a++) tf[a] += X[a];
    // pcl::rad2deg(float) = alpha * 57.29578f; pow(float, 2) is evaluated in double; the sums and sqrt are double (:1040-1048)
    const float r0 = X[0] * 57.29578f, r1 = X[1] * 57.29578f, r2 = X[2] * 57.29578f, c0 = X[3] * 100, c1 = X[4] * 100, c2 = X[5] * 100;
    const float deltaR = (float)std::sqrt((double)r0 * r0 + (double)r1 * r1 + (double)r2 * r2);
    const float deltaT = (float)std::sqrt((double)c0 * c0 + (double)c1 * c1 + (double)c2 * c2);
    if (deltaR < 0.05 && deltaT < 0.05) { *converged = 1; break; }
  }
  *degenerate = s.degenerate ? 1 : 0;
  return it;
}

/* host helpers exposed for unit tests */
void ro_s2m_get_transformation(const float tf[6], float t12[12]) { get_transformation(tf, t12); }
void ro_s2m_eigen(int n, const float* A, float* W, float* V) { std::vector<float> a(A, A + n * n); jacobi_eigen(n, a.data(), W, V); }
void ro_s2m_qr_solve(int rows, int cols, const float* A, const float* b, float* x, int pivot) {
  std::vector<float> a(A, A + rows * cols), bb(b, b + rows); qr_solve(rows, cols, a.data(), bb.data(), x, pivot != 0);
}
int ro_s2m_invert(int n, const float* A, float* inv) { return invert_lu(n, A, inv) ? 1 : 0; }
}
