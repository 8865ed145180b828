/*
 * scancontext_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code) of the Scan Context place recogniser.
 * Follows src/scancontext/Scancontext.cpp (+ include/scancontext/Scancontext.h:80-99) of the reference:
 *   xy2theta :23-36, circshift :39-59, eig2stdvec :62-66, distDirectSC :69-90, fastAlignUsingVkey :93-113,
 *   distanceBtnScanContext :116-148, makeScancontext :151-195, makeRingkeyFromScancontext :198-211,
 *   makeSectorkeyFromScancontext :214-227, makeAndSaveScancontextAndKeys :236-250, detectLoopClosureID :253-344.
 * Matrices are Eigen::MatrixXd in the reference; here plain row-major [ring][sector] doubles.
 *
 * PINNING: the reference has no test for this code, but the code itself compiles here: oracle/_ref/librolo_ref_sc.so is the
 * reference's own Scancontext.cpp with its vendored nanoflann kd-tree, built from /root/reference against oracle/ref_shim
 * (`make -C oracle ref`, oracle/ref_sc_driver.cpp), and tests/test_scancontext_vs_ref_cpu.py requires descriptors, keys,
 * distances / shifts and loop detections of this restatement to EQUAL it (they do, bit for bit, on 45-frame sequences).
 * What stays restated: Eigen's vectorised norm()/dot()/mean() reductions (in-order sums in ref_shim/Eigen/Dense and here)
 * and libm's atan(float), here the correctly rounded value (double atan rounded to float) so that the GPU can reproduce it.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace {
const int NR = 20, NS = 60;
const double PC_MAX_RADIUS = 80.0, SEARCH_RATIO = 0.1, SC_DIST_THRES = 0.4;
const int NUM_EXCLUDE_RECENT = 30, NUM_CANDIDATES_FROM_TREE = 3, TREE_MAKING_PERIOD_ = 10;

float atan_f(float v) { return (float)std::atan((double)v); }

float xy2theta(float x, float y) {
  if ((x >= 0) & (y >= 0)) return (float)((180 / M_PI) * atan_f(y / x));
  if ((x < 0) & (y >= 0)) return (float)(180 - ((180 / M_PI) * atan_f(y / (-x))));
  if ((x < 0) & (y < 0)) return (float)(180 + ((180 / M_PI) * atan_f(y / x)));
  if ((x >= 0) & (y < 0)) return (float)(360 - ((180 / M_PI) * atan_f((-y) / x)));
  return std::nanf("");
}

int to_int_x86(double v) { return (v != v || v >= 2147483648.0 || v < -2147483648.0) ? INT32_MIN : (int)v; }   // cvttsd2si

struct SC { std::vector<double> desc, ringkey, sectorkey; std::vector<float> invkey; };

SC make_scancontext(const float* pts, size_t n, size_t stride, double lidar_height) {
  const int NO_POINT = -1000;
  SC s; s.desc.assign(NR * NS, (double)NO_POINT);
  for (size_t i = 0; i < n; i++) {
    const float x = pts[i * stride], y = pts[i * stride + 1];
    const float z = (float)(pts[i * stride + 2] + lidar_height);
    const float azim_range = std::sqrt(x * x + y * y);
    const float azim_angle = xy2theta(x, y);
    if (azim_range > PC_MAX_RADIUS) continue;
    const int ring_idx = std::max(std::min(NR, to_int_x86(std::ceil((azim_range / PC_MAX_RADIUS) * NR))), 1);
    const int sctor_idx = std::max(std::min(NS, to_int_x86(std::ceil((azim_angle / 360.0) * NS))), 1);
    if (s.desc[(ring_idx - 1) * NS + (sctor_idx - 1)] < z) s.desc[(ring_idx - 1) * NS + (sctor_idx - 1)] = z;
  }
  for (double& v : s.desc) if (v == NO_POINT) v = 0;
  s.ringkey.resize(NR); s.sectorkey.resize(NS); s.invkey.resize(NR);
  for (int r = 0; r < NR; r++) { double a = 0; for (int c = 0; c < NS; c++) a += s.desc[r * NS + c]; s.ringkey[r] = a / NS; s.invkey[r] = (float)s.ringkey[r]; }
  for (int c = 0; c < NS; c++) { double a = 0; for (int r = 0; r < NR; r++) a += s.desc[r * NS + c]; s.sectorkey[c] = a / NR; }
  return s;
}

std::vector<double> circshift(const std::vector<double>& m, int rows, int shift) {
  const int cols = (int)m.size() / rows;
  if (shift == 0) return m;
  std::vector<double> o(m.size(), 0.0);
  for (int c = 0; c < cols; c++) { const int nl = (c + shift) % cols; for (int r = 0; r < rows; r++) o[r * cols + nl] = m[r * cols + c]; }
  return o;
}

double dist_direct(const std::vector<double>& a, const std::vector<double>& b) {
  int num_eff_cols = 0; double sum = 0;
  for (int c = 0; c < NS; c++) {
    double na = 0, nb = 0, dot = 0;
    for (int r = 0; r < NR; r++) { na += a[r * NS + c] * a[r * NS + c]; nb += b[r * NS + c] * b[r * NS + c]; dot += a[r * NS + c] * b[r * NS + c]; }
    na = std::sqrt(na); nb = std::sqrt(nb);
    if ((na == 0) | (nb == 0)) continue;
    sum = sum + dot / (na * nb);
    num_eff_cols = num_eff_cols + 1;
  }
  return 1.0 - sum / num_eff_cols;
}

int fast_align(const std::vector<double>& v1, const std::vector<double>& v2) {
  int arg = 0; double best = 10000000;
  for (int shift = 0; shift < NS; shift++) {
    const std::vector<double> sh = circshift(v2, 1, shift);
    double ss = 0; for (int c = 0; c < NS; c++) { const double d = v1[c] - sh[c]; ss += d * d; }
    const double nrm = std::sqrt(ss);
    if (nrm < best) { arg = shift; best = nrm; }
  }
  return arg;
}

std::pair<double, int> distance_btn(const SC& a, const SC& b) {
  const int arg = fast_align(a.sectorkey, b.sectorkey);
  const int SEARCH_RADIUS = (int)std::round(0.5 * SEARCH_RATIO * NS);
  std::vector<int> space{arg};
  for (int ii = 1; ii < SEARCH_RADIUS + 1; ii++) { space.push_back((arg + ii + NS) % NS); space.push_back((arg - ii + NS) % NS); }
  std::sort(space.begin(), space.end());
  int argmin = 0; double best = 10000000;
  for (int sft : space) { const double d = dist_direct(a.desc, circshift(b.desc, NR, sft)); if (d < best) { argmin = sft; best = d; } }
  return {best, argmin};
}

struct Manager { std::vector<SC> all; std::vector<int> to_search; int counter = 0; };
}  // namespace

extern "C" {
void* ro_sc_create() { return new Manager(); }
void ro_sc_destroy(void* h) { delete (Manager*)h; }
void ro_sc_make_and_save(void* h, const float* pts, size_t n, size_t stride_floats, double lidar_height) {
  ((Manager*)h)->all.push_back(make_scancontext(pts, n, stride_floats, lidar_height));
}
int ro_sc_size(void* h) { return (int)((Manager*)h)->all.size(); }
int ro_sc_get(void* h, int index, double* desc, double* ringkey, double* sectorkey) {
  Manager& m = *(Manager*)h;
  if (index < 0) index += (int)m.all.size();
  if (index < 0 || index >= (int)m.all.size()) return -1;
  std::memcpy(desc, m.all[index].desc.data(), sizeof(double) * NR * NS);
  std::memcpy(ringkey, m.all[index].ringkey.data(), sizeof(double) * NR);
  std::memcpy(sectorkey, m.all[index].sectorkey.data(), sizeof(double) * NS);
  return 0;
}
int ro_sc_distance(void* h, int i, int j, double* dist, int* shift) {
  Manager& m = *(Manager*)h;
  const int n = (int)m.all.size();
  if (i < 0) i += n; if (j < 0) j += n;
  if (i < 0 || j < 0 || i >= n || j >= n) return -1;
  const std::pair<double, int> r = distance_btn(m.all[i], m.all[j]);
  *dist = r.first; *shift = r.second;
  return 0;
}
/* detectLoopClosureID: returns loop id (-1: none), *yaw_diff_rad */
int ro_sc_detect(void* h, float* yaw_diff_rad) {
  Manager& m = *(Manager*)h;
  *yaw_diff_rad = 0.0f;
  const int n = (int)m.all.size();
  if (n < NUM_EXCLUDE_RECENT + 1) return -1;
  if (m.counter % TREE_MAKING_PERIOD_ == 0) { m.to_search.clear(); for (int i = 0; i < n - NUM_EXCLUDE_RECENT; i++) m.to_search.push_back(i); }
  m.counter = m.counter + 1;
  const SC& cur = m.all.back();
  std::vector<std::pair<float, int>> cand;
  for (int i : m.to_search) {
    float d = 0.f;
    for (int k = 0; k < NR; k++) { const float diff = cur.invkey[k] - m.all[i].invkey[k]; d += diff * diff; }
    cand.push_back({d, i});
  }
  std::sort(cand.begin(), cand.end());
  double min_dist = 10000000; int nn_align = 0, nn_idx = 0;
  for (int c = 0; c < NUM_CANDIDATES_FROM_TREE && c < (int)cand.size(); c++) {
    const std::pair<double, int> r = distance_btn(cur, m.all[cand[c].second]);
    if (r.first < min_dist) { min_dist = r.first; nn_align = r.second; nn_idx = cand[c].second; }
  }
  const float deg = (float)(nn_align * (360.0 / double(NS)));
  *yaw_diff_rad = (float)(deg * M_PI / 180.0);
  return min_dist < SC_DIST_THRES ? nn_idx : -1;
}
}
