// ref_sc_driver.cpp — C API over the REFERENCE's own SCManager (src/scancontext/Scancontext.cpp compiled from /root/reference,
// with its vendored nanoflann) for pinning oracle/scancontext_oracle.cpp. Test infrastructure; built only where the reference
// checkout exists (`make -C oracle ref`), output in oracle/_ref/.
#include <cstring>

#include "scancontext/Scancontext.h"

extern "C" {
void* rsc_create() { return new SCManager(); }
void rsc_destroy(void* h) { delete (SCManager*)h; }
void rsc_make_and_save(void* h, const float* pts, size_t n, size_t stride_floats) {
  pcl::PointCloud<SCPointType> cloud;
  cloud.points.resize(n);
  for (size_t i = 0; i < n; i++) { cloud.points[i].x = pts[i * stride_floats]; cloud.points[i].y = pts[i * stride_floats + 1]; cloud.points[i].z = pts[i * stride_floats + 2]; }
  ((SCManager*)h)->makeAndSaveScancontextAndKeys(cloud);
}
int rsc_size(void* h) { return (int)((SCManager*)h)->polarcontexts_.size(); }
// row-major [ring][sector] descriptor, ring key, sector key of entry `index` (-1 = last)
int rsc_get(void* h, int index, double* desc, double* ringkey, double* sectorkey) {
  SCManager& m = *(SCManager*)h;
  const int n = (int)m.polarcontexts_.size();
  if (index < 0) index += n;
  if (index < 0 || index >= n) return -1;
  const Eigen::MatrixXd& d = m.polarcontexts_[index];
  for (int r = 0; r < d.rows(); r++) for (int c = 0; c < d.cols(); c++) desc[r * d.cols() + c] = d(r, c);
  for (int r = 0; r < d.rows(); r++) ringkey[r] = m.polarcontext_invkeys_[index](r, 0);
  for (int c = 0; c < d.cols(); c++) sectorkey[c] = m.polarcontext_vkeys_[index](0, c);
  return 0;
}
int rsc_distance(void* h, int i, int j, double* dist, int* shift) {
  SCManager& m = *(SCManager*)h;
  const int n = (int)m.polarcontexts_.size();
  if (i < 0) i += n; if (j < 0) j += n;
  if (i < 0 || j < 0 || i >= n || j >= n) return -1;
  std::pair<double, int> r = m.distanceBtnScanContext(m.polarcontexts_[i], m.polarcontexts_[j]);
  *dist = r.first; *shift = r.second;
  return 0;
}
int rsc_detect(void* h, float* yaw_diff_rad) {
  std::pair<int, float> r = ((SCManager*)h)->detectLoopClosureID();
  *yaw_diff_rad = r.second;
  return r.first;
}
}
