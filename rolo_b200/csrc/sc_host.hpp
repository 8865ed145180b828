// sc_host.hpp — host side of the Scan Context place recogniser (src/scancontext/Scancontext.cpp:69-149,236-344).
// Descriptors are 20 x 60 doubles (built on the device, scancontext.cu); everything here is a few thousand flops per
// query and stays on the host like the 6x6 solves: column-shift alignment, cosine distance, ring-key candidate search,
// threshold. Sums run sequentially (Eigen's vectorised reductions differ in the last bit; the oracle defines parity).
#pragma once
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

namespace rolo {
namespace sc {

constexpr int NR = 20, NS = 60;                 // PC_NUM_RING, PC_NUM_SECTOR (Scancontext.h:82-83)
constexpr int NUM_EXCLUDE_RECENT = 30;          // :89
constexpr int NUM_CANDIDATES_FROM_TREE = 3;     // :90
constexpr double SEARCH_RATIO = 0.1;            // :93
constexpr double SC_DIST_THRES = 0.4;           // :95
constexpr int TREE_MAKING_PERIOD = 10;          // :99

struct Entry { double desc[NR * NS]; double ring_key[NR]; double sector_key[NS]; float ring_key_f[NR]; };

// column c of `b` compared with column (c - shift) of the unshifted matrix == circshift(b, shift) (:39-59)
inline double dist_direct(const double* a, const double* b, int shift) {   // distDirectSC on (a, circshift(b, shift))
  int eff = 0; double sum = 0.0;
  for (int c = 0; c < NS; c++) {
    const int cb = ((c - shift) % NS + NS) % NS;
    double na = 0.0, nb = 0.0, dot = 0.0;
    for (int r = 0; r < NR; r++) { const double x = a[r * NS + c], y = b[r * NS + cb]; na += x * x; nb += y * y; dot += x * y; }
    na = std::sqrt(na); nb = std::sqrt(nb);
    if ((na == 0) | (nb == 0)) continue;
    sum = sum + dot / (na * nb);
    eff = eff + 1;
  }
  return 1.0 - sum / eff;          // eff == 0 gives NaN exactly like the reference's 0/0
}

inline int fast_align(const double* vkey1, const double* vkey2) {          // fastAlignUsingVkey (:93-113)
  int arg = 0; double best = 10000000;
  for (int shift = 0; shift < NS; shift++) {
    double ss = 0.0;
    for (int c = 0; c < NS; c++) { const double d = vkey1[c] - vkey2[((c - shift) % NS + NS) % NS]; ss += d * d; }
    const double nrm = std::sqrt(ss);
    if (nrm < best) { arg = shift; best = nrm; }
  }
  return arg;
}

inline std::pair<double, int> distance(const Entry& a, const Entry& b) {   // distanceBtnScanContext (:116-148)
  const int arg = fast_align(a.sector_key, b.sector_key);
  const int radius = (int)std::round(0.5 * SEARCH_RATIO * NS);
  std::vector<int> space{arg};
  for (int ii = 1; ii < radius + 1; ii++) { space.push_back((arg + ii + NS) % NS); space.push_back((arg - ii + NS) % NS); }
  std::sort(space.begin(), space.end());
  int best_shift = 0; double best = 10000000;
  for (int s : space) { const double d = dist_direct(a.desc, b.desc, s); if (d < best) { best_shift = s; best = d; } }
  return {best, best_shift};
}

struct Manager {
  std::vector<Entry> entries;
  std::vector<int> search_set;       // snapshot of the searchable entries (the reference rebuilds its kd-tree every 10th call)
  int tree_counter = 0;

  // detectLoopClosureID (:253-344): returns (loop id or -1, yaw difference in radians)
  std::pair<int, float> detect() {
    const int n = (int)entries.size();
    if (n < NUM_EXCLUDE_RECENT + 1) return {-1, 0.0f};
    if (tree_counter % TREE_MAKING_PERIOD == 0) { search_set.clear(); for (int i = 0; i < n - NUM_EXCLUDE_RECENT; i++) search_set.push_back(i); }
    tree_counter = tree_counter + 1;
    const Entry& cur = entries.back();
    // exact 3-NN over the float ring keys (nanoflann L2_Simple: float squared distance summed in order); ties by index
    std::vector<std::pair<float, int>> cand;
    for (int i : search_set) {
      float d = 0.f;
      for (int k = 0; k < NR; k++) { const float diff = cur.ring_key_f[k] - entries[i].ring_key_f[k]; d += diff * diff; }
      cand.push_back({d, i});
    }
    const int kk = std::min<int>(NUM_CANDIDATES_FROM_TREE, (int)cand.size());
    std::partial_sort(cand.begin(), cand.begin() + kk, cand.end());
    double min_dist = 10000000; int nn_align = 0, nn_idx = 0;
    for (int c = 0; c < kk; c++) {
      const std::pair<double, int> r = distance(cur, entries[cand[c].second]);
      if (r.first < min_dist) { min_dist = r.first; nn_align = r.second; nn_idx = cand[c].second; }
    }
    const int loop_id = (min_dist < SC_DIST_THRES) ? nn_idx : -1;
    const float yaw = (float)((double)(float)(nn_align * (360.0 / NS)) * M_PI / 180.0);    // deg2rad(float) (:17-20)
    return {loop_id, yaw};
  }
};

}  // namespace sc
}  // namespace rolo
