// kdtree.hpp — exact kNN of the CPU ORACLE (test infrastructure), shared by rolo_oracle.cpp and scan2map_oracle.cpp.
#pragma once
#include <algorithm>
#include <vector>

namespace {
// ---------------------------------------------------------------------------------------
// exact kNN: stands in for pcl::search::KdTree -> FLANN single kd-tree, L2_Simple<float>
// (call site impl.hpp:436). Distances are float, accumulated x,y,z in order without FMA;
// result = k smallest under (distance, index) lexicographic order, sorted ascending.
// ---------------------------------------------------------------------------------------
struct KdTree {
  struct Node { int left, right, begin, end, dim; float split; };
  const float* pts = nullptr;  // stride 4
  int n = 0;
  std::vector<int> idx;
  std::vector<Node> nodes;
  static constexpr int LEAF = 12;

  void build(const float* p, int count) {
    pts = p; n = count; idx.resize(n); nodes.clear(); nodes.reserve(2 * n / LEAF + 16);
    for (int i = 0; i < n; i++) idx[i] = i;
    if (n > 0) build_rec(0, n);
  }
  int build_rec(int b, int e) {
    int id = (int)nodes.size();
    nodes.push_back(Node{-1, -1, b, e, -1, 0.f});
    if (e - b <= LEAF) return id;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = b; i < e; i++) for (int d = 0; d < 3; d++) { float v = pts[idx[i]*4+d]; lo[d] = std::min(lo[d], v); hi[d] = std::max(hi[d], v); }
    int dim = 0; float ext = hi[0]-lo[0];
    for (int d = 1; d < 3; d++) if (hi[d]-lo[d] > ext) { ext = hi[d]-lo[d]; dim = d; }
    if (!(ext > 0.f)) return id;  // all identical: keep as a (big) leaf
    int mid = (b + e) / 2;
    const float* P = pts;
    std::nth_element(idx.begin()+b, idx.begin()+mid, idx.begin()+e, [P, dim](int a, int c) {
      float va = P[a*4+dim], vc = P[c*4+dim]; return va < vc || (va == vc && a < c); });
    float split = pts[idx[mid]*4+dim];
    int l = build_rec(b, mid);
    int r = build_rec(mid, e);
    nodes[id].left = l; nodes[id].right = r; nodes[id].dim = dim; nodes[id].split = split;
    return id;
  }
  struct Cand { float d; int i; };
  static inline bool less(const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }
  void search(const float q[3], int k, Cand* heap, int& hn, int node) const {
    const Node& nd = nodes[node];
    if (nd.left < 0) {
      for (int t = nd.begin; t < nd.end; t++) {
        int j = idx[t];
        float dx = q[0]-pts[j*4+0], dy = q[1]-pts[j*4+1], dz = q[2]-pts[j*4+2];
        float d = dx*dx; d += dy*dy; d += dz*dz;
        Cand c{d, j};
        if (hn < k) { heap[hn++] = c; std::push_heap(heap, heap+hn, less); }
        else if (less(c, heap[0])) { std::pop_heap(heap, heap+hn, less); heap[hn-1] = c; std::push_heap(heap, heap+hn, less); }
      }
      return;
    }
    float diff = q[nd.dim] - nd.split;
    int first = diff < 0 ? nd.left : nd.right, second = diff < 0 ? nd.right : nd.left;
    search(q, k, heap, hn, first);
    float d2 = diff*diff;
    if (hn < k || d2 <= heap[0].d) search(q, k, heap, hn, second);
  }
  // returns number found (<= k), sorted ascending
  int knn(const float q[3], int k, int* out) const {
    std::vector<Cand> heap(k);
    int hn = 0;
    if (n > 0) search(q, k, heap.data(), hn, 0);
    std::sort_heap(heap.data(), heap.data()+hn, less);
    for (int i = 0; i < hn; i++) out[i] = heap[i].i;
    return hn;
  }
  // as knn(), also returning the float squared distances (nearestKSearch's k_sqr_distances)
  int knn_dist(const float q[3], int k, int* out, float* dist) const {
    std::vector<Cand> heap(k);
    int hn = 0;
    if (n > 0) search(q, k, heap.data(), hn, 0);
    std::sort_heap(heap.data(), heap.data()+hn, less);
    for (int i = 0; i < hn; i++) { out[i] = heap[i].i; dist[i] = heap[i].d; }
    return hn;
  }
};

}  // namespace
