// ref_shim/pcl/search/kdtree.h — exact kNN stand-in for pcl::search::KdTree -> FLANN (L2_Simple<float>):
// k nearest under (float distance accumulated x,y,z; index) order, ascending. Test infrastructure only.
#ifndef ROLO_REF_SHIM_PCL_KDTREE
#define ROLO_REF_SHIM_PCL_KDTREE
#include <pcl/search/search.h>
namespace pcl { namespace search {
template <typename PointT> class KdTree : public Search<PointT> {
 public:
  using PointCloudConstPtr = typename PointCloud<PointT>::ConstPtr;
  struct Node { int left, right, begin, end, dim; float split; };
  struct Cand { float d; int i; };
  static bool less(const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }
  void setInputCloud(const PointCloudConstPtr& cloud) override {
    cloud_ = cloud; const int n = (int)cloud->size();
    idx_.resize(n); nodes_.clear(); for (int i = 0; i < n; i++) idx_[i] = i;
    if (n) build(0, n);
  }
  PointCloudConstPtr getInputCloud() const override { return cloud_; }
  int nearestKSearch(const PointT& p, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const override {
    std::vector<Cand> heap(k); int hn = 0;
    if (!nodes_.empty()) search(p.data, k, heap.data(), hn, 0);
    std::sort_heap(heap.data(), heap.data() + hn, less);
    k_indices.resize(hn); k_sqr_distances.resize(hn);
    for (int i = 0; i < hn; i++) { k_indices[i] = heap[i].i; k_sqr_distances[i] = heap[i].d; }
    return hn;
  }
 private:
  const float* pt(int i) const { return cloud_->points[i].data; }
  int build(int b, int e) {
    int id = (int)nodes_.size(); nodes_.push_back(Node{-1, -1, b, e, -1, 0.f});
    if (e - b <= 12) return id;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = b; i < e; i++) for (int d = 0; d < 3; d++) { float v = pt(idx_[i])[d]; lo[d] = std::min(lo[d], v); hi[d] = std::max(hi[d], v); }
    int dim = 0; float ext = hi[0] - lo[0];
    for (int d = 1; d < 3; d++) if (hi[d] - lo[d] > ext) { ext = hi[d] - lo[d]; dim = d; }
    if (!(ext > 0.f)) return id;
    int mid = (b + e) / 2;
    std::nth_element(idx_.begin() + b, idx_.begin() + mid, idx_.begin() + e, [this, dim](int a, int c) { float va = pt(a)[dim], vc = pt(c)[dim]; return va < vc || (va == vc && a < c); });
    float split = pt(idx_[mid])[dim];
    int l = build(b, mid), r = build(mid, e);
    nodes_[id].left = l; nodes_[id].right = r; nodes_[id].dim = dim; nodes_[id].split = split;
    return id;
  }
  void search(const float* q, int k, Cand* heap, int& hn, int node) const {
    const Node& nd = nodes_[node];
    if (nd.left < 0) {
      for (int t = nd.begin; t < nd.end; t++) {
        int j = idx_[t]; const float* p = pt(j);
        float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
        float d = dx * dx; d += dy * dy; d += dz * dz;
        Cand c{d, j};
        if (hn < k) { heap[hn++] = c; std::push_heap(heap, heap + hn, less); }
        else if (less(c, heap[0])) { std::pop_heap(heap, heap + hn, less); heap[hn - 1] = c; std::push_heap(heap, heap + hn, less); }
      }
      return;
    }
    float diff = q[nd.dim] - nd.split;
    search(q, k, heap, hn, diff < 0 ? nd.left : nd.right);
    if (hn < k || diff * diff <= heap[0].d) search(q, k, heap, hn, diff < 0 ? nd.right : nd.left);
  }
  PointCloudConstPtr cloud_;
  std::vector<int> idx_;
  std::vector<Node> nodes_;
};
}}  // namespace pcl::search
#endif
