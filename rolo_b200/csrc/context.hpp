// context.hpp — internal definitions shared by the translation units that implement the C-ABI (include/rolo_gicp.h):
// the process-global caching pool, device buffers, the registration context and its helpers. Not installed, not part of the ABI.
#pragma once
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/rolo_gicp.h"
#include "common.cuh"
#include "covariance.cuh"
#include "frontend.cuh"
#include "linearize.cuh"
#include "lsq_host.hpp"
#include "cloud_info.hpp"
#include "scan2map.cuh"
#include "scancontext.cuh"
#include "sc_host.hpp"
#include "s2m_math.cuh"
#include "voxelmap.cuh"

using namespace rolo;

extern std::mutex g_mutex;
extern std::atomic<unsigned long long> g_launches;   // own kernels launched (bench.py's gpu_launches claim)

// ------------------------------------------------------------------------------------------------
// process-global caching pool: the caller constructs a fresh registration object every frame
// (lidarOdometry.cpp:460), so device buffers, pinned blocks and streams must outlive contexts.
// ------------------------------------------------------------------------------------------------
struct Pool {
  std::multimap<std::pair<int, size_t>, void*> free_dev;    // (device, bytes) -> ptr
  std::multimap<size_t, void*> free_pinned;
  std::multimap<int, cudaStream_t> free_streams;
  std::map<void*, std::pair<int, size_t>> live_dev;
  std::map<void*, size_t> live_pinned;

  static size_t round_up(size_t b) {
    size_t g = 256;
    while (g < b) g <<= 1;
    if (b > (size_t)1 << 20) { const size_t q = (size_t)1 << 20; g = (b + q - 1) / q * q; }
    return g;
  }
  cudaError_t dev_alloc(int device, size_t bytes, void** out) {
    bytes = round_up(std::max<size_t>(bytes, 256));
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = free_dev.find({device, bytes});
    if (it != free_dev.end()) { *out = it->second; free_dev.erase(it); live_dev[*out] = {device, bytes}; return cudaSuccess; }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {   // release cached blocks and retry once
      for (auto& kv : free_dev) cudaFree(kv.second);
      free_dev.clear();
      e = cudaMalloc(out, bytes);
      if (e != cudaSuccess) return e;
    }
    live_dev[*out] = {device, bytes};
    return cudaSuccess;
  }
  void dev_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = live_dev.find(p);
    if (it == live_dev.end()) return;
    free_dev.insert({it->second, p});
    live_dev.erase(it);
  }
  cudaError_t pinned_alloc(size_t bytes, void** out) {
    bytes = round_up(std::max<size_t>(bytes, 256));
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = free_pinned.find(bytes);
    if (it != free_pinned.end()) { *out = it->second; free_pinned.erase(it); live_pinned[*out] = bytes; return cudaSuccess; }
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) return e;
    live_pinned[*out] = bytes;
    return cudaSuccess;
  }
  void pinned_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = live_pinned.find(p);
    if (it == live_pinned.end()) return;
    free_pinned.insert({it->second, p});
    live_pinned.erase(it);
  }
  cudaError_t stream_get(int device, cudaStream_t* s) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = free_streams.find(device);
    if (it != free_streams.end()) { *s = it->second; free_streams.erase(it); return cudaSuccess; }
    return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
  }
  void stream_put(int device, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_mutex);
    free_streams.insert({device, s});
  }
};
Pool& pool();

template <typename T>
struct DevBuf {
  T* p = nullptr; size_t cap = 0; int device = 0;
  cudaError_t ensure(int dev, size_t count) {
    if (count <= cap && p) return cudaSuccess;
    // growth: the old block goes back to the process-global pool, where any stream of any context may pick it up at once.
    // Kernels queued on it may still be running, so the device is drained first (growth is rare: buffers keep their size class
    // from frame to frame).
    if (p) { cudaSetDevice(device); cudaDeviceSynchronize(); }
    release();
    device = dev;
    void* q = nullptr;
    cudaError_t e = pool().dev_alloc(dev, std::max<size_t>(count, 1) * sizeof(T), &q);
    if (e != cudaSuccess) return e;
    p = (T*)q; cap = count;
    return cudaSuccess;
  }
  void release() { if (p) pool().dev_free(p); p = nullptr; cap = 0; }
};

// scratch of one covariance-estimation group (covariance.cuh: KnnScratch), sized for `total` points in `n_clouds` clouds
struct KnnBufs {
  DevBuf<unsigned long long> codes_a, codes_b;
  DevBuf<int> idx_a, idx_b, bbox, bucket, fail;
  DevBuf<float4> sorted;
  DevBuf<unsigned char> cub_temp;
  DevBuf<KnnCloud> clouds;
  size_t cub_bytes = 0;
  cudaError_t ensure(int dev, size_t total, int n_clouds) {
    cudaError_t e;
    if ((e = codes_a.ensure(dev, total)) != cudaSuccess) return e;
    if ((e = codes_b.ensure(dev, total)) != cudaSuccess) return e;
    if ((e = idx_a.ensure(dev, total)) != cudaSuccess) return e;
    if ((e = idx_b.ensure(dev, total)) != cudaSuccess) return e;
    if ((e = sorted.ensure(dev, total)) != cudaSuccess) return e;
    if ((e = bbox.ensure(dev, (size_t)n_clouds * 8)) != cudaSuccess) return e;
    if ((e = bucket.ensure(dev, (size_t)n_clouds * cov_bucket_table_entries())) != cudaSuccess) return e;
    if ((e = fail.ensure(dev, 2 + 2 * total)) != cudaSuccess) return e;
    if ((e = clouds.ensure(dev, n_clouds)) != cudaSuccess) return e;
    cub_bytes = cov_cub_temp_bytes((int)total, n_clouds);
    return cub_temp.ensure(dev, cub_bytes);
  }
  KnnScratch view() const {
    KnnScratch s;
    s.codes_a = codes_a.p; s.codes_b = codes_b.p; s.idx_a = idx_a.p; s.idx_b = idx_b.p; s.sorted = sorted.p;
    s.cub_temp = cub_temp.p; s.cub_temp_bytes = cub_bytes; s.bbox = bbox.p; s.bucket_start = bucket.p; s.fail = fail.p; s.clouds = clouds.p;
    return s;
  }
  void release() {
    codes_a.release(); codes_b.release(); idx_a.release(); idx_b.release(); bbox.release(); bucket.release(); fail.release();
    sorted.release(); cub_temp.release(); clouds.release();
  }
};

struct Cloud {
  DevBuf<float4> pts;
  DevBuf<float> cov6;
  int n = 0, n_pad = 0;
  bool has_cov = false;
  const void* host_identity = nullptr;   // the reference compares cloud pointers (rot_vgicp_impl.hpp:113,134)
  // target-sharded registration: covariances exist only for the points this rank owns (blocks of the Morton order dealt round-robin,
  // covariance.cuh: knn_owns); sorted_view = this cloud's Morton-ordered copy (xyz + original index) inside the covariance scratch,
  // valid until that scratch is reused
  int shard_rank = 0, shard_world = 1;
  const float4* sorted_view = nullptr;
  void clear() { n = 0; n_pad = 0; has_cov = false; host_identity = nullptr; shard_rank = 0; shard_world = 1; sorted_view = nullptr; }
  void release() { pts.release(); cov6.release(); clear(); }
};


// ------------------------------------------------------------------------------------------------
struct rg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t upload_stream = nullptr;   // H2D cloud uploads (== stream, or the pair's setup stream inside a batch)
  cudaEvent_t ev_upload = nullptr;
  bool side_pending = false;              // work queued on upload_stream that `stream` has not been ordered after yet
  mutable std::string err;

  // RotVGICP members (rot_vgicp_impl.hpp:20-39)
  int k_correspondences = 20;
  int regularization_method = RG_REG_PLANE;
  double voxel_resolution = 1.0;
  float lambda_ = 1.0f;
  int search_method = RG_DIRECT1;
  int voxel_mode = RG_ACC_ADDITIVE;
  int voxel_type = RG_VOXEL_POLAR;
  double polar_resolution[3] = {1.0, 0.0, 0.0};   // Vector3d::Identity() (reference quirk Q4: unusable until set)
  int num_threads = 0;

  host::LsqDriver lsq;

  Cloud src, tgt;

  // voxel map
  DevBuf<VoxSlot> table; unsigned mask = 0;
  DevBuf<VoxPayload> vox;
  DevBuf<double> sums;
  DevBuf<int> slot_of_point;
  DevBuf<int> counters;
  bool voxel_ok = false;
  int n_vox = -1;

  // covariance scratch
  DevBuf<unsigned long long> codes_in, codes_out;
  DevBuf<int> idx_in, idx_out, bbox;
  DevBuf<float4> sorted;
  DevBuf<unsigned char> cub_temp;
  DevBuf<int> knn_idx, bucket_start;
  KnnBufs knn;                          // group scratch of calculate_covariances (stand-alone contexts; a batch owns its own)

  // linearize
  DevBuf<int> corr; int corr_noff = 0;
  DevBuf<double> partials;
  DevBuf<unsigned> counter;
  DevBuf<double> result_dev;
  DevBuf<LinStatic> st_dev; bool st_ok = false;
  bool st_standalone_ok = false;        // own schedule + descriptor uploaded (stand-alone launches); a batch skips it
  double* result_host = nullptr;        // mapped pinned [RES_N] (owned; stand-alone contexts)
  unsigned* flag_host = nullptr;        // mapped pinned
  double* result_view = nullptr;        // where the kernel writes this pair's result (own block or the batch's)
  unsigned* flag_view = nullptr;
  int lin_grid = 0, lin_tiles = 0;      // persistent-CTA grid and tile count of the stand-alone launch
  DevBuf<int> sched_dev;
  double R_lin[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  bool have_corr = false;
  size_t num_corr = 0;

  DevBuf<unsigned char> staging;        // raw host cloud upload / misc
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool profile = false;
  double last_kernel_us = -1.0;

  rg_reduce_fn reduce_fn = nullptr; void* reduce_user = nullptr; int reduce_on_device = 0;
  bool deterministic = false;           // rg_set_deterministic: sort + fixed-order voxel sums instead of fp64 atomics
  DevBuf<int> det_flags;
  int tgt_rank = 0, tgt_world = 0;      // target-sharded setup (rg_set_target_shard): this rank's slice of the target's Morton order
  rg_voxel_merge_fn voxel_merge_fn = nullptr; void* voxel_merge_user = nullptr;
  DevBuf<unsigned long long> vox_keys;
  int shard_lo = 0, shard_hi = 0;       // evaluation range within the source cloud (0,0 = all)
  // fused NVLink reduction (rg_shard_connect)
  double* xchg = nullptr;               // own exchange buffer (plain cudaMalloc: exported through CUDA IPC)
  double* peer_xchg[XCHG_MAX_RANKS] = {nullptr};
  int p2p_rank = 0, p2p_world = 0;
  unsigned shard_seq = 0;

  // front end (range-image projection + feature extraction) scratch
  DevBuf<unsigned char> fe_raw;
  DevBuf<int> fe_owner, fe_row_count, fe_col, fe_start, fe_end, fe_counts, fe_picked, fe_label, fe_corner_cnt, fe_surf_cnt;
  DevBuf<float> fe_range_image, fe_prange, fe_curv;
  DevBuf<float4> fe_full, fe_extracted, fe_corner_ring, fe_surf_ring, fe_corner_out, fe_surf_out;
  // scan-to-map optimisation (backMapping.cpp:681-1054): clouds 0/1 = corner/surface feature MAP, 2/3 = corner/surface SCAN
  Cloud s2m_cloud[4];
  DevBuf<float4> s2m_sorted[2], s2m_coeff[2];
  DevBuf<unsigned long long> s2m_codes[2];
  DevBuf<int> s2m_bucket[2], s2m_bbox[2];
  DevBuf<unsigned char> s2m_flag[2];
  DevBuf<double> s2m_partial, s2m_result;
  DevBuf<unsigned> s2m_counter;
  DevBuf<int> s2m_nn;
  // getFitnessScore: Morton index of the target + per-source-point nearest neighbour
  DevBuf<float4> fit_sorted;
  DevBuf<unsigned long long> fit_codes;
  DevBuf<int> fit_bucket, fit_bbox, fit_idx;
  DevBuf<float> fit_d2, fit_T;
  // Scan Context place recognition (src/scancontext/Scancontext.cpp): descriptor kernel + host manager
  Cloud sc_cloud;
  DevBuf<int> sc_bins;
  DevBuf<double> sc_out;
  sc::Manager sc_mgr;
  bool s2m_map_ok = false, s2m_scan_ok = false, s2m_lin_ok = false;
  ProjDeskew fe_deskew = {0, 0, {0.f, 0.f, 0.f}, 0.1f, -1.0};
  DevBuf<unsigned char> fe_wire;        // packed PointXYZI records / uploaded message sections (cloud_info.hpp)
  bool fe_frame_ok = false;             // rg_front_end_frame products are resident
  int fe_n_scan = 0, fe_n[3] = {0, 0, 0};   // extracted, corner, surface counts of that frame

  rg_batch* batch = nullptr;            // non-null when owned by a batch
  int batch_index = 0;
};

struct rg_batch {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<rg_ctx*> pairs;
  DevBuf<LinStatic> st_dev;
  DevBuf<LinDyn> dyn_dev;
  LinDyn* dyn_host = nullptr;           // pinned staging [P]
  double* result_host = nullptr;        // mapped pinned [P][RES_N]
  unsigned* flag_host = nullptr;        // mapped pinned [P]
  std::vector<int> active;
  DevBuf<LinStatic> st_sub; DevBuf<LinDyn> dyn_sub;
  DevBuf<int> sched_dev; std::vector<int> sched_group; int sched_grid = 0, sched_tiles = 0;   // schedule of the last launched group
  KnnBufs knn;                          // covariance estimation of many clouds per launch chain
  DevBuf<VoxelBuild> vox_builds;        // voxel-map build descriptors of all pairs
  std::vector<cudaStream_t> setup_streams;   // covariance / voxel-map stages of different pairs overlap on these
  std::vector<cudaEvent_t> setup_events;
  cudaEvent_t ev_inputs = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double lin_us_sum = 0; int lin_launches = 0;
  // stage timing of the last rg_batch_align (CUDA events on the batch stream)
  cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};   // start, covariances done, voxel maps done, iterations done
  std::vector<cudaEvent_t> ev_knn;                                    // pairs around the search kernels of every covariance group
  int knn_groups = 0; double knn_queries = 0; double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::string err;
};


int set_err(const rg_ctx* c, int code, const std::string& msg);
int use_device(const rg_ctx* c);
int upload_cloud(rg_ctx* c, Cloud& cl, const float* host, size_t n, size_t stride_bytes);
int upload_cloud_device(rg_ctx* c, Cloud& cl, const float* dev, size_t n, cudaStream_t user_stream);

#define RG_CUDA(ctx, expr)                                                                          \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return set_err(ctx, RG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)
#define CHECK_CTX(c) do { if (!(c)) return set_err(nullptr, RG_ERR_INVALID, "null context"); } while (0)
