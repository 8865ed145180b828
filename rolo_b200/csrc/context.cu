// context.cu — registration context, device-memory pool, host drivers and the C-ABI (include/rolo_gicp.h).
//
// The product path is CUDA-only: every entry point fails with RG_ERR_NO_DEVICE / RG_ERR_CUDA when no
// B200-class device is usable; there is no CPU fallback anywhere in this library.

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

namespace {

std::mutex g_mutex;
std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};   // own kernels launched (bench.py's gpu_launches claim)
std::atomic<unsigned> g_seq{0};   // launch sequence numbers written to the mapped completion flags

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
Pool& pool() { static Pool* p = new Pool; return *p; }   // intentionally leaked: safe at process exit

template <typename T>
struct DevBuf {
  T* p = nullptr; size_t cap = 0; int device = 0;
  cudaError_t ensure(int dev, size_t count) {
    if (count <= cap && p) return cudaSuccess;
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

__global__ void extract_points_kernel(const unsigned char* __restrict__ raw, size_t stride, int n, int n_pad, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  float4 p = make_float4(0.f, 0.f, 0.f, 1.f);
  if (i < n) {
    const float* s = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    p.x = s[0]; p.y = s[1]; p.z = s[2];
  }
  out[i] = p;
}
__global__ void copy_points_kernel(const float4* __restrict__ in, int n, int n_pad, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  float4 p = make_float4(0.f, 0.f, 0.f, 1.f);
  if (i < n) { p = in[i]; p.w = 1.f; }
  out[i] = p;
}
__global__ void cov_from_double_kernel(const double* __restrict__ cov16, int n, int n_pad, float* __restrict__ cov6) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const size_t base = (size_t)(i / TILE_PTS) * COV_PLANES * TILE_PTS + (i % TILE_PTS);
  float v[6] = {0, 0, 0, 0, 0, 0};
  if (i < n) {
    const double* c = cov16 + (size_t)i * 16;
    v[0] = (float)c[0]; v[1] = (float)c[1]; v[2] = (float)c[2]; v[3] = (float)c[5]; v[4] = (float)c[6]; v[5] = (float)c[10];
  }
  for (int t = 0; t < 6; t++) cov6[base + (size_t)t * TILE_PTS] = v[t];
}
__global__ void cov_to_double_kernel(const float* __restrict__ cov6, int n, double* __restrict__ cov16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t base = (size_t)(i / TILE_PTS) * COV_PLANES * TILE_PTS + (i % TILE_PTS);
  double v[6];
  for (int t = 0; t < 6; t++) v[t] = (double)cov6[base + (size_t)t * TILE_PTS];
  double* c = cov16 + (size_t)i * 16;
  c[0] = v[0]; c[1] = v[1]; c[2] = v[2]; c[3] = 0;
  c[4] = v[1]; c[5] = v[3]; c[6] = v[4]; c[7] = 0;
  c[8] = v[2]; c[9] = v[4]; c[10] = v[5]; c[11] = 0;
  c[12] = c[13] = c[14] = c[15] = 0;
}
__global__ void transform_points_kernel(const float4* __restrict__ in, int n, const float* __restrict__ T /*16*/, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  float4 o;
  o.x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
  o.y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
  o.z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
  o.w = 1.f;
  out[i] = o;
}
__global__ void table_dump_kernel(const VoxSlot* __restrict__ table, unsigned cap, const VoxPayload* __restrict__ vox,
                                  unsigned long long* __restrict__ keys, VoxPayload* __restrict__ out) {
  const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  const VoxSlot sl = table[s];
  if (sl.key == KEY_EMPTY) return;
  keys[sl.id] = sl.key;
  out[sl.id] = vox[sl.id];
}

struct Cloud {
  DevBuf<float4> pts;
  DevBuf<float> cov6;
  int n = 0, n_pad = 0;
  bool has_cov = false;
  const void* host_identity = nullptr;   // the reference compares cloud pointers (rot_vgicp_impl.hpp:113,134)
  void clear() { n = 0; n_pad = 0; has_cov = false; host_identity = nullptr; }
  void release() { pts.release(); cov6.release(); clear(); }
};

}  // namespace

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

  // linearize
  DevBuf<int> corr; int corr_noff = 0;
  DevBuf<double> partials;
  DevBuf<unsigned> counter;
  DevBuf<double> result_dev;
  DevBuf<LinStatic> st_dev; bool st_ok = false;
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
  std::vector<cudaStream_t> setup_streams;   // covariance / voxel-map stages of different pairs overlap on these
  std::vector<cudaEvent_t> setup_events;
  cudaEvent_t ev_inputs = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double lin_us_sum = 0; int lin_launches = 0;
  std::string err;
};

namespace {

int set_err(const rg_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  { std::lock_guard<std::mutex> lk(g_mutex); g_last_error = msg; }
  return code;
}
#define RG_CUDA(ctx, expr)                                                                          \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return set_err(ctx, RG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)

int use_device(const rg_ctx* cc) {
  rg_ctx* c = const_cast<rg_ctx*>(cc);
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return set_err(c, RG_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  if (c->side_pending && c->upload_stream && c->stream != c->upload_stream) {   // lazily order `stream` after side-stream work of this context
    cudaStreamWaitEvent(c->stream, c->ev_upload, 0);
    c->side_pending = false;
  }
  return RG_OK;
}

int noff_of(int method) { return method == RG_DIRECT1 ? 1 : (method == RG_DIRECT7 ? 7 : (method == RG_DIRECT27 ? 27 : 0)); }

GridParams make_grid(const rg_ctx* c) {
  GridParams g;
  g.voxel_type = c->voxel_type;
  for (int i = 0; i < 3; i++) {
    g.res[i] = (c->voxel_type == RG_VOXEL_POLAR) ? c->polar_resolution[i] : c->voxel_resolution;
    g.inv_res[i] = 1.0 / g.res[i];
  }
  return g;
}

int upload_cloud(rg_ctx* c, Cloud& cl, const float* host, size_t n, size_t stride_bytes) {
  if (use_device(c)) return RG_ERR_CUDA;
  if (n > 0 && !host) return set_err(c, RG_ERR_INVALID, "null cloud pointer");
  if (n > 0 && (stride_bytes < 12 || stride_bytes % 4)) return set_err(c, RG_ERR_INVALID, "stride_bytes must be a multiple of 4 and >= 12");
  if (n > (size_t)1 << 30) return set_err(c, RG_ERR_INVALID, "cloud too large");
  const int n_pad = (int)((n + TILE_PTS - 1) / TILE_PTS * TILE_PTS);
  RG_CUDA(c, cl.pts.ensure(c->device, std::max(n_pad, TILE_PTS)));
  RG_CUDA(c, cl.cov6.ensure(c->device, (size_t)std::max(n_pad, TILE_PTS) * COV_PLANES));
  cl.n = (int)n; cl.n_pad = n_pad; cl.has_cov = false; cl.host_identity = host;
  if (n == 0) return RG_OK;
  cudaStream_t us = c->upload_stream ? c->upload_stream : c->stream;
  if (us != c->stream) {   // order after whatever the context queued before (e.g. the previous frame still using these buffers)
    RG_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));
    RG_CUDA(c, cudaStreamWaitEvent(us, c->ev_upload, 0));
  }
  if (c->staging.cap < n * stride_bytes) RG_CUDA(c, cudaStreamSynchronize(us));   // about to re-allocate a buffer in-flight work may read
  RG_CUDA(c, c->staging.ensure(c->device, n * stride_bytes));
  // pageable memory: the runtime stages the copy before returning; page-locked memory: truly asynchronous, the caller keeps
  // the cloud alive until the next solving call returns (pcl::Registration holds the cloud's shared_ptr, as in the reference)
  RG_CUDA(c, cudaMemcpyAsync(c->staging.p, host, n * stride_bytes, cudaMemcpyHostToDevice, us));
  g_launches++;
  extract_points_kernel<<<(n_pad + 255) / 256, 256, 0, us>>>(c->staging.p, stride_bytes, (int)n, n_pad, cl.pts.p);
  RG_CUDA(c, cudaGetLastError());
  if (us != c->stream) {
    RG_CUDA(c, cudaEventRecord(c->ev_upload, us));
    RG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_upload, 0));
  }
  return RG_OK;
}

int upload_cloud_device(rg_ctx* c, Cloud& cl, const float* dev, size_t n, cudaStream_t user_stream) {
  if (use_device(c)) return RG_ERR_CUDA;
  if (n > 0 && !dev) return set_err(c, RG_ERR_INVALID, "null cloud pointer");
  const int n_pad = (int)((n + TILE_PTS - 1) / TILE_PTS * TILE_PTS);
  RG_CUDA(c, cl.pts.ensure(c->device, std::max(n_pad, TILE_PTS)));
  RG_CUDA(c, cl.cov6.ensure(c->device, (size_t)std::max(n_pad, TILE_PTS) * COV_PLANES));
  cl.n = (int)n; cl.n_pad = n_pad; cl.has_cov = false; cl.host_identity = dev;
  if (n == 0) return RG_OK;
  if (user_stream && user_stream != c->stream) {   // order after the producer stream
    cudaEvent_t ev;
    RG_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaEventRecord(ev, user_stream);
    cudaStreamWaitEvent(c->stream, ev, 0);
    cudaEventDestroy(ev);
  }
  g_launches++;
  copy_points_kernel<<<(n_pad + 255) / 256, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(dev), (int)n, n_pad, cl.pts.p);
  RG_CUDA(c, cudaGetLastError());
  return RG_OK;
}

int calc_cov(rg_ctx* c, Cloud& cl) {
  if (use_device(c)) return RG_ERR_CUDA;
  if (cl.n == 0) return set_err(c, RG_ERR_INVALID, "calculate_covariances: empty cloud");
  if (c->regularization_method < RG_REG_NONE || c->regularization_method > RG_REG_PLANE_S)
    return set_err(c, RG_ERR_UNSUPPORTED, "here must not be reached (unknown regularization method; reference abort()s, rot_vgicp_impl.hpp:472-474)");
  if (c->k_correspondences < 1 || c->k_correspondences > 64) return set_err(c, RG_ERR_INVALID, "k_correspondences must be in [1,64]");
  if (cl.n < c->k_correspondences)
    return set_err(c, RG_ERR_INVALID, "cloud has fewer points than k_correspondences (reference reads uninitialised neighbours here, rot_vgicp_impl.hpp:438-444)");
  const int n = cl.n;
  RG_CUDA(c, c->codes_in.ensure(c->device, n)); RG_CUDA(c, c->codes_out.ensure(c->device, n));
  RG_CUDA(c, c->idx_in.ensure(c->device, n)); RG_CUDA(c, c->idx_out.ensure(c->device, n));
  RG_CUDA(c, c->sorted.ensure(c->device, n)); RG_CUDA(c, c->bbox.ensure(c->device, 8)); RG_CUDA(c, c->bucket_start.ensure(c->device, cov_bucket_table_entries()));
  const size_t tb = cov_cub_temp_bytes(n);
  RG_CUDA(c, c->cub_temp.ensure(c->device, tb));
  RG_CUDA(c, c->knn_idx.ensure(c->device, (size_t)n * c->k_correspondences));
  CovLaunch L;
  L.pts = cl.pts.p; L.n = n; L.k = c->k_correspondences; L.method = c->regularization_method;
  L.cov6 = cl.cov6.p; L.knn = c->knn_idx.p;
  L.s.codes_in = c->codes_in.p; L.s.codes_out = c->codes_out.p; L.s.idx_in = c->idx_in.p; L.s.idx_out = c->idx_out.p;
  L.s.sorted = c->sorted.p; L.s.cub_temp = c->cub_temp.p; L.s.cub_temp_bytes = tb; L.s.bbox = c->bbox.p; L.s.bucket_start = c->bucket_start.p;
  // zero the padding rows of the covariance tiles (only the last tile has any)
  if (cl.n_pad > n) RG_CUDA(c, cudaMemsetAsync(cl.cov6.p + (size_t)(cl.n_pad / TILE_PTS - 1) * COV_PLANES * TILE_PTS, 0, sizeof(float) * COV_PLANES * TILE_PTS, c->stream));
  RG_CUDA(c, launch_covariances(L, c->stream));
  g_launches += 7;   // bbox_init, bbox, morton, gather, bucket table, knn, cov (cub's radix-sort kernels not counted)
  cl.has_cov = true;
  return RG_OK;
}

int check_grid(rg_ctx* c) {
  if (c->voxel_type == RG_VOXEL_POLAR) {
    for (int i = 0; i < 3; i++)
      if (!(c->polar_resolution[i] > 0.0))
        return set_err(c, RG_ERR_INVALID, "polar resolution not set: call setPolarResolution/setResolution first (the reference default (1,0,0) divides by zero, rot_vgicp_impl.hpp:38, vmp_voxel.hpp:210)");
  } else if (!(c->voxel_resolution > 0.0)) {
    return set_err(c, RG_ERR_INVALID, "voxel resolution must be positive");
  }
  if (noff_of(c->search_method) == 0)
    return set_err(c, RG_ERR_UNSUPPORTED, "unsupported neighbor search method (reference abort()s, vmp_voxel.hpp:16-18)");
  return RG_OK;
}

int build_voxelmap(rg_ctx* c) {
  if (use_device(c)) return RG_ERR_CUDA;
  if (int rc = check_grid(c)) return rc;
  if (c->tgt.n == 0) return set_err(c, RG_ERR_INVALID, "create_voxelmap: no target cloud");
  if (!c->tgt.has_cov) return set_err(c, RG_ERR_INVALID, "create_voxelmap: target covariances missing");
  const int n = c->tgt.n;
  size_t cap = 1024;
  while (cap < (size_t)2 * n) cap <<= 1;
  RG_CUDA(c, c->table.ensure(c->device, cap));
  c->mask = (unsigned)(cap - 1);
  RG_CUDA(c, c->vox.ensure(c->device, n));
  RG_CUDA(c, c->sums.ensure(c->device, (size_t)n * 10));
  RG_CUDA(c, c->slot_of_point.ensure(c->device, n));
  RG_CUDA(c, c->counters.ensure(c->device, 2));
  VoxelBuild b;
  b.pts = c->tgt.pts.p; b.cov6 = c->tgt.cov6.p; b.n = n; b.grid = make_grid(c);
  b.table = c->table.p; b.mask = c->mask; b.slot_of_point = c->slot_of_point.p; b.counters = c->counters.p;
  b.sums = c->sums.p; b.vox = c->vox.p;
  RG_CUDA(c, launch_voxel_build(b, c->stream));
  g_launches += 3;
  c->voxel_ok = true; c->n_vox = -1; c->st_ok = false; c->have_corr = false;
  return RG_OK;
}

int fetch_voxel_count(rg_ctx* c) {
  if (!c->voxel_ok) return 0;
  if (c->n_vox >= 0) return c->n_vox;
  int h[2] = {0, 0};
  if (cudaMemcpyAsync(h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
  if (h[1] != 0) { set_err(c, RG_ERR_INVALID, "voxel coordinate out of the +-2^20 key range (or table overflow)"); return -1; }
  c->n_vox = h[0];
  return c->n_vox;
}

int sm_count(int device) {
  static std::map<int, int> cache;
  std::lock_guard<std::mutex> lk(g_mutex);
  auto it = cache.find(device);
  if (it != cache.end()) return it->second;
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  cache[device] = n;
  return n;
}

int lin_slots(int device) { return sm_count(device) * 2; }   // 2 resident CTAs per SM (launch bounds of the fused kernel)

// make sure the static descriptor + scratch for the fused kernel exist
int prepare_linearize(rg_ctx* c) {
  if (use_device(c)) return RG_ERR_CUDA;
  if (int rc = check_grid(c)) return rc;
  if (c->src.n == 0) return set_err(c, RG_ERR_INVALID, "no source cloud");
  if (!c->src.has_cov) return set_err(c, RG_ERR_INVALID, "source covariances missing");
  if (!c->voxel_ok) { if (int rc = build_voxelmap(c)) return rc; }   // lazily, as rot_vgicp_impl.hpp:227-230
  const int noff = noff_of(c->search_method);
  if (c->st_ok && c->corr_noff == noff) return RG_OK;
  const int n_pairs = c->batch ? (int)c->batch->pairs.size() : 1;
  const int lo = (c->shard_hi > c->shard_lo) ? c->shard_lo : 0;
  const int n_eval = (c->shard_hi > c->shard_lo) ? (c->shard_hi - c->shard_lo) : c->src.n;
  (void)n_pairs;
  {   // stand-alone launch shape: one pair, persistent CTAs
    std::vector<int> sched;
    const int ntiles = (n_eval + TILE_PTS - 1) / TILE_PTS;
    c->lin_grid = std::max(1, std::min(ntiles, lin_slots(c->device)));
    build_schedule(&n_eval, 1, c->lin_grid, sched, c->lin_tiles);
    RG_CUDA(c, c->sched_dev.ensure(c->device, sched.size()));
    RG_CUDA(c, cudaMemcpyAsync(c->sched_dev.p, sched.data(), sched.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    RG_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  RG_CUDA(c, c->corr.ensure(c->device, (size_t)n_eval * noff));
  c->corr_noff = noff; c->have_corr = false;
  RG_CUDA(c, c->partials.ensure(c->device, (size_t)lin_slots(c->device) * RES_N));   // any launch grid fits
  if (!c->counter.p) { RG_CUDA(c, c->counter.ensure(c->device, 4)); RG_CUDA(c, cudaMemsetAsync(c->counter.p, 0, 4 * sizeof(unsigned), c->stream)); }
  RG_CUDA(c, c->result_dev.ensure(c->device, RES_N));
  RG_CUDA(c, c->st_dev.ensure(c->device, 1));
  if (!c->batch) {
    if (!c->result_host) {
      void* p = nullptr;
      RG_CUDA(c, pool().pinned_alloc(RES_N * sizeof(double) + 64, &p));
      c->result_host = (double*)p;
      c->flag_host = (unsigned*)((char*)p + RES_N * sizeof(double));
      *c->flag_host = 0;
    }
  }
  LinStatic st;
  st.src = c->src.pts.p + lo; st.cov6 = c->src.cov6.p + (size_t)(lo / TILE_PTS) * COV_PLANES * TILE_PTS; st.n = n_eval; st.mask = c->mask; st.table = c->table.p; st.vox = c->vox.p;
  st.corr = c->corr.p; st.partials = c->partials.p; st.counter = c->counter.p;
  const bool to_host = !c->batch && !(c->reduce_fn && c->reduce_on_device);
  if (c->batch) { c->result_view = c->batch->result_host + (size_t)c->batch_index * RES_N; c->flag_view = c->batch->flag_host + c->batch_index; }
  else { c->result_view = c->result_host; c->flag_view = c->flag_host; }
  st.result = to_host || c->batch ? c->result_view : c->result_dev.p;
  st.done_flag = to_host || c->batch ? c->flag_view : nullptr;
  st.grid = make_grid(c);
  st.shard_rank = c->p2p_rank; st.shard_world = (c->p2p_world > 1 && !c->batch) ? c->p2p_world : 0;
  for (int p = 0; p < XCHG_MAX_RANKS; p++) st.peer_xchg[p] = c->peer_xchg[p];
  RG_CUDA(c, cudaMemcpyAsync(c->st_dev.p, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
  if (c->batch) RG_CUDA(c, cudaMemcpyAsync(c->batch->st_dev.p + c->batch_index, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));   // `st` is a stack object
  c->st_ok = true;
  return RG_OK;
}

void fill_dyn_pose(LinDyn& d, const double* T, const double* Rm) {
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) d.R[i * 3 + j] = T[i * 4 + j]; d.t[i] = T[i * 4 + 3]; }
  for (int i = 0; i < 9; i++) d.Rm[i] = Rm[i];
  d.ct[0] = d.ct[1] = d.ct[2] = 0.0; d.lam_c = 0.0; d.h_scale = 1.0; d.b_ct_scale = 0.0; d.shard_seq = 0;
}

// Build the launch description of one evaluation request; returns mode/refresh.
int make_dyn(rg_ctx* c, const host::Request& r, const double* t3_guess, const double* /*t3_last*/, double dtn, double dtn1,
             LinDyn& d, int& mode, int& refresh) {
  double T[16];
  switch (r.kind) {
    case host::REQ_LIN6:
    case host::REQ_LIN3: {
      double Rm[9];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rm[i * 3 + j] = r.pose[i * 4 + j];
      fill_dyn_pose(d, r.pose, Rm);
      std::memcpy(c->R_lin, Rm, sizeof(Rm));
      mode = (r.kind == host::REQ_LIN6) ? LIN_SE3 : LIN_SO3; refresh = 1;
      return RG_OK;
    }
    case host::REQ_ERROR:
      if (!c->have_corr) return set_err(c, RG_ERR_INVALID, "compute_error before any linearize (no correspondences)");
      fill_dyn_pose(d, r.pose, c->R_lin);
      mode = LIN_COST; refresh = 0;
      return RG_OK;
    case host::REQ_T3_LIN:
    case host::REQ_T_ERROR: {
      if (!c->have_corr) return set_err(c, RG_ERR_INVALID, "translation stage before any linearize (no correspondences)");
      host::mat4_identity(T);
      T[3] = r.t[0]; T[7] = r.t[1]; T[11] = r.t[2];
      fill_dyn_pose(d, T, c->R_lin);
      // ct = ((p - g) - (p + t))/dtn - last_transform/dtn1 (rot_vgicp_impl.hpp:544-548 / 642-646).
      // Reference quirk Q3: `last_transform.matrix().col(3).head<3>() = last_t0` indexes column 3 of a 4x1
      // vector, so last_t0 never reaches last_transform: it stays Zero in t3_linearize (impl.hpp:539) and
      // Vector4d::Identity() = (1,0,0,0) in compute_t_error (impl.hpp:637). Replicated.
      const double lx = (r.kind == host::REQ_T_ERROR) ? 1.0 : 0.0;
      d.ct[0] = -(t3_guess[0] + r.t[0]) / dtn - lx / dtn1;
      d.ct[1] = -(t3_guess[1] + r.t[1]) / dtn;
      d.ct[2] = -(t3_guess[2] + r.t[2]) / dtn;
      // `lambda_/pt_size` is float / size_t in the reference: a float division, promoted to double afterwards (impl.hpp:555,582)
      d.lam_c = (double)(c->lambda_ / (float)c->num_corr);
      d.h_scale = 1.0 + d.lam_c / (dtn * dtn);
      d.b_ct_scale = d.lam_c / dtn;
      mode = (r.kind == host::REQ_T3_LIN) ? LIN_SE3 : LIN_COST; refresh = 0;
      return RG_OK;
    }
  }
  return set_err(c, RG_ERR_INVALID, "bad request");
}

int wait_flag(rg_ctx* c, volatile unsigned* flag, unsigned seq, cudaStream_t stream) {
  const auto t0 = std::chrono::steady_clock::now();
  unsigned spins = 0;
  while (*flag != seq) {
    if ((++spins & 0x3FFF) == 0) {
      cudaError_t q = cudaStreamQuery(stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) return set_err(c, RG_ERR_CUDA, std::string("kernel failed: ") + cudaGetErrorString(q));
      if (q == cudaSuccess && *flag != seq) {
        // stream drained but flag not set: give the mapped write a moment, then fail
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) return set_err(c, RG_ERR_CUDA, "result flag never arrived");
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) return set_err(c, RG_ERR_CUDA, "timeout waiting for the fused kernel");
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return RG_OK;
}

// one evaluation on a stand-alone context; res receives RES_N doubles
int evaluate_single(rg_ctx* c, const host::Request& r, const double* g3, const double* l3, double dtn, double dtn1, double* res) {
  if (int rc = prepare_linearize(c)) return rc;
  LinDyn d; int mode = 0, refresh = 0;
  if (int rc = make_dyn(c, r, g3, l3, dtn, dtn1, d, mode, refresh)) return rc;
  d.seq = ++g_seq;
  d.shard_seq = (c->p2p_world > 1) ? ++c->shard_seq : 0u;
  LinLaunch L;
  L.mode = mode; L.refresh = refresh; L.noff = c->corr_noff; L.ct = (r.kind == host::REQ_T3_LIN || r.kind == host::REQ_T_ERROR) ? 1 : 0; L.n_pairs = 1; L.grid = c->lin_grid; L.total_tiles = c->lin_tiles;
  L.st = c->st_dev.p; L.dyn = nullptr; L.sched = c->sched_dev.p;
  if (c->profile) cudaEventRecord(c->ev0, c->stream);
  RG_CUDA(c, launch_linearize(L, &d, c->stream));
  g_launches++;
  if (c->profile) cudaEventRecord(c->ev1, c->stream);
  const bool on_dev = c->reduce_fn && c->reduce_on_device && !c->batch;
  if (on_dev) {
    if (int rc = c->reduce_fn(c->reduce_user, c->result_dev.p, RES_N, 1, (void*)c->stream)) return set_err(c, RG_ERR_CUDA, "reduce hook failed (" + std::to_string(rc) + ")");
    RG_CUDA(c, cudaMemcpyAsync(res, c->result_dev.p, RES_N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RG_CUDA(c, cudaStreamSynchronize(c->stream));
  } else {
    if (int rc = wait_flag(c, c->flag_view, d.seq, c->stream)) return rc;
    std::memcpy(res, c->result_view, RES_N * sizeof(double));
    if (c->reduce_fn) if (int rc = c->reduce_fn(c->reduce_user, res, RES_N, 0, nullptr)) return set_err(c, RG_ERR_CUDA, "reduce hook failed (" + std::to_string(rc) + ")");
  }
  if (c->profile) { cudaEventSynchronize(c->ev1); float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->last_kernel_us = ms * 1e3; }
  if (c->p2p_world > 1 && res[28] != res[28]) return set_err(c, RG_ERR_CUDA, "sharded registration: a peer rank never delivered its partial sums (bounded wait expired)");
  if (refresh) { c->have_corr = true; c->num_corr = (size_t)(res[28] + 0.5); }
  return RG_OK;
}

int ensure_covariances(rg_ctx* c) {
  // rot_vgicp_impl.hpp:153-158: recompute when the stored count differs from the cloud size
  if (!c->src.has_cov) { if (int rc = calc_cov(c, c->src)) return rc; }
  if (!c->tgt.has_cov) { if (int rc = calc_cov(c, c->tgt)) return rc; }
  return RG_OK;
}

int run_driver(rg_ctx* c, const double* g3, const double* l3, double dtn, double dtn1) {
  host::Request r;
  double res[RES_N], H[36], b[6], cost;
  while (c->lsq.want(r)) {
    if (int rc = evaluate_single(c, r, g3, l3, dtn, dtn1, res)) return rc;
    const int ndof = (r.kind == host::REQ_LIN3) ? 3 : 6;
    if (r.kind == host::REQ_ERROR || r.kind == host::REQ_T_ERROR) { cost = res[27]; c->lsq.feed(nullptr, nullptr, cost); }
    else { unpack_result(res, ndof, H, b, &cost); c->lsq.feed(H, b, cost); }
  }
  return RG_OK;
}

}  // namespace

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

int rg_abi_version(void) { return ROLO_GICP_ABI_VERSION; }

int rg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* rg_last_error(const rg_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_mutex);
  copy = g_last_error;
  return copy.c_str();
}

static int create_ctx(int device, cudaStream_t shared_stream, rg_ctx** out) {
  if (!out) return set_err(nullptr, RG_ERR_INVALID, "rg_create: null out");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(nullptr, RG_ERR_NO_DEVICE, "no CUDA device: the B200 core has no CPU fallback");
  }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= ndev) return set_err(nullptr, RG_ERR_INVALID, "device index out of range");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return set_err(nullptr, RG_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  rg_ctx* c = new rg_ctx;
  c->device = device;
  if (shared_stream) { c->stream = shared_stream; c->own_stream = false; }
  else if ((e = pool().stream_get(device, &c->stream)) != cudaSuccess) { delete c; return set_err(nullptr, RG_ERR_CUDA, std::string("stream: ") + cudaGetErrorString(e)); }
  cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
  cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming);
  c->upload_stream = c->stream;
  *out = c;
  return RG_OK;
}

int rg_create(int device, rg_ctx** out) { return create_ctx(device, nullptr, out); }

void rg_destroy(rg_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->upload_stream && c->upload_stream != c->stream) cudaStreamSynchronize(c->upload_stream);
  cudaStreamSynchronize(c->stream);
  c->src.release(); c->tgt.release();
  c->table.release(); c->vox.release(); c->sums.release(); c->slot_of_point.release(); c->counters.release();
  c->codes_in.release(); c->codes_out.release(); c->idx_in.release(); c->idx_out.release(); c->bbox.release(); c->sorted.release(); c->cub_temp.release(); c->knn_idx.release(); c->bucket_start.release();
  c->fe_raw.release(); c->fe_owner.release(); c->fe_row_count.release(); c->fe_col.release(); c->fe_start.release(); c->fe_end.release();
  c->fe_counts.release(); c->fe_picked.release(); c->fe_label.release(); c->fe_corner_cnt.release(); c->fe_surf_cnt.release();
  c->fe_range_image.release(); c->fe_prange.release(); c->fe_curv.release(); c->fe_full.release(); c->fe_extracted.release();
  c->fe_corner_ring.release(); c->fe_surf_ring.release(); c->fe_corner_out.release(); c->fe_surf_out.release(); c->fe_wire.release();
  for (int k = 0; k < 4; k++) c->s2m_cloud[k].release();
  for (int k = 0; k < 2; k++) { c->s2m_sorted[k].release(); c->s2m_coeff[k].release(); c->s2m_codes[k].release(); c->s2m_bucket[k].release(); c->s2m_bbox[k].release(); c->s2m_flag[k].release(); }
  c->s2m_partial.release(); c->s2m_result.release(); c->s2m_counter.release(); c->s2m_nn.release();
  c->sc_cloud.release(); c->sc_bins.release(); c->sc_out.release();
  c->fit_sorted.release(); c->fit_codes.release(); c->fit_bucket.release(); c->fit_bbox.release(); c->fit_idx.release(); c->fit_d2.release(); c->fit_T.release();
  c->corr.release(); c->partials.release(); c->counter.release(); c->result_dev.release(); c->st_dev.release(); c->staging.release(); c->sched_dev.release();
  for (int p = 0; p < XCHG_MAX_RANKS; p++) if (c->peer_xchg[p] && c->peer_xchg[p] != c->xchg) cudaIpcCloseMemHandle(c->peer_xchg[p]);
  if (c->xchg) cudaFree(c->xchg);
  if (c->result_host) pool().pinned_free(c->result_host);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_upload) cudaEventDestroy(c->ev_upload);
  if (c->own_stream && c->stream) pool().stream_put(c->device, c->stream);
  delete c;
}

#define CHECK_CTX(c) do { if (!(c)) return set_err(nullptr, RG_ERR_INVALID, "null context"); } while (0)

int rg_set_resolution(rg_ctx* c, double r) { CHECK_CTX(c); c->voxel_resolution = r; c->voxel_type = RG_VOXEL_UNIFORM; c->voxel_ok = false; c->st_ok = false; return RG_OK; }
int rg_set_polar_resolution(rg_ctx* c, double th, double ph, double r) {
  CHECK_CTX(c); c->polar_resolution[0] = th; c->polar_resolution[1] = ph; c->polar_resolution[2] = r; c->voxel_type = RG_VOXEL_POLAR; c->voxel_ok = false; c->st_ok = false; return RG_OK;
}
int rg_set_voxel_accumulation_mode(rg_ctx* c, int m) {
  CHECK_CTX(c);
  if (m < RG_ACC_ADDITIVE || m > RG_ACC_MULTIPLICATIVE) return set_err(c, RG_ERR_INVALID, "bad accumulation mode");
  c->voxel_mode = m;   // all three modes use the additive voxel in the reference (vmp_voxel.hpp:176-185, quirk Q6)
  return RG_OK;
}
int rg_set_neighbor_search_method(rg_ctx* c, int m) { CHECK_CTX(c); c->search_method = m; c->st_ok = false; return RG_OK; }
int rg_set_num_threads(rg_ctx* c, int n) { CHECK_CTX(c); c->num_threads = n; return RG_OK; }
int rg_set_correspondence_randomness(rg_ctx* c, int k) { CHECK_CTX(c); c->k_correspondences = k; return RG_OK; }
int rg_set_regularization_method(rg_ctx* c, int m) { CHECK_CTX(c); c->regularization_method = m; return RG_OK; }
int rg_set_optimizer_type(rg_ctx* c, int t) {
  CHECK_CTX(c);
  if (t < RG_OPT_GAUSS_NEWTON || t > RG_OPT_SO3_LM) t = RG_OPT_LM;   // lsq.hpp:204 default branch
  c->lsq.cfg.optimizer = t; return RG_OK;
}
int rg_set_rotation_epsilon(rg_ctx* c, double e) { CHECK_CTX(c); c->lsq.cfg.rotation_epsilon = e; return RG_OK; }
int rg_set_transformation_epsilon(rg_ctx* c, double e) { CHECK_CTX(c); c->lsq.cfg.transformation_epsilon = e; return RG_OK; }
int rg_set_maximum_iterations(rg_ctx* c, int n) { CHECK_CTX(c); c->lsq.cfg.max_iterations = n; return RG_OK; }
int rg_set_initial_lambda_factor(rg_ctx* c, double f) { CHECK_CTX(c); c->lsq.cfg.lm_init_lambda_factor = f; return RG_OK; }
int rg_set_debug_print(rg_ctx* c, int on) { CHECK_CTX(c); c->lsq.cfg.debug_print = on != 0; return RG_OK; }
int rg_set_ct_lambda(rg_ctx* c, float l) { CHECK_CTX(c); c->lambda_ = l; return RG_OK; }
int rg_set_profile(rg_ctx* c, int on) { CHECK_CTX(c); c->profile = on != 0; return RG_OK; }
double rg_last_kernel_us(const rg_ctx* c) { return c ? c->last_kernel_us : -1.0; }

int rg_set_input_source(rg_ctx* c, const float* pts, size_t n, size_t stride) {
  CHECK_CTX(c);
  int rc = upload_cloud(c, c->src, pts, n, stride);
  c->st_ok = false; c->have_corr = false; c->shard_lo = c->shard_hi = 0;
  return rc;
}
int rg_set_input_target(rg_ctx* c, const float* pts, size_t n, size_t stride) {
  CHECK_CTX(c);
  int rc = upload_cloud(c, c->tgt, pts, n, stride);
  c->voxel_ok = false; c->st_ok = false; c->have_corr = false;
  return rc;
}
int rg_set_input_source_device(rg_ctx* c, const float* dev, size_t n, void* stream) {
  CHECK_CTX(c);
  int rc = upload_cloud_device(c, c->src, dev, n, (cudaStream_t)stream);
  c->st_ok = false; c->have_corr = false; c->shard_lo = c->shard_hi = 0;
  return rc;
}
int rg_set_input_target_device(rg_ctx* c, const float* dev, size_t n, void* stream) {
  CHECK_CTX(c);
  int rc = upload_cloud_device(c, c->tgt, dev, n, (cudaStream_t)stream);
  c->voxel_ok = false; c->st_ok = false; c->have_corr = false;
  return rc;
}
int rg_clear_source(rg_ctx* c) { CHECK_CTX(c); c->src.clear(); c->st_ok = false; c->have_corr = false; return RG_OK; }
int rg_clear_target(rg_ctx* c) { CHECK_CTX(c); c->tgt.clear(); c->voxel_ok = false; c->st_ok = false; c->have_corr = false; return RG_OK; }
int rg_swap_source_and_target(rg_ctx* c) {
  CHECK_CTX(c);
  std::swap(c->src, c->tgt);
  c->voxel_ok = false; c->st_ok = false; c->have_corr = false; c->num_corr = 0;
  return RG_OK;
}
size_t rg_source_size(const rg_ctx* c) { return c ? (size_t)c->src.n : 0; }
size_t rg_target_size(const rg_ctx* c) { return c ? (size_t)c->tgt.n : 0; }

int rg_set_covariances(rg_ctx* c, int which, const double* cov, size_t n) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  Cloud& cl = which == RG_SOURCE ? c->src : c->tgt;
  if (n != (size_t)cl.n) {
    // the reference stores any vector and recomputes at align() when the size mismatches (rot_vgicp_impl.hpp:153-158)
    cl.has_cov = false;
    return RG_OK;
  }
  if (n == 0) return RG_OK;
  if (!cov) return set_err(c, RG_ERR_INVALID, "null covariance pointer");
  RG_CUDA(c, c->staging.ensure(c->device, n * 16 * sizeof(double)));
  RG_CUDA(c, cudaMemcpyAsync(c->staging.p, cov, n * 16 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  cov_from_double_kernel<<<(cl.n_pad + 255) / 256, 256, 0, c->stream>>>((const double*)c->staging.p, cl.n, cl.n_pad, cl.cov6.p);
  RG_CUDA(c, cudaGetLastError());
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  cl.has_cov = true;
  if (which == RG_TARGET) { c->voxel_ok = false; c->st_ok = false; }
  return RG_OK;
}

size_t rg_get_covariances(rg_ctx* c, int which, double* out) {
  if (!c) return 0;
  Cloud& cl = which == RG_SOURCE ? c->src : c->tgt;
  if (!cl.has_cov) return 0;
  if (!out) return (size_t)cl.n;
  if (use_device(c)) return 0;
  if (c->staging.ensure(c->device, (size_t)cl.n * 16 * sizeof(double)) != cudaSuccess) return 0;
  cov_to_double_kernel<<<(cl.n + 255) / 256, 256, 0, c->stream>>>(cl.cov6.p, cl.n, (double*)c->staging.p);
  if (cudaMemcpyAsync(out, c->staging.p, (size_t)cl.n * 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return 0;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return 0;
  return (size_t)cl.n;
}

int rg_calculate_covariances(rg_ctx* c, int which) {
  CHECK_CTX(c);
  // inside a batch the estimation runs on the pair's own setup stream (the one its uploads use), so that the uploads and
  // kNN kernels of different pairs overlap; the batch stream is ordered after it
  cudaStream_t saved = c->stream;
  const bool side = c->batch && c->upload_stream && c->upload_stream != c->stream;
  if (side) { cudaEventRecord(c->ev_upload, c->stream); cudaStreamWaitEvent(c->upload_stream, c->ev_upload, 0); c->stream = c->upload_stream; }
  int rc = calc_cov(c, which == RG_SOURCE ? c->src : c->tgt);
  if (side) { cudaEventRecord(c->ev_upload, c->stream); c->stream = saved; c->side_pending = true; }   // joined lazily (use_device) or by rg_batch_align
  if (rc == RG_OK && which == RG_TARGET) { c->voxel_ok = false; c->st_ok = false; }
  return rc;
}

int rg_create_voxelmap(rg_ctx* c) {
  CHECK_CTX(c);
  if (!c->tgt.has_cov) { if (int rc = calc_cov(c, c->tgt)) return rc; }
  if (int rc = build_voxelmap(c)) return rc;
  return fetch_voxel_count(c) < 0 ? RG_ERR_INVALID : RG_OK;
}

static int eval_pose(rg_ctx* c, int kind, const double T[16], double* H, double* b, double* cost, int ndof) {
  CHECK_CTX(c);
  if (!T) return set_err(c, RG_ERR_INVALID, "null pose");
  host::Request r; r.kind = kind; std::memcpy(r.pose, T, sizeof(r.pose));
  double res[RES_N];
  if (kind != host::REQ_ERROR) { if (int rc = ensure_covariances(c)) return rc; }
  if (int rc = evaluate_single(c, r, nullptr, nullptr, 0, 0, res)) return rc;
  unpack_result(res, ndof, H, b, cost);
  return RG_OK;
}
int rg_linearize(rg_ctx* c, const double T[16], double* H, double* b, double* cost) { return eval_pose(c, host::REQ_LIN6, T, H, b, cost, 6); }
int rg_so3_linearize(rg_ctx* c, const double T[16], double* H, double* b, double* cost) { return eval_pose(c, host::REQ_LIN3, T, H, b, cost, 3); }
int rg_compute_error(rg_ctx* c, const double T[16], double* cost) { return eval_pose(c, host::REQ_ERROR, T, nullptr, nullptr, cost, 6); }

int rg_t3_linearize(rg_ctx* c, const double t[3], const double g[3], const double l[3], double dtn, double dtn1, double* H, double* b, double* cost) {
  CHECK_CTX(c);
  host::Request r; r.kind = host::REQ_T3_LIN; std::memcpy(r.t, t, sizeof(r.t));
  double res[RES_N];
  if (int rc = evaluate_single(c, r, g, l, dtn, dtn1, res)) return rc;
  unpack_result(res, 6, H, b, cost);
  return RG_OK;
}
int rg_compute_t_error(rg_ctx* c, const double t[3], const double g[3], const double l[3], double dtn, double dtn1, double* cost) {
  CHECK_CTX(c);
  host::Request r; r.kind = host::REQ_T_ERROR; std::memcpy(r.t, t, sizeof(r.t));
  double res[RES_N];
  if (int rc = evaluate_single(c, r, g, l, dtn, dtn1, res)) return rc;
  if (cost) *cost = res[27];
  return RG_OK;
}
size_t rg_num_correspondences(const rg_ctx* c) { return c ? c->num_corr : 0; }

int rg_align(rg_ctx* c, const float guess[16], float out_T[16]) {
  CHECK_CTX(c);
  if (!out_T) return set_err(c, RG_ERR_INVALID, "null output");
  if (c->src.n == 0 || c->tgt.n == 0) return set_err(c, RG_ERR_INVALID, "align: source/target cloud not set");
  c->voxel_ok = false; c->st_ok = false; c->have_corr = false;      // voxelmap_.reset() (rot_vgicp_impl.hpp:147)
  if (int rc = check_grid(c)) return rc;
  if (int rc = ensure_covariances(c)) return rc;
  double g[16];
  if (guess) for (int i = 0; i < 16; i++) g[i] = (double)guess[i]; else host::mat4_identity(g);
  c->lsq.begin_align(g);
  if (int rc = run_driver(c, nullptr, nullptr, 0, 0)) return rc;
  if (fetch_voxel_count(c) < 0) return RG_ERR_INVALID;               // target voxel keys outside the +-2^20 cell range are an error, not dropped
  for (int i = 0; i < 16; i++) out_T[i] = (float)c->lsq.x0[i];      // final_transformation_ = x0.cast<float>() (lsq.hpp:177)
  return c->lsq.failed ? RG_ERR_NOT_CONVERGED : RG_OK;
}

int rg_compute_translation(rg_ctx* c, double trans[3], const double g[3], const double l[3], double dtn, double dtn1, float ct_lambda) {
  CHECK_CTX(c);
  if (!trans || !g || !l) return set_err(c, RG_ERR_INVALID, "null argument");
  c->lambda_ = ct_lambda;                                           // rot_vgicp_impl.hpp:166
  c->lsq.begin_translation(trans);
  if (int rc = run_driver(c, g, l, dtn, dtn1)) return rc;
  trans[0] = c->lsq.t0[0]; trans[1] = c->lsq.t0[1]; trans[2] = c->lsq.t0[2];
  return c->lsq.failed ? RG_ERR_NOT_CONVERGED : RG_OK;
}

int rg_transform_source(rg_ctx* c, const float T[16], float* out, size_t stride) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  const int n = c->src.n;
  if (n == 0) return RG_OK;
  if (!out || !T || stride < 12 || stride % 4) return set_err(c, RG_ERR_INVALID, "bad arguments");
  RG_CUDA(c, c->staging.ensure(c->device, (size_t)n * sizeof(float4) + 64));
  float* dT = (float*)c->staging.p;
  float4* dout = (float4*)((char*)c->staging.p + 64);
  RG_CUDA(c, cudaMemcpyAsync(dT, T, 16 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  transform_points_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->src.pts.p, n, dT, dout);
  RG_CUDA(c, cudaGetLastError());
  std::vector<float4> tmp((size_t)n);
  RG_CUDA(c, cudaMemcpyAsync(tmp.data(), dout, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; i++) {
    float* o = (float*)((char*)out + (size_t)i * stride);
    o[0] = tmp[i].x; o[1] = tmp[i].y; o[2] = tmp[i].z;
  }
  return RG_OK;
}

int rg_has_converged(const rg_ctx* c) { return c && c->lsq.converged ? 1 : 0; }
int rg_nr_iterations(const rg_ctx* c) { return c ? c->lsq.nr_iterations : 0; }
int rg_get_final_hessian(const rg_ctx* c, double H[36]) { CHECK_CTX(c); std::memcpy(H, c->lsq.final_hessian, sizeof(double) * 36); return RG_OK; }
int rg_get_final_pose(const rg_ctx* c, double T[16]) { CHECK_CTX(c); std::memcpy(T, c->lsq.x0, sizeof(double) * 16); return RG_OK; }

size_t rg_voxel_count(const rg_ctx* c) { int v = fetch_voxel_count(const_cast<rg_ctx*>(c)); return v < 0 ? 0 : (size_t)v; }

int rg_get_voxels(rg_ctx* c, int32_t* keys, double* mean, double* cov6, int32_t* npts) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  const int V = fetch_voxel_count(c);
  if (V < 0) return RG_ERR_INVALID;
  if (V == 0) return RG_OK;
  DevBuf<unsigned long long> dk; DevBuf<VoxPayload> dp;
  RG_CUDA(c, dk.ensure(c->device, V)); RG_CUDA(c, dp.ensure(c->device, V));
  const unsigned cap = c->mask + 1;
  table_dump_kernel<<<(cap + 255) / 256, 256, 0, c->stream>>>(c->table.p, cap, c->vox.p, dk.p, dp.p);
  std::vector<unsigned long long> hk(V); std::vector<VoxPayload> hp(V);
  RG_CUDA(c, cudaMemcpyAsync(hk.data(), dk.p, sizeof(unsigned long long) * V, cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaMemcpyAsync(hp.data(), dp.p, sizeof(VoxPayload) * V, cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  dk.release(); dp.release();
  std::vector<int> order(V);
  std::vector<int> kx(V), ky(V), kz(V);
  for (int i = 0; i < V; i++) { order[i] = i; unpack_key(hk[i], kx[i], ky[i], kz[i]); }
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    if (kx[a] != kx[b]) return kx[a] < kx[b];
    if (ky[a] != ky[b]) return ky[a] < ky[b];
    return kz[a] < kz[b]; });
  for (int o = 0; o < V; o++) {
    const int i = order[o];
    if (keys) { keys[o * 3 + 0] = kx[i]; keys[o * 3 + 1] = ky[i]; keys[o * 3 + 2] = kz[i]; }
    if (mean) { mean[o * 3 + 0] = hp[i].mean[0]; mean[o * 3 + 1] = hp[i].mean[1]; mean[o * 3 + 2] = hp[i].mean[2]; }
    if (cov6) for (int t = 0; t < 6; t++) cov6[o * 6 + t] = hp[i].cov[t];
    if (npts) npts[o] = (int32_t)(hp[i].w * hp[i].w + 0.5);
  }
  return RG_OK;
}

int rg_get_correspondence_keys(rg_ctx* c, int32_t* keys) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!c->have_corr) return set_err(c, RG_ERR_INVALID, "no correspondences yet");
  const int V = fetch_voxel_count(c);
  if (V < 0) return RG_ERR_INVALID;
  const size_t m = (size_t)((c->shard_hi > c->shard_lo) ? (c->shard_hi - c->shard_lo) : c->src.n) * c->corr_noff;
  std::vector<int> hc(m);
  RG_CUDA(c, cudaMemcpyAsync(hc.data(), c->corr.p, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
  std::vector<VoxSlot> ht((size_t)c->mask + 1);
  RG_CUDA(c, cudaMemcpyAsync(ht.data(), c->table.p, sizeof(VoxSlot) * ht.size(), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<unsigned long long> id2key(std::max(V, 1));
  for (auto& s : ht) if (s.key != KEY_EMPTY && s.id >= 0 && s.id < V) id2key[s.id] = s.key;
  for (size_t i = 0; i < m; i++) {
    if (hc[i] < 0) { keys[i * 3] = keys[i * 3 + 1] = keys[i * 3 + 2] = INT32_MIN; continue; }
    int x, y, z; unpack_key(id2key[hc[i]], x, y, z);
    keys[i * 3] = x; keys[i * 3 + 1] = y; keys[i * 3 + 2] = z;
  }
  return RG_OK;
}

int rg_knn(rg_ctx* c, int which, int k, int32_t* out) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  Cloud& cl = which == RG_SOURCE ? c->src : c->tgt;
  const int n = cl.n;
  if (n == 0 || k < 1 || k > 64 || n < k || !out) return set_err(c, RG_ERR_INVALID, "rg_knn: bad arguments");
  RG_CUDA(c, c->codes_in.ensure(c->device, n)); RG_CUDA(c, c->codes_out.ensure(c->device, n));
  RG_CUDA(c, c->idx_in.ensure(c->device, n)); RG_CUDA(c, c->idx_out.ensure(c->device, n));
  RG_CUDA(c, c->sorted.ensure(c->device, n)); RG_CUDA(c, c->bbox.ensure(c->device, 8)); RG_CUDA(c, c->bucket_start.ensure(c->device, cov_bucket_table_entries()));
  const size_t tb = cov_cub_temp_bytes(n);
  RG_CUDA(c, c->cub_temp.ensure(c->device, tb));
  DevBuf<int> dk;
  RG_CUDA(c, dk.ensure(c->device, (size_t)n * k));
  CovLaunch L;
  L.pts = cl.pts.p; L.n = n; L.k = k; L.method = c->regularization_method; L.cov6 = nullptr; L.knn = dk.p;
  L.s.codes_in = c->codes_in.p; L.s.codes_out = c->codes_out.p; L.s.idx_in = c->idx_in.p; L.s.idx_out = c->idx_out.p;
  L.s.sorted = c->sorted.p; L.s.cub_temp = c->cub_temp.p; L.s.cub_temp_bytes = tb; L.s.bbox = c->bbox.p; L.s.bucket_start = c->bucket_start.p;
  RG_CUDA(c, launch_covariances(L, c->stream));
  RG_CUDA(c, cudaMemcpyAsync(out, dk.p, sizeof(int) * (size_t)n * k, cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  dk.release();
  return RG_OK;
}

int rg_set_source_shard(rg_ctx* c, size_t lo, size_t hi) {
  CHECK_CTX(c);
  if (lo == 0 && hi == 0) { c->shard_lo = c->shard_hi = 0; c->st_ok = false; c->have_corr = false; return RG_OK; }
  if (lo % TILE_PTS != 0 || hi <= lo || hi > (size_t)c->src.n) return set_err(c, RG_ERR_INVALID, "rg_set_source_shard: need lo % 256 == 0 and lo < hi <= source size");
  c->shard_lo = (int)lo; c->shard_hi = (int)hi; c->st_ok = false; c->have_corr = false;
  return RG_OK;
}

int rg_shard_export(rg_ctx* c, void* handle) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!handle) return set_err(c, RG_ERR_INVALID, "rg_shard_export: null handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!c->xchg) {
    RG_CUDA(c, cudaMalloc((void**)&c->xchg, XCHG_DOUBLES * sizeof(double)));
    RG_CUDA(c, cudaMemset(c->xchg, 0, XCHG_DOUBLES * sizeof(double)));
  }
  cudaIpcMemHandle_t h;
  RG_CUDA(c, cudaIpcGetMemHandle(&h, c->xchg));
  std::memcpy(handle, &h, sizeof(h));
  return RG_OK;
}

int rg_shard_connect(rg_ctx* c, int rank, int world, const void* handles) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (world < 1 || world > XCHG_MAX_RANKS || rank < 0 || rank >= world || !handles || !c->xchg)
    return set_err(c, RG_ERR_INVALID, "rg_shard_connect: bad arguments (call rg_shard_export first; world <= 8)");
  for (int p = 0; p < world; p++) {
    if (p == rank) { c->peer_xchg[p] = c->xchg; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + (size_t)p * 64, sizeof(h));
    void* ptr = nullptr;
    RG_CUDA(c, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_xchg[p] = (double*)ptr;
  }
  RG_CUDA(c, cudaMemset(c->xchg, 0, XCHG_DOUBLES * sizeof(double)));
  RG_CUDA(c, cudaDeviceSynchronize());
  c->p2p_rank = rank; c->p2p_world = world; c->shard_seq = 0; c->st_ok = false;
  c->reduce_fn = nullptr;
  return RG_OK;
}

int rg_set_reduce_hook(rg_ctx* c, rg_reduce_fn fn, void* user, int on_device) {
  CHECK_CTX(c);
  c->reduce_fn = fn; c->reduce_user = user; c->reduce_on_device = on_device; c->st_ok = false;
  return RG_OK;
}

// ------------------------------------------------------------------------------------------------
// batch of independent scan pairs
// ------------------------------------------------------------------------------------------------
int rg_batch_create(int device, int n_pairs, rg_batch** out) {
  if (!out || n_pairs < 1) return set_err(nullptr, RG_ERR_INVALID, "rg_batch_create: bad arguments");
  *out = nullptr;
  rg_ctx* first = nullptr;
  if (int rc = create_ctx(device, nullptr, &first)) return rc;
  rg_batch* b = new rg_batch;
  b->device = first->device;
  b->stream = first->stream;          // the batch owns the stream through pair 0
  first->batch = b; first->batch_index = 0;
  b->pairs.push_back(first);
  for (int i = 1; i < n_pairs; i++) {
    rg_ctx* c = nullptr;
    if (int rc = create_ctx(b->device, b->stream, &c)) { rg_batch_destroy(b); return rc; }
    c->batch = b; c->batch_index = i;
    b->pairs.push_back(c);
  }
  void* p = nullptr;
  bool ok = b->st_dev.ensure(b->device, n_pairs) == cudaSuccess && b->dyn_dev.ensure(b->device, n_pairs) == cudaSuccess &&
            b->st_sub.ensure(b->device, n_pairs) == cudaSuccess && b->dyn_sub.ensure(b->device, n_pairs) == cudaSuccess;
  if (ok && pool().pinned_alloc(sizeof(LinDyn) * n_pairs, &p) == cudaSuccess) b->dyn_host = (LinDyn*)p; else ok = false;
  if (ok && pool().pinned_alloc(sizeof(double) * RES_N * n_pairs, &p) == cudaSuccess) b->result_host = (double*)p; else ok = false;
  if (ok && pool().pinned_alloc(sizeof(unsigned) * n_pairs + 64, &p) == cudaSuccess) b->flag_host = (unsigned*)p; else ok = false;
  if (!ok) { rg_batch_destroy(b); return set_err(nullptr, RG_ERR_CUDA, "rg_batch_create: allocation failed"); }
  std::memset(b->flag_host, 0, sizeof(unsigned) * n_pairs);
  cudaEventCreate(&b->ev0); cudaEventCreate(&b->ev1);
  cudaEventCreateWithFlags(&b->ev_inputs, cudaEventDisableTiming);
  const int n_setup = std::min(n_pairs, 4);
  for (int i = 0; i < n_setup; i++) {
    cudaStream_t st = nullptr; cudaEvent_t ev = nullptr;
    if (pool().stream_get(b->device, &st) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
      rg_batch_destroy(b); return set_err(nullptr, RG_ERR_CUDA, "rg_batch_create: stream/event creation failed");
    }
    b->setup_streams.push_back(st); b->setup_events.push_back(ev);
  }
  for (int i = 0; i < n_pairs; i++) b->pairs[i]->upload_stream = b->setup_streams[i % n_setup];
  *out = b;
  return RG_OK;
}

void rg_batch_destroy(rg_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (size_t i = b->pairs.size(); i-- > 0;) rg_destroy(b->pairs[i]);   // pair 0 (stream owner) last
  b->st_dev.release(); b->dyn_dev.release(); b->st_sub.release(); b->dyn_sub.release(); b->sched_dev.release();
  if (b->dyn_host) pool().pinned_free(b->dyn_host);
  if (b->result_host) pool().pinned_free(b->result_host);
  if (b->flag_host) pool().pinned_free(b->flag_host);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->ev_inputs) cudaEventDestroy(b->ev_inputs);
  for (auto ev : b->setup_events) cudaEventDestroy(ev);
  for (auto st : b->setup_streams) { cudaStreamSynchronize(st); pool().stream_put(b->device, st); }
  delete b;
}

rg_ctx* rg_batch_ctx(rg_batch* b, int pair) { return (b && pair >= 0 && pair < (int)b->pairs.size()) ? b->pairs[pair] : nullptr; }
double rg_batch_last_linearize_us(const rg_batch* b) { return (b && b->lin_launches) ? b->lin_us_sum / b->lin_launches : -1.0; }
int rg_batch_last_linearize_launches(const rg_batch* b) { return b ? b->lin_launches : 0; }
void* rg_batch_stream(const rg_batch* b) { return b ? (void*)b->stream : nullptr; }
void* rg_stream(const rg_ctx* c) { return c ? (void*)c->stream : nullptr; }
unsigned long long rg_kernel_launch_count(void) { return g_launches.load(); }

int rg_batch_align(rg_batch* b, const float* guesses, float* out_T) {
  if (!b || !out_T) return set_err(nullptr, RG_ERR_INVALID, "rg_batch_align: bad arguments");
  const int P = (int)b->pairs.size();
  if (cudaSetDevice(b->device) != cudaSuccess) return set_err(nullptr, RG_ERR_CUDA, "cudaSetDevice failed");
  // stage 1+2: covariances and voxel maps of every pair. Pairs are spread over a few setup streams so that the
  // small, latency-bound kernels of one pair (radix-sort passes, bbox, Morton codes, voxel insert) overlap the
  // big kNN kernels of the others; the batch stream joins them before the lock-step iterations start.
  int noff = -1;
  const int S = (int)b->setup_streams.size();
  RG_CUDA(nullptr, cudaEventRecord(b->ev_inputs, b->stream));
  for (int si = 0; si < S; si++) RG_CUDA(nullptr, cudaStreamWaitEvent(b->setup_streams[si], b->ev_inputs, 0));
  int rc_setup = RG_OK;
  for (int i = 0; i < P && rc_setup == RG_OK; i++) {
    rg_ctx* c = b->pairs[i];
    if (c->src.n == 0 || c->tgt.n == 0) { rc_setup = set_err(c, RG_ERR_INVALID, "batch align: source/target cloud not set"); break; }
    c->voxel_ok = false; c->st_ok = false; c->have_corr = false;
    c->stream = b->setup_streams[i % S];
    if ((rc_setup = check_grid(c)) == RG_OK && (rc_setup = ensure_covariances(c)) == RG_OK) rc_setup = build_voxelmap(c);
    c->stream = b->stream;
  }
  for (int si = 0; si < S; si++) {
    cudaEventRecord(b->setup_events[si], b->setup_streams[si]);
    cudaStreamWaitEvent(b->stream, b->setup_events[si], 0);
  }
  if (rc_setup != RG_OK) return rc_setup;
  for (int i = 0; i < P; i++) {
    rg_ctx* c = b->pairs[i];
    if (int rc = prepare_linearize(c)) return rc;
    if (noff < 0) noff = c->corr_noff;
    if (c->corr_noff != noff) return set_err(c, RG_ERR_INVALID, "batch align: all pairs must use the same neighbor search method");
    double g[16];
    if (guesses) for (int k = 0; k < 16; k++) g[k] = (double)guesses[i * 16 + k]; else host::mat4_identity(g);
    c->lsq.begin_align(g);
  }
  b->lin_us_sum = 0; b->lin_launches = 0;
  b->sched_group.clear();   // cloud sizes may have changed since the last call
  std::vector<host::Request> req(P);
  std::vector<int> group;
  std::vector<LinStatic> st_host(P);
  RG_CUDA(nullptr, cudaMemcpyAsync(st_host.data(), b->st_dev.p, sizeof(LinStatic) * P, cudaMemcpyDeviceToHost, b->stream));
  RG_CUDA(nullptr, cudaStreamSynchronize(b->stream));
  for (;;) {
    // collect the pending request of every unfinished pair and serve them grouped by kernel variant
    std::vector<int> pend;
    for (int i = 0; i < P; i++) if (b->pairs[i]->lsq.want(req[i])) pend.push_back(i);
    if (pend.empty()) break;
    for (int variant = 0; variant < 3; variant++) {
      const int kind = variant == 0 ? host::REQ_LIN6 : (variant == 1 ? host::REQ_LIN3 : host::REQ_ERROR);
      group.clear();
      for (int i : pend) if (req[i].kind == kind) group.push_back(i);
      if (group.empty()) continue;
      const int G = (int)group.size();
      int mode = 0, refresh = 0;
      const unsigned seq = ++g_seq;
      for (int gi = 0; gi < G; gi++) {
        rg_ctx* c = b->pairs[group[gi]];
        if (int rc = make_dyn(c, req[group[gi]], nullptr, nullptr, 0, 0, b->dyn_host[gi], mode, refresh)) return rc;
        b->dyn_host[gi].seq = seq;
      }
      const LinStatic* st_arr = b->st_dev.p;
      if (G != P) {   // compact the static descriptors of the active subset
        std::vector<LinStatic> sub(G);
        for (int gi = 0; gi < G; gi++) sub[gi] = st_host[group[gi]];
        RG_CUDA(nullptr, cudaMemcpyAsync(b->st_sub.p, sub.data(), sizeof(LinStatic) * G, cudaMemcpyHostToDevice, b->stream));
        RG_CUDA(nullptr, cudaStreamSynchronize(b->stream));
        st_arr = b->st_sub.p;
      }
      RG_CUDA(nullptr, cudaMemcpyAsync(b->dyn_dev.p, b->dyn_host, sizeof(LinDyn) * G, cudaMemcpyHostToDevice, b->stream));
      if (b->sched_group != group) {   // (re)build the persistent-CTA schedule for this set of pairs
        std::vector<int> npts(G), sched;
        long long tiles = 0;
        for (int gi = 0; gi < G; gi++) { npts[gi] = st_host[group[gi]].n; tiles += (npts[gi] + TILE_PTS - 1) / TILE_PTS; }
        b->sched_grid = (int)std::max<long long>(1, std::min<long long>(tiles, lin_slots(b->device)));
        build_schedule(npts.data(), G, b->sched_grid, sched, b->sched_tiles);
        RG_CUDA(nullptr, b->sched_dev.ensure(b->device, (size_t)3 * P + 1));
        RG_CUDA(nullptr, cudaMemcpyAsync(b->sched_dev.p, sched.data(), sched.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
        RG_CUDA(nullptr, cudaStreamSynchronize(b->stream));
        b->sched_group = group;
      }
      LinLaunch L;
      L.mode = mode; L.refresh = refresh; L.noff = noff; L.ct = 0; L.n_pairs = G; L.grid = b->sched_grid; L.total_tiles = b->sched_tiles;
      L.st = st_arr; L.dyn = b->dyn_dev.p; L.sched = b->sched_dev.p;
      cudaEventRecord(b->ev0, b->stream);
      RG_CUDA(nullptr, launch_linearize(L, nullptr, b->stream));
      g_launches++;
      cudaEventRecord(b->ev1, b->stream);
      for (int gi = 0; gi < G; gi++) {
        rg_ctx* c = b->pairs[group[gi]];
        if (int rc = wait_flag(c, b->flag_host + group[gi], seq, b->stream)) return rc;
      }
      if (cudaEventSynchronize(b->ev1) == cudaSuccess) { float ms = 0; cudaEventElapsedTime(&ms, b->ev0, b->ev1); b->lin_us_sum += ms * 1e3; b->lin_launches++; }
      for (int gi = 0; gi < G; gi++) {
        const int i = group[gi];
        rg_ctx* c = b->pairs[i];
        const double* res = b->result_host + (size_t)i * RES_N;
        double H[36], bb[6], cost;
        if (kind == host::REQ_ERROR) c->lsq.feed(nullptr, nullptr, res[27]);
        else {
          c->have_corr = true; c->num_corr = (size_t)(res[28] + 0.5);
          unpack_result(res, kind == host::REQ_LIN3 ? 3 : 6, H, bb, &cost);
          c->lsq.feed(H, bb, cost);
        }
      }
    }
  }
  int rc_all = RG_OK;
  {   // voxel-key range / table overflow flags of all pairs: P small async copies, one sync
    std::vector<int> flags((size_t)P * 2, 0);
    for (int i = 0; i < P; i++) RG_CUDA(nullptr, cudaMemcpyAsync(&flags[(size_t)i * 2], b->pairs[i]->counters.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    RG_CUDA(nullptr, cudaStreamSynchronize(b->stream));
    for (int i = 0; i < P; i++) {
      b->pairs[i]->n_vox = flags[(size_t)i * 2];
      if (flags[(size_t)i * 2 + 1] != 0) return set_err(b->pairs[i], RG_ERR_INVALID, "voxel coordinate out of the +-2^20 key range (or table overflow)");
    }
  }
  for (int i = 0; i < P; i++) {
    rg_ctx* c = b->pairs[i];
    for (int k = 0; k < 16; k++) out_T[i * 16 + k] = (float)c->lsq.x0[k];
    if (c->lsq.failed) rc_all = RG_ERR_NOT_CONVERGED;
  }
  return rc_all;
}

// ------------------------------------------------------------------------------------------------
// upstream range-image passes
// ------------------------------------------------------------------------------------------------
static int fe_alloc(rg_ctx* c, int n_scan, int horizon) {
  const size_t cells = (size_t)n_scan * horizon;
  RG_CUDA(c, c->fe_owner.ensure(c->device, cells)); RG_CUDA(c, c->fe_range_image.ensure(c->device, cells));
  RG_CUDA(c, c->fe_full.ensure(c->device, cells)); RG_CUDA(c, c->fe_row_count.ensure(c->device, n_scan));
  RG_CUDA(c, c->fe_extracted.ensure(c->device, cells)); RG_CUDA(c, c->fe_col.ensure(c->device, cells));
  RG_CUDA(c, c->fe_prange.ensure(c->device, cells)); RG_CUDA(c, c->fe_start.ensure(c->device, n_scan)); RG_CUDA(c, c->fe_end.ensure(c->device, n_scan));
  RG_CUDA(c, c->fe_counts.ensure(c->device, 8));
  RG_CUDA(c, c->fe_curv.ensure(c->device, cells)); RG_CUDA(c, c->fe_picked.ensure(c->device, cells)); RG_CUDA(c, c->fe_label.ensure(c->device, cells));
  RG_CUDA(c, c->fe_corner_ring.ensure(c->device, (size_t)n_scan * FE_MAX_CORNERS_PER_RING)); RG_CUDA(c, c->fe_corner_cnt.ensure(c->device, n_scan));
  RG_CUDA(c, c->fe_surf_ring.ensure(c->device, cells)); RG_CUDA(c, c->fe_surf_cnt.ensure(c->device, n_scan));
  RG_CUDA(c, c->fe_corner_out.ensure(c->device, (size_t)n_scan * FE_MAX_CORNERS_PER_RING)); RG_CUDA(c, c->fe_surf_out.ensure(c->device, cells));
  return RG_OK;
}

int rg_image_projection(rg_ctx* c, const void* raw, size_t n_points, const rg_projection_params* p, float* range_image, float* extracted,
                        int32_t* col_ind, float* point_range, int32_t* start_ring, int32_t* end_ring, int32_t* n_extracted) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!p || (!raw && n_points) || !n_extracted) return set_err(c, RG_ERR_INVALID, "rg_image_projection: null argument");
  if (p->n_scan < 1 || p->horizon_scan < 1 || p->downsample_rate < 1 || p->point_stride_bytes < 12 || p->point_stride_bytes % 4 ||
      p->ring_offset_bytes < 0 || p->ring_offset_bytes + 2 > p->point_stride_bytes || n_points > (size_t)INT32_MAX / 2)
    return set_err(c, RG_ERR_INVALID, "rg_image_projection: bad parameters");
  if (int rc = fe_alloc(c, p->n_scan, p->horizon_scan)) return rc;
  const size_t cells = (size_t)p->n_scan * p->horizon_scan;
  RG_CUDA(c, c->fe_raw.ensure(c->device, std::max<size_t>(n_points * p->point_stride_bytes, 16)));
  if (n_points) RG_CUDA(c, cudaMemcpyAsync(c->fe_raw.p, raw, n_points * p->point_stride_bytes, cudaMemcpyHostToDevice, c->stream));
  ProjLaunch L;
  L.raw = c->fe_raw.p; L.n_points = (int)n_points; L.stride = p->point_stride_bytes; L.ring_offset = p->ring_offset_bytes;
  L.n_scan = p->n_scan; L.horizon = p->horizon_scan; L.downsample = p->downsample_rate; L.min_range = p->lidar_min_range; L.max_range = p->lidar_max_range;
  L.owner = c->fe_owner.p; L.range_image = c->fe_range_image.p; L.full = c->fe_full.p; L.row_count = c->fe_row_count.p;
  L.extracted = c->fe_extracted.p; L.col_ind = c->fe_col.p; L.point_range = c->fe_prange.p; L.start_ring = c->fe_start.p; L.end_ring = c->fe_end.p;
  L.n_extracted = c->fe_counts.p;
  RG_CUDA(c, launch_projection(L, c->stream));
  g_launches += 5;
  int n = 0;
  RG_CUDA(c, cudaMemcpyAsync(&n, c->fe_counts.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_extracted = n;
  if (range_image) RG_CUDA(c, cudaMemcpyAsync(range_image, c->fe_range_image.p, cells * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (start_ring) RG_CUDA(c, cudaMemcpyAsync(start_ring, c->fe_start.p, p->n_scan * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (end_ring) RG_CUDA(c, cudaMemcpyAsync(end_ring, c->fe_end.p, p->n_scan * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (n > 0) {
    if (extracted) RG_CUDA(c, cudaMemcpyAsync(extracted, c->fe_extracted.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    if (col_ind) RG_CUDA(c, cudaMemcpyAsync(col_ind, c->fe_col.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (point_range) RG_CUDA(c, cudaMemcpyAsync(point_range, c->fe_prange.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  }
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  return RG_OK;
}

int rg_feature_extraction(rg_ctx* c, const float* extracted, const int32_t* col_ind, const float* point_range, const int32_t* start_ring,
                          const int32_t* end_ring, int32_t n, const rg_feature_params* p, float* curvature, int32_t* picked, int32_t* label,
                          float* corner, int32_t* n_corner, float* surface, int32_t* n_surface) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!p || !extracted || !col_ind || !point_range || !start_ring || !end_ring || !n_corner || !n_surface)
    return set_err(c, RG_ERR_INVALID, "rg_feature_extraction: null argument");
  if (p->n_scan < 1 || p->horizon_scan < 1 || n < 1 || (size_t)n > (size_t)p->n_scan * p->horizon_scan || !(p->surf_leaf_size > 0.f))
    return set_err(c, RG_ERR_INVALID, "rg_feature_extraction: bad parameters");
  if (p->horizon_scan > FE_RING_CAP) return set_err(c, RG_ERR_UNSUPPORTED, "rg_feature_extraction: horizon_scan above the per-ring capacity (2048)");
  if (int rc = fe_alloc(c, p->n_scan, p->horizon_scan)) return rc;
  RG_CUDA(c, cudaMemcpyAsync(c->fe_extracted.p, extracted, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, cudaMemcpyAsync(c->fe_col.p, col_ind, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, cudaMemcpyAsync(c->fe_prange.p, point_range, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, cudaMemcpyAsync(c->fe_start.p, start_ring, p->n_scan * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, cudaMemcpyAsync(c->fe_end.p, end_ring, p->n_scan * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, cudaMemsetAsync(c->fe_counts.p, 0, 8 * sizeof(int), c->stream));
  FeatLaunch L;
  L.extracted = c->fe_extracted.p; L.col_ind = c->fe_col.p; L.point_range = c->fe_prange.p; L.start_ring = c->fe_start.p; L.end_ring = c->fe_end.p;
  L.n_dev = nullptr; L.n = n; L.n_scan = p->n_scan; L.horizon = p->horizon_scan;
  L.edge_thr = p->edge_threshold; L.surf_thr = p->surf_threshold; L.leaf = p->surf_leaf_size;
  L.curvature = c->fe_curv.p; L.picked = c->fe_picked.p; L.label = c->fe_label.p;
  L.corner_ring = c->fe_corner_ring.p; L.corner_cnt = c->fe_corner_cnt.p; L.surf_ring = c->fe_surf_ring.p; L.surf_cnt = c->fe_surf_cnt.p;
  L.corner_out = c->fe_corner_out.p; L.n_corner = c->fe_counts.p + 1; L.surf_out = c->fe_surf_out.p; L.n_surf = c->fe_counts.p + 2;
  L.error_flag = c->fe_counts.p + 3;
  RG_CUDA(c, launch_features(L, c->stream));
  g_launches += 5;
  int h[4] = {0, 0, 0, 0};
  RG_CUDA(c, cudaMemcpyAsync(h, c->fe_counts.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (h[3] != 0) return set_err(c, RG_ERR_UNSUPPORTED, "rg_feature_extraction: sector/ring capacity exceeded or inconsistent ring indices");
  *n_corner = h[1]; *n_surface = h[2];
  if (curvature) RG_CUDA(c, cudaMemcpyAsync(curvature, c->fe_curv.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (picked) RG_CUDA(c, cudaMemcpyAsync(picked, c->fe_picked.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (label) RG_CUDA(c, cudaMemcpyAsync(label, c->fe_label.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (corner && h[1] > 0) RG_CUDA(c, cudaMemcpyAsync(corner, c->fe_corner_out.p, (size_t)h[1] * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  if (surface && h[2] > 0) RG_CUDA(c, cudaMemcpyAsync(surface, c->fe_surf_out.p, (size_t)h[2] * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  return RG_OK;
}

}  // extern "C"

__global__ void concat_clouds_kernel(const float4* __restrict__ a, int na, const float4* __restrict__ b, int nb, int n_pad, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  float4 p = make_float4(0.f, 0.f, 0.f, 1.f);
  if (i < na) p = a[i]; else if (i < na + nb) p = b[i - na];
  p.w = 1.f;
  out[i] = p;
}

extern "C" int rg_front_end_frame(rg_ctx* c, const void* raw, size_t n_points, const rg_projection_params* pp, const rg_feature_params* fp, int which,
                                  int32_t* n_corner, int32_t* n_surface) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!pp || !fp || (!raw && n_points)) return set_err(c, RG_ERR_INVALID, "rg_front_end_frame: null argument");
  if (pp->n_scan != fp->n_scan || pp->horizon_scan != fp->horizon_scan || pp->n_scan < 1 || pp->horizon_scan < 1 || pp->downsample_rate < 1 ||
      pp->point_stride_bytes < 12 || pp->point_stride_bytes % 4 || pp->ring_offset_bytes < 0 || pp->ring_offset_bytes + 2 > pp->point_stride_bytes ||
      !(fp->surf_leaf_size > 0.f))
    return set_err(c, RG_ERR_INVALID, "rg_front_end_frame: bad parameters");
  if (pp->horizon_scan > FE_RING_CAP) return set_err(c, RG_ERR_UNSUPPORTED, "rg_front_end_frame: horizon_scan above the per-ring capacity (2048)");
  if (int rc = fe_alloc(c, pp->n_scan, pp->horizon_scan)) return rc;
  const int cells = pp->n_scan * pp->horizon_scan;
  RG_CUDA(c, c->fe_raw.ensure(c->device, std::max<size_t>(n_points * pp->point_stride_bytes, 16)));
  if (n_points) RG_CUDA(c, cudaMemcpyAsync(c->fe_raw.p, raw, n_points * pp->point_stride_bytes, cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, cudaMemsetAsync(c->fe_counts.p, 0, 8 * sizeof(int), c->stream));
  ProjLaunch P;
  P.raw = c->fe_raw.p; P.n_points = (int)n_points; P.stride = pp->point_stride_bytes; P.ring_offset = pp->ring_offset_bytes;
  P.n_scan = pp->n_scan; P.horizon = pp->horizon_scan; P.downsample = pp->downsample_rate; P.min_range = pp->lidar_min_range; P.max_range = pp->lidar_max_range;
  P.owner = c->fe_owner.p; P.range_image = c->fe_range_image.p; P.full = c->fe_full.p; P.row_count = c->fe_row_count.p;
  P.extracted = c->fe_extracted.p; P.col_ind = c->fe_col.p; P.point_range = c->fe_prange.p; P.start_ring = c->fe_start.p; P.end_ring = c->fe_end.p;
  P.n_extracted = c->fe_counts.p;
  RG_CUDA(c, launch_projection(P, c->stream));
  FeatLaunch F;
  F.extracted = c->fe_extracted.p; F.col_ind = c->fe_col.p; F.point_range = c->fe_prange.p; F.start_ring = c->fe_start.p; F.end_ring = c->fe_end.p;
  F.n_dev = c->fe_counts.p; F.n = cells; F.n_scan = fp->n_scan; F.horizon = fp->horizon_scan;
  F.edge_thr = fp->edge_threshold; F.surf_thr = fp->surf_threshold; F.leaf = fp->surf_leaf_size;
  F.curvature = c->fe_curv.p; F.picked = c->fe_picked.p; F.label = c->fe_label.p;
  F.corner_ring = c->fe_corner_ring.p; F.corner_cnt = c->fe_corner_cnt.p; F.surf_ring = c->fe_surf_ring.p; F.surf_cnt = c->fe_surf_cnt.p;
  F.corner_out = c->fe_corner_out.p; F.n_corner = c->fe_counts.p + 1; F.surf_out = c->fe_surf_out.p; F.n_surf = c->fe_counts.p + 2;
  F.error_flag = c->fe_counts.p + 3;
  RG_CUDA(c, launch_features(F, c->stream));
  g_launches += 10;
  int h[4] = {0, 0, 0, 0};
  RG_CUDA(c, cudaMemcpyAsync(h, c->fe_counts.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (h[3] != 0) return set_err(c, RG_ERR_UNSUPPORTED, "rg_front_end_frame: sector/ring capacity exceeded");
  if (n_corner) *n_corner = h[1];
  if (n_surface) *n_surface = h[2];
  c->fe_frame_ok = true; c->fe_n_scan = pp->n_scan; c->fe_n[0] = h[0]; c->fe_n[1] = h[1]; c->fe_n[2] = h[2];
  Cloud& cl = which == RG_SOURCE ? c->src : c->tgt;
  const int n = h[1] + h[2];
  const int n_pad = (n + TILE_PTS - 1) / TILE_PTS * TILE_PTS;
  RG_CUDA(c, cl.pts.ensure(c->device, std::max(n_pad, TILE_PTS)));
  RG_CUDA(c, cl.cov6.ensure(c->device, (size_t)std::max(n_pad, TILE_PTS) * COV_PLANES));
  cl.n = n; cl.n_pad = n_pad; cl.has_cov = false; cl.host_identity = nullptr;
  if (n > 0) {
    g_launches++;
    concat_clouds_kernel<<<(n_pad + 255) / 256, 256, 0, c->stream>>>(c->fe_corner_out.p, h[1], c->fe_surf_out.p, h[2], n_pad, cl.pts.p);
    RG_CUDA(c, cudaGetLastError());
  }
  if (which == RG_TARGET) c->voxel_ok = false;
  c->st_ok = false; c->have_corr = false;
  if (which == RG_SOURCE) c->shard_lo = c->shard_hi = 0;
  return RG_OK;
}

// ---- rolo/CloudInfoStamp wire image at the ROS edge (cloud_info.hpp) ------------------------------------------------

// float4 (x, y, z, intensity) -> 32-byte pcl::PointXYZI record (x y z 1 | intensity 0 0 0)
__global__ void pack_xyzi_kernel(const float4* __restrict__ in, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  out[2 * i] = make_float4(p.x, p.y, p.z, 1.f);
  out[2 * i + 1] = make_float4(p.w, 0.f, 0.f, 0.f);
}

struct WireCloud { const unsigned char* data; int n, step, ox, oy, oz; };

// corner + surface point records of a serialised message -> the context's float4 cloud (featureLast = corner + surf)
__global__ void unpack_features_kernel(WireCloud a, WireCloud b, int n_pad, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  float4 p = make_float4(0.f, 0.f, 0.f, 1.f);
  const WireCloud* w = nullptr; int j = 0;
  if (i < a.n) { w = &a; j = i; } else if (i < a.n + b.n) { w = &b; j = i - a.n; }
  if (w) {
    const unsigned char* r = w->data + (size_t)j * w->step;
    p.x = *reinterpret_cast<const float*>(r + w->ox); p.y = *reinterpret_cast<const float*>(r + w->oy); p.z = *reinterpret_cast<const float*>(r + w->oz);
  }
  out[i] = p;
}

extern "C" int rg_cloud_info_serialize(rg_ctx* c, int stage, const rg_msg_header* hdr, const rg_cloud_info_scalars* sc, uint8_t* out, size_t capacity,
                                       size_t* written) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!hdr || !written || (stage != 0 && stage != 1)) return set_err(c, RG_ERR_INVALID, "rg_cloud_info_serialize: bad arguments");
  if (!c->fe_frame_ok) return set_err(c, RG_ERR_INVALID, "rg_cloud_info_serialize: no front-end frame resident (call rg_front_end_frame first)");
  static const float no_cov[1] = {0.f};
  wire::Scalars s{};
  if (sc) {
    s.startOrientation = sc->start_orientation; s.endOrientation = sc->end_orientation; s.orientationDiff = sc->orientation_diff;
    for (int i = 0; i < 6; i++) s.initialGuess[i] = sc->initial_guess[i];
    s.covariance = sc->n_covariance ? sc->covariance : no_cov; s.n_covariance = sc->n_covariance; s.odomAvailable = sc->odom_available;
    if (sc->n_covariance && !sc->covariance) return set_err(c, RG_ERR_INVALID, "rg_cloud_info_serialize: null covariance");
  } else { s.covariance = no_cov; }
  const char* frame = hdr->frame_id ? hdr->frame_id : "";
  const wire::Header h{hdr->seq, hdr->stamp_sec, hdr->stamp_nsec, frame, std::strlen(frame)};
  wire::Sections sec;
  wire::Writer w(out, capacity);
  const size_t total = wire::write_cloud_info(w, stage, h, s, (uint32_t)c->fe_n_scan, (uint32_t)c->fe_n[0], (uint32_t)c->fe_n[1], (uint32_t)c->fe_n[2], sec);
  *written = total;
  if (!out) return RG_OK;
  if (w.overflow) return set_err(c, RG_ERR_INVALID, "rg_cloud_info_serialize: output buffer too small");
  const int n0 = c->fe_n[0], n1 = stage == 1 ? c->fe_n[1] : 0, n2 = stage == 1 ? c->fe_n[2] : 0;
  const size_t recs = (size_t)n0 + n1 + n2;
  RG_CUDA(c, c->fe_wire.ensure(c->device, std::max<size_t>(recs * wire::POINT_STEP, 32)));
  float4* stage_buf = reinterpret_cast<float4*>(c->fe_wire.p);
  const struct { const float4* src; int n; size_t off; } part[3] = {{c->fe_extracted.p, n0, sec.projected}, {c->fe_corner_out.p, n1, sec.corner}, {c->fe_surf_out.p, n2, sec.surface}};
  size_t rec0 = 0;
  for (int k = 0; k < 3; k++) {
    if (part[k].n <= 0) continue;
    g_launches++;
    pack_xyzi_kernel<<<(part[k].n + 255) / 256, 256, 0, c->stream>>>(part[k].src, part[k].n, stage_buf + 2 * rec0);
    RG_CUDA(c, cudaGetLastError());
    RG_CUDA(c, cudaMemcpyAsync(out + part[k].off, stage_buf + 2 * rec0, (size_t)part[k].n * wire::POINT_STEP, cudaMemcpyDeviceToHost, c->stream));
    rec0 += part[k].n;
  }
  if (stage == 0) {
    RG_CUDA(c, cudaMemcpyAsync(out + sec.start_ring, c->fe_start.p, 4ull * c->fe_n_scan, cudaMemcpyDeviceToHost, c->stream));
    RG_CUDA(c, cudaMemcpyAsync(out + sec.end_ring, c->fe_end.p, 4ull * c->fe_n_scan, cudaMemcpyDeviceToHost, c->stream));
    if (n0 > 0) {
      RG_CUDA(c, cudaMemcpyAsync(out + sec.col_ind, c->fe_col.p, 4ull * n0, cudaMemcpyDeviceToHost, c->stream));
      RG_CUDA(c, cudaMemcpyAsync(out + sec.point_range, c->fe_prange.p, 4ull * n0, cudaMemcpyDeviceToHost, c->stream));
    }
  }
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  return RG_OK;
}

extern "C" int rg_set_input_from_cloud_info(rg_ctx* c, int which, const uint8_t* msg, size_t len, int32_t* n_corner, int32_t* n_surface) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!msg || (which != RG_SOURCE && which != RG_TARGET)) return set_err(c, RG_ERR_INVALID, "rg_set_input_from_cloud_info: bad arguments");
  wire::Parsed P;
  if (!wire::parse_cloud_info(msg, len, P)) return set_err(c, RG_ERR_INVALID, "rg_set_input_from_cloud_info: malformed CloudInfoStamp message");
  const wire::CloudView* cv[2] = {&P.corner, &P.surface};
  for (int k = 0; k < 2; k++) {
    const wire::CloudView& v = *cv[k];
    if (v.n_points == 0) continue;
    if (v.off_x < 0 || v.off_y < 0 || v.off_z < 0 || (v.off_x | v.off_y | v.off_z | (int)v.point_step) % 4 != 0 ||
        (uint32_t)std::max(v.off_x, std::max(v.off_y, v.off_z)) + 4 > v.point_step)
      return set_err(c, RG_ERR_UNSUPPORTED, "rg_set_input_from_cloud_info: feature clouds need 4-byte aligned float32 x/y/z fields");
  }
  const size_t bytes_a = (size_t)P.corner.n_points * P.corner.point_step, bytes_b = (size_t)P.surface.n_points * P.surface.point_step;
  const size_t off_b = (bytes_a + 15) / 16 * 16;
  RG_CUDA(c, c->fe_wire.ensure(c->device, std::max<size_t>(off_b + bytes_b, 32)));
  if (bytes_a) RG_CUDA(c, cudaMemcpyAsync(c->fe_wire.p, msg + P.corner.data, bytes_a, cudaMemcpyHostToDevice, c->stream));
  if (bytes_b) RG_CUDA(c, cudaMemcpyAsync(c->fe_wire.p + off_b, msg + P.surface.data, bytes_b, cudaMemcpyHostToDevice, c->stream));
  Cloud& cl = which == RG_SOURCE ? c->src : c->tgt;
  const int n = (int)(P.corner.n_points + P.surface.n_points);
  const int n_pad = (n + TILE_PTS - 1) / TILE_PTS * TILE_PTS;
  RG_CUDA(c, cl.pts.ensure(c->device, std::max(n_pad, TILE_PTS)));
  RG_CUDA(c, cl.cov6.ensure(c->device, (size_t)std::max(n_pad, TILE_PTS) * COV_PLANES));
  cl.n = n; cl.n_pad = n_pad; cl.has_cov = false; cl.host_identity = nullptr;
  if (n > 0) {
    const WireCloud a{c->fe_wire.p, (int)P.corner.n_points, (int)P.corner.point_step, P.corner.off_x, P.corner.off_y, P.corner.off_z};
    const WireCloud b{c->fe_wire.p + off_b, (int)P.surface.n_points, (int)P.surface.point_step, P.surface.off_x, P.surface.off_y, P.surface.off_z};
    g_launches++;
    unpack_features_kernel<<<(n_pad + 255) / 256, 256, 0, c->stream>>>(a, b, n_pad, cl.pts.p);
    RG_CUDA(c, cudaGetLastError());
  }
  RG_CUDA(c, cudaStreamSynchronize(c->stream));      // msg may be pageable and reused by the caller
  if (n_corner) *n_corner = (int32_t)P.corner.n_points;
  if (n_surface) *n_surface = (int32_t)P.surface.n_points;
  if (which == RG_TARGET) c->voxel_ok = false;
  c->st_ok = false; c->have_corr = false;
  if (which == RG_SOURCE) c->shard_lo = c->shard_hi = 0;
  return RG_OK;
}

// ---- backMapping scan-to-map optimisation (backMapping.cpp:681-1054) ------------------------------

extern "C" int rg_s2m_set_map(rg_ctx* c, const float* corner, size_t n_corner, const float* surf, size_t n_surf, size_t stride_bytes) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  c->s2m_map_ok = false; c->s2m_lin_ok = false;
  const float* src[2] = {corner, surf}; const size_t cnt[2] = {n_corner, n_surf};
  for (int m = 0; m < 2; m++) {
    if (int rc = upload_cloud(c, c->s2m_cloud[m], src[m], cnt[m], stride_bytes)) return rc;
    const int n = (int)cnt[m];
    if (n == 0) continue;
    // kdtree*FromMap->setInputCloud (backMapping.cpp:689-690): here the Morton index of covariance.cu
    RG_CUDA(c, c->codes_in.ensure(c->device, n)); RG_CUDA(c, c->idx_in.ensure(c->device, n)); RG_CUDA(c, c->idx_out.ensure(c->device, n));
    RG_CUDA(c, c->s2m_codes[m].ensure(c->device, n)); RG_CUDA(c, c->s2m_sorted[m].ensure(c->device, n));
    RG_CUDA(c, c->s2m_bbox[m].ensure(c->device, 8)); RG_CUDA(c, c->s2m_bucket[m].ensure(c->device, cov_bucket_table_entries()));
    const size_t tb = cov_cub_temp_bytes(n);
    RG_CUDA(c, c->cub_temp.ensure(c->device, tb));
    CovScratch sc;
    sc.codes_in = c->codes_in.p; sc.codes_out = c->s2m_codes[m].p; sc.idx_in = c->idx_in.p; sc.idx_out = c->idx_out.p; sc.sorted = c->s2m_sorted[m].p;
    sc.cub_temp = c->cub_temp.p; sc.cub_temp_bytes = tb; sc.bbox = c->s2m_bbox[m].p; sc.bucket_start = c->s2m_bucket[m].p;
    RG_CUDA(c, launch_sort_cloud(c->s2m_cloud[m].pts.p, n, sc, true, c->stream));
    g_launches += 5;
  }
  c->s2m_map_ok = true;
  return RG_OK;
}

extern "C" int rg_s2m_set_scan(rg_ctx* c, const float* corner, size_t n_corner, const float* surf, size_t n_surf, size_t stride_bytes) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  c->s2m_scan_ok = false; c->s2m_lin_ok = false;
  // the intensity column is not used by the optimisation: xyz only
  if (int rc = upload_cloud(c, c->s2m_cloud[2], corner, n_corner, stride_bytes)) return rc;
  if (int rc = upload_cloud(c, c->s2m_cloud[3], surf, n_surf, stride_bytes)) return rc;
  for (int m = 0; m < 2; m++) {
    RG_CUDA(c, c->s2m_coeff[m].ensure(c->device, std::max(c->s2m_cloud[2 + m].n, 1)));
    RG_CUDA(c, c->s2m_flag[m].ensure(c->device, std::max(c->s2m_cloud[2 + m].n, 1)));
  }
  c->s2m_scan_ok = true;
  return RG_OK;
}

// one kernel launch = cornerOptimization + surfOptimization + combineOptimizationCoeffs + the AtA/AtB assembly at pose tf
static int s2m_evaluate(rg_ctx* c, const float tf[6], float AtA[36], float AtB[6], int* n_sel) {
  if (!c->s2m_map_ok || !c->s2m_scan_ok) return set_err(c, RG_ERR_INVALID, "scan-to-map: call rg_s2m_set_map and rg_s2m_set_scan first");
  const size_t s2m_total = (size_t)c->s2m_cloud[2].n + c->s2m_cloud[3].n;
  RG_CUDA(c, c->s2m_partial.ensure(c->device, ((s2m_total + S2M_FIT_THREADS_PER_BLOCK - 1) / S2M_FIT_THREADS_PER_BLOCK + 1) * S2M_RES));
  RG_CUDA(c, c->s2m_nn.ensure(c->device, std::max<size_t>(s2m_total * 5, 8)));
  RG_CUDA(c, c->s2m_result.ensure(c->device, S2M_RES));
  if (!c->s2m_counter.p) {
    RG_CUDA(c, c->s2m_counter.ensure(c->device, 1));
    RG_CUDA(c, cudaMemsetAsync(c->s2m_counter.p, 0, sizeof(unsigned), c->stream));
  }
  S2MLaunch L;
  for (int m = 0; m < 2; m++) {
    L.map[m].pts = c->s2m_cloud[m].pts.p; L.map[m].sorted = c->s2m_sorted[m].p; L.map[m].codes = c->s2m_codes[m].p;
    L.map[m].bucket_start = c->s2m_bucket[m].p; L.map[m].bbox = c->s2m_bbox[m].p; L.map[m].n = c->s2m_cloud[m].n;
    L.scan[m] = c->s2m_cloud[2 + m].pts.p; L.n_scan[m] = c->s2m_cloud[2 + m].n;
    L.coeff[m] = c->s2m_coeff[m].p; L.flag[m] = c->s2m_flag[m].p;
  }
  s2m::get_transformation(tf, L.t);
  // LMOptimization's six sines / cosines (backMapping.cpp:949-954), evaluated on the host like the reference does
  L.sc[0] = sinf(tf[1]); L.sc[1] = cosf(tf[1]); L.sc[2] = sinf(tf[2]); L.sc[3] = cosf(tf[2]); L.sc[4] = sinf(tf[0]); L.sc[5] = cosf(tf[0]);
  L.nn = c->s2m_nn.p; L.partial = c->s2m_partial.p; L.counter = c->s2m_counter.p; L.result = c->s2m_result.p;
  RG_CUDA(c, launch_scan2map(L, c->stream));
  g_launches += 2;
  double r[S2M_RES];
  RG_CUDA(c, cudaMemcpyAsync(r, c->s2m_result.p, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  int t = 0;
  for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++, t++) AtA[a * 6 + b] = AtA[b * 6 + a] = (float)r[t];   // cv::gemm: double sums, float result
  for (int a = 0; a < 6; a++) AtB[a] = (float)r[21 + a];
  *n_sel = (int)(r[27] + 0.5);
  c->s2m_lin_ok = true;
  return RG_OK;
}

extern "C" int rg_s2m_linearize(rg_ctx* c, const float tf[6], float AtA[36], float AtB[6], int32_t* n_selected) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!tf || !AtA || !AtB || !n_selected) return set_err(c, RG_ERR_INVALID, "rg_s2m_linearize: null argument");
  int n = 0;
  if (int rc = s2m_evaluate(c, tf, AtA, AtB, &n)) return rc;
  *n_selected = n;
  return RG_OK;
}

extern "C" int rg_s2m_get_coefficients(rg_ctx* c, float* coeff_corner, uint8_t* flag_corner, float* coeff_surf, uint8_t* flag_surf) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!c->s2m_lin_ok) return set_err(c, RG_ERR_INVALID, "rg_s2m_get_coefficients: no evaluation yet");
  float* cf[2] = {coeff_corner, coeff_surf}; uint8_t* fl[2] = {flag_corner, flag_surf};
  for (int m = 0; m < 2; m++) {
    const int n = c->s2m_cloud[2 + m].n;
    if (n == 0) continue;
    if (cf[m]) RG_CUDA(c, cudaMemcpyAsync(cf[m], c->s2m_coeff[m].p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (fl[m]) RG_CUDA(c, cudaMemcpyAsync(fl[m], c->s2m_flag[m].p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  }
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  return RG_OK;
}

// scan2MapOptimization (backMapping.cpp:681-710) + LMOptimization (:937-1054); transformUpdate()'s IMU blend is not part of it
extern "C" int rg_s2m_optimize(rg_ctx* c, float tf[6], int max_iterations, int edge_feature_min, int surf_feature_min, rg_s2m_result* out) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!tf || !out || max_iterations < 0) return set_err(c, RG_ERR_INVALID, "rg_s2m_optimize: bad arguments");
  if (!c->s2m_map_ok || !c->s2m_scan_ok) return set_err(c, RG_ERR_INVALID, "scan-to-map: call rg_s2m_set_map and rg_s2m_set_scan first");
  out->iterations = 0; out->converged = 0; out->degenerate = 0; out->enough_features = 0;
  if (!(c->s2m_cloud[2].n > edge_feature_min && c->s2m_cloud[3].n > surf_feature_min)) return RG_OK;   // "Not enough features!" (:687,705-707)
  out->enough_features = 1;
  bool degenerate = false;
  float matP[36];
  for (int iter = 0; iter < max_iterations; iter++) {
    float AtA[36], AtB[6], X[6];
    int sel = 0;
    if (int rc = s2m_evaluate(c, tf, AtA, AtB, &sel)) return rc;
    out->iterations++;
    if (sel < 50) continue;                                                       // :956-958
    { float A[36], b[6]; std::memcpy(A, AtA, sizeof(A)); std::memcpy(b, AtB, sizeof(b)); s2m::qr_solve<6, 6, false>(A, b, X); }   // cv::solve(DECOMP_QR)
    if (iter == 0) {                                                              // degeneracy analysis, first iteration only (:1008-1029)
      float A[36], E[6], V[36], V2[36], Vi[36];
      std::memcpy(A, AtA, sizeof(A));
      s2m::jacobi_eigen<6>(A, E, V);
      std::memcpy(V2, V, sizeof(V2));
      degenerate = false;
      for (int i = 5; i >= 0; i--) {
        if (E[i] < 100.f) { for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0.f; degenerate = true; } else break;
      }
      if (!s2m::invert_lu<6>(V, Vi)) std::memset(Vi, 0, sizeof(Vi));
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) { double acc = 0.0; for (int k = 0; k < 6; k++) acc += (double)Vi[a * 6 + k] * (double)V2[k * 6 + b]; matP[a * 6 + b] = (float)acc; }
    }
    if (degenerate) {
      float X2[6]; std::memcpy(X2, X, sizeof(X2));
      for (int a = 0; a < 6; a++) { double acc = 0.0; for (int k = 0; k < 6; k++) acc += (double)matP[a * 6 + k] * (double)X2[k]; X[a] = (float)acc; }
    }
    for (int a = 0; a < 6; a++) tf[a] += X[a];
    const float r0 = X[0] * 57.29578f, r1 = X[1] * 57.29578f, r2 = X[2] * 57.29578f, c0 = X[3] * 100, c1 = X[4] * 100, c2 = X[5] * 100;
    const float deltaR = (float)std::sqrt((double)r0 * r0 + (double)r1 * r1 + (double)r2 * r2);
    const float deltaT = (float)std::sqrt((double)c0 * c0 + (double)c1 * c1 + (double)c2 * c2);
    if ((double)deltaR < 0.05 && (double)deltaT < 0.05) { out->converged = 1; break; }
  }
  out->degenerate = degenerate ? 1 : 0;
  return RG_OK;
}

// ---- Scan Context place recognition (src/scancontext/Scancontext.cpp) ------------------------------

extern "C" int rg_sc_make_and_save(rg_ctx* c, const float* pts, size_t n, size_t stride_bytes, double lidar_height) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (int rc = upload_cloud(c, c->sc_cloud, pts, n, stride_bytes)) return rc;
  RG_CUDA(c, c->sc_bins.ensure(c->device, SC_BINS));
  RG_CUDA(c, c->sc_out.ensure(c->device, SC_BINS + SC_NUM_RING + SC_NUM_SECTOR));
  SCLaunch L;
  L.pts = c->sc_cloud.pts.p; L.n = (int)n; L.lidar_height = lidar_height; L.max_radius = 80.0;   // PC_MAX_RADIUS (Scancontext.h:84)
  L.bins = c->sc_bins.p; L.desc = c->sc_out.p; L.ring_key = c->sc_out.p + SC_BINS; L.sector_key = c->sc_out.p + SC_BINS + SC_NUM_RING;
  RG_CUDA(c, launch_scancontext(L, c->stream));
  g_launches += 3;
  double h[SC_BINS + SC_NUM_RING + SC_NUM_SECTOR];
  RG_CUDA(c, cudaMemcpyAsync(h, c->sc_out.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  sc::Entry e;
  std::memcpy(e.desc, h, sizeof(e.desc)); std::memcpy(e.ring_key, h + SC_BINS, sizeof(e.ring_key)); std::memcpy(e.sector_key, h + SC_BINS + SC_NUM_RING, sizeof(e.sector_key));
  for (int k = 0; k < SC_NUM_RING; k++) e.ring_key_f[k] = (float)e.ring_key[k];      // eig2stdvec (Scancontext.cpp:62-66)
  c->sc_mgr.entries.push_back(e);
  return RG_OK;
}

extern "C" int rg_sc_size(rg_ctx* c) { return c ? (int)c->sc_mgr.entries.size() : 0; }

extern "C" int rg_sc_clear(rg_ctx* c) {
  CHECK_CTX(c);
  c->sc_mgr = sc::Manager();
  return RG_OK;
}

static const sc::Entry* sc_entry(rg_ctx* c, int index) {
  const int n = (int)c->sc_mgr.entries.size();
  if (index < 0) index += n;
  return (index >= 0 && index < n) ? &c->sc_mgr.entries[index] : nullptr;
}

extern "C" int rg_sc_get(rg_ctx* c, int index, double* desc, double* ring_key, double* sector_key) {
  CHECK_CTX(c);
  const sc::Entry* e = sc_entry(c, index);
  if (!e) return set_err(c, RG_ERR_INVALID, "rg_sc_get: no such descriptor");
  if (desc) std::memcpy(desc, e->desc, sizeof(e->desc));
  if (ring_key) std::memcpy(ring_key, e->ring_key, sizeof(e->ring_key));
  if (sector_key) std::memcpy(sector_key, e->sector_key, sizeof(e->sector_key));
  return RG_OK;
}

extern "C" int rg_sc_distance(rg_ctx* c, int i, int j, double* dist, int32_t* shift) {
  CHECK_CTX(c);
  const sc::Entry *a = sc_entry(c, i), *b = sc_entry(c, j);
  if (!a || !b || !dist || !shift) return set_err(c, RG_ERR_INVALID, "rg_sc_distance: bad arguments");
  const std::pair<double, int> r = sc::distance(*a, *b);
  *dist = r.first; *shift = r.second;
  return RG_OK;
}

extern "C" int rg_sc_detect(rg_ctx* c, int32_t* loop_id, float* yaw_diff_rad) {
  CHECK_CTX(c);
  if (!loop_id || !yaw_diff_rad) return set_err(c, RG_ERR_INVALID, "rg_sc_detect: null argument");
  if (c->sc_mgr.entries.empty()) return set_err(c, RG_ERR_INVALID, "rg_sc_detect: no descriptor saved yet");
  const std::pair<int, float> r = c->sc_mgr.detect();
  *loop_id = r.first; *yaw_diff_rad = r.second;
  return RG_OK;
}

// ---- pcl::Registration::getFitnessScore / loop-closure verification (backMapping.cpp:2351,2428) -------

// nearest target point of every source point moved by T (row-major 4x4 float)
static int nearest_target(rg_ctx* c, const float T[16]) {
  if (c->src.n == 0 || c->tgt.n == 0) return set_err(c, RG_ERR_INVALID, "fitness score: source and target clouds must be set");
  const int n = c->tgt.n, nq = c->src.n;
  RG_CUDA(c, c->codes_in.ensure(c->device, n)); RG_CUDA(c, c->idx_in.ensure(c->device, n)); RG_CUDA(c, c->idx_out.ensure(c->device, n));
  RG_CUDA(c, c->fit_codes.ensure(c->device, n)); RG_CUDA(c, c->fit_sorted.ensure(c->device, n));
  RG_CUDA(c, c->fit_bbox.ensure(c->device, 8)); RG_CUDA(c, c->fit_bucket.ensure(c->device, cov_bucket_table_entries()));
  const size_t tb = cov_cub_temp_bytes(n);
  RG_CUDA(c, c->cub_temp.ensure(c->device, tb));
  RG_CUDA(c, c->fit_idx.ensure(c->device, nq)); RG_CUDA(c, c->fit_d2.ensure(c->device, nq)); RG_CUDA(c, c->fit_T.ensure(c->device, 12));
  CovScratch sc;
  sc.codes_in = c->codes_in.p; sc.codes_out = c->fit_codes.p; sc.idx_in = c->idx_in.p; sc.idx_out = c->idx_out.p; sc.sorted = c->fit_sorted.p;
  sc.cub_temp = c->cub_temp.p; sc.cub_temp_bytes = tb; sc.bbox = c->fit_bbox.p; sc.bucket_start = c->fit_bucket.p;
  RG_CUDA(c, launch_sort_cloud(c->tgt.pts.p, n, sc, true, c->stream));          // tree_->setInputCloud(target_) of pcl::Registration
  RG_CUDA(c, cudaMemcpyAsync(c->fit_T.p, T, 12 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, launch_nearest(c->src.pts.p, nq, c->fit_T.p, sc, n, c->fit_idx.p, c->fit_d2.p, c->stream));
  g_launches += 6;
  return RG_OK;
}

extern "C" int rg_nearest_target(rg_ctx* c, const float T[16], int32_t* idx, float* sq_dist) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!T) return set_err(c, RG_ERR_INVALID, "rg_nearest_target: null transform");
  if (int rc = nearest_target(c, T)) return rc;
  if (idx) RG_CUDA(c, cudaMemcpyAsync(idx, c->fit_idx.p, sizeof(int) * (size_t)c->src.n, cudaMemcpyDeviceToHost, c->stream));
  if (sq_dist) RG_CUDA(c, cudaMemcpyAsync(sq_dist, c->fit_d2.p, sizeof(float) * (size_t)c->src.n, cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  return RG_OK;
}

extern "C" int rg_get_fitness_score(rg_ctx* c, const float T[16], double max_range, double* score) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (!T || !score) return set_err(c, RG_ERR_INVALID, "rg_get_fitness_score: null argument");
  if (int rc = nearest_target(c, T)) return rc;
  std::vector<float> d2((size_t)c->src.n);
  RG_CUDA(c, cudaMemcpyAsync(d2.data(), c->fit_d2.p, sizeof(float) * d2.size(), cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  // the reference's accumulation: doubles, in point order, squared distance compared with max_range as is
  double fitness = 0.0; int nr = 0;
  for (float d : d2) if ((double)d <= max_range) { fitness += (double)d; nr++; }
  *score = nr > 0 ? fitness / nr : 1.7976931348623157e308;
  return RG_OK;
}
