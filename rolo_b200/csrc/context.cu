// context.cu — registration context, device-memory pool, host drivers and the registration part of the C-ABI
// (include/rolo_gicp.h); the front-end and mapping entry points live in abi_frontend.cu / abi_mapping.cu over context.hpp.
//
// The product path is CUDA-only: every entry point fails with RG_ERR_NO_DEVICE / RG_ERR_CUDA when no
// B200-class device is usable; there is no CPU fallback anywhere in this library.

#include "context.hpp"
#include <limits>

std::mutex g_mutex;
std::atomic<unsigned long long> g_launches{0};
Pool& pool() { static Pool* p = new Pool; return *p; }   // intentionally leaked: safe at process exit

namespace {

std::string g_last_error;
std::atomic<unsigned> g_seq{0};   // launch sequence numbers written to the mapped completion flags

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

}  // namespace

// ---- helpers shared with abi_frontend.cu / abi_mapping.cu (declared in context.hpp) and local ones ----

int set_err(const rg_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  { std::lock_guard<std::mutex> lk(g_mutex); g_last_error = msg; }
  return code;
}

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
  // no lazy join here (use_device would order the shared batch stream after this pair's previous upload: the source upload after the
  // target upload of every pair, i.e. the batch stream after nearly all copies — the overlap of copies and kernels would be gone)
  { cudaError_t e = cudaSetDevice(c->device); if (e != cudaSuccess) return set_err(c, RG_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
  if (n > 0 && !host) return set_err(c, RG_ERR_INVALID, "null cloud pointer");
  if (n > 0 && (stride_bytes < 12 || stride_bytes % 4)) return set_err(c, RG_ERR_INVALID, "stride_bytes must be a multiple of 4 and >= 12");
  if (n > (size_t)1 << 30) return set_err(c, RG_ERR_INVALID, "cloud too large");
  const int n_pad = (int)((n + TILE_PTS - 1) / TILE_PTS * TILE_PTS);
  RG_CUDA(c, cl.pts.ensure(c->device, std::max(n_pad, TILE_PTS)));
  RG_CUDA(c, cl.cov6.ensure(c->device, (size_t)std::max(n_pad, TILE_PTS) * COV_PLANES));
  cl.n = (int)n; cl.n_pad = n_pad; cl.has_cov = false; cl.host_identity = host; cl.shard_rank = 0; cl.shard_world = 1; cl.sorted_view = nullptr;
  if (n == 0) return RG_OK;
  cudaStream_t us = c->upload_stream ? c->upload_stream : c->stream;
  if (us != c->stream) {   // order after whatever the context queued before (e.g. the previous frame still using these buffers)
    RG_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));
    RG_CUDA(c, cudaStreamWaitEvent(us, c->ev_upload, 0));
  }
  // pageable memory: the runtime stages the copy before returning; page-locked memory: truly asynchronous, the caller keeps
  // the cloud alive until the next solving call returns (pcl::Registration holds the cloud's shared_ptr, as in the reference)
  if (stride_bytes == sizeof(float4)) {
    // 16-byte records (pcl::PointXYZ, float4): the copy engine writes the device layout directly — no staging buffer and no kernel
    // that would have to squeeze in beside the persistent search kernels of an earlier group (no kernel reads the 4th component)
    RG_CUDA(c, cudaMemcpyAsync(cl.pts.p, host, n * sizeof(float4), cudaMemcpyHostToDevice, us));
    if ((size_t)n_pad > n) RG_CUDA(c, cudaMemsetAsync(cl.pts.p + n, 0, ((size_t)n_pad - n) * sizeof(float4), us));
  } else {
    if (c->staging.cap < n * stride_bytes) RG_CUDA(c, cudaStreamSynchronize(us));   // about to re-allocate a buffer in-flight work may read
    RG_CUDA(c, c->staging.ensure(c->device, n * stride_bytes));
    RG_CUDA(c, cudaMemcpyAsync(c->staging.p, host, n * stride_bytes, cudaMemcpyHostToDevice, us));
    g_launches++;
    extract_points_kernel<<<(n_pad + 255) / 256, 256, 0, us>>>(c->staging.p, stride_bytes, (int)n, n_pad, cl.pts.p);
    RG_CUDA(c, cudaGetLastError());
  }
  if (us != c->stream) {
    RG_CUDA(c, cudaEventRecord(c->ev_upload, us));
    // inside a batch the shared batch stream is NOT made to wait here: rg_batch_align joins the uploads group by group
    // (and use_device() joins lazily for any other call), so later pairs' copies overlap earlier pairs' kernels
    if (c->batch) c->side_pending = true;
    else RG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_upload, 0));
  }
  return RG_OK;
}

int upload_cloud_device(rg_ctx* c, Cloud& cl, const float* dev, size_t n, cudaStream_t user_stream) {
  if (use_device(c)) return RG_ERR_CUDA;
  if (n > 0 && !dev) return set_err(c, RG_ERR_INVALID, "null cloud pointer");
  const int n_pad = (int)((n + TILE_PTS - 1) / TILE_PTS * TILE_PTS);
  RG_CUDA(c, cl.pts.ensure(c->device, std::max(n_pad, TILE_PTS)));
  RG_CUDA(c, cl.cov6.ensure(c->device, (size_t)std::max(n_pad, TILE_PTS) * COV_PLANES));
  cl.n = (int)n; cl.n_pad = n_pad; cl.has_cov = false; cl.host_identity = dev; cl.shard_rank = 0; cl.shard_world = 1; cl.sorted_view = nullptr;
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

// RotVGICP::calculate_covariances (rot_vgicp_impl.hpp:421-496) of a GROUP of clouds (one launch chain for all of them) on `stream`,
// with the scratch `bufs`. knn_out (optional, single cloud only) receives the unordered neighbour sets.
int calc_cov_group(rg_ctx* c, const std::vector<Cloud*>& clouds, KnnBufs& bufs, cudaStream_t stream, int* knn_out = nullptr,
                   cudaEvent_t ev_knn0 = nullptr, cudaEvent_t ev_knn1 = nullptr) {
  if (clouds.empty()) return RG_OK;
  if (c->regularization_method < RG_REG_NONE || c->regularization_method > RG_REG_PLANE_S)
    return set_err(c, RG_ERR_UNSUPPORTED, "here must not be reached (unknown regularization method; reference abort()s, rot_vgicp_impl.hpp:472-474)");
  const int k = c->k_correspondences;
  if (k < 1 || k > 64) return set_err(c, RG_ERR_INVALID, "k_correspondences must be in [1,64]");
  size_t total = 0; int max_n = 0;
  for (Cloud* cl : clouds) {
    if (cl->n == 0) return set_err(c, RG_ERR_INVALID, "calculate_covariances: empty cloud");
    if (cl->n < k)
      return set_err(c, RG_ERR_INVALID, "cloud has fewer points than k_correspondences (reference reads uninitialised neighbours here, rot_vgicp_impl.hpp:438-444)");
    total += (size_t)cl->n; max_n = std::max(max_n, cl->n);
  }
  if (total > (size_t)1 << 30) return set_err(c, RG_ERR_INVALID, "calculate_covariances: group too large");
  // zero the padding rows of the covariance tiles (only the last tile of a cloud has any)
  for (Cloud* cl : clouds)
    if (cl->n_pad > cl->n) RG_CUDA(c, cudaMemsetAsync(cl->cov6.p + (size_t)(cl->n_pad / TILE_PTS - 1) * COV_PLANES * TILE_PTS, 0, sizeof(float) * COV_PLANES * TILE_PTS, stream));
  if (k > 32) {   // generic path: thread-per-query search into explicit neighbour lists, one cloud at a time
    for (Cloud* cl : clouds) {
      const int n = cl->n;
      RG_CUDA(c, c->codes_in.ensure(c->device, n)); RG_CUDA(c, c->codes_out.ensure(c->device, n));
      RG_CUDA(c, c->idx_in.ensure(c->device, n)); RG_CUDA(c, c->idx_out.ensure(c->device, n));
      RG_CUDA(c, c->sorted.ensure(c->device, n)); RG_CUDA(c, c->bbox.ensure(c->device, 8)); RG_CUDA(c, c->bucket_start.ensure(c->device, cov_bucket_table_entries()));
      const size_t tb = cov_cub_temp_bytes(n);
      RG_CUDA(c, c->cub_temp.ensure(c->device, tb));
      RG_CUDA(c, c->knn_idx.ensure(c->device, (size_t)n * k));
      CovLaunch L;
      L.pts = cl->pts.p; L.n = n; L.k = k; L.method = c->regularization_method;
      L.cov6 = cl->cov6.p; L.knn = c->knn_idx.p;
      L.s.codes_in = c->codes_in.p; L.s.codes_out = c->codes_out.p; L.s.idx_in = c->idx_in.p; L.s.idx_out = c->idx_out.p;
      L.s.sorted = c->sorted.p; L.s.cub_temp = c->cub_temp.p; L.s.cub_temp_bytes = tb; L.s.bbox = c->bbox.p; L.s.bucket_start = c->bucket_start.p;
      RG_CUDA(c, launch_covariances_generic(L, stream));
      g_launches += 6;
      if (knn_out) RG_CUDA(c, cudaMemcpyAsync(knn_out, c->knn_idx.p, sizeof(int) * (size_t)n * k, cudaMemcpyDeviceToDevice, stream));
      cl->has_cov = true;
    }
    return RG_OK;
  }
  RG_CUDA(c, bufs.ensure(c->device, total, (int)clouds.size()));
  if (&bufs == &c->knn) { c->src.sorted_view = nullptr; c->tgt.sorted_view = nullptr; }   // the Morton-ordered copies in this scratch are about to be overwritten
  std::vector<KnnCloud> desc(clouds.size());
  int off = 0, chunk_off = 0;
  for (size_t i = 0; i < clouds.size(); i++) {
    desc[i].pts = clouds[i]->pts.p; desc[i].cov6 = clouds[i]->cov6.p; desc[i].knn_out = (knn_out && clouds.size() == 1) ? knn_out : nullptr;
    desc[i].n = clouds[i]->n; desc[i].off = off; desc[i].chunk_off = chunk_off;
    desc[i].shard_rank = clouds[i]->shard_rank; desc[i].shard_world = clouds[i]->shard_world;
    clouds[i]->sorted_view = bufs.sorted.p + off;
    off += clouds[i]->n; chunk_off += knn_owned_chunks(clouds[i]->n, clouds[i]->shard_rank, clouds[i]->shard_world);
  }
  KnnGroup g;
  g.clouds = desc.data(); g.n_clouds = (int)desc.size(); g.total = (int)total; g.max_n = max_n; g.k = k; g.method = c->regularization_method;
  g.s = bufs.view();
  g.ev_knn0 = ev_knn0; g.ev_knn1 = ev_knn1;
  static const bool want_stats = std::getenv("ROLO_KNN_STATS") != nullptr;   // development aid: work counters of the chunk kernel on stderr
  DevBuf<unsigned long long> stats;
  if (want_stats) { RG_CUDA(c, stats.ensure(c->device, 8)); RG_CUDA(c, cudaMemsetAsync(stats.p, 0, 64, stream)); g.stats = stats.p; }
  int own = 0;
  RG_CUDA(c, launch_covariances_group(g, stream, &own));
  g_launches += own;
  if (want_stats) {
    unsigned long long h[8]; int nf[2] = {0, 0};
    cudaMemcpyAsync(h, stats.p, 64, cudaMemcpyDeviceToHost, stream); cudaMemcpyAsync(nf, bufs.fail.p, 8, cudaMemcpyDeviceToHost, stream); cudaStreamSynchronize(stream);
    const double ch = (double)std::max<unsigned long long>(h[6], 1);
    fprintf(stderr, "[knn stats] clouds %d points %zu chunks %llu | per chunk: raw %.1f staged %.1f events %.1f groups %.3f | pass/query %.1f | max raw %llu | fallback queries %d\n",
            (int)clouds.size(), total, h[6], h[0] / ch, h[1] / ch, h[2] / ch, h[3] / ch, h[4] / (ch * 32.0), h[5], nf[0]);
    stats.release();
  }
  for (Cloud* cl : clouds) cl->has_cov = true;
  return RG_OK;
}

// the share of the target this rank owns in a target-sharded registration
void apply_target_shard(rg_ctx* c) {
  Cloud& t = c->tgt;
  if (c->tgt_world > 1 && !c->batch) { t.shard_rank = c->tgt_rank; t.shard_world = c->tgt_world; }
  else { t.shard_rank = 0; t.shard_world = 1; }
}

int calc_cov(rg_ctx* c, Cloud& cl) {
  if (use_device(c)) return RG_ERR_CUDA;
  if (&cl == &c->tgt) apply_target_shard(c);
  std::vector<Cloud*> one{&cl};
  return calc_cov_group(c, one, c->knn, c->stream);
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

// buffers + launch description of VmfVoxelMap::create_voxelmap over this context's target
int prepare_voxel_build(rg_ctx* c, VoxelBuild& b) {
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
  b.pts = c->tgt.pts.p; b.cov6 = c->tgt.cov6.p; b.n = n; b.grid = make_grid(c);
  b.table = c->table.p; b.mask = c->mask; b.slot_of_point = c->slot_of_point.p; b.counters = c->counters.p;
  b.sums = c->sums.p; b.vox = c->vox.p;
  return RG_OK;
}

int build_voxelmap(rg_ctx* c) {
  if (use_device(c)) return RG_ERR_CUDA;
  VoxelBuild b;
  if (int rc = prepare_voxel_build(c, b)) return rc;
  if (c->tgt_world > 1 && !c->batch) {
    // target-sharded setup: accumulate this rank's slice, hand the compact voxel list to the merge hook, which gathers the lists
    // of all ranks and calls rg_voxel_import (the finalize pass runs there)
    if (!c->voxel_merge_fn) return set_err(c, RG_ERR_INVALID, "target-sharded registration without a voxel merge hook (rg_set_voxel_merge_hook)");
    if (!c->tgt.sorted_view) {   // covariances installed from outside, or the scratch was reused since: estimate this rank's slice again
      c->tgt.has_cov = false;
      if (int rc = calc_cov(c, c->tgt)) return rc;
      if (int rc = prepare_voxel_build(c, b)) return rc;
    }
    b.order = c->tgt.sorted_view; b.shard_rank = c->tgt.shard_rank; b.shard_world = c->tgt.shard_world;
    RG_CUDA(c, launch_voxel_build(b, c->stream, false));
    g_launches += 3;
    int h[2] = {0, 0};
    RG_CUDA(c, cudaMemcpyAsync(h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    RG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (h[1] != 0) return set_err(c, RG_ERR_INVALID, "voxel coordinate out of the +-2^20 key range (or table overflow)");
    RG_CUDA(c, c->vox_keys.ensure(c->device, std::max(h[0], 1)));
    RG_CUDA(c, launch_voxel_keys(b, c->vox_keys.p, c->stream));
    g_launches += 1;
    c->voxel_ok = false;
    if (int rc = c->voxel_merge_fn(c->voxel_merge_user, c->vox_keys.p, c->sums.p, h[0], (void*)c->stream))
      return set_err(c, RG_ERR_CUDA, "voxel merge hook failed (" + std::to_string(rc) + ")");
    if (!c->voxel_ok) return set_err(c, RG_ERR_INVALID, "voxel merge hook returned without calling rg_voxel_import");
    return RG_OK;
  }
  if (c->deterministic) {
    const int n = b.n;
    RG_CUDA(c, c->codes_in.ensure(c->device, n)); RG_CUDA(c, c->codes_out.ensure(c->device, n));
    RG_CUDA(c, c->idx_in.ensure(c->device, (size_t)n + 1)); RG_CUDA(c, c->idx_out.ensure(c->device, n));
    RG_CUDA(c, c->det_flags.ensure(c->device, (size_t)n + 1));
    VoxelDetScratch s;
    s.cub_temp_bytes = voxel_det_temp_bytes(n);
    RG_CUDA(c, c->cub_temp.ensure(c->device, s.cub_temp_bytes));
    s.keys_a = c->codes_in.p; s.keys_b = c->codes_out.p; s.idx_a = c->idx_in.p; s.idx_b = c->idx_out.p; s.flags = c->det_flags.p; s.cub_temp = c->cub_temp.p;
    RG_CUDA(c, launch_voxel_build_deterministic(b, s, c->stream));
    g_launches += 4;
  } else {
    RG_CUDA(c, launch_voxel_build(b, c->stream));
    g_launches += 4;
  }
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
// make sure the static descriptor + scratch for the fused kernel exist. Inside rg_batch_align (`batch_out` != null) only the
// descriptor is produced — the batch uploads all of them with one copy and builds its own persistent-CTA schedule; the stand-alone
// launch shape (own schedule + descriptor on the device, two synchronising copies) is then set up lazily, if the pair is ever
// evaluated on its own.
int prepare_linearize(rg_ctx* c, LinStatic* batch_out = nullptr) {
  if (use_device(c)) return RG_ERR_CUDA;
  if (int rc = check_grid(c)) return rc;
  if (c->src.n == 0) return set_err(c, RG_ERR_INVALID, "no source cloud");
  if (!c->src.has_cov) return set_err(c, RG_ERR_INVALID, "source covariances missing");
  if (!c->voxel_ok) { if (int rc = build_voxelmap(c)) return rc; }   // lazily, as rot_vgicp_impl.hpp:227-230
  const int noff = noff_of(c->search_method);
  if (c->st_ok && c->corr_noff == noff && !batch_out && c->st_standalone_ok) return RG_OK;
  const int lo = (c->shard_hi > c->shard_lo) ? c->shard_lo : 0;
  const int n_eval = (c->shard_hi > c->shard_lo) ? (c->shard_hi - c->shard_lo) : c->src.n;
  if (!(c->st_ok && c->corr_noff == noff)) {
    RG_CUDA(c, c->corr.ensure(c->device, (size_t)n_eval * noff));
    c->corr_noff = noff; c->have_corr = false;
    RG_CUDA(c, c->partials.ensure(c->device, (size_t)lin_slots(c->device) * RES_N));   // any launch grid fits
    if (!c->counter.p) { RG_CUDA(c, c->counter.ensure(c->device, 4)); RG_CUDA(c, cudaMemsetAsync(c->counter.p, 0, 4 * sizeof(unsigned), c->stream)); }
    RG_CUDA(c, c->result_dev.ensure(c->device, RES_N));
    RG_CUDA(c, c->st_dev.ensure(c->device, 1));
    c->st_standalone_ok = false;
  }
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
  c->st_ok = true;
  if (batch_out) { *batch_out = st; return RG_OK; }
  {   // stand-alone launch shape: one pair, persistent CTAs
    std::vector<int> sched;
    const int ntiles = (n_eval + TILE_PTS - 1) / TILE_PTS;
    c->lin_grid = std::max(1, std::min(ntiles, lin_slots(c->device)));
    build_schedule(&n_eval, 1, c->lin_grid, sched, c->lin_tiles);
    RG_CUDA(c, c->sched_dev.ensure(c->device, sched.size()));
    RG_CUDA(c, cudaMemcpyAsync(c->sched_dev.p, sched.data(), sched.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    RG_CUDA(c, cudaMemcpyAsync(c->st_dev.p, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
    if (c->batch) RG_CUDA(c, cudaMemcpyAsync(c->batch->st_dev.p + c->batch_index, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
    RG_CUDA(c, cudaStreamSynchronize(c->stream));   // `st` and `sched` are stack objects
  }
  c->st_standalone_ok = true;
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
  if (use_device(c)) return RG_ERR_CUDA;
  std::vector<Cloud*> todo;
  if (!c->src.has_cov) todo.push_back(&c->src);
  if (!c->tgt.has_cov) { apply_target_shard(c); todo.push_back(&c->tgt); }
  return calc_cov_group(c, todo, c->knn, c->stream);
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
  c->codes_in.release(); c->codes_out.release(); c->idx_in.release(); c->idx_out.release(); c->bbox.release(); c->sorted.release(); c->cub_temp.release(); c->knn_idx.release(); c->bucket_start.release(); c->knn.release(); c->vox_keys.release(); c->det_flags.release();
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

int rg_set_deterministic(rg_ctx* c, int on) { CHECK_CTX(c); c->deterministic = on != 0; c->voxel_ok = false; c->st_ok = false; c->have_corr = false; return RG_OK; }

int rg_reset_voxelmap(rg_ctx* c) { CHECK_CTX(c); c->voxel_ok = false; c->st_ok = false; c->have_corr = false; return RG_OK; }

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
  static const bool trace = std::getenv("ROLO_TRACE") != nullptr;   // development aid: stage wall times of every align on stderr
  const auto t0 = std::chrono::steady_clock::now();
  if (int rc = ensure_covariances(c)) return rc;
  if (trace) cudaStreamSynchronize(c->stream);
  const auto t1 = std::chrono::steady_clock::now();
  if (trace) { if (int rc = prepare_linearize(c)) return rc; cudaStreamSynchronize(c->stream); }
  const auto t2 = std::chrono::steady_clock::now();
  double g[16];
  if (guess) for (int i = 0; i < 16; i++) g[i] = (double)guess[i]; else host::mat4_identity(g);
  c->lsq.begin_align(g);
  if (int rc = run_driver(c, nullptr, nullptr, 0, 0)) return rc;
  if (trace) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[rolo trace] align: covariances %.3f ms, voxel map + launch set-up %.3f ms, %d iterations %.3f ms\n", ms(t0, t1), ms(t1, t2), c->lsq.nr_iterations + 1, ms(t2, t3));
  }
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
  // the product kernels (the same launch chain calculate_covariances runs) emit the neighbour SET of every point; the rows are
  // put into the documented ascending (distance, index) order on the host, with the float distance of the reference's kd-tree
  DevBuf<int> dk;
  RG_CUDA(c, dk.ensure(c->device, (size_t)n * k));
  const int saved_k = c->k_correspondences;
  const bool had_cov = cl.has_cov;
  DevBuf<float> keep_cov;                      // rg_knn must not disturb installed covariances
  RG_CUDA(c, keep_cov.ensure(c->device, (size_t)cl.n_pad * COV_PLANES));
  RG_CUDA(c, cudaMemcpyAsync(keep_cov.p, cl.cov6.p, sizeof(float) * (size_t)cl.n_pad * COV_PLANES, cudaMemcpyDeviceToDevice, c->stream));
  c->k_correspondences = k;
  std::vector<Cloud*> one{&cl};
  const int rc = calc_cov_group(c, one, c->knn, c->stream, dk.p);
  c->k_correspondences = saved_k;
  cudaMemcpyAsync(cl.cov6.p, keep_cov.p, sizeof(float) * (size_t)cl.n_pad * COV_PLANES, cudaMemcpyDeviceToDevice, c->stream);
  cl.has_cov = had_cov;
  if (rc) { cudaStreamSynchronize(c->stream); dk.release(); keep_cov.release(); return rc; }
  std::vector<float4> hp((size_t)n);
  RG_CUDA(c, cudaMemcpyAsync(out, dk.p, sizeof(int) * (size_t)n * k, cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaMemcpyAsync(hp.data(), cl.pts.p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  RG_CUDA(c, cudaStreamSynchronize(c->stream));
  dk.release(); keep_cov.release();
  std::vector<std::pair<float, int>> row((size_t)k);
  for (int i = 0; i < n; i++) {
    int32_t* r = out + (size_t)i * k;
    for (int t = 0; t < k; t++) {
      const int id = r[t];
      float d = std::numeric_limits<float>::infinity();
      if (id >= 0 && id < n) {
        volatile float dx = hp[i].x - hp[id].x, dy = hp[i].y - hp[id].y, dz = hp[i].z - hp[id].z;
        volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;   // every product and sum rounded separately (no contraction)
        volatile float s1 = xx + yy;
        d = s1 + zz;
      }
      row[t] = {d, id};
    }
    std::sort(row.begin(), row.end());
    for (int t = 0; t < k; t++) r[t] = row[t].second;
  }
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

int rg_set_target_shard(rg_ctx* c, int rank, int world) {
  CHECK_CTX(c);
  if (world <= 1) { c->tgt_rank = 0; c->tgt_world = 0; }
  else {
    if (rank < 0 || rank >= world) return set_err(c, RG_ERR_INVALID, "rg_set_target_shard: rank out of range");
    c->tgt_rank = rank; c->tgt_world = world;
  }
  c->tgt.has_cov = false; c->voxel_ok = false; c->st_ok = false; c->have_corr = false;
  return RG_OK;
}

int rg_set_voxel_merge_hook(rg_ctx* c, rg_voxel_merge_fn fn, void* user) {
  CHECK_CTX(c);
  c->voxel_merge_fn = fn; c->voxel_merge_user = user;
  return RG_OK;
}

int rg_voxel_import(rg_ctx* c, const void* keys_dev, const void* sums_dev, int n_total) {
  CHECK_CTX(c);
  if (use_device(c)) return RG_ERR_CUDA;
  if (n_total < 0 || (n_total > 0 && (!keys_dev || !sums_dev))) return set_err(c, RG_ERR_INVALID, "rg_voxel_import: bad arguments");
  if (n_total > c->tgt.n) return set_err(c, RG_ERR_INVALID, "rg_voxel_import: more voxel entries than target points");
  VoxelBuild b;
  if (int rc = prepare_voxel_build(c, b)) return rc;
  const double* sums_in = (const double*)sums_dev;
  if (n_total > 0 && sums_in >= c->sums.p && sums_in < c->sums.p + c->sums.cap) {   // the caller passed our own accumulator rows: they are rewritten below
    RG_CUDA(c, c->staging.ensure(c->device, (size_t)n_total * 10 * sizeof(double)));
    RG_CUDA(c, cudaMemcpyAsync(c->staging.p, sums_in, (size_t)n_total * 10 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    sums_in = (const double*)c->staging.p;
  }
  RG_CUDA(c, launch_voxel_import(b, (const unsigned long long*)keys_dev, sums_in, n_total, c->stream));
  g_launches += 4;
  c->voxel_ok = true; c->n_vox = -1; c->st_ok = false; c->have_corr = false;
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
  const int n_setup = 1;   // ONE upload stream: the copy engine then serves the pairs strictly in issue order, so the first covariance group's inputs land first
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
  b->st_dev.release(); b->dyn_dev.release(); b->st_sub.release(); b->dyn_sub.release(); b->sched_dev.release(); b->knn.release(); b->vox_builds.release();
  if (b->dyn_host) pool().pinned_free(b->dyn_host);
  if (b->result_host) pool().pinned_free(b->result_host);
  if (b->flag_host) pool().pinned_free(b->flag_host);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->ev_inputs) cudaEventDestroy(b->ev_inputs);
  for (auto e : b->ev_stage) if (e) cudaEventDestroy(e);
  for (auto e : b->ev_knn) cudaEventDestroy(e);
  for (auto ev : b->setup_events) cudaEventDestroy(ev);
  for (auto st : b->setup_streams) { cudaStreamSynchronize(st); pool().stream_put(b->device, st); }
  delete b;
}

rg_ctx* rg_batch_ctx(rg_batch* b, int pair) { return (b && pair >= 0 && pair < (int)b->pairs.size()) ? b->pairs[pair] : nullptr; }
double rg_batch_last_linearize_us(const rg_batch* b) { return (b && b->lin_launches) ? b->lin_us_sum / b->lin_launches : -1.0; }
int rg_batch_last_linearize_launches(const rg_batch* b) { return b ? b->lin_launches : 0; }
int rg_batch_last_stages(const rg_batch* b, double out[8]) { if (!b || !out) return RG_ERR_INVALID; for (int i = 0; i < 8; i++) out[i] = b->stage_ms[i]; return RG_OK; }
void* rg_batch_stream(const rg_batch* b) { return b ? (void*)b->stream : nullptr; }
void* rg_stream(const rg_ctx* c) { return c ? (void*)c->stream : nullptr; }
unsigned long long rg_kernel_launch_count(void) { return g_launches.load(); }

int rg_batch_align(rg_batch* b, const float* guesses, float* out_T) {
  if (!b || !out_T) return set_err(nullptr, RG_ERR_INVALID, "rg_batch_align: bad arguments");
  const int P = (int)b->pairs.size();
  if (cudaSetDevice(b->device) != cudaSuccess) return set_err(nullptr, RG_ERR_CUDA, "cudaSetDevice failed");
  // stage 1: covariances of every cloud that lacks them, a GROUP of pairs per launch chain (one radix sort, one kNN + covariance
  // launch for all clouds of the group). A group only waits for the uploads of its own pairs, so the H2D copies of later
  // groups overlap the kernels of earlier ones.
  int noff = -1;
  for (int i = 0; i < P; i++) {
    rg_ctx* c = b->pairs[i];
    if (c->src.n == 0 || c->tgt.n == 0) return set_err(c, RG_ERR_INVALID, "batch align: source/target cloud not set");
    if (int rc = check_grid(c)) return rc;
    c->voxel_ok = false; c->st_ok = false; c->have_corr = false;
  }
  const int GROUP_PAIRS = 32;
  for (auto& e : b->ev_stage) if (!e) cudaEventCreate(&e);
  b->knn_groups = 0; b->knn_queries = 0;
  cudaEventRecord(b->ev_stage[0], b->stream);
  for (int g0 = 0; g0 < P; g0 += GROUP_PAIRS) {
    const int g1 = std::min(P, g0 + GROUP_PAIRS);
    std::vector<char> done(g1 - g0, 0);
    for (int i = g0; i < g1; i++) {
      rg_ctx* c = b->pairs[i];
      if (c->side_pending && c->upload_stream && c->upload_stream != b->stream) { RG_CUDA(c, cudaStreamWaitEvent(b->stream, c->ev_upload, 0)); c->side_pending = false; }
    }
    for (int i = g0; i < g1; i++) {
      if (done[i - g0]) continue;
      rg_ctx* lead = b->pairs[i];
      std::vector<Cloud*> todo;
      for (int j = i; j < g1; j++) {   // pairs of this group with the same (k, regulariser) share a launch chain
        rg_ctx* c = b->pairs[j];
        if (done[j - g0] || c->k_correspondences != lead->k_correspondences || c->regularization_method != lead->regularization_method) continue;
        done[j - g0] = 1;
        if (!c->tgt.has_cov) todo.push_back(&c->tgt);
        if (!c->src.has_cov) todo.push_back(&c->src);
      }
      if (todo.empty()) continue;
      while ((int)b->ev_knn.size() < 2 * (b->knn_groups + 1)) { cudaEvent_t e = nullptr; cudaEventCreate(&e); b->ev_knn.push_back(e); }
      if (int rc = calc_cov_group(lead, todo, b->knn, b->stream, nullptr, b->ev_knn[2 * b->knn_groups], b->ev_knn[2 * b->knn_groups + 1])) return rc;
      if (lead->k_correspondences <= 32) b->knn_groups++;
      for (Cloud* cl : todo) b->knn_queries += cl->n;
    }
  }
  cudaEventRecord(b->ev_stage[1], b->stream);
  // stage 2: every voxel map by one launch chain
  {
    std::vector<VoxelBuild> builds(P);
    for (int i = 0; i < P; i++) if (int rc = prepare_voxel_build(b->pairs[i], builds[i])) return rc;
    RG_CUDA(nullptr, b->vox_builds.ensure(b->device, P));
    RG_CUDA(nullptr, launch_voxel_build_batch(builds.data(), b->vox_builds.p, P, b->stream));
    g_launches += 4;
    for (int i = 0; i < P; i++) { rg_ctx* c = b->pairs[i]; c->voxel_ok = true; c->n_vox = -1; c->st_ok = false; c->have_corr = false; }
  }
  cudaEventRecord(b->ev_stage[2], b->stream);
  std::vector<LinStatic> st_host(P);
  for (int i = 0; i < P; i++) {
    rg_ctx* c = b->pairs[i];
    if (int rc = prepare_linearize(c, &st_host[i])) return rc;
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
  RG_CUDA(nullptr, cudaMemcpyAsync(b->st_dev.p, st_host.data(), sizeof(LinStatic) * P, cudaMemcpyHostToDevice, b->stream));   // pageable: staged before the call returns
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
  cudaEventRecord(b->ev_stage[3], b->stream);
  if (cudaEventSynchronize(b->ev_stage[3]) == cudaSuccess) {
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev_stage[0], b->ev_stage[1]); b->stage_ms[0] = ms;
    double knn = 0; for (int g = 0; g < b->knn_groups; g++) { cudaEventElapsedTime(&ms, b->ev_knn[2 * g], b->ev_knn[2 * g + 1]); knn += ms; }
    b->stage_ms[1] = knn;
    cudaEventElapsedTime(&ms, b->ev_stage[1], b->ev_stage[2]); b->stage_ms[2] = ms;
    cudaEventElapsedTime(&ms, b->ev_stage[2], b->ev_stage[3]); b->stage_ms[3] = ms;
    b->stage_ms[4] = b->lin_us_sum * 1e-3; b->stage_ms[5] = b->knn_queries; b->stage_ms[6] = b->lin_launches;
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

}  // extern "C"
