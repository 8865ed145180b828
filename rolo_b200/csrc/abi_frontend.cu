// abi_frontend.cu — C-ABI of the upstream range-image passes (imageProjection, featureExtraction), the device-resident front-end
// frame and the rolo/CloudInfoStamp wire image at the ROS edge (include/rolo_gicp.h). Kernels: frontend.cu.

#include "context.hpp"

extern "C" {

// ------------------------------------------------------------------------------------------------
// upstream range-image passes
// ------------------------------------------------------------------------------------------------
int rg_set_deskew(rg_ctx* c, const rg_deskew_params* p) {
  CHECK_CTX(c);
  if (!p) { c->fe_deskew.enabled = 0; return RG_OK; }
  if (p->time_offset_bytes < 0 || p->time_offset_bytes % 4 || !(p->scan_period > 0.f) || p->odom_time_diff == 0.0)
    return set_err(c, RG_ERR_INVALID, "rg_set_deskew: bad parameters");
  c->fe_deskew.enabled = 1; c->fe_deskew.time_offset = p->time_offset_bytes;
  c->fe_deskew.incre_rpy[0] = p->odom_incre_roll; c->fe_deskew.incre_rpy[1] = p->odom_incre_pitch; c->fe_deskew.incre_rpy[2] = p->odom_incre_yaw;
  c->fe_deskew.scan_period = p->scan_period; c->fe_deskew.odom_time_diff = p->odom_time_diff;
  return RG_OK;
}

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
  if (c->fe_deskew.enabled && c->fe_deskew.time_offset + 4 > p->point_stride_bytes) return set_err(c, RG_ERR_INVALID, "deskew: time offset outside the point record");
  if (int rc = fe_alloc(c, p->n_scan, p->horizon_scan)) return rc;
  const size_t cells = (size_t)p->n_scan * p->horizon_scan;
  RG_CUDA(c, c->fe_raw.ensure(c->device, std::max<size_t>(n_points * p->point_stride_bytes, 16)));
  if (n_points) RG_CUDA(c, cudaMemcpyAsync(c->fe_raw.p, raw, n_points * p->point_stride_bytes, cudaMemcpyHostToDevice, c->stream));
  ProjLaunch L;
  L.raw = c->fe_raw.p; L.n_points = (int)n_points; L.stride = p->point_stride_bytes; L.ring_offset = p->ring_offset_bytes;
  L.n_scan = p->n_scan; L.horizon = p->horizon_scan; L.downsample = p->downsample_rate; L.min_range = p->lidar_min_range; L.max_range = p->lidar_max_range; L.deskew = c->fe_deskew;
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
  if (c->fe_deskew.enabled && c->fe_deskew.time_offset + 4 > pp->point_stride_bytes) return set_err(c, RG_ERR_INVALID, "deskew: time offset outside the point record");
  if (int rc = fe_alloc(c, pp->n_scan, pp->horizon_scan)) return rc;
  const int cells = pp->n_scan * pp->horizon_scan;
  RG_CUDA(c, c->fe_raw.ensure(c->device, std::max<size_t>(n_points * pp->point_stride_bytes, 16)));
  if (n_points) RG_CUDA(c, cudaMemcpyAsync(c->fe_raw.p, raw, n_points * pp->point_stride_bytes, cudaMemcpyHostToDevice, c->stream));
  RG_CUDA(c, cudaMemsetAsync(c->fe_counts.p, 0, 8 * sizeof(int), c->stream));
  ProjLaunch P;
  P.raw = c->fe_raw.p; P.n_points = (int)n_points; P.stride = pp->point_stride_bytes; P.ring_offset = pp->ring_offset_bytes;
  P.n_scan = pp->n_scan; P.horizon = pp->horizon_scan; P.downsample = pp->downsample_rate; P.min_range = pp->lidar_min_range; P.max_range = pp->lidar_max_range; P.deskew = c->fe_deskew;
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

