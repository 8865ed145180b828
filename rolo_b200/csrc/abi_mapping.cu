// abi_mapping.cu — C-ABI of the back-end (mapping) paths of SURVEY.md §8(f): scan-to-map optimisation, Scan Context place
// recognition, loop-closure fitness score (include/rolo_gicp.h). Kernels: scan2map.cu, scancontext.cu, covariance.cu.

#include "context.hpp"

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
