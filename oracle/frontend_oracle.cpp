// frontend_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code) for the two upstream
// range-image passes of sdwyc/ROLO @ 3920bb7:
//   src/imageProjection.cpp   projectPointCloud :399-474, cloudExtraction :477-505 (deskew disabled by default,
//                             config/params.yaml:25; deskewPoint :368-396 returns the point unchanged then)
//   src/featureExtraction.cpp calculateSmoothness :87-110, markOccludedPoints :112-150,
//                             extractFeatures :153-266
// plus pcl::VoxelGrid<PointXYZI>::applyFilter (PCL >= 1.10, not in the tree; call site
// featureExtraction.cpp:256-258) restated from its published algorithm: leaf index
// floor(p * inverse_leaf) - min_b in float, centroid of all fields per leaf, output in ascending
// leaf-index order.
//
// PINNED against the reference's own code: no golden data exists for these passes, but the two node sources compile
// from where they lie into oracle/_ref/librolo_ref_fe.so (oracle/ref_fe_driver.cpp against oracle/ref_shim_ros), and
// tests/test_frontend_vs_ref_cpu.py requires this restatement to equal them: range image, compaction, index arrays, curvature,
// marks, labels and corner clouds exactly (labels / marks up to the sort ties of Q7), deskewed coordinates to one float ulp
// (glibc sinf / cosf vs correctly rounded), surface centroids to 1e-6 (summation order inside a leaf). What stays third-party
// and restated on both sides: pcl::VoxelGrid, pcl::getTransformation.
// Single-frame, stateless restatement. Reference quirks handled explicitly:
//  Q7  std::sort tie order is unspecified -> ties broken by point index here (and on the GPU).
//  Q8  the reference's work arrays are `new`-allocated once and never cleared: entries outside
//      [5, n-5) keep stale values, cloudSmoothness[4] is never written although ring 0 / sector 0
//      scans k = 4, and neighbour marking can index -1 / n. Restated as: all work arrays start at
//      zero each frame, cloudSmoothness[i].ind = i for every i, and neighbour marking stops at the
//      array ends.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

// deskewPoint (imageProjection.cpp:368-396) for deskewEnabled && odomAvailable: rotation-only correction by the odometry
// increment scaled to the point's time; sin/cos are the correctly rounded float values (double libm rounded once)
static void deskew_point(float& x, float& y, float& z, double relTime, const float incre_rpy[3], float scanPeriod, double odomTimeDiff) {
  const float ratio = relTime / scanPeriod;                                       // :380
  const float k = scanPeriod / odomTimeDiff;                                      // Eigen takes the scalar as float (:383)
  const float roll = -((incre_rpy[0] * k) * ratio), pitch = -((incre_rpy[1] * k) * ratio), yaw = -((incre_rpy[2] * k) * ratio);
  auto cs = [](float a) { return (float)std::cos((double)a); };
  auto sn = [](float a) { return (float)std::sin((double)a); };
  const float A = cs(yaw), B = sn(yaw), C = cs(pitch), D = sn(pitch), E = cs(roll), F = sn(roll), DE = D * E, DF = D * F;   // pcl::getTransformation
  const float t00 = A * C, t01 = A * DF - B * E, t02 = B * F + A * DE, t10 = B * C, t11 = A * E + B * DF, t12 = B * DE - A * F, t20 = -D, t21 = C * F, t22 = C * E;
  const float nx = t00 * x + t01 * y + t02 * z + 0.0f, ny = t10 * x + t11 * y + t12 * z + 0.0f, nz = t20 * x + t21 * y + t22 * z + 0.0f;   // :388-390
  x = nx; y = ny; z = nz;
}

int ro_image_projection_deskew(const void* raw, size_t n_points, int stride_bytes, int ring_offset_bytes, int n_scan, int horizon_scan,
                               int downsample_rate, float lidar_min_range, float lidar_max_range, int deskew, int time_offset_bytes,
                               const float* incre_rpy, float scan_period, double odom_time_diff, float* range_image,
                               float* extracted_xyzi, int* point_col_ind, float* point_range, int* start_ring_index, int* end_ring_index);

// returns the number of extracted points
int ro_image_projection(const void* raw, size_t n_points, int stride_bytes, int ring_offset_bytes, int n_scan, int horizon_scan,
                        int downsample_rate, float lidar_min_range, float lidar_max_range, float* range_image,
                        float* extracted_xyzi, int* point_col_ind, float* point_range, int* start_ring_index, int* end_ring_index) {
  return ro_image_projection_deskew(raw, n_points, stride_bytes, ring_offset_bytes, n_scan, horizon_scan, downsample_rate, lidar_min_range,
                                    lidar_max_range, 0, 0, nullptr, 0.1f, -1.0, range_image, extracted_xyzi, point_col_ind, point_range,
                                    start_ring_index, end_ring_index);
}

int ro_image_projection_deskew(const void* raw, size_t n_points, int stride_bytes, int ring_offset_bytes, int n_scan, int horizon_scan,
                               int downsample_rate, float lidar_min_range, float lidar_max_range, int deskew, int time_offset_bytes,
                               const float* incre_rpy, float scan_period, double odom_time_diff, float* range_image,
                               float* extracted_xyzi, int* point_col_ind, float* point_range, int* start_ring_index, int* end_ring_index) {
  const unsigned char* base = (const unsigned char*)raw;
  const size_t cells = (size_t)n_scan * horizon_scan;
  std::vector<float> full(cells * 4, 0.0f);
  for (size_t i = 0; i < cells; i++) range_image[i] = FLT_MAX;                    // resetParameters
  // imageProjection.cpp:399-474
  for (size_t i = 0; i < n_points; i++) {
    const float* p = (const float*)(base + i * stride_bytes);
    uint16_t ring; std::memcpy(&ring, base + i * stride_bytes + ring_offset_bytes, sizeof(ring));
    const float x = p[0], y = p[1], z = p[2];
    const float intensity = ring * z;                                             // :410
    const float range = std::sqrt(x * x + y * y + z * z);                         // utility.h:462-465
    if (range < lidar_min_range || range > lidar_max_range) continue;             // :414
    const int rowIdn = ring;
    if (rowIdn < 0 || rowIdn >= n_scan) continue;                                 // :428
    if (rowIdn % downsample_rate != 0) continue;                                  // :431
    int columnIdn = -1;
    {
      float horizonAngle = std::atan2(x, y) * 180 / M_PI;                         // :438 (float atan2, double division)
      static thread_local float ang_res_x; ang_res_x = 360.0 / float(horizon_scan);  // :439
      columnIdn = -round((horizonAngle - 90.0) / ang_res_x) + horizon_scan / 2;   // :441
      if (columnIdn >= horizon_scan) columnIdn -= horizon_scan;                   // :443-444
    }
    if (columnIdn < 0 || columnIdn >= horizon_scan) continue;                     // :447
    const size_t index = (size_t)columnIdn + (size_t)rowIdn * horizon_scan;
    if (range_image[index] != FLT_MAX) continue;                                  // :451 first writer wins
    range_image[index] = range;
    float dx = x, dy = y, dz = z;
    if (deskew) {                                                                 // :454 thisPoint = deskewPoint(&thisPoint, relTime)
      float rel; std::memcpy(&rel, base + i * stride_bytes + time_offset_bytes, sizeof(rel));
      deskew_point(dx, dy, dz, (double)rel, incre_rpy, scan_period, odom_time_diff);
    }
    full[index * 4 + 0] = dx; full[index * 4 + 1] = dy; full[index * 4 + 2] = dz; full[index * 4 + 3] = intensity;
  }
  // imageProjection.cpp:477-505
  int count = 0;
  for (int i = 0; i < n_scan; ++i) {
    start_ring_index[i] = count - 1 + 5;
    for (int j = 0; j < horizon_scan; ++j) {
      const size_t idx = (size_t)j + (size_t)i * horizon_scan;
      if (range_image[idx] != FLT_MAX) {
        point_col_ind[count] = j;
        point_range[count] = range_image[idx];
        std::memcpy(extracted_xyzi + (size_t)count * 4, &full[idx * 4], 4 * sizeof(float));
        ++count;
      }
    }
    end_ring_index[i] = count - 1 - 5;
  }
  return count;
}

// pcl::VoxelGrid<PointXYZI> (downsample_all_data = true, min_points_per_voxel = 0), appends to out; returns #out
static int voxel_grid(const std::vector<float>& in /*xyzi*/, float leaf, float* out) {
  const size_t n = in.size() / 4;
  if (n == 0) return 0;
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (size_t i = 0; i < n; i++) for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], in[i * 4 + d]); mx[d] = std::max(mx[d], in[i * 4 + d]); }
  int min_b[3], max_b[3], div_b[3];
  for (int d = 0; d < 3; d++) { min_b[d] = (int)std::floor(mn[d] * inv); max_b[d] = (int)std::floor(mx[d] * inv); div_b[d] = max_b[d] - min_b[d] + 1; }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<int, int>> idx(n);
  for (size_t i = 0; i < n; i++) {
    int ijk[3];
    for (int d = 0; d < 3; d++) ijk[d] = (int)(std::floor(in[i * 4 + d] * inv) - (float)min_b[d]);
    idx[i] = {ijk[0] * mul[0] + ijk[1] * mul[1] + ijk[2] * mul[2], (int)i};
  }
  std::sort(idx.begin(), idx.end());    // (leaf index, point index): deterministic centroid summation order
  int m = 0;
  size_t s = 0;
  while (s < n) {
    size_t e = s;
    float c[4] = {0, 0, 0, 0};
    while (e < n && idx[e].first == idx[s].first) { for (int d = 0; d < 4; d++) c[d] += in[(size_t)idx[e].second * 4 + d]; e++; }
    const float cnt = (float)(e - s);
    for (int d = 0; d < 4; d++) out[(size_t)m * 4 + d] = c[d] / cnt;
    m++;
    s = e;
  }
  return m;
}

int ro_feature_extraction(const float* extracted_xyzi, const int* point_col_ind, const float* point_range, const int* start_ring_index,
                          const int* end_ring_index, int n, int n_scan, int /*horizon_scan*/, float edge_threshold, float surf_threshold,
                          float surf_leaf_size, float* curvature, int* neighbor_picked, int* label, float* corner_xyzi, int* n_corner,
                          float* surface_xyzi, int* n_surface) {
  struct Smooth { float value; int ind; };
  std::vector<Smooth> sm(n);
  for (int i = 0; i < n; i++) { curvature[i] = 0.f; neighbor_picked[i] = 0; label[i] = 0; sm[i] = {0.f, i}; }   // Q8
  // featureExtraction.cpp:87-110
  for (int i = 5; i < n - 5; i++) {
    float diffRange = point_range[i - 5] + point_range[i - 4] + point_range[i - 3] + point_range[i - 2] + point_range[i - 1] -
                      point_range[i] * 10 + point_range[i + 1] + point_range[i + 2] + point_range[i + 3] + point_range[i + 4] +
                      point_range[i + 5];
    curvature[i] = diffRange * diffRange;
    sm[i].value = curvature[i];
  }
  // featureExtraction.cpp:112-150
  for (int i = 5; i < n - 6; ++i) {
    float depth1 = point_range[i], depth2 = point_range[i + 1];
    int columnDiff = std::abs(int(point_col_ind[i + 1] - point_col_ind[i]));
    if (columnDiff < 10) {
      if (depth1 - depth2 > 0.3) { for (int l = -5; l <= 0; l++) neighbor_picked[i + l] = 1; }
      else if (depth2 - depth1 > 0.3) { for (int l = 1; l <= 6; l++) neighbor_picked[i + l] = 1; }
    }
    float diff1 = std::abs(float(point_range[i - 1] - point_range[i]));
    float diff2 = std::abs(float(point_range[i + 1] - point_range[i]));
    if (diff1 > 0.02 * point_range[i] && diff2 > 0.02 * point_range[i]) neighbor_picked[i] = 1;
  }
  // featureExtraction.cpp:153-266
  auto mark_neighbors = [&](int ind) {
    for (int l = 1; l <= 5; l++) {
      if (ind + l >= n) break;                                                                           // Q8
      int columnDiff = std::abs(int(point_col_ind[ind + l] - point_col_ind[ind + l - 1]));
      if (columnDiff > 10) break;
      neighbor_picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; l--) {
      if (ind + l < 0) break;                                                                            // Q8
      int columnDiff = std::abs(int(point_col_ind[ind + l] - point_col_ind[ind + l + 1]));
      if (columnDiff > 10) break;
      neighbor_picked[ind + l] = 1;
    }
  };
  int nc = 0, ns = 0;
  std::vector<float> surf_scan;
  for (int i = 0; i < n_scan; i++) {
    surf_scan.clear();
    for (int j = 0; j < 6; j++) {
      int sp = (start_ring_index[i] * (6 - j) + end_ring_index[i] * j) / 6;
      int ep = (start_ring_index[i] * (5 - j) + end_ring_index[i] * (j + 1)) / 6 - 1;
      if (sp >= ep) continue;
      std::sort(sm.begin() + sp, sm.begin() + ep, [](const Smooth& a, const Smooth& b) { return a.value < b.value || (a.value == b.value && a.ind < b.ind); });   // :179 + Q7
      int largestPickedNum = 0;
      for (int k = ep; k >= sp; k--) {
        int ind = sm[k].ind;
        if (neighbor_picked[ind] == 0 && curvature[ind] > edge_threshold) {
          largestPickedNum++;
          if (largestPickedNum <= 20) {
            label[ind] = 1;
            std::memcpy(corner_xyzi + (size_t)nc * 4, extracted_xyzi + (size_t)ind * 4, 4 * sizeof(float));
            nc++;
          } else {
            break;
          }
          neighbor_picked[ind] = 1;
          mark_neighbors(ind);
        }
      }
      for (int k = sp; k <= ep; k++) {
        int ind = sm[k].ind;
        if (neighbor_picked[ind] == 0 && curvature[ind] < surf_threshold) {
          label[ind] = -1;
          neighbor_picked[ind] = 1;
          mark_neighbors(ind);
        }
      }
      for (int k = sp; k <= ep; k++)
        if (label[k] <= 0) surf_scan.insert(surf_scan.end(), extracted_xyzi + (size_t)k * 4, extracted_xyzi + (size_t)k * 4 + 4);
    }
    ns += voxel_grid(surf_scan, surf_leaf_size, surface_xyzi + (size_t)ns * 4);   // :256-258, :263
  }
  *n_corner = nc; *n_surface = ns;
  return 0;
}

}  // extern "C"
