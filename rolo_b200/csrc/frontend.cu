// frontend.cu — the two upstream range-image passes as sm_100a kernels.
//
//  ImageProjection::projectPointCloud + cloudExtraction   (src/imageProjection.cpp:399-505, optional rotation-only deskewPoint :368-396)
//  FeatureExtraction::calculateSmoothness / markOccludedPoints / extractFeatures
//                                                          (src/featureExtraction.cpp:87-266)
//  pcl::VoxelGrid<PointXYZI> per ring                      (call site featureExtraction.cpp:256-258)
//
// These are HBM-bound integer/float passes over a 128 x 2048 image (no dense contraction):
//  * "first point to claim a cell wins" (imageProjection.cpp:451) = atomicMin over the point index;
//  * row-major compaction = per-row counts + an intra-row ballot scan (one CTA per ring);
//  * per ring the six sectors are coupled through cloudNeighborPicked marks that cross sector borders, so
//    one CTA owns a ring: each sector is bitonic-sorted in shared memory by (curvature, index) and the greedy
//    corner / surface selection is run by one warp, 32 sorted candidates at a time: ballot the still-eligible
//    lanes, take the first, mark its +-5 neighbours (with the column-gap break), re-test. Rings never
//    interact (their selectable ranges are 9 indices apart, marks reach 5), so rings run in parallel;
//  * the per-ring VoxelGrid is a shared-memory sort by leaf index + a segmented centroid.
// Float expressions use explicit _rn intrinsics wherever nvcc's FMA contraction would change a rounding
// the CPU reference performs (curvature sum, range, centroid sums), so results match bit for bit.

#include "frontend.cuh"

#include <float.h>
#include <limits.h>

namespace rolo {

namespace {

struct RawPoint { float x, y, z; int ring; };

__device__ __forceinline__ RawPoint load_raw(const unsigned char* raw, int i, int stride, int ring_offset) {
  const unsigned char* p = raw + (size_t)i * stride;
  RawPoint r;
  r.x = *reinterpret_cast<const float*>(p);
  r.y = *reinterpret_cast<const float*>(p + 4);
  r.z = *reinterpret_cast<const float*>(p + 8);
  r.ring = (int)*reinterpret_cast<const unsigned short*>(p + ring_offset);
  return r;
}
__device__ __forceinline__ float point_range(float x, float y, float z) {   // utility.h:462-465
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}
// imageProjection.cpp:438-444 with the reference's float/double promotions
__device__ __forceinline__ int column_index(float x, float y, int horizon) {
  const float horizonAngle = (float)((double)__fmul_rn(atan2f(x, y), 180.0f) / 3.14159265358979323846);
  const float ang_res_x = (float)(360.0 / (double)(float)horizon);
  int col = (int)(-round(((double)horizonAngle - 90.0) / (double)ang_res_x) + (double)(horizon / 2));
  if (col >= horizon) col -= horizon;
  return col;
}

__global__ void proj_init_kernel(int* owner, int cells) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cells) owner[c] = INT_MAX;
}

__global__ void proj_claim_kernel(const unsigned char* __restrict__ raw, int n, int stride, int ring_offset, int n_scan, int horizon,
                                  int downsample, float min_range, float max_range, int* owner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RawPoint p = load_raw(raw, i, stride, ring_offset);
  const float range = point_range(p.x, p.y, p.z);
  if (range < min_range || range > max_range) return;
  const int row = p.ring;
  if (row < 0 || row >= n_scan) return;
  if (row % downsample != 0) return;
  const int col = column_index(p.x, p.y, horizon);
  if (col < 0 || col >= horizon) return;
  atomicMin(&owner[(size_t)row * horizon + col], i);      // first writer wins == smallest point index
}

// deskewPoint (imageProjection.cpp:368-396): rotate the point back by the odometry increment scaled to its time. sin / cos are
// the correctly rounded float values (double libm rounded once), identical to the oracle's.
__device__ __forceinline__ void deskew_point(float& x, float& y, float& z, float rel, const ProjDeskew& dk) {
  const float ratio = (float)((double)rel / (double)dk.scan_period);              // :380 (relTime is a double parameter)
  const float k = (float)((double)dk.scan_period / dk.odom_time_diff);            // Eigen takes the scalar as float (:383)
  const float roll = -__fmul_rn(__fmul_rn(dk.incre_rpy[0], k), ratio), pitch = -__fmul_rn(__fmul_rn(dk.incre_rpy[1], k), ratio),
              yaw = -__fmul_rn(__fmul_rn(dk.incre_rpy[2], k), ratio);
  const float A = (float)cos((double)yaw), B = (float)sin((double)yaw), C = (float)cos((double)pitch), D = (float)sin((double)pitch),
              E = (float)cos((double)roll), F = (float)sin((double)roll);
  const float DE = __fmul_rn(D, E), DF = __fmul_rn(D, F);                         // pcl::getTransformation(0,0,0,roll,pitch,yaw)
  const float t00 = __fmul_rn(A, C), t01 = __fsub_rn(__fmul_rn(A, DF), __fmul_rn(B, E)), t02 = __fadd_rn(__fmul_rn(B, F), __fmul_rn(A, DE));
  const float t10 = __fmul_rn(B, C), t11 = __fadd_rn(__fmul_rn(A, E), __fmul_rn(B, DF)), t12 = __fsub_rn(__fmul_rn(B, DE), __fmul_rn(A, F));
  const float t20 = -D, t21 = __fmul_rn(C, F), t22 = __fmul_rn(C, E);
  const float nx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00, x), __fmul_rn(t01, y)), __fmul_rn(t02, z)), 0.0f);
  const float ny = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t10, x), __fmul_rn(t11, y)), __fmul_rn(t12, z)), 0.0f);
  const float nz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t20, x), __fmul_rn(t21, y)), __fmul_rn(t22, z)), 0.0f);
  x = nx; y = ny; z = nz;
}

__global__ void proj_write_kernel(const unsigned char* __restrict__ raw, int stride, int ring_offset, const int* __restrict__ owner, int cells,
                                  ProjDeskew dk, float* __restrict__ range_image, float4* __restrict__ full) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const int o = owner[c];
  if (o == INT_MAX) { range_image[c] = FLT_MAX; full[c] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
  const RawPoint p = load_raw(raw, o, stride, ring_offset);
  range_image[c] = point_range(p.x, p.y, p.z);
  float x = p.x, y = p.y, z = p.z;
  if (dk.enabled) deskew_point(x, y, z, *reinterpret_cast<const float*>(raw + (size_t)o * stride + dk.time_offset), dk);   // :454
  full[c] = make_float4(x, y, z, __fmul_rn((float)p.ring, p.z));           // intensity := ring * z of the raw point (:410)
}

__global__ void proj_count_kernel(const float* __restrict__ range_image, int horizon, int* row_count) {
  const int row = blockIdx.x;
  int cnt = 0;
  for (int j = threadIdx.x; j < horizon; j += blockDim.x) cnt += range_image[(size_t)row * horizon + j] != FLT_MAX;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  __shared__ int s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s[w]; row_count[row] = t; }
}

__global__ void proj_extract_kernel(const float* __restrict__ range_image, const float4* __restrict__ full, const int* __restrict__ row_count,
                                    int n_scan, int horizon, float4* __restrict__ extracted, int* __restrict__ col_ind,
                                    float* __restrict__ prange, int* start_ring, int* end_ring, int* n_extracted) {
  const int row = blockIdx.x;
  __shared__ int s_off, s_warp[32], s_base;
  if (threadIdx.x == 0) {
    int off = 0;
    for (int r = 0; r < row; r++) off += row_count[r];
    s_off = off;
    start_ring[row] = off - 1 + 5;                           // imageProjection.cpp:485
    end_ring[row] = off + row_count[row] - 1 - 5;            // :503
    if (row == n_scan - 1) *n_extracted = off + row_count[row];
    s_base = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j0 = 0; j0 < horizon; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    const size_t idx = (size_t)row * horizon + j;
    const bool valid = j < horizon && range_image[idx] != FLT_MAX;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; w++) before += s_warp[w];
    if (valid) {
      const int pos = s_off + s_base + before + __popc(m & ((1u << lane) - 1u));
      col_ind[pos] = j;
      prange[pos] = range_image[idx];
      extracted[pos] = full[idx];
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < nw; w++) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
}

// ---- feature extraction -------------------------------------------------------------------------------------

__global__ void feat_init_kernel(const float* __restrict__ r, const int* __restrict__ n_dev, int n_max, float* curvature, int* picked, int* label) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_dev ? *n_dev : n_max;
  if (i >= n) return;
  float c = 0.f;
  if (i >= 5 && i < n - 5) {                                  // featureExtraction.cpp:91-99, summed left to right
    float d = __fadd_rn(r[i - 5], r[i - 4]);
    d = __fadd_rn(d, r[i - 3]); d = __fadd_rn(d, r[i - 2]); d = __fadd_rn(d, r[i - 1]);
    d = __fsub_rn(d, __fmul_rn(r[i], 10.0f));
    d = __fadd_rn(d, r[i + 1]); d = __fadd_rn(d, r[i + 2]); d = __fadd_rn(d, r[i + 3]); d = __fadd_rn(d, r[i + 4]); d = __fadd_rn(d, r[i + 5]);
    c = __fmul_rn(d, d);
  }
  curvature[i] = c; picked[i] = 0; label[i] = 0;
}

__global__ void feat_occlusion_kernel(const float* __restrict__ r, const int* __restrict__ col, const int* __restrict__ n_dev, int n_max, int* picked) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_dev ? *n_dev : n_max;
  if (i < 5 || i >= n - 6) return;                            // featureExtraction.cpp:116
  const float depth1 = r[i], depth2 = r[i + 1];
  const int columnDiff = abs(col[i + 1] - col[i]);
  if (columnDiff < 10) {
    if ((double)__fsub_rn(depth1, depth2) > 0.3) {
      for (int l = -5; l <= 0; l++) picked[i + l] = 1;
    } else if ((double)__fsub_rn(depth2, depth1) > 0.3) {
      for (int l = 1; l <= 6; l++) picked[i + l] = 1;
    }
  }
  const float diff1 = fabsf(__fsub_rn(r[i - 1], r[i]));
  const float diff2 = fabsf(__fsub_rn(r[i + 1], r[i]));
  if ((double)diff1 > 0.02 * (double)r[i] && (double)diff2 > 0.02 * (double)r[i]) picked[i] = 1;
}

__device__ __forceinline__ void block_bitonic_sort(unsigned long long* keys, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i + j;
        const bool up = (i & k) == 0;
        const unsigned long long a = keys[i], b = keys[ixj];
        if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
      }
      __syncthreads();
    }
  }
}

// marks the +-5 neighbours of `ind` (featureExtraction.cpp:199-212): warp-cooperative, lanes 0..4 forward, 8..12 backward
__device__ __forceinline__ void mark_neighbors(int ind, int n, const int* __restrict__ col, volatile int* picked, int lane) {
  bool ok = false;
  int target = -1;
  if (lane < 5) {
    const int l = lane + 1;
    if (ind + l < n) { ok = abs(col[ind + l] - col[ind + l - 1]) <= 10; target = ind + l; }
  } else if (lane >= 8 && lane < 13) {
    const int l = -(lane - 8 + 1);
    if (ind + l >= 0) { ok = abs(col[ind + l] - col[ind + l + 1]) <= 10; target = ind + l; }
  }
  const unsigned okm = __ballot_sync(0xffffffffu, ok);
  const unsigned fwd = okm & 0x1Fu, bwd = (okm >> 8) & 0x1Fu;
  const int nf = __ffs(~fwd & 0x3Fu) - 1;      // number of leading ok steps before the first break
  const int nb = __ffs(~bwd & 0x3Fu) - 1;
  if (lane < 5 && lane < nf) picked[target] = 1;
  if (lane >= 8 && lane < 13 && (lane - 8) < nb) picked[target] = 1;
  __syncwarp();
}

__global__ void __launch_bounds__(256)
feat_select_kernel(const float4* __restrict__ extracted, const int* __restrict__ col, const int* __restrict__ start_ring, const int* __restrict__ end_ring,
                   const int* __restrict__ n_dev, int n_max, float edge_thr, float surf_thr, const float* __restrict__ curvature, int* picked_g,
                   int* label, float4* __restrict__ corner_ring, int* __restrict__ corner_cnt, int* error_flag) {
  __shared__ unsigned long long keys[FE_SECTOR_CAP];
  __shared__ int s_corner;
  const int ring = blockIdx.x;
  const int n = n_dev ? *n_dev : n_max;
  const int lane = threadIdx.x & 31;
  volatile int* picked = picked_g;
  if (threadIdx.x == 0) s_corner = 0;
  __syncthreads();
  const int sr = start_ring[ring], er = end_ring[ring];
  for (int j = 0; j < 6; j++) {
    const int sp = (sr * (6 - j) + er * j) / 6;                       // featureExtraction.cpp:171-172
    const int ep = (sr * (5 - j) + er * (j + 1)) / 6 - 1;
    if (sp >= ep) continue;                                           // block-uniform
    const int len = ep - sp;
    if (len > FE_SECTOR_CAP || sp < 0 || ep >= n) { if (threadIdx.x == 0) *error_flag = 1; continue; }
    int P = 2; while (P < len) P <<= 1;
    for (int t = threadIdx.x; t < P; t += blockDim.x)
      keys[t] = (t < len) ? (((unsigned long long)__float_as_uint(curvature[sp + t]) << 32) | (unsigned)(sp + t)) : ~0ull;
    __syncthreads();
    block_bitonic_sort(keys, P);                                      // ascending (curvature, index): :179 + Q7
    if (threadIdx.x < 32) {
      // ---- corners: largest curvature first, k = ep .. sp (:182-213) ----
      int n_picked = 0;
      bool stop = false;
      for (int base = ep; base >= sp && !stop; base -= 32) {
        const int k = base - lane;
        const bool valid = k >= sp;
        const int ind = !valid ? 0 : (k == ep ? ep : (int)(unsigned)(keys[k - sp] & 0xFFFFFFFFull));
        bool cand = valid && curvature[ind] > edge_thr;
        for (;;) {
          const bool elig = cand && picked[ind] == 0;
          const unsigned m = __ballot_sync(0xffffffffu, elig);
          if (m == 0) break;
          const int sel = __ffs(m) - 1;
          n_picked++;
          if (n_picked > 20) { stop = true; break; }                  // the 21st eligible point ends the sector scan
          const int sind = __shfl_sync(0xffffffffu, ind, sel);
          if (lane == 0) {
            label[sind] = 1;
            corner_ring[ring * FE_MAX_CORNERS_PER_RING + s_corner] = extracted[sind];
            s_corner = s_corner + 1;
            picked[sind] = 1;
          }
          __syncwarp();
          mark_neighbors(sind, n, col, picked, lane);
          if (lane == sel) cand = false;
        }
      }
      // ---- surface points: smallest curvature first, k = sp .. ep (:214-239) ----
      for (int base = sp; base <= ep; base += 32) {
        const int k = base + lane;
        const bool valid = k <= ep;
        const int ind = !valid ? 0 : (k == ep ? ep : (int)(unsigned)(keys[k - sp] & 0xFFFFFFFFull));
        bool cand = valid && curvature[ind] < surf_thr;
        for (;;) {
          const bool elig = cand && picked[ind] == 0;
          const unsigned m = __ballot_sync(0xffffffffu, elig);
          if (m == 0) break;
          const int sel = __ffs(m) - 1;
          const int sind = __shfl_sync(0xffffffffu, ind, sel);
          if (lane == 0) { label[sind] = -1; picked[sind] = 1; }
          __syncwarp();
          mark_neighbors(sind, n, col, picked, lane);
          if (lane == sel) cand = false;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) corner_cnt[ring] = s_corner;
}

__global__ void __launch_bounds__(256)
feat_surface_kernel(const float4* __restrict__ extracted, const int* __restrict__ start_ring, const int* __restrict__ end_ring,
                    const int* __restrict__ label, float leaf, int horizon, float4* __restrict__ surf_ring, int* __restrict__ surf_cnt, int* error_flag) {
  __shared__ unsigned long long keys[FE_RING_CAP];
  __shared__ int s_list[FE_RING_CAP];
  __shared__ int s_cnt, s_warp[8], s_heads;
  __shared__ float s_mn[3][8], s_mx[3][8];
  __shared__ int s_minb[3], s_div[2];
  const int ring = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_cnt = 0; s_heads = 0; }
  __syncthreads();
  const int sr = start_ring[ring], er = end_ring[ring];
  // candidates: every k of every non-empty sector with label <= 0, in index order (featureExtraction.cpp:241-247)
  for (int j = 0; j < 6; j++) {
    const int sp = (sr * (6 - j) + er * j) / 6;
    const int ep = (sr * (5 - j) + er * (j + 1)) / 6 - 1;
    if (sp >= ep) continue;
    for (int k0 = sp; k0 <= ep; k0 += blockDim.x) {
      const int k = k0 + threadIdx.x;
      const bool take = k <= ep && label[k] <= 0;
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      int before = 0;
      for (int w = 0; w < warp; w++) before += s_warp[w];
      const int pos = s_cnt + before + __popc(m & ((1u << lane) - 1u));
      if (take && pos < FE_RING_CAP) s_list[pos] = k;
      __syncthreads();
      if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += s_warp[w]; s_cnt += t; }
      __syncthreads();
    }
  }
  int cnt = s_cnt;
  if (cnt > FE_RING_CAP) { if (threadIdx.x == 0) { *error_flag = 2; surf_cnt[ring] = 0; } return; }
  if (cnt == 0) { if (threadIdx.x == 0) surf_cnt[ring] = 0; return; }
  // pcl::VoxelGrid: bounding box -> min_b / div_b
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
    const float4 p = extracted[s_list[t]];
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  #pragma unroll
  for (int d = 0; d < 3; d++) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o)); }
    if (lane == 0) { s_mn[d][warp] = mn[d]; s_mx[d][warp] = mx[d]; }
  }
  __syncthreads();
  const float inv = __fdiv_rn(1.0f, leaf);
  if (threadIdx.x == 0) {
    int minb[3], maxb[3];
    for (int d = 0; d < 3; d++) {
      float a = s_mn[d][0], b = s_mx[d][0];
      for (int w = 1; w < 8; w++) { a = fminf(a, s_mn[d][w]); b = fmaxf(b, s_mx[d][w]); }
      minb[d] = (int)floorf(__fmul_rn(a, inv)); maxb[d] = (int)floorf(__fmul_rn(b, inv));
      s_minb[d] = minb[d];
    }
    s_div[0] = maxb[0] - minb[0] + 1;
    s_div[1] = (maxb[0] - minb[0] + 1) * (maxb[1] - minb[1] + 1);
  }
  __syncthreads();
  int P = 2; while (P < cnt) P <<= 1;
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    unsigned long long key = ~0ull;
    if (t < cnt) {
      const float4 p = extracted[s_list[t]];
      const int i0 = (int)(floorf(__fmul_rn(p.x, inv)) - (float)s_minb[0]);
      const int i1 = (int)(floorf(__fmul_rn(p.y, inv)) - (float)s_minb[1]);
      const int i2 = (int)(floorf(__fmul_rn(p.z, inv)) - (float)s_minb[2]);
      const int idx = i0 + i1 * s_div[0] + i2 * s_div[1];
      key = ((unsigned long long)(unsigned)idx << 32) | (unsigned)t;     // (leaf index, position in the candidate list)
    }
    keys[t] = key;
  }
  __syncthreads();
  block_bitonic_sort(keys, P);
  // segmented centroid: one thread per run head, float sums in sorted order, output ordered by leaf index
  for (int t0 = 0; t0 < cnt; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const bool head = t < cnt && (t == 0 || (keys[t] >> 32) != (keys[t - 1] >> 32));
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; w++) before += s_warp[w];
    if (head) {
      const int rank = s_heads + before + __popc(m & ((1u << lane) - 1u));
      const unsigned long long leafid = keys[t] >> 32;
      float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
      int e = t;
      while (e < cnt && (keys[e] >> 32) == leafid) {
        const float4 p = extracted[s_list[(int)(unsigned)(keys[e] & 0xFFFFFFFFull)]];
        cx = __fadd_rn(cx, p.x); cy = __fadd_rn(cy, p.y); cz = __fadd_rn(cz, p.z); ci = __fadd_rn(ci, p.w);
        e++;
      }
      const float c = (float)(e - t);
      surf_ring[(size_t)ring * horizon + rank] = make_float4(__fdiv_rn(cx, c), __fdiv_rn(cy, c), __fdiv_rn(cz, c), __fdiv_rn(ci, c));
    }
    __syncthreads();
    if (threadIdx.x == 0) { int tt = 0; for (int w = 0; w < 8; w++) tt += s_warp[w]; s_heads += tt; }
    __syncthreads();
  }
  if (threadIdx.x == 0) surf_cnt[ring] = s_heads;
}

__global__ void feat_compact_kernel(const float4* __restrict__ corner_ring, const int* __restrict__ corner_cnt, const float4* __restrict__ surf_ring,
                                    const int* __restrict__ surf_cnt, int n_scan, int horizon, float4* __restrict__ corner_out, int* n_corner,
                                    float4* __restrict__ surf_out, int* n_surf) {
  const int ring = blockIdx.x;
  __shared__ int s_co, s_so;
  if (threadIdx.x == 0) {
    int co = 0, so = 0;
    for (int r = 0; r < ring; r++) { co += corner_cnt[r]; so += surf_cnt[r]; }
    s_co = co; s_so = so;
    if (ring == n_scan - 1) { *n_corner = co + corner_cnt[ring]; *n_surf = so + surf_cnt[ring]; }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < corner_cnt[ring]; t += blockDim.x) corner_out[s_co + t] = corner_ring[ring * FE_MAX_CORNERS_PER_RING + t];
  for (int t = threadIdx.x; t < surf_cnt[ring]; t += blockDim.x) surf_out[s_so + t] = surf_ring[(size_t)ring * horizon + t];
}

}  // namespace

cudaError_t launch_projection(const ProjLaunch& L, cudaStream_t stream) {
  const int cells = L.n_scan * L.horizon;
  proj_init_kernel<<<(cells + 255) / 256, 256, 0, stream>>>(L.owner, cells);
  if (L.n_points > 0)
    proj_claim_kernel<<<(L.n_points + 255) / 256, 256, 0, stream>>>(L.raw, L.n_points, L.stride, L.ring_offset, L.n_scan, L.horizon, L.downsample,
                                                                   L.min_range, L.max_range, L.owner);
  proj_write_kernel<<<(cells + 255) / 256, 256, 0, stream>>>(L.raw, L.stride, L.ring_offset, L.owner, cells, L.deskew, L.range_image, L.full);
  proj_count_kernel<<<L.n_scan, 256, 0, stream>>>(L.range_image, L.horizon, L.row_count);
  proj_extract_kernel<<<L.n_scan, 256, 0, stream>>>(L.range_image, L.full, L.row_count, L.n_scan, L.horizon, L.extracted, L.col_ind, L.point_range,
                                                    L.start_ring, L.end_ring, L.n_extracted);
  return cudaGetLastError();
}

cudaError_t launch_features(const FeatLaunch& L, cudaStream_t stream) {
  if (L.n <= 0) return cudaErrorInvalidValue;
  const int nb = (L.n + 255) / 256;
  feat_init_kernel<<<nb, 256, 0, stream>>>(L.point_range, L.n_dev, L.n, L.curvature, L.picked, L.label);
  feat_occlusion_kernel<<<nb, 256, 0, stream>>>(L.point_range, L.col_ind, L.n_dev, L.n, L.picked);
  feat_select_kernel<<<L.n_scan, 256, 0, stream>>>(L.extracted, L.col_ind, L.start_ring, L.end_ring, L.n_dev, L.n, L.edge_thr, L.surf_thr, L.curvature,
                                                   L.picked, L.label, L.corner_ring, L.corner_cnt, L.error_flag);
  feat_surface_kernel<<<L.n_scan, 256, 0, stream>>>(L.extracted, L.start_ring, L.end_ring, L.label, L.leaf, L.horizon, L.surf_ring, L.surf_cnt, L.error_flag);
  feat_compact_kernel<<<L.n_scan, 256, 0, stream>>>(L.corner_ring, L.corner_cnt, L.surf_ring, L.surf_cnt, L.n_scan, L.horizon, L.corner_out, L.n_corner,
                                                    L.surf_out, L.n_surf);
  return cudaGetLastError();
}

}  // namespace rolo
