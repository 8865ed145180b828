// covariance.cu — RotVGICP::calculate_covariances on the GPU (rot_vgicp_impl.hpp:421-496) and the Morton-ordered spatial index.
//
// Reference: per point, pcl::search::KdTree::nearestKSearch(k=20) (FLANN, float L2), mean-centre the
// 4 x k neighbour matrix, cov = N N^T / k, regularise through JacobiSVD (PLANE: U diag(1,1,1e-3) V^T).
// Here (exact, no approximation of the neighbour set), for a whole GROUP of clouds per launch chain:
//   1. per cloud bounding box, 39-bit Morton codes (13 bits/axis), ONE radix sort of (cloud id | code, index) over the
//      concatenated group [cub::DeviceRadixSort is the only library call: plumbing, not the hot op], Morton-ordered copy of the
//      points, bucket table over the top code bits.
//   2. knn_chunk_kernel: one WARP owns 32 Morton-consecutive queries (one per lane) and searches them TOGETHER; persistent warps draw
//      chunks from a global ticket:
//        a. the 64 Morton-order neighbours of the chunk are staged once; every lane visits them outwards from its own position. The
//           k smallest DISTANCES live in registers as an ascending list (no index payload), updated in batches by sort-merge on a
//           comparator network (sort_networks.cuh); the window's k-th distance tau0 bounds the lane's search radius;
//        b. consecutive lanes whose balls' boxes are disjoint share no candidate: the chunk is cut into groups there. Per group the
//           union box is covered by <= 32 octree cells of one level, each one contiguous range of the sorted array (one bucketed
//           binary-search pair per LANE, not per query); a group whose cells hold more than KNN_RAW_CAP points sheds its widest
//           lanes to the fallback kernel;
//        c. the candidates of those ranges are clipped against the box and staged in shared memory; every lane scans all of them
//           (broadcast reads) against its fixed threshold tau0 and records position + distance of what passes in a per-lane list
//           in shared memory; the list is merged into the distance list once at the end (a list that runs full is brought up to
//           date early);
//        d. tau* = the exact k-th smallest distance. The neighbour SET is {d < tau*} plus the (k - m) smallest original
//           indices among {d == tau*}, taken from the window flags and the pass list: the same set the oracle's (distance, index)
//           order selects;
//        e. epilogue in the same kernel: mean / scatter in fp64, cyclic-Jacobi eigen-decomposition, regulariser, tile-planar
//           fp32 store. The neighbour lists (21 MB per 262144-point cloud) are never written.
//      Anything unusual (shed lanes, pass-list or tie overflow, non-finite query) puts the query on a fail list;
//   3. knn_fallback_kernel serves that list one warp per query (sorted top-32 across the lanes): Morton window, then the covering
//      cells of the ball's box — scanned directly when small, else refined by an octree descent with distance pruning (per-warp
//      stack, nearest cell first) — + the same epilogue. 0.3-0.9 % of the queries.
// Float distances are accumulated x,y,z without FMA contraction so the neighbour sets match the CPU oracle bit for bit.

#include "covariance.cuh"
#include "morton.cuh"
#include "sort_networks.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <float.h>
#include <algorithm>
#include <math_constants.h>
#include <limits.h>

namespace rolo {

namespace {

// a cloud descriptor reaches the index kernels either from a device array (group path, one per blockIdx.y) or by value
struct CloudRef {
  const KnnCloud* arr;
  KnnCloud one;
};
__device__ __forceinline__ KnnCloud get_cloud(const CloudRef& r) { return r.arr ? r.arr[blockIdx.y] : r.one; }

__global__ void bbox_init_kernel(int* bbox, int n_clouds) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_clouds * 8) return;
  const int d = t & 7;
  bbox[t] = d < 3 ? INT_MAX : (d < 6 ? INT_MIN : 0);
}

__global__ void bbox_kernel(CloudRef ref, int* bbox_all) {
  const KnnCloud C = get_cloud(ref);
  int* bbox = bbox_all + blockIdx.y * 8;
  const float4* __restrict__ pts = C.pts;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C.n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!((p.x - p.x == 0.f) && (p.y - p.y == 0.f) && (p.z - p.z == 0.f))) continue;   // NaN / Inf points must not blow up the box
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  #pragma unroll
  for (int d = 0; d < 3; d++) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  __shared__ float s_mn[3][8], s_mx[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    #pragma unroll
    for (int d = 0; d < 3; d++) { s_mn[d][warp] = mn[d]; s_mx[d][warp] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    float a = s_mn[d][0], b = s_mx[d][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { a = fminf(a, s_mn[d][w]); b = fmaxf(b, s_mx[d][w]); }
    atomicMin(&bbox[d], float_to_ordered(a));
    atomicMax(&bbox[3 + d], float_to_ordered(b));
  }
}

// key = cloud id << 39 | Morton code: one radix sort orders every cloud of the group
__global__ void morton_kernel(CloudRef ref, const int* __restrict__ bbox_all, unsigned long long* __restrict__ codes, int* __restrict__ idx) {
  const KnnCloud C = get_cloud(ref);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.n) return;
  const Quant q = make_quant(bbox_all + blockIdx.y * 8);
  const float4 p = C.pts[i];
  unsigned ix, iy, iz;
  quantize(q, p.x, p.y, p.z, ix, iy, iz);
  codes[C.off + i] = ((unsigned long long)blockIdx.y << CODE_BITS) | morton3(ix, iy, iz);
  idx[C.off + i] = i;
}

// Morton-ordered copy of the points (w = original index) + the sorted codes with the cloud bits stripped
__global__ void gather_sorted_kernel(CloudRef ref, const int* __restrict__ idx_sorted, const unsigned long long* __restrict__ codes_sorted,
                                     float4* __restrict__ sorted, unsigned long long* __restrict__ codes_clean) {
  const KnnCloud C = get_cloud(ref);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C.n) return;
  const int i = idx_sorted[C.off + j];
  float4 p = C.pts[i];
  p.w = __int_as_float(i);
  sorted[C.off + j] = p;
  if (codes_clean) codes_clean[C.off + j] = codes_sorted[C.off + j] & ((1ull << CODE_BITS) - 1ull);
}

// bucket_start[b] = first sorted position whose code has top bits >= b (b = 0 .. N_BUCKETS): every later
// lower_bound only searches inside one bucket
__global__ void bucket_table_kernel(CloudRef ref, const unsigned long long* __restrict__ codes_clean, int* __restrict__ bucket_all) {
  const KnnCloud C = get_cloud(ref);
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > N_BUCKETS) return;
  int* bucket_start = bucket_all + (size_t)blockIdx.y * (N_BUCKETS + 1);
  bucket_start[b] = (b == N_BUCKETS) ? C.n : lower_bound_u64(codes_clean + C.off, C.n, (unsigned long long)b << BUCKET_SHIFT);
}

__device__ __forceinline__ bool cand_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

// generic k (<= KCAP): local-memory arrays, insertion with early exit.
constexpr int KCAP = 64;
struct TopKDynamic {
  float d[KCAP]; int id[KCAP];
  __device__ __forceinline__ void reset(int k) { for (int t = 0; t < k; t++) { d[t] = FLT_MAX; id[t] = INT_MAX; } }
  __device__ __forceinline__ bool accepts(float dd, int ii, int k) const { return cand_less(dd, ii, d[k - 1], id[k - 1]); }
  __device__ __forceinline__ void insert(float dd, int ii, int k) {
    int t = k - 1;
    while (t > 0 && cand_less(dd, ii, d[t - 1], id[t - 1])) { d[t] = d[t - 1]; id[t] = id[t - 1]; t--; }
    d[t] = dd; id[t] = ii;
  }
  __device__ __forceinline__ float worst(int k) const { return d[k - 1]; }
  __device__ __forceinline__ bool full(int k) const { return id[k - 1] != INT_MAX; }
  __device__ __forceinline__ int get(int t) const { return id[t]; }
};

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 matrix (fp64). Q columns = eigenvectors.
__device__ __forceinline__ void jacobi_rot(double& app, double& aqq, double& apq, double& arp, double& arq,
                                           double& q0p, double& q0q, double& q1p, double& q1q, double& q2p, double& q2q) {
  if (apq == 0.0) return;
  const double theta = (aqq - app) / (2.0 * apq);
  const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = rsqrt(t * t + 1.0);
  const double s = t * c;
  app -= t * apq; aqq += t * apq; apq = 0.0;
  const double rp = c * arp - s * arq, rq = s * arp + c * arq; arp = rp; arq = rq;
  double a, b;
  a = c * q0p - s * q0q; b = s * q0p + c * q0q; q0p = a; q0q = b;
  a = c * q1p - s * q1q; b = s * q1p + c * q1q; q1p = a; q1q = b;
  a = c * q2p - s * q2q; b = s * q2p + c * q2q; q2p = a; q2q = b;
}

__device__ void eig3_sym(const double c[6], double lam[3], double Q[9]) {
  double a00 = c[0], a01 = c[1], a02 = c[2], a11 = c[3], a12 = c[4], a22 = c[5];
  double q00 = 1, q01 = 0, q02 = 0, q10 = 0, q11 = 1, q12 = 0, q20 = 0, q21 = 0, q22 = 1;
  #pragma unroll 1
  for (int sweep = 0; sweep < 12; sweep++) {
    const double off = a01 * a01 + a02 * a02 + a12 * a12;
    const double dg = a00 * a00 + a11 * a11 + a22 * a22;
    if (off <= 1e-36 * dg || off == 0.0) break;
    jacobi_rot(a00, a11, a01, a02, a12, q00, q01, q10, q11, q20, q21);   // (p,q) = (0,1), r = 2
    jacobi_rot(a00, a22, a02, a01, a12, q00, q02, q10, q12, q20, q22);   // (0,2), r = 1
    jacobi_rot(a11, a22, a12, a01, a02, q01, q02, q11, q12, q21, q22);   // (1,2), r = 0
  }
  lam[0] = a00; lam[1] = a11; lam[2] = a22;
  Q[0] = q00; Q[1] = q01; Q[2] = q02; Q[3] = q10; Q[4] = q11; Q[5] = q12; Q[6] = q20; Q[7] = q21; Q[8] = q22;
}

__device__ __forceinline__ void inv3_sym(const double c[6], double o[6]) {
  const double k0 = c[3] * c[5] - c[4] * c[4], k1 = c[2] * c[4] - c[1] * c[5], k2 = c[1] * c[4] - c[2] * c[3];
  const double idet = 1.0 / (c[0] * k0 + c[1] * k1 + c[2] * k2);
  o[0] = k0 * idet; o[1] = k1 * idet; o[2] = k2 * idet;
  o[3] = (c[0] * c[5] - c[2] * c[2]) * idet; o[4] = (c[1] * c[2] - c[0] * c[4]) * idet; o[5] = (c[0] * c[3] - c[1] * c[1]) * idet;
}

// rot_vgicp_impl.hpp:458-492
__device__ void regularize(const double cov[6], int method, double out[6]) {
  if (method == 0 /*NONE*/) {
    #pragma unroll
    for (int t = 0; t < 6; t++) out[t] = cov[t];
    return;
  }
  if (method == 4 /*FROBENIUS*/) {
    double C[6] = {cov[0] + 1e-3, cov[1], cov[2], cov[3] + 1e-3, cov[4], cov[5] + 1e-3}, Ci[6];
    inv3_sym(C, Ci);
    const double fro = sqrt(Ci[0] * Ci[0] + Ci[3] * Ci[3] + Ci[5] * Ci[5] + 2.0 * (Ci[1] * Ci[1] + Ci[2] * Ci[2] + Ci[4] * Ci[4]));
    #pragma unroll
    for (int t = 0; t < 6; t++) Ci[t] /= fro;
    inv3_sym(Ci, out);
    return;
  }
  double lam[3], Q[9];
  eig3_sym(cov, lam, Q);
  // singular values |lam| sorted descending; U = Q sign(lam), V = Q
  int o0 = 0, o1 = 1, o2 = 2;
  double s0 = fabs(lam[0]), s1 = fabs(lam[1]), s2 = fabs(lam[2]);
  if (s1 > s0) { double t = s0; s0 = s1; s1 = t; int u = o0; o0 = o1; o1 = u; }
  if (s2 > s0) { double t = s0; s0 = s2; s2 = t; int u = o0; o0 = o2; o2 = u; }
  if (s2 > s1) { double t = s1; s1 = s2; s2 = t; int u = o1; o1 = o2; o2 = u; }
  double v0, v1, v2;
  switch (method) {
    case 3: /*PLANE*/ v0 = 1.0; v1 = 1.0; v2 = 1e-3; break;
    case 1: /*MIN_EIG*/ v0 = fmax(s0, 1e-3); v1 = fmax(s1, 1e-3); v2 = fmax(s2, 1e-3); break;
    case 2: /*NORMALIZED_MIN_EIG*/ v0 = fmax(s0 / s0, 1e-3); v1 = fmax(s1 / s0, 1e-3); v2 = fmax(s2 / s0, 1e-3); break;
    default: /*PLANE_S*/ { const double sum = s0 + s1 + s2; v0 = s0 / sum; v1 = s1 / sum; v2 = 1e-3; } break;
  }
  const int ord[3] = {o0, o1, o2};
  const double val[3] = {v0, v1, v2};
  #pragma unroll
  for (int t = 0; t < 6; t++) out[t] = 0.0;
  #pragma unroll
  for (int e = 0; e < 3; e++) {
    const int col = ord[e];
    double qx = 0, qy = 0, qz = 0, l = 0;
    #pragma unroll
    for (int cidx = 0; cidx < 3; cidx++) if (cidx == col) { qx = Q[0 + cidx]; qy = Q[3 + cidx]; qz = Q[6 + cidx]; l = lam[cidx]; }
    const double sv = (l < 0.0) ? -val[e] : val[e];
    out[0] += sv * qx * qx; out[1] += sv * qx * qy; out[2] += sv * qx * qz;
    out[3] += sv * qy * qy; out[4] += sv * qy * qz; out[5] += sv * qz * qz;
  }
}



__device__ __forceinline__ float dist2_exact(const float4& q, const float4& p) {
  const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ void store_cov6(float* __restrict__ cov6, int i, const double out[6]) {
  const size_t base = (size_t)(i / TILE_PTS) * COV_PLANES * TILE_PTS + (i % TILE_PTS);
  #pragma unroll
  for (int t = 0; t < 6; t++) cov6[base + (size_t)t * TILE_PTS] = (float)out[t];
}

// ---------------------------------------------------------------------------------------------------
// chunk-cooperative exact kNN + covariance (see the file header). KC = compile-time capacity of the distance chain (k <= KC).
// ---------------------------------------------------------------------------------------------------
constexpr int KNN_WARPS = 8;          // warps (= chunks of 32 queries) per CTA; warps are independent (no CTA barrier)
constexpr int KNN_STAGE = 128;        // staged candidates per warp (float4), + 4 entries of padding
constexpr int KNN_PL = 28;            // per-lane list of (position, distance): a window phase queues <= 22 distances, the region phase
                                      // records the candidates within the window's k-th distance
constexpr int KNN_TIE = 16;           // tie entries per lane (aliases the stage buffer: 128 float4 = 16 x 32 ints)
#ifndef ROLO_KNN_RAW_CAP
#define ROLO_KNN_RAW_CAP 4096
#endif
constexpr int KNN_RAW_CAP = ROLO_KNN_RAW_CAP;     // candidates of one group before clipping; above it the widest lanes go to the fallback kernel
constexpr int KNN_WINDOW = 64;        // Morton-order neighbours scanned first (the chunk itself + 16 before + 16 after)
#ifndef ROLO_KNN_FILL
#define ROLO_KNN_FILL 20
#endif
constexpr int KNN_FILL = ROLO_KNN_FILL;          // window entries visited before the first list update (every lane's nearest Morton neighbours)
constexpr size_t KNN_STAGE_BYTES = (KNN_STAGE + 4) * sizeof(float4);
constexpr size_t KNN_SMEM_WARP = KNN_STAGE_BYTES + (size_t)KNN_PL * 32 * (sizeof(int) + sizeof(float));
constexpr size_t KNN_SMEM = KNN_WARPS * KNN_SMEM_WARP;
static_assert(KNN_TIE * 32 * sizeof(int) <= KNN_STAGE * sizeof(float4), "tie buffer aliases the stage buffer");
static_assert(KNN_WINDOW <= KNN_STAGE && (KNN_WINDOW & (KNN_WINDOW - 1)) == 0, "the window is staged as one tile and indexed modulo its size");
static_assert((KNN_WINDOW - KNN_FILL + 1) / 2 <= KNN_PL && KNN_FILL <= KNN_PL && 2 * KNN_PL >= 32, "list capacity");

template <int KC>
__device__ __forceinline__ void chain_insert(float (&a)[KC], float v) {
  // a is ascending; afterwards it holds the KC smallest of (a, v): slot i becomes min(a[i], max(a[i-1], v))
  #pragma unroll
  for (int i = KC - 1; i >= 1; i--) a[i] = fminf(a[i], fmaxf(a[i - 1], v));
  a[0] = fminf(a[0], v);
}

template <int KC>
__global__ void __launch_bounds__(KNN_WARPS * 32, 3)
knn_chunk_kernel(const KnnCloud* __restrict__ clouds, int n_clouds, int total_chunks, const float4* __restrict__ sorted_all,
                 const unsigned long long* __restrict__ codes_all, const int* __restrict__ bucket_all, const int* __restrict__ bbox_all,
                 int* __restrict__ fail, int k, int method, unsigned long long* stats) {
  extern __shared__ __align__(16) unsigned char knn_smem[];
  const unsigned FULL = 0xffffffffu;
  const float INF = CUDART_INF_F;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wsm = knn_smem + (size_t)warp * KNN_SMEM_WARP;
  float4* stage = reinterpret_cast<float4*>(wsm);
  int* pl_pos = reinterpret_cast<int*>(wsm + KNN_STAGE_BYTES) + lane;                                           // [slot * 32]
  float* pl_d = reinterpret_cast<float*>(wsm + KNN_STAGE_BYTES + (size_t)KNN_PL * 32 * sizeof(int)) + lane;     // [slot * 32]
  int* tie = reinterpret_cast<int*>(stage) + lane;                                                                         // [slot * 32]

  // persistent warps: the work per chunk varies by orders of magnitude (sparse far returns, chunks that straddle a jump of the
  // Morton curve), so every warp draws its next chunk from a global ticket instead of owning a fixed one
  #pragma unroll 1
  for (;;) {
  int chunk = 0;
  if (lane == 0) chunk = atomicAdd(fail + 1, 1);
  chunk = __shfl_sync(FULL, chunk, 0);
  if (chunk >= total_chunks) break;
  int cloud;
  {
    int lo_c = 0, hi_c = n_clouds - 1;                      // last cloud whose first chunk is <= chunk
    while (lo_c < hi_c) { const int mid = (lo_c + hi_c + 1) >> 1; if (__ldg(&clouds[mid].chunk_off) <= chunk) lo_c = mid; else hi_c = mid - 1; }
    cloud = lo_c;
  }
  const KnnCloud C = clouds[cloud];
  const int n = C.n;
  int j0;
  {
    const int cl = chunk - C.chunk_off;                      // this cloud's cl-th owned chunk
    if (C.shard_world > 1) { const int per = KNN_SHARD_BLOCK / 32; j0 = (((cl / per) * C.shard_world + C.shard_rank) * per + cl % per) * 32; }
    else j0 = cl * 32;
  }
  const float4* __restrict__ sorted = sorted_all + C.off;
  const unsigned long long* __restrict__ codes = codes_all + C.off;
  const int* __restrict__ bucket_start = bucket_all + (size_t)cloud * (N_BUCKETS + 1);

  const int j = j0 + lane;
  const bool in_cloud = j < n;
  const float4 q = __ldg(sorted + min(j, n - 1));            // lanes past the end repeat the last point and write nothing
  const bool finite = (q.x - q.x == 0.f) && (q.y - q.y == 0.f) && (q.z - q.z == 0.f);
  bool kicked = false;                                       // handed to the fallback kernel (warp-per-query search)
  bool overflow = false;

  float a[KC];                                               // ascending; a[KC-1] = k-th smallest distance seen so far
  #pragma unroll
  for (int i = 0; i < KC; i++) a[i] = (i < KC - k) ? -1.0f : INF;   // the low KC-k slots are pinned below every distance
  int npass = 0, nmerged = 0;                                // list entries; those below nmerged are already in `a`
  unsigned long long wmask = 0;                              // window entries that were within the running k-th distance when visited
  unsigned st_staged = 0, st_events = 0, st_groups = 0, st_raw = 0;

  // the Morton-order neighbourhood of the chunk: 64 candidates, staged once
  const int wcnt = min(KNN_WINDOW, n);
  const int wlo = max(0, min(j0 - 16, n - wcnt));
  #pragma unroll
  for (int h = 0; h < KNN_WINDOW / 32; h++) {
    const int w = h * 32 + lane;
    float4 c = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, 0.f);
    if (w < wcnt) c = __ldg(sorted + wlo + w);
    stage[w] = c;
  }
  __syncwarp();

  // Four phases, each ending in the same list update (one copy of the sorting network in the code):
  //   0: every lane visits its 20 nearest Morton neighbours (j, j+1, j-1, j+2, ...): near in Morton order is mostly near in space,
  //      so the first list is already good;
  //   1, 2: the other 44 window entries in two halves, each against the fixed threshold of the phase before;
  //   3: the candidates of the covering cells (the region), against the fixed threshold of phase 2: the lock-step scan only compares
  //      and records, no list is touched inside it.
  #pragma unroll 1
  for (int phase = 0; phase < 4; phase++) {
    if (phase < 3) {
      const int wself = min(j, n - 1) - wlo;
      const float thr = a[KC - 1];
      const int tmid = KNN_FILL + (KNN_WINDOW - KNN_FILL) / 2;
      const int t0 = phase == 0 ? 0 : (phase == 1 ? KNN_FILL : tmid), t1 = phase == 0 ? KNN_FILL : (phase == 1 ? tmid : KNN_WINDOW);
      #pragma unroll 1
      for (int t = t0; t < t1; t++) {
        const int off = (t & 1) ? ((t + 1) >> 1) : -(t >> 1);
        const int w = (wself + off) & (KNN_WINDOW - 1);
        const float d = dist2_exact(q, stage[w]);
        if (finite && d <= thr) { pl_d[npass * 32] = d; npass++; wmask |= 1ull << w; }
      }
    } else {
      __syncwarp();                                          // the stage buffer is about to be reused
      // ---- search box of every lane (rounded outwards); consecutive lanes whose boxes are disjoint share no candidate, so the chunk is
      //      cut there and the groups are searched one after the other (a chunk that straddles a jump of the Morton curve would
      //      otherwise drag in everything between its two ends)
      const float thr = a[KC - 1];
      float l0, l1, l2, h0, h1, h2;
      {
        const float R = __fmul_ru(__fsqrt_ru(thr), 1.000001f);
        l0 = __fsub_rd(q.x, R); l1 = __fsub_rd(q.y, R); l2 = __fsub_rd(q.z, R);
        h0 = __fadd_ru(q.x, R); h1 = __fadd_ru(q.y, R); h2 = __fadd_ru(q.z, R);
      }
      unsigned cuts;
      {
        const float nl0 = __shfl_down_sync(FULL, l0, 1), nl1 = __shfl_down_sync(FULL, l1, 1), nl2 = __shfl_down_sync(FULL, l2, 1);
        const float nh0 = __shfl_down_sync(FULL, h0, 1), nh1 = __shfl_down_sync(FULL, h1, 1), nh2 = __shfl_down_sync(FULL, h2, 1);
        const bool nfin = __shfl_down_sync(FULL, finite ? 1 : 0, 1) != 0;
        const bool overlap = nl0 <= h0 && l0 <= nh0 && nl1 <= h1 && l1 <= nh1 && nl2 <= h2 && l2 <= nh2;   // false for NaN boxes
        cuts = __ballot_sync(FULL, lane < 31 && !(overlap && finite && nfin)) | 0x80000000u;
      }
      const Quant qn = make_quant(bbox_all + cloud * 8);
      #pragma unroll 1
      for (int gstart = 0; gstart < 32;) {
        const int gend = gstart + __ffs(cuts >> gstart) - 1;    // the group is lanes gstart..gend
        const bool act = finite && lane >= gstart && lane <= gend;
        gstart = gend + 1;
        if (!__any_sync(FULL, act)) continue;
        st_groups++;
        float lox, loy, loz, hix, hiy, hiz;
        int lo, cnt, incl, total, excl;
        bool act_g = act;
        #pragma unroll 1
        for (int attempt = 0;; attempt++) {
          lox = ordered_to_float(__reduce_min_sync(FULL, float_to_ordered(act_g ? l0 : INF)));
          loy = ordered_to_float(__reduce_min_sync(FULL, float_to_ordered(act_g ? l1 : INF)));
          loz = ordered_to_float(__reduce_min_sync(FULL, float_to_ordered(act_g ? l2 : INF)));
          hix = ordered_to_float(__reduce_max_sync(FULL, float_to_ordered(act_g ? h0 : -INF)));
          hiy = ordered_to_float(__reduce_max_sync(FULL, float_to_ordered(act_g ? h1 : -INF)));
          hiz = ordered_to_float(__reduce_max_sync(FULL, float_to_ordered(act_g ? h2 : -INF)));
          // covering cells: the finest level at which the box spans <= 32 cells; lane c looks up cell c (one range of the sorted array)
          lo = 0; cnt = 0;
          {
            unsigned ix0, iy0, iz0, ix1, iy1, iz1;
            quantize(qn, lox, loy, loz, ix0, iy0, iz0);
            quantize(qn, hix, hiy, hiz, ix1, iy1, iz1);
            const unsigned ext = max(max(ix1 - ix0, iy1 - iy0), iz1 - iz0);
            int sh = max(0, 30 - __clz((int)ext));          // ext >> sh < 4
            int ncx, ncy, ncz;
            #pragma unroll 1
            for (;; sh++) {
              ncx = (int)(ix1 >> sh) - (int)(ix0 >> sh) + 1; ncy = (int)(iy1 >> sh) - (int)(iy0 >> sh) + 1; ncz = (int)(iz1 >> sh) - (int)(iz0 >> sh) + 1;
              if (ncx * ncy * ncz <= 32 || sh >= MORTON_BITS) break;
            }
            if (lane < ncx * ncy * ncz) {
              const unsigned cx = (ix0 >> sh) + (unsigned)(lane % ncx), cy = (iy0 >> sh) + (unsigned)((lane / ncx) % ncy), cz = (iz0 >> sh) + (unsigned)(lane / (ncx * ncy));
              const unsigned long long pre = morton3(cx, cy, cz);
              lo = lower_bound_bucketed(codes, bucket_start, pre << (3 * sh));
              const int hi = (sh >= MORTON_BITS) ? n : lower_bound_bucketed(codes, bucket_start, (pre + 1ull) << (3 * sh));
              cnt = hi - lo;
            }
          }
          incl = cnt;
          #pragma unroll
          for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += v; }
          total = __shfl_sync(FULL, incl, 31);
          excl = incl - cnt;
          if (total <= KNN_RAW_CAP) break;
          // too many candidates for a lock-step scan: the lane with the widest box goes to the fallback kernel, which searches one
          // query with all 32 lanes; after a few of those the whole group goes
          const float myR = act_g ? (h0 - l0) : -1.0f;
          const float maxR = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(fmaxf(myR, 0.f))));
          const bool all_go = attempt >= 3;
          if (act_g && (all_go || myR == maxR)) { kicked = true; act_g = false; }
          if (!__any_sync(FULL, act_g)) { total = 0; break; }
        }
        st_raw += total;

        // ---- stage the candidates of the ranges (clipped against the box, the window left out) and scan them tile by tile.
        //      Two candidates per lane and step: two independent loads in flight
        int ns = 0;
        #pragma unroll 1
        for (int base = 0; base < total; base += 64) {
          float4 c[2];
          bool keep[2];
          #pragma unroll
          for (int h = 0; h < 2; h++) {
            const int g = base + h * 32 + lane;
            int o = 0, b = 31;                               // owning cell: the first lane whose inclusive count exceeds g
            #pragma unroll
            for (int s = 0; s < 5; s++) {
              const int mid = (o + b) >> 1;
              const int v = __shfl_sync(FULL, incl, mid);
              if (v <= g) o = mid + 1; else b = mid;
            }
            const int cell_lo = __shfl_sync(FULL, lo, o), cell_ex = __shfl_sync(FULL, excl, o);
            const int pos = cell_lo + (g - cell_ex);
            c[h] = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, 0.f);
            if (g < total) c[h] = __ldg(sorted + pos);
            keep[h] = c[h].x >= lox && c[h].x <= hix && c[h].y >= loy && c[h].y <= hiy && c[h].z >= loz && c[h].z <= hiz &&
                      (unsigned)(pos - wlo) >= (unsigned)wcnt;   // false for the NaN filler
            c[h].w = __int_as_float(pos);
          }
          #pragma unroll
          for (int h = 0; h < 2; h++) {
            const unsigned km = __ballot_sync(FULL, keep[h]);
            if (keep[h]) stage[ns + __popc(km & ((1u << lane) - 1u))] = c[h];
            ns += __popc(km);
          }
          if (ns > KNN_STAGE - 64 || base + 64 >= total) {
            if (lane < 3) stage[ns + lane] = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, 0.f);   // pad to a multiple of 4
            __syncwarp();
            st_staged += ns;
            #pragma unroll 1
            for (int s = 0; s < ns; s += 4) {
              float cur = a[KC - 1];                        // == thr unless a list ran full (below)
              float4 p[4];
              float d[4];
              #pragma unroll
              for (int h = 0; h < 4; h++) p[h] = stage[s + h];   // broadcast reads
              #pragma unroll
              for (int h = 0; h < 4; h++) d[h] = dist2_exact(q, p[h]);
              const float dmin = fminf(fminf(d[0], d[1]), fminf(d[2], d[3]));   // NaN-safe: fminf drops the NaN filler
              if (__any_sync(FULL, act_g && dmin <= cur)) {
                st_events++;
                #pragma unroll
                for (int h = 0; h < 4; h++) {
                  if (act_g && d[h] <= cur) {
                    if (npass < KNN_PL) { pl_pos[npass * 32] = __float_as_int(p[h].w); pl_d[npass * 32] = d[h]; npass++; } else overflow = true;
                  }
                }
                // a list that is nearly full is brought up to date by single insertions and stripped of what fell behind; rare
                if (__any_sync(FULL, npass > KNN_PL - 4)) {
                  const int mx = __reduce_max_sync(FULL, npass - nmerged);
                  #pragma unroll 1
                  for (int e = 0; e < mx; e++) chain_insert<KC>(a, nmerged + e < npass ? pl_d[(nmerged + e) * 32] : INF);
                  const float lim = a[KC - 1];
                  int m2 = 0;
                  #pragma unroll 1
                  for (int e = 0; e < npass; e++) {
                    const float dd = pl_d[e * 32];
                    if (dd <= lim) { pl_pos[m2 * 32] = pl_pos[e * 32]; pl_d[m2 * 32] = dd; m2++; }
                  }
                  npass = m2; nmerged = m2;
                  if (npass > KNN_PL - 4) overflow = true;  // that many candidates tie inside the k-th distance
                }
              }
            }
            ns = 0;
            __syncwarp();
          }
        }
      }
    }

    // ---- list update: sort up to KC pending distances (comparator network on registers), keep the KC smallest of the two sorted lists
    //      (min(a[i], x[KC-1-i]) is that set) and sort again; repeated while a lane has distances pending
    #pragma unroll 1
    for (;;) {
      const int pending = npass - nmerged;
      if (!__any_sync(FULL, pending > 0)) break;
      float x[KC];
      #pragma unroll
      for (int i = 0; i < KC; i++) x[i] = i < pending ? pl_d[(nmerged + i) * 32] : INF;
      nmerged += min(pending, KC);
      #pragma unroll 1
      for (int pass = 0; pass < 2; pass++) {
        sort_network<KC>(x);
        if (pass == 0) {
          #pragma unroll
          for (int i = 0; i < KC / 2; i++) { const float u = x[i], v = x[KC - 1 - i]; x[i] = fminf(a[i], v); x[KC - 1 - i] = fminf(a[KC - 1 - i], u); }
          if (KC & 1) x[KC / 2] = fminf(a[KC / 2], x[KC / 2]);
        }
      }
      #pragma unroll
      for (int i = 0; i < KC; i++) a[i] = x[i];
    }
    if (phase < 3) { npass = 0; nmerged = 0; }              // window entries are tracked by wmask, not by the list
  }
  if (stats) {
    if (lane == 0) {
      atomicAdd(stats + 0, (unsigned long long)st_raw); atomicAdd(stats + 1, (unsigned long long)st_staged); atomicAdd(stats + 2, (unsigned long long)st_events);
      atomicAdd(stats + 3, (unsigned long long)st_groups); atomicAdd(stats + 6, 1ull); atomicMax(stats + 5, (unsigned long long)st_raw);
    }
    atomicAdd(stats + 4, (unsigned long long)npass);
  }

  // ---- the neighbour set: everything below tau*, then the smallest original indices among the ties at tau*; the candidates are the
  //      window entries flagged in `wmask` and the list (whose distances are stored). The selected positions overwrite the list from
  //      the front (m never runs ahead of the read position; the window part is capped at the k-1 entries that can be below tau*)
  const float tau = a[KC - 1];
  int m = 0, nt = 0;
  // 1. list entries first, compacted in place
  {
    int mm = 0;
    #pragma unroll 1
    for (int e = 0; e < npass; e++) {
      const float d = pl_d[e * 32];
      const int pos = pl_pos[e * 32];
      if (d < tau) { pl_pos[mm * 32] = pos; mm++; }
      else if (d == tau) { if (nt < KNN_TIE) tie[nt * 32] = pos; nt++; }
    }
    m = mm;
  }
  // 2. window entries appended behind them: at most k - 1 entries in all; the position half and the distance half of the list are
  //    contiguous, so the selection may use both (2 * KNN_PL >= 32 slots)
  #pragma unroll 1
  while (wmask) {
    const int w = __ffsll((long long)wmask) - 1;
    wmask &= wmask - 1;
    const float d = dist2_exact(q, __ldg(sorted + wlo + w));
    if (d < tau) { if (m < 2 * KNN_PL) pl_pos[m * 32] = wlo + w; m++; }      // past KNN_PL the entries run on into the (now dead) distance half
    else if (d == tau) { if (nt < KNN_TIE) tie[nt * 32] = wlo + w; nt++; }
  }
  const int need = k - m;
  bool ok = finite && !kicked && !overflow && need >= 1 && nt >= need && nt <= KNN_TIE;
  if (ok) {
    if (nt == need) {
      for (int u = 0; u < nt; u++) pl_pos[(m + u) * 32] = tie[u * 32];
    } else {
      int prev = -1;
      for (int u = 0; u < need; u++) {
        int best = INT_MAX, best_pos = 0;
        for (int v = 0; v < nt; v++) {
          const int pos = tie[v * 32];
          const int oi = __float_as_int(__ldg(sorted + pos).w);
          if (oi > prev && oi < best) { best = oi; best_pos = pos; }
        }
        pl_pos[(m + u) * 32] = best_pos; prev = best;
      }
    }
  }

  // ---- covariance of the k neighbours (rot_vgicp_impl.hpp:446-456) + regulariser (:458-492)
  if (in_cloud) {
    const int self = __float_as_int(q.w);
    if (ok) {
      double sx = 0, sy = 0, sz = 0;
      #pragma unroll 4
      for (int u = 0; u < k; u++) { const float4 p = __ldg(sorted + pl_pos[u * 32]); sx += (double)p.x; sy += (double)p.y; sz += (double)p.z; }
      const double mx = sx / k, my = sy / k, mz = sz / k;
      double cc[6] = {0, 0, 0, 0, 0, 0};
      #pragma unroll 4
      for (int u = 0; u < k; u++) {
        const float4 p = __ldg(sorted + pl_pos[u * 32]);
        const double dx = (double)p.x - mx, dy = (double)p.y - my, dz = (double)p.z - mz;
        cc[0] += dx * dx; cc[1] += dx * dy; cc[2] += dx * dz; cc[3] += dy * dy; cc[4] += dy * dz; cc[5] += dz * dz;
        if (C.knn_out) C.knn_out[(size_t)self * k + u] = __float_as_int(p.w);
      }
      if (C.cov6) {
        #pragma unroll
        for (int t = 0; t < 6; t++) cc[t] /= (double)k;
        double out[6];
        regularize(cc, method, out);
        store_cov6(C.cov6, self, out);
      }
    } else {
      const int slot = atomicAdd(fail, 1);
      fail[2 + 2 * slot] = cloud; fail[3 + 2 * slot] = j;
    }
  }
  __syncwarp();
  }   // persistent chunk loop
}

// ---------------------------------------------------------------------------------------------------
// exact kNN of ONE query by a whole warp (k <= 32): the fallback behind the chunk kernel, for the queries whose candidate region is too
// large for a lock-step scan (a stray point next to a dense surface, a far return whose neighbours are metres away) or whose lists
// overflowed (exact ties en masse).
//  * the current top-32 lives one element per lane, sorted by (distance, index); candidates arrive 32 at a time (coalesced); a batch in
//    which nobody beats the k-th key costs one ballot; otherwise the winners are inserted by shuffle (few) or the batch is
//    bitonic-sorted and merged (many);
//  * the 64 Morton-order neighbours of the query go first: their k-th distance bounds the search radius, so ONE pass over the cells
//    that cover the ball's box (<= 32 cells of one level, one binary-search pair per lane) is exact — no level iteration, no restart.
// Returns this lane's element of the ascending (distance, index) list (KEY_INF = none).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void knn_warp_merge(unsigned long long& top, unsigned long long key, int k, bool first) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned long long kth = __shfl_sync(FULL, top, k - 1);
  const bool pass = key < kth;
  unsigned m = __ballot_sync(FULL, pass);
  if (m == 0) return;
  if (__popc(m) > 10) {
    // bitonic sort of the batch (non-passing lanes carry +inf), then bitonic merge with the top-32
    if (!pass) key = KEY_INF;
    #pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
      #pragma unroll
      for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) bitonic_cmpx(key, j2, ((lane & k2) == 0) == ((lane & j2) == 0));
    }
    if (first) {
      top = key;                                     // the top-32 is still empty: the sorted batch is the new top-32
    } else {
      const unsigned long long rev = __shfl_sync(FULL, key, 31 - lane);
      if (rev < top) top = rev;
      #pragma unroll
      for (int j2 = 16; j2 > 0; j2 >>= 1) bitonic_cmpx(top, j2, (lane & j2) == 0);
    }
  } else {
    while (m) {
      const int srcl = __ffs(m) - 1;
      m &= m - 1;
      const unsigned long long ck = __shfl_sync(FULL, key, srcl);
      const int pos = __popc(__ballot_sync(FULL, top < ck));      // elements that stay in front
      const unsigned long long up = __shfl_up_sync(FULL, top, 1);
      if (lane > pos) top = up; else if (lane == pos) top = ck;
    }
  }
}

// Descent: the cells that cover the ball's box are refined where it pays. Work items (a cell = one range of the sorted array) live on a
// per-warp stack, nearest cell on top; a popped cell farther from the query than the current k-th distance is dropped, a small one is
// scanned (32 candidates per step into the top-32), a large one is split into its 8 children (7 binary searches inside its range, one
// per lane), which are pushed farthest first. The k-th distance shrinks as the near cells are scanned, so a stray point beside a dense
// surface touches the few hundred candidates around its neighbours instead of the thousands inside its first box.
constexpr int KNN_DESC_STACK = 160;   // 32 covering cells + 7 per level of refinement (13 levels)
#ifndef ROLO_KNN_FALLBACK_GRID
#define ROLO_KNN_FALLBACK_GRID 592     // CTAs of 8 warps (one query per warp at a time): 4 per SM — the descent waits on dependent loads, more warps hide them
#endif
#ifndef ROLO_KNN_DESC_LEAF
#define ROLO_KNN_DESC_LEAF 128
#endif
#ifndef ROLO_KNN_DESC_MIN
#define ROLO_KNN_DESC_MIN 512
#endif
constexpr int KNN_DESC_LEAF = ROLO_KNN_DESC_LEAF;    // cells up to this size are scanned, larger ones split
constexpr int KNN_DESC_SPLITS = 4096; // splits per query before the descent stops refining (a guard, never reached on real clouds)
constexpr int KNN_DESC_MIN = ROLO_KNN_DESC_MIN;     // covering cells holding at most this many points are scanned directly (no stack)

struct DescItem { int lo, cnt; unsigned long long cell; };   // cell = cx | cy << 13 | cz << 26 | shift << 39

// squared distance from the query (rel = q - bbox min, double) to the cell, a lower bound: the cell is widened by a slack that covers the
// rounding of quantize()
__device__ __forceinline__ double cell_min_d2(const Quant& qn, const double rel[3], unsigned cx, unsigned cy, unsigned cz, int sh) {
  const double sz = qn.cell * (double)(1u << sh), slack = qn.cell * 1e-6;
  const unsigned cc[3] = {cx, cy, cz};
  double d2 = 0.0;
  #pragma unroll
  for (int t = 0; t < 3; t++) {
    const double lo = (double)cc[t] * sz - slack, hi = (double)(cc[t] + 1u) * sz + slack;
    const double d = fmax(fmax(lo - rel[t], rel[t] - hi), 0.0);
    d2 += d * d;
  }
  return d2;
}

// the k-th distance as a pruning bound: a candidate whose FLOAT distance ties or beats it lies at a true distance within a few ulp of it
__device__ __forceinline__ double prune_bound(unsigned long long kth) {
  return kth == KEY_INF ? 1e300 : (double)__uint_as_float((unsigned)(kth >> 32)) * (1.0 + 1e-6);
}

// lane c < 32 may own one covering cell (lo, cnt, cx, cy, cz at shift sh); `top` holds the window's result. Positions inside the window
// [wlo, wlo + wcnt) are in the list already and are skipped.
__device__ unsigned long long knn_warp_descent(const float4* __restrict__ sorted, const unsigned long long* __restrict__ codes, int k, const Quant& qn,
                                               const float4 q, int wlo, int wcnt, unsigned long long top, DescItem* stack,
                                               int lo, int cnt, unsigned cx, unsigned cy, unsigned cz, int sh) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const double rel[3] = {(double)q.x - qn.mn[0], (double)q.y - qn.mn[1], (double)q.z - qn.mn[2]};
  int sp;
  {
    const double bound0 = prune_bound(__shfl_sync(FULL, top, k - 1));   // every lane takes part in the shuffle
    const double m2 = cnt > 0 ? cell_min_d2(qn, rel, cx, cy, cz, sh) : 0.0;
    const bool valid = cnt > 0 && m2 <= bound0;
    const unsigned vm = __ballot_sync(FULL, valid);
    int r = 0;                                           // farther valid cells go below this one
    #pragma unroll 4
    for (int s2 = 0; s2 < 32; s2++) {
      const double o = __shfl_sync(FULL, m2, s2);
      if (((vm >> s2) & 1u) && (o > m2 || (o == m2 && s2 > lane))) r++;
    }
    if (valid) { DescItem it; it.lo = lo; it.cnt = cnt; it.cell = (unsigned long long)cx | ((unsigned long long)cy << 13) | ((unsigned long long)cz << 26) | ((unsigned long long)sh << 39); stack[r] = it; }
    sp = __popc(vm);
    __syncwarp();
  }
  int splits = 0;                                       // bounded: past the budget the remaining cells are scanned as they are
  #pragma unroll 1
  while (sp > 0) {
    sp--;
    const DescItem it = stack[sp];                       // every lane reads the same entry
    __syncwarp();
    const unsigned icx = (unsigned)(it.cell & 0x1FFFu), icy = (unsigned)((it.cell >> 13) & 0x1FFFu), icz = (unsigned)((it.cell >> 26) & 0x1FFFu);
    const int ish = (int)(it.cell >> 39);
    const double bound = prune_bound(__shfl_sync(FULL, top, k - 1));
    if (cell_min_d2(qn, rel, icx, icy, icz, ish) > bound) continue;
    if (it.cnt <= KNN_DESC_LEAF || ish == 0 || sp + 8 > KNN_DESC_STACK || ++splits > KNN_DESC_SPLITS) {
      #pragma unroll 1
      for (int base = 0; base < it.cnt; base += 32) {
        const int pos = it.lo + base + lane;
        unsigned long long key = KEY_INF;
        if (base + lane < it.cnt && (unsigned)(pos - wlo) >= (unsigned)wcnt) {
          const float4 p = __ldg(sorted + pos);
          const float d = dist2_exact(q, p);
          if (d == d) key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
        }
        knn_warp_merge(top, key, k, false);
      }
      continue;
    }
    // split: the children are consecutive in Morton order, child c (x bit 0, y bit 1, z bit 2) = [b_c, b_(c+1))
    const unsigned long long P = morton3(icx, icy, icz);
    int b = lane == 0 ? it.lo : it.lo + it.cnt;
    if (lane >= 1 && lane <= 7) {
      const unsigned long long key = ((P << 3) + (unsigned long long)lane) << (3 * (ish - 1));
      int l2 = it.lo, h2 = it.lo + it.cnt;
      while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (__ldg(codes + mid) < key) l2 = mid + 1; else h2 = mid; }
      b = l2;
    }
    const int bn = __shfl_down_sync(FULL, b, 1);
    const int ccnt = lane < 8 ? bn - b : 0;
    const unsigned ccx = 2u * icx + (lane & 1), ccy = 2u * icy + ((lane >> 1) & 1), ccz = 2u * icz + ((lane >> 2) & 1);
    const double m2 = ccnt > 0 ? cell_min_d2(qn, rel, ccx, ccy, ccz, ish - 1) : 0.0;
    const bool valid = ccnt > 0 && m2 <= bound;
    const unsigned vm = __ballot_sync(FULL, valid);
    int r = 0;
    #pragma unroll
    for (int s2 = 0; s2 < 8; s2++) {
      const double o = __shfl_sync(FULL, m2, s2);
      if (((vm >> s2) & 1u) && (o > m2 || (o == m2 && s2 > lane))) r++;
    }
    if (valid) { DescItem ch; ch.lo = b; ch.cnt = ccnt; ch.cell = (unsigned long long)ccx | ((unsigned long long)ccy << 13) | ((unsigned long long)ccz << 26) | ((unsigned long long)(ish - 1) << 39); stack[sp + r] = ch; }
    sp += __popc(vm);
    __syncwarp();
  }
  return top;
}

__device__ unsigned long long knn_warp_query(const float4* __restrict__ sorted, const unsigned long long* __restrict__ codes,
                                             const int* __restrict__ bucket_start, int n, int k, const Quant& qn, int j, const float4 q,
                                             DescItem* __restrict__ stack) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  unsigned long long top = KEY_INF;                   // this lane's element of the sorted top-32
  auto key_of = [&](const float4& p) {
    const float d = dist2_exact(q, p);
    return (d == d) ? (((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w)) : KEY_INF;
  };
  // the Morton window first
  const int wcnt = min(KNN_WINDOW, n);
  const int wlo = max(0, min(j - KNN_WINDOW / 2, n - wcnt));
  #pragma unroll
  for (int h = 0; h < KNN_WINDOW / 32; h++) {
    const int w = h * 32 + lane;
    unsigned long long key = KEY_INF;
    if (w < wcnt) key = key_of(__ldg(sorted + wlo + w));
    knn_warp_merge(top, key, k, h == 0);
  }
  const unsigned long long kthw = __shfl_sync(FULL, top, k - 1);
  if (kthw == KEY_INF) return top;                    // fewer than k finite candidates in the window: only a non-finite query gets here
  // the ball's box and its covering cells
  const float tau = __uint_as_float((unsigned)(kthw >> 32));
  const float R = __fmul_ru(__fsqrt_ru(tau), 1.000001f);
  const float lox = __fsub_rd(q.x, R), loy = __fsub_rd(q.y, R), loz = __fsub_rd(q.z, R);
  const float hix = __fadd_ru(q.x, R), hiy = __fadd_ru(q.y, R), hiz = __fadd_ru(q.z, R);
  int lo = 0, cnt = 0, sh;
  unsigned cx = 0, cy = 0, cz = 0;
  {
    unsigned ix0, iy0, iz0, ix1, iy1, iz1;
    quantize(qn, lox, loy, loz, ix0, iy0, iz0);
    quantize(qn, hix, hiy, hiz, ix1, iy1, iz1);
    const unsigned ext = max(max(ix1 - ix0, iy1 - iy0), iz1 - iz0);
    sh = max(0, 30 - __clz((int)ext));
    int ncx, ncy, ncz;
    #pragma unroll 1
    for (;; sh++) {
      ncx = (int)(ix1 >> sh) - (int)(ix0 >> sh) + 1; ncy = (int)(iy1 >> sh) - (int)(iy0 >> sh) + 1; ncz = (int)(iz1 >> sh) - (int)(iz0 >> sh) + 1;
      if (ncx * ncy * ncz <= 32 || sh >= MORTON_BITS) break;
    }
    if (lane < ncx * ncy * ncz) {
      cx = (ix0 >> sh) + (unsigned)(lane % ncx); cy = (iy0 >> sh) + (unsigned)((lane / ncx) % ncy); cz = (iz0 >> sh) + (unsigned)(lane / (ncx * ncy));
      const unsigned long long pre = morton3(cx, cy, cz);
      lo = lower_bound_bucketed(codes, bucket_start, pre << (3 * sh));
      const int hi = (sh >= MORTON_BITS) ? n : lower_bound_bucketed(codes, bucket_start, (pre + 1ull) << (3 * sh));
      cnt = hi - lo;
    }
  }
  int incl = cnt;
  #pragma unroll
  for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += v; }
  const int total = __shfl_sync(FULL, incl, 31);
  const int excl = incl - cnt;
  if (total > KNN_DESC_MIN) return knn_warp_descent(sorted, codes, k, qn, q, wlo, wcnt, top, stack, lo, cnt, cx, cy, cz, sh);
  #pragma unroll 1
  for (int base = 0; base < total; base += 32) {
    const int g = base + lane;
    int o = 0, b = 31;
    #pragma unroll
    for (int s = 0; s < 5; s++) {
      const int mid = (o + b) >> 1;
      const int v = __shfl_sync(FULL, incl, mid);
      if (v <= g) o = mid + 1; else b = mid;
    }
    const int cell_lo = __shfl_sync(FULL, lo, o), cell_ex = __shfl_sync(FULL, excl, o);
    unsigned long long key = KEY_INF;
    if (g < total) {
      const int pos = cell_lo + (g - cell_ex);
      if ((unsigned)(pos - wlo) >= (unsigned)wcnt) {            // window entries are in the list already
        const float4 p = __ldg(sorted + pos);
        if (p.x >= lox && p.x <= hix && p.y >= loy && p.y <= hiy && p.z >= loz && p.z <= hiz) key = key_of(p);
      }
    }
    knn_warp_merge(top, key, k, false);
  }
  return top;
}

// The queries the chunk kernel could not finish (fail list), one warp each: exact search, then the same covariance epilogue with the
// k neighbours held one per lane (fp64 sums by butterfly). A query without k valid neighbours (non-finite coordinates) gets a
// zero covariance (such a point never finds a voxel, so the value is never used).
__global__ void __launch_bounds__(256)
knn_fallback_kernel(const KnnCloud* __restrict__ clouds, const float4* __restrict__ sorted_all, const unsigned long long* __restrict__ codes_all,
                    const int* __restrict__ bucket_all, const int* __restrict__ bbox_all, const int* __restrict__ fail, int k, int method) {
  __shared__ DescItem desc_stack[8][KNN_DESC_STACK];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int nfail = fail[0];
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  #pragma unroll 1
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < nfail; e += warps_total) {
    const int cloud = fail[2 + 2 * e], j = fail[3 + 2 * e];
    const KnnCloud C = clouds[cloud];
    const float4* __restrict__ sorted = sorted_all + C.off;
    const Quant qn = make_quant(bbox_all + cloud * 8);
    const float4 q = __ldg(sorted + j);
    const int self = __float_as_int(q.w);
    const unsigned long long top = knn_warp_query(sorted, codes_all + C.off, bucket_all + (size_t)cloud * (N_BUCKETS + 1), C.n, k, qn, j, q, desc_stack[threadIdx.x >> 5]);
    const bool mine = lane < k;
    const unsigned idx = (unsigned)(top & 0xFFFFFFFFull);
    const bool good = !mine || (top != KEY_INF && idx < (unsigned)C.n);
    const bool all_good = __all_sync(FULL, good);
    double v[3] = {0, 0, 0};
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mine && all_good) { p = __ldg(C.pts + idx); v[0] = (double)p.x; v[1] = (double)p.y; v[2] = (double)p.z; }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      #pragma unroll
      for (int t = 0; t < 3; t++) v[t] += __shfl_xor_sync(FULL, v[t], o);
    }
    double cc[6] = {0, 0, 0, 0, 0, 0};
    if (mine && all_good) {
      const double dx = (double)p.x - v[0] / k, dy = (double)p.y - v[1] / k, dz = (double)p.z - v[2] / k;
      cc[0] = dx * dx; cc[1] = dx * dy; cc[2] = dx * dz; cc[3] = dy * dy; cc[4] = dy * dz; cc[5] = dz * dz;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      #pragma unroll
      for (int t = 0; t < 6; t++) cc[t] += __shfl_xor_sync(FULL, cc[t], o);
    }
    if (C.knn_out && mine) C.knn_out[(size_t)self * k + lane] = all_good ? (int)idx : -1;
    if (C.cov6 && lane == 0) {
      double out[6];
      if (all_good) {
        #pragma unroll
        for (int t = 0; t < 6; t++) cc[t] /= (double)k;
        regularize(cc, method, out);
      } else {
        #pragma unroll
        for (int t = 0; t < 6; t++) out[t] = 0.0;
      }
      store_cov6(C.cov6, self, out);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// exact nearest neighbour of EXTERNAL query points (pcl::Registration::getFitnessScore's nearestKSearch(1) over the target),
// one warp per query. The query is optionally moved by a 3x4 float transform first (row-major; ((a x + b y) + c z) + d with
// every operation rounded, like the oracle). Start level: the insertion position of the query's Morton code among the sorted
// codes gives, through the common prefix with its two neighbours, the finest level at which the query's cell is occupied;
// the 3x3x3 block there is scanned (lane-wise minimum, one warp reduction), the result is exact once the distance is inside
// the radius the block covers, otherwise one level coarser. Queries outside the bounding box are clamped for the cell
// lookup; the coverage test uses their true coordinates, so they stay exact (and cost more the farther away they are).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nn1_warp_kernel(const float4* __restrict__ queries, int nq, const float* __restrict__ T12, const float4* __restrict__ sorted,
                const unsigned long long* __restrict__ codes, const int* __restrict__ bucket_start, int n, const int* __restrict__ bbox,
                int* __restrict__ out_idx, float* __restrict__ out_d2) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int j = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (j >= nq) return;
  const Quant qn = make_quant(bbox);
  float4 q = __ldg(queries + j);
  if (T12) {
    const float x = q.x, y = q.y, z = q.z;
    q.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T12[0], x), __fmul_rn(T12[1], y)), __fmul_rn(T12[2], z)), T12[3]);
    q.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T12[4], x), __fmul_rn(T12[5], y)), __fmul_rn(T12[6], z)), T12[7]);
    q.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T12[8], x), __fmul_rn(T12[9], y)), __fmul_rn(T12[10], z)), T12[11]);
  }
  unsigned ix, iy, iz;
  quantize(qn, q.x, q.y, q.z, ix, iy, iz);
  const unsigned long long cq = morton3(ix, iy, iz);
  const int pos = lower_bound_bucketed(codes, bucket_start, cq);
  const unsigned long long xl = pos > 0 ? (cq ^ __ldg(codes + pos - 1)) : ~0ull, xr = pos < n ? (cq ^ __ldg(codes + pos)) : ~0ull;
  const int bl = pos > 0 ? (xl ? (__clzll((long long)xl) - (64 - CODE_BITS)) : CODE_BITS) : 0;
  const int br = pos < n ? (xr ? (__clzll((long long)xr) - (64 - CODE_BITS)) : CODE_BITS) : 0;
  int L = min(max(bl, br) / 3, MORTON_BITS);
  const double rel[3] = {(double)q.x - qn.mn[0], (double)q.y - qn.mn[1], (double)q.z - qn.mn[2]};
  unsigned long long best = KEY_INF;
  #pragma unroll 1
  for (;; L--) {
    const int sh = MORTON_BITS - L;
    const int cx = (int)(ix >> sh), cy = (int)(iy >> sh), cz = (int)(iz >> sh);
    const int ncell = 1 << L;
    int lo = 0, cnt = 0;
    if (lane < 27) {
      const int nx = cx + lane % 3 - 1, ny = cy + (lane / 3) % 3 - 1, nz = cz + lane / 9 - 1;
      if (nx >= 0 && ny >= 0 && nz >= 0 && nx < ncell && ny < ncell && nz < ncell) {
        const unsigned long long pre = morton3((unsigned)nx, (unsigned)ny, (unsigned)nz);
        lo = lower_bound_bucketed(codes, bucket_start, pre << (3 * sh));
        const int hi = (L == 0) ? n : lower_bound_bucketed(codes, bucket_start, (pre + 1ull) << (3 * sh));
        cnt = hi - lo;
      }
    }
    int incl = cnt;
    #pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += v; }
    const int total = __shfl_sync(FULL, incl, 31);
    const int excl = incl - cnt;
    best = KEY_INF;
    #pragma unroll 1
    for (int base = 0; base < total; base += 32) {
      const int g = base + lane;
      int a = 0, b = 26;
      #pragma unroll
      for (int s = 0; s < 5; s++) {
        const int mid = (a + b) >> 1;
        const int v = __shfl_sync(FULL, incl, mid);
        if (v <= g) a = mid + 1; else b = mid;
      }
      a = min(a, 26);
      const int cell_lo = __shfl_sync(FULL, lo, a), cell_ex = __shfl_sync(FULL, excl, a);
      if (g < total) {
        const float4 p = __ldg(sorted + cell_lo + (g - cell_ex));
        const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
        float d = __fmul_rn(dx, dx);
        d = __fadd_rn(d, __fmul_rn(dy, dy));
        d = __fadd_rn(d, __fmul_rn(dz, dz));
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
        if (key < best) best = key;
      }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(FULL, best, o); if (other < best) best = other; }
    if (L == 0) break;
    if (best != KEY_INF) {
      const double sL = qn.cell * (double)(1 << sh);
      double rs = 1e300;
      const int cc[3] = {cx, cy, cz};
      #pragma unroll
      for (int dd = 0; dd < 3; dd++) {
        if (cc[dd] - 1 >= 0) rs = fmin(rs, rel[dd] - (double)(cc[dd] - 1) * sL);
        if (cc[dd] + 1 < ncell) rs = fmin(rs, (double)(cc[dd] + 2) * sL - rel[dd]);
      }
      if (rs > 0.0 && (double)__uint_as_float((unsigned)(best >> 32)) <= rs * rs * (1.0 - 1e-6)) break;
    }
  }
  if (lane == 0) {
    out_idx[j] = best == KEY_INF ? -1 : (int)(unsigned)(best & 0xFFFFFFFFull);
    out_d2[j] = best == KEY_INF ? CUDART_INF_F : __uint_as_float((unsigned)(best >> 32));
  }
}

// generic fallback (32 < k <= 64): one thread per query, local-memory top-k
__global__ void __launch_bounds__(128)
knn_thread_kernel(const float4* __restrict__ sorted, const unsigned long long* __restrict__ codes, int n, int k,
                  const int* __restrict__ bbox, int* __restrict__ knn_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const Quant qn = make_quant(bbox);
  const float4 q = sorted[j];
  const int self = __float_as_int(q.w);
  unsigned ix, iy, iz;
  quantize(qn, q.x, q.y, q.z, ix, iy, iz);
  const double rel[3] = {(double)q.x - qn.mn[0], (double)q.y - qn.mn[1], (double)q.z - qn.mn[2]};
  int L;
  {
    const unsigned long long cq = codes[j];
    const unsigned long long xl = cq ^ codes[max(j - 4, 0)], xr = cq ^ codes[min(j + 4, n - 1)];
    const int bl = (j - 4 >= 0) ? (xl ? (__clzll((long long)xl) - (64 - CODE_BITS)) : CODE_BITS) : 0;
    const int br = (j + 4 < n) ? (xr ? (__clzll((long long)xr) - (64 - CODE_BITS)) : CODE_BITS) : 0;
    L = min(max(bl, br) / 3, MORTON_BITS);
  }
  TopKDynamic top;
  #pragma unroll 1
  for (;; L--) {
    top.reset(k);
    const int sh = MORTON_BITS - L;
    const int cx = (int)(ix >> sh), cy = (int)(iy >> sh), cz = (int)(iz >> sh);
    const int ncell = 1 << L;
    #pragma unroll 1
    for (int o = 0; o < 27; o++) {
      const int nx = cx + o % 3 - 1, ny = cy + (o / 3) % 3 - 1, nz = cz + o / 9 - 1;
      if (nx < 0 || ny < 0 || nz < 0 || nx >= ncell || ny >= ncell || nz >= ncell) continue;
      const unsigned long long pre = morton3((unsigned)nx, (unsigned)ny, (unsigned)nz);
      const int lo = lower_bound_u64(codes, n, pre << (3 * sh));
      const int hi = (L == 0) ? n : lower_bound_u64(codes, n, (pre + 1ull) << (3 * sh));
      for (int c = lo; c < hi; c++) {
        const float4 p = __ldg(sorted + c);
        const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
        float d = __fmul_rn(dx, dx);
        d = __fadd_rn(d, __fmul_rn(dy, dy));
        d = __fadd_rn(d, __fmul_rn(dz, dz));
        const int pid = __float_as_int(p.w);
        if (top.accepts(d, pid, k)) top.insert(d, pid, k);
      }
    }
    if (L == 0) break;
    const double sL = qn.cell * (double)(1 << sh);
    double rs = 1e300;
    const int cc[3] = {cx, cy, cz};
    #pragma unroll
    for (int d = 0; d < 3; d++) {
      if (cc[d] - 1 >= 0) rs = fmin(rs, rel[d] - (double)(cc[d] - 1) * sL);
      if (cc[d] + 1 < ncell) rs = fmin(rs, (double)(cc[d] + 2) * sL - rel[d]);
    }
    if (top.full(k) && (double)top.worst(k) <= rs * rs * (1.0 - 1e-6)) break;
  }
  for (int t = 0; t < k; t++) knn_out[(size_t)self * k + t] = top.get(t);
}

// mean-centred scatter / k over the neighbours in (distance, index) order, then the regulariser
// (rot_vgicp_impl.hpp:438-492). One thread per point.
__global__ void __launch_bounds__(128)
cov_from_knn_kernel(const float4* __restrict__ pts, const int* __restrict__ knn, int n, int k, int method, float* __restrict__ cov6) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* row = knn + (size_t)i * k;
  double sx = 0, sy = 0, sz = 0;
  for (int t = 0; t < k; t++) { const float4 p = __ldg(pts + __ldg(row + t)); sx += (double)p.x; sy += (double)p.y; sz += (double)p.z; }
  const double mx = sx / k, my = sy / k, mz = sz / k;
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int t = 0; t < k; t++) {
    const float4 p = __ldg(pts + __ldg(row + t));
    const double dx = (double)p.x - mx, dy = (double)p.y - my, dz = (double)p.z - mz;
    c[0] += dx * dx; c[1] += dx * dy; c[2] += dx * dz; c[3] += dy * dy; c[4] += dy * dz; c[5] += dz * dz;
  }
  #pragma unroll
  for (int t = 0; t < 6; t++) c[t] /= (double)k;
  double out[6];
  regularize(c, method, out);
  const size_t base = (size_t)(i / TILE_PTS) * COV_PLANES * TILE_PTS + (i % TILE_PTS);
  #pragma unroll
  for (int t = 0; t < 6; t++) cov6[base + (size_t)t * TILE_PTS] = (float)out[t];
}

}  // namespace

size_t cov_bucket_table_entries() { return (size_t)N_BUCKETS + 1; }

static int cloud_bits(int n_clouds) { int b = 0; while ((1 << b) < n_clouds) b++; return b; }

size_t cov_cub_temp_bytes(int n, int n_clouds) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, n, 0, CODE_BITS + cloud_bits(n_clouds));
  return bytes;
}

// bounding box -> Morton codes -> radix sort -> Morton-ordered copy (+ bucket table) of ONE cloud: the spatial index of the
// scan-to-map 5-NN (scan2map.cu), the fitness score and the generic k > 32 path
cudaError_t launch_sort_cloud(const float4* pts, int n, const CovScratch& s, bool bucket_table, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int tb = 256, nb = (n + tb - 1) / tb;
  CloudRef ref; ref.arr = nullptr; ref.one.pts = pts; ref.one.cov6 = nullptr; ref.one.knn_out = nullptr; ref.one.n = n; ref.one.off = 0; ref.one.chunk_off = 0; ref.one.shard_rank = 0; ref.one.shard_world = 1;
  bbox_init_kernel<<<1, 32, 0, stream>>>(s.bbox, 1);
  bbox_kernel<<<min(nb, 148), tb, 0, stream>>>(ref, s.bbox);
  morton_kernel<<<nb, tb, 0, stream>>>(ref, s.bbox, s.codes_in, s.idx_in);
  size_t temp = s.cub_temp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(s.cub_temp, temp, s.codes_in, s.codes_out, s.idx_in, s.idx_out, n, 0, CODE_BITS, stream);
  if (e != cudaSuccess) return e;
  gather_sorted_kernel<<<nb, tb, 0, stream>>>(ref, s.idx_out, s.codes_out, s.sorted, nullptr);
  if (bucket_table) bucket_table_kernel<<<(N_BUCKETS + 1 + tb - 1) / tb, tb, 0, stream>>>(ref, s.codes_out, s.bucket_start);
  return cudaGetLastError();
}

cudaError_t launch_nearest(const float4* queries, int nq, const float* T12_dev, const CovScratch& s, int n, int* out_idx, float* out_d2, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  if (n <= 0) return cudaErrorInvalidValue;
  nn1_warp_kernel<<<(unsigned)(((long long)nq * 32 + 255) / 256), 256, 0, stream>>>(queries, nq, T12_dev, s.sorted, s.codes_out, s.bucket_start, n, s.bbox, out_idx, out_d2);
  return cudaGetLastError();
}

cudaError_t launch_covariances_generic(const CovLaunch& L, cudaStream_t stream) {
  if (L.n <= 0) return cudaSuccess;
  if (L.k > KCAP || L.k < 1 || L.n < L.k || !L.knn) return cudaErrorInvalidValue;
  const int n = L.n;
  cudaError_t e = launch_sort_cloud(L.pts, n, L.s, false, stream);
  if (e != cudaSuccess) return e;
  knn_thread_kernel<<<(n + 127) / 128, 128, 0, stream>>>(L.s.sorted, L.s.codes_out, n, L.k, L.s.bbox, L.knn);
  if (L.cov6) cov_from_knn_kernel<<<(n + 127) / 128, 128, 0, stream>>>(L.pts, L.knn, n, L.k, L.method, L.cov6);
  return cudaGetLastError();
}

template <int KC>
static cudaError_t launch_chunk(const KnnGroup& g, int total_chunks, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(knn_chunk_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KNN_SMEM);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::max(1, std::min(3 * sms, (total_chunks + KNN_WARPS - 1) / KNN_WARPS));   // 3 resident CTAs per SM
  knn_chunk_kernel<KC><<<grid, KNN_WARPS * 32, KNN_SMEM, stream>>>(g.s.clouds, g.n_clouds, total_chunks, g.s.sorted, g.s.codes_a, g.s.bucket_start, g.s.bbox,
                                                                    g.s.fail, g.k, g.method, g.stats);
  return cudaGetLastError();
}

cudaError_t launch_covariances_group(const KnnGroup& g, cudaStream_t stream, int* own_launches) {
  if (own_launches) *own_launches = 0;
  if (g.n_clouds <= 0 || g.total <= 0) return cudaSuccess;
  if (g.k < 1 || g.k > 32 || g.n_clouds > KNN_MAX_CLOUDS) return cudaErrorInvalidValue;
  for (int c = 0; c < g.n_clouds; c++) if (g.clouds[c].n < g.k) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemcpyAsync(g.s.clouds, g.clouds, sizeof(KnnCloud) * g.n_clouds, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(g.s.fail, 0, 2 * sizeof(int), stream)) != cudaSuccess) return e;
  const int tb = 256;
  const unsigned nc = (unsigned)g.n_clouds;
  const unsigned nb = (unsigned)((g.max_n + tb - 1) / tb);
  CloudRef ref; ref.arr = g.s.clouds; ref.one = g.clouds[0];
  bbox_init_kernel<<<(g.n_clouds * 8 + tb - 1) / tb, tb, 0, stream>>>(g.s.bbox, g.n_clouds);
  bbox_kernel<<<dim3(max(1u, min(nb, 1184u / nc)), nc), tb, 0, stream>>>(ref, g.s.bbox);
  morton_kernel<<<dim3(nb, nc), tb, 0, stream>>>(ref, g.s.bbox, g.s.codes_a, g.s.idx_a);
  size_t temp = g.s.cub_temp_bytes;
  e = cub::DeviceRadixSort::SortPairs(g.s.cub_temp, temp, g.s.codes_a, g.s.codes_b, g.s.idx_a, g.s.idx_b, g.total, 0, CODE_BITS + cloud_bits(g.n_clouds), stream);
  if (e != cudaSuccess) return e;
  gather_sorted_kernel<<<dim3(nb, nc), tb, 0, stream>>>(ref, g.s.idx_b, g.s.codes_b, g.s.sorted, g.s.codes_a);
  bucket_table_kernel<<<dim3((N_BUCKETS + 1 + tb - 1) / tb, nc), tb, 0, stream>>>(ref, g.s.codes_a, g.s.bucket_start);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (g.ev_knn0) cudaEventRecord(g.ev_knn0, stream);
  int total_chunks = 0;
  for (int c = 0; c < g.n_clouds; c++) total_chunks += knn_owned_chunks(g.clouds[c].n, g.clouds[c].shard_rank, g.clouds[c].shard_world);   // == the chunk_off prefix the caller filled in
  if (g.k <= 8) e = launch_chunk<8>(g, total_chunks, stream);
  else if (g.k <= 12) e = launch_chunk<12>(g, total_chunks, stream);
  else if (g.k <= 16) e = launch_chunk<16>(g, total_chunks, stream);
  else if (g.k <= 20) e = launch_chunk<20>(g, total_chunks, stream);
  else if (g.k <= 24) e = launch_chunk<24>(g, total_chunks, stream);
  else e = launch_chunk<32>(g, total_chunks, stream);
  if (e != cudaSuccess) return e;
  knn_fallback_kernel<<<ROLO_KNN_FALLBACK_GRID, 256, 0, stream>>>(g.s.clouds, g.s.sorted, g.s.codes_a, g.s.bucket_start, g.s.bbox, g.s.fail, g.k, g.method);
  if (g.ev_knn1) cudaEventRecord(g.ev_knn1, stream);
  if (own_launches) *own_launches = 7;   // bbox_init, bbox, morton, gather, bucket table, chunk kNN + covariance, fallback (+ 1 memset not counted; cub's kernels not counted)
  return cudaGetLastError();
}

}  // namespace rolo
