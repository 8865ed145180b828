// covariance.cu — RotVGICP::calculate_covariances on the GPU (rot_vgicp_impl.hpp:421-496).
//
// Reference: per point, pcl::search::KdTree::nearestKSearch(k=20) (FLANN, float L2), mean-centre the
// 4 x k neighbour matrix, cov = N N^T / k, regularise through JacobiSVD (PLANE: U diag(1,1,1e-3) V^T).
// Here (exact, no approximation of the neighbour set):
//   1. bounding box, 39-bit Morton codes (13 bits/axis: 5 radix passes), radix sort of (code, index)  [cub::DeviceRadixSort
//      is the only library call: plumbing, not the hot op]
//   2. one thread per query walks an implicit octree over the sorted codes: at level L the 3x3x3 block
//      of level-L cells around the query is scanned (each cell = one contiguous range found by binary
//      search); the result is exact once the k-th distance is below the distance to the block
//      boundary, otherwise the level is coarsened. The start level comes from the common prefix of
//      the query's Morton neighbours, so dense (near) and sparse (far) regions both scan O(100) points.
//   3. neighbours ordered by (float distance, index) exactly like the oracle; covariance in fp64;
//      symmetric 3x3 eigen-decomposition by cyclic Jacobi (fp64) gives the SVD of the PSD matrix.
// Float distances are accumulated x,y,z without FMA contraction so the neighbour sets match the
// CPU oracle bit for bit.

#include "covariance.cuh"
#include "morton.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <float.h>
#include <math_constants.h>
#include <limits.h>

namespace rolo {

namespace {

__global__ void bbox_init_kernel(int* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = INT_MIN;
}

__global__ void bbox_kernel(const float4* __restrict__ pts, int n, int* bbox) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
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

__global__ void morton_kernel(const float4* __restrict__ pts, int n, const int* __restrict__ bbox,
                              unsigned long long* __restrict__ codes, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Quant q = make_quant(bbox);
  const float4 p = pts[i];
  unsigned ix, iy, iz;
  quantize(q, p.x, p.y, p.z, ix, iy, iz);
  codes[i] = morton3(ix, iy, iz);
  idx[i] = i;
}

// Morton-ordered copy of the points (w = original index) + the packed cell coordinates (ix | iy << 21 | iz << 42) of each
// sorted point: the kNN kernel needs them per query, and de-interleaving the code there would cost ~100 warp-instructions
// per query instead of ~100 per 32 points here. `cells` reuses the unsorted-code buffer, which is dead after the sort.
__global__ void gather_sorted_kernel(const float4* __restrict__ pts, const int* __restrict__ idx, int n, const int* __restrict__ bbox,
                                     float4* __restrict__ sorted, unsigned long long* __restrict__ cells) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int i = idx[j];
  float4 p = pts[i];
  const Quant q = make_quant(bbox);
  unsigned ix, iy, iz;
  quantize(q, p.x, p.y, p.z, ix, iy, iz);          // identical to morton_kernel's
  cells[j] = (unsigned long long)ix | ((unsigned long long)iy << 21) | ((unsigned long long)iz << 42);
  p.w = __int_as_float(i);
  sorted[j] = p;
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


// ---------------------------------------------------------------------------------------------------
// exact kNN, one WARP per query (k <= 32).
//  * lanes 0..26 each locate one cell of the 3x3x3 block (two binary searches in parallel instead of 54
//    in sequence); a warp scan of the cell counts gives the candidate total, which also steers the level:
//    too many candidates -> finer level, fewer than k -> coarser, before anything is scanned;
//  * candidates are scanned 32 at a time (coalesced), the current top-32 lives one element per lane, sorted
//    by (distance, index); a batch in which nobody beats the k-th distance costs one ballot; otherwise the
//    winners are inserted by shuffle (few) or the batch is bitonic-sorted and merged (many).
// ---------------------------------------------------------------------------------------------------
constexpr int KNN_QPW = 4;            // queries handled sequentially by one warp
constexpr int KNN_CAND_CAP = 1024;    // prefer a finer level when the block holds more candidates than this
// bucket_start[b] = first sorted position whose code has top bits >= b (b = 0 .. N_BUCKETS): every later
// lower_bound only searches inside one bucket
__global__ void bucket_table_kernel(const unsigned long long* __restrict__ codes, int n, int* __restrict__ bucket_start) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > N_BUCKETS) return;
  bucket_start[b] = (b == N_BUCKETS) ? n : lower_bound_u64(codes, n, (unsigned long long)b << BUCKET_SHIFT);
}

__global__ void __launch_bounds__(256, 5)        // 48 registers: 40 warps/SM measured 4 % faster than 32
knn_warp_kernel(const float4* __restrict__ sorted, const unsigned long long* __restrict__ codes, const unsigned long long* __restrict__ cells,
                const int* __restrict__ bucket_start, int n, int k, const int* __restrict__ bbox, int* __restrict__ knn_out) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const Quant qn = make_quant(bbox);

  // ranges of the previous query of this warp: consecutive Morton neighbours usually share the block
  int prevL = -1, prev_cx = 0, prev_cy = 0, prev_cz = 0, lo = 0, cnt = 0, incl = 0, total = 0;

  #pragma unroll 1
  for (int qq = 0; qq < KNN_QPW; qq++) {
    const int j = warp_global * KNN_QPW + qq;
    if (j >= n) break;                                   // warp-uniform
    const float4 q = __ldg(sorted + j);
    const int self = __float_as_int(q.w);
    const unsigned long long cq = __ldg(codes + j);
    const unsigned long long cc3 = __ldg(cells + j);
    const unsigned ix = (unsigned)(cc3 & 0x1FFFFFull), iy = (unsigned)((cc3 >> 21) & 0x1FFFFFull), iz = (unsigned)(cc3 >> 42);

    // start level: reuse the previous query's final level when this query lies in the same cell there; otherwise
    // the level-cpl(code[j], code[j-4]) cell holds the sorted entries j-4..j (likewise on the right), so the
    // larger of the two common prefixes is a boundary-robust first guess
    int L;
    bool reuse = false;
    if (prevL >= 0 && (int)(ix >> (MORTON_BITS - prevL)) == prev_cx && (int)(iy >> (MORTON_BITS - prevL)) == prev_cy &&
        (int)(iz >> (MORTON_BITS - prevL)) == prev_cz) {
      L = prevL; reuse = true;
    } else {
      const unsigned long long xl = cq ^ __ldg(codes + max(j - 4, 0)), xr = cq ^ __ldg(codes + min(j + 4, n - 1));
      const int bl = (j - 4 >= 0) ? (xl ? (__clzll((long long)xl) - (64 - CODE_BITS)) : CODE_BITS) : 0;
      const int br = (j + 4 < n) ? (xr ? (__clzll((long long)xr) - (64 - CODE_BITS)) : CODE_BITS) : 0;
      L = min(max(bl, br) / 3, MORTON_BITS);
    }

    unsigned long long top = KEY_INF;                   // this lane's element of the sorted top-32
    bool no_finer = reuse;
    #pragma unroll 1
    for (;;) {
      const int sh = MORTON_BITS - L;
      const int cx = (int)(ix >> sh), cy = (int)(iy >> sh), cz = (int)(iz >> sh);
      const int ncell = 1 << L;
      if (!reuse) {
        lo = 0; cnt = 0;
        if (lane < 27) {
          const int oo = (lane == 0) ? 13 : (lane <= 13 ? lane - 1 : lane);   // centre cell first
          const int nx = cx + oo % 3 - 1, ny = cy + (oo / 3) % 3 - 1, nz = cz + oo / 9 - 1;
          if (nx >= 0 && ny >= 0 && nz >= 0 && nx < ncell && ny < ncell && nz < ncell) {
            const unsigned long long pre = morton3((unsigned)nx, (unsigned)ny, (unsigned)nz);
            lo = lower_bound_bucketed(codes, bucket_start, pre << (3 * sh));
            const int hi = (L == 0) ? n : lower_bound_bucketed(codes, bucket_start, (pre + 1ull) << (3 * sh));
            cnt = hi - lo;
          }
        }
        incl = cnt;
        #pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += v; }
        total = __shfl_sync(FULL, incl, 31);
        if (!no_finer && total > KNN_CAND_CAP && L < MORTON_BITS) { L++; continue; }
        if (total < k && L > 0) { no_finer = true; L--; continue; }
      }
      reuse = false;
      prevL = L; prev_cx = cx; prev_cy = cy; prev_cz = cz;
      const int excl = incl - cnt;

      top = KEY_INF;
      #pragma unroll 1
      for (int base = 0; base < total; base += 32) {
        const int g = base + lane;
        const bool valid = g < total;
        // owning cell: the first lane whose inclusive count exceeds g
        int a = 0, b = 26;
        #pragma unroll
        for (int s = 0; s < 5; s++) {
          const int mid = (a + b) >> 1;
          const int v = __shfl_sync(FULL, incl, mid);
          if (v <= g) a = mid + 1; else b = mid;
        }
        a = min(a, 26);
        const int cell_lo = __shfl_sync(FULL, lo, a), cell_ex = __shfl_sync(FULL, excl, a);
        unsigned long long key = KEY_INF;
        if (valid) {
          const float4 p = __ldg(sorted + cell_lo + (g - cell_ex));
          const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
          float d = __fmul_rn(dx, dx);
          d = __fadd_rn(d, __fmul_rn(dy, dy));
          d = __fadd_rn(d, __fmul_rn(dz, dz));
          key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
        }
        const unsigned long long kth = __shfl_sync(FULL, top, k - 1);
        const bool pass = key < kth;
        unsigned m = __ballot_sync(FULL, pass);
        if (m == 0) continue;
        if (__popc(m) > 10) {
          // bitonic sort of the batch (non-passing lanes carry +inf), then bitonic merge with the top-32
          if (!pass) key = KEY_INF;
          #pragma unroll
          for (int k2 = 2; k2 <= 32; k2 <<= 1) {
            #pragma unroll
            for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) bitonic_cmpx(key, j2, ((lane & k2) == 0) == ((lane & j2) == 0));
          }
          if (base == 0) {
            top = key;                                   // the top-32 is still empty: the sorted batch is the new top-32
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
      if (L == 0) break;
      // exact iff the k-th neighbour lies inside the guaranteed-complete radius of the scanned block
      const double sL = qn.cell * (double)(1 << sh);
      const double rel[3] = {(double)q.x - qn.mn[0], (double)q.y - qn.mn[1], (double)q.z - qn.mn[2]};
      double rs = 1e300;
      const int cc[3] = {cx, cy, cz};
      #pragma unroll
      for (int dd = 0; dd < 3; dd++) {
        if (cc[dd] - 1 >= 0) rs = fmin(rs, rel[dd] - (double)(cc[dd] - 1) * sL);
        if (cc[dd] + 1 < ncell) rs = fmin(rs, (double)(cc[dd] + 2) * sL - rel[dd]);
      }
      const unsigned long long kth = __shfl_sync(FULL, top, k - 1);
      if (kth != KEY_INF && (double)__uint_as_float((unsigned)(kth >> 32)) <= rs * rs * (1.0 - 1e-6)) break;
      no_finer = true;
      L--;
    }
    if (lane < k) knn_out[(size_t)self * k + lane] = (int)(unsigned)(top & 0xFFFFFFFFull);
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

size_t cov_cub_temp_bytes(int n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, n, 0, 3 * MORTON_BITS);
  return bytes;
}

// bounding box -> Morton codes -> radix sort -> Morton-ordered copy (+ bucket table): the spatial index shared by the
// covariance kNN and the scan-to-map 5-NN (scan2map.cu)
cudaError_t launch_sort_cloud(const float4* pts, int n, const CovScratch& s, bool bucket_table, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int tb = 256, nb = (n + tb - 1) / tb;
  bbox_init_kernel<<<1, 32, 0, stream>>>(s.bbox);
  bbox_kernel<<<min(nb, 148), tb, 0, stream>>>(pts, n, s.bbox);
  morton_kernel<<<nb, tb, 0, stream>>>(pts, n, s.bbox, s.codes_in, s.idx_in);
  size_t temp = s.cub_temp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(s.cub_temp, temp, s.codes_in, s.codes_out, s.idx_in, s.idx_out, n, 0, 3 * MORTON_BITS, stream);
  if (e != cudaSuccess) return e;
  gather_sorted_kernel<<<nb, tb, 0, stream>>>(pts, s.idx_out, n, s.bbox, s.sorted, s.codes_in);   // codes_in now holds the packed cell coordinates
  if (bucket_table) bucket_table_kernel<<<(N_BUCKETS + 1 + tb - 1) / tb, tb, 0, stream>>>(s.codes_out, n, s.bucket_start);
  return cudaGetLastError();
}

cudaError_t launch_nearest(const float4* queries, int nq, const float* T12_dev, const CovScratch& s, int n, int* out_idx, float* out_d2, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  if (n <= 0) return cudaErrorInvalidValue;
  nn1_warp_kernel<<<(unsigned)(((long long)nq * 32 + 255) / 256), 256, 0, stream>>>(queries, nq, T12_dev, s.sorted, s.codes_out, s.bucket_start, n, s.bbox, out_idx, out_d2);
  return cudaGetLastError();
}

cudaError_t launch_covariances(const CovLaunch& L, cudaStream_t stream) {
  if (L.n <= 0) return cudaSuccess;
  if (L.k > KCAP || L.k < 1 || L.n < L.k || !L.knn) return cudaErrorInvalidValue;
  const int n = L.n;
  cudaError_t e = launch_sort_cloud(L.pts, n, L.s, L.k <= 32, stream);
  if (e != cudaSuccess) return e;
  if (L.k <= 32) {
    const int queries_per_block = (256 / 32) * KNN_QPW;
    const int gb = (n + queries_per_block - 1) / queries_per_block;
    knn_warp_kernel<<<gb, 256, 0, stream>>>(L.s.sorted, L.s.codes_out, L.s.codes_in, L.s.bucket_start, n, L.k, L.s.bbox, L.knn);
  } else {
    knn_thread_kernel<<<(n + 127) / 128, 128, 0, stream>>>(L.s.sorted, L.s.codes_out, n, L.k, L.s.bbox, L.knn);
  }
  if (L.cov6) cov_from_knn_kernel<<<(n + 127) / 128, 128, 0, stream>>>(L.pts, L.knn, n, L.k, L.method, L.cov6);
  return cudaGetLastError();
}

}  // namespace rolo
