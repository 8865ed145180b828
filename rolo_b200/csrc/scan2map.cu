// scan2map.cu — backMapping's cornerOptimization + surfOptimization + combineOptimizationCoeffs + the matA/matB assembly of
// LMOptimization (src/backMapping.cpp:719-1006): two kernels per Gauss-Newton iteration.
//
// Reference: per scan feature point, transform to the map frame, kd-tree 5-NN in the corner / surface map, accept when the
// 5th neighbour is closer than 1 m, then either a PCA line test + point-to-line coefficient (corners, cv::eigen) or a plane
// fit + point-to-plane coefficient (surfaces, Eigen QR); the selected points are compacted, a Jacobian row is built per
// point and AtA (6x6) / AtB (6) are formed with cv::gemm.
// Here: two kernels per iteration — a warp per scan point for the 5-NN, then a thread per scan point for the fit.
//   * 5-NN: both maps carry the Morton index of covariance.cu. Only neighbours within 1 m matter, so the search is a
//     fixed-level 3x3x3 block scan at the level whose cells are >= 1.001 m: every point within 1 m of the query lies in
//     that block, hence the five nearest are exact whenever the reference would accept the point, and a point the
//     reference rejects (5th neighbour >= 1 m) finds fewer than five candidates. Lanes 0..26 locate the 27 cell ranges
//     with two bucketed binary searches each; candidates are scanned 32 per step and merged into a sorted top list held
//     one element per lane, ordered by (float distance, index) like the oracle's kd-tree.
//     The block one level finer is tried first (a quarter of the candidates) and kept when its 5th distance is inside
//     the radius that block provably covers.
//   * the per-point algebra (3x3 Jacobi eigen / 5x3 pivoted Householder QR, s2m_math.cuh) runs one thread per point,
//     float, no FMA contraction (-fmad=false for this file): results are bit-equal to the CPU oracle's.
//   * no compaction: unselected points contribute zero rows; the 21 + 6 + 1 sums (AtA upper triangle, AtB, count) are
//     formed in double: warp shuffle tree -> block -> per-block partial -> the last block adds the partials in block
//     order, so the result is deterministic.
// HBM traffic is a few MB per iteration (the maps are L2-resident): the 5-NN kernel is instruction-issue bound, the fit kernel
// latency bound (profiles/r01_ncu_s2m_v2.txt).

#include "scan2map.cuh"
#include "morton.cuh"
#include "s2m_math.cuh"

#include <algorithm>

namespace rolo {

namespace {

constexpr int S2M_THREADS = 256;
constexpr int S2M_WARPS = S2M_THREADS / 32;

struct Params {
  S2MMap map[2];
  const float4* scan[2];
  int n_scan[2];
  float t[12];
  float sc[6];
  float4* coeff[2];
  unsigned char* flag[2];
  int* nn;
  double* partial;
  unsigned* counter;
  double* result;
};

// shift (MORTON_BITS - level) of the coarsest-resolution level whose cells are still >= 1.001 m
__device__ __forceinline__ int search_shift(const Quant& qn) {
  int sh = 0;
  while (sh < MORTON_BITS && qn.cell * (double)(1 << sh) < 1.001) sh++;
  return sh;
}

// one 3x3x3 block scan at shift `sh`: merges every candidate closer than 1 m into the sorted top list (one element per lane)
__device__ __forceinline__ void scan_block(const S2MMap& m, int sh, unsigned ix, unsigned iy, unsigned iz, float qx, float qy, float qz, int lane,
                                           unsigned long long& top) {
  const unsigned FULL = 0xffffffffu;
  const int cx = (int)(ix >> sh), cy = (int)(iy >> sh), cz = (int)(iz >> sh);
  const int L = MORTON_BITS - sh, ncell = 1 << L;
  int lo = 0, cnt = 0;
  if (lane < 27) {
    const int nx = cx + lane % 3 - 1, ny = cy + (lane / 3) % 3 - 1, nz = cz + lane / 9 - 1;
    if (nx >= 0 && ny >= 0 && nz >= 0 && nx < ncell && ny < ncell && nz < ncell) {
      const unsigned long long pre = morton3((unsigned)nx, (unsigned)ny, (unsigned)nz);
      lo = lower_bound_bucketed(m.codes, m.bucket_start, pre << (3 * sh));
      const int hi = (L == 0) ? m.n : lower_bound_bucketed(m.codes, m.bucket_start, (pre + 1ull) << (3 * sh));
      cnt = hi - lo;
    }
  }
  int incl = cnt;
  #pragma unroll
  for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += v; }
  const int total = __shfl_sync(FULL, incl, 31);
  const int excl = incl - cnt;
  top = KEY_INF;
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
    unsigned long long key = KEY_INF;
    if (g < total) {
      const float4 p = __ldg(m.sorted + cell_lo + (g - cell_ex));
      const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
      float d = dx * dx; d = d + dy * dy; d = d + dz * dz;             // no FMA in this translation unit
      if (d < 1.0f) key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
    }
    const unsigned long long kth = __shfl_sync(FULL, top, 4);
    unsigned msk = __ballot_sync(FULL, key < kth);
    while (msk) {                                                       // few candidates survive the 5th-distance filter
      const int srcl = __ffs(msk) - 1;
      msk &= msk - 1;
      const unsigned long long ck = __shfl_sync(FULL, key, srcl);
      const int pos = __popc(__ballot_sync(FULL, top < ck));
      const unsigned long long up = __shfl_up_sync(FULL, top, 1);
      if (lane > pos) top = up; else if (lane == pos) top = ck;
    }
  }
}

// exact five nearest map points within 1 m of the query (one warp). `sh` is the level whose cells are >= 1.001 m: its
// 3x3x3 block holds every point within 1 m. The block one level finer (a quarter of the candidates on a surface) is tried
// first and accepted when its 5th distance lies inside the radius that block is guaranteed to cover.
__device__ __forceinline__ bool five_nearest(const S2MMap& m, const Quant& qn, int sh, float qx, float qy, float qz, int lane, int idx[5]) {
  const unsigned FULL = 0xffffffffu;
  unsigned ix, iy, iz;
  quantize(qn, qx, qy, qz, ix, iy, iz);
  unsigned long long top = KEY_INF;
  bool done = false;
  if (sh > 0) {
    const int sf = sh - 1;
    scan_block(m, sf, ix, iy, iz, qx, qy, qz, lane, top);
    const unsigned long long k5 = __shfl_sync(FULL, top, 4);
    if (k5 != KEY_INF) {
      const double sL = qn.cell * (double)(1 << sf);
      const int ncell = 1 << (MORTON_BITS - sf);
      const double rel[3] = {(double)qx - qn.mn[0], (double)qy - qn.mn[1], (double)qz - qn.mn[2]};
      const int cc[3] = {(int)(ix >> sf), (int)(iy >> sf), (int)(iz >> sf)};
      double rs = 1e300;
      #pragma unroll
      for (int dd = 0; dd < 3; dd++) {
        if (cc[dd] - 1 >= 0) rs = fmin(rs, rel[dd] - (double)(cc[dd] - 1) * sL);
        if (cc[dd] + 1 < ncell) rs = fmin(rs, (double)(cc[dd] + 2) * sL - rel[dd]);
      }
      done = rs > 0.0 && (double)__uint_as_float((unsigned)(k5 >> 32)) <= rs * rs * (1.0 - 1e-6);
    }
  }
  if (!done) scan_block(m, sh, ix, iy, iz, qx, qy, qz, lane, top);
  const bool ok = __shfl_sync(FULL, top, 4) != KEY_INF;
  #pragma unroll
  for (int j = 0; j < 5; j++) idx[j] = (int)(unsigned)(__shfl_sync(FULL, top, j) & 0xFFFFFFFFull);
  return ok;
}

// backMapping.cpp:741-822 — coefficient of a corner point given its five neighbours; false when rejected
__device__ __forceinline__ bool corner_coeff(const float4 nb[5], float x0, float y0, float z0, float4& coeff) {
  float cx = 0, cy = 0, cz = 0;
  #pragma unroll
  for (int j = 0; j < 5; j++) { cx += nb[j].x; cy += nb[j].y; cz += nb[j].z; }
  cx /= 5; cy /= 5; cz /= 5;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
  #pragma unroll
  for (int j = 0; j < 5; j++) {
    const float ax = nb[j].x - cx, ay = nb[j].y - cy, az = nb[j].z - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az; a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  float A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33}, D[3], V[9];
  s2m::jacobi_eigen<3>(A, D, V);
  if (!(D[0] > 3 * D[1])) return false;
  // the reference writes `cx + 0.1 * v` with a double literal: the sum is formed in double and rounded once (:777-783)
  const float x1 = (float)((double)cx + 0.1 * (double)V[0]), y1 = (float)((double)cy + 0.1 * (double)V[1]), z1 = (float)((double)cz + 0.1 * (double)V[2]);
  const float x2 = (float)((double)cx - 0.1 * (double)V[0]), y2 = (float)((double)cy - 0.1 * (double)V[1]), z2 = (float)((double)cz - 0.1 * (double)V[2]);
  const float m11 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
  const float m22 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
  const float m33 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
  const float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);
  const float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
  const float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
  const float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
  const float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
  const float ld2 = a012 / l12;
  const float s = (float)(1.0 - 0.9 * (double)fabsf(ld2));
  coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
  return (double)s > 0.1;
}

// backMapping.cpp:851-905 — coefficient of a surface point
__device__ __forceinline__ bool surf_coeff(const float4 nb[5], float4 ori, float sx, float sy, float sz, float4& coeff) {
  float A0[15], B0[5] = {-1.f, -1.f, -1.f, -1.f, -1.f}, X0[3];
  #pragma unroll
  for (int j = 0; j < 5; j++) { A0[j * 3] = nb[j].x; A0[j * 3 + 1] = nb[j].y; A0[j * 3 + 2] = nb[j].z; }
  s2m::qr_solve<5, 3, true>(A0, B0, X0);
  float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
  const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  bool valid = true;
  #pragma unroll
  for (int j = 0; j < 5; j++)
    if ((double)fabsf(pa * nb[j].x + pb * nb[j].y + pc * nb[j].z + pd) > 0.2) valid = false;
  if (!valid) return false;
  const float pd2 = pa * sx + pb * sy + pc * sz + pd;
  const float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(ori.x * ori.x + ori.y * ori.y + ori.z * ori.z)));
  coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
  return (double)s > 0.1;
}

// stage 1, one WARP per scan point: pointAssociateToMap + nearestKSearch(5) -> five neighbour indices, or -1
__global__ void __launch_bounds__(S2M_THREADS)
s2m_knn_kernel(const __grid_constant__ Params P) {
  const int lane = threadIdx.x & 31;
  const int q = (int)((blockIdx.x * (unsigned)S2M_THREADS + threadIdx.x) >> 5);
  const int n0 = P.n_scan[0], ntot = n0 + P.n_scan[1];
  if (q >= ntot) return;
  const int which = q >= n0 ? 1 : 0, i = q - which * n0;
  int idx[5] = {-1, -1, -1, -1, -1};
  if (P.map[which].n >= 5) {
    const Quant qn = make_quant(P.map[which].bbox);
    const int sh = search_shift(qn);
    const float4 ori = __ldg(P.scan[which] + i);
    const float sx = P.t[0] * ori.x + P.t[1] * ori.y + P.t[2] * ori.z + P.t[3];
    const float sy = P.t[4] * ori.x + P.t[5] * ori.y + P.t[6] * ori.z + P.t[7];
    const float sz = P.t[8] * ori.x + P.t[9] * ori.y + P.t[10] * ori.z + P.t[11];
    if (!five_nearest(P.map[which], qn, sh, sx, sy, sz, lane, idx)) idx[0] = -1;
  }
  if (lane < 5) {
    int v = idx[0];
    #pragma unroll
    for (int j = 1; j < 5; j++) if (lane == j) v = idx[0] < 0 ? -1 : idx[j];
    P.nn[(size_t)q * 5 + lane] = v;
  }
}

// stage 2, one THREAD per scan point: line / plane fit, coefficient, Jacobian row; 21 + 6 + 1 sums reduced warp -> block ->
// per-block partial -> last block (block order): deterministic
constexpr int S2M_FIT_THREADS = S2M_FIT_THREADS_PER_BLOCK;
__global__ void __launch_bounds__(S2M_FIT_THREADS)
s2m_fit_kernel(const __grid_constant__ Params P) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * S2M_FIT_THREADS + threadIdx.x;
  const int n0 = P.n_scan[0], ntot = n0 + P.n_scan[1];
  float row[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, bi = 0.f;
  bool sel = false;
  if (q < ntot) {
    const int which = q >= n0 ? 1 : 0, i = q - which * n0;
    const float4 ori = __ldg(P.scan[which] + i);
    // pointAssociateToMap (backMapping.cpp:293-299)
    const float sx = P.t[0] * ori.x + P.t[1] * ori.y + P.t[2] * ori.z + P.t[3];
    const float sy = P.t[4] * ori.x + P.t[5] * ori.y + P.t[6] * ori.z + P.t[7];
    const float sz = P.t[8] * ori.x + P.t[9] * ori.y + P.t[10] * ori.z + P.t[11];
    float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
    const int* nn = P.nn + (size_t)q * 5;
    if (__ldg(nn) >= 0) {
      float4 nb[5];
      #pragma unroll
      for (int j = 0; j < 5; j++) nb[j] = __ldg(P.map[which].pts + __ldg(nn + j));
      float4 cf;
      sel = which == 0 ? corner_coeff(nb, sx, sy, sz, cf) : surf_coeff(nb, ori, sx, sy, sz, cf);
      if (sel) coeff = cf;
    }
    P.coeff[which][i] = coeff; P.flag[which][i] = sel ? 1 : 0;
    if (sel) {
      // LMOptimization's Jacobian row (backMapping.cpp:968-1005): lidar -> camera axis swap, then back
      const float srx = P.sc[0], crx = P.sc[1], sry = P.sc[2], cry = P.sc[3], srz = P.sc[4], crz = P.sc[5];
      const float px = ori.y, py = ori.z, pz = ori.x;
      const float cx = coeff.y, cy = coeff.z, cz = coeff.x;
      const float arx = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * cx + (-srx * srz * px - crz * srx * py - crx * pz) * cy +
                        (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * cz;
      const float ary = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * cx +
                        ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * cz;
      const float arz = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * cx + (crx * crz * px - crx * srz * py) * cy +
                        ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * cz;
      row[0] = arz; row[1] = arx; row[2] = ary; row[3] = cz; row[4] = cx; row[5] = cy;
      bi = -coeff.w;
    }
  }
  __shared__ double s_acc[S2M_FIT_THREADS / 32][S2M_RES];
  __shared__ bool s_last;
  int t = 0;
  #pragma unroll
  for (int a = 0; a < 6; a++) {
    #pragma unroll
    for (int b = a; b < 6; b++, t++) {
      double v = (double)row[a] * (double)row[b];       // exact: float x float fits a double
      #pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      if (lane == 0) s_acc[warp][t] = v;
    }
  }
  #pragma unroll
  for (int a = 0; a < 6; a++) {
    double v = (double)row[a] * (double)bi;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0) s_acc[warp][21 + a] = v;
  }
  {
    const unsigned m = __ballot_sync(FULL, sel);
    if (lane == 0) { s_acc[warp][27] = (double)__popc(m); for (int k = 28; k < S2M_RES; k++) s_acc[warp][k] = 0.0; }
  }
  __syncthreads();
  if (threadIdx.x < S2M_RES) {
    double v = 0.0;
    #pragma unroll
    for (int w = 0; w < S2M_FIT_THREADS / 32; w++) v += s_acc[w][threadIdx.x];
    P.partial[(size_t)blockIdx.x * S2M_RES + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(P.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x < S2M_RES) {
      double v = 0.0;
      for (unsigned b = 0; b < gridDim.x; b++) v += __ldcg(P.partial + (size_t)b * S2M_RES + threadIdx.x);   // block order
      P.result[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *P.counter = 0u;
  }
}

}  // namespace

cudaError_t launch_scan2map(const S2MLaunch& L, cudaStream_t stream) {
  Params P;
  for (int w = 0; w < 2; w++) { P.map[w] = L.map[w]; P.scan[w] = L.scan[w]; P.n_scan[w] = L.n_scan[w]; P.coeff[w] = L.coeff[w]; P.flag[w] = L.flag[w]; }
  for (int k = 0; k < 12; k++) P.t[k] = L.t[k];
  for (int k = 0; k < 6; k++) P.sc[k] = L.sc[k];
  P.nn = L.nn; P.partial = L.partial; P.counter = L.counter; P.result = L.result;
  const long long ntot = (long long)L.n_scan[0] + L.n_scan[1];
  if (ntot > 0) s2m_knn_kernel<<<(unsigned)((ntot * 32 + S2M_THREADS - 1) / S2M_THREADS), S2M_THREADS, 0, stream>>>(P);
  const int fit_blocks = (int)std::max<long long>(1, (ntot + S2M_FIT_THREADS - 1) / S2M_FIT_THREADS);
    s2m_fit_kernel<<<fit_blocks, S2M_FIT_THREADS, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace rolo
