// linearize.cu — the fused per-Gauss-Newton-iteration kernel (sm_100a).
//
// One launch evaluates, for every source point of every scan pair in the batch:
//   p' = R p + t                                   (rot_vgicp_impl.hpp:184-185, 254)
//   voxel key of p' -> hash probe (+ neighbour offsets) (rot_vgicp_impl.hpp:186-193, vmp_voxel.hpp:199-233)
//   M = (B_v + Rm A Rm^T)^-1                        (rot_vgicp_impl.hpp:204-221; 3x3 block of the 4x4 form)
//   e = mu_v - p', w = sqrt(n_v), cost += w e^T M e (rot_vgicp_impl.hpp:256-259)
//   H += w J^T M J, b += w J^T M e, J = [skew(p') | -I]   (rot_vgicp_impl.hpp:265-277 / 345-372 / 574-590)
// and tree-reduces the 28 unique sums (+ correspondence count) to one 32-double result per pair:
// warp shuffles -> shared memory -> per-CTA partial -> last-CTA fixed-order final pass (deterministic).
//
// Data movement: source points (float4) and covariances (tile-planar 6 x fp32) stream through shared
// memory as 4 KB + 6 KB tiles with 1-D TMA bulk copies (cp.async.bulk + full/empty mbarrier ring, 4 stages,
// no CTA-wide barrier in the main loop);
// the voxel table/payload gathers go through the read-only path (L1/L2 resident). All arithmetic
// that decides a voxel or enters a sum is fp64 (fp32 is only a pre-filter for the key).
// No tensor cores: there is no dense contraction in this path.

#include "linearize.cuh"
#include <string.h>
#include <vector>

namespace rolo {

namespace {

constexpr int NTHREADS = TILE_PTS;               // 256
constexpr int NWARPS = NTHREADS / 32;
#ifndef LIN_MINBLOCKS
#define LIN_MINBLOCKS 2
#endif
constexpr int NSTAGES = 4;                       // TMA ring depth (4 x 10 KB per CTA)
constexpr uint32_t PTS_TILE_BYTES = TILE_PTS * sizeof(float4);             // 4096
constexpr uint32_t COV_TILE_BYTES = TILE_PTS * COV_PLANES * sizeof(float); // 6144

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

__device__ __forceinline__ double warp_sum(double v) {
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NOFF>
__device__ __forceinline__ void neighbor_offset(int o, int& dx, int& dy, int& dz) {
  if (NOFF == 1) { dx = dy = dz = 0; }
  else if (NOFF == 7) {   // vmp_voxel.hpp:23-31 order
    dx = (o == 1) - (o == 2); dy = (o == 3) - (o == 4); dz = (o == 5) - (o == 6);
  } else {                // vmp_voxel.hpp:38-46 order
    dx = o / 9 - 1; dy = (o / 3) % 3 - 1; dz = o % 3 - 1;
  }
}

// rare path of the table probe: the first slot held another key; continue linearly (vmp_voxel.hpp:226-233 semantics)
__device__ __noinline__ int probe_collision(const VoxSlot* __restrict__ table, unsigned mask, unsigned h, unsigned long long key) {
  for (unsigned probe = 0; probe < mask; probe++) {
    h = (h + 1) & mask;
    const ulonglong2 sl = __ldg(reinterpret_cast<const ulonglong2*>(table + h));
    if (sl.x == key) return (int)(unsigned)(sl.y & 0xFFFFFFFFull);
    if (sl.x == KEY_EMPTY) return -1;
  }
  return -1;
}

// number of accumulators actually used per mode
template <int MODE> struct AccN { static constexpr int value = (MODE == LIN_SE3) ? 29 : (MODE == LIN_SO3 ? 11 : 2); };
// SE3 : [0..5] Mw, [6..14] G, [15..20] StG, [21..23] q, [24..26] Stq, [27] cost, [28] count
// SO3 : [0..5] StG, [6..8] Stq, [9] cost, [10] count
// COST: [0] cost, [1] count

template <int MODE, bool REFRESH, int NOFF, bool CT>
__global__ void __launch_bounds__(NTHREADS, LIN_MINBLOCKS)
linearize_kernel(const LinStatic* __restrict__ st_arr, const LinDyn* __restrict__ dyn_arr, const __grid_constant__ LinDyn dyn0,
                 const int* __restrict__ sched, int n_pairs, int total_tiles) {
  // Persistent CTAs: the batch's tiles form one global sequence (pair 0's tiles, pair 1's, ...); CTA g owns the contiguous
  // range [g*T/G, (g+1)*T/G), so any number of pairs fills every SM evenly. A CTA flushes its accumulators whenever its
  // range crosses into the next pair; per pair the last arriving CTA sums the partials in fixed order.
  constexpr int NACC = AccN<MODE>::value;
  __shared__ __align__(128) float4 s_pts[NSTAGES][TILE_PTS];
  __shared__ __align__(128) float s_cov[NSTAGES][COV_PLANES * TILE_PTS];
  __shared__ __align__(8) uint64_t s_full[NSTAGES];    // TMA bytes landed
  __shared__ __align__(8) uint64_t s_empty[NSTAGES];   // every warp has copied its operands out of the stage
  __shared__ LinDyn s_dyn;
  __shared__ GridParams s_grid;
  __shared__ double s_red[NWARPS][RES_USED];
  __shared__ int s_is_last;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int* __restrict__ tile_prefix = sched;                  // [n_pairs + 1]
  const int* __restrict__ cta_first = sched + n_pairs + 1;      // [n_pairs] first CTA touching the pair
  const int* __restrict__ cta_count = cta_first + n_pairs;      // [n_pairs] number of CTAs touching the pair
  const int G = (int)gridDim.x, g = (int)blockIdx.x;
  const int t0 = (int)(((long long)g * total_tiles) / G), t1 = (int)(((long long)(g + 1) * total_tiles) / G);
  if (t0 >= t1) return;

  if (tid == 0) {
    #pragma unroll
    for (int sidx = 0; sidx < NSTAGES; sidx++) { mbar_init(&s_full[sidx], 1); mbar_init(&s_empty[sidx], NWARPS); }
    fence_mbar_init();
  }
  __syncthreads();

  // pair that owns tile t0: the largest p with tile_prefix[p] <= t0
  int pair;
  {
    int lo = 0, hi = n_pairs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tile_prefix[mid] <= t0) lo = mid; else hi = mid; }
    pair = lo;
  }
  int pf_pair = pair;   // producer cursor (thread 0)
  auto issue = [&](int tl, int sidx) {
    while (tile_prefix[pf_pair + 1] <= tl) pf_pair++;
    const LinStatic& ps = st_arr[pf_pair];
    const int plt = tl - tile_prefix[pf_pair];
    mbar_arrive_expect_tx(&s_full[sidx], PTS_TILE_BYTES + COV_TILE_BYTES);
    tma_load_1d(&s_pts[sidx][0], ps.src + (size_t)plt * TILE_PTS, PTS_TILE_BYTES, &s_full[sidx]);
    tma_load_1d(&s_cov[sidx][0], ps.cov6 + (size_t)plt * COV_PLANES * TILE_PTS, COV_TILE_BYTES, &s_full[sidx]);
  };
  if (tid == 0) {   // prologue: fill the ring (TMA 1-D bulk copies, one elected thread)
    #pragma unroll
    for (int sidx = 0; sidx < NSTAGES; sidx++) if (t0 + sidx < t1) issue(t0 + sidx, sidx);
  }

  double acc[NACC];
  #pragma unroll
  for (int k = 0; k < NACC; k++) acc[k] = 0.0;

  int it = 0;
  int t = t0;
  #pragma unroll 1
  while (t < t1) {
    while (tile_prefix[pair + 1] <= t) pair++;             // (pairs always own >= 1 tile; kept for safety)
    const LinStatic& st = st_arr[pair];
    const int pair_begin = tile_prefix[pair];
    const int pair_end = min(tile_prefix[pair + 1], t1);
    // stage this pair's pose block and grid parameters into shared memory (uniform operands for every thread)
    {
      const double* src = dyn_arr ? reinterpret_cast<const double*>(dyn_arr + pair) : reinterpret_cast<const double*>(&dyn0);
      double* dst = reinterpret_cast<double*>(&s_dyn);
      for (int k = tid; k < (int)(sizeof(LinDyn) / sizeof(double)); k += NTHREADS) dst[k] = src[k];
      if (tid == 0) s_grid = st.grid;
    }
    __syncthreads();
    const int n = st.n;
    const VoxSlot* __restrict__ table = st.table;
    const VoxPayload* __restrict__ vox = st.vox;
    int* __restrict__ corr = st.corr;
    const unsigned mask = st.mask;

    #pragma unroll 1
    for (; t < pair_end; t++, it++) {
      const int stage = it % NSTAGES;
      const uint32_t phase = (uint32_t)(it / NSTAGES) & 1u;
      const int lt = t - pair_begin;
      mbar_wait(&s_full[stage], phase);

      // copy this thread's operands out of the stage, then hand the stage back to the producer: no CTA-wide barrier
      const float4 pf = s_pts[stage][tid];
      const float* cv = &s_cov[stage][tid];
      const float af0 = cv[0 * TILE_PTS], af1 = cv[1 * TILE_PTS], af2 = cv[2 * TILE_PTS];
      const float af3 = cv[3 * TILE_PTS], af4 = cv[4 * TILE_PTS], af5 = cv[5 * TILE_PTS];
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[stage]);
      // producer duty first, compute second: refill the stage that was consumed in the PREVIOUS iteration (every warp
      // released it at the top of that iteration, so this wait is short) with the tile NSTAGES-1 ahead. Issuing it here
      // instead of after this thread's own math keeps the ring full even though thread 0 also computes.
      if (tid == 0 && it > 0) {
        const int tl = t - 1 + NSTAGES;
        if (tl < t1) {
          const int ps = (it - 1) % NSTAGES;
          mbar_wait(&s_empty[ps], (uint32_t)((it - 1) / NSTAGES) & 1u);
          issue(tl, ps);
        }
      }

      const int i = lt * TILE_PTS + tid;
      if (i < n) {
        const double px = (double)pf.x, py = (double)pf.y, pz = (double)pf.z;
        // p' = R p + t
        double x = s_dyn.R[0] * px + s_dyn.R[1] * py + s_dyn.R[2] * pz + s_dyn.t[0];
        double y = s_dyn.R[3] * px + s_dyn.R[4] * py + s_dyn.R[5] * pz + s_dyn.t[1];
        double z = s_dyn.R[6] * px + s_dyn.R[7] * py + s_dyn.R[8] * pz + s_dyn.t[2];
        // a point without a usable voxel key never matches in the reference (NaN / Inf coordinates from a non-dense cloud, or the exact
        // origin of a polar grid, whose acos(z/r) is NaN: x86 turns those into the unreachable cell INT_MIN). It is taken out here,
        // with its coordinates zeroed, so that 0 * NaN cannot poison the sums of the whole pair
        const bool usable = (x - x == 0.0) && (y - y == 0.0) && (z - z == 0.0) && (s_grid.voxel_type != 0 || x != 0.0 || y != 0.0 || z != 0.0);
        if (!usable) { x = 0.0; y = 0.0; z = 0.0; }

        int cx = 0, cy = 0, cz = 0;
        ulonglong2 slot0 = make_ulonglong2(KEY_EMPTY, 0ull);
        unsigned h0 = 0;
        bool key_ok = false;
        if (REFRESH) {
          voxel_coord(s_grid, x, y, z, cx, cy, cz);
          // issue the first probe (offset 0) now; it is consumed after C_A has been computed
          int dx0, dy0, dz0;
          neighbor_offset<NOFF>(0, dx0, dy0, dz0);
          key_ok = usable && key_in_range(cx + dx0, cy + dy0, cz + dz0);
          if (key_ok) { h0 = hash_key(pack_key(cx + dx0, cy + dy0, cz + dz0)) & mask; slot0 = __ldg(reinterpret_cast<const ulonglong2*>(table + h0)); }
        } else if (NOFF == 1) {
          slot0.y = (unsigned long long)(unsigned)corr[i];
        }

        // C_A = Rm A Rm^T (symmetric): independent of the lookup, overlaps the table probe latency
        double ca0, ca1, ca2, ca3, ca4, ca5;
        {
          const double a0 = (double)af0, a1 = (double)af1, a2 = (double)af2, a3 = (double)af3, a4 = (double)af4, a5 = (double)af5;
          const double r0 = s_dyn.Rm[0], r1 = s_dyn.Rm[1], r2 = s_dyn.Rm[2];
          const double r3 = s_dyn.Rm[3], r4 = s_dyn.Rm[4], r5 = s_dyn.Rm[5];
          const double r6 = s_dyn.Rm[6], r7 = s_dyn.Rm[7], r8 = s_dyn.Rm[8];
          const double t00 = r0 * a0 + r1 * a1 + r2 * a2, t01 = r0 * a1 + r1 * a3 + r2 * a4, t02 = r0 * a2 + r1 * a4 + r2 * a5;
          const double t10 = r3 * a0 + r4 * a1 + r5 * a2, t11 = r3 * a1 + r4 * a3 + r5 * a4, t12 = r3 * a2 + r4 * a4 + r5 * a5;
          const double t20 = r6 * a0 + r7 * a1 + r8 * a2, t21 = r6 * a1 + r7 * a3 + r8 * a4, t22 = r6 * a2 + r7 * a4 + r8 * a5;
          ca0 = t00 * r0 + t01 * r1 + t02 * r2;
          ca1 = t00 * r3 + t01 * r4 + t02 * r5;
          ca2 = t00 * r6 + t01 * r7 + t02 * r8;
          ca3 = t10 * r3 + t11 * r4 + t12 * r5;
          ca4 = t10 * r6 + t11 * r7 + t12 * r8;
          ca5 = t20 * r6 + t21 * r7 + t22 * r8;
        }

        #pragma unroll 1
        for (int o = 0; o < NOFF; o++) {
          int id;
          if (REFRESH) {
            if (o == 0) {
              // resolve the pre-issued probe (vmp_voxel.hpp:226-233): first slot hits almost always (load factor << 1)
              id = -1;
              if (key_ok) {
                int dx0, dy0, dz0;
                neighbor_offset<NOFF>(0, dx0, dy0, dz0);
                const unsigned long long key = pack_key(cx + dx0, cy + dy0, cz + dz0);
                if (slot0.x == key) id = (int)(unsigned)(slot0.y & 0xFFFFFFFFull);
                else if (slot0.x != KEY_EMPTY) id = probe_collision(table, mask, h0, key);
              }
            } else {
              int dx, dy, dz;
              neighbor_offset<NOFF>(o, dx, dy, dz);
              id = usable ? table_lookup(table, mask, cx + dx, cy + dy, cz + dz) : -1;
            }
            corr[(size_t)i * NOFF + o] = id;
          } else {
            id = (NOFF == 1) ? (int)(unsigned)slot0.y : corr[(size_t)i * NOFF + o];
          }
          // DIRECT1: no divergent skip — a miss runs the same straight-line math on voxel 0 with weight 0
          // (misses are rare); DIRECT7/27: most neighbour cells are empty, skip them
          const bool hit = id >= 0;
          if (NOFF > 1 && !hit) continue;
          if (!hit) id = 0;

          // voxel payload gather: 80 B as 5 x 16 B read-only loads
          const double2* vp = reinterpret_cast<const double2*>(vox + id);
          const double2 v0 = __ldg(vp + 0), v1 = __ldg(vp + 1), v2 = __ldg(vp + 2), v3 = __ldg(vp + 3), v4 = __ldg(vp + 4);
          const double mx = v0.x, my = v0.y, mz = v1.x;
          const double c0 = v1.y + ca0, c1 = v2.x + ca1, c2 = v2.y + ca2, c3 = v3.x + ca3, c4 = v3.y + ca4, c5 = v4.x + ca5;
          const double w = v4.y;

          // M = C^-1 = adj(C) / det. Every sum is linear in M, so all products are formed on the adjugate and the
          // reciprocal of the determinant is folded into the weights: the division (rcp.approx + 2 Newton steps, no
          // branchy IEEE slow path) leaves the critical path and overlaps the Jacobian products.
          const double m0 = c3 * c5 - c4 * c4;
          const double m1 = c2 * c4 - c1 * c5;
          const double m2 = c1 * c4 - c2 * c3;
          const double m3 = c0 * c5 - c2 * c2;
          const double m4 = c1 * c2 - c0 * c4;
          const double m5 = c0 * c3 - c1 * c1;
          const double det = c0 * m0 + c1 * m1 + c2 * m2;
          double idet;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(idet) : "d"(det));
          idet = fma(idet, fma(-det, idet, 1.0), idet);
          idet = fma(idet, fma(-det, idet, 1.0), idet);
          const double wi = hit ? w * idet : 0.0;           // weight of terms built on adj(C); exactly 0 for a miss (no NaN leak)

          // residuals (on the adjugate)
          const double ex = mx - x, ey = my - y, ez = mz - z;
          const double qx = m0 * ex + m1 * ey + m2 * ez;
          const double qy = m1 * ex + m3 * ey + m4 * ez;
          const double qz = m2 * ex + m4 * ey + m5 * ez;
          double cost = ex * qx + ey * qy + ez * qz;
          // continuous-time term (t3 modes; lam_c = 0 otherwise)  rot_vgicp_impl.hpp:548,555
          const double lam_c = s_dyn.lam_c;
          double bx = qx, by = qy, bz = qz;      // adj(C) e_b with e_b = e + b_ct_scale * ct
          if (CT) {
            const double tx = s_dyn.ct[0], ty = s_dyn.ct[1], tz = s_dyn.ct[2];
            const double ux = m0 * tx + m1 * ty + m2 * tz;
            const double uy = m1 * tx + m3 * ty + m4 * tz;
            const double uz = m2 * tx + m4 * ty + m5 * tz;
            cost += lam_c * (tx * ux + ty * uy + tz * uz);
            bx += s_dyn.b_ct_scale * ux; by += s_dyn.b_ct_scale * uy; bz += s_dyn.b_ct_scale * uz;
          }

          if (MODE == LIN_COST) {
            acc[0] += wi * cost;
            acc[1] += hit ? 1.0 : 0.0;
          } else {
            const double ws = wi * s_dyn.h_scale;
            // G = M S, S = skew(p') = [[0,-z,y],[z,0,-x],[-y,x,0]]
            const double g00 = m1 * z - m2 * y, g01 = m2 * x - m0 * z, g02 = m0 * y - m1 * x;
            const double g10 = m3 * z - m4 * y, g11 = m4 * x - m1 * z, g12 = m1 * y - m3 * x;
            const double g20 = m4 * z - m5 * y, g21 = m5 * x - m2 * z, g22 = m2 * y - m4 * x;
            // S^T G (symmetric), S^T = [[0,z,-y],[-z,0,x],[y,-x,0]]
            const double h00 = z * g10 - y * g20, h01 = z * g11 - y * g21, h02 = z * g12 - y * g22;
            const double h11 = x * g21 - z * g01, h12 = x * g22 - z * g02;
            const double h22 = y * g02 - x * g12;
            // S^T (M e_b)
            const double sx = z * by - y * bz, sy = x * bz - z * bx, sz = y * bx - x * by;
            if (MODE == LIN_SE3) {
              acc[0] += ws * m0; acc[1] += ws * m1; acc[2] += ws * m2; acc[3] += ws * m3; acc[4] += ws * m4; acc[5] += ws * m5;
              acc[6] += ws * g00; acc[7] += ws * g01; acc[8] += ws * g02;
              acc[9] += ws * g10; acc[10] += ws * g11; acc[11] += ws * g12;
              acc[12] += ws * g20; acc[13] += ws * g21; acc[14] += ws * g22;
              acc[15] += ws * h00; acc[16] += ws * h01; acc[17] += ws * h02; acc[18] += ws * h11; acc[19] += ws * h12; acc[20] += ws * h22;
              acc[21] += wi * bx; acc[22] += wi * by; acc[23] += wi * bz;
              acc[24] += wi * sx; acc[25] += wi * sy; acc[26] += wi * sz;
              acc[27] += wi * cost;
              acc[28] += hit ? 1.0 : 0.0;
            } else {
              acc[0] += ws * h00; acc[1] += ws * h01; acc[2] += ws * h02; acc[3] += ws * h11; acc[4] += ws * h12; acc[5] += ws * h22;
              acc[6] += wi * sx; acc[7] += wi * sy; acc[8] += wi * sz;
              acc[9] += wi * cost;
              acc[10] += hit ? 1.0 : 0.0;
            }
          }
        }
      }
    }

    // ---- flush this pair's sums: warp shuffle tree -> shared -> CTA partial -> last-CTA fixed-order final pass ----
    #pragma unroll
    for (int k = 0; k < NACC; k++) {
      const double v = warp_sum(acc[k]);
      if (lane == 0) s_red[warp][k] = v;
      acc[k] = 0.0;
    }
    __syncthreads();
    const int n_cta = cta_count[pair];
    double* partial = st.partials + (size_t)(g - cta_first[pair]) * RES_N;
    if (tid < NACC) {
      double v = 0.0;
      #pragma unroll
      for (int wq = 0; wq < NWARPS; wq++) v += s_red[wq][tid];
      partial[tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned ticket = atomicAdd(st.counter, 1u);
      s_is_last = (ticket == (unsigned)n_cta - 1u);
    }
    __syncthreads();
    if (s_is_last) {
      __threadfence();
      // last CTA of this pair: fixed-order sum over the CTA partials (deterministic for a given launch shape)
      const int c = tid & 31, gq0 = tid >> 5;
      double v = 0.0;
      if (c < NACC) {
        const volatile double* pp = st.partials;
        for (int b = gq0; b < n_cta; b += NWARPS) v += pp[(size_t)b * RES_N + c];
        s_red[gq0][c] = v;
      }
      __syncthreads();
      double tot = 0.0;
      int dst = -1;
      if (tid < RES_N) {
        if (tid < NACC) {
          #pragma unroll
          for (int gq = 0; gq < NWARPS; gq++) tot += s_red[gq][tid];
        }
        // scatter the mode-specific accumulator layout into the common 32-double result layout
        if (MODE == LIN_SE3) dst = tid < NACC ? tid : -1;
        else if (MODE == LIN_SO3) dst = tid < 6 ? 15 + tid : (tid < 9 ? 24 + (tid - 6) : (tid == 9 ? 27 : (tid == 10 ? 28 : -1)));
        else dst = tid == 0 ? 27 : (tid == 1 ? 28 : -1);
      }
      if (st.shard_world <= 1) {
        if (dst >= 0) st.result[dst] = tot;
      } else {
        // ---- fused reduction over NVLink peer memory (point-sharded registration) ----
        __shared__ double s_out[RES_N];
        __shared__ int s_timeout;
        const int W = st.shard_world, me = st.shard_rank;
        const unsigned long long sq = (unsigned long long)s_dyn.shard_seq;
        const size_t row_mine = (size_t)((sq & 1ull) * XCHG_MAX_RANKS + me) * XCHG_STRIDE;
        if (tid < RES_N) s_out[tid] = 0.0;
        if (tid == 0) s_timeout = 0;
        __syncthreads();
        if (dst >= 0) s_out[dst] = tot;
        __syncthreads();
        if (tid < RES_N) {
          const double v = s_out[tid];
          for (int p = 0; p < W; p++) st.peer_xchg[p][row_mine + tid] = v;               // local store or NVLink P2P store
        }
        __threadfence_system();
        __syncthreads();
        if (tid < W) *reinterpret_cast<volatile unsigned long long*>(st.peer_xchg[tid] + row_mine + RES_N) = sq;   // publish my row to peer `tid`
        if (tid < W) {                                                                    // wait for source rank `tid` in my own buffer
          const volatile unsigned long long* f =
              reinterpret_cast<const volatile unsigned long long*>(st.peer_xchg[me] + (size_t)((sq & 1ull) * XCHG_MAX_RANKS + tid) * XCHG_STRIDE + RES_N);
          long long spins = 0;
          while (*f != sq) { if (++spins > 20000000ll) { s_timeout = 1; break; } }       // bounded: a missing peer must not hang the GPU
        }
        __threadfence_system();
        __syncthreads();
        if (tid < RES_N) {
          double v = 0.0;
          const volatile double* mine = st.peer_xchg[me];
          for (int r = 0; r < W; r++) v += mine[(size_t)((sq & 1ull) * XCHG_MAX_RANKS + r) * XCHG_STRIDE + tid];   // fixed rank order
          if (s_timeout && tid == 28) v = __longlong_as_double(0x7ff8000000000000ll);    // NaN count => host reports the timeout
          st.result[tid] = v;
        }
      }
      __syncthreads();
      if (tid == 0) {
        *st.counter = 0u;                       // re-arm for the next launch
        if (st.done_flag) {
          __threadfence_system();
          *reinterpret_cast<volatile unsigned*>(st.done_flag) = s_dyn.seq;
        }
      }
    }
    __syncthreads();   // s_red / s_is_last / s_dyn are reused by the next pair segment
  }
}

template <int MODE, bool REFRESH, int NOFF, bool CT>
cudaError_t launch_t(const LinLaunch& L, const LinDyn& dyn0, cudaStream_t stream) {
  linearize_kernel<MODE, REFRESH, NOFF, CT><<<(unsigned)L.grid, NTHREADS, 0, stream>>>(L.st, L.dyn, dyn0, L.sched, L.n_pairs, L.total_tiles);
  return cudaGetLastError();
}

template <int NOFF>
cudaError_t launch_n(const LinLaunch& L, const LinDyn& dyn0, cudaStream_t stream) {
  // CT (continuous-time term of the translation stage) only occurs with cached correspondences
  if (L.refresh) {
    if (L.mode == LIN_SE3) return launch_t<LIN_SE3, true, NOFF, false>(L, dyn0, stream);
    if (L.mode == LIN_SO3) return launch_t<LIN_SO3, true, NOFF, false>(L, dyn0, stream);
    return launch_t<LIN_COST, true, NOFF, false>(L, dyn0, stream);
  }
  if (L.mode == LIN_SE3) return launch_t<LIN_SE3, false, NOFF, true>(L, dyn0, stream);
  if (L.mode == LIN_SO3) return launch_t<LIN_SO3, false, NOFF, false>(L, dyn0, stream);
  return L.ct ? launch_t<LIN_COST, false, NOFF, true>(L, dyn0, stream) : launch_t<LIN_COST, false, NOFF, false>(L, dyn0, stream);
}

}  // namespace

cudaError_t launch_linearize(const LinLaunch& L, const LinDyn* dyn0, cudaStream_t stream) {
  LinDyn zero;
  if (!dyn0) { memset(&zero, 0, sizeof(zero)); dyn0 = &zero; }
  if (!L.dyn && dyn0 == &zero) return cudaErrorInvalidValue;
  switch (L.noff) {
    case 1: return launch_n<1>(L, *dyn0, stream);
    case 7: return launch_n<7>(L, *dyn0, stream);
    case 27: return launch_n<27>(L, *dyn0, stream);
  }
  return cudaErrorInvalidValue;
}

void build_schedule(const int* n_points, int n_pairs, int grid, std::vector<int>& sched, int& total_tiles) {
  sched.assign((size_t)3 * n_pairs + 1, 0);
  int* prefix = sched.data(); int* first = prefix + n_pairs + 1; int* count = first + n_pairs;
  for (int p = 0; p < n_pairs; p++) prefix[p + 1] = prefix[p] + (n_points[p] + TILE_PTS - 1) / TILE_PTS;
  total_tiles = prefix[n_pairs];
  const long long T = total_tiles, G = grid;
  for (int p = 0; p < n_pairs; p++) {
    const long long a = prefix[p], e = prefix[p + 1];
    if (e <= a) { first[p] = 0; count[p] = 0; continue; }
    // CTA g owns [g*T/G, (g+1)*T/G): first CTA whose range end exceeds a, last CTA whose range start is below e
    long long gf = (a * G) / T; while (((gf + 1) * T) / G <= a) gf++; while (gf > 0 && (gf * T) / G > a) gf--;
    long long gl = ((e - 1) * G) / T; while (((gl + 1) * T) / G <= e - 1) gl++; while (gl > 0 && (gl * T) / G > e - 1) gl--;
    first[p] = (int)gf; count[p] = (int)(gl - gf + 1);
  }
}

int linearize_max_blocks_per_sm() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, linearize_kernel<LIN_SE3, true, 1, false>, NTHREADS, 0);
  return nb;
}

}  // namespace rolo
