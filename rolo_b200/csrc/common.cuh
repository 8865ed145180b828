// common.cuh — shared device/host definitions of the B200 (sm_100a) voxelized-GICP core.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace rolo {

// ---- data layout in HBM -------------------------------------------------------
// points   : float4 xyzw[Npad]           (w = 1; Npad = N rounded up to TILE_PTS, pad = far sentinel)
// cov6     : tile-planar float: tile b (TILE_PTS points) holds 6 planes [xx|xy|xz|yy|yz|zz] of
//            TILE_PTS floats each, i.e. element (i, c) at  (i / TILE) * 6*TILE + c*TILE + i % TILE.
//            One 1-D TMA bulk copy moves a whole tile (6 KB); shared-memory reads are conflict-free.
// table    : open-addressing hash, VoxSlot[cap] (16 B), cap = pow2 >= 2 * Ntarget
// payload  : VoxPayload[V] (80 B, fp64): mean, covariance (6 unique), w = sqrt(n)
constexpr int TILE_PTS = 256;          // points per TMA tile == threads per CTA in the fused kernel
constexpr int COV_PLANES = 6;
constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr int KEY_BIAS = 1 << 20;      // 21 bits per axis

struct __align__(16) VoxSlot {
  unsigned long long key;   // packed biased (x,y,z), KEY_EMPTY = free
  int id;                   // compact voxel index into the payload array, -1 until assigned
  int pad;
};

struct __align__(16) VoxPayload {
  double mean[3];
  double cov[6];            // xx xy xz yy yz zz (mean of the member points' covariances, voxel.hpp:107)
  double w;                 // sqrt(num_points) (impl.hpp:258)
};

struct VoxAccum {            // accumulation buffer (struct-of-arrays by field, fp64 atomics)
  double* sum;               // [V_cap][10]: mean xyz (3), cov6 (6), count (1)
};

struct GridParams {
  int voxel_type;            // RG_VOXEL_POLAR = 0 / RG_VOXEL_UNIFORM = 1
  double res[3];             // polar (theta, phi, r) or uniform (res, res, res)
  double inv_res[3];
};

// Result vector of one linearize / compute_error launch (per pair): 32 doubles.
//  [0..5]   Mw   = sum w*s M            (xx xy xz yy yz zz)       -> H[3:6,3:6]
//  [6..14]  G    = sum w*s M S          (row-major 3x3)           -> H[3:6,0:3] = -G, H[0:3,3:6] = -G^T
//  [15..20] StG  = sum w*s S^T M S      (xx xy xz yy yz zz)       -> H[0:3,0:3]
//  [21..23] q    = sum w M e_b                                     -> b[3:6] = -q
//  [24..26] Stq  = sum w S^T M e_b                                 -> b[0:3]
//  [27]     cost = sum w (e^T M e + lam_c ct^T M ct)
//  [28]     number of correspondences
//  [29..31] reserved
constexpr int RES_N = 32;
constexpr int RES_USED = 29;

inline __host__ __device__ unsigned long long pack_key(int x, int y, int z) {
  return ((unsigned long long)(unsigned)(x + KEY_BIAS) << 42) | ((unsigned long long)(unsigned)(y + KEY_BIAS) << 21) |
         (unsigned long long)(unsigned)(z + KEY_BIAS);
}
inline __host__ __device__ bool key_in_range(int x, int y, int z) {
  return (unsigned)(x + KEY_BIAS) < (2u << 20) && (unsigned)(y + KEY_BIAS) < (2u << 20) && (unsigned)(z + KEY_BIAS) < (2u << 20);
}
inline __host__ __device__ void unpack_key(unsigned long long k, int& x, int& y, int& z) {
  x = (int)((k >> 42) & 0x1FFFFF) - KEY_BIAS;
  y = (int)((k >> 21) & 0x1FFFFF) - KEY_BIAS;
  z = (int)(k & 0x1FFFFF) - KEY_BIAS;
}
inline __host__ __device__ unsigned hash_key(unsigned long long k) {   // 32-bit mix of the three 21-bit fields (murmur3 finaliser)
  unsigned h = (unsigned)(k >> 42) * 0x9E3779B1u ^ (unsigned)((k >> 21) & 0x1FFFFF) * 0x85EBCA77u ^ (unsigned)(k & 0x1FFFFF) * 0xC2B2AE3Du;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}

#ifdef __CUDACC__
// Voxel coordinate of a point given in double, bit-compatible with the reference:
//  UNIFORM: floor(x/res - 0.5) per axis                      (vmp_voxel.hpp:199-201)
//  POLAR  : floor([atan2(y,x)+pi, acos(z/r), r] / res)       (vmp_voxel.hpp:208-211)
// Exact fp64 path (always correct, slow on the transcendental units).
__device__ __forceinline__ void voxel_coord_exact(const GridParams& g, double x, double y, double z, int& cx, int& cy, int& cz) {
  if (g.voxel_type == 0) {
    double r = sqrt(x * x + y * y + z * z);
    cx = (int)floor((atan2(y, x) + 3.14159265358979323846) / g.res[0]);
    cy = (int)floor(acos(z / r) / g.res[1]);
    cz = (int)floor(r / g.res[2]);
  } else {
    cx = (int)floor(x / g.res[0] - 0.5);
    cy = (int)floor(y / g.res[1] - 0.5);
    cz = (int)floor(z / g.res[2] - 0.5);
  }
}

// fp32 minimax approximations used ONLY by the key pre-filter below (coefficients fitted offline; maximum error incl. fp32
// evaluation measured over 4e6 samples: atan2 2.2e-6 rad, acos 9.4e-7 rad on |u| <= 0.95). They never decide a key alone:
// a value within the margin of a cell border is re-evaluated with the exact fp64 expression.
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = __fdividef(mn, mx);
  const float s = a * a;
  float q = -0.011465809307992458f;
  q = fmaf(q, s, 0.05195053294301033f);
  q = fmaf(q, s, -0.1157294437289238f);
  q = fmaf(q, s, 0.19323529303073883f);
  q = fmaf(q, s, -0.33256831765174866f);
  q = fmaf(q, s, 0.9999744892120361f);
  float r = a * q;
  r = (ay > ax) ? 1.57079637f - r : r;
  r = (x < 0.f) ? 3.14159274f - r : r;
  return (y < 0.f) ? -r : r;
}
__device__ __forceinline__ float acos_fast(float u) {
  const float au = fabsf(u);
  float q = -0.00482112355530262f;
  q = fmaf(q, au, 0.020417358726263046f);
  q = fmaf(q, au, -0.04576948285102844f);
  q = fmaf(q, au, 0.08812171965837479f);
  q = fmaf(q, au, -0.21453753113746643f);
  q = fmaf(q, au, 1.5707956552505493f);
  const float r = sqrtf(1.0f - au) * q;
  return (u < 0.f) ? 3.14159274f - r : r;
}

// Filtered path: evaluate in fp32 (polar) or with a reciprocal multiply (uniform), and only fall
// back to the exact fp64 evaluation when the value lies within a conservative error bound of a cell
// boundary. The returned coordinate always equals voxel_coord_exact().
// Error budgets (absolute, before dividing by the resolution; >= 3x the measured approximation error):
//   theta: fp32 inputs (6e-8 rel) + atan2_fast (2.2e-6 + approximate divide 3e-7) + fp32 pi/add -> 9e-6 rad
//   phi  : u = z * rsqrt(r2) (4e-7 abs) * 1/sqrt(1-u^2) (<= 2.3 for |u|<=0.9) + acos_fast (9.4e-7) -> 5e-6 rad; |u|>0.9 -> exact
//   r    : r2 * rsqrt(r2) chain 6e-7 relative
__device__ __forceinline__ void voxel_coord(const GridParams& g, double x, double y, double z, int& cx, int& cy, int& cz) {
  if (g.voxel_type == 0) {
    // everything in fp32; the margins include the fp32 rounding of the final scale (|value| * 2^-22)
    const float xf = (float)x, yf = (float)y, zf = (float)z;
    const float r2 = xf * xf + yf * yf + zf * zf;
    const float ir = rsqrtf(r2);
    const float rf = r2 * ir;
    const float u = zf * ir;
    const float ir0 = (float)g.inv_res[0], ir1 = (float)g.inv_res[1], ir2 = (float)g.inv_res[2];
    const float th = (atan2_fast(yf, xf) + 3.14159274f) * ir0;
    const float ph = acos_fast(u) * ir1;
    const float rr = rf * ir2;
    const float fth = floorf(th), fph = floorf(ph), frr = floorf(rr);
    const float dth = 9.0e-6f * ir0 + 4e-7f * th;
    const float dph = 5.0e-6f * ir1 + 4e-7f * ph;
    const float drr = 1.2e-6f * rr + 1e-9f;
    bool safe = (th - fth > dth) && (fth + 1.0f - th > dth) && (ph - fph > dph) && (fph + 1.0f - ph > dph) &&
                (rr - frr > drr) && (frr + 1.0f - rr > drr) && (fabsf(u) <= 0.9f) && (rf > 1e-3f) && (rf < 1e6f) &&
                (th < 1e6f) && (ph < 1e6f) && (rr < 1e6f);
    if (safe) { cx = (int)fth; cy = (int)fph; cz = (int)frr; }
    else voxel_coord_exact(g, x, y, z, cx, cy, cz);
  } else {
    const double vx = x * g.inv_res[0] - 0.5, vy = y * g.inv_res[1] - 0.5, vz = z * g.inv_res[2] - 0.5;
    const double fx = floor(vx), fy = floor(vy), fz = floor(vz);
    const double ex = 1e-12 * (fabs(vx) + 1.0), ey = 1e-12 * (fabs(vy) + 1.0), ez = 1e-12 * (fabs(vz) + 1.0);
    bool safe = (vx - fx > ex) && (fx + 1.0 - vx > ex) && (vy - fy > ey) && (fy + 1.0 - vy > ey) && (vz - fz > ez) && (fz + 1.0 - vz > ez);
    if (safe) { cx = (int)fx; cy = (int)fy; cz = (int)fz; }
    else voxel_coord_exact(g, x, y, z, cx, cy, cz);
  }
}

// Hash lookup: returns compact voxel id or -1. Linear probing, table is read-only here.
__device__ __forceinline__ int table_lookup(const VoxSlot* __restrict__ table, unsigned mask, int cx, int cy, int cz) {
  if (!key_in_range(cx, cy, cz)) return -1;
  const unsigned long long key = pack_key(cx, cy, cz);
  unsigned h = hash_key(key) & mask;
  #pragma unroll 1
  for (unsigned probe = 0; probe <= mask; probe++) {
    const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2*>(table + h));
    if (s.x == key) return (int)(unsigned)(s.y & 0xFFFFFFFFull);
    if (s.x == KEY_EMPTY) return -1;
    h = (h + 1) & mask;
  }
  return -1;
}
#endif  // __CUDACC__

}  // namespace rolo
