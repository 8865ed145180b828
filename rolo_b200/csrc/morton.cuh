// morton.cuh — device helpers of the Morton-ordered spatial index (shared by covariance.cu and scan2map.cu).
// A cloud is indexed by: its bounding box (ordered-int min/max), 39-bit Morton codes of the quantised coordinates, the
// radix-sorted (code, index) pairs, a Morton-ordered copy of the points and a bucket table over the top 15 code bits.
// A level-L cell (L <= MORTON_BITS) is one contiguous range of the sorted codes, found by two binary searches.
#pragma once
#include <cuda_runtime.h>

namespace rolo {

constexpr int MORTON_BITS = 13;           // bits per axis: 39-bit codes = 5 radix passes; finest cell = extent / 8192
constexpr int CODE_BITS = 3 * MORTON_BITS;

__device__ __forceinline__ int float_to_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

struct Quant {           // derived from the bbox on the fly by every kernel (no host round trip)
  double mn[3];
  double scale;          // finest cells per metre
  double cell;           // finest cell size (metres)
};

__device__ __forceinline__ Quant make_quant(const int* bbox) {
  Quant q;
  double ext = 0.0;
  #pragma unroll
  for (int d = 0; d < 3; d++) {
    q.mn[d] = (double)ordered_to_float(bbox[d]);
    ext = fmax(ext, (double)ordered_to_float(bbox[3 + d]) - q.mn[d]);
  }
  ext = fmax(ext, 1e-6) * (1.0 + 1e-6);
  q.cell = ext / (double)(1 << MORTON_BITS);
  q.scale = 1.0 / q.cell;
  return q;
}

__device__ __forceinline__ unsigned long long expand21(unsigned v) {   // spread 21 bits to every third bit
  unsigned long long x = v & 0x1FFFFFull;
  x = (x | x << 32) & 0x1F00000000FFFFull;
  x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full;
  x = (x | x << 4) & 0x10C30C30C30C30C3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ unsigned long long morton3(unsigned x, unsigned y, unsigned z) {
  return expand21(x) | (expand21(y) << 1) | (expand21(z) << 2);
}
__device__ __forceinline__ void quantize(const Quant& q, float x, float y, float z, unsigned& ix, unsigned& iy, unsigned& iz) {
  const double m = (double)((1 << MORTON_BITS) - 1);
  ix = (unsigned)fmin(fmax(floor(((double)x - q.mn[0]) * q.scale), 0.0), m);
  iy = (unsigned)fmin(fmax(floor(((double)y - q.mn[1]) * q.scale), 0.0), m);
  iz = (unsigned)fmin(fmax(floor(((double)z - q.mn[2]) * q.scale), 0.0), m);
}

__device__ __forceinline__ int lower_bound_u64(const unsigned long long* __restrict__ a, int n, unsigned long long key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int BUCKET_BITS = 15;       // top Morton bits indexed by the bucket table (octree level 5): 128 KB per cloud, L2-resident for a whole group
constexpr int BUCKET_SHIFT = 3 * MORTON_BITS - BUCKET_BITS;
constexpr int N_BUCKETS = 1 << BUCKET_BITS;

__device__ __forceinline__ int lower_bound_bucketed(const unsigned long long* __restrict__ codes, const int* __restrict__ bucket_start,
                                                    unsigned long long key) {
  const unsigned long long top = key >> BUCKET_SHIFT;
  if (top >= (unsigned long long)N_BUCKETS) return __ldg(bucket_start + N_BUCKETS);
  int lo = __ldg(bucket_start + top);
  if ((key & ((1ull << BUCKET_SHIFT) - 1ull)) == 0ull) return lo;     // bucket-aligned key (levels <= 5)
  int hi = __ldg(bucket_start + top + 1);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(codes + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ unsigned compact21(unsigned long long x) {   // inverse of expand21
  x &= 0x1249249249249249ull;
  x = (x ^ (x >> 2)) & 0x10C30C30C30C30C3ull;
  x = (x ^ (x >> 4)) & 0x100F00F00F00F00Full;
  x = (x ^ (x >> 8)) & 0x1F0000FF0000FFull;
  x = (x ^ (x >> 16)) & 0x1F00000000FFFFull;
  x = (x ^ (x >> 32)) & 0x1FFFFFull;
  return (unsigned)x;
}

// candidates are ordered by the packed key (float distance bits << 32 | index): distances are >= +0, so the
// unsigned bit pattern is monotone, and the index breaks ties exactly like the oracle's (distance, index) order
__device__ __forceinline__ void bitonic_cmpx(unsigned long long& key, int j2, bool take_min) {
  const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j2);
  if (take_min == (other < key)) key = other;
}
constexpr unsigned long long KEY_INF = 0x7F7FFFFF7FFFFFFFull;            // (FLT_MAX, INT_MAX)


}  // namespace rolo
