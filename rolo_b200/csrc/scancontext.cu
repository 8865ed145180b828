// scancontext.cu — SCManager::makeScancontext + makeRingkeyFromScancontext + makeSectorkeyFromScancontext
// (src/scancontext/Scancontext.cpp:151-227) on the GPU.
//
// Reference: a serial loop over the scan; each point goes to a (ring, sector) polar bin, the bin keeps the maximum of
// z + LIDAR_HEIGHT; empty bins (-1000) become 0; the ring key is the mean of each row, the sector key the mean of each column.
// Here: one thread per point, the maximum through atomicMax on the order-preserving integer image of the float (a maximum
// does not depend on the order of the points, so the result is bit-equal to the serial loop), then one small block turns
// the bins into the double descriptor and the two keys (sequential sums, like the oracle).
// xy2theta's atan is evaluated in double and rounded to float: that is the correctly rounded atanf the reference's
// std::atan(float) approximates to < 1 ulp; device and oracle agree on it exactly (a 1-ulp libm difference would flip a
// sector bin for points within 1e-5 degree of a bin edge).
#include "scancontext.cuh"

#include <math_constants.h>

namespace rolo {

namespace {

constexpr float SC_NO_POINT = -1000.f;

__device__ __forceinline__ int f2o(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ float atan_rn(float v) { return (float)atan((double)v); }

// Scancontext.cpp:23-36 (float in, float out; the products with 180/M_PI are formed in double)
__device__ __forceinline__ float xy2theta(float x, float y) {
  const double k = 180.0 / 3.14159265358979323846;
  if ((x >= 0) & (y >= 0)) return (float)(k * (double)atan_rn(y / x));
  if ((x < 0) & (y >= 0)) return (float)(180.0 - (k * (double)atan_rn(y / (-x))));
  if ((x < 0) & (y < 0)) return (float)(180.0 + (k * (double)atan_rn(y / x)));
  if ((x >= 0) & (y < 0)) return (float)(360.0 - (k * (double)atan_rn((-y) / x)));
  return CUDART_NAN_F;              // NaN coordinates: the reference falls off the end of the function
}

// int(ceil(v)) with the x86 conversion of NaN / out-of-range values (INT_MIN), then clamped to [1, hi]
__device__ __forceinline__ int clamp_bin(double v, int hi) {
  const double c = ceil(v);
  const int i = (c != c || c >= 2147483648.0 || c < -2147483648.0) ? (int)0x80000000 : (int)c;
  return max(min(hi, i), 1);
}

__global__ void sc_init_kernel(int* bins) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < SC_BINS) bins[t] = f2o(SC_NO_POINT);
}

__global__ void sc_bin_kernel(const float4* __restrict__ pts, int n, double lidar_height, double max_radius, int* __restrict__ bins) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const float z = (float)((double)p.z + lidar_height);
  const float azim_range = sqrtf(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)));
  const float azim_angle = xy2theta(p.x, p.y);
  if ((double)azim_range > max_radius) return;
  const int ring = clamp_bin(((double)azim_range / max_radius) * (double)SC_NUM_RING, SC_NUM_RING);
  const int sector = clamp_bin(((double)azim_angle / 360.0) * (double)SC_NUM_SECTOR, SC_NUM_SECTOR);
  if (z != z) return;               // `desc < NaN` is false: a NaN height never updates the bin
  atomicMax(&bins[(ring - 1) * SC_NUM_SECTOR + (sector - 1)], f2o(z));
}

__global__ void sc_finalize_kernel(const int* __restrict__ bins, double* __restrict__ desc, double* __restrict__ ring_key, double* __restrict__ sector_key) {
  __shared__ double s[SC_BINS];
  for (int t = threadIdx.x; t < SC_BINS; t += blockDim.x) {
    const float v = o2f(bins[t]);
    const double d = (v == SC_NO_POINT) ? 0.0 : (double)v;     // "reset no points to zero" (:183-187)
    s[t] = d; desc[t] = d;
  }
  __syncthreads();
  if (threadIdx.x < SC_NUM_RING) {
    double a = 0.0;
    for (int c = 0; c < SC_NUM_SECTOR; c++) a += s[threadIdx.x * SC_NUM_SECTOR + c];
    ring_key[threadIdx.x] = a / (double)SC_NUM_SECTOR;
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + SC_NUM_SECTOR) {
    const int c = threadIdx.x - 32;
    double a = 0.0;
    for (int r = 0; r < SC_NUM_RING; r++) a += s[r * SC_NUM_SECTOR + c];
    sector_key[c] = a / (double)SC_NUM_RING;
  }
}

}  // namespace

cudaError_t launch_scancontext(const SCLaunch& L, cudaStream_t stream) {
  sc_init_kernel<<<(SC_BINS + 255) / 256, 256, 0, stream>>>(L.bins);
  if (L.n > 0) sc_bin_kernel<<<(L.n + 255) / 256, 256, 0, stream>>>(L.pts, L.n, L.lidar_height, L.max_radius, L.bins);
  sc_finalize_kernel<<<1, 128, 0, stream>>>(L.bins, L.desc, L.ring_key, L.sector_key);
  return cudaGetLastError();
}

}  // namespace rolo
