// voxelmap.cu — VmfVoxelMap::create_voxelmap on the GPU (vmp_voxel.hpp:167-197, 93-108).
//
// The reference builds an unordered_map<Vector3i, shared_ptr<Voxel>> serially and accumulates, per
// voxel, the sum of member points (4d), the sum of member covariances (4x4) and the count, then
// divides by the count. Here:
//   1. vox_insert   : key of every target point -> open-addressing hash (atomicCAS on the packed key);
//                     the CAS winner takes the next compact voxel id and zeroes its accumulator row.
//   2. vox_accumulate: fp64 sums per voxel. Points arrive in scan order, so consecutive lanes mostly
//                     share a voxel: a segmented warp scan folds each run and only run tails issue
//                     the 10 fp64 atomics.
//   3. vox_finalize : mean = sum/n, cov = sum/n, w = sqrt(n) -> 80-byte payload rows.
// kappa / r_bar / dir_reg (vmp_voxel.hpp:97,103-106) are never read by any cost or Jacobian path and
// are not materialised.

#include "voxelmap.cuh"

namespace rolo {

namespace {

__global__ void vox_insert_kernel(const float4* __restrict__ pts, int n, GridParams grid, VoxSlot* table, unsigned mask,
                                  int* __restrict__ slot_of_point, int* counters /* [0]=voxel count, [1]=error flag */,
                                  double* __restrict__ sums, int pair_stride_unused) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  int cx, cy, cz;
  voxel_coord(grid, (double)p.x, (double)p.y, (double)p.z, cx, cy, cz);
  if (!key_in_range(cx, cy, cz)) { atomicExch(&counters[1], 1); slot_of_point[i] = -1; return; }
  const unsigned long long key = pack_key(cx, cy, cz);
  unsigned h = hash_key(key) & mask;
  for (unsigned probe = 0; probe <= mask; probe++) {
    unsigned long long* kp = &table[h].key;
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
    if (cur == KEY_EMPTY) {
      cur = atomicCAS(kp, KEY_EMPTY, key);
      if (cur == KEY_EMPTY) {              // we created the voxel
        const int id = atomicAdd(&counters[0], 1);
        table[h].id = id;
        double* row = sums + (size_t)id * 10;
        #pragma unroll
        for (int k = 0; k < 10; k++) row[k] = 0.0;
        slot_of_point[i] = (int)h;
        return;
      }
    }
    if (cur == key) { slot_of_point[i] = (int)h; return; }
    h = (h + 1) & mask;
  }
  atomicExch(&counters[1], 2);   // table full (cannot happen with cap >= 2n)
  slot_of_point[i] = -1;
}

__global__ void vox_accumulate_kernel(const float4* __restrict__ pts, const float* __restrict__ cov6, int n,
                                      const int* __restrict__ slot_of_point, const VoxSlot* __restrict__ table,
                                      double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  int id = -1;
  double v[10];
  #pragma unroll
  for (int k = 0; k < 10; k++) v[k] = 0.0;
  if (i < n) {
    const int slot = slot_of_point[i];
    if (slot >= 0) {
      id = table[slot].id;
      const float4 p = pts[i];
      const size_t base = (size_t)(i / TILE_PTS) * COV_PLANES * TILE_PTS + (i % TILE_PTS);
      v[0] = (double)p.x; v[1] = (double)p.y; v[2] = (double)p.z;
      #pragma unroll
      for (int c = 0; c < 6; c++) v[3 + c] = (double)cov6[base + (size_t)c * TILE_PTS];
      v[9] = 1.0;
    }
  }
  // segmented inclusive scan over runs of equal id among consecutive lanes
  const int prev = __shfl_up_sync(0xffffffffu, id, 1);
  const bool head = (lane == 0) || (prev != id);
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
  #pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    #pragma unroll
    for (int k = 0; k < 10; k++) {
      const double o = __shfl_up_sync(0xffffffffu, v[k], off);
      if ((int)lane - off >= start) v[k] += o;
    }
  }
  const int nxt = __shfl_down_sync(0xffffffffu, id, 1);
  const bool tail = (lane == 31) || (nxt != id);
  if (tail && id >= 0) {
    double* row = sums + (size_t)id * 10;
    #pragma unroll
    for (int k = 0; k < 10; k++) atomicAdd(row + k, v[k]);
  }
}

__global__ void vox_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ counters, VoxPayload* __restrict__ vox, int max_v) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= max_v || id >= counters[0]) return;
  const double* row = sums + (size_t)id * 10;
  const double cnt = row[9];
  VoxPayload p;
  p.mean[0] = row[0] / cnt; p.mean[1] = row[1] / cnt; p.mean[2] = row[2] / cnt;
  #pragma unroll
  for (int c = 0; c < 6; c++) p.cov[c] = row[3 + c] / cnt;
  p.w = sqrt(cnt);
  vox[id] = p;
}

}  // namespace

cudaError_t launch_voxel_build(const VoxelBuild& b, cudaStream_t stream) {
  cudaError_t e;
  const size_t cap = (size_t)b.mask + 1;
  if ((e = cudaMemsetAsync(b.table, 0xFF, cap * sizeof(VoxSlot), stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(b.counters, 0, 2 * sizeof(int), stream)) != cudaSuccess) return e;
  if (b.n == 0) return cudaSuccess;
  const int threads = 256;
  const int blocks = (b.n + threads - 1) / threads;
  vox_insert_kernel<<<blocks, threads, 0, stream>>>(b.pts, b.n, b.grid, b.table, b.mask, b.slot_of_point, b.counters, b.sums, 0);
  vox_accumulate_kernel<<<blocks, threads, 0, stream>>>(b.pts, b.cov6, b.n, b.slot_of_point, b.table, b.sums);
  vox_finalize_kernel<<<blocks, threads, 0, stream>>>(b.sums, b.counters, b.vox, b.n);
  return cudaGetLastError();
}

}  // namespace rolo
