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
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace rolo {

namespace {

// a build descriptor reaches the kernels either from a device array (batch: one per blockIdx.y) or by value
struct BuildRef {
  const VoxelBuild* arr;
  VoxelBuild one;
};
__device__ __forceinline__ const VoxelBuild& get_build(const BuildRef& r) { return r.arr ? r.arr[blockIdx.y] : r.one; }

__global__ void vox_clear_kernel(BuildRef ref) {
  const VoxelBuild& B = get_build(ref);
  const size_t cap = (size_t)B.mask + 1;
  if (blockIdx.x == 0 && threadIdx.x < 2) B.counters[threadIdx.x] = 0;
  ulonglong2* t = reinterpret_cast<ulonglong2*>(B.table);
  for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (size_t)gridDim.x * blockDim.x)
    t[s] = make_ulonglong2(KEY_EMPTY, 0xFFFFFFFFFFFFFFFFull);
}

__global__ void vox_insert_kernel(BuildRef ref) {
  const VoxelBuild& B = get_build(ref);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // work item: point i, or sorted position i of a restricted build
  if (i >= B.n) return;
  VoxSlot* table = B.table;
  const unsigned mask = B.mask;
  int* counters = B.counters;
  if (B.order && !knn_owns(i, B.shard_rank, B.shard_world)) { B.slot_of_point[i] = -1; return; }
  const float4 p = B.order ? B.order[i] : B.pts[i];
  int cx, cy, cz;
  voxel_coord(B.grid, (double)p.x, (double)p.y, (double)p.z, cx, cy, cz);
  // a point without a finite key (NaN / Inf coordinates, the origin of a polar grid) sits in an unreachable cell of the
  // reference's map (x86 converts the NaN to INT_MIN): it is simply not inserted
  const bool finite = (p.x - p.x == 0.f) && (p.y - p.y == 0.f) && (p.z - p.z == 0.f) &&
                      (B.grid.voxel_type != 0 || (p.x != 0.f || p.y != 0.f || p.z != 0.f));
  if (!finite) { B.slot_of_point[i] = -1; return; }
  if (!key_in_range(cx, cy, cz)) { atomicExch(&counters[1], 1); B.slot_of_point[i] = -1; return; }
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
        double* row = B.sums + (size_t)id * 10;
        #pragma unroll
        for (int k = 0; k < 10; k++) row[k] = 0.0;
        B.slot_of_point[i] = (int)h;
        return;
      }
    }
    if (cur == key) { B.slot_of_point[i] = (int)h; return; }
    h = (h + 1) & mask;
  }
  atomicExch(&counters[1], 2);   // table full (cannot happen with cap >= 2n)
  B.slot_of_point[i] = -1;
}

__global__ void vox_accumulate_kernel(BuildRef ref) {
  const VoxelBuild& B = get_build(ref);
  const int n = B.n;
  if ((int)(blockIdx.x * blockDim.x) >= n) return;          // block-uniform
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  int id = -1;
  double v[10];
  #pragma unroll
  for (int k = 0; k < 10; k++) v[k] = 0.0;
  if (t < n) {
    const int slot = B.slot_of_point[t];
    if (slot >= 0) {
      id = B.table[slot].id;
      const float4 p = B.order ? B.order[t] : B.pts[t];
      const int i = B.order ? __float_as_int(p.w) : t;      // the covariance lives at the point's original index
      const size_t base = (size_t)(i / TILE_PTS) * COV_PLANES * TILE_PTS + (i % TILE_PTS);
      v[0] = (double)p.x; v[1] = (double)p.y; v[2] = (double)p.z;
      #pragma unroll
      for (int c = 0; c < 6; c++) v[3 + c] = (double)B.cov6[base + (size_t)c * TILE_PTS];
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
    double* row = B.sums + (size_t)id * 10;
    #pragma unroll
    for (int k = 0; k < 10; k++) atomicAdd(row + k, v[k]);
  }
}

__global__ void vox_finalize_kernel(BuildRef ref) {
  const VoxelBuild& B = get_build(ref);
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= B.n || id >= B.counters[0]) return;
  const double* row = B.sums + (size_t)id * 10;
  const double cnt = row[9];
  VoxPayload p;
  p.mean[0] = row[0] / cnt; p.mean[1] = row[1] / cnt; p.mean[2] = row[2] / cnt;
  #pragma unroll
  for (int c = 0; c < 6; c++) p.cov[c] = row[3 + c] / cnt;
  p.w = sqrt(cnt);
  B.vox[id] = p;
}

__global__ void vox_keys_kernel(const VoxSlot* __restrict__ table, size_t cap, unsigned long long* __restrict__ keys) {
  for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (size_t)gridDim.x * blockDim.x) {
    const VoxSlot sl = table[s];
    if (sl.key != KEY_EMPTY && sl.id >= 0) keys[sl.id] = sl.key;
  }
}

// import, pass 1: every gathered entry claims (or finds) the slot of its key; the creator takes the next id and zeroes its row
__global__ void vox_import_insert_kernel(VoxelBuild B, const unsigned long long* __restrict__ keys, int n_entries, int* __restrict__ slot_of_entry) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const unsigned long long key = keys[e];
  unsigned h = hash_key(key) & B.mask;
  for (unsigned probe = 0; probe <= B.mask; probe++) {
    unsigned long long* kp = &B.table[h].key;
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
    if (cur == KEY_EMPTY) {
      cur = atomicCAS(kp, KEY_EMPTY, key);
      if (cur == KEY_EMPTY) {
        const int id = atomicAdd(&B.counters[0], 1);
        B.table[h].id = id;
        double* row = B.sums + (size_t)id * 10;
        #pragma unroll
        for (int k = 0; k < 10; k++) row[k] = 0.0;
        slot_of_entry[e] = (int)h;
        return;
      }
    }
    if (cur == key) { slot_of_entry[e] = (int)h; return; }
    h = (h + 1) & B.mask;
  }
  atomicExch(&B.counters[1], 2);
  slot_of_entry[e] = -1;
}
// import, pass 2: add the entry's 10 sums to its voxel's row
__global__ void vox_import_add_kernel(VoxelBuild B, const double* __restrict__ sums_in, int n_entries, const int* __restrict__ slot_of_entry) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = t / 10, k = t % 10;
  if (e >= n_entries) return;
  const int slot = slot_of_entry[e];
  if (slot < 0) return;
  atomicAdd(B.sums + (size_t)B.table[slot].id * 10 + k, sums_in[(size_t)e * 10 + k]);
}

// ---- deterministic build --------------------------------------------------------------------------------------------------
__global__ void vox_det_key_kernel(VoxelBuild B, unsigned long long* __restrict__ keys, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n) return;
  const float4 p = B.pts[i];
  int cx, cy, cz;
  voxel_coord(B.grid, (double)p.x, (double)p.y, (double)p.z, cx, cy, cz);
  const bool finite = (p.x - p.x == 0.f) && (p.y - p.y == 0.f) && (p.z - p.z == 0.f) &&
                      (B.grid.voxel_type != 0 || (p.x != 0.f || p.y != 0.f || p.z != 0.f));
  unsigned long long key = KEY_EMPTY;                       // sorts behind every real key (21-bit fields: the top bit is never set)
  if (finite) {
    if (key_in_range(cx, cy, cz)) key = pack_key(cx, cy, cz); else atomicExch(&B.counters[1], 1);
  }
  keys[i] = key; idx[i] = i;
}
__global__ void vox_det_head_kernel(const unsigned long long* __restrict__ keys, int n, int* __restrict__ flags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  flags[j] = (j < n && keys[j] != KEY_EMPTY && (j == 0 || keys[j] != keys[j - 1])) ? 1 : 0;   // flags[n] = 0: its exclusive sum is the voxel count
}
// one thread per voxel: sequential fp64 sums over its points in index order, payload row, table entry with id = rank of the key
__global__ void vox_det_reduce_kernel(VoxelBuild B, const unsigned long long* __restrict__ keys, const int* __restrict__ idx,
                                      const int* __restrict__ rank /* exclusive sum of the head flags */) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) B.counters[0] = rank[B.n];
  if (j >= B.n) return;
  const unsigned long long key = keys[j];
  if (key == KEY_EMPTY || (j > 0 && keys[j - 1] == key)) return;      // not the first point of a voxel
  const int id = rank[j];
  double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = j; e < B.n && keys[e] == key; e++) {
    const int i = idx[e];
    const float4 p = B.pts[i];
    const size_t base = (size_t)(i / TILE_PTS) * COV_PLANES * TILE_PTS + (i % TILE_PTS);
    v[0] += (double)p.x; v[1] += (double)p.y; v[2] += (double)p.z;
    #pragma unroll
    for (int c = 0; c < 6; c++) v[3 + c] += (double)B.cov6[base + (size_t)c * TILE_PTS];
    v[9] += 1.0;
  }
  double* row = B.sums + (size_t)id * 10;
  #pragma unroll
  for (int k = 0; k < 10; k++) row[k] = v[k];
  VoxPayload pl;
  pl.mean[0] = v[0] / v[9]; pl.mean[1] = v[1] / v[9]; pl.mean[2] = v[2] / v[9];
  #pragma unroll
  for (int c = 0; c < 6; c++) pl.cov[c] = v[3 + c] / v[9];
  pl.w = sqrt(v[9]);
  B.vox[id] = pl;
  unsigned h = hash_key(key) & B.mask;                              // keys are unique here: plain linear probing with CAS
  for (unsigned probe = 0; probe <= B.mask; probe++) {
    if (atomicCAS(&B.table[h].key, KEY_EMPTY, key) == KEY_EMPTY) { B.table[h].id = id; return; }
    h = (h + 1) & B.mask;
  }
  atomicExch(&B.counters[1], 2);
}

cudaError_t launch_all(const BuildRef& ref, int n_builds, int max_n, unsigned max_mask, cudaStream_t stream, bool finalize = true) {
  if (n_builds <= 0) return cudaSuccess;
  const int threads = 256;
  const unsigned nb = (unsigned)std::max(1, (max_n + threads - 1) / threads);
  const size_t cap = (size_t)max_mask + 1;
  const unsigned cb = (unsigned)std::min<size_t>((cap + threads - 1) / threads, 2048);
  vox_clear_kernel<<<dim3(cb, n_builds), threads, 0, stream>>>(ref);
  if (max_n > 0) {
    vox_insert_kernel<<<dim3(nb, n_builds), threads, 0, stream>>>(ref);
    vox_accumulate_kernel<<<dim3(nb, n_builds), threads, 0, stream>>>(ref);
    if (finalize) vox_finalize_kernel<<<dim3(nb, n_builds), threads, 0, stream>>>(ref);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_voxel_build(const VoxelBuild& b, cudaStream_t stream, bool finalize) {
  BuildRef ref; ref.arr = nullptr; ref.one = b;
  return launch_all(ref, 1, b.n, b.mask, stream, finalize);
}

size_t voxel_det_temp_bytes(int n) {
  size_t a = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, n, 0, 64);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (const int*)nullptr, (int*)nullptr, n + 1);
  return std::max(a, c);
}

cudaError_t launch_voxel_build_deterministic(const VoxelBuild& b, const VoxelDetScratch& s, cudaStream_t stream) {
  if (b.order) return cudaErrorInvalidValue;
  BuildRef ref; ref.arr = nullptr; ref.one = b;
  const size_t cap = (size_t)b.mask + 1;
  vox_clear_kernel<<<dim3((unsigned)std::min<size_t>((cap + 255) / 256, 2048), 1), 256, 0, stream>>>(ref);
  if (b.n <= 0) return cudaGetLastError();
  const int nb = (b.n + 255) / 256;
  vox_det_key_kernel<<<nb, 256, 0, stream>>>(b, s.keys_a, s.idx_a);
  size_t temp = s.cub_temp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(s.cub_temp, temp, s.keys_a, s.keys_b, s.idx_a, s.idx_b, b.n, 0, 64, stream);   // stable
  if (e != cudaSuccess) return e;
  vox_det_head_kernel<<<(b.n + 1 + 255) / 256, 256, 0, stream>>>(s.keys_b, b.n, s.flags);
  temp = s.cub_temp_bytes;
  int* rank = s.idx_a;                                                   // idx_a is free after the sort: [n + 1] needed -> the caller sizes it n + 1
  e = cub::DeviceScan::ExclusiveSum(s.cub_temp, temp, s.flags, rank, b.n + 1, stream);
  if (e != cudaSuccess) return e;
  vox_det_reduce_kernel<<<nb, 256, 0, stream>>>(b, s.keys_b, s.idx_b, rank);
  return cudaGetLastError();
}

cudaError_t launch_voxel_keys(const VoxelBuild& b, unsigned long long* keys, cudaStream_t stream) {
  const size_t cap = (size_t)b.mask + 1;
  vox_keys_kernel<<<(unsigned)std::min<size_t>((cap + 255) / 256, 2048), 256, 0, stream>>>(b.table, cap, keys);
  return cudaGetLastError();
}

cudaError_t launch_voxel_import(const VoxelBuild& b, const unsigned long long* keys, const double* sums_in, int n_entries, cudaStream_t stream) {
  BuildRef ref; ref.arr = nullptr; ref.one = b;
  const size_t cap = (size_t)b.mask + 1;
  vox_clear_kernel<<<dim3((unsigned)std::min<size_t>((cap + 255) / 256, 2048), 1), 256, 0, stream>>>(ref);
  if (n_entries > 0) {
    vox_import_insert_kernel<<<(n_entries + 255) / 256, 256, 0, stream>>>(b, keys, n_entries, b.slot_of_point);
    vox_import_add_kernel<<<(n_entries * 10 + 255) / 256, 256, 0, stream>>>(b, sums_in, n_entries, b.slot_of_point);
    BuildRef fin = ref; fin.one.order = nullptr; fin.one.n = std::min(n_entries, b.n);
    vox_finalize_kernel<<<dim3((unsigned)((fin.one.n + 255) / 256), 1), 256, 0, stream>>>(fin);
  }
  return cudaGetLastError();
}

// every target of a batch in one launch chain: `dev_builds` is the device copy of `host_builds`
cudaError_t launch_voxel_build_batch(const VoxelBuild* host_builds, VoxelBuild* dev_builds, int n_builds, cudaStream_t stream) {
  if (n_builds <= 0) return cudaSuccess;
  int max_n = 0; unsigned max_mask = 0;
  for (int i = 0; i < n_builds; i++) { max_n = std::max(max_n, host_builds[i].n); max_mask = std::max(max_mask, host_builds[i].mask); }
  cudaError_t e = cudaMemcpyAsync(dev_builds, host_builds, sizeof(VoxelBuild) * n_builds, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  BuildRef ref; ref.arr = dev_builds; ref.one = host_builds[0];
  return launch_all(ref, n_builds, max_n, max_mask, stream);
}

}  // namespace rolo
