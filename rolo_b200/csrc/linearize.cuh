// linearize.cuh — launch interface of the fused correspondence + Mahalanobis + Jacobian + reduction kernel.
#pragma once
#include "common.cuh"
#include <vector>

namespace rolo {

enum LinMode { LIN_SE3 = 0, LIN_SO3 = 1, LIN_COST = 2 };

// Static (per align) description of one scan pair.
struct LinStatic {
  const float4* src;          // [n_pad] source points
  const float* cov6;          // tile-planar source covariances
  int n;                      // source points
  unsigned mask;              // table capacity - 1
  const VoxSlot* table;
  const VoxPayload* vox;
  int* corr;                  // [n * noff] cached compact voxel id (-1 = none)
  double* partials;           // [CTAs touching this pair][RES_N] (capacity: the launch grid)
  unsigned* counter;          // ticket for the last-block reduction (zero before launch, self-resetting)
  double* result;             // [RES_N] (device memory, or mapped pinned host memory)
  unsigned* done_flag;        // optional mapped host flag: receives `seq` after result is visible
  GridParams grid;
  // point-sharded registration fused with its reduction over NVLink peer memory (optional, shard_world <= 1: off).
  // peer_xchg[p] is rank p's exchange buffer (CUDA-IPC mapped): [2 parities][XCHG_MAX_RANKS source ranks][XCHG_STRIDE doubles];
  // the last CTA stores this rank's 32 sums + a sequence word into every peer's buffer, waits for every peer's row in its
  // own buffer and adds the rows in rank order: all ranks obtain bit-identical totals without an NCCL call or a host copy.
  int shard_rank, shard_world;
  double* peer_xchg[8];
};
constexpr int XCHG_MAX_RANKS = 8;
constexpr int XCHG_STRIDE = 40;      // 32 values, 1 sequence word, pad (320 B rows)
constexpr size_t XCHG_DOUBLES = (size_t)2 * XCHG_MAX_RANKS * XCHG_STRIDE;

// Dynamic (per evaluation) part: poses and continuous-time terms.
struct LinDyn {
  double R[9];                // rotation applied to source points (row-major)
  double t[3];                // translation applied to source points
  double Rm[9];               // rotation of the linearisation pose: M = (B + Rm A Rm^T)^-1 (frozen for LM trials / t3)
  double ct[3];               // continuous-time residual vector (t3 modes), else 0
  double lam_c;               // lambda / pt_size (t3 modes), else 0
  double h_scale;             // 1 + lam_c / dtn^2
  double b_ct_scale;          // lam_c / dtn
  unsigned seq;               // value written to done_flag
  unsigned shard_seq;         // evaluation counter of a sharded registration (identical on every rank)
};

struct LinLaunch {
  int mode;                   // LinMode
  int refresh;                // 1: probe the voxel table and rewrite corr; 0: use cached corr
  int noff;                   // 1, 7 or 27
  int ct;                     // 1: continuous-time term active (translation stage evaluations)
  int n_pairs;
  int grid;                   // persistent CTAs; CTA g owns tiles [g*T/G, (g+1)*T/G) of the batch's global tile sequence
  int total_tiles;            // T
  const LinStatic* st;        // device array [n_pairs]
  const LinDyn* dyn;          // device array [n_pairs]
  const int* sched;           // device int[3*n_pairs + 1] from build_schedule(): tile_prefix | cta_first | cta_count
};

// host: per-pair tile prefix and the range of CTAs that touch each pair, for `grid` persistent CTAs
void build_schedule(const int* n_points, int n_pairs, int grid, std::vector<int>& sched, int& total_tiles);

// dyn0: host copy of the pose block, passed by value as a kernel parameter when L.dyn == nullptr
// (single-pair path: no H2D copy per evaluation). Ignored when L.dyn is a device array.
cudaError_t launch_linearize(const LinLaunch& L, const LinDyn* dyn0, cudaStream_t stream);
int linearize_max_blocks_per_sm();

// Expand the packed 32-double result into H (6x6 or 3x3 row-major), b and cost.
inline void unpack_result(const double* r, int ndof, double* H, double* b, double* cost) {
  if (cost) *cost = r[27];
  if (!H || !b) return;
  const double* Mw = r; const double* G = r + 6; const double* StG = r + 15; const double* q = r + 21; const double* Stq = r + 24;
  if (ndof == 3) {
    H[0] = StG[0]; H[1] = StG[1]; H[2] = StG[2];
    H[3] = StG[1]; H[4] = StG[3]; H[5] = StG[4];
    H[6] = StG[2]; H[7] = StG[4]; H[8] = StG[5];
    b[0] = Stq[0]; b[1] = Stq[1]; b[2] = Stq[2];
    return;
  }
  const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    H[i * 6 + j] = StG[sym[i][j]];              // S^T M S
    H[(3 + i) * 6 + 3 + j] = Mw[sym[i][j]];     // M
    H[(3 + i) * 6 + j] = -G[i * 3 + j];         // -M S
    H[j * 6 + 3 + i] = -G[i * 3 + j];           // -(M S)^T
  }
  for (int i = 0; i < 3; i++) { b[i] = Stq[i]; b[3 + i] = -q[i]; }
}

}  // namespace rolo
