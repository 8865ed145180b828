// scancontext.cuh — launch interface of the Scan Context descriptor kernel (src/scancontext/Scancontext.cpp:151-227).
#pragma once
#include "common.cuh"

namespace rolo {

constexpr int SC_NUM_RING = 20;        // Scancontext.h:82-84
constexpr int SC_NUM_SECTOR = 60;
constexpr int SC_BINS = SC_NUM_RING * SC_NUM_SECTOR;

struct SCLaunch {
  const float4* pts; int n;            // scan in the lidar frame (w ignored)
  double lidar_height;                 // LIDAR_HEIGHT (Scancontext.h:80)
  double max_radius;                   // PC_MAX_RADIUS
  int* bins;                           // [SC_BINS] scratch (ordered-int encoded float maxima)
  double* desc;                        // out [SC_NUM_RING][SC_NUM_SECTOR] row-major
  double* ring_key;                    // out [SC_NUM_RING]   row means   (makeRingkeyFromScancontext)
  double* sector_key;                  // out [SC_NUM_SECTOR] column means (makeSectorkeyFromScancontext)
};
cudaError_t launch_scancontext(const SCLaunch& L, cudaStream_t stream);

}  // namespace rolo
