/* back_end_sequence.c — the back-end (backMapping) call sequence written against the C-ABI in plain C:
 *   scan-to-map optimisation   (src/backMapping.cpp:681-710)   rg_s2m_set_map / rg_s2m_set_scan / rg_s2m_optimize
 *   Scan Context bookkeeping   (:1184-1213, 2381-2392)         rg_sc_make_and_save / rg_sc_detect
 *   loop-closure verification  (:2331-2352)                    GICP core in 6-DoF LM mode + rg_get_fitness_score
 * Build:  gcc -O2 -I include examples/back_end_sequence.c -o build/back_end_sequence \
 *             -L rolo_b200/lib -lrolo_gicp_b200 -Wl,-rpath,$PWD/rolo_b200/lib -lm
 * Run:    build/back_end_sequence map_corner.bin map_surf.bin scan_corner.bin scan_surf.bin roll pitch yaw x y z
 *         (float32 xyzw records; the six numbers are the odometry guess, the reference's transformTobeMapped)
 */
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include "rolo_gicp.h"

static float* read_cloud(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  float* p = (float*)malloc(bytes > 0 ? (size_t)bytes : 16);
  if (fread(p, 1, (size_t)bytes, f) != (size_t)bytes) { perror("fread"); exit(2); }
  fclose(f);
  *n = (size_t)bytes / 16;
  return p;
}

#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != RG_OK && rc_ != RG_ERR_NOT_CONVERGED) {                                \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, rg_last_error(ctx));        \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 11) { fprintf(stderr, "usage: %s map_corner.bin map_surf.bin scan_corner.bin scan_surf.bin roll pitch yaw x y z\n", argv[0]); return 2; }
  size_t nmc, nms, nsc, nss;
  float* map_corner = read_cloud(argv[1], &nmc);
  float* map_surf = read_cloud(argv[2], &nms);
  float* scan_corner = read_cloud(argv[3], &nsc);
  float* scan_surf = read_cloud(argv[4], &nss);
  float tf[6];
  for (int k = 0; k < 6; k++) tf[k] = (float)atof(argv[5 + k]);

  rg_ctx* ctx = NULL;
  if (rg_create(-1, &ctx) != RG_OK) { fprintf(stderr, "rg_create: %s\n", rg_last_error(NULL)); return 1; }

  /* scan2MapOptimization */
  CHECK(rg_s2m_set_map(ctx, map_corner, nmc, map_surf, nms, 16));
  CHECK(rg_s2m_set_scan(ctx, scan_corner, nsc, scan_surf, nss, 16));
  rg_s2m_result res;
  CHECK(rg_s2m_optimize(ctx, tf, 30, /*edgeFeatureMinValidNum*/ 10, /*surfFeatureMinValidNum*/ 100, &res));
  printf("scan2map iterations=%d converged=%d degenerate=%d enough=%d pose=[% .6f % .6f % .6f % .6f % .6f % .6f]\n", res.iterations, res.converged,
         res.degenerate, res.enough_features, tf[0], tf[1], tf[2], tf[3], tf[4], tf[5]);

  /* Scan Context: one descriptor per keyframe, then the loop query */
  CHECK(rg_sc_make_and_save(ctx, scan_surf, nss, 16, /*LIDAR_HEIGHT*/ 2.0));
  int32_t loop_id; float yaw_diff;
  CHECK(rg_sc_detect(ctx, &loop_id, &yaw_diff));
  printf("scancontext entries=%d loop_id=%d yaw_diff=% .6f\n", rg_sc_size(ctx), (int)loop_id, yaw_diff);

  /* loop-closure verification: align the scan's surface features to the map's with the GICP core (LM), accept on fitness */
  CHECK(rg_set_polar_resolution(ctx, 0.175, 0.175, 2.0));
  CHECK(rg_set_optimizer_type(ctx, RG_OPT_LM));
  CHECK(rg_set_maximum_iterations(ctx, 100));
  CHECK(rg_set_input_target(ctx, map_surf, nms, 16));
  CHECK(rg_set_input_source(ctx, scan_surf, nss, 16));
  float T[16];
  CHECK(rg_align(ctx, NULL, T));
  double fitness = 0.0;
  CHECK(rg_get_fitness_score(ctx, T, DBL_MAX, &fitness));
  printf("loop_check converged=%d fitness=%.9e\n", rg_has_converged(ctx), fitness);

  rg_destroy(ctx);
  free(map_corner); free(map_surf); free(scan_corner); free(scan_surf);
  return 0;
}
