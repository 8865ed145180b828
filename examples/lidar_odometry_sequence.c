/* lidar_odometry_sequence.c — the call sequence of LidarOdometry::scanRegeistration
 * (src/lidarOdometry.cpp:448-501) written against the C-ABI (plain C, no PCL/Eigen/torch):
 *   construct per frame -> setPolarResolution -> setInputTarget/Source -> align -> computeTranslation.
 * Build:  gcc -O2 -I include examples/lidar_odometry_sequence.c -o build/lidar_odometry_sequence \
 *             -L rolo_b200/lib -lrolo_gicp_b200 -Wl,-rpath,$PWD/rolo_b200/lib -lm
 * Run:    build/lidar_odometry_sequence target.bin source.bin   (float32 xyzw records; tools/dump_pair.py writes them)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rolo_gicp.h"

static float* read_cloud(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  float* p = (float*)malloc((size_t)bytes);
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
  if (argc < 3) { fprintf(stderr, "usage: %s target.bin source.bin\n", argv[0]); return 2; }
  size_t nt = 0, ns = 0;
  float* target = read_cloud(argv[1], &nt);
  float* source = read_cloud(argv[2], &ns);

  for (int frame = 0; frame < 3; frame++) {           /* a fresh object every frame, like lidarOdometry.cpp:460 */
    rg_ctx* ctx = NULL;
    if (rg_create(-1, &ctx) != RG_OK) { fprintf(stderr, "rg_create: %s\n", rg_last_error(NULL)); return 1; }
    CHECK(rg_set_polar_resolution(ctx, 0.175, 0.175, 2.0));
    CHECK(rg_set_num_threads(ctx, 0));
    CHECK(rg_clear_target(ctx));
    CHECK(rg_clear_source(ctx));
    CHECK(rg_set_input_target(ctx, target, nt, 16));
    CHECK(rg_set_input_source(ctx, source, ns, 16));
    float T[16];
    CHECK(rg_align(ctx, NULL, T));                      /* rotation estimate (default SO3-LM) */
    double reg_t[3] = {0, 0, 0};
    const double guess[3] = {T[3], T[7], T[11]}, last[3] = {0, 0, 0};
    CHECK(rg_compute_translation(ctx, reg_t, guess, last, 0.1, 0.1, 0.3f));
    printf("frame %d: converged=%d iters=%d  R row0 = [% .6f % .6f % .6f]  reg_translation = [% .6f % .6f % .6f]\n", frame,
           rg_has_converged(ctx), rg_nr_iterations(ctx), T[0], T[1], T[2], reg_t[0], reg_t[1], reg_t[2]);
    rg_destroy(ctx);
  }
  free(target); free(source);
  return 0;
}
