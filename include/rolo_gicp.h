/*
 * rolo_gicp.h — C-ABI of the B200-native voxelized-GICP core (librolo_gicp_b200.so).
 *
 * This is the drop-in boundary for the one data-parallel hot path of sdwyc/ROLO: the
 * fast_gicp::RotVGICP scan matcher ("rot_gicp") plus the two upstream range-image passes.
 * Every entry point replaces one reference interface (cited file:line, paths relative to the
 * reference checkout). The C++ class template in include/rot_gicp/gicp/rot_vgicp.hpp forwards to
 * these; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; all functions return 0 on success and a
 *    negative rg_status on failure (no exception crosses the ABI); rg_last_error() gives text.
 *  - matrices are ROW-MAJOR. Poses are 4x4 (16 values); H6 = 6x6 (36), b6 = 6; H3 = 3x3 (9), b3 = 3.
 *    (Eigen is column-major: the C++ wrapper transposes; H is symmetric.)
 *  - "host" pointers are ordinary host memory; "_device" variants take device pointers of the
 *    context's GPU and a cudaStream_t passed as void* (0 = the context's own stream).
 *  - enum values equal the reference's enum class order (gicp_settings.hpp:6-12,
 *    lsq_registration.hpp:13) so the C++ wrapper can static_cast.
 *  - the CUDA device is mandatory: there is NO CPU fallback; rg_create fails without a GPU.
 */
#ifndef ROLO_GICP_H
#define ROLO_GICP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden: only this ABI is exported */
#endif

#define ROLO_GICP_ABI_VERSION 1

typedef struct rg_ctx rg_ctx;      /* one registration object == one fast_gicp::RotVGICP instance   */
typedef struct rg_batch rg_batch;  /* a lock-step batch of independent scan pairs on one GPU         */

typedef enum rg_status {
  RG_OK = 0,
  RG_ERR_INVALID = -1,      /* bad argument / call order (maps to std::invalid_argument)            */
  RG_ERR_CUDA = -2,         /* CUDA runtime error (maps to std::runtime_error)                      */
  RG_ERR_NO_DEVICE = -3,    /* no usable GPU: the product path refuses to run                       */
  RG_ERR_UNSUPPORTED = -4,  /* enum value the reference abort()s on (voxel.hpp:16-18, impl.hpp:472) */
  RG_ERR_NOT_CONVERGED = -5 /* "lm not converged!!" break (lsq.hpp:66-69,168-171); pose still valid */
} rg_status;

/* gicp_settings.hpp:6 */
enum { RG_REG_NONE = 0, RG_REG_MIN_EIG, RG_REG_NORMALIZED_MIN_EIG, RG_REG_PLANE, RG_REG_FROBENIUS, RG_REG_PLANE_S };
/* gicp_settings.hpp:8 */
enum { RG_DIRECT27 = 0, RG_DIRECT7, RG_DIRECT1, RG_DIRECT_RADIUS };
/* gicp_settings.hpp:10 */
enum { RG_ACC_ADDITIVE = 0, RG_ACC_ADDITIVE_WEIGHTED, RG_ACC_MULTIPLICATIVE };
/* gicp_settings.hpp:12 */
enum { RG_VOXEL_POLAR = 0, RG_VOXEL_UNIFORM };
/* lsq_registration.hpp:13 */
enum { RG_OPT_GAUSS_NEWTON = 0, RG_OPT_LM, RG_OPT_SO3_LM };
enum { RG_SOURCE = 0, RG_TARGET = 1 };

/* ---- lifetime -------------------------------------------------------------------------------- */
int  rg_abi_version(void);
/* RotVGICP::RotVGICP() (rot_vgicp_impl.hpp:20-39) + LsqRegistration() (lsq_registration_impl.hpp:9-24):
 * same defaults. device < 0 => current CUDA device. Device buffers come from a process-global pool
 * because the caller constructs a fresh object every frame (lidarOdometry.cpp:460). */
int  rg_create(int device, rg_ctx** out);
void rg_destroy(rg_ctx* ctx);                               /* ~RotVGICP (rot_vgicp_impl.hpp:42)     */
const char* rg_last_error(const rg_ctx* ctx);               /* ctx may be NULL: last global error    */
int  rg_device_count(void);

/* ---- setters (rot_vgicp.hpp:72-93, lsq_registration.hpp:51-62, pcl::Registration) -------------- */
int rg_set_resolution(rg_ctx*, double resolution);                         /* rot_vgicp_impl.hpp:45-49  */
int rg_set_polar_resolution(rg_ctx*, double theta, double phi, double r);  /* rot_vgicp_impl.hpp:52-56  */
int rg_set_voxel_accumulation_mode(rg_ctx*, int mode);                     /* rot_vgicp_impl.hpp:64-66  */
int rg_set_neighbor_search_method(rg_ctx*, int method);                    /* rot_vgicp_impl.hpp:59-61  */
int rg_set_num_threads(rg_ctx*, int n);                                    /* rot_vgicp_impl.hpp:79-87 (kept for API parity; ignored) */
int rg_set_correspondence_randomness(rg_ctx*, int k);                      /* rot_vgicp_impl.hpp:90-92  */
int rg_set_regularization_method(rg_ctx*, int method);                     /* rot_vgicp_impl.hpp:95-97  */
int rg_set_optimizer_type(rg_ctx*, int type);                              /* lsq_registration_impl.hpp:338-340 */
int rg_set_rotation_epsilon(rg_ctx*, double eps);                          /* lsq_registration_impl.hpp:30-32   */
int rg_set_transformation_epsilon(rg_ctx*, double eps);                    /* pcl::Registration::setTransformationEpsilon */
int rg_set_maximum_iterations(rg_ctx*, int n);                             /* pcl::Registration::setMaximumIterations     */
int rg_set_initial_lambda_factor(rg_ctx*, double f);                       /* lsq_registration_impl.hpp:35-37   */
int rg_set_debug_print(rg_ctx*, int on);                                   /* lsq_registration_impl.hpp:40-42   */

/* ---- clouds ------------------------------------------------------------------------------------ */
/* setInputSource / setInputTarget (rot_vgicp_impl.hpp:111-120, 132-143). xyz are the first three
 * floats of each point; stride_bytes = sizeof(PointT) (32 for pcl::PointXYZI, 16 for PointXYZ).
 * The H2D copy is queued on the context's stream: from pageable memory the host buffer is consumed before the call
 * returns; from page-locked memory the copy is asynchronous and the buffer must stay valid and unmodified until the next
 * solving call (rg_align, rg_linearize, rg_calculate_covariances, ...) has returned — pcl::Registration keeps the cloud's
 * shared_ptr for exactly that long in the reference.                                               */
int rg_set_input_source(rg_ctx*, const float* host_points, size_t n, size_t stride_bytes);
int rg_set_input_target(rg_ctx*, const float* host_points, size_t n, size_t stride_bytes);
/* same, from device-resident float4 (x,y,z,*) arrays: no host traffic */
int rg_set_input_source_device(rg_ctx*, const float* dev_xyzw, size_t n, void* stream);
int rg_set_input_target_device(rg_ctx*, const float* dev_xyzw, size_t n, void* stream);
int rg_clear_source(rg_ctx*);                                              /* rot_vgicp_impl.hpp:99-103  */
int rg_clear_target(rg_ctx*);                                              /* rot_vgicp_impl.hpp:105-109 */
int rg_swap_source_and_target(rg_ctx*);                                    /* rot_vgicp_impl.hpp:68-76   */
size_t rg_source_size(const rg_ctx*);
size_t rg_target_size(const rg_ctx*);

/* setSourceCovariances / setTargetCovariances (rot_vgicp_impl.hpp:122-130): n 4x4 double matrices */
int rg_set_covariances(rg_ctx*, int which, const double* cov4x4, size_t n);
/* getSourceCovariances / getTargetCovariances (rot_vgicp.hpp:94-100): returns count; writes n*16
 * doubles when out != NULL. Device storage is 6 x fp32 per point.               */
size_t rg_get_covariances(rg_ctx*, int which, double* out_cov4x4);

/* ---- stages of computeTransformation ------------------------------------------------------------ */
/* RotVGICP::calculate_covariances (rot_vgicp_impl.hpp:421-496): exact kNN(k) + regularised covariance */
int rg_calculate_covariances(rg_ctx*, int which);
/* VmfVoxelMap::create_voxelmap (vmp_voxel.hpp:167-197) over the target cloud + covariances        */
int rg_create_voxelmap(rg_ctx*);
/* RotVGICP::linearize (rot_vgicp_impl.hpp:224-290) = update_correspondences (172-222) + 6-DoF normal
 * equations, one fused kernel. H36/b6 may be NULL (cost only, as evaluateCost allows).          */
int rg_linearize(rg_ctx*, const double T[16], double* H36, double* b6, double* cost);
/* RotVGICP::so3_linearize (rot_vgicp_impl.hpp:292-388)                                             */
int rg_so3_linearize(rg_ctx*, const double T[16], double* H9, double* b3, double* cost);
/* RotVGICP::compute_error (rot_vgicp_impl.hpp:390-417): cached correspondences + Mahalanobis      */
int rg_compute_error(rg_ctx*, const double T[16], double* cost);
/* RotVGICP::t3_linearize (rot_vgicp_impl.hpp:498-607) / compute_t_error (609-658)                  */
int rg_t3_linearize(rg_ctx*, const double t[3], const double init_guess[3], const double last_t0[3],
                    double interval_tn, double interval_tn_1, double* H36, double* b6, double* cost);
int rg_compute_t_error(rg_ctx*, const double t[3], const double init_guess[3], const double last_t0[3],
                       double interval_tn, double interval_tn_1, double* cost);
int rg_set_ct_lambda(rg_ctx*, float lambda);                 /* lambda_ member, rot_vgicp_impl.hpp:166 */
size_t rg_num_correspondences(const rg_ctx*);                /* voxel_correspondences_.size()          */

/* ---- drivers -------------------------------------------------------------------------------------- */
/* pcl::Registration::align(output, guess) -> RotVGICP::computeTransformation (rot_vgicp_impl.hpp:146-160)
 * -> LsqRegistration::computeTransformation (lsq_registration_impl.hpp:152-179). The GN/LM step logic
 * (6x6 / 3x3 LDLT, damping, convergence) runs on the host, one fused kernel per evaluation.
 * Returns RG_OK or RG_ERR_NOT_CONVERGED (pose is written in both cases, as the reference does).  */
int rg_align(rg_ctx*, const float guess[16], float final_transformation[16]);
/* RotVGICP::computeTranslation (rot_vgicp_impl.hpp:162-169 -> lsq_registration_impl.hpp:55-80)     */
int rg_compute_translation(rg_ctx*, double trans[3], const double init_guess[3], const double last_t0[3],
                           double interval_tn, double interval_tn_1, float ct_lambda);
/* pcl::transformPointCloud(*input_, output, T) (lsq_registration_impl.hpp:78,178): writes n points,
 * xyz transformed, to host memory with the given stride (other bytes untouched).                  */
int rg_transform_source(rg_ctx*, const float T[16], float* host_out, size_t stride_bytes);
int rg_has_converged(const rg_ctx*);                          /* pcl::Registration::hasConverged       */
int rg_nr_iterations(const rg_ctx*);
int rg_get_final_hessian(const rg_ctx*, double H36[36]);      /* lsq_registration_impl.hpp:45-47       */
int rg_get_final_pose(const rg_ctx*, double T[16]);           /* x0 in double, before the float cast   */

/* ---- introspection used by the parity tests --------------------------------------------------------- */
size_t rg_voxel_count(const rg_ctx*);
/* voxels sorted by key: keys int32[V*3], mean double[V*3], cov6 double[V*6] (xx,xy,xz,yy,yz,zz), npts int32[V] */
int rg_get_voxels(rg_ctx*, int32_t* keys, double* mean, double* cov6, int32_t* npts);
/* cached correspondence voxel key per (source point, offset): int32[n*noff*3], INT32_MIN rows = none */
int rg_get_correspondence_keys(rg_ctx*, int32_t* keys);
/* exact kNN indices (k nearest incl. self, ascending (distance, index)): int32[n*k] */
int rg_knn(rg_ctx*, int which, int k, int32_t* out_idx);
/* kernel time of the last linearize/compute_error launch in microseconds (CUDA events), or <0 */
double rg_last_kernel_us(const rg_ctx*);
int rg_set_profile(rg_ctx*, int on);

/* ---- batch of independent scan pairs (BASELINE config 3; one rg_batch per GPU, no collective) ------- */
int  rg_batch_create(int device, int n_pairs, rg_batch** out);
void rg_batch_destroy(rg_batch*);
rg_ctx* rg_batch_ctx(rg_batch*, int pair);      /* per-pair context for setters / inputs / results  */
/* align() of every pair in lock step: the covariances are estimated in groups of 32 pairs per launch chain (one radix sort, one
 * kNN + covariance launch per group), every voxel map by one launch chain, then one fused linearize launch per GN/LM evaluation
 * covers all pairs.                                                                              */
int  rg_batch_align(rg_batch*, const float* guesses /* n_pairs*16 or NULL */, float* final_transformations /* n_pairs*16 */);
double rg_batch_last_linearize_us(const rg_batch*);   /* mean device time of the fused launches (CUDA events) */
int    rg_batch_last_linearize_launches(const rg_batch*);
/* stage times of the last rg_batch_align, CUDA events on the batch stream: out[0] covariance estimation (index build + search + epilogue, all
 * groups) ms, out[1] of which the search kernels (chunk + fallback) ms, out[2] voxel maps ms, out[3] iterations ms (first launch to last result,
 * host solves included), out[4] sum of the fused linearize kernels ms, out[5] points whose covariance was estimated, out[6] fused launches */
int    rg_batch_last_stages(const rg_batch*, double out[8]);
void*  rg_batch_stream(const rg_batch*);               /* cudaStream_t the batch launches on (for event timing) */
void*  rg_stream(const rg_ctx*);                       /* cudaStream_t of a stand-alone context           */
/* number of this library's own kernels launched so far in this process (library-internal kernels of
 * cub's radix sort are not counted) */
unsigned long long rg_kernel_launch_count(void);

/* ---- sharded single registration (BASELINE config 4): reduction hook ----------------------------------
 * When set, every linearize/compute_error result (32 doubles: 28 packed sums + count + pad) is passed
 * through `fn` before the host step, so N ranks each holding a shard of the source cloud behave as one
 * registration. on_device != 0: buf is a device pointer (NCCL path); else host pointer (gloo / tests).   */
typedef int (*rg_reduce_fn)(void* user, double* buf, int count, int on_device, void* stream);
int rg_set_reduce_hook(rg_ctx*, rg_reduce_fn fn, void* user, int on_device);
/* Restrict every evaluation (linearize / compute_error / t3) to source points [lo, hi) while covariances stay those of
 * the FULL cloud (kNN neighbourhoods must not be cut at shard borders). lo must be a multiple of 256 (the TMA tile);
 * lo == hi == 0 restores the whole cloud. Reset by rg_set_input_source*. */
int rg_set_source_shard(rg_ctx*, size_t lo, size_t hi);
/* Fused reduction over NVLink peer memory (alternative to the reduce hook, one process per GPU on one node):
 * every rank exports a CUDA-IPC handle of its exchange buffer (rg_shard_export, 64 bytes), the handles are all-gathered by the
 * caller (MPI / torch.distributed / any transport), and rg_shard_connect maps the peers. From then on the LAST CTA of each
 * linearize/compute_error launch stores this rank's 32 partial sums into every peer's buffer with P2P stores, waits for the
 * peers' rows and adds them in rank order inside the same kernel: no NCCL call, no extra launch, no host copy; all ranks
 * obtain bit-identical sums. Every rank must issue the same sequence of evaluations (they do: identical sums => identical
 * optimiser decisions). world <= 8. A peer that never arrives is reported as RG_ERR_CUDA after a bounded wait. */
int rg_shard_export(rg_ctx*, void* ipc_handle_64_bytes);
int rg_shard_connect(rg_ctx*, int rank, int world, const void* ipc_handles /* world x 64 bytes */);
/* Target-sharded SETUP of such a registration ("target is sharded, partial accumulators built locally and merged once per
 * align"). After rg_set_target_shard(rank, world) this rank estimates the covariances of ONE share of the target only (the target's
 * Morton order is dealt out in blocks of 2048 points, round-robin over the ranks — the expensive queries cluster in space, contiguous
 * slices would be badly balanced; the spatial index still covers the whole cloud, so the neighbourhoods are un-cut) and accumulates only those points into its voxel sums (vmp_voxel.hpp:167-197 up to, not including,
 * finalize). Once per voxel-map build the merge hook receives the rank's compact voxel list — device pointers keys uint64[n_local]
 * (packed voxel key) and sums double[n_local * 10] (sum of points xyz, sum of covariances xx xy xz yy yz zz, point count) — gathers the
 * lists of all ranks by whatever transport it likes (NCCL all-gather in rolo_b200/dist.py) and hands the concatenation to
 * rg_voxel_import, which adds entries with equal keys and finalizes (vmp_voxel.hpp:102-107). world <= 1 restores the full setup.
 * The target covariances must be computed by the context (rg_calculate_covariances / rg_align), not installed from outside. */
typedef int (*rg_voxel_merge_fn)(void* user, const void* keys_dev, const void* sums_dev, int n_local, void* stream);
int rg_set_target_shard(rg_ctx*, int rank, int world);
int rg_set_voxel_merge_hook(rg_ctx*, rg_voxel_merge_fn fn, void* user);
int rg_voxel_import(rg_ctx*, const void* keys_dev, const void* sums_dev, int n_total);

/* Deterministic voxel build: the target points are radix-sorted by voxel key and every voxel adds its points in index order (the order
 * of the reference's serial loop, vmp_voxel.hpp:167-197) instead of fp64 atomics in arrival order: means, covariances and voxel ids
 * are then bit-reproducible from run to run (and the sums equal the oracle's bit for bit). Slower; for tests and for ranks that must
 * agree to the last bit. Stand-alone contexts only (a batch always uses the atomic build). */
int rg_set_deterministic(rg_ctx*, int on);
/* voxelmap_.reset() (rot_vgicp_impl.hpp:147): the next linearize rebuilds the voxel map from the target and its covariances */
int rg_reset_voxelmap(rg_ctx*);

/* ---- the optimiser alone ----------------------------------------------------------------------------------
 * LsqRegistration's Gauss-Newton / LM / SO3-LM / translation-LM step logic (lsq_registration_impl.hpp:55-335: LDLT solves, damping,
 * rho test, convergence) as a resumable request / feed state machine. rg_align drives it internally; the C++ base class
 * fast_gicp::LsqRegistration (include/rot_gicp/gicp/lsq_registration.hpp) drives it through the reference's protected virtuals
 * (linearize / so3_linearize / compute_error / t3_linearize / compute_t_error), so user subclasses keep working:
 *   rg_lsq_begin_align(q, guess) or rg_lsq_begin_translation(q, trans);
 *   while ((kind = rg_lsq_want(q, pose, t3)) != RG_REQ_NONE) { evaluate the request; rg_lsq_feed(q, H, b, cost); }
 *   rg_lsq_state(q, ...).
 * H / b are row-major 6x6 / 6 (RG_REQ_LIN6, RG_REQ_T3_LIN) or 3x3 / 3 (RG_REQ_LIN3) and NULL for the two error requests. */
typedef struct rg_lsq rg_lsq;
enum { RG_REQ_NONE = 0, RG_REQ_LIN6, RG_REQ_LIN3, RG_REQ_ERROR, RG_REQ_T3_LIN, RG_REQ_T_ERROR };
rg_lsq* rg_lsq_create(void);
void rg_lsq_destroy(rg_lsq*);
void rg_lsq_configure(rg_lsq*, int optimizer, int max_iterations, double rotation_epsilon, double transformation_epsilon,
                      double lm_init_lambda_factor, int debug_print);
void rg_lsq_begin_align(rg_lsq*, const double guess[16]);
void rg_lsq_begin_translation(rg_lsq*, const double trans[3]);
int  rg_lsq_want(rg_lsq*, double pose16[16], double t3[3]);
void rg_lsq_feed(rg_lsq*, const double* H, const double* b, double cost);
void rg_lsq_state(const rg_lsq*, double pose16[16], double t3[3], int* nr_iterations, int* converged, int* failed, double final_hessian36[36]);

/* ---- upstream range-image passes ------------------------------------------------------------------------ */
/* ImageProjection::projectPointCloud + cloudExtraction (imageProjection.cpp:399-505); deskew off unless rg_set_deskew was called
 * (config/params.yaml:25 ships it disabled). Raw points: x,y,z float at byte offsets 0,4,8; ring uint16 at ring_offset.
 * Outputs (host): range image float[n_scan*horizon] (FLT_MAX = empty), extracted cloud float4 xyz+intensity
 * (intensity = ring*z, imageProjection.cpp:410), point_col_ind int32, point_range float, start/end ring index. */
typedef struct rg_projection_params {
  int n_scan, horizon_scan, downsample_rate;
  float lidar_min_range, lidar_max_range;
  int point_stride_bytes, ring_offset_bytes;
} rg_projection_params;
/* deskewPoint (imageProjection.cpp:368-396): with deskewEnabled && odomAvailable every projected point is rotated back by the
 * odometry increment (roll, pitch, yaw) scaled to its time within the sweep; translation is not compensated in the reference.
 * NULL (the default, config/params.yaml:25) disables it. Applies to rg_image_projection and rg_front_end_frame. */
typedef struct rg_deskew_params {
  int time_offset_bytes;                                     /* float relative time of the point (VelodynePointXYZIRT::time: 24) */
  float odom_incre_roll, odom_incre_pitch, odom_incre_yaw;   /* odomIncreRoll/Pitch/Yaw (:299,351) */
  float scan_period;                                         /* :79 */
  double odom_time_diff;                                     /* :298,350 */
} rg_deskew_params;
int rg_set_deskew(rg_ctx*, const rg_deskew_params* p);
int rg_image_projection(rg_ctx*, const void* host_raw_points, size_t n_points, const rg_projection_params* p,
                        float* range_image, float* extracted_xyzi, int32_t* point_col_ind, float* point_range,
                        int32_t* start_ring_index, int32_t* end_ring_index, int32_t* n_extracted);

/* FeatureExtraction::calculateSmoothness / markOccludedPoints / extractFeatures
 * (featureExtraction.cpp:87-266). Inputs are rg_image_projection's outputs.
 * Outputs: curvature float[n], neighbor_picked int32[n], label int32[n] after extraction,
 * corner cloud float4[<=n] and (un-downsampled) surface candidate mask.                          */
typedef struct rg_feature_params {
  int n_scan, horizon_scan;
  float edge_threshold, surf_threshold, surf_leaf_size;
} rg_feature_params;
int rg_feature_extraction(rg_ctx*, const float* extracted_xyzi, const int32_t* point_col_ind, const float* point_range,
                          const int32_t* start_ring_index, const int32_t* end_ring_index, int32_t n_extracted,
                          const rg_feature_params* p, float* curvature, int32_t* neighbor_picked, int32_t* label,
                          float* corner_xyzi, int32_t* n_corner, float* surface_xyzi, int32_t* n_surface);

/* One streamed front-end frame, device-resident end to end (BASELINE config 5): raw driver points (host) ->
 * projectPointCloud/cloudExtraction -> calculateSmoothness/markOccludedPoints/extractFeatures -> the feature cloud
 * `corner + surface` (lidarOdometry.cpp:521-524: featureLast = corner + surf) installed directly as this context's
 * target (which = RG_TARGET) or source cloud. Only the two counts cross back to the host. */
int rg_front_end_frame(rg_ctx*, const void* host_raw_points, size_t n_points, const rg_projection_params* pp,
                       const rg_feature_params* fp, int which, int32_t* n_corner, int32_t* n_surface);

/* ---- rolo/CloudInfoStamp at the ROS edge (msg/CloudInfoStamp.msg:1-28) ------------------------------
 * The stages hand their products over on the device (rg_front_end_frame); these two calls produce / consume the ROS1
 * wire image (roscpp serialisation, every cloud as pcl::toROSMsg writes a PointCloud<PointXYZI>: point_step 32) only
 * where a ROS publisher / subscriber needs it.
 *   stage 0 = the message imageProjection publishes (imageProjection.cpp:507-512): index arrays + cloud_projected;
 *   stage 1 = the message featureExtraction publishes (featureExtraction.cpp:268-287): index arrays cleared,
 *             cloud_projected kept, extracted_corner / extracted_surface filled, extracted_normal an empty cloud.
 * The point records are packed on the device and copied straight into `out` (pinned memory: no staging copy).
 * out == NULL: only *written (the size) is computed. */
typedef struct rg_msg_header { uint32_t seq, stamp_sec, stamp_nsec; const char* frame_id; } rg_msg_header;
typedef struct rg_cloud_info_scalars {
  float start_orientation, end_orientation, orientation_diff;
  float initial_guess[6];              /* X Y Z Roll Pitch Yaw */
  const float* covariance; uint32_t n_covariance;
  uint8_t odom_available;
} rg_cloud_info_scalars;
int rg_cloud_info_serialize(rg_ctx*, int stage, const rg_msg_header* hdr, const rg_cloud_info_scalars* scalars /* NULL: zeros */,
                            uint8_t* out, size_t capacity, size_t* written);
/* Consumer edge (lidarOdometry.cpp:521-524: fromROSMsg(extracted_corner), fromROSMsg(extracted_surface),
 * featureLast = corner + surf): parses a serialised CloudInfoStamp and installs corner + surface as this context's
 * source or target cloud; the point records are unpacked on the device. */
int rg_set_input_from_cloud_info(rg_ctx*, int which, const uint8_t* msg, size_t len, int32_t* n_corner, int32_t* n_surface);

/* ---- backMapping scan-to-map optimisation (src/backMapping.cpp:681-1054) -------------------------------
 * LOAM-style refinement of the odometry pose against the local feature map: per scan feature a 5-NN search in the
 * corner / surface map, a line (PCA) or plane (QR) fit, a point-to-line / point-to-plane coefficient, the 6x6 normal
 * equations, a Gauss-Newton step with the degeneracy projection of the first iteration. One fused kernel per iteration
 * (scan2map.cu); the 6x6 algebra stays on the host. Pose vectors are the reference's transformTobeMapped:
 * {roll, pitch, yaw, x, y, z} float. Clouds: xyz at the start of each record, stride_bytes apart.
 *   rg_s2m_set_map    = kdtreeCornerFromMap/kdtreeSurfFromMap->setInputCloud(laserCloud{Corner,Surf}FromMapDS)  (:689-690)
 *   rg_s2m_set_scan   = laserCloudCornerLastDS / laserCloudSurfLastDS                                            (:670-678)
 *   rg_s2m_linearize  = cornerOptimization + surfOptimization + combineOptimizationCoeffs + matAtA/matAtB at one pose
 *                       (:719-1006); n_selected = laserCloudSelNum
 *   rg_s2m_get_coefficients = coeffSel / laserCloudOri*Flag of the last evaluation, per scan point (4 floats, 1 byte)
 *   rg_s2m_optimize   = scan2MapOptimization's loop (:691-703) without transformUpdate(); enough_features = 0 reproduces
 *                       the "Not enough features!" branch (:687,705-707), tf is then left untouched. */
typedef struct rg_s2m_result { int32_t iterations, converged, degenerate, enough_features; } rg_s2m_result;
int rg_s2m_set_map(rg_ctx*, const float* corner, size_t n_corner, const float* surf, size_t n_surf, size_t stride_bytes);
int rg_s2m_set_scan(rg_ctx*, const float* corner, size_t n_corner, const float* surf, size_t n_surf, size_t stride_bytes);
int rg_s2m_linearize(rg_ctx*, const float tf[6], float AtA[36], float AtB[6], int32_t* n_selected);
int rg_s2m_get_coefficients(rg_ctx*, float* coeff_corner, uint8_t* flag_corner, float* coeff_surf, uint8_t* flag_surf);
int rg_s2m_optimize(rg_ctx*, float tf[6], int max_iterations, int edge_feature_min, int surf_feature_min, rg_s2m_result* out);

/* ---- Scan Context place recognition (src/scancontext/Scancontext.cpp) ----------------------------------
 * The context owns one SCManager. The 20 x 60 descriptor (max height per polar bin) and its ring / sector keys are built on
 * the device; alignment, cosine distance and the candidate search are a few thousand flops and run on the host.
 *   rg_sc_make_and_save = SCManager::makeAndSaveScancontextAndKeys(scan)      (Scancontext.cpp:151-250)
 *   rg_sc_detect        = SCManager::detectLoopClosureID()                    (:253-344): loop_id (-1 = none), yaw difference
 *   rg_sc_distance      = SCManager::distanceBtnScanContext(sc_i, sc_j)       (:116-148)
 *   rg_sc_get           = polarcontexts_[index] (row-major [ring][sector]), its ring key (20) and sector key (60); index -1 = last
 * lidar_height is Scancontext.h:80's LIDAR_HEIGHT (2.0 in the reference). */
int rg_sc_make_and_save(rg_ctx*, const float* pts, size_t n, size_t stride_bytes, double lidar_height);
int rg_sc_size(rg_ctx*);
int rg_sc_clear(rg_ctx*);
int rg_sc_get(rg_ctx*, int index, double* desc1200, double* ring_key20, double* sector_key60);
int rg_sc_distance(rg_ctx*, int i, int j, double* dist, int32_t* shift);
int rg_sc_detect(rg_ctx*, int32_t* loop_id, float* yaw_diff_rad);

/* ---- loop-closure verification (src/backMapping.cpp:2331-2352,2408-2431) -------------------------------
 * The reference aligns the current keyframe cloud to the candidate's with pcl::IterativeClosestPoint and accepts the loop when
 * hasConverged() && getFitnessScore() <= historyKeyframeFitnessScore. With this library the alignment is the GICP core in its
 * 6-DoF LM mode (rg_set_optimizer_type(RG_OPT_LM) + rg_align on the two keyframe clouds) and the acceptance test is
 *   rg_get_fitness_score = pcl::Registration::getFitnessScore(max_range): mean of the squared distances from every source
 *   point, moved by T, to its exact nearest target point, over the pairs with squared distance <= max_range (the reference
 *   compares the SQUARED distance with max_range; default DBL_MAX); DBL_MAX when no pair qualifies.
 *   rg_nearest_target = the per-point result behind it: target index int32[n_source], squared distance float[n_source]. */
int rg_get_fitness_score(rg_ctx*, const float T[16], double max_range, double* score);
int rg_nearest_target(rg_ctx*, const float T[16], int32_t* idx, float* sq_dist);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ROLO_GICP_H */
