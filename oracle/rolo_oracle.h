/*
 * rolo_oracle.h — C API of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a structure-faithful CPU/OpenMP restatement of the reference's
 * rot_gicp hot path (sdwyc/ROLO @ 3920bb7):
 *   include/rot_gicp/gicp/impl/rot_vgicp_impl.hpp
 *   include/rot_gicp/gicp/impl/lsq_registration_impl.hpp
 *   include/rot_gicp/gicp/vmp_voxel.hpp
 *   include/rot_gicp/so3/so3.hpp
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under rolo_b200/ links or imports it.
 *
 * PINNING: the reference ships no golden vectors or numeric tests for this path
 * (SURVEY.md §4, §8c). The restatement is pinned against the reference ITSELF run
 * here: oracle/_ref/librolo_ref.so is the reference's own rot_vgicp_impl.hpp /
 * lsq_registration_impl.hpp / vmp_voxel.hpp / so3.hpp compiled from /root/reference
 * (oracle/ref_driver.cpp, `make -C oracle ref`) and tests/test_oracle_vs_ref_cpu.py
 * requires covariances, per-call {H,b,cost}, align() and computeTranslation() to agree
 * to 1e-11. What stays restated (not pinned by reference code, absent from the tree):
 * Eigen's JacobiSVD/LDLT/inverse and PCL/FLANN's kNN, provided by oracle/ref_shim for
 * both sides.
 *
 * Conventions: all matrices ROW-MAJOR. Poses are 4x4 (double[16]).
 * H6 is 6x6 row-major (36 doubles), b6 6 doubles; H3 3x3 (9), b3 (3).
 */
#ifndef ROLO_ORACLE_H
#define ROLO_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ro_ctx ro_ctx;

/* enum values follow include/rot_gicp/gicp/gicp_settings.hpp:6-12 and lsq_registration.hpp:13 */
enum { RO_REG_NONE = 0, RO_REG_MIN_EIG, RO_REG_NORMALIZED_MIN_EIG, RO_REG_PLANE, RO_REG_FROBENIUS, RO_REG_PLANE_S };
enum { RO_DIRECT27 = 0, RO_DIRECT7, RO_DIRECT1, RO_DIRECT_RADIUS };
enum { RO_VOXEL_POLAR = 0, RO_VOXEL_UNIFORM };
enum { RO_OPT_GAUSS_NEWTON = 0, RO_OPT_LM, RO_OPT_SO3_LM };

ro_ctx* ro_create(void);
void ro_destroy(ro_ctx*);

void ro_set_num_threads(ro_ctx*, int n);                 /* 0 => omp_get_max_threads() */
int  ro_get_num_threads(ro_ctx*);
void ro_set_resolution(ro_ctx*, double res);             /* => UNIFORM */
void ro_set_polar_resolution(ro_ctx*, double th, double ph, double r); /* => POLAR */
void ro_set_neighbor_search_method(ro_ctx*, int m);
void ro_set_regularization_method(ro_ctx*, int m);
void ro_set_correspondence_randomness(ro_ctx*, int k);
void ro_set_optimizer_type(ro_ctx*, int t);
void ro_set_max_iterations(ro_ctx*, int n);
void ro_set_rotation_epsilon(ro_ctx*, double e);
void ro_set_transformation_epsilon(ro_ctx*, double e);
void ro_set_initial_lambda_factor(ro_ctx*, double f);

/* clouds: n points, xyz at pts[i*stride_floats + 0..2]; w forced to 1 */
void ro_set_input_source(ro_ctx*, const float* pts, size_t n, size_t stride_floats);
void ro_set_input_target(ro_ctx*, const float* pts, size_t n, size_t stride_floats);
void ro_set_covariances(ro_ctx*, int which /*0 src,1 tgt*/, const double* cov4x4, size_t n);
size_t ro_get_covariances(ro_ctx*, int which, double* cov4x4 /* n*16 or NULL */);
int  ro_calculate_covariances(ro_ctx*, int which);
/* kNN only (k nearest incl. self, float distances, ties by lower index); idx int32[n*k] */
void ro_knn(ro_ctx*, int which, int k, int* idx);

void   ro_create_voxelmap(ro_ctx*);
size_t ro_voxel_count(ro_ctx*);
/* voxels sorted by (key0,key1,key2): keys int[V*3], mean double[V*4], cov double[V*16], npts int[V] */
void   ro_get_voxels(ro_ctx*, int* keys, double* mean, double* cov, int* npts);
void   ro_reset_voxelmap(ro_ctx*);

double ro_linearize(ro_ctx*, const double T[16], double* H36, double* b6);
double ro_so3_linearize(ro_ctx*, const double T[16], double* H9, double* b3);
double ro_compute_error(ro_ctx*, const double T[16]);
double ro_t3_linearize(ro_ctx*, const double t[3], const double guess[3], const double last_t0[3],
                       double dtn, double dtn1, double* H36, double* b6);
double ro_compute_t_error(ro_ctx*, const double t[3], const double guess[3], const double last_t0[3],
                          double dtn, double dtn1);
void   ro_set_ct_lambda(ro_ctx*, float lambda);
size_t ro_num_correspondences(ro_ctx*);
/* correspondence list in reference order: src index int[M], voxel key int[M*3] */
void   ro_get_correspondences(ro_ctx*, int* src_idx, int* vox_key);

/* align(): pcl::Registration::align + RotVGICP::computeTransformation. guess/out 4x4 float row-major.
 * returns 0 ok, 1 "lm not converged" break happened. */
int  ro_align(ro_ctx*, const float guess[16], float out_T[16]);
int  ro_compute_translation(ro_ctx*, double trans[3] /*inout*/, const double guess[3], const double last_t0[3],
                            double dtn, double dtn1, float ct_lambda);
/* pcl::Registration::getFitnessScore(max_range) at transform T (row-major 4x4 float); idx int[n_source] / d2 float[n_source] optional */
double ro_get_fitness_score(ro_ctx*, const float T[16], double max_range, int* idx, float* d2);
int  ro_has_converged(ro_ctx*);
int  ro_nr_iterations(ro_ctx*);
void ro_get_final_hessian(ro_ctx*, double* H36);
void ro_get_final_pose_d(ro_ctx*, double T[16]);   /* double-precision x0 before the float cast */

/* trace of every linearize/compute_error call during the last align/compute_translation.
 * record = 64 doubles: [0]=kind (0 linearize6,1 so3_linearize,2 compute_error,3 t3_linearize,4 compute_t_error),
 * [1]=cost, [2..17]=pose (row-major 4x4; for t3 kinds only translation column meaningful),
 * [18..53]=H (6x6 or 3x3 in the top-left), [54..59]=b, [60]=lambda at call time, [61..63] reserved */
size_t ro_trace_size(ro_ctx*);
void   ro_get_trace(ro_ctx*, double* out);

/* host helpers exposed for unit tests */
void ro_so3_exp(const double w[3], double q_wxyz[4]);
void ro_se3_exp(const double a[6], double T[16]);
void ro_svd3(const double A[9], double U[9], double S[3], double V[9]);
int  ro_ldlt_solve(const double* A, const double* rhs, int n, double* x);

#ifdef __cplusplus
}
#endif
#endif
