/*
 * rolo_gicp_testing.h — test hooks exported by librolo_gicp_b200.so for the CPU-only test suite (no CUDA call inside).
 * Not part of the drop-in boundary: they expose host-side logic (the optimiser state machine of
 * rolo_b200/csrc/lsq_host.hpp and the persistent-CTA schedule of linearize.cu) so it can be checked without a GPU.
 */
#ifndef ROLO_GICP_TESTING_H
#define ROLO_GICP_TESTING_H
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden: only this ABI is exported */
#endif

typedef struct rgt_lsq rgt_lsq;   /* one host::LsqDriver */
enum { RGT_REQ_NONE = 0, RGT_REQ_LIN6, RGT_REQ_LIN3, RGT_REQ_ERROR, RGT_REQ_T3_LIN, RGT_REQ_T_ERROR };

rgt_lsq* rgt_lsq_create(int optimizer, int max_iterations, double rotation_epsilon, double transformation_epsilon);
void rgt_lsq_destroy(rgt_lsq*);
void rgt_lsq_begin_align(rgt_lsq*, const double guess[16]);
void rgt_lsq_begin_translation(rgt_lsq*, const double trans[3]);
/* returns the pending request kind (RGT_REQ_NONE when finished); pose16 (pose requests) or t3 (translation requests) filled */
int  rgt_lsq_want(rgt_lsq*, double pose16[16], double t3[3]);
void rgt_lsq_feed(rgt_lsq*, const double* H, const double* b, double cost);
void rgt_lsq_state(const rgt_lsq*, double pose16[16], double t3[3], int* nr_iterations, int* converged, int* failed, double final_hessian36[36]);

/* persistent-CTA schedule: sched int[3*n_pairs+1] = tile_prefix | cta_first | cta_count; returns total tiles */
int rgt_build_schedule(const int* n_points, int n_pairs, int grid, int* sched);

/* host math */
void rgt_se3_exp(const double a[6], double T[16]);
void rgt_ldlt_solve(const double* A, const double* rhs, int n, double* x);

/* scan-to-map host algebra (s2m_math.cuh): W descending + rows of V eigenvectors; least squares by Householder QR;
 * inverse by LU (returns 0 when singular); pcl::getTransformation rows */
void rgt_s2m_eigen(int n /* 3 or 6 */, const float* A, float* W, float* V);
void rgt_s2m_qr_solve(int rows, int cols /* 5x3 pivoted or 6x6 */, const float* A, const float* b, float* x);
int  rgt_s2m_invert6(const float* A, float* inv);
void rgt_s2m_get_transformation(const float tf[6], float t12[12]);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
