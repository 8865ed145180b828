// s2m_math.cuh — small dense float routines of the scan-to-map optimisation (backMapping.cpp:681-1054), usable on the host
// and inside the kernel so both sides of the C-ABI run the same operation sequence.
//
// The reference calls OpenCV / Eigen / PCL for these (none in /root/reference); what is implemented is their published
// algorithm: cv::eigen's classical max-pivot Jacobi with a descending sort (3x3 per corner point on the device, 6x6 once per
// optimisation on the host), Householder QR least squares (column-pivoted 5x3 plane fit on the device as
// Eigen::colPivHouseholderQr().solve, unpivoted 6x6 as cv::solve(DECOMP_QR) on the host), cv::Mat::inv by LU with partial
// pivoting, pcl::getTransformation's closed form. All float; the translation unit is compiled with -fmad=false so the
// device results equal a plain IEEE evaluation.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace rolo {
namespace s2m {

__host__ __device__ inline float hyp(float a, float b) {
  a = fabsf(a); b = fabsf(b);
  if (a > b) { b /= a; return a * sqrtf(1 + b * b); }
  if (b > 0) { a /= b; return b * sqrtf(1 + a * a); }
  return 0;
}

template <int N>
__host__ __device__ inline int row_argmax(const float* A, int k) {     // largest |A[k][i]|, i > k
  int m = k + 1; float mv = fabsf(A[k * N + m]);
  for (int i = k + 2; i < N; i++) { const float v = fabsf(A[k * N + i]); if (mv < v) { mv = v; m = i; } }
  return m;
}
template <int N>
__host__ __device__ inline int col_argmax(const float* A, int k) {     // largest |A[i][k]|, i < k
  int m = 0; float mv = fabsf(A[k]);
  for (int i = 1; i < k; i++) { const float v = fabsf(A[i * N + k]); if (mv < v) { mv = v; m = i; } }
  return m;
}

// symmetric eigen-decomposition: W descending, rows of V the eigenvectors; A (row-major) is overwritten
template <int N>
__host__ __device__ inline void jacobi_eigen(float* A, float* W, float* V) {
  int indR[N], indC[N];
  for (int i = 0; i < N; i++) { for (int j = 0; j < N; j++) V[i * N + j] = 0.f; V[i * N + i] = 1.f; }
  for (int k = 0; k < N; k++) {
    W[k] = A[k * N + k];
    if (k < N - 1) indR[k] = row_argmax<N>(A, k);
    if (k > 0) indC[k] = col_argmax<N>(A, k);
  }
  for (int iters = 0; iters < N * N * 30; iters++) {
    int k = 0; float mv = fabsf(A[indR[0]]);
    for (int i = 1; i < N - 1; i++) { const float v = fabsf(A[i * N + indR[i]]); if (mv < v) { mv = v; k = i; } }
    int l = indR[k];
    for (int i = 1; i < N; i++) { const float v = fabsf(A[indC[i] * N + i]); if (mv < v) { mv = v; k = indC[i]; l = i; } }
    const float p = A[k * N + l];
    if (fabsf(p) <= 1.1920929e-07f) break;
    const float y = (W[l] - W[k]) * 0.5f;
    float t = fabsf(y) + hyp(p, y);
    float s = hyp(p, t);
    const float c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) { s = -s; t = -t; }
    A[k * N + l] = 0;
    W[k] -= t; W[l] += t;
    for (int i = 0; i < k; i++) { const float a0 = A[i * N + k], b0 = A[i * N + l]; A[i * N + k] = a0 * c - b0 * s; A[i * N + l] = a0 * s + b0 * c; }
    for (int i = k + 1; i < l; i++) { const float a0 = A[k * N + i], b0 = A[i * N + l]; A[k * N + i] = a0 * c - b0 * s; A[i * N + l] = a0 * s + b0 * c; }
    for (int i = l + 1; i < N; i++) { const float a0 = A[k * N + i], b0 = A[l * N + i]; A[k * N + i] = a0 * c - b0 * s; A[l * N + i] = a0 * s + b0 * c; }
    for (int i = 0; i < N; i++) { const float a0 = V[k * N + i], b0 = V[l * N + i]; V[k * N + i] = a0 * c - b0 * s; V[l * N + i] = a0 * s + b0 * c; }
    for (int j = 0; j < 2; j++) {
      const int idx = j == 0 ? k : l;
      if (idx < N - 1) indR[idx] = row_argmax<N>(A, idx);
      if (idx > 0) indC[idx] = col_argmax<N>(A, idx);
    }
  }
  for (int k = 0; k < N - 1; k++) {
    int m = k;
    for (int i = k + 1; i < N; i++) if (W[m] < W[i]) m = i;
    if (k != m) {
      float tmp = W[m]; W[m] = W[k]; W[k] = tmp;
      for (int i = 0; i < N; i++) { tmp = V[m * N + i]; V[m * N + i] = V[k * N + i]; V[k * N + i] = tmp; }
    }
  }
}

// least squares min |A x - b| through Householder reflections; A is ROWS x COLS row-major, A and b are overwritten
template <int ROWS, int COLS, bool PIVOT>
__host__ __device__ inline void qr_solve(float* A, float* b, float* x) {
  int perm[COLS];
  for (int j = 0; j < COLS; j++) perm[j] = j;
  for (int k = 0; k < COLS; k++) {
    if (PIVOT) {       // bring the remaining column of largest norm forward
      int m = k; float best = 0.f;
      for (int j = k; j < COLS; j++) {
        float s = 0.f;
        for (int i = k; i < ROWS; i++) s += A[i * COLS + j] * A[i * COLS + j];
        if (j == k || s > best) { best = s; m = j; }
      }
      if (m != k) {
        for (int i = 0; i < ROWS; i++) { const float tmp = A[i * COLS + k]; A[i * COLS + k] = A[i * COLS + m]; A[i * COLS + m] = tmp; }
        const int tp = perm[k]; perm[k] = perm[m]; perm[m] = tp;
      }
    }
    // reflector of column k below the diagonal: A[k][k] <- beta, A[k+1..][k] <- essential part
    float tail = 0.f;
    for (int i = k + 1; i < ROWS; i++) tail += A[i * COLS + k] * A[i * COLS + k];
    const float c0 = A[k * COLS + k];
    float tau = 0.f;
    if (tail <= 1.17549435e-38f) {
      for (int i = k + 1; i < ROWS; i++) A[i * COLS + k] = 0.f;
    } else {
      float beta = sqrtf(c0 * c0 + tail);
      if (c0 >= 0.f) beta = -beta;
      for (int i = k + 1; i < ROWS; i++) A[i * COLS + k] = A[i * COLS + k] / (c0 - beta);
      A[k * COLS + k] = beta;
      tau = (beta - c0) / beta;
    }
    if (tau != 0.f) {
      for (int j = k + 1; j < COLS; j++) {
        float dot = A[k * COLS + j];
        for (int i = k + 1; i < ROWS; i++) dot += A[i * COLS + k] * A[i * COLS + j];
        dot *= tau;
        A[k * COLS + j] -= dot;
        for (int i = k + 1; i < ROWS; i++) A[i * COLS + j] -= dot * A[i * COLS + k];
      }
      float dot = b[k];
      for (int i = k + 1; i < ROWS; i++) dot += A[i * COLS + k] * b[i];
      dot *= tau;
      b[k] -= dot;
      for (int i = k + 1; i < ROWS; i++) b[i] -= dot * A[i * COLS + k];
    }
  }
  float y[COLS];
  for (int k = COLS - 1; k >= 0; k--) {
    float s = b[k];
    for (int j = k + 1; j < COLS; j++) s -= A[k * COLS + j] * y[j];
    const float d = A[k * COLS + k];
    y[k] = d != 0.f ? s / d : 0.f;
  }
  for (int k = 0; k < COLS; k++) x[perm[k]] = y[k];
}

// inverse by elimination with partial pivoting on [A | I]; false (inv untouched beyond scratch) when a pivot vanishes
template <int N>
inline bool invert_lu(const float* Ain, float* inv) {
  float A[N * N];
  for (int i = 0; i < N * N; i++) A[i] = Ain[i];
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) inv[i * N + j] = i == j ? 1.f : 0.f;
  for (int i = 0; i < N; i++) {
    int k = i;
    for (int j = i + 1; j < N; j++) if (fabsf(A[j * N + i]) > fabsf(A[k * N + i])) k = j;
    if (fabsf(A[k * N + i]) < 1.1920929e-07f * 100) return false;
    if (k != i)
      for (int j = 0; j < N; j++) {
        float tmp = A[i * N + j]; A[i * N + j] = A[k * N + j]; A[k * N + j] = tmp;
        tmp = inv[i * N + j]; inv[i * N + j] = inv[k * N + j]; inv[k * N + j] = tmp;
      }
    const float d = -1.f / A[i * N + i];
    for (int j = i + 1; j < N; j++) {
      const float alpha = A[j * N + i] * d;
      for (int c = i + 1; c < N; c++) A[j * N + c] += alpha * A[i * N + c];
      for (int c = 0; c < N; c++) inv[j * N + c] += alpha * inv[i * N + c];
    }
  }
  for (int i = N - 1; i >= 0; i--)
    for (int c = 0; c < N; c++) {
      float s = inv[i * N + c];
      for (int k = i + 1; k < N; k++) s -= A[i * N + k] * inv[k * N + c];
      inv[i * N + c] = s / A[i * N + i];
    }
  return true;
}

// rows of the 3x4 matrix pcl::getTransformation(x, y, z, roll, pitch, yaw) builds; tf = roll pitch yaw x y z
// (trans2Affine3f, backMapping.cpp:339-342)
inline void get_transformation(const float tf[6], float t[12]) {
  const float A = cosf(tf[2]), B = sinf(tf[2]), C = cosf(tf[1]), D = sinf(tf[1]), E = cosf(tf[0]), F = sinf(tf[0]);
  const float DE = D * E, DF = D * F;
  t[0] = A * C; t[1] = A * DF - B * E; t[2] = B * F + A * DE; t[3] = tf[3];
  t[4] = B * C; t[5] = A * E + B * DF; t[6] = B * DE - A * F; t[7] = tf[4];
  t[8] = -D;    t[9] = C * F;          t[10] = C * E;         t[11] = tf[5];
}

}  // namespace s2m
}  // namespace rolo
