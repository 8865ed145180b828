/*
 * s2m_numerics.hpp — the THIRD-PARTY numerics behind backMapping's scan-to-map optimisation, restated from the published algorithms
 * (test infrastructure, NOT product code). Shared by scan2map_oracle.cpp (the restatement of the reference's own code) and by
 * ref_shim_bm/ (the stand-ins the reference's own member functions are compiled against in oracle/_ref/librolo_ref_bm.so), so a
 * comparison of those two isolates the transcription of the REFERENCE's code; these routines themselves are pinned against the real
 * libraries where the image has them: cv::eigen (3x3, 6x6) and cv::Mat::inv bit-equal to OpenCV 4.13 (tests/golden/opencv_s2m_golden.npz),
 * cv::solve(DECOMP_QR) to rounding (that build dispatches to LAPACK), Eigen::ColPivHouseholderQR against numpy lstsq.
 */
#pragma once
#include <cmath>
#include <utility>
#include <vector>

namespace s2m_num {

inline float hyp(float a, float b) {   // OpenCV's hypot helper
  a = std::fabs(a); b = std::fabs(b);
  if (a > b) { b /= a; return a * std::sqrt(1 + b * b); }
  if (b > 0) { a /= b; return b * std::sqrt(1 + a * a); }
  return 0;
}

// cv::eigen for a symmetric CV_32F matrix (JacobiImpl_): classical Jacobi choosing the largest off-diagonal pivot, rows of V
// are the eigenvectors, eigenvalues sorted descending. A (n x n, row-major) is destroyed.
inline void jacobi_eigen(int n, float* A, float* W, float* V) {
  const float eps = 1.1920929e-07f;
  std::vector<int> indR(n), indC(n);
  for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) V[i * n + j] = 0.f; V[i * n + i] = 1.f; }
  for (int k = 0; k < n; k++) {
    W[k] = A[k * n + k];
    if (k < n - 1) { int m = k + 1; float mv = std::fabs(A[k * n + m]); for (int i = k + 2; i < n; i++) { float v = std::fabs(A[k * n + i]); if (mv < v) { mv = v; m = i; } } indR[k] = m; }
    if (k > 0) { int m = 0; float mv = std::fabs(A[k]); for (int i = 1; i < k; i++) { float v = std::fabs(A[i * n + k]); if (mv < v) { mv = v; m = i; } } indC[k] = m; }
  }
  if (n > 1) for (int iters = 0, maxIters = n * n * 30; iters < maxIters; iters++) {
    int k = 0, l; float mv = std::fabs(A[indR[0]]);
    for (int i = 1; i < n - 1; i++) { float v = std::fabs(A[i * n + indR[i]]); if (mv < v) { mv = v; k = i; } }
    l = indR[k];
    for (int i = 1; i < n; i++) { float v = std::fabs(A[indC[i] * n + i]); if (mv < v) { mv = v; k = indC[i]; l = i; } }
    float p = A[k * n + l];
    if (std::fabs(p) <= eps) break;
    float y = (W[l] - W[k]) * 0.5f;
    float t = std::fabs(y) + hyp(p, y);
    float s = hyp(p, t);
    float c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) { s = -s; t = -t; }
    A[k * n + l] = 0;
    W[k] -= t; W[l] += t;
    float a0, b0;
#define ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
    for (int i = 0; i < k; i++) ROT(A[i * n + k], A[i * n + l]);
    for (int i = k + 1; i < l; i++) ROT(A[k * n + i], A[i * n + l]);
    for (int i = l + 1; i < n; i++) ROT(A[k * n + i], A[l * n + i]);
    for (int i = 0; i < n; i++) ROT(V[k * n + i], V[l * n + i]);
#undef ROT
    for (int j = 0; j < 2; j++) {
      int idx = j == 0 ? k : l;
      if (idx < n - 1) { int m = idx + 1; float mv2 = std::fabs(A[idx * n + m]); for (int i = idx + 2; i < n; i++) { float v = std::fabs(A[idx * n + i]); if (mv2 < v) { mv2 = v; m = i; } } indR[idx] = m; }
      if (idx > 0) { int m = 0; float mv2 = std::fabs(A[idx]); for (int i = 1; i < idx; i++) { float v = std::fabs(A[i * n + idx]); if (mv2 < v) { mv2 = v; m = i; } } indC[idx] = m; }
    }
  }
  for (int k = 0; k < n - 1; k++) {       // descending
    int m = k;
    for (int i = k + 1; i < n; i++) if (W[m] < W[i]) m = i;
    if (k != m) { std::swap(W[m], W[k]); for (int i = 0; i < n; i++) std::swap(V[m * n + i], V[k * n + i]); }
  }
}

// Householder reflector of x[0..m) (Eigen makeHouseholderInPlace): x[0] <- beta, x[1..] <- essential part, returns tau
inline float householder(float* x, int m, int stride) {
  float tail = 0.f;
  for (int i = 1; i < m; i++) tail += x[i * stride] * x[i * stride];
  const float c0 = x[0];
  if (tail <= 1.17549435e-38f) { for (int i = 1; i < m; i++) x[i * stride] = 0.f; return 0.f; }
  float beta = std::sqrt(c0 * c0 + tail);
  if (c0 >= 0.f) beta = -beta;
  for (int i = 1; i < m; i++) x[i * stride] = x[i * stride] / (c0 - beta);
  x[0] = beta;
  return (beta - c0) / beta;
}

// least squares min |A x - b| by Householder QR, optional column pivoting (rows x cols, row-major, A and b destroyed).
// pivot = true: Eigen::ColPivHouseholderQR::solve (the 5x3 plane fit); false: cv::solve(DECOMP_QR) (the 6x6 normal equations)
inline void qr_solve(int rows, int cols, float* A, float* b, float* x, bool pivot) {
  std::vector<int> perm(cols);
  std::vector<float> cn(cols);
  for (int j = 0; j < cols; j++) { perm[j] = j; float s = 0.f; for (int i = 0; i < rows; i++) s += A[i * cols + j] * A[i * cols + j]; cn[j] = s; }
  for (int k = 0; k < cols; k++) {
    if (pivot) {
      int m = k; float best = 0.f;
      for (int j = k; j < cols; j++) { float s = 0.f; for (int i = k; i < rows; i++) s += A[i * cols + j] * A[i * cols + j]; cn[j] = s; if (j == k || s > best) { best = s; m = j; } }
      if (m != k) { for (int i = 0; i < rows; i++) std::swap(A[i * cols + k], A[i * cols + m]); std::swap(perm[k], perm[m]); }
    }
    const float tau = householder(&A[k * cols + k], rows - k, cols);
    if (tau != 0.f) {
      for (int j = k + 1; j < cols; j++) {     // apply H = I - tau v v^T (v = [1, essential]) to the remaining columns
        float dot = A[k * cols + j];
        for (int i = k + 1; i < rows; i++) dot += A[i * cols + k] * A[i * cols + j];
        dot *= tau;
        A[k * cols + j] -= dot;
        for (int i = k + 1; i < rows; i++) A[i * cols + j] -= dot * A[i * cols + k];
      }
      float dot = b[k];
      for (int i = k + 1; i < rows; i++) dot += A[i * cols + k] * b[i];
      dot *= tau;
      b[k] -= dot;
      for (int i = k + 1; i < rows; i++) b[i] -= dot * A[i * cols + k];
    }
  }
  std::vector<float> y(cols);
  for (int k = cols - 1; k >= 0; k--) {      // R y = (Q^T b)[0..cols)
    float s = b[k];
    for (int j = k + 1; j < cols; j++) s -= A[k * cols + j] * y[j];
    const float d = A[k * cols + k];
    y[k] = d != 0.f ? s / d : 0.f;
  }
  for (int k = 0; k < cols; k++) x[perm[k]] = y[k];
}

// cv::Mat::inv() (DECOMP_LU): Gauss-Jordan elimination with partial pivoting on [A | I]; returns false when singular
inline bool invert_lu(int n, const float* Ain, float* inv) {
  std::vector<float> A(Ain, Ain + n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = i == j ? 1.f : 0.f;
  for (int i = 0; i < n; i++) {
    int k = i;
    for (int j = i + 1; j < n; j++) if (std::fabs(A[j * n + i]) > std::fabs(A[k * n + i])) k = j;
    if (std::fabs(A[k * n + i]) < 1.1920929e-07f * 100) return false;
    if (k != i) for (int j = 0; j < n; j++) { std::swap(A[i * n + j], A[k * n + j]); std::swap(inv[i * n + j], inv[k * n + j]); }
    const float d = -1.f / A[i * n + i];
    for (int j = i + 1; j < n; j++) {
      const float alpha = A[j * n + i] * d;
      for (int c = i + 1; c < n; c++) A[j * n + c] += alpha * A[i * n + c];
      for (int c = 0; c < n; c++) inv[j * n + c] += alpha * inv[i * n + c];
    }
  }
  for (int i = n - 1; i >= 0; i--)
    for (int c = 0; c < n; c++) {
      float s = inv[i * n + c];
      for (int k = i + 1; k < n; k++) s -= A[i * n + k] * inv[k * n + c];
      inv[i * n + c] = s / A[i * n + i];
    }
  return true;
}

// cv::gemm for CV_32F operands: C (m x n) = A (m x k) * B (k x n), products accumulated in double, rounded to float once
inline void gemm_f32(int m, int n, int k, const float* A, const float* B, float* C) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double acc = 0.0;
      for (int t = 0; t < k; t++) acc += (double)A[i * k + t] * (double)B[t * n + j];
      C[i * n + j] = (float)acc;
    }
}

}  // namespace s2m_num
