// ref_shim_bm/cv_ops.hpp — the cv:: operations backMapping's scan-to-map member functions call on CV_32F matrices, for the stand-in
// cv::Mat of oracle/ref_shim_ros/rolo_ros_shim.h: cv::eigen (symmetric, Jacobi), cv::transpose, operator* (cv::gemm), cv::solve(DECOMP_QR),
// cv::Mat::inv (DECOMP_LU). The arithmetic is oracle/s2m_numerics.hpp (restated from OpenCV's published algorithms and pinned against the
// real library where the image has it — see that header). Test infrastructure only.
#pragma once
#include "rolo_ros_shim.h"
#include "s2m_numerics.hpp"
namespace cv {
enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4 };
inline Mat make_mat(int r, int c) { return Mat(r, c, CV_32F, Scalar::all(0)); }
// eigenvalues descending in an n x 1 column, eigenvectors in the ROWS of an n x n matrix (OpenCV's convention)
inline bool eigen(const Mat& src, Mat& eigenvalues, Mat& eigenvectors) {
  const int n = src.rows;
  std::vector<float> a(src.d);
  eigenvalues = make_mat(n, 1); eigenvectors = make_mat(n, n);
  s2m_num::jacobi_eigen(n, a.data(), eigenvalues.d.data(), eigenvectors.d.data());
  return true;
}
inline void transpose(const Mat& src, Mat& dst) {
  Mat t = make_mat(src.cols, src.rows);
  for (int i = 0; i < src.rows; i++) for (int j = 0; j < src.cols; j++) t.d[(size_t)j * src.rows + i] = src.d[(size_t)i * src.cols + j];
  dst = t;
}
inline Mat operator*(const Mat& a, const Mat& b) {
  Mat c = make_mat(a.rows, b.cols);
  s2m_num::gemm_f32(a.rows, b.cols, a.cols, a.d.data(), b.d.data(), c.d.data());
  return c;
}
inline bool solve(const Mat& A, const Mat& B, Mat& X, int flags) {
  if (flags != DECOMP_QR || B.cols != 1) throw std::runtime_error("cv::solve stand-in: DECOMP_QR with one right-hand side only");
  std::vector<float> a(A.d), b(B.d);
  X = make_mat(A.cols, 1);
  s2m_num::qr_solve(A.rows, A.cols, a.data(), b.data(), X.d.data(), false);
  return true;
}
inline Mat Mat::inv() const {
  Mat r = make_mat(rows, cols);
  if (!s2m_num::invert_lu(rows, d.data(), r.d.data())) std::fill(r.d.begin(), r.d.end(), 0.f);   // singular: cv::Mat::inv returns zeros
  return r;
}
}  // namespace cv
