// ref_shim_bm/eigen_colpivqr.hpp — Eigen::ColPivHouseholderQR::solve for the stand-in Eigen (oracle/ref_shim/Eigen/Core), over the
// restated Householder QR of oracle/s2m_numerics.hpp. Test infrastructure only.
#pragma once
#include <Eigen/Core>
#include "s2m_numerics.hpp"
namespace Eigen { namespace internal {
template <typename S, int R, int C> Matrix<S, C, 1> ColPivQr<S, R, C>::solve(const Matrix<S, R, 1>& b) const {
  static_assert(std::is_same<S, float>::value, "float only");
  float a[R * C], bb[R], x[C];
  for (int i = 0; i < R; i++) { for (int j = 0; j < C; j++) a[i * C + j] = A(i, j); bb[i] = b(i); }
  s2m_num::qr_solve(R, C, a, bb, x, true);
  Matrix<S, C, 1> out;
  for (int j = 0; j < C; j++) out(j) = x[j];
  return out;
}
}}  // namespace Eigen::internal
