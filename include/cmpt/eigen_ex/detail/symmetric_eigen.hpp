// detail/symmetric_eigen.hpp — host solver for small dense real symmetric matrices (cyclic Jacobi).
//
// Used by the thick-restart driver (thick_restart.hpp): after a restart the projected matrix is an arrowhead block
// followed by a tridiagonal tail, not tridiagonal, so detail/tridiag_eigen.hpp does not apply.  Sizes are the Krylov
// basis size (tens to a few hundred); Jacobi is simple, unconditionally convergent and accurate to working
// precision for symmetric matrices (Golub & Van Loan, section 8.5).
#ifndef CMPT_EIGEN_EX_DETAIL_SYMMETRIC_EIGEN_HPP_
#define CMPT_EIGEN_EX_DETAIL_SYMMETRIC_EIGEN_HPP_

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

namespace cmpt {
namespace EigenEx {
namespace detail {

/// a: n x n column-major symmetric matrix (destroyed).  w: eigenvalues ascending.  z: n x n column-major, columns =
/// eigenvectors in the order of w.  Returns false when 60 sweeps did not reach the tolerance.
template <class Real>
bool symmetric_eigensystem(int n, std::vector<Real>& a, std::vector<Real>& w, std::vector<Real>& z) {
  w.assign(static_cast<std::size_t>(n), Real(0));
  z.assign(static_cast<std::size_t>(n) * n, Real(0));
  for (int i = 0; i < n; ++i) z[static_cast<std::size_t>(i) * n + i] = Real(1);
  auto A = [&](int i, int j) -> Real& { return a[static_cast<std::size_t>(j) * n + i]; };
  auto Z = [&](int i, int j) -> Real& { return z[static_cast<std::size_t>(j) * n + i]; };
  bool converged = (n <= 1);
  for (int sweep = 0; sweep < 60 && !converged; ++sweep) {
    Real off = 0, diag = 0;
    for (int j = 0; j < n; ++j) {
      diag += A(j, j) * A(j, j);
      for (int i = 0; i < j; ++i) off += A(i, j) * A(i, j);
    }
    if (off <= std::numeric_limits<Real>::epsilon() * std::numeric_limits<Real>::epsilon() * (diag + off)) {
      converged = true;
      break;
    }
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const Real apq = A(p, q);
        if (apq == Real(0)) continue;
        const Real theta = (A(q, q) - A(p, p)) / (Real(2) * apq);
        const Real t = (theta >= 0 ? Real(1) : Real(-1)) / (std::abs(theta) + std::sqrt(theta * theta + Real(1)));
        const Real c = Real(1) / std::sqrt(t * t + Real(1)), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns p, q
          const Real akp = A(k, p), akq = A(k, q);
          A(k, p) = c * akp - s * akq;
          A(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p, q
          const Real apk = A(p, k), aqk = A(q, k);
          A(p, k) = c * apk - s * aqk;
          A(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const Real zkp = Z(k, p), zkq = Z(k, q);
          Z(k, p) = c * zkp - s * zkq;
          Z(k, q) = s * zkp + c * zkq;
        }
      }
    }
  }
  std::vector<int> order(static_cast<std::size_t>(n));
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return A(x, x) < A(y, y); });
  std::vector<Real> zs(z.size());
  for (int j = 0; j < n; ++j) {
    w[static_cast<std::size_t>(j)] = A(order[j], order[j]);
    std::copy(z.begin() + static_cast<std::size_t>(order[j]) * n, z.begin() + static_cast<std::size_t>(order[j] + 1) * n,
              zs.begin() + static_cast<std::size_t>(j) * n);
  }
  z.swap(zs);
  return converged;
}

}  // namespace detail
}  // namespace EigenEx
}  // namespace cmpt

#endif
