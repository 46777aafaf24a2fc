// detail/tridiag_eigen.hpp — host solver for the m x m tridiagonal Ritz problem.
//
// The reference calls Eigen::SelfAdjointEigenSolver<...>::computeFromTridiagonal(alpha, beta)
// (lanczos.hpp:637,741,779-781) — Eigen is a third-party dependency that is not part of the reference
// tree.  This is an independent implementation of the same published algorithm (implicit symmetric QR
// with Wilkinson shift, Golub & Van Loan alg. 8.3.3): eigenvalues ascending, eigenvectors as columns,
// only the first n-1 entries of the sub-diagonal are read.  It is part of the algorithm (the Ritz
// problem is solved on the host by design), not a fallback for the device path.
#ifndef CMPT_EIGEN_EX_DETAIL_TRIDIAG_EIGEN_HPP_
#define CMPT_EIGEN_EX_DETAIL_TRIDIAG_EIGEN_HPP_

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

namespace cmpt {
namespace EigenEx {
namespace detail {

// d[0..n): diagonal (overwritten by the eigenvalues, unsorted); e[0..n-1): sub-diagonal (destroyed).
// z: optional n x n column-major accumulator (must be the identity on entry).  Returns false if the
// iteration limit (30 n sweeps) is hit.
template <class Real>
bool tridiagonal_qr(int n, Real* d, Real* e, Real* z) {
  if (n <= 1) return true;
  const Real eps = std::numeric_limits<Real>::epsilon();
  const Real tiny = std::numeric_limits<Real>::min();
  int end = n - 1;
  int iter = 0;
  const int max_iter = 30 * n;
  while (end > 0) {
    for (int i = 0; i < end; ++i)
      if (std::abs(e[i]) <= eps * (std::abs(d[i]) + std::abs(d[i + 1])) || std::abs(e[i]) <= tiny) e[i] = Real(0);
    while (end > 0 && e[end - 1] == Real(0)) --end;
    if (end <= 0) break;
    if (++iter > max_iter) return false;
    int start = end - 1;
    while (start > 0 && e[start - 1] != Real(0)) --start;
    // Wilkinson shift from the trailing 2x2 of the unreduced block
    const Real td = (d[end - 1] - d[end]) * Real(0.5);
    const Real ee = e[end - 1];
    Real mu = d[end];
    if (td == Real(0)) {
      mu -= std::abs(ee);
    } else {
      const Real h = std::hypot(td, ee);
      mu -= ee * (ee / (td + (td > Real(0) ? h : -h)));
    }
    Real x = d[start] - mu;
    Real zb = e[start];  // the bulge
    for (int k = start; k < end; ++k) {
      // rotation [c -s; s c] with s x + c zb = 0
      // sqrt(x^2+z^2) directly (std::hypot is ~10x slower); fall back to hypot only when the squares
      // over/underflow
      Real r = std::sqrt(x * x + zb * zb);
      if (!(r > tiny) || !(r < std::numeric_limits<Real>::max())) r = std::hypot(x, zb);
      Real c = Real(1), s = Real(0);
      if (r != Real(0)) {
        c = x / r;
        s = -zb / r;
      }
      const Real a = d[k], b = e[k], g = d[k + 1];
      d[k] = c * c * a - Real(2) * c * s * b + s * s * g;
      d[k + 1] = s * s * a + Real(2) * c * s * b + c * c * g;
      e[k] = c * s * (a - g) + (c * c - s * s) * b;
      if (k > start) e[k - 1] = c * e[k - 1] - s * zb;
      x = e[k];
      if (k < end - 1) {
        zb = -s * e[k + 1];
        e[k + 1] = c * e[k + 1];
      }
      if (z) {
        Real* zk = z + static_cast<std::size_t>(k) * n;
        Real* zk1 = zk + n;
        for (int i = 0; i < n; ++i) {
          const Real p = zk[i], q = zk1[i];
          zk[i] = c * p - s * q;
          zk1[i] = s * p + c * q;
        }
      }
    }
  }
  return true;
}

// eigenvalues only, ascending
template <class Real>
bool tridiagonal_eigenvalues(const Real* alpha, const Real* beta, int n, std::vector<Real>& w) {
  w.assign(alpha, alpha + n);
  std::vector<Real> e(beta, beta + (n > 0 ? n - 1 : 0));
  const bool ok = tridiagonal_qr<Real>(n, w.data(), e.data(), nullptr);
  std::sort(w.begin(), w.end());
  return ok;
}

// eigenvalues ascending + eigenvectors (n x n column-major in z)
template <class Real>
bool tridiagonal_eigensystem(const Real* alpha, const Real* beta, int n, std::vector<Real>& w, std::vector<Real>& z) {
  w.assign(alpha, alpha + n);
  std::vector<Real> e(beta, beta + (n > 0 ? n - 1 : 0));
  std::vector<Real> q(static_cast<std::size_t>(n) * n, Real(0));
  for (int i = 0; i < n; ++i) q[static_cast<std::size_t>(i) * n + i] = Real(1);
  const bool ok = tridiagonal_qr<Real>(n, w.data(), e.data(), q.data());
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[a] < w[b]; });
  std::vector<Real> ws(n);
  z.assign(static_cast<std::size_t>(n) * n, Real(0));
  for (int j = 0; j < n; ++j) {
    ws[j] = w[order[j]];
    std::copy(q.begin() + static_cast<std::size_t>(order[j]) * n, q.begin() + static_cast<std::size_t>(order[j] + 1) * n,
              z.begin() + static_cast<std::size_t>(j) * n);
  }
  w.swap(ws);
  return ok;
}

}  // namespace detail
}  // namespace EigenEx
}  // namespace cmpt

#endif
