// detail/hessenberg_eigen.hpp — host solver for the m x m upper-Hessenberg Ritz problem of Arnoldi.
//
// The reference hands the Hessenberg matrix to Eigen::ComplexEigenSolver / Eigen::EigenSolver
// (arnoldi.hpp:472-501,811-821); Eigen is a third-party dependency outside the reference tree.  This is
// an independent implementation of the published algorithm those solvers use for the complex case:
// single-shift QR on the Hessenberg matrix with Givens rotations (complex Schur form H = U T U^H,
// Wilkinson shift, exceptional shifts at iterations 10 and 20), eigenvalues = diag(T), eigenvectors by
// back-substitution on T multiplied by U, each column scaled to unit 2-norm.  A real Hessenberg matrix
// is embedded in complex arithmetic.  Part of the algorithm, not a fallback.
#ifndef CMPT_EIGEN_EX_DETAIL_HESSENBERG_EIGEN_HPP_
#define CMPT_EIGEN_EX_DETAIL_HESSENBERG_EIGEN_HPP_

#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>
#include <vector>

namespace cmpt {
namespace EigenEx {
namespace detail {

template <class Real>
struct HessenbergQR {
  using C = std::complex<Real>;
  int n;
  std::vector<C> t;  // n x n column-major, overwritten by the Schur factor T
  std::vector<C> u;  // n x n Schur vectors (only when want_vectors)
  bool want_vectors;

  static Real norm1(const C& z) { return std::abs(z.real()) + std::abs(z.imag()); }
  C& T(int i, int j) { return t[static_cast<std::size_t>(j) * n + i]; }
  C& U(int i, int j) { return u[static_cast<std::size_t>(j) * n + i]; }

  // unitary J = [c s; -conj(s) c] with J^H [p; q] = [r; 0]
  static void givens(const C& p, const C& q, Real& c, C& s, C& r) {
    const Real ap = std::abs(p), aq = std::abs(q);
    if (aq == Real(0)) {
      c = Real(1);
      s = C(0);
      r = p;
      return;
    }
    const Real nrm = std::hypot(ap, aq);
    const C ph = (ap == Real(0)) ? C(1) : p / ap;
    c = ap / nrm;
    s = -ph * std::conj(q) / nrm;
    r = ph * nrm;
  }
  void rot_left(int i, Real c, const C& s, int j0, int j1) {  // rows i, i+1, columns [j0, j1)
    for (int j = j0; j < j1; ++j) {
      const C a = T(i, j), b = T(i + 1, j);
      T(i, j) = c * a - s * b;
      T(i + 1, j) = std::conj(s) * a + c * b;
    }
  }
  void rot_right(int i, Real c, const C& s, int r0, int r1) {  // columns i, i+1, rows [r0, r1)
    for (int r = r0; r < r1; ++r) {
      const C a = T(r, i), b = T(r, i + 1);
      T(r, i) = c * a - std::conj(s) * b;
      T(r, i + 1) = s * a + c * b;
    }
    if (want_vectors)
      for (int r = 0; r < n; ++r) {
        const C a = U(r, i), b = U(r, i + 1);
        U(r, i) = c * a - std::conj(s) * b;
        U(r, i + 1) = s * a + c * b;
      }
  }
  bool negligible(int i) {  // sub-diagonal entry T(i+1, i)
    const Real d = norm1(T(i, i)) + norm1(T(i + 1, i + 1));
    const Real sd = norm1(T(i + 1, i));
    if (sd <= std::numeric_limits<Real>::epsilon() * d || sd <= std::numeric_limits<Real>::min()) {
      T(i + 1, i) = C(0);
      return true;
    }
    return false;
  }
  C shift(int iu, int iter) {
    if (iter == 10 || iter == 20)  // exceptional shift
      return C(std::abs(T(iu, iu - 1).real()) + (iu >= 2 ? std::abs(T(iu - 1, iu - 2).real()) : Real(0)));
    C t00 = T(iu - 1, iu - 1), t01 = T(iu - 1, iu), t10 = T(iu, iu - 1), t11 = T(iu, iu);
    const Real nt = std::abs(t00) + std::abs(t01) + std::abs(t10) + std::abs(t11);
    if (nt == Real(0)) return C(0);
    t00 /= nt;
    t01 /= nt;
    t10 /= nt;
    t11 /= nt;
    const C b = t01 * t10, cc = t00 - t11;
    const C disc = std::sqrt(cc * cc + Real(4) * b);
    const C det = t00 * t11 - b, tr = t00 + t11;
    C e1 = (tr + disc) / Real(2), e2 = (tr - disc) / Real(2);
    if (norm1(e1) > norm1(e2)) {
      if (norm1(e1) > Real(0)) e2 = det / e1;
    } else if (norm1(e2) > Real(0)) {
      e1 = det / e2;
    }
    return nt * ((norm1(e1 - t11) < norm1(e2 - t11)) ? e1 : e2);
  }

  // Reduce to Schur form.  Without vectors only the active window is updated (eigenvalues only).
  bool reduce() {
    if (want_vectors) {
      u.assign(static_cast<std::size_t>(n) * n, C(0));
      for (int i = 0; i < n; ++i) U(i, i) = C(1);
    }
    int iu = n - 1, iter = 0, total = 0;
    const int max_total = 30 * std::max(n, 1);
    while (true) {
      while (iu > 0) {
        if (!negligible(iu - 1)) break;
        iter = 0;
        --iu;
      }
      if (iu <= 0) break;
      ++iter;
      if (++total > max_total) return false;
      int il = iu - 1;
      while (il > 0 && !negligible(il - 1)) --il;
      const C mu = shift(iu, iter);
      const int cend = want_vectors ? n : iu + 1;  // left rotations: columns up to here
      const int rbeg = want_vectors ? 0 : il;       // right rotations: rows from here
      Real c;
      C s, r;
      givens(T(il, il) - mu, T(il + 1, il), c, s, r);
      rot_left(il, c, s, il, cend);
      rot_right(il, c, s, rbeg, std::min(il + 2, iu) + 1);
      for (int i = il + 1; i < iu; ++i) {
        givens(T(i, i - 1), T(i + 1, i - 1), c, s, r);
        T(i, i - 1) = r;
        T(i + 1, i - 1) = C(0);
        rot_left(i, c, s, i, cend);
        rot_right(i, c, s, rbeg, std::min(i + 2, iu) + 1);
      }
    }
    return true;
  }
};

// Eigenvalues (unsorted, order of the Schur diagonal) of an upper-Hessenberg matrix given column-major
// as complex.  If vectors != nullptr also the unit-norm right eigenvectors (n x n column-major).
template <class Real>
bool hessenberg_eigen(int n, const std::complex<Real>* h, std::vector<std::complex<Real>>& w,
                      std::vector<std::complex<Real>>* vectors) {
  using C = std::complex<Real>;
  HessenbergQR<Real> q;
  q.n = n;
  q.t.assign(h, h + static_cast<std::size_t>(n) * n);
  q.want_vectors = vectors != nullptr;
  // entries below the first sub-diagonal are structurally zero
  for (int j = 0; j < n; ++j)
    for (int i = j + 2; i < n; ++i) q.T(i, j) = C(0);
  const bool ok = q.reduce();
  w.resize(n);
  for (int i = 0; i < n; ++i) w[i] = q.T(i, i);
  if (!vectors) return ok;
  // back-substitution on T (x_k = 1), then V = U X, unit columns
  Real tnorm = 0;
  for (int j = 0; j < n; ++j) {
    Real s = 0;
    for (int i = 0; i <= j; ++i) s += std::abs(q.T(i, j));
    tnorm = std::max(tnorm, s);
  }
  std::vector<C> x(static_cast<std::size_t>(n) * n, C(0));
  auto X = [&](int i, int j) -> C& { return x[static_cast<std::size_t>(j) * n + i]; };
  for (int k = n - 1; k >= 0; --k) {
    X(k, k) = C(1);
    for (int i = k - 1; i >= 0; --i) {
      C acc = -q.T(i, k);
      for (int j = i + 1; j < k; ++j) acc -= q.T(i, j) * X(j, k);
      C z = q.T(i, i) - q.T(k, k);
      if (z == C(0)) z = C(std::numeric_limits<Real>::epsilon() * tnorm);
      X(i, k) = acc / z;
    }
  }
  vectors->assign(static_cast<std::size_t>(n) * n, C(0));
  for (int k = 0; k < n; ++k) {
    C* vk = vectors->data() + static_cast<std::size_t>(k) * n;
    for (int j = 0; j <= k; ++j) {
      const C xj = X(j, k);
      const C* uj = q.u.data() + static_cast<std::size_t>(j) * n;
      for (int i = 0; i < n; ++i) vk[i] += uj[i] * xj;
    }
    Real nrm = 0;
    for (int i = 0; i < n; ++i) nrm += std::norm(vk[i]);
    nrm = std::sqrt(nrm);
    if (nrm > Real(0))
      for (int i = 0; i < n; ++i) vk[i] /= nrm;
  }
  return ok;
}


// Unitary reduction of a general square complex matrix to upper-Hessenberg form by Householder reflectors
// (Golub & Van Loan alg. 7.4.2): A = Q H Q^H.  a is n x n column-major and is overwritten by H; q receives Q.
// Needed after a thick restart of the Arnoldi iteration, when the projected matrix is no longer Hessenberg.
template <class Real>
inline void hessenberg_reduce(int n, std::vector<std::complex<Real>>& a, std::vector<std::complex<Real>>& q) {
  using C = std::complex<Real>;
  auto A = [&](int i, int j) -> C& { return a[static_cast<std::size_t>(j) * n + i]; };
  auto Q = [&](int i, int j) -> C& { return q[static_cast<std::size_t>(j) * n + i]; };
  q.assign(static_cast<std::size_t>(n) * n, C(0));
  for (int i = 0; i < n; ++i) Q(i, i) = C(1);
  std::vector<C> v(static_cast<std::size_t>(n));
  for (int k = 0; k + 2 < n; ++k) {
    Real tail = 0;
    for (int i = k + 2; i < n; ++i) tail += std::norm(A(i, k));
    if (tail == Real(0)) continue;  // column already in Hessenberg form
    const Real alpha = std::sqrt(std::norm(A(k + 1, k)) + tail);
    const C x0 = A(k + 1, k);
    const C phase = (std::abs(x0) == Real(0)) ? C(1) : x0 / std::abs(x0);
    // v = x + phase*alpha*e1, reflector P = I - 2 v v^H / (v^H v)
    for (int i = 0; i < n; ++i) v[static_cast<std::size_t>(i)] = C(0);
    for (int i = k + 1; i < n; ++i) v[static_cast<std::size_t>(i)] = A(i, k);
    v[static_cast<std::size_t>(k + 1)] += phase * alpha;
    Real vn = 0;
    for (int i = k + 1; i < n; ++i) vn += std::norm(v[static_cast<std::size_t>(i)]);
    if (vn == Real(0)) continue;
    for (int j = 0; j < n; ++j) {  // A <- P A
      C s(0);
      for (int i = k + 1; i < n; ++i) s += std::conj(v[static_cast<std::size_t>(i)]) * A(i, j);
      s *= Real(2) / vn;
      for (int i = k + 1; i < n; ++i) A(i, j) -= v[static_cast<std::size_t>(i)] * s;
    }
    for (int i = 0; i < n; ++i) {  // A <- A P, Q <- Q P
      C s(0), t(0);
      for (int j = k + 1; j < n; ++j) {
        s += A(i, j) * v[static_cast<std::size_t>(j)];
        t += Q(i, j) * v[static_cast<std::size_t>(j)];
      }
      s *= Real(2) / vn;
      t *= Real(2) / vn;
      for (int j = k + 1; j < n; ++j) {
        A(i, j) -= s * std::conj(v[static_cast<std::size_t>(j)]);
        Q(i, j) -= t * std::conj(v[static_cast<std::size_t>(j)]);
      }
    }
    for (int i = k + 2; i < n; ++i) A(i, k) = C(0);
  }
}

// Eigenvalues (and unit-norm right eigenvectors) of a general square complex matrix, column-major: Hessenberg
// reduction when the matrix is not Hessenberg already, then hessenberg_eigen, then the back-transformation.
template <class Real>
bool general_eigen(int n, const std::complex<Real>* a, std::vector<std::complex<Real>>& w,
                   std::vector<std::complex<Real>>* vectors) {
  using C = std::complex<Real>;
  bool is_hessenberg = true;
  for (int j = 0; j < n && is_hessenberg; ++j)
    for (int i = j + 2; i < n; ++i)
      if (a[static_cast<std::size_t>(j) * n + i] != C(0)) {
        is_hessenberg = false;
        break;
      }
  if (is_hessenberg) return hessenberg_eigen<Real>(n, a, w, vectors);
  std::vector<C> h(a, a + static_cast<std::size_t>(n) * n), q, y;
  hessenberg_reduce<Real>(n, h, q);
  const bool ok = hessenberg_eigen<Real>(n, h.data(), w, vectors ? &y : nullptr);
  if (vectors) {
    vectors->assign(static_cast<std::size_t>(n) * n, C(0));
    for (int j = 0; j < n; ++j)
      for (int p = 0; p < n; ++p) {
        const C yp = y[static_cast<std::size_t>(j) * n + p];
        const C* qp = q.data() + static_cast<std::size_t>(p) * n;
        C* vj = vectors->data() + static_cast<std::size_t>(j) * n;
        for (int i = 0; i < n; ++i) vj[i] += qp[i] * yp;
      }
  }
  return ok;
}

}  // namespace detail
}  // namespace EigenEx
}  // namespace cmpt

#endif
