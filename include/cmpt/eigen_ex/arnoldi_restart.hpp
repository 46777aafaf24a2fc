// arnoldi_restart.hpp — thick-restart (Krylov-Schur) Arnoldi on the device-resident basis (SURVEY.md §8(f) rank 3 and
// BASELINE cfg 3, "Arnoldi (restarted, m=50)"; additive, the reference has no restarted solver).
//
// Algorithm: G. W. Stewart, "A Krylov-Schur algorithm for large eigenproblems", SIAM J. Matrix Anal. Appl. 23 (2001),
// in the orthonormal-basis form that also serves a real Scalar: the Arnoldi basis never holds more than maxBasis()
// vectors.  When it is full (A Q = Q H + residue q_next e_m^T) the projected matrix H is decomposed on the host, the
// wanted Ritz vectors y_i (plus a few more) span an invariant subspace of H; an orthonormal basis Z of it is formed
// (for a real Scalar a complex-conjugate pair contributes Re y and Im y, so Z stays real) and the basis is compressed on
// the device to Q Z (ArnoldiBase::thickRestart, cmb_arnoldi_thick_restart).  The Krylov decomposition survives:
//     A (Q Z) = (Q Z) T + q_next b^H,   T = Z^H H Z,   b_j = residue * Z(m-1, j),
// and the iteration continues from q_next.  After a restart the projected matrix is no longer Hessenberg (T is full,
// row k is b^T); its eigenproblem goes through detail::general_eigen (Householder reduction + Hessenberg QR).
// Every step uses the same kernels as ArnoldiEigenSolver (operator apply fused with the normalisation, CGS2 against
// the whole basis), so deflation vectors, the shift and row-partitioned operators work unchanged.
//
// Wanted eigenvalues: largest modulus (the reference's order, arnoldi.hpp:813-819), largest real part or smallest real
// part.  Convergence: a wanted pair counts as converged when its Ritz residual residue * |y_i(m-1)| (= ||A x - theta x||
// for the unit Ritz vector x) is at most tolerance() * max(|theta_i|, tiny).
#ifndef CMPT_EIGEN_EX_ARNOLDI_RESTART_HPP_
#define CMPT_EIGEN_EX_ARNOLDI_RESTART_HPP_

#include <algorithm>
#include <cmath>
#include <complex>
#include <numeric>
#include <string>
#include <vector>

#include "arnoldi.hpp"

namespace cmpt {
namespace EigenEx {

template <class Scalar_>
class ThickRestartArnoldi {
 public:
  using Index = EigenEx::Index;
  using Scalar = Scalar_;
  using RealScalar = typename RealOf<Scalar>::type;
  using ComplexScalar = std::complex<RealScalar>;
  using VectorType = Vector<Scalar>;
  using ComplexVectorType = Vector<ComplexScalar>;
  using ComplexMatrixType = Matrix<ComplexScalar>;
  using RealVectorType = Vector<RealScalar>;
  using MatMulFunction = std::function<void(const Scalar*, Scalar*)>;
  enum Which { LargestMagnitude = 0, LargestReal = 1, SmallestReal = 2 };

  static std::string headWARN() { return std::string("WARN      "); }
  static std::string headINFO() { return std::string("INFO      "); }

  ThickRestartArnoldi() { setAllSettingsDefault(); }

  ThickRestartArnoldi& setAllSettingsDefault() {
    wanted_ = 1;
    maxBasis_ = 50;
    keep_ = -1;
    maxRestarts_ = 200;
    tolerance_ = 1.0e-10;
    which_ = LargestMagnitude;
    computeEigenvectorsOn_ = true;
    return *this;
  }

  // ---- settings of the restarted iteration ----
  Index wanted() const { return wanted_; }
  ThickRestartArnoldi& setWanted(Index n) {
    wanted_ = n;
    return *this;
  }
  /// largest number of Arnoldi vectors held on the device (>= wanted + 3)
  Index maxBasis() const { return maxBasis_; }
  ThickRestartArnoldi& setMaxBasis(Index m) {
    maxBasis_ = m;
    return *this;
  }
  /// Ritz vectors kept at a restart; -1 = wanted + (maxBasis - wanted) / 2 (rounded so that conjugate pairs stay together)
  Index keep() const { return keep_; }
  ThickRestartArnoldi& setKeep(Index k) {
    keep_ = k;
    return *this;
  }
  Index maxRestarts() const { return maxRestarts_; }
  ThickRestartArnoldi& setMaxRestarts(Index r) {
    maxRestarts_ = r;
    return *this;
  }
  RealScalar tolerance() const { return tolerance_; }
  ThickRestartArnoldi& setTolerance(RealScalar t) {
    tolerance_ = t;
    return *this;
  }
  Which which() const { return which_; }
  ThickRestartArnoldi& setWhich(Which w) {
    which_ = w;
    return *this;
  }
  ThickRestartArnoldi& setComputeEigenvectorsOn(bool on) {
    computeEigenvectorsOn_ = on;
    return *this;
  }

  // ---- pass-throughs to the Arnoldi basis (same names as ArnoldiEigenSolver) ----
  ThickRestartArnoldi& setMatrixMultiplication(const MatMulFunction& matmul, Index height) {
    base_.setMatrixMultiplication(matmul, height);
    return *this;
  }
  ThickRestartArnoldi& setMatrixMultiplication(const DeviceOperator<Scalar>& op) {
    base_.setMatrixMultiplication(op);
    return *this;
  }
  ThickRestartArnoldi& setInitialVector(const VectorType& v) {
    base_.setInitialVector(v);
    return *this;
  }
  ThickRestartArnoldi& setInitialVector() {
    base_.setInitialVector();
    return *this;
  }
  ThickRestartArnoldi& setOrthogonalizingVectors(const std::vector<VectorType>& o) {
    base_.setOrthogonalizingVectors(o);
    return *this;
  }
  ThickRestartArnoldi& setEigenvalueShift(Scalar s) {
    base_.setEigenvalueShift(s);
    return *this;
  }
  ThickRestartArnoldi& setThreshold(RealScalar t) {
    base_.setThreshold(t);
    return *this;
  }
  const ArnoldiBase<Scalar>& arnoldiBase() const { return base_; }
  Index localHeight() const { return base_.localHeight(); }

  // ---- results ----
  const ComplexVectorType& eigenvalues() const { return eigenvalues_; }    ///< the wanted() Ritz values, in `which` order
  const ComplexMatrixType& eigenvectors() const { return eigenvectors_; }  ///< local rows x wanted(), unit norm, phase-fixed
  const RealVectorType& residuals() const { return residuals_; }           ///< Ritz residuals ||A x - theta x||
  Index restarts() const { return restarts_; }
  Index operatorApplications() const { return applies_; }
  Index converged() const { return nconverged_; }
  const std::vector<std::string>& log() const { return log_; }
  double deviceBytes() const { return base_.deviceBytes(); }

  Index compute() {
    log_.clear();
    restarts_ = 0;
    applies_ = 0;
    nconverged_ = 0;
    eigenvalues_.resize(0);
    residuals_.resize(0);
    eigenvectors_.resize(0, 0);
    base_.clearArnoldiSteps();
    if (wanted_ < 1) throw ArnoldiException("ThickRestartArnoldi: wanted() must be >= 1");
    const Index height = base_.matrixHeight();
    const Index mb = std::min<Index>(maxBasis_, height);
    if (mb < std::min<Index>(wanted_ + 3, height))
      throw ArnoldiException("ThickRestartArnoldi: maxBasis() must be at least wanted() + 3");
    base_.setReserveSize(2 * mb + 2);  // the compressed vectors are assembled behind the basis
    Index keep = keep_;
    if (keep < 0) keep = wanted_ + std::max<Index>(0, (mb - wanted_) / 2);
    keep = std::max<Index>(wanted_, std::min<Index>(keep, mb - 2));

    std::vector<ComplexScalar> w, y;
    std::vector<Index> order;
    for (;;) {
      const Index before = base_.arnoldivectorsSize();
      const Index done = base_.updateArnoldiSteps(mb - before);
      applies_ += done;
      const Index m = base_.arnoldivectorsSize();
      if (m == 0) {
        log_.push_back(headINFO() + "initial arnoldivector generation fail");
        return 0;
      }
      const bool exhausted = base_.arnoldiStepIsUtmost();  // invariant subspace found (or the whole space spanned)
      solveProjected_(m, w, y, order);
      const RealScalar res = base_.residue();
      const Index nw = std::min<Index>(wanted_, m);
      nconverged_ = 0;
      residuals_.resize(nw);
      for (Index i = 0; i < nw; ++i) {
        const Index c = order[static_cast<std::size_t>(i)];
        residuals_[i] = res * std::abs(y[static_cast<std::size_t>(c) * m + (m - 1)]);
        const RealScalar scale = std::max<RealScalar>(std::abs(w[static_cast<std::size_t>(c)]), std::numeric_limits<RealScalar>::min());
        if (residuals_[i] <= tolerance_ * scale) ++nconverged_;
      }
      const bool all = (nconverged_ == nw);
      if (all || exhausted || restarts_ >= maxRestarts_ || m <= keep) {
        if (all)
          log_.push_back(headINFO() + "thick-restart arnoldi converged");
        else if (exhausted)
          log_.push_back(headINFO() + "arnoldi steps achieved full of Krylov subspace");
        else
          log_.push_back(headWARN() + "thick-restart arnoldi achieved maxRestarts");
        finish_(m, nw, w, y, order);
        return 0;
      }
      restart_(m, keep, w, y, order, res);
      ++restarts_;
    }
  }

 protected:
  static bool isReal_() { return !IsComplex<Scalar>::value; }
  template <class S, class Dummy = void>
  struct IsComplex {
    static constexpr bool value = false;
  };
  template <class R, class Dummy>
  struct IsComplex<std::complex<R>, Dummy> {
    static constexpr bool value = true;
  };
  template <class S, class Dummy = void>
  struct Cast {
    static S from(const ComplexScalar& z) { return z.real(); }
  };
  template <class R, class Dummy>
  struct Cast<std::complex<R>, Dummy> {
    static std::complex<R> from(const ComplexScalar& z) { return z; }
  };

  /// the projected matrix of the first m vectors as a dense complex matrix (column-major)
  std::vector<ComplexScalar> projected_(Index m) const {
    const auto h = base_.makeHessenbergMatrix(m);
    std::vector<ComplexScalar> g(static_cast<std::size_t>(m) * m);
    for (Index c = 0; c < m; ++c)
      for (Index r = 0; r < m; ++r) g[static_cast<std::size_t>(c) * m + r] = ComplexScalar(h(r, c));
    return g;
  }

  /// eigen-decomposition of the projected matrix; order = indices of the Ritz values sorted by `which`
  void solveProjected_(Index m, std::vector<ComplexScalar>& w, std::vector<ComplexScalar>& y, std::vector<Index>& order) const {
    const std::vector<ComplexScalar> g = projected_(m);
    detail::general_eigen<RealScalar>(static_cast<int>(m), g.data(), w, &y);
    order.resize(static_cast<std::size_t>(m));
    std::iota(order.begin(), order.end(), Index(0));
    const Which wh = which_;
    std::stable_sort(order.begin(), order.end(), [&](Index a, Index b) {
      const ComplexScalar &za = w[static_cast<std::size_t>(a)], &zb = w[static_cast<std::size_t>(b)];
      if (wh == LargestReal) return za.real() > zb.real();
      if (wh == SmallestReal) return za.real() < zb.real();
      return std::abs(za) > std::abs(zb);
    });
  }

  /// orthonormalises the columns of z (m x k, column-major) in place with two sweeps of modified Gram-Schmidt; columns
  /// that turn out dependent are dropped.  Returns the number of columns kept.
  static Index orthonormalize_(Index m, Index k, std::vector<Scalar>& z) {
    Index kept = 0;
    for (Index j = 0; j < k; ++j) {
      Scalar* zj = z.data() + static_cast<std::size_t>(j) * m;
      RealScalar n0 = 0;
      for (Index r = 0; r < m; ++r) n0 += std::norm(ComplexScalar(zj[r]));
      n0 = std::sqrt(n0);
      for (int sweep = 0; sweep < 2; ++sweep)
        for (Index i = 0; i < kept; ++i) {
          const Scalar* zi = z.data() + static_cast<std::size_t>(i) * m;
          Scalar d = Scalar(0);
          for (Index r = 0; r < m; ++r) d += detail::conj_(zi[r]) * zj[r];
          for (Index r = 0; r < m; ++r) zj[r] -= d * zi[r];
        }
      RealScalar n1 = 0;
      for (Index r = 0; r < m; ++r) n1 += std::norm(ComplexScalar(zj[r]));
      n1 = std::sqrt(n1);
      if (!(n1 > RealScalar(1e-8) * n0) || n0 == RealScalar(0)) continue;  // dependent column
      Scalar* dst = z.data() + static_cast<std::size_t>(kept) * m;
      for (Index r = 0; r < m; ++r) dst[r] = zj[r] / n1;
      ++kept;
    }
    return kept;
  }

  void restart_(Index m, Index keep, const std::vector<ComplexScalar>& w, const std::vector<ComplexScalar>& y,
                const std::vector<Index>& order, RealScalar res) {
    // columns of the subspace to keep: the `keep` best Ritz vectors; for a real Scalar a complex Ritz value brings its
    // conjugate along (Re y and Im y span the same real subspace as y and conj(y))
    std::vector<Scalar> z;
    z.reserve(static_cast<std::size_t>(m) * (keep + 1));
    Index cols = 0;
    std::vector<char> used(static_cast<std::size_t>(m), 0);
    const RealScalar tiny = RealScalar(64) * std::numeric_limits<RealScalar>::epsilon();
    for (Index i = 0; i < m && cols < keep; ++i) {
      const Index c = order[static_cast<std::size_t>(i)];
      if (used[static_cast<std::size_t>(c)]) continue;
      used[static_cast<std::size_t>(c)] = 1;
      const ComplexScalar* yc = y.data() + static_cast<std::size_t>(c) * m;
      const ComplexScalar th = w[static_cast<std::size_t>(c)];
      if (!isReal_() || std::abs(th.imag()) <= tiny * std::abs(th)) {
        if (isReal_()) {
          // a real eigenvalue of a real matrix has a real eigenvector up to a phase: rotate the largest entry onto the real axis
          Index big = 0;
          for (Index r = 1; r < m; ++r)
            if (std::abs(yc[r]) > std::abs(yc[big])) big = r;
          const ComplexScalar ph = std::abs(yc[big]) > 0 ? std::conj(yc[big]) / std::abs(yc[big]) : ComplexScalar(1);
          for (Index r = 0; r < m; ++r) z.push_back(Cast<Scalar>::from(yc[r] * ph));
        } else {
          for (Index r = 0; r < m; ++r) z.push_back(Cast<Scalar>::from(yc[r]));
        }
        ++cols;
      } else {
        // complex pair of a real problem: mark the conjugate as used, keep Re y and Im y
        Index partner = -1;
        RealScalar best = std::numeric_limits<RealScalar>::max();
        for (Index j = 0; j < m; ++j) {
          if (used[static_cast<std::size_t>(j)]) continue;
          const RealScalar d = std::abs(w[static_cast<std::size_t>(j)] - std::conj(th));
          if (d < best) {
            best = d;
            partner = j;
          }
        }
        if (partner >= 0) used[static_cast<std::size_t>(partner)] = 1;
        for (Index r = 0; r < m; ++r) z.push_back(Cast<Scalar>::from(ComplexScalar(yc[r].real(), 0)));
        for (Index r = 0; r < m; ++r) z.push_back(Cast<Scalar>::from(ComplexScalar(yc[r].imag(), 0)));
        cols += 2;
      }
    }
    const Index k = orthonormalize_(m, cols, z);
    z.resize(static_cast<std::size_t>(m) * k);
    // T = Z^H H Z, b = residue * Z(m-1, :)
    const auto h = base_.makeHessenbergMatrix(m);
    std::vector<Scalar> hz(static_cast<std::size_t>(m) * k, Scalar(0)), t(static_cast<std::size_t>(k) * k, Scalar(0)),
        b(static_cast<std::size_t>(k));
    for (Index j = 0; j < k; ++j)
      for (Index c = 0; c < m; ++c) {
        const Scalar zc = z[static_cast<std::size_t>(j) * m + c];
        if (zc == Scalar(0)) continue;
        for (Index r = 0; r < m; ++r) hz[static_cast<std::size_t>(j) * m + r] += h(r, c) * zc;
      }
    for (Index j = 0; j < k; ++j)
      for (Index i = 0; i < k; ++i) {
        Scalar s = Scalar(0);
        for (Index r = 0; r < m; ++r) s += detail::conj_(z[static_cast<std::size_t>(i) * m + r]) * hz[static_cast<std::size_t>(j) * m + r];
        t[static_cast<std::size_t>(j) * k + i] = s;
      }
    for (Index j = 0; j < k; ++j) b[static_cast<std::size_t>(j)] = Scalar(res) * z[static_cast<std::size_t>(j) * m + (m - 1)];
    base_.thickRestart(z, k, t, b);
  }

  void finish_(Index m, Index nw, const std::vector<ComplexScalar>& w, const std::vector<ComplexScalar>& y,
               const std::vector<Index>& order) {
    eigenvalues_.resize(nw);
    for (Index i = 0; i < nw; ++i)
      eigenvalues_[i] = w[static_cast<std::size_t>(order[static_cast<std::size_t>(i)])] - ComplexScalar(base_.eigenvalueShift());
    if (!computeEigenvectorsOn_ || nw == 0) {
      eigenvectors_.resize(0, 0);
      return;
    }
    detail::resize_result(eigenvectors_, localHeight(), nw);
    std::vector<ComplexScalar> coef(static_cast<std::size_t>(m) * nw);
    for (Index i = 0; i < nw; ++i) {
      const ComplexScalar* yc = y.data() + static_cast<std::size_t>(order[static_cast<std::size_t>(i)]) * m;
      std::copy(yc, yc + m, coef.begin() + static_cast<std::size_t>(i) * m);
    }
    detail::check(cmb_krylov_ritz_vectors(base_.deviceState(), CMB_C64, coef.data(), m, m, nw, eigenvectors_.data(),
                                          localHeight()),
                  "cmb_krylov_ritz_vectors");
  }

  ArnoldiBase<Scalar> base_;
  Index wanted_, maxBasis_, keep_, maxRestarts_;
  RealScalar tolerance_;
  Which which_;
  bool computeEigenvectorsOn_;
  ComplexVectorType eigenvalues_;
  ComplexMatrixType eigenvectors_;
  RealVectorType residuals_;
  Index restarts_ = 0, applies_ = 0, nconverged_ = 0;
  std::vector<std::string> log_;
};

}  // namespace EigenEx
}  // namespace cmpt

#endif
