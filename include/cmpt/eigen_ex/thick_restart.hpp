// thick_restart.hpp — thick-restart Lanczos on the device-resident basis (additive, the
// reference has no restarted solver).
//
// Algorithm: K. Wu and H. Simon, "Thick-restart Lanczos method for large symmetric eigenvalue problems",
// SIAM J. Matrix Anal. Appl. 22 (2000).  The Krylov basis never holds more than maxBasis() vectors: when it is
// full, the projected matrix T of the first m = maxBasis()-1 vectors is diagonalised on the host, the keep() lowest
// Ritz pairs are kept and the basis is compressed on the device to [y_1..y_k, u_m] (LanczosBase::thickRestart,
// cmb_lanczos_thick_restart).  After a restart T is an arrowhead (diag theta, last row/column beta_{m-1} S(m-1,i))
// followed by the usual tridiagonal tail; it is solved with detail/symmetric_eigen.hpp.  Every step uses the same
// kernels as LanczosEigenSolver (operator apply fused with alpha, CGS2 against the whole basis), so deflation
// vectors (setOrthogonalizingVectors), eigenvalueShift and row-partitioned operators work unchanged.
//
// Convergence: a wanted pair counts as converged when its Ritz residual bound |beta_last * S(last, i)| is at most
// tolerance() * max(1, |theta_i|) — the quantity ritzResiduals() reports for LanczosEigenSolver.
#ifndef CMPT_EIGEN_EX_THICK_RESTART_HPP_
#define CMPT_EIGEN_EX_THICK_RESTART_HPP_

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "detail/symmetric_eigen.hpp"
#include "lanczos.hpp"

namespace cmpt {
namespace EigenEx {

template <class Scalar_>
class ThickRestartLanczos {
 public:
  using Index = EigenEx::Index;
  using Scalar = Scalar_;
  using RealScalar = typename RealOf<Scalar>::type;
  using VectorType = Vector<Scalar>;
  using RealVectorType = Vector<RealScalar>;
  using MatrixType = Matrix<Scalar>;
  using MatMulFunction = std::function<void(const Scalar*, Scalar*)>;

  static std::string headWARN() { return std::string("WARN      "); }
  static std::string headINFO() { return std::string("INFO      "); }

  ThickRestartLanczos() { setAllSettingsDefault(); }

  ThickRestartLanczos& setAllSettingsDefault() {
    wanted_ = 1;
    maxBasis_ = 40;
    keep_ = -1;
    maxRestarts_ = 100;
    tolerance_ = 1.0e-10;
    computeEigenvectorsOn_ = true;
    return *this;
  }

  // ---- settings of the restarted iteration ----
  /// number of lowest eigenpairs to converge
  Index wanted() const { return wanted_; }
  ThickRestartLanczos& setWanted(Index n) {
    wanted_ = n;
    return *this;
  }
  /// largest number of Lanczos vectors held on the device (>= wanted + 3)
  Index maxBasis() const { return maxBasis_; }
  ThickRestartLanczos& setMaxBasis(Index m) {
    maxBasis_ = m;
    return *this;
  }
  /// Ritz pairs kept at a restart; -1 = wanted + min(wanted, (maxBasis - wanted) / 2)  (a few extra pairs speed up
  /// convergence markedly, Wu & Simon section 4)
  Index keep() const { return keep_; }
  ThickRestartLanczos& setKeep(Index k) {
    keep_ = k;
    return *this;
  }
  Index maxRestarts() const { return maxRestarts_; }
  ThickRestartLanczos& setMaxRestarts(Index r) {
    maxRestarts_ = r;
    return *this;
  }
  RealScalar tolerance() const { return tolerance_; }
  ThickRestartLanczos& setTolerance(RealScalar t) {
    tolerance_ = t;
    return *this;
  }
  ThickRestartLanczos& setComputeEigenvectorsOn(bool on) {
    computeEigenvectorsOn_ = on;
    return *this;
  }

  // ---- pass-throughs to the Lanczos basis (same names as LanczosEigenSolver) ----
  ThickRestartLanczos& setMatrixMultiplication(const MatMulFunction& matmul, Index height) {
    base_.setMatrixMultiplication(matmul, height);
    return *this;
  }
  ThickRestartLanczos& setMatrixMultiplication(const DeviceOperator<Scalar>& op) {
    base_.setMatrixMultiplication(op);
    return *this;
  }
  ThickRestartLanczos& setInitialVector(const VectorType& v) {
    base_.setInitialVector(v);
    return *this;
  }
  ThickRestartLanczos& setInitialVector() {
    base_.setInitialVector();
    return *this;
  }
  ThickRestartLanczos& setOrthogonalizingVectors(const std::vector<VectorType>& o) {
    base_.setOrthogonalizingVectors(o);
    return *this;
  }
  ThickRestartLanczos& setEigenvalueShift(RealScalar s) {
    base_.setEigenvalueShift(s);
    return *this;
  }
  ThickRestartLanczos& setThreshold(RealScalar t) {
    base_.setThreshold(t);
    return *this;
  }
  const LanczosBase<Scalar>& lanczosBase() const { return base_; }
  Index localHeight() const { return base_.localHeight(); }

  // ---- results ----
  const RealVectorType& eigenvalues() const { return eigenvalues_; }  ///< lowest wanted() Ritz values, ascending
  const MatrixType& eigenvectors() const { return eigenvectors_; }    ///< local rows x wanted()
  const RealVectorType& residuals() const { return residuals_; }      ///< Ritz residual bounds of eigenvalues()
  Index restarts() const { return restarts_; }
  Index operatorApplications() const { return applies_; }
  Index converged() const { return nconverged_; }  ///< how many of the wanted pairs met the tolerance
  const std::vector<std::string>& log() const { return log_; }
  double deviceBytes() const { return base_.deviceBytes(); }

  /// runs the restarted iteration from the initial vector; returns 0 like LanczosEigenSolver::compute()
  Index compute() {
    log_.clear();
    restarts_ = 0;
    applies_ = 0;
    nconverged_ = 0;
    eigenvalues_.resize(0);
    residuals_.resize(0);
    eigenvectors_.resize(0, 0);
    base_.clearLanczosSteps();
    base_.setReorthogonalizeInterval(1);
    if (wanted_ < 1) throw LanczosException("ThickRestartLanczos: wanted() must be >= 1");
    const Index height = base_.matrixHeight();
    Index mb = std::min<Index>(maxBasis_, height);
    if (mb < std::min<Index>(wanted_ + 3, height))
      throw LanczosException("ThickRestartLanczos: maxBasis() must be at least wanted() + 3");
    base_.setReserveSize(mb + wanted_ + 8);
    Index keep = keep_;
    if (keep < 0) keep = wanted_ + std::min<Index>(wanted_, std::max<Index>(0, (mb - wanted_) / 2));
    keep = std::max<Index>(wanted_, std::min<Index>(keep, mb - 2));

    std::vector<RealScalar> w, z;
    bool exhausted = false;
    for (;;) {
      // fill the basis up to mb vectors (one device chain, one host synchronisation)
      const Index before = base_.lanczosvectorsSize();
      const Index done = base_.updateLanczosSteps(mb - before);
      applies_ += done;
      const Index nv = base_.lanczosvectorsSize();
      if (nv == 0) {
        log_.push_back(headINFO() + "initial lanczosvector generation fail");
        return 0;
      }
      exhausted = base_.lanczosStepIsUtmost() && nv < mb;  // breakdown: the Krylov space is invariant
      // projected matrix of the first m vectors; u_m (when it exists) is the residual direction
      const bool have_residual = !exhausted && static_cast<Index>(base_.beta().size()) >= nv - 1 && nv >= 2;
      const Index m = have_residual ? nv - 1 : nv;
      solveProjected_(m, w, z);
      const RealScalar blast = have_residual ? base_.beta()[static_cast<std::size_t>(m - 1)] : RealScalar(0);
      const Index nw = std::min<Index>(wanted_, m);
      nconverged_ = 0;
      residuals_.resize(nw);
      for (Index i = 0; i < nw; ++i) {
        residuals_[i] = std::abs(blast * z[static_cast<std::size_t>(i) * m + (m - 1)]);
        const RealScalar scale = std::max<RealScalar>(RealScalar(1), std::abs(w[static_cast<std::size_t>(i)]));
        if (residuals_[i] <= tolerance_ * scale) ++nconverged_;
      }
      const bool all = (nconverged_ == nw);
      if (all || exhausted || restarts_ >= maxRestarts_ || m <= keep) {
        if (all)
          log_.push_back(headINFO() + "thick-restart lanczos converged");
        else if (exhausted)
          log_.push_back(headINFO() + "lanczos steps achieved full of Krylov subspace");
        else
          log_.push_back(headWARN() + "thick-restart lanczos achieved maxRestarts");
        finish_(m, nw, w, z);
        return 0;
      }
      // keep the `keep` lowest Ritz pairs
      std::vector<RealScalar> coef(static_cast<std::size_t>(m) * keep), theta(static_cast<std::size_t>(keep)),
          coupling(static_cast<std::size_t>(keep));
      for (Index i = 0; i < keep; ++i) {
        theta[static_cast<std::size_t>(i)] = w[static_cast<std::size_t>(i)];
        coupling[static_cast<std::size_t>(i)] = blast * z[static_cast<std::size_t>(i) * m + (m - 1)];
        std::copy(z.begin() + static_cast<std::size_t>(i) * m, z.begin() + static_cast<std::size_t>(i + 1) * m,
                  coef.begin() + static_cast<std::size_t>(i) * m);
      }
      base_.thickRestart(coef, theta, coupling);
      ++restarts_;
    }
  }

 protected:
  /// eigen-decomposition of the projected matrix of the first m basis vectors (arrowhead + tridiagonal tail)
  void solveProjected_(Index m, std::vector<RealScalar>& w, std::vector<RealScalar>& z) const {
    const std::vector<RealScalar>& a = base_.alpha();
    const std::vector<RealScalar>& b = base_.beta();
    const Index k = std::min<Index>(base_.arrowSize(), m - 1 >= 0 ? m : 0);
    if (base_.arrowSize() == 0) {
      detail::tridiagonal_eigensystem<RealScalar>(a.data(), b.data(), static_cast<int>(m), w, z);
      return;
    }
    std::vector<RealScalar> t(static_cast<std::size_t>(m) * m, RealScalar(0));
    auto T = [&](Index i, Index j) -> RealScalar& { return t[static_cast<std::size_t>(j) * m + i]; };
    for (Index i = 0; i < m; ++i) T(i, i) = a[static_cast<std::size_t>(i)];
    for (Index i = 0; i + 1 < m; ++i) {
      const Index j = (i < k) ? k : i + 1;  // arrowhead rows couple to column k
      if (j < m) T(i, j) = T(j, i) = b[static_cast<std::size_t>(i)];
    }
    detail::symmetric_eigensystem<RealScalar>(static_cast<int>(m), t, w, z);
  }

  void finish_(Index m, Index nw, const std::vector<RealScalar>& w, const std::vector<RealScalar>& z) {
    eigenvalues_.resize(nw);
    for (Index i = 0; i < nw; ++i) eigenvalues_[i] = w[static_cast<std::size_t>(i)] - base_.eigenvalueShift();
    if (!computeEigenvectorsOn_ || nw == 0) {
      eigenvectors_.resize(0, 0);
      return;
    }
    detail::resize_result(eigenvectors_, localHeight(), nw);
    std::vector<Scalar> coef(static_cast<std::size_t>(m) * nw);
    for (Index i = 0; i < nw; ++i)
      for (Index r = 0; r < m; ++r)
        coef[static_cast<std::size_t>(i) * m + r] = Scalar(z[static_cast<std::size_t>(i) * m + r]);
    detail::check(cmb_krylov_ritz_vectors(base_.deviceState(), detail::DTypeOf<Scalar>::value, coef.data(), m, m, nw,
                                          eigenvectors_.data(), localHeight()),
                  "cmb_krylov_ritz_vectors");
  }

  LanczosBase<Scalar> base_;
  Index wanted_, maxBasis_, keep_, maxRestarts_;
  RealScalar tolerance_;
  bool computeEigenvectorsOn_;
  RealVectorType eigenvalues_, residuals_;
  MatrixType eigenvectors_;
  Index restarts_ = 0, applies_ = 0, nconverged_ = 0;
  std::vector<std::string> log_;
};

}  // namespace EigenEx
}  // namespace cmpt

#endif
