// arnoldi.hpp — drop-in for versmc/cmpt-eigenex include/cmpt/eigen_ex/arnoldi.hpp on B200.
//
// Same namespace, class names, setters/getters, log strings and stop logic as the reference (file:line
// cited at each member).  The Arnoldi vectors live in HBM; each step is the fused operator apply plus
// CGS2 (three passes over the basis) behind the C-ABI of include/cmpt_b200.h.  The m x m Hessenberg Ritz
// problem is solved on the host (detail/hessenberg_eigen.hpp), as in the reference.
//
// Differences a user can observe:
//  * orthogonalisation is CGS2 (h = h1 + h2) instead of one modified Gram-Schmidt sweep (arnoldi.hpp:380-383);
//  * Scalar = double works (the reference only compiles for complex Scalar, arnoldi.hpp:857,864): eigenvalues
//    and eigenvectors are complex either way;
//  * eigenvectors_h() and convergenceLog() compile (the reference's have a typo / wrong type, :666,:671);
//  * additive: setMatrixMultiplication(DeviceOperator), ritzResiduals(), explicit restart (restart()).
#ifndef CMPT_EIGEN_EX_ARNOLDI_HPP_
#define CMPT_EIGEN_EX_ARNOLDI_HPP_

#include <algorithm>
#include <cmath>
#include <complex>
#include <functional>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "detail/hessenberg_eigen.hpp"
#include "detail/convergence.hpp"
#include "detail/krylov_device.hpp"
#include "device.hpp"
#include "lanczos.hpp"
#include "random.hpp"

namespace cmpt {
namespace EigenEx {

using ArnoldiException = EigenEx::LanczosException;  // arnoldi.hpp:45

/// Stand-in for Eigen::ComplexEigenSolver / Eigen::EigenSolver on the Hessenberg matrix (arnoldi.hpp:472-501).
template <class Scalar>
class HessenbergEigenSolver {
 public:
  using RealScalar = typename RealOf<Scalar>::type;
  using ComplexScalar = std::complex<RealScalar>;
  using ComplexVectorType = Vector<ComplexScalar>;
  using ComplexMatrixType = Matrix<ComplexScalar>;
  using MatrixType = Matrix<Scalar>;

  HessenbergEigenSolver& compute(const MatrixType& h, bool computeEigenvectors = true) {
    const int n = static_cast<int>(h.rows());
    std::vector<ComplexScalar> hc(static_cast<std::size_t>(n) * n);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) hc[static_cast<std::size_t>(j) * n + i] = ComplexScalar(h(i, j));
    std::vector<ComplexScalar> w, v;
    converged_ = detail::hessenberg_eigen<RealScalar>(n, hc.data(), w, computeEigenvectors ? &v : nullptr);
    eivals_.resize(n);
    for (int i = 0; i < n; ++i) eivals_[i] = w[i];
    if (computeEigenvectors) {
      eivecs_.resize(n, n);
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) eivecs_(i, j) = v[static_cast<std::size_t>(j) * n + i];
    } else {
      eivecs_.resize(0, 0);
    }
    return *this;
  }
  const ComplexVectorType& eigenvalues() const { return eivals_; }
  const ComplexMatrixType& eigenvectors() const { return eivecs_; }
  bool converged() const { return converged_; }

 private:
  ComplexVectorType eivals_;
  ComplexMatrixType eivecs_;
  bool converged_ = true;
};

/// This class generates the basis of the Krylov subspace (arnoldi.hpp:53-438): h_ij and the Arnoldi vectors.
template <class Scalar_>
class ArnoldiBase {
 public:
  using Index = EigenEx::Index;
  using Scalar = Scalar_;
  using RealScalar = typename RealOf<Scalar>::type;
  using VectorType = Vector<Scalar>;
  using RealVectorType = Vector<RealScalar>;
  using MatrixType = Matrix<Scalar>;
  using ScalarDistribution = typename NormalDistributionGen<Scalar>::Type;
  using VectorDistribution = typename EigenEx::VectorDistribution<ScalarDistribution>;
  using MatMulFunction = std::function<void(const Scalar*, Scalar*)>;

  template <class URBG>
  static VectorType makeRandomVector(URBG& g, Index size) {
    VectorDistribution vdist(ScalarDistribution(), size, true);
    return vdist(g);
  }

 protected:
  Index reserveSize_;
  std::vector<VectorType> orthogonalizingVectors_;
  MatMulFunction matrixMultiplication_;
  DeviceOperator<Scalar> deviceOperator_;
  Scalar eigenvalueShift_;
  Index matrixHeight_;
  VectorType initialVector_;
  std::uint64_t initialVersion_ = 1;  // bumped by every setInitialVector: the device keeps the last uploaded version
  RealScalar threshold_;

 public:
  // accessors for parameters of settings of arnoldi computing (arnoldi.hpp:113-173)
  Index reserveSize() const { return reserveSize_; }
  ArnoldiBase& setReserveSize(Index resSize) {
    reserveSize_ = resSize;
    return *this;
  }
  const std::vector<VectorType>& orthogonalizingVectors() const { return orthogonalizingVectors_; }
  std::vector<VectorType>& refOrthogonalizingVectors() { return orthogonalizingVectors_; }
  ArnoldiBase& setOrthogonalizingVectors(const std::vector<VectorType>& orthoVec) {
    orthogonalizingVectors_ = orthoVec;
    return *this;
  }
  ArnoldiBase& setOrthogonalizingVectors(std::vector<VectorType>&& orthoVec) {
    orthogonalizingVectors_.swap(orthoVec);
    return *this;
  }
  const MatMulFunction& matrixMultiplication() const { return matrixMultiplication_; }
  ArnoldiBase& setMatrixMultiplication(const MatMulFunction& matmul, Index height) {
    matrixMultiplication_ = matmul;
    deviceOperator_ = DeviceOperator<Scalar>();
    matrixHeight_ = height;
    return *this;
  }
  ArnoldiBase& setMatrixMultiplication(MatMulFunction&& matmul, Index height) {
    std::swap(matrixMultiplication_, matmul);
    deviceOperator_ = DeviceOperator<Scalar>();
    matrixHeight_ = height;
    return *this;
  }
  /// additive overload: operator resident in HBM
  ArnoldiBase& setMatrixMultiplication(const DeviceOperator<Scalar>& op) {
    deviceOperator_ = op;
    DeviceOperator<Scalar> held = op;
    matrixMultiplication_ = [held](const Scalar* in, Scalar* out) { held.apply(in, out); };
    matrixHeight_ = op.height();
    return *this;
  }
  const DeviceOperator<Scalar>& deviceOperator() const { return deviceOperator_; }
  Index matrixHeight() const { return matrixHeight_; }
  /// additive: rows held by this process (see LanczosBase::localHeight)
  Index localHeight() const { return deviceOperator_ ? deviceOperator_.rows() : matrixHeight_; }
  Index localRowBegin() const { return deviceOperator_ ? deviceOperator_.rowBegin() : 0; }
  Scalar eigenvalueShift() const { return eigenvalueShift_; }
  ArnoldiBase& setEigenvalueShift(Scalar eishift) {
    eigenvalueShift_ = eishift;
    return *this;
  }
  const VectorType& initialVector() const { return initialVector_; }
  ArnoldiBase& setInitialVector(const VectorType& inivec) {
    initialVector_ = inivec;
    ++initialVersion_;
    return *this;
  }
  ArnoldiBase& setInitialVector(VectorType&& inivec) {
    initialVector_ = std::move(inivec);
    ++initialVersion_;
    return *this;
  }
  /// additive: copy n scalars straight into the (pinned, reused) start-vector storage
  ArnoldiBase& setInitialVector(const Scalar* data, Index n) {
    detail::assign_upload(initialVector_, data, n);
    ++initialVersion_;
    return *this;
  }
  ArnoldiBase& setInitialVector() {  // arnoldi.hpp:162-166
    std::mt19937 rengine;
    VectorType full = makeRandomVector(rengine, matrixHeight_);
    if (localHeight() == matrixHeight_) {
      setInitialVector(std::move(full));
    } else {
      VectorType slab(localHeight());
      for (Index i = 0; i < localHeight(); ++i) slab[i] = full[localRowBegin() + i];
      setInitialVector(std::move(slab));
    }
    return *this;
  }
  RealScalar threshold() const { return threshold_; }
  ArnoldiBase& setThreshold(RealScalar thre) {
    threshold_ = thre;
    return *this;
  }

 protected:
  Index iterations_;
  Index nvectors_;
  mutable std::vector<VectorType> arnoldivectors_;  // host cache, filled on demand
  RealScalar residue_;
  std::vector<std::vector<Scalar>> h_;
  Index restartedAt_ = -1;  // > 0 after thickRestart(k): column k-1 of h_ already holds its full (k+1)-entry column
  detail::KrylovDevice<Scalar> dev_;

 public:
  Index iterations() const { return iterations_; }
  Index arnoldivectorsSize() const { return nvectors_; }
  const std::vector<VectorType>& arnoldivectors() const {
    if (static_cast<Index>(arnoldivectors_.size()) > nvectors_) arnoldivectors_.resize(nvectors_);
    while (static_cast<Index>(arnoldivectors_.size()) < nvectors_) {
      VectorType v(localHeight());
      detail::check(cmb_krylov_get_col(dev_.handle(), static_cast<std::int64_t>(arnoldivectors_.size()), v.data()),
                    "cmb_krylov_get_col");
      arnoldivectors_.push_back(std::move(v));
    }
    return arnoldivectors_;
  }
  const std::vector<std::vector<Scalar>>& h() const { return h_; }
  RealScalar residue() const { return residue_; }
  cmb_krylov* deviceState() const { return dev_.handle(); }
  double deviceBytes() const { return dev_.ready() ? cmb_krylov_bytes(dev_.handle()) : 0.0; }

 public:
  ArnoldiBase() : iterations_(0), nvectors_(0), residue_(0) { setAllSettingsDefault(); }

  ArnoldiBase& setAllSettingsDefault() {  // arnoldi.hpp:208-218
    setReserveSize(128);
    setOrthogonalizingVectors(std::vector<VectorType>());
    setMatrixMultiplication([](const Scalar*, Scalar*) {}, 0);
    setEigenvalueShift(0.0);
    setInitialVector();
    setThreshold(DefaultTolerance<RealScalar>::value());
    return *this;
  }

  void clearArnoldiSteps() {  // arnoldi.hpp:224-229
    iterations_ = 0;
    nvectors_ = 0;
    arnoldivectors_.clear();
    h_.clear();
    residue_ = 0;
    restartedAt_ = -1;
    if (dev_.ready()) detail::check(cmb_krylov_clear(dev_.handle()), "cmb_krylov_clear");
  }

  void clear() {  // arnoldi.hpp:235-238
    clearArnoldiSteps();
    setAllSettingsDefault();
    dev_.release();
  }

  /// arnoldi.hpp:277-288
  bool arnoldiStepIsUtmost() const {
    if (nvectors_ == 0) return false;
    if (nvectors_ == matrixHeight_) return true;
    if (residue_ <= threshold_) return true;
    return false;
  }

  /// one Arnoldi step (arnoldi.hpp:312-392)
  bool updateArnoldiSteps() { return updateArnoldiSteps(1) == 1; }

  /// additive: up to `count` consecutive calls of updateArnoldiSteps() enqueued on the device with a single host
  /// synchronisation at the end.  Returns how many of them returned true.
  Index updateArnoldiSteps(Index count) {
    if (matrixHeight_ <= 0) return 0;
    if (!matrixMultiplication_ && !deviceOperator_) return 0;
    if (count <= 0) return 0;
    dev_.prepare(deviceOperator_, matrixMultiplication_, matrixHeight_, reserveSize_);
    if (nvectors_ == 0) {
      // setInitialArnoldivector (arnoldi.hpp:245-269)
      if (matrixHeight_ < 0) throw ArnoldiException("matrixHeight_ < 0");
      if (localHeight() != static_cast<Index>(initialVector_.size())) setInitialVector();
      dev_.setDeflation(orthogonalizingVectors_, localHeight());
      int st = 0;
      dev_.start(initialVector_, initialVersion_, threshold_, &st);
      if (st != CMB_STEP_OK) return 0;
    } else if (arnoldiStepIsUtmost()) {
      return 0;
    }
    const Index k0 = nvectors_;
    if (count > matrixHeight_ - k0) count = matrixHeight_ - k0;
    if (count <= 0) return 0;
    const Index ldh = k0 + count + 2;
    std::vector<Scalar> cols(static_cast<std::size_t>(ldh) * static_cast<std::size_t>(count), Scalar(0));
    std::vector<double> res(static_cast<std::size_t>(count), 0.0);
    std::int64_t done = 0;
    int status = 0;
    int rc = cmb_arnoldi_run(dev_.handle(), dev_.op(), &eigenvalueShift_, threshold_, count, cols.data(), ldh, res.data(),
                             &done, &status);
    dev_.rethrowCallbackError();
    detail::check(rc, "cmb_arnoldi_run");
    for (Index s = 0; s < static_cast<Index>(done); ++s) {
      const Index k = k0 + s;
      if (k > 0 && k != restartedAt_) {
        h_[k - 1].resize(k + 1);
        h_[k - 1][k] = Scalar(residue_);  // sub-diagonal entry of the previous column (arnoldi.hpp:362-363)
      }
      std::vector<Scalar> col(static_cast<std::size_t>(k) + 2, Scalar(0));
      for (Index i = 0; i <= k; ++i) col[i] = cols[static_cast<std::size_t>(s) * ldh + i];
      h_.push_back(col);
      residue_ = res[static_cast<std::size_t>(s)];
      nvectors_ = k + 1;
      ++iterations_;
    }
    return static_cast<Index>(done);
  }

  /// additive (thick restart, Krylov-Schur style; arnoldi_restart.hpp): with m = arnoldivectorsSize() vectors and the
  /// residual of the last step pending, replace the basis by Q Z, where Z (m x k, column-major in coef) is an orthonormal
  /// basis of an invariant subspace of the projected matrix H.  The projected matrix becomes [T ; b^T] in its first k
  /// columns (T = Z^H H Z supplied by the caller column-major in t, b_j = residue * Z(m-1, j) in b); the next
  /// updateArnoldiSteps() continues from the pending residual vector, which becomes vector number k.
  void thickRestart(const std::vector<Scalar>& coef, Index k, const std::vector<Scalar>& t, const std::vector<Scalar>& b) {
    const Index m = nvectors_;
    if (m < 1 || k < 1 || k > m || static_cast<Index>(coef.size()) != m * k || static_cast<Index>(t.size()) != k * k ||
        static_cast<Index>(b.size()) != k)
      throw ArnoldiException("thickRestart: bad shapes");
    detail::check(cmb_arnoldi_thick_restart(dev_.handle(), coef.data(), m, m, k), "cmb_arnoldi_thick_restart");
    h_.assign(static_cast<std::size_t>(k), std::vector<Scalar>());
    for (Index c = 0; c < k; ++c) {
      std::vector<Scalar>& col = h_[static_cast<std::size_t>(c)];
      col.resize(static_cast<std::size_t>(k) + 1);
      for (Index r = 0; r < k; ++r) col[static_cast<std::size_t>(r)] = t[static_cast<std::size_t>(c) * k + r];
      col[static_cast<std::size_t>(k)] = b[static_cast<std::size_t>(c)];
    }
    nvectors_ = k;
    restartedAt_ = k;
    arnoldivectors_.clear();
  }

  /// dense copy of the basis, n x (number of Hessenberg columns) (arnoldi.hpp:398-409)
  MatrixType makeArnoldiMatrix() const {
    Index nr = localHeight();
    Index nc = static_cast<Index>(h_.size());
    if (nc > matrixHeight_) nc = matrixHeight_;
    MatrixType V(nr, nc);
    const auto& vecs = arnoldivectors();
    for (Index c = 0; c < nc; ++c)
      for (Index r = 0; r < nr; ++r) V(r, c) = vecs[c][r];
    return V;
  }

  /// Hessenberg matrix of the current step (arnoldi.hpp:415-432)
  MatrixType makeHessenbergMatrix() const { return makeHessenbergMatrix(static_cast<Index>(h_.size())); }
  /// additive: the leading hsize x hsize block, i.e. the Hessenberg matrix as it was after hsize steps
  MatrixType makeHessenbergMatrix(Index hsize) const {
    if (hsize > static_cast<Index>(h_.size())) hsize = static_cast<Index>(h_.size());
    if (hsize > matrixHeight_) hsize = matrixHeight_;
    MatrixType hess = MatrixType::Zero(hsize, hsize);
    for (Index c = 0, nc = hess.cols(); c < nc; ++c) {
      Index nr = hess.rows();
      Index nr_ = static_cast<Index>(h_[c].size());
      if (nr_ < nr) nr = nr_;
      for (Index r = 0; r < nr; ++r) hess(r, c) = h_[c][r];
    }
    return hess;
  }
};

/// eigen solver with Arnoldi (arnoldi.hpp:444-1027)
template <class Scalar_>
class ArnoldiEigenSolver {
 public:
  using Index = EigenEx::Index;
  using Scalar = Scalar_;
  using RealScalar = typename RealOf<Scalar>::type;
  using ComplexScalar = std::complex<RealScalar>;
  using VectorType = Vector<Scalar>;
  using RealVectorType = Vector<RealScalar>;
  using ComplexVectorType = Vector<ComplexScalar>;
  using MatrixType = Matrix<Scalar>;
  using RealMatrixType = Matrix<RealScalar>;
  using ComplexMatrixType = Matrix<ComplexScalar>;
  using ScalarDistribution = typename NormalDistributionGen<Scalar>::Type;
  using VectorDistribution = typename EigenEx::VectorDistribution<ScalarDistribution>;
  using MatMulFunction = std::function<void(const Scalar*, Scalar*)>;
  using DenseEigenSolver = HessenbergEigenSolver<Scalar>;

  static std::string headERROR() { return std::string("ERROR     "); }
  static std::string headWARN() { return std::string("WARN      "); }
  static std::string headINFO() { return std::string("INFO      "); }
  static std::string headDEBUG() { return std::string("DEBUG     "); }

  static constexpr Index unlimited = -1;

  template <class URBG>
  static VectorType makeRandomVector(URBG& g, Index size) {
    return ArnoldiBase<Scalar>::makeRandomVector(g, size);
  }

 protected:
  Index minIterations_;
  Index maxIterations_;
  RealScalar tolerance_;
  std::vector<Index> indicesForConvergence_;  // order of eigenvalues: descending by absolute value
  Index maxEigenvalues_;
  bool computeEigenvectorsOn_;

 public:
  Index minIterations() const { return minIterations_; }
  ArnoldiEigenSolver& setMinIterations(Index miniter) {
    minIterations_ = miniter;
    return *this;
  }
  Index maxIterations() const { return maxIterations_; }
  ArnoldiEigenSolver& setMaxIterations(Index maxiter) {
    maxIterations_ = maxiter;
    return *this;
  }
  RealScalar tolerance() const { return tolerance_; }
  ArnoldiEigenSolver& setTolerance(RealScalar toler) {
    tolerance_ = toler;
    return *this;
  }
  const std::vector<Index>& indicesForConvergence() const { return indicesForConvergence_; }
  ArnoldiEigenSolver& setIndicesForConvergence(const std::vector<Index>& iCovs) {
    resolvePending_();  // pending log entries belong to the old index set
    indicesForConvergence_ = iCovs;
    return *this;
  }
  Index maxEigenvalues() const { return maxEigenvalues_; }
  ArnoldiEigenSolver& setMaxEigenvalues(Index maxeivals) {
    maxEigenvalues_ = maxeivals;
    return *this;
  }
  Index computeEigenvectorsOn() const { return computeEigenvectorsOn_; }
  ArnoldiEigenSolver& setComputeEigenvectorsOn(bool cEivecOn) {
    computeEigenvectorsOn_ = cEivecOn;
    return *this;
  }

 protected:
  ArnoldiBase<Scalar> arnoldiBase_;

 public:  // transparent accessors (arnoldi.hpp:582-642)
  const ArnoldiBase<Scalar>& arnoldiBase() const { return arnoldiBase_; }
  Index reserveSize() const { return arnoldiBase_.reserveSize(); }
  ArnoldiEigenSolver& setReserveSize(Index resSize) {
    arnoldiBase_.setReserveSize(resSize);
    return *this;
  }
  const std::vector<VectorType>& orthogonalizingVectors() const { return arnoldiBase_.orthogonalizingVectors(); }
  std::vector<VectorType>& refOrthogonalizingVectors() { return arnoldiBase_.refOrthogonalizingVectors(); }
  ArnoldiEigenSolver& setOrthogonalizingVectors(const std::vector<VectorType>& orthoVec) {
    arnoldiBase_.setOrthogonalizingVectors(orthoVec);
    return *this;
  }
  ArnoldiEigenSolver& setOrthogonalizingVectors(std::vector<VectorType>&& orthoVec) {
    arnoldiBase_.setOrthogonalizingVectors(std::move(orthoVec));
    return *this;
  }
  const MatMulFunction& matrixMultiplication() const { return arnoldiBase_.matrixMultiplication(); }
  ArnoldiEigenSolver& setMatrixMultiplication(const MatMulFunction& matmul, Index height) {
    arnoldiBase_.setMatrixMultiplication(matmul, height);
    return *this;
  }
  ArnoldiEigenSolver& setMatrixMultiplication(MatMulFunction&& matmul, Index height) {
    arnoldiBase_.setMatrixMultiplication(std::move(matmul), height);
    return *this;
  }
  ArnoldiEigenSolver& setMatrixMultiplication(const DeviceOperator<Scalar>& op) {
    arnoldiBase_.setMatrixMultiplication(op);
    return *this;
  }
  Index matrixHeight() const { return arnoldiBase_.matrixHeight(); }
  Index localHeight() const { return arnoldiBase_.localHeight(); }
  Scalar eigenvalueShift() const { return arnoldiBase_.eigenvalueShift(); }
  ArnoldiEigenSolver& setEigenvalueShift(Scalar eishift) {
    arnoldiBase_.setEigenvalueShift(eishift);
    return *this;
  }
  const VectorType& initialVector() const { return arnoldiBase_.initialVector(); }
  ArnoldiEigenSolver& setInitialVector(const VectorType& inivec) {
    arnoldiBase_.setInitialVector(inivec);
    return *this;
  }
  ArnoldiEigenSolver& setInitialVector(VectorType&& inivec) {
    arnoldiBase_.setInitialVector(std::move(inivec));
    return *this;
  }
  ArnoldiEigenSolver& setInitialVector(const Scalar* data, Index n) {
    arnoldiBase_.setInitialVector(data, n);
    return *this;
  }
  ArnoldiEigenSolver& setInitialVector() {
    arnoldiBase_.setInitialVector();
    return *this;
  }
  RealScalar threshold() const { return arnoldiBase_.threshold(); }
  ArnoldiEigenSolver& setThreshold(RealScalar thre) {
    arnoldiBase_.setThreshold(thre);
    return *this;
  }
  Index iterations() const { return arnoldiBase_.iterations(); }
  const std::vector<VectorType>& arnoldivectors() const { return arnoldiBase_.arnoldivectors(); }

 protected:
  ComplexVectorType eigenvalues_;
  ComplexMatrixType eigenvectors_;
  ComplexMatrixType eigenvectors_h_;
  std::vector<std::string> log_;
  MatrixType hessenbergMatrix_;
  DenseEigenSolver des_;
  mutable std::map<Index, std::vector<ComplexScalar>> convergenceLog_;
  // trips skipped by batched stepping whose convergence-log entries have not been computed yet (resolved on the
  // first access to the log): {first state, number of calls, insert position per tracked index}
  struct PendingReplay {
    Index before, done;
    std::map<Index, std::size_t> pos;
  };
  mutable std::vector<PendingReplay> pendingReplay_;

 public:
  const ComplexVectorType& eigenvalues() const { return eigenvalues_; }
  const ComplexMatrixType& eigenvectors() const { return eigenvectors_; }
  const ComplexMatrixType& eigenvectors_h() const { return eigenvectors_h_; }
  const std::vector<std::string>& log() const { return log_; }
  const MatrixType& hessenbergMatrix() const { return hessenbergMatrix_; }
  const DenseEigenSolver& des() const { return des_; }
  const std::map<Index, std::vector<ComplexScalar>>& convergenceLog() const {
    resolvePending_();
    return convergenceLog_;
  }

  /// additive: Ritz residual bounds residue * |Y(last, i)| of the returned eigenpairs
  RealVectorType ritzResiduals() const {
    RealVectorType r(eigenvalues_.size());
    const Index m = eigenvectors_h_.rows();
    for (Index i = 0; i < static_cast<Index>(eigenvalues_.size()); ++i)
      r[i] = (m > 0 && i < eigenvectors_h_.cols()) ? arnoldiBase_.residue() * std::abs(eigenvectors_h_(m - 1, i)) : RealScalar(0);
    return r;
  }

 public:
  ArnoldiEigenSolver() { setAllSettingsDefault(); }

  ArnoldiEigenSolver& setAllSettingsDefault() {  // arnoldi.hpp:681-692
    setMinIterations(1);
    setMaxIterations(unlimited);
    setTolerance(DefaultTolerance<RealScalar>::value());
    setIndicesForConvergence(std::vector<Index>{0});
    setMaxEigenvalues(unlimited);
    setComputeEigenvectorsOn(true);
    arnoldiBase_.setAllSettingsDefault();
    return *this;
  }

  ArnoldiEigenSolver& clearComputedData() {  // arnoldi.hpp:699-706
    arnoldiBase_.clearArnoldiSteps();
    eigenvalues_.resize(0);
    eigenvectors_.resize(0, 0);
    log_.clear();
    convergenceLog_.clear();
    pendingReplay_.clear();
    return *this;
  }

  ArnoldiEigenSolver& clear() {  // arnoldi.hpp:713-717
    clearComputedData();
    setAllSettingsDefault();
    return *this;
  }

  Index continueToCompute() {  // arnoldi.hpp:725-736
    log_.push_back(headINFO() + "ArnoldiEigenSolver<ScalarType>::continueToCompute(...) was called");
    if (arnoldiBase_.arnoldivectorsSize() == 0) return compute();
    Index ret = mainCalculation_();
    log_.push_back(headINFO() + "ArnoldiEigenSolver<ScalarType>::compute(...) finish computing");
    return ret;
  }

  Index compute() {  // arnoldi.hpp:741-760
    log_.push_back(headINFO() + "ArnoldiEigenSolver<ScalarType>::compute(...) was called");
    clearComputedData();
    if (static_cast<Index>(initialVector().size()) != localHeight()) {
      log_.push_back(headINFO() + "in compute(), initial_vector is empty or invalid, then set at random");
      setInitialVector();
    }
    Index ret = mainCalculation_();
    log_.push_back(headINFO() + "ArnoldiEigenSolver<ScalarType>::compute(...) finish computing");
    return ret;
  }

  /// additive (BASELINE cfg 3, "restarted"): explicit restart — `cycles` runs of compute(), each started from the
  /// real part (real Scalar) or the value (complex Scalar) of the leading Ritz vector of the previous run.
  /// The reference has no restart; one cycle is exactly compute().
  Index computeWithRestarts(Index cycles) {
    Index ret = 0;
    for (Index c = 0; c < cycles; ++c) {
      if (c > 0) {
        if (eigenvectors_.cols() == 0) break;
        VectorType next(localHeight());
        for (Index i = 0; i < localHeight(); ++i) next[i] = fromComplex_(eigenvectors_(i, 0));
        setInitialVector(std::move(next));
      }
      ret = compute();
    }
    return ret;
  }

  Index mainCalculation_() {  // arnoldi.hpp:764-873
    bool set_initialvector_is_fail = false;
    while (true) {
      recordTrip_();
      {
        if (set_initialvector_is_fail) {
          log_.push_back(headINFO() + "initial arnoldivector generation fail");
          break;
        }
        if (arnoldiBase_.arnoldiStepIsUtmost()) {
          log_.push_back(headINFO() + "arnoldi steps finished with threshold");
          log_.push_back(headINFO() + "arnoldi steps achieved full of Krylov subspace");
          break;
        }
        if (arnoldiBase_.iterations() >= minIterations()) {
          if (arnoldiBase_.iterations() == maxIterations()) {
            log_.push_back(headWARN() + "arnoldi steps achieved maxIterations");
            break;
          }
          if (watchedValuesSettled_()) {
            log_.push_back(headINFO() + "arnoldi steps converged with tolerance");
            break;
          }
        }
      }
      // steps no stop rule can interrupt (iterations < minIterations) go to the device as one batch; their
      // per-trip bookkeeping is replayed afterwards from the Hessenberg columns (same values, same logs)
      Index batch = 1;
      if (arnoldiBase_.iterations() < minIterations()) {
        batch = minIterations() - arnoldiBase_.iterations();
        const Index room = matrixHeight() - arnoldiBase_.arnoldivectorsSize();
        if (batch > room) batch = room;
        if (batch < 1) batch = 1;
      }
      const Index before = arnoldiBase_.arnoldivectorsSize();
      const Index done = arnoldiBase_.updateArnoldiSteps(batch);
      if (arnoldiBase_.arnoldivectorsSize() == 0) set_initialvector_is_fail = true;
      deferReplay_(before, done);
      solveHessenberg_(false);
    }
    // eigenvectors of H are needed once, at exit (the reference recomputes them every trip)
    solveHessenberg_(true);

    // back eigen value to original one (arnoldi.hpp:828-838)
    Index eivalsize = eigenvalues_.size();
    if (maxEigenvalues_ != unlimited) {
      if (maxEigenvalues_ < eivalsize) eivalsize = maxEigenvalues_;
    }
    ComplexVectorType eivals_temp = eigenvalues_;
    eigenvalues_.resize(eivalsize);
    for (Index k = 0; k < eivalsize; ++k) eigenvalues_[k] = eivals_temp[k] - ComplexScalar(arnoldiBase_.eigenvalueShift());

    // Ritz vectors X = Q Y, normalised, phase-fixed (arnoldi.hpp:841-865) — assembled on the device
    if (computeEigenvectorsOn_) {
      if (!(eivalsize > 0 && eigenvectors_h_.rows() > 0)) {
        eigenvectors_ = ComplexMatrixType::Zero(localHeight(), eivalsize);
      } else {
        detail::resize_result(eigenvectors_, localHeight(), eivalsize);
        const Index nm = eigenvectors_h_.rows();
        std::vector<ComplexScalar> coef(static_cast<std::size_t>(nm) * eivalsize);
        for (Index c = 0; c < eivalsize; ++c)
          for (Index j = 0; j < nm; ++j) coef[static_cast<std::size_t>(c) * nm + j] = eigenvectors_h_(j, c);
        detail::check(cmb_krylov_ritz_vectors(arnoldiBase_.deviceState(), CMB_C64, coef.data(), nm, nm, eivalsize,
                                              eigenvectors_.data(), localHeight()),
                      "cmb_krylov_ritz_vectors");
      }
    } else {
      eigenvectors_.resize(0, 0);
    }
    return 0;
  }

 protected:
  static Scalar fromComplex_(const ComplexScalar& z) { return FromComplex<Scalar>::get(z); }
  template <class S, class Dummy = void>
  struct FromComplex {
    static S get(const ComplexScalar& z) { return z.real(); }
  };
  template <class R, class Dummy>
  struct FromComplex<std::complex<R>, Dummy> {
    static std::complex<R> get(const ComplexScalar& z) { return z; }
  };

  /// The trips a batched updateArnoldiSteps(done) skipped (states with before+1 .. before+done-1 Arnoldi vectors)
  /// owe the convergence log one entry per tracked index each.  They are recorded here and computed only when the
  /// log is read (convergenceLog(), the convergence test): a fixed-length run never pays for them inside compute().
  void deferReplay_(Index before, Index done) {
    if (done - 1 <= 0) return;
    PendingReplay p;
    p.before = before;
    p.done = done;
    for (auto& idx : indicesForConvergence_) {
      auto it = convergenceLog_.find(idx);
      p.pos[idx] = (it == convergenceLog_.end()) ? 0 : it->second.size();
    }
    pendingReplay_.push_back(p);
  }
  void resolvePending_() const {
    // later batches first: inserting them does not move the insert positions of earlier ones
    for (auto it = pendingReplay_.rbegin(); it != pendingReplay_.rend(); ++it) replayTrips_(*it);
    pendingReplay_.clear();
  }
  /// Eigenvalues of the leading Hessenberg blocks of the skipped states, sorted by descending |lambda|, computed on
  /// a few host threads and inserted in trip order at the recorded positions.
  void replayTrips_(const PendingReplay& pr) const {
    const Index before = pr.before, ntrips = pr.done - 1;
    if (ntrips <= 0) return;
    std::vector<std::vector<ComplexScalar>> ritz(static_cast<std::size_t>(ntrips));
    auto work = [&](Index t) {
      const MatrixType h = arnoldiBase_.makeHessenbergMatrix(before + 1 + t);
      HessenbergEigenSolver<Scalar> solver;
      solver.compute(h, false);
      std::vector<ComplexScalar>& ev = ritz[static_cast<std::size_t>(t)];
      ev.resize(static_cast<std::size_t>(solver.eigenvalues().size()));
      for (std::size_t i = 0; i < ev.size(); ++i) ev[i] = solver.eigenvalues()[static_cast<Index>(i)];
      std::stable_sort(ev.begin(), ev.end(),
                       [](const ComplexScalar& a, const ComplexScalar& b) { return std::abs(a) > std::abs(b); });
    };
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads > 8) nthreads = 8;
    if (nthreads < 1 || ntrips < 4) nthreads = 1;
    if (nthreads == 1) {
      for (Index t = 0; t < ntrips; ++t) work(t);
    } else {
      std::vector<std::thread> pool;
      for (unsigned w = 0; w < nthreads; ++w)
        pool.emplace_back([&, w]() {
          for (Index t = static_cast<Index>(w); t < ntrips; t += static_cast<Index>(nthreads)) work(t);
        });
      for (auto& th : pool) th.join();
    }
    for (auto& kv : pr.pos) {
      std::vector<ComplexScalar> vals;
      for (Index t = 0; t < ntrips; ++t) {
        const std::vector<ComplexScalar>& ev = ritz[static_cast<std::size_t>(t)];
        Index i = detail::wrap_index(kv.first, static_cast<Index>(ev.size()));
        if (i < 0) continue;
        vals.push_back(ev[static_cast<std::size_t>(i)]);
      }
      if (vals.empty()) continue;
      std::vector<ComplexScalar>& dst = convergenceLog_[kv.first];
      dst.insert(dst.begin() + static_cast<std::ptrdiff_t>(std::min(kv.second, dst.size())), vals.begin(), vals.end());
    }
  }

  /// Hessenberg eigenproblem of the current step, sorted by descending |lambda| (arnoldi.hpp:805-822)
  void solveHessenberg_(bool vectors) {
    hessenbergMatrix_ = arnoldiBase_.makeHessenbergMatrix();
    if (hessenbergMatrix_.rows() == 0) {
      eigenvalues_.resize(0);
      eigenvectors_h_.resize(0, 0);
      return;
    }
    des_.compute(hessenbergMatrix_, vectors);
    eigenvalues_ = des_.eigenvalues();
    auto order = compute_sorted_indices(
        eigenvalues_.data(), eigenvalues_.data() + eigenvalues_.size(),
        [](const ComplexScalar& a, const ComplexScalar& b) -> bool { return std::abs(a) > std::abs(b); });
    EigenEx::cwiseShuffle(eigenvalues_, order);
    if (vectors) {
      eigenvectors_h_ = des_.eigenvectors();
      EigenEx::rowwiseShuffle(eigenvectors_h_, order);
    }
  }

  /// argsort (arnoldi.hpp:893-924)
  template <class ConstRandomIter>
  std::vector<std::size_t> compute_sorted_indices(
      const ConstRandomIter& begin, const ConstRandomIter& end,
      const std::function<bool(const typename std::iterator_traits<ConstRandomIter>::value_type&,
                               const typename std::iterator_traits<ConstRandomIter>::value_type&)>& pred) {
    std::vector<std::size_t> indices(static_cast<std::size_t>(end - begin));
    for (std::size_t i = 0; i < indices.size(); ++i) indices[i] = i;
    std::stable_sort(indices.begin(), indices.end(),
                     [&](std::size_t a, std::size_t b) { return pred(*(begin + a), *(begin + b)); });
    return indices;
  }

  /// one driver trip: the watched Ritz values join their histories
  void recordTrip_() {
    detail::record_trip(convergenceLog_, indicesForConvergence_, eigenvalues_, static_cast<Index>(eigenvalues_.size()));
  }

  /// stop rule (arnoldi.hpp:969-996): all watched Ritz values moved by at most tolerance * |first - last Ritz value|
  bool watchedValuesSettled_() {
    resolvePending_();
    const Index n = static_cast<Index>(eigenvalues_.size());
    if (n < 2) return false;
    const RealScalar spread = std::abs(eigenvalues_[0] - eigenvalues_[n - 1]);
    return detail::histories_settled(convergenceLog_, indicesForConvergence_, spread, tolerance_);
  }

 public:
  /// number of error / warning lines in the log
  Index hasERROR() const { return detail::count_tagged(log_, headERROR()); }
  Index hasWARN() const { return detail::count_tagged(log_, headWARN()); }
};

}  // namespace EigenEx
}  // namespace cmpt

#endif
