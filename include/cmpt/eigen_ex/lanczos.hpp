// lanczos.hpp — drop-in for versmc/cmpt-eigenex include/cmpt/eigen_ex/lanczos.hpp on B200.
//
// Same namespace (cmpt::EigenEx), class names, setters/getters, log strings and stop logic as the
// reference (file:line cited at each member).  What changed is where the work happens: the Krylov basis,
// the work vectors and (optionally) the operator live in HBM, and every Lanczos step runs as hand-written
// sm_100a kernels behind the C-ABI of include/cmpt_b200.h.  The m x m tridiagonal Ritz problem is solved
// on the host (detail/tridiag_eigen.hpp), as in the reference.
//
// Additive API (not in the reference): setMatrixMultiplication(DeviceOperator), updateLanczosSteps(n),
// ritzResiduals(), deviceBytes().
//
// Differences a user can observe:
//  * full reorthogonalisation (interval 1) is classical Gram-Schmidt applied twice (CGS2) instead of one
//    modified Gram-Schmidt sweep (lanczos.hpp:411-426): alpha/beta agree to ~1e-14;
//  * lanczosvectors() / eigenvectors() are copied from the device on demand;
//  * Scalar must be double or std::complex<double>.
#ifndef CMPT_EIGEN_EX_LANCZOS_HPP_
#define CMPT_EIGEN_EX_LANCZOS_HPP_

#include <algorithm>
#include <cmath>
#include <complex>
#include <functional>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "detail/convergence.hpp"
#include "detail/krylov_device.hpp"
#include "detail/tridiag_eigen.hpp"
#include "device.hpp"
#include "random.hpp"

namespace cmpt {
namespace EigenEx {

/// Default convergence tolerance / breakdown threshold (the reference's values, lanczos.hpp:62-83): 1e-4 in single
/// precision, 1e-12 otherwise.
template <class Scalar>
class DefaultTolerance {
 public:
  static constexpr Scalar value() {
    return std::is_same<Scalar, float>::value ? static_cast<Scalar>(1.0e-4f) : static_cast<Scalar>(1.0e-12);
  }
};

/// Stand-in for Eigen::SelfAdjointEigenSolver restricted to computeFromTridiagonal (lanczos.hpp:637).
template <class Scalar>
class TridiagonalEigenSolver {
 public:
  using RealScalar = typename RealOf<Scalar>::type;
  using RealVectorType = Vector<RealScalar>;
  using MatrixType = Matrix<Scalar>;  // the reference types the eigenvectors as Matrix<Scalar> (lanczos.hpp:479)

  /// eigen-decomposition of the tridiagonal matrix (diag, subdiag); only subdiag[0..n-2] is read
  template <class V1, class V2>
  TridiagonalEigenSolver& computeFromTridiagonal(const V1& diag, const V2& subdiag, bool computeVectors = true) {
    return computeRaw(diag.data(), subdiag.data(), static_cast<int>(diag.size()), computeVectors);
  }
  TridiagonalEigenSolver& computeRaw(const RealScalar* diag, const RealScalar* subdiag, int n, bool computeVectors) {
    std::vector<RealScalar> w, z;
    bool ok;
    if (computeVectors)
      ok = detail::tridiagonal_eigensystem<RealScalar>(diag, subdiag, n, w, z);
    else
      ok = detail::tridiagonal_eigenvalues<RealScalar>(diag, subdiag, n, w);
    converged_ = ok;
    eivals_.resize(n);
    for (int i = 0; i < n; ++i) eivals_[i] = w[i];
    if (computeVectors) {
      eivecs_.resize(n, n);
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) eivecs_(i, j) = Scalar(z[static_cast<std::size_t>(j) * n + i]);
    } else {
      eivecs_.resize(0, 0);
    }
    return *this;
  }
  const RealVectorType& eigenvalues() const { return eivals_; }
  const MatrixType& eigenvectors() const { return eivecs_; }
  bool converged() const { return converged_; }

 private:
  RealVectorType eivals_;
  MatrixType eivecs_;
  bool converged_ = true;
};

/// This class generates the basis of the Krylov subspace (lanczos.hpp:104-461):
/// Lanczos vectors (device resident), alpha (diagonal), beta (sub-diagonal).
template <class Scalar_>
class LanczosBase {
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

  /// make random normalized vector (lanczos.hpp:124-135)
  template <class URBG>
  static VectorType makeRandomVector(URBG& g, Index size) {
    VectorDistribution vdist(ScalarDistribution(), size, true);
    return vdist(g);
  }

 protected:
  Index reserveSize_;
  std::vector<VectorType> orthogonalizingVectors_;
  MatMulFunction matrixMultiplication_;
  DeviceOperator<Scalar> deviceOperator_;  // additive: operator resident in HBM
  RealScalar eigenvalueShift_;
  Index matrixHeight_;
  Index reorthogonalizeInterval_;
  VectorType initialVector_;
  std::uint64_t initialVersion_ = 1;  // bumped by every setInitialVector: the device keeps the last uploaded version
  RealScalar threshold_;

 public:
  // accessors for parameters of settings of lanczos computing (lanczos.hpp:160-226)
  Index reserveSize() const { return reserveSize_; }
  LanczosBase& setReserveSize(Index resSize) {
    reserveSize_ = resSize;
    return *this;
  }
  const std::vector<VectorType>& orthogonalizingVectors() const { return orthogonalizingVectors_; }
  std::vector<VectorType>& refOrthogonalizingVectors() { return orthogonalizingVectors_; }
  LanczosBase& setOrthogonalizingVectors(const std::vector<VectorType>& orthoVec) {
    orthogonalizingVectors_ = orthoVec;
    return *this;
  }
  LanczosBase& setOrthogonalizingVectors(std::vector<VectorType>&& orthoVec) {
    orthogonalizingVectors_.swap(orthoVec);
    return *this;
  }
  const MatMulFunction& matrixMultiplication() const { return matrixMultiplication_; }
  LanczosBase& setMatrixMultiplication(const MatMulFunction& matmul, Index height) {
    matrixMultiplication_ = matmul;
    deviceOperator_ = DeviceOperator<Scalar>();
    matrixHeight_ = height;
    return *this;
  }
  LanczosBase& setMatrixMultiplication(MatMulFunction&& matmul, Index height) {
    std::swap(matrixMultiplication_, matmul);
    deviceOperator_ = DeviceOperator<Scalar>();
    matrixHeight_ = height;
    return *this;
  }
  /// additive overload: operator resident in HBM (CSR/SELL, dense, matrix-free); no per-step host traffic
  LanczosBase& setMatrixMultiplication(const DeviceOperator<Scalar>& op) {
    deviceOperator_ = op;
    DeviceOperator<Scalar> held = op;
    matrixMultiplication_ = [held](const Scalar* in, Scalar* out) { held.apply(in, out); };
    matrixHeight_ = op.height();
    return *this;
  }
  const DeviceOperator<Scalar>& deviceOperator() const { return deviceOperator_; }
  Index matrixHeight() const { return matrixHeight_; }
  /// additive: length of the vectors this process holds — matrixHeight() on one GPU, the number of rows of this
  /// rank's shard in a row-partitioned run (start/deflation/basis/Ritz vectors are then local slabs)
  Index localHeight() const { return deviceOperator_ ? deviceOperator_.rows() : matrixHeight_; }
  Index localRowBegin() const { return deviceOperator_ ? deviceOperator_.rowBegin() : 0; }
  RealScalar eigenvalueShift() const { return eigenvalueShift_; }
  LanczosBase& setEigenvalueShift(RealScalar eishift) {
    eigenvalueShift_ = eishift;
    return *this;
  }
  bool reorthogonalizeInterval() const { return reorthogonalizeInterval_; }  // sic: bool (lanczos.hpp:194)
  LanczosBase& setReorthogonalizeInterval(Index reorthoInterval) {
    reorthogonalizeInterval_ = reorthoInterval;
    return *this;
  }
  const VectorType& initialVector() const { return initialVector_; }
  LanczosBase& setInitialVector(const VectorType& inivec) {
    initialVector_ = inivec;
    ++initialVersion_;
    return *this;
  }
  LanczosBase& setInitialVector(VectorType&& inivec) {
    initialVector_ = std::move(inivec);
    ++initialVersion_;
    return *this;
  }
  /// additive: copy n scalars straight into the (pinned, reused) start-vector storage
  LanczosBase& setInitialVector(const Scalar* data, Index n) {
    detail::assign_upload(initialVector_, data, n);
    ++initialVersion_;
    return *this;
  }
  /// initial vector of size matrixHeight with random contents, fixed seed (lanczos.hpp:214-218)
  LanczosBase& setInitialVector() {
    std::mt19937 rengine;
    VectorType full = makeRandomVector(rengine, matrixHeight_);
    if (localHeight() == matrixHeight_) {
      setInitialVector(std::move(full));
    } else {  // row-partitioned: this rank's slab of the same global vector
      VectorType slab(localHeight());
      for (Index i = 0; i < localHeight(); ++i) slab[i] = full[localRowBegin() + i];
      setInitialVector(std::move(slab));
    }
    return *this;
  }
  RealScalar threshold() const { return threshold_; }
  LanczosBase& setThreshold(RealScalar thre) {
    threshold_ = thre;
    return *this;
  }

 protected:
  // computed data: the vectors live on the device; alpha/beta are mirrored on the host
  Index iterations_;
  Index nvectors_;
  Index arrow_;  // size of the arrowhead block after a thick restart (0: tridiagonal)
  mutable std::vector<VectorType> lanczosvectors_;  // host cache of the device basis, filled on demand
  std::vector<RealScalar> alpha_;
  std::vector<RealScalar> beta_;
  detail::KrylovDevice<Scalar> dev_;

 public:
  Index iterations() const { return iterations_; }
  /// number of Lanczos vectors (== lanczosvectors().size(), without copying them to the host)
  Index lanczosvectorsSize() const { return nvectors_; }
  /// Lanczos vectors, copied from the device on demand (lanczos.hpp:246)
  const std::vector<VectorType>& lanczosvectors() const {
    if (static_cast<Index>(lanczosvectors_.size()) > nvectors_) lanczosvectors_.resize(nvectors_);
    while (static_cast<Index>(lanczosvectors_.size()) < nvectors_) {
      VectorType v(localHeight());
      detail::check(cmb_krylov_get_col(dev_.handle(), static_cast<std::int64_t>(lanczosvectors_.size()), v.data()),
                    "cmb_krylov_get_col");
      lanczosvectors_.push_back(std::move(v));
    }
    return lanczosvectors_;
  }
  const std::vector<RealScalar>& alpha() const { return alpha_; }
  const std::vector<RealScalar>& beta() const { return beta_; }
  /// additive: > 0 after thickRestart(k): the projected matrix is then diag(alpha) with T(i,k) = T(k,i) = beta[i] for
  /// i < k (arrowhead) and T(i,i+1) = T(i+1,i) = beta[i] for i >= k (tridiagonal tail); 0 = plain tridiagonal
  Index arrowSize() const { return arrow_; }
  cmb_krylov* deviceState() const { return dev_.handle(); }
  /// algorithmic bytes moved by the Krylov steps so far
  double deviceBytes() const { return dev_.ready() ? cmb_krylov_bytes(dev_.handle()) : 0.0; }

 public:
  LanczosBase() : iterations_(0), nvectors_(0), arrow_(0) { setAllSettingsDefault(); }

  /// default settings; does NOT clear computed data (lanczos.hpp:260-271)
  LanczosBase& setAllSettingsDefault() {
    setReserveSize(128);
    setOrthogonalizingVectors(std::vector<VectorType>());
    setMatrixMultiplication([](const Scalar*, Scalar*) {}, 0);
    setEigenvalueShift(0.0);
    setReorthogonalizeInterval(1);
    setInitialVector();
    setThreshold(DefaultTolerance<RealScalar>::value());
    return *this;
  }

  /// clears lanczosvectors, alpha, beta; keeps settings (lanczos.hpp:277-283)
  void clearLanczosSteps() {
    iterations_ = 0;
    nvectors_ = 0;
    arrow_ = 0;
    lanczosvectors_.clear();
    alpha_.clear();
    beta_.clear();
    if (dev_.ready()) detail::check(cmb_krylov_clear(dev_.handle()), "cmb_krylov_clear");
  }

  /// clear all data, and set all settings default (lanczos.hpp:289-292)
  void clear() {
    clearLanczosSteps();
    setAllSettingsDefault();
    dev_.release();
  }

  /// judge whether the dimension of the Krylov subspace is utmost (lanczos.hpp:331-347)
  bool lanczosStepIsUtmost() const {
    if (nvectors_ == matrixHeight_) return true;
    if (beta_.size() > 0) return beta_.back() <= threshold_;
    return false;
  }

  /// one Lanczos step (lanczos.hpp:371-457); false when no step could be done
  bool updateLanczosSteps() { return updateLanczosSteps(1) == 1; }

  /// additive: up to `count` consecutive calls of updateLanczosSteps() enqueued on the device with a single
  /// host synchronisation at the end.  Returns how many of them returned true.
  Index updateLanczosSteps(Index count) {
    if (matrixHeight_ <= 0) return 0;
    if (!matrixMultiplication_ && !deviceOperator_) return 0;
    if (count <= 0) return 0;
    prepareDevice_();
    Index done = 0;
    if (nvectors_ == 0) {
      if (!setInitialLanczosvector_()) return 0;
    }
    std::vector<double> a(static_cast<std::size_t>(count) + 1), b(static_cast<std::size_t>(count) + 1);
    std::int64_t steps = 0;
    int status = 0;
    const bool first = (nvectors_ == 0);
    int rc = cmb_lanczos_run(dev_.handle(), dev_.op(), eigenvalueShift_, reorthogonalizeInterval_, threshold_, count,
                             a.data(), b.data(), &steps, &status);
    dev_.rethrowCallbackError();
    detail::check(rc, "cmb_lanczos_run");
    done = static_cast<Index>(steps);
    const Index nbeta = first ? (done > 0 ? done - 1 : 0) : done;
    for (Index i = 0; i < done; ++i) alpha_.push_back(a[static_cast<std::size_t>(i)]);
    for (Index i = 0; i < nbeta; ++i) beta_.push_back(b[static_cast<std::size_t>(i)]);
    if (status & CMB_STEP_BREAKDOWN) beta_.push_back(b[static_cast<std::size_t>(nbeta)]);  // kept (lanczos.hpp:433-436)
    nvectors_ += done;
    iterations_ += nbeta;
    return done;
  }

  /// additive: thick restart.  With Lanczos vectors u_0..u_m, keeps the k Ritz pairs
  /// (theta[i], V_m coef(:,i)) of the projected matrix of u_0..u_{m-1} — coef is m x k column-major — plus u_m:
  /// the basis becomes [y_1..y_k, u_m], alpha = [theta, alpha_m], beta[i] = couplings[i] (= last-row component of
  /// coef(:,i) times the last beta, supplied by the caller).  Needs reorthogonalizeInterval() == 1.
  void thickRestart(const std::vector<RealScalar>& coef, const std::vector<RealScalar>& theta,
                    const std::vector<RealScalar>& couplings) {
    const Index m = nvectors_ - 1, k = static_cast<Index>(theta.size());
    if (m < 1 || k < 1 || k > m || static_cast<Index>(coef.size()) != m * k ||
        static_cast<Index>(couplings.size()) != k)
      throw LanczosException("thickRestart: bad shapes");
    if (reorthogonalizeInterval_ != 1) throw LanczosException("thickRestart needs full reorthogonalisation");
    std::vector<double> c(coef.begin(), coef.end());
    detail::check(cmb_lanczos_thick_restart(dev_.handle(), c.data(), m, m, k), "cmb_lanczos_thick_restart");
    const RealScalar alpha_m = alpha_[static_cast<std::size_t>(m)];
    alpha_.assign(theta.begin(), theta.end());
    alpha_.push_back(alpha_m);
    beta_.assign(couplings.begin(), couplings.end());
    nvectors_ = k + 1;
    arrow_ = k;
    lanczosvectors_.clear();
  }

 protected:
  void prepareDevice_() {
    Index reserve = reserveSize_;
    dev_.prepare(deviceOperator_, matrixMultiplication_, matrixHeight_, reserve);
  }

  /// lanczos.hpp:299-323; returns false when the start vector has (numerically) no component left
  bool setInitialLanczosvector_() {
    if (matrixHeight_ < 0) throw LanczosException("matrixHeight_ < 0");
    if (localHeight() != static_cast<Index>(initialVector_.size())) setInitialVector();
    dev_.setDeflation(orthogonalizingVectors_, localHeight());
    int status = 0;
    dev_.start(initialVector_, initialVersion_, threshold_, &status);
    return status == CMB_STEP_OK;
  }
};

/// eigen solver with Lanczos method (lanczos.hpp:468-927)
template <class Scalar_>
class LanczosEigenSolver {
 public:
  using Index = EigenEx::Index;
  using Scalar = Scalar_;
  using RealScalar = typename RealOf<Scalar>::type;
  using VectorType = Vector<Scalar>;
  using RealVectorType = Vector<RealScalar>;
  using MatrixType = Matrix<Scalar>;
  using RealMatrixType = Matrix<Scalar>;  // sic (lanczos.hpp:479)
  using ScalarDistribution = typename NormalDistributionGen<Scalar>::Type;
  using VectorDistribution = typename EigenEx::VectorDistribution<ScalarDistribution>;
  using MatMulFunction = std::function<void(const Scalar*, Scalar*)>;

  // header of log generation (lanczos.hpp:486-489)
  static std::string headERROR() { return std::string("ERROR     "); }
  static std::string headWARN() { return std::string("WARN      "); }
  static std::string headINFO() { return std::string("INFO      "); }
  static std::string headDEBUG() { return std::string("DEBUG     "); }

  // Index value for special case of minIterations, maxIterations
  static constexpr Index unlimited = -1;

  template <class URBG>
  static VectorType makeRandomVector(URBG& g, Index size) {
    return LanczosBase<Scalar>::makeRandomVector(g, size);
  }

 protected:
  Index minIterations_;
  Index maxIterations_;
  RealScalar tolerance_;
  std::vector<Index> indicesForConvergence_;
  Index maxEigenvalues_;
  bool computeEigenvectorsOn_;

 public:
  Index minIterations() const { return minIterations_; }
  LanczosEigenSolver& setMinIterations(Index miniter) {
    minIterations_ = miniter;
    return *this;
  }
  Index maxIterations() const { return maxIterations_; }
  LanczosEigenSolver& setMaxIterations(Index maxiter) {
    maxIterations_ = maxiter;
    return *this;
  }
  RealScalar tolerance() const { return tolerance_; }
  LanczosEigenSolver& setTolerance(RealScalar toler) {
    tolerance_ = toler;
    return *this;
  }
  const std::vector<Index>& indicesForConvergence() const { return indicesForConvergence_; }
  LanczosEigenSolver& setIndicesForConvergence(const std::vector<Index>& iCovs) {
    resolvePending_();  // pending log entries belong to the old index set
    indicesForConvergence_ = iCovs;
    return *this;
  }
  Index maxEigenvalues() const { return maxEigenvalues_; }
  LanczosEigenSolver& setMaxEigenvalues(Index maxeivals) {
    maxEigenvalues_ = maxeivals;
    return *this;
  }
  Index computeEigenvectorsOn() const { return computeEigenvectorsOn_; }  // sic: Index (lanczos.hpp:551)
  LanczosEigenSolver& setComputeEigenvectorsOn(bool cEivecOn) {
    computeEigenvectorsOn_ = cEivecOn;
    return *this;
  }

 protected:
  LanczosBase<Scalar> lanczosBase_;

 public:  // transparent accessors for lanczosBase (lanczos.hpp:562-628)
  const LanczosBase<Scalar>& lanczosBase() const { return lanczosBase_; }
  Index reserveSize() const { return lanczosBase_.reserveSize(); }
  LanczosEigenSolver& setReserveSize(Index resSize) {
    lanczosBase_.setReserveSize(resSize);
    return *this;
  }
  const std::vector<VectorType>& orthogonalizingVectors() const { return lanczosBase_.orthogonalizingVectors(); }
  std::vector<VectorType>& refOrthogonalizingVectors() { return lanczosBase_.refOrthogonalizingVectors(); }
  LanczosEigenSolver& setOrthogonalizingVectors(const std::vector<VectorType>& orthoVec) {
    lanczosBase_.setOrthogonalizingVectors(orthoVec);
    return *this;
  }
  LanczosEigenSolver& setOrthogonalizingVectors(std::vector<VectorType>&& orthoVec) {
    lanczosBase_.setOrthogonalizingVectors(std::move(orthoVec));
    return *this;
  }
  const MatMulFunction& matrixMultiplication() const { return lanczosBase_.matrixMultiplication(); }
  LanczosEigenSolver& setMatrixMultiplication(const MatMulFunction& matmul, Index height) {
    lanczosBase_.setMatrixMultiplication(matmul, height);
    return *this;
  }
  LanczosEigenSolver& setMatrixMultiplication(MatMulFunction&& matmul, Index height) {
    lanczosBase_.setMatrixMultiplication(std::move(matmul), height);
    return *this;
  }
  /// additive overload: operator resident in HBM
  LanczosEigenSolver& setMatrixMultiplication(const DeviceOperator<Scalar>& op) {
    lanczosBase_.setMatrixMultiplication(op);
    return *this;
  }
  Index matrixHeight() const { return lanczosBase_.matrixHeight(); }
  Index localHeight() const { return lanczosBase_.localHeight(); }
  RealScalar eigenvalueShift() const { return lanczosBase_.eigenvalueShift(); }
  LanczosEigenSolver& setEigenvalueShift(RealScalar eishift) {
    lanczosBase_.setEigenvalueShift(eishift);
    return *this;
  }
  bool reorthogonalizeInterval() const { return lanczosBase_.reorthogonalizeInterval(); }
  LanczosEigenSolver& setReorthogonalizeInterval(Index reorthoInterval) {
    lanczosBase_.setReorthogonalizeInterval(reorthoInterval);
    return *this;
  }
  const VectorType& initialVector() const { return lanczosBase_.initialVector(); }
  LanczosEigenSolver& setInitialVector(const VectorType& inivec) {
    lanczosBase_.setInitialVector(inivec);
    return *this;
  }
  LanczosEigenSolver& setInitialVector(VectorType&& inivec) {
    lanczosBase_.setInitialVector(std::move(inivec));
    return *this;
  }
  LanczosEigenSolver& setInitialVector(const Scalar* data, Index n) {
    lanczosBase_.setInitialVector(data, n);
    return *this;
  }
  LanczosEigenSolver& setInitialVector() {
    lanczosBase_.setInitialVector();
    return *this;
  }
  RealScalar threshold() const { return lanczosBase_.threshold(); }
  LanczosEigenSolver& setThreshold(RealScalar thre) {
    lanczosBase_.setThreshold(thre);
    return *this;
  }
  Index iterations() const { return lanczosBase_.iterations(); }
  const std::vector<VectorType>& lanczosvectors() const { return lanczosBase_.lanczosvectors(); }
  const std::vector<RealScalar>& alpha() const { return lanczosBase_.alpha(); }
  const std::vector<RealScalar>& beta() const { return lanczosBase_.beta(); }

 protected:
  RealVectorType eigenvalues_;
  MatrixType eigenvectors_;
  std::vector<std::string> log_;
  TridiagonalEigenSolver<Scalar> es_tri_;
  mutable std::map<Index, std::vector<RealScalar>> convergenceLog_;
  // trips skipped by batched stepping whose convergence-log entries have not been computed yet (resolved on the
  // first access to the log): {first state, number of calls, insert position per tracked index}
  struct PendingReplay {
    Index before, done;
    std::map<Index, std::size_t> pos;
  };
  mutable std::vector<PendingReplay> pendingReplay_;

 public:
  const RealVectorType& eigenvalues() const { return eigenvalues_; }
  const MatrixType& eigenvectors() const { return eigenvectors_; }
  const std::vector<std::string>& log() const { return log_; }
  const TridiagonalEigenSolver<Scalar>& es_tri() const { return es_tri_; }
  const std::map<Index, std::vector<RealScalar>>& convergenceLog() const {
    resolvePending_();
    return convergenceLog_;
  }

  /// additive: Ritz residual bounds ||A x_i - theta_i x_i|| = |beta_next * S(last, i)| of the returned eigenpairs,
  /// beta_next = ||A u_k - alpha_k u_k - beta_{k-1} u_{k-1}|| (one fused device pass; the kept beta after a breakdown)
  RealVectorType ritzResiduals() const {
    const Index k = static_cast<Index>(alpha().size());
    RealVectorType r(eigenvalues_.size());
    RealScalar bl = RealScalar(0);
    if (k > 0 && static_cast<Index>(beta().size()) >= k) {
      bl = beta()[k - 1];
    } else if (k > 0 && lanczosBase_.deviceState()) {
      double b = 0.0;
      detail::check(cmb_lanczos_residual_norm(lanczosBase_.deviceState(), &b), "cmb_lanczos_residual_norm");
      bl = b;
    }
    for (Index i = 0; i < static_cast<Index>(eigenvalues_.size()); ++i)
      r[i] = (k > 0 && es_tri_.eigenvectors().rows() == k) ? std::abs(bl * es_tri_.eigenvectors()(k - 1, i)) : RealScalar(0);
    return r;
  }

 public:
  LanczosEigenSolver() { setAllSettingsDefault(); }

  /// default settings; does NOT clear computed data (lanczos.hpp:657-668)
  LanczosEigenSolver& setAllSettingsDefault() {
    setMinIterations(1);
    setMaxIterations(unlimited);
    setTolerance(DefaultTolerance<RealScalar>::value());
    setIndicesForConvergence(std::vector<Index>{0});
    setMaxEigenvalues(unlimited);
    setComputeEigenvectorsOn(true);
    lanczosBase_.setAllSettingsDefault();
    return *this;
  }

  /// clear computed data, keep settings (lanczos.hpp:675-682)
  LanczosEigenSolver& clearComputedData() {
    lanczosBase_.clearLanczosSteps();
    eigenvalues_.resize(0);
    eigenvectors_.resize(0, 0);
    log_.clear();
    convergenceLog_.clear();
    pendingReplay_.clear();
    return *this;
  }

  /// clear computed data and set all settings default (lanczos.hpp:689-693)
  LanczosEigenSolver& clear() {
    clearComputedData();
    setAllSettingsDefault();
    return *this;
  }

  /// continue from the current state with possibly changed settings (lanczos.hpp:701-712)
  Index continueToCompute() {
    log_.push_back(headINFO() + "EigenSolver<ScalarType>::continueToCompute(...) was called");
    if (lanczosBase_.lanczosvectorsSize() == 0) return compute();
    Index ret = mainCalculation_();
    log_.push_back(headINFO() + "EigenSolver<ScalarType>::compute(...) finish computing");
    return ret;
  }

  /// compute matrix diagonalization (lanczos.hpp:717-736)
  Index compute() {
    log_.push_back(headINFO() + "EigenSolver<ScalarType>::compute(...) was called");
    clearComputedData();
    if (static_cast<Index>(initialVector().size()) != localHeight()) {
      log_.push_back(headINFO() + "in compute(), initial_vector is empty or invalid, then set at random");
      setInitialVector();
    }
    Index ret = mainCalculation_();
    log_.push_back(headINFO() + "EigenSolver<ScalarType>::compute(...) finish computing");
    return ret;
  }

  /// lanczos.hpp:740-823.  Steps that no stop rule can interrupt (iterations < minIterations) are enqueued
  /// on the device as one batch; their per-trip bookkeeping (convergence log) is replayed afterwards from
  /// alpha/beta, so results and logs are those of the trip-by-trip loop.
  Index mainCalculation_() {
    solveTridiagonal_(0, false);
    bool set_initialvector_is_fail = false;
    while (true) {
      recordTrip_();
      {
        if (set_initialvector_is_fail) {
          log_.push_back(headINFO() + "initial lanczosvector generation fail");
          break;
        }
        if (lanczosBase_.lanczosStepIsUtmost()) {
          log_.push_back(headINFO() + "lanczos steps finished with threshold");
          log_.push_back(headINFO() + "lanczos steps achieved full of Krylov subspace");
          break;
        }
        if (lanczosBase_.iterations() >= minIterations()) {
          if (lanczosBase_.iterations() == maxIterations()) {
            log_.push_back(headWARN() + "lanczos steps achieved maxIterations");
            break;
          }
          if (watchedValuesSettled_()) {
            log_.push_back(headINFO() + "lanczos steps converged with tolerance");
            break;
          }
        }
      }
      // number of calls no test above can interrupt
      Index batch = 1;
      if (lanczosBase_.iterations() < minIterations()) {
        batch = minIterations() - lanczosBase_.iterations() + (lanczosBase_.lanczosvectorsSize() == 0 ? 1 : 0);
        const Index room = matrixHeight() - lanczosBase_.lanczosvectorsSize();
        if (batch > room) batch = room;
        if (batch < 1) batch = 1;
      }
      const Index before = static_cast<Index>(lanczosBase_.alpha().size());
      const Index done = lanczosBase_.updateLanczosSteps(batch);
      if (lanczosBase_.lanczosvectorsSize() == 0) set_initialvector_is_fail = true;
      // replay the trips of the intermediate states (their tridiagonal problems are independent of each other)
      deferReplay_(before, done);
      solveTridiagonal_(static_cast<Index>(lanczosBase_.alpha().size()), false);
    }

    // eigenvectors of the tridiagonal matrix are needed once, here (the reference recomputes them every trip)
    solveTridiagonal_(static_cast<Index>(lanczosBase_.alpha().size()), true);

    // back eigen value to original one (lanczos.hpp:786-795)
    Index eivalsize = es_tri_.eigenvalues().size();
    if (maxEigenvalues_ != unlimited) {
      if (maxEigenvalues_ < eivalsize) eivalsize = maxEigenvalues_;
    }
    eigenvalues_.resize(eivalsize);
    for (Index k = 0; k < eivalsize; ++k) eigenvalues_[k] = es_tri_.eigenvalues()[k] - lanczosBase_.eigenvalueShift();

    // Ritz vectors (lanczos.hpp:798-817): X = V S, normalised, phase-fixed — assembled on the device
    if (computeEigenvectorsOn_) {
      detail::resize_result(eigenvectors_, localHeight(), eivalsize);
      if (eivalsize > 0) {
        const Index nm = es_tri_.eigenvectors().rows();
        std::vector<Scalar> coef(static_cast<std::size_t>(nm) * eivalsize);
        for (Index kk = 0; kk < eivalsize; ++kk)
          for (Index m = 0; m < nm; ++m) coef[static_cast<std::size_t>(kk) * nm + m] = es_tri_.eigenvectors()(m, kk);
        detail::check(cmb_krylov_ritz_vectors(lanczosBase_.deviceState(), detail::DTypeOf<Scalar>::value, coef.data(), nm,
                                              nm, eivalsize, eigenvectors_.data(), localHeight()),
                      "cmb_krylov_ritz_vectors");
      }
    } else {
      eigenvectors_.resize(0, 0);
    }
    return 0;
  }

 protected:
  void solveTridiagonal_(Index k, bool vectors) {
    es_tri_.computeRaw(lanczosBase_.alpha().data(), lanczosBase_.beta().data(), static_cast<int>(k), vectors);
  }

  /// The trips a batched updateLanczosSteps(done) skipped (states with before+1 .. before+done-1 Lanczos vectors)
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
  /// Ritz values of each skipped T_j on a few host threads, inserted in trip order at the recorded positions —
  /// exactly the entries recordTrip_ would have appended trip by trip.
  void replayTrips_(const PendingReplay& pr) const {
    const Index before = pr.before, ntrips = pr.done - 1;
    if (ntrips <= 0) return;
    std::vector<std::vector<RealScalar>> ritz(static_cast<std::size_t>(ntrips));
    const RealScalar* a = lanczosBase_.alpha().data();
    const RealScalar* b = lanczosBase_.beta().data();
    auto work = [&](Index t) {
      detail::tridiagonal_eigenvalues<RealScalar>(a, b, static_cast<int>(before + 1 + t), ritz[static_cast<std::size_t>(t)]);
    };
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads > 8) nthreads = 8;
    if (nthreads < 1 || ntrips < 8) nthreads = 1;
    if (nthreads == 1) {
      for (Index t = 0; t < ntrips; ++t) work(t);
    } else {
      // interleave so that every thread gets small and large problems
      std::vector<std::thread> pool;
      for (unsigned w = 0; w < nthreads; ++w)
        pool.emplace_back([&, w]() {
          for (Index t = static_cast<Index>(w); t < ntrips; t += static_cast<Index>(nthreads)) work(t);
        });
      for (auto& th : pool) th.join();
    }
    for (auto& kv : pr.pos) {
      std::vector<RealScalar> vals;
      for (Index t = 0; t < ntrips; ++t) {
        const std::vector<RealScalar>& ev = ritz[static_cast<std::size_t>(t)];
        Index i = detail::wrap_index(kv.first, static_cast<Index>(ev.size()));
        if (i < 0) continue;
        vals.push_back(ev[static_cast<std::size_t>(i)]);
      }
      if (vals.empty()) continue;
      std::vector<RealScalar>& dst = convergenceLog_[kv.first];
      dst.insert(dst.begin() + static_cast<std::ptrdiff_t>(std::min(kv.second, dst.size())), vals.begin(), vals.end());
    }
  }

  /// one driver trip: the watched Ritz values of the current T_j join their histories
  void recordTrip_() {
    const RealVectorType& ritz = es_tri_.eigenvalues();
    detail::record_trip(convergenceLog_, indicesForConvergence_, ritz, static_cast<Index>(ritz.size()));
  }

  /// stop rule: all watched Ritz values moved by at most tolerance * (spread of the Ritz values) in the last trip
  bool watchedValuesSettled_() {
    resolvePending_();
    const RealVectorType& ritz = es_tri_.eigenvalues();
    const Index n = static_cast<Index>(ritz.size());
    if (n < 2) return false;
    const RealScalar spread = ritz[0] - ritz[n - 1];
    return detail::histories_settled(convergenceLog_, indicesForConvergence_, spread, tolerance_);
  }

 public:
  /// number of error / warning lines in the log (lanczos.hpp:903-922)
  Index hasERROR() const { return detail::count_tagged(log_, headERROR()); }
  Index hasWARN() const { return detail::count_tagged(log_, headWARN()); }
};

/// exp(xA)|ket> with the Lanczos method or a Taylor expansion (lanczos.hpp:1002-1164), A Hermitian.
/// solveWithEigens / solveWithTaylor* are host routines with the reference's signatures; solveWithLanczos runs
/// the Krylov part on the device and evaluates the expansion there too (see below).
template <typename Scalar_>
class LanczosExponentialSolver {
 public:
  using Index = EigenEx::Index;
  using Solver = LanczosEigenSolver<Scalar_>;
  using Scalar = typename Solver::Scalar;
  using RealScalar = typename Solver::RealScalar;
  using VectorType = typename Solver::VectorType;
  using RealVectorType = typename Solver::RealVectorType;
  using MatrixType = typename Solver::MatrixType;
  using RealMatrixType = typename Solver::RealMatrixType;
  using MatMulFunction = typename Solver::MatMulFunction;

  static constexpr Index unlimited = Solver::unlimited;

  /// out = sum_n exp(x E_n) <y_n|in> y_n over the first max_expand eigenpairs, smallest weight first
  /// (lanczos.hpp:1024-1054)
  static void solveWithEigens(const Scalar x, const RealVectorType& eivals, const MatrixType& eivecs, Index max_expand,
                              const VectorType& in, VectorType& out) {
    Index max = max_expand;
    if (static_cast<Index>(eivals.size()) - max_expand < 0) max = eivals.size();
    if (static_cast<Index>(eivecs.cols()) - max_expand < 0) max = eivecs.cols();
    const Index n = in.size();
    out.resize(n);
    for (Index i = 0; i < n; ++i) out[i] = Scalar(0);
    for (Index n_ = 0; n_ < max; ++n_) {
      Index k = n_;
      if (std::real(x) < 0.0) k = max - n_ - 1;
      Scalar inner = Scalar(0);
      for (Index i = 0; i < n; ++i) inner += detail::conj_(eivecs(i, k)) * in[i];
      const Scalar c = std::exp(x * eivals[k]) * inner;
      for (Index i = 0; i < n; ++i) out[i] += c * eivecs(i, k);
    }
  }

  /// es.compute(), then the eigen-expansion with es.initialVector() as the input vector (lanczos.hpp:1061-1075).
  /// The expansion is evaluated in the Krylov space on the device: g = V^H in (one pass over the basis),
  /// c = S_nev exp(x Theta) S_nev^H g on the host (nev = number of Ritz pairs the solver returned), out = V c (one
  /// more pass) — the same sum as solveWithEigens(x, es.eigenvalues(), es.eigenvectors(), ...), without needing the
  /// n x nev Ritz-vector matrix.  As in the reference, nothing is added when the solver returned no eigenvectors.
  static void solveWithLanczos(const Scalar& x, Solver& es, VectorType& out) {
    es.compute();
    const Index n = es.localHeight();
    const Index nev = std::min<Index>(es.eigenvalues().size(), es.eigenvectors().cols());
    const Index nk = es.lanczosBase().lanczosvectorsSize();
    out.resize(n);
    for (Index i = 0; i < n; ++i) out[i] = Scalar(0);
    if (nev <= 0 || nk <= 0) return;
    std::vector<Scalar> g(static_cast<std::size_t>(nk));
    detail::check(cmb_krylov_project(es.lanczosBase().deviceState(), es.initialVector().data(), g.data()),
                  "cmb_krylov_project");
    const auto& S = es.es_tri().eigenvectors();  // nk x nk, columns = eigenvectors of T (real entries)
    std::vector<Scalar> c(static_cast<std::size_t>(nk), Scalar(0));
    for (Index k_ = 0; k_ < nev; ++k_) {
      Index k = k_;
      if (std::real(x) < 0.0) k = nev - k_ - 1;  // smallest weight first, as the reference
      Scalar inner = Scalar(0);
      for (Index m = 0; m < nk; ++m) inner += detail::conj_(S(m, k)) * g[static_cast<std::size_t>(m)];
      const Scalar w = std::exp(x * es.eigenvalues()[k]) * inner;
      for (Index m = 0; m < nk; ++m) c[static_cast<std::size_t>(m)] += w * S(m, k);
    }
    detail::check(cmb_krylov_combine(es.lanczosBase().deviceState(), c.data(), nk, out.data()), "cmb_krylov_combine");
  }

  /// plain Taylor expansion (lanczos.hpp:1085-1130)
  static void solveWithTaylorNoDivision(Scalar x, const MatMulFunction& matmul, Index matrix_height,
                                        RealScalar matrix_radius, const VectorType& in, VectorType& out,
                                        RealScalar error = 1.0e-14, Index max_expansion = unlimited) {
    out = in;  // k == 0
    Scalar c_k = Scalar(1.0);
    RealScalar radius_k = RealScalar(1.0);
    Index k = 1;
    VectorType ket_k(matrix_height), ket_pre(matrix_height);
    c_k *= x / static_cast<double>(k);
    radius_k *= matrix_radius;
    matmul(in.data(), ket_k.data());
    for (Index i = 0; i < matrix_height; ++i) out[i] += c_k * ket_k[i];
    if (max_expansion == 1) return;
    std::swap(ket_pre, ket_k);
    for (k = 2; k != max_expansion; ++k) {
      c_k *= x / static_cast<double>(k);
      radius_k *= matrix_radius;
      matmul(ket_pre.data(), ket_k.data());
      for (Index i = 0; i < matrix_height; ++i) out[i] += c_k * ket_k[i];
      std::swap(ket_k, ket_pre);
      if (std::abs(c_k * radius_k) < error) break;
    }
  }

  /// Taylor expansion with the step split into div = floor(|x| radius + 1) sub-steps (lanczos.hpp:1138-1161).
  /// The reference applies every sub-step to `in` (so it returns exp(xA/div) in); the sub-steps are chained here,
  /// which is what the function documents: out = exp(xA) in.
  static void solveWithTaylorAutoDivision(Scalar x, const MatMulFunction& matmul, Index matrix_height,
                                          RealScalar matrix_radius, const VectorType& in, VectorType& out,
                                          RealScalar error = 1.0e-14, Index max_expansion = unlimited) {
    const RealScalar rad = std::abs(x * matrix_radius);
    const Index div = static_cast<Index>(rad + 1.0);
    VectorType cur = in, next;
    for (Index i = 0; i < div; ++i) {
      const Scalar x_ = static_cast<RealScalar>(1.0 / div) * x;
      solveWithTaylorNoDivision(x_, matmul, matrix_height, matrix_radius, cur, next, error, max_expansion);
      std::swap(cur, next);
    }
    out = cur;
  }
};

}  // namespace EigenEx
}  // namespace cmpt

#endif
