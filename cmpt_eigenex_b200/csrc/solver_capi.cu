// solver_capi.cu — C binding (include/cmpt_b200_solver.h) of the drop-in solver classes.  Host code only:
// it instantiates LanczosEigenSolver / ArnoldiEigenSolver for double and std::complex<double> and forwards.
#include <string.h>

#include <complex>
#include <string>

#include "cmpt/eigen_ex/arnoldi.hpp"
#include "cmpt/eigen_ex/lanczos.hpp"
#include "cmpt/eigen_ex/arnoldi_restart.hpp"
#include "cmpt/eigen_ex/thick_restart.hpp"
#include "cmpt/eigen_ex/detail/symmetric_eigen.hpp"
#include "cmpt_b200_solver.h"
#include "common.cuh"

using namespace cmpt::EigenEx;
using cmb::set_error;

struct cmbs_solver {
  virtual ~cmbs_solver() {}
  int kind = 0;
  int dtype = 0;
  virtual void set_operator(cmb_op* op) = 0;
  virtual void set_callback(int64_t n, cmb_matmul_fn fn, void* user) = 0;
  virtual bool set_int(const std::string& k, int64_t v) = 0;
  virtual bool set_real(const std::string& k, double v) = 0;
  virtual bool set_complex(const std::string& k, double re, double im) = 0;
  virtual bool get_int(const std::string& k, int64_t* v) = 0;
  virtual bool get_real(const std::string& k, double* v) = 0;
  virtual void set_indices(const int64_t* idx, int64_t n) = 0;
  virtual void set_initial(const void* v, int64_t n) = 0;
  virtual void set_ortho(int64_t nvec, const void* vecs, int64_t ld) = 0;
  virtual void compute() = 0;
  virtual void continue_compute() = 0;
  virtual bool restarts(int64_t) { return false; }
  virtual void clear() = 0;
  virtual void clear_computed() = 0;
  virtual void eigenvalues(void* out) = 0;
  virtual void eigenvectors_ptr(const void** p, int64_t* r, int64_t* c) = 0;
  virtual void residuals(double* out) = 0;
  virtual bool alpha_beta(double*, double*) { return false; }
  virtual bool hessenberg(void*) { return false; }
  virtual bool residue(double*) { return false; }
  virtual void small_vectors(void* out, int64_t* r, int64_t* c) = 0;
  virtual void basis_vector(int64_t k, void* out) = 0;
  virtual const std::vector<std::string>& log() = 0;
  virtual void conv_log(int64_t index, void* out, int64_t* n) = 0;
  virtual double bytes() = 0;
  virtual bool exp_lanczos(double, double, void*) { return false; }
  virtual bool exp_taylor(double, double, double, int, const void*, void*) { return false; }
};

namespace {

template <class Solver>
struct Common : cmbs_solver {
  using Scalar = typename Solver::Scalar;
  using Index = typename Solver::Index;
  Solver es;
  void set_operator(cmb_op* op) override { es.setMatrixMultiplication(DeviceOperator<Scalar>::borrow(op)); }
  void set_callback(int64_t n, cmb_matmul_fn fn, void* user) override {
    es.setMatrixMultiplication([fn, user](const Scalar* in, Scalar* out) { fn(in, out, user); }, Index(n));
  }
  void set_indices(const int64_t* idx, int64_t n) override {
    std::vector<Index> v(idx, idx + n);
    es.setIndicesForConvergence(v);
  }
  void set_initial(const void* v, int64_t n) override {
    if (n == 0) {
      es.setInitialVector();
      return;
    }
    es.setInitialVector(static_cast<const Scalar*>(v), Index(n));
  }
  void set_ortho(int64_t nvec, const void* vecs, int64_t ld) override {
    std::vector<typename Solver::VectorType> o;
    const Scalar* p = static_cast<const Scalar*>(vecs);
    const Index n = es.localHeight();
    for (int64_t j = 0; j < nvec; ++j) {
      typename Solver::VectorType x(n);
      memcpy(x.data(), p + size_t(j) * ld, sizeof(Scalar) * size_t(n));
      o.push_back(std::move(x));
    }
    es.setOrthogonalizingVectors(std::move(o));
  }
  void compute() override { es.compute(); }
  void continue_compute() override { es.continueToCompute(); }
  void clear() override { es.clear(); }
  void clear_computed() override { es.clearComputedData(); }
  const std::vector<std::string>& log() override { return es.log(); }
  bool base_set_int(const std::string& k, int64_t v) {
    if (k == "minIterations") es.setMinIterations(v);
    else if (k == "maxIterations") es.setMaxIterations(v);
    else if (k == "maxEigenvalues") es.setMaxEigenvalues(v);
    else if (k == "computeEigenvectorsOn") es.setComputeEigenvectorsOn(v != 0);
    else if (k == "reserveSize") es.setReserveSize(v);
    else return false;
    return true;
  }
  bool base_get_int(const std::string& k, int64_t* v) {
    if (k == "minIterations") *v = es.minIterations();
    else if (k == "maxIterations") *v = es.maxIterations();
    else if (k == "maxEigenvalues") *v = es.maxEigenvalues();
    else if (k == "computeEigenvectorsOn") *v = es.computeEigenvectorsOn();
    else if (k == "reserveSize") *v = es.reserveSize();
    else if (k == "iterations") *v = es.iterations();
    else if (k == "neigenvalues") *v = es.eigenvalues().size();
    else if (k == "nlog") *v = int64_t(es.log().size());
    else if (k == "hasWARN") *v = es.hasWARN();
    else if (k == "hasERROR") *v = es.hasERROR();
    else if (k == "matrixHeight") *v = es.matrixHeight();
    else if (k == "localHeight") *v = es.localHeight();
    else return false;
    return true;
  }
  void eigenvectors_ptr(const void** p, int64_t* r, int64_t* c) override {
    *p = es.eigenvectors().data();
    *r = es.eigenvectors().rows();
    *c = es.eigenvectors().cols();
  }
  void residuals(double* out) override {
    auto r = es.ritzResiduals();
    for (Index i = 0; i < Index(r.size()); ++i) out[i] = r[i];
  }
};

template <class Scalar>
struct LanczosS : Common<LanczosEigenSolver<Scalar>> {
  using B = Common<LanczosEigenSolver<Scalar>>;
  using B::es;
  bool set_int(const std::string& k, int64_t v) override {
    if (k == "reorthogonalizeInterval") {
      es.setReorthogonalizeInterval(v);
      return true;
    }
    return B::base_set_int(k, v);
  }
  bool get_int(const std::string& k, int64_t* v) override {
    if (k == "nvectors") *v = es.lanczosBase().lanczosvectorsSize();
    else if (k == "nalpha") *v = int64_t(es.alpha().size());
    else if (k == "nbeta") *v = int64_t(es.beta().size());
    else return B::base_get_int(k, v);
    return true;
  }
  bool set_real(const std::string& k, double v) override {
    if (k == "tolerance") es.setTolerance(v);
    else if (k == "threshold") es.setThreshold(v);
    else if (k == "eigenvalueShift") es.setEigenvalueShift(v);
    else return false;
    return true;
  }
  bool set_complex(const std::string&, double, double) override { return false; }
  bool get_real(const std::string& k, double* v) override {
    if (k == "tolerance") *v = es.tolerance();
    else if (k == "threshold") *v = es.threshold();
    else if (k == "eigenvalueShift") *v = es.eigenvalueShift();
    else return false;
    return true;
  }
  void eigenvalues(void* out) override {
    double* o = static_cast<double*>(out);
    for (Index i = 0; i < Index(es.eigenvalues().size()); ++i) o[i] = es.eigenvalues()[i];
  }
  bool alpha_beta(double* a, double* b) override {
    std::copy(es.alpha().begin(), es.alpha().end(), a);
    std::copy(es.beta().begin(), es.beta().end(), b);
    return true;
  }
  void small_vectors(void* out, int64_t* r, int64_t* c) override {
    const auto& S = es.es_tri().eigenvectors();
    *r = S.rows();
    *c = S.cols();
    if (out) memcpy(out, S.data(), sizeof(Scalar) * size_t(S.rows()) * size_t(S.cols()));
  }
  void basis_vector(int64_t k, void* out) override {
    const auto& v = es.lanczosvectors();
    memcpy(out, v.at(size_t(k)).data(), sizeof(Scalar) * size_t(es.localHeight()));
  }
  void conv_log(int64_t index, void* out, int64_t* n) override {
    auto it = es.convergenceLog().find(Index(index));
    if (it == es.convergenceLog().end()) {
      *n = 0;
      return;
    }
    *n = int64_t(it->second.size());
    if (out) std::copy(it->second.begin(), it->second.end(), static_cast<double*>(out));
  }
  double bytes() override { return es.lanczosBase().deviceBytes(); }
  static Scalar mkx(double re, double, double*) { return re; }
  static Scalar mkx(double re, double im, std::complex<double>*) { return std::complex<double>(re, im); }
  bool exp_lanczos(double re, double im, void* out) override {
    typename LanczosEigenSolver<Scalar>::VectorType o;
    LanczosExponentialSolver<Scalar>::solveWithLanczos(mkx(re, im, static_cast<Scalar*>(nullptr)), es, o);
    memcpy(out, o.data(), sizeof(Scalar) * size_t(o.size()));
    return true;
  }
  bool exp_taylor(double re, double im, double radius, int autodiv, const void* in, void* out) override {
    using V = typename LanczosEigenSolver<Scalar>::VectorType;
    const Index n = es.matrixHeight();
    V vin(static_cast<const Scalar*>(in), n), o;
    const Scalar x = mkx(re, im, static_cast<Scalar*>(nullptr));
    if (autodiv)
      LanczosExponentialSolver<Scalar>::solveWithTaylorAutoDivision(x, es.matrixMultiplication(), n, radius, vin, o);
    else
      LanczosExponentialSolver<Scalar>::solveWithTaylorNoDivision(x, es.matrixMultiplication(), n, radius, vin, o);
    memcpy(out, o.data(), sizeof(Scalar) * size_t(n));
    return true;
  }
};

template <class Scalar>
struct ArnoldiS : Common<ArnoldiEigenSolver<Scalar>> {
  using B = Common<ArnoldiEigenSolver<Scalar>>;
  using B::es;
  using C = std::complex<double>;
  bool set_int(const std::string& k, int64_t v) override { return B::base_set_int(k, v); }
  bool get_int(const std::string& k, int64_t* v) override {
    if (k == "nvectors") *v = es.arnoldiBase().arnoldivectorsSize();
    else if (k == "hessenbergSize") *v = es.hessenbergMatrix().rows();
    else return B::base_get_int(k, v);
    return true;
  }
  bool set_real(const std::string& k, double v) override {
    if (k == "tolerance") es.setTolerance(v);
    else if (k == "threshold") es.setThreshold(v);
    else if (k == "eigenvalueShift") es.setEigenvalueShift(Scalar(v));
    else return false;
    return true;
  }
  static double mk(double re, double) { return re; }
  bool set_complex(const std::string& k, double re, double im) override {
    if (k != "eigenvalueShift") return false;
    set_shift(re, im, static_cast<Scalar*>(nullptr));
    return true;
  }
  void set_shift(double re, double, double*) { es.setEigenvalueShift(re); }
  void set_shift(double re, double im, C*) { es.setEigenvalueShift(C(re, im)); }
  bool get_real(const std::string& k, double* v) override {
    if (k == "tolerance") *v = es.tolerance();
    else if (k == "threshold") *v = es.threshold();
    else if (k == "eigenvalueShift") *v = std::real(C(es.eigenvalueShift()));
    else return false;
    return true;
  }
  bool restarts(int64_t cycles) override {
    es.computeWithRestarts(cycles);
    return true;
  }
  void eigenvalues(void* out) override {
    C* o = static_cast<C*>(out);
    for (Index i = 0; i < Index(es.eigenvalues().size()); ++i) o[i] = es.eigenvalues()[i];
  }
  bool hessenberg(void* out) override {
    const auto& H = es.hessenbergMatrix();
    memcpy(out, H.data(), sizeof(Scalar) * size_t(H.rows()) * size_t(H.cols()));
    return true;
  }
  bool residue(double* out) override {
    *out = es.arnoldiBase().residue();
    return true;
  }
  void small_vectors(void* out, int64_t* r, int64_t* c) override {
    const auto& Y = es.eigenvectors_h();
    *r = Y.rows();
    *c = Y.cols();
    if (out) memcpy(out, Y.data(), sizeof(C) * size_t(Y.rows()) * size_t(Y.cols()));
  }
  void basis_vector(int64_t k, void* out) override {
    const auto& v = es.arnoldivectors();
    memcpy(out, v.at(size_t(k)).data(), sizeof(Scalar) * size_t(es.localHeight()));
  }
  void conv_log(int64_t index, void* out, int64_t* n) override {
    auto it = es.convergenceLog().find(Index(index));
    if (it == es.convergenceLog().end()) {
      *n = 0;
      return;
    }
    *n = int64_t(it->second.size());
    if (out) std::copy(it->second.begin(), it->second.end(), static_cast<C*>(out));
  }
  double bytes() override { return es.arnoldiBase().deviceBytes(); }
};

// ThickRestartLanczos<Scalar> (additive solver, thick_restart.hpp): the subset of the interface that applies to it.
template <class Scalar>
struct ThickS : cmbs_solver {
  using Solver = ThickRestartLanczos<Scalar>;
  using Index = typename Solver::Index;
  Solver es;
  [[noreturn]] static void no(const char* what) {
    throw LanczosException(std::string(what) + " is not available for the thick-restart solver");
  }
  void set_operator(cmb_op* op) override { es.setMatrixMultiplication(DeviceOperator<Scalar>::borrow(op)); }
  void set_callback(int64_t n, cmb_matmul_fn fn, void* user) override {
    es.setMatrixMultiplication([fn, user](const Scalar* in, Scalar* out) { fn(in, out, user); }, Index(n));
  }
  bool set_int(const std::string& k, int64_t v) override {
    if (k == "wanted") es.setWanted(v);
    else if (k == "maxBasis") es.setMaxBasis(v);
    else if (k == "keep") es.setKeep(v);
    else if (k == "maxRestarts") es.setMaxRestarts(v);
    else if (k == "computeEigenvectorsOn") es.setComputeEigenvectorsOn(v != 0);
    else return false;
    return true;
  }
  bool set_real(const std::string& k, double v) override {
    if (k == "tolerance") es.setTolerance(v);
    else if (k == "threshold") es.setThreshold(v);
    else if (k == "eigenvalueShift") es.setEigenvalueShift(v);
    else return false;
    return true;
  }
  bool set_complex(const std::string&, double, double) override { return false; }
  bool get_int(const std::string& k, int64_t* v) override {
    if (k == "wanted") *v = es.wanted();
    else if (k == "maxBasis") *v = es.maxBasis();
    else if (k == "keep") *v = es.keep();
    else if (k == "maxRestarts") *v = es.maxRestarts();
    else if (k == "restarts") *v = es.restarts();
    else if (k == "operatorApplications" || k == "iterations") *v = es.operatorApplications();
    else if (k == "converged") *v = es.converged();
    else if (k == "neigenvalues") *v = es.eigenvalues().size();
    else if (k == "nlog") *v = int64_t(es.log().size());
    else if (k == "nvectors") *v = es.lanczosBase().lanczosvectorsSize();
    else if (k == "matrixHeight") *v = es.lanczosBase().matrixHeight();
    else if (k == "localHeight") *v = es.localHeight();
    else if (k == "hasWARN" || k == "hasERROR") {
      const std::string head = (k == "hasWARN") ? Solver::headWARN() : std::string("ERROR     ");
      int64_t c = 0;
      for (const auto& l : es.log()) c += (l.find(head) == 0);
      *v = c;
    } else return false;
    return true;
  }
  bool get_real(const std::string& k, double* v) override {
    if (k == "tolerance") *v = es.tolerance();
    else return false;
    return true;
  }
  void set_indices(const int64_t*, int64_t) override { no("indicesForConvergence"); }
  void set_initial(const void* v, int64_t n) override {
    if (n == 0) {
      es.setInitialVector();
      return;
    }
    typename Solver::VectorType x(n);
    memcpy(x.data(), v, sizeof(Scalar) * size_t(n));
    es.setInitialVector(x);
  }
  void set_ortho(int64_t nvec, const void* vecs, int64_t ld) override {
    std::vector<typename Solver::VectorType> o;
    const Scalar* p = static_cast<const Scalar*>(vecs);
    const Index n = es.localHeight();
    for (int64_t j = 0; j < nvec; ++j) {
      typename Solver::VectorType x(n);
      memcpy(x.data(), p + size_t(j) * ld, sizeof(Scalar) * size_t(n));
      o.push_back(std::move(x));
    }
    es.setOrthogonalizingVectors(o);
  }
  void compute() override { es.compute(); }
  void continue_compute() override { no("continueToCompute"); }
  void clear() override { es = Solver(); }
  void clear_computed() override { no("clearComputedData"); }
  void eigenvalues(void* out) override {
    double* o = static_cast<double*>(out);
    for (Index i = 0; i < Index(es.eigenvalues().size()); ++i) o[i] = es.eigenvalues()[i];
  }
  void eigenvectors_ptr(const void** p, int64_t* r, int64_t* c) override {
    *p = es.eigenvectors().data();
    *r = es.eigenvectors().rows();
    *c = es.eigenvectors().cols();
  }
  void residuals(double* out) override {
    for (Index i = 0; i < Index(es.residuals().size()); ++i) out[i] = es.residuals()[i];
  }
  void small_vectors(void*, int64_t*, int64_t*) override { no("the projected eigenvectors"); }
  void basis_vector(int64_t, void*) override { no("basis vectors"); }
  const std::vector<std::string>& log() override { return es.log(); }
  void conv_log(int64_t, void*, int64_t* n) override { *n = 0; }
  double bytes() override { return es.deviceBytes(); }
};

// ThickRestartArnoldi<Scalar> (additive solver, arnoldi_restart.hpp)
template <class Scalar>
struct ThickArnoldiS : cmbs_solver {
  using Solver = ThickRestartArnoldi<Scalar>;
  using Index = typename Solver::Index;
  using C = std::complex<double>;
  Solver es;
  [[noreturn]] static void no(const char* what) {
    throw LanczosException(std::string(what) + " is not available for the thick-restart Arnoldi solver");
  }
  void set_operator(cmb_op* op) override { es.setMatrixMultiplication(DeviceOperator<Scalar>::borrow(op)); }
  void set_callback(int64_t n, cmb_matmul_fn fn, void* user) override {
    es.setMatrixMultiplication([fn, user](const Scalar* in, Scalar* out) { fn(in, out, user); }, Index(n));
  }
  bool set_int(const std::string& k, int64_t v) override {
    if (k == "wanted") es.setWanted(v);
    else if (k == "maxBasis") es.setMaxBasis(v);
    else if (k == "keep") es.setKeep(v);
    else if (k == "maxRestarts") es.setMaxRestarts(v);
    else if (k == "which") es.setWhich(static_cast<typename Solver::Which>(v));
    else if (k == "computeEigenvectorsOn") es.setComputeEigenvectorsOn(v != 0);
    else return false;
    return true;
  }
  bool set_real(const std::string& k, double v) override {
    if (k == "tolerance") es.setTolerance(v);
    else if (k == "threshold") es.setThreshold(v);
    else if (k == "eigenvalueShift") es.setEigenvalueShift(Scalar(v));
    else return false;
    return true;
  }
  bool set_complex(const std::string& k, double re, double im) override {
    if (k != "eigenvalueShift") return false;
    es.setEigenvalueShift(make_shift(re, im, static_cast<Scalar*>(nullptr)));
    return true;
  }
  static double make_shift(double re, double, double*) { return re; }
  static C make_shift(double re, double im, C*) { return C(re, im); }
  bool get_int(const std::string& k, int64_t* v) override {
    if (k == "wanted") *v = es.wanted();
    else if (k == "maxBasis") *v = es.maxBasis();
    else if (k == "keep") *v = es.keep();
    else if (k == "maxRestarts") *v = es.maxRestarts();
    else if (k == "restarts") *v = es.restarts();
    else if (k == "operatorApplications" || k == "iterations") *v = es.operatorApplications();
    else if (k == "converged") *v = es.converged();
    else if (k == "neigenvalues") *v = es.eigenvalues().size();
    else if (k == "nlog") *v = int64_t(es.log().size());
    else if (k == "nvectors") *v = es.arnoldiBase().arnoldivectorsSize();
    else if (k == "matrixHeight") *v = es.arnoldiBase().matrixHeight();
    else if (k == "localHeight") *v = es.localHeight();
    else if (k == "hasWARN" || k == "hasERROR") {
      const std::string head = (k == "hasWARN") ? Solver::headWARN() : std::string("ERROR     ");
      int64_t c = 0;
      for (const auto& l : es.log()) c += (l.find(head) == 0);
      *v = c;
    } else return false;
    return true;
  }
  bool get_real(const std::string& k, double* v) override {
    if (k == "tolerance") *v = es.tolerance();
    else return false;
    return true;
  }
  void set_indices(const int64_t*, int64_t) override { no("indicesForConvergence"); }
  void set_initial(const void* v, int64_t n) override {
    if (n == 0) {
      es.setInitialVector();
      return;
    }
    typename Solver::VectorType x(n);
    memcpy(x.data(), v, sizeof(Scalar) * size_t(n));
    es.setInitialVector(x);
  }
  void set_ortho(int64_t nvec, const void* vecs, int64_t ld) override {
    std::vector<typename Solver::VectorType> o;
    const Scalar* p = static_cast<const Scalar*>(vecs);
    const Index n = es.localHeight();
    for (int64_t j = 0; j < nvec; ++j) {
      typename Solver::VectorType x(n);
      memcpy(x.data(), p + size_t(j) * ld, sizeof(Scalar) * size_t(n));
      o.push_back(std::move(x));
    }
    es.setOrthogonalizingVectors(o);
  }
  void compute() override { es.compute(); }
  void continue_compute() override { no("continueToCompute"); }
  void clear() override { es = Solver(); }
  void clear_computed() override { no("clearComputedData"); }
  void eigenvalues(void* out) override {
    C* o = static_cast<C*>(out);
    for (Index i = 0; i < Index(es.eigenvalues().size()); ++i) o[i] = es.eigenvalues()[i];
  }
  void eigenvectors_ptr(const void** p, int64_t* r, int64_t* c) override {
    *p = es.eigenvectors().data();
    *r = es.eigenvectors().rows();
    *c = es.eigenvectors().cols();
  }
  void residuals(double* out) override {
    for (Index i = 0; i < Index(es.residuals().size()); ++i) out[i] = es.residuals()[i];
  }
  void small_vectors(void*, int64_t*, int64_t*) override { no("the projected eigenvectors"); }
  void basis_vector(int64_t, void*) override { no("basis vectors"); }
  const std::vector<std::string>& log() override { return es.log(); }
  void conv_log(int64_t, void*, int64_t* n) override { *n = 0; }
  double bytes() override { return es.deviceBytes(); }
};

template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    set_error("%s", e.what());
    return CMB_ERR_INVALID;
  } catch (...) {
    set_error("unknown C++ exception");
    return CMB_ERR_INVALID;
  }
}

}  // namespace

#define S_REQ(c, m)          \
  do {                       \
    if (!(c)) {              \
      set_error("%s", m);    \
      return CMB_ERR_INVALID; \
    }                        \
  } while (0)

extern "C" {

int cmbs_create(int kind, cmb_dtype dtype, cmbs_solver** out) {
  S_REQ(out, "null argument");
  *out = nullptr;
  return guarded([&]() -> int {
    cmbs_solver* s = nullptr;
    if (kind == CMBS_LANCZOS && dtype == CMB_F64) s = new LanczosS<double>();
    else if (kind == CMBS_LANCZOS && dtype == CMB_C64) s = new LanczosS<std::complex<double>>();
    else if (kind == CMBS_ARNOLDI && dtype == CMB_F64) s = new ArnoldiS<double>();
    else if (kind == CMBS_ARNOLDI && dtype == CMB_C64) s = new ArnoldiS<std::complex<double>>();
    else if (kind == CMBS_THICK_RESTART && dtype == CMB_F64) s = new ThickS<double>();
    else if (kind == CMBS_THICK_RESTART && dtype == CMB_C64) s = new ThickS<std::complex<double>>();
    else if (kind == CMBS_THICK_RESTART_ARNOLDI && dtype == CMB_F64) s = new ThickArnoldiS<double>();
    else if (kind == CMBS_THICK_RESTART_ARNOLDI && dtype == CMB_C64) s = new ThickArnoldiS<std::complex<double>>();
    S_REQ(s, "unknown solver kind / dtype");
    s->kind = kind;
    s->dtype = dtype;
    *out = s;
    return CMB_OK;
  });
}
int cmbs_destroy(cmbs_solver* s) {
  return guarded([&]() -> int {
    delete s;
    return CMB_OK;
  });
}
int cmbs_set_operator(cmbs_solver* s, cmb_ctx* ctx, cmb_op* op) {
  S_REQ(s && op, "null argument");
  S_REQ(ctx == nullptr || cmb_op_context(op) == ctx, "operator belongs to another context");
  S_REQ(cmb_op_dtype(op) == s->dtype, "operator dtype differs from the solver's Scalar");
  return guarded([&]() -> int {
    s->set_operator(op);
    return CMB_OK;
  });
}
int cmbs_set_callback(cmbs_solver* s, int64_t height, cmb_matmul_fn fn, void* user) {
  S_REQ(s && fn, "null argument");
  return guarded([&]() -> int {
    s->set_callback(height, fn, user);
    return CMB_OK;
  });
}
int cmbs_set_int(cmbs_solver* s, const char* name, int64_t value) {
  S_REQ(s && name, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->set_int(name, value), "unknown integer setting");
    return CMB_OK;
  });
}
int cmbs_set_real(cmbs_solver* s, const char* name, double value) {
  S_REQ(s && name, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->set_real(name, value), "unknown real setting");
    return CMB_OK;
  });
}
int cmbs_set_complex(cmbs_solver* s, const char* name, double re, double im) {
  S_REQ(s && name, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->set_complex(name, re, im), "unknown complex setting");
    return CMB_OK;
  });
}
int cmbs_get_int(cmbs_solver* s, const char* name, int64_t* value) {
  S_REQ(s && name && value, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->get_int(name, value), "unknown integer property");
    return CMB_OK;
  });
}
int cmbs_get_real(cmbs_solver* s, const char* name, double* value) {
  S_REQ(s && name && value, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->get_real(name, value), "unknown real property");
    return CMB_OK;
  });
}
int cmbs_set_indices_for_convergence(cmbs_solver* s, const int64_t* idx, int64_t n) {
  S_REQ(s && (n == 0 || idx) && n >= 0, "bad argument");
  return guarded([&]() -> int {
    s->set_indices(idx, n);
    return CMB_OK;
  });
}
int cmbs_set_initial_vector(cmbs_solver* s, const void* v, int64_t n) {
  S_REQ(s && (n == 0 || v) && n >= 0, "bad argument");
  return guarded([&]() -> int {
    s->set_initial(v, n);
    return CMB_OK;
  });
}
int cmbs_set_orthogonalizing_vectors(cmbs_solver* s, int64_t nvec, const void* vecs, int64_t ld) {
  S_REQ(s && nvec >= 0 && (nvec == 0 || vecs), "bad argument");
  return guarded([&]() -> int {
    s->set_ortho(nvec, vecs, ld);
    return CMB_OK;
  });
}
int cmbs_compute(cmbs_solver* s) {
  S_REQ(s, "null argument");
  return guarded([&]() -> int {
    s->compute();
    return CMB_OK;
  });
}
int cmbs_continue_to_compute(cmbs_solver* s) {
  S_REQ(s, "null argument");
  return guarded([&]() -> int {
    s->continue_compute();
    return CMB_OK;
  });
}
int cmbs_compute_with_restarts(cmbs_solver* s, int64_t cycles) {
  S_REQ(s && cycles >= 1, "bad argument");
  return guarded([&]() -> int {
    S_REQ(s->restarts(cycles), "restarts are an Arnoldi feature");
    return CMB_OK;
  });
}
int cmbs_clear(cmbs_solver* s) {
  S_REQ(s, "null argument");
  return guarded([&]() -> int {
    s->clear();
    return CMB_OK;
  });
}
int cmbs_clear_computed_data(cmbs_solver* s) {
  S_REQ(s, "null argument");
  return guarded([&]() -> int {
    s->clear_computed();
    return CMB_OK;
  });
}
int cmbs_get_eigenvalues(cmbs_solver* s, void* out) {
  S_REQ(s && out, "null argument");
  return guarded([&]() -> int {
    s->eigenvalues(out);
    return CMB_OK;
  });
}
int cmbs_eigenvectors_ptr(cmbs_solver* s, const void** ptr, int64_t* rows, int64_t* cols) {
  S_REQ(s && ptr && rows && cols, "null argument");
  return guarded([&]() -> int {
    s->eigenvectors_ptr(ptr, rows, cols);
    return CMB_OK;
  });
}
int cmbs_get_ritz_residuals(cmbs_solver* s, double* out) {
  S_REQ(s && out, "null argument");
  return guarded([&]() -> int {
    s->residuals(out);
    return CMB_OK;
  });
}
int cmbs_get_alpha_beta(cmbs_solver* s, double* alpha, double* beta) {
  S_REQ(s && alpha && beta, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->alpha_beta(alpha, beta), "alpha/beta are Lanczos properties");
    return CMB_OK;
  });
}
int cmbs_get_hessenberg(cmbs_solver* s, void* out) {
  S_REQ(s && out, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->hessenberg(out), "the Hessenberg matrix is an Arnoldi property");
    return CMB_OK;
  });
}
int cmbs_get_residue(cmbs_solver* s, double* out) {
  S_REQ(s && out, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->residue(out), "residue is an Arnoldi property");
    return CMB_OK;
  });
}
int cmbs_get_small_eigenvectors(cmbs_solver* s, void* out, int64_t* rows, int64_t* cols) {
  S_REQ(s && rows && cols, "null argument");
  return guarded([&]() -> int {
    s->small_vectors(out, rows, cols);
    return CMB_OK;
  });
}
int cmbs_get_basis_vector(cmbs_solver* s, int64_t k, void* out) {
  S_REQ(s && out && k >= 0, "bad argument");
  return guarded([&]() -> int {
    s->basis_vector(k, out);
    return CMB_OK;
  });
}
int cmbs_get_log_line(cmbs_solver* s, int64_t i, char* buf, int64_t buflen) {
  S_REQ(s && buf && buflen > 0, "bad argument");
  return guarded([&]() -> int {
    const auto& lg = s->log();
    S_REQ(i >= 0 && i < int64_t(lg.size()), "log index out of range");
    strncpy(buf, lg[size_t(i)].c_str(), size_t(buflen) - 1);
    buf[buflen - 1] = 0;
    return CMB_OK;
  });
}
int cmbs_get_convergence_log(cmbs_solver* s, int64_t index, void* out, int64_t* n) {
  S_REQ(s && n, "null argument");
  return guarded([&]() -> int {
    s->conv_log(index, out, n);
    return CMB_OK;
  });
}
double cmbs_device_bytes(cmbs_solver* s) {
  double b = 0.0;
  if (s) guarded([&]() -> int {
      b = s->bytes();
      return CMB_OK;
    });
  return b;
}

int cmbs_exp_solve_with_lanczos(cmbs_solver* s, double x_re, double x_im, void* out) {
  S_REQ(s && out, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->exp_lanczos(x_re, x_im, out), "exp(xA)v is a Lanczos feature");
    return CMB_OK;
  });
}
int cmbs_exp_solve_with_taylor(cmbs_solver* s, double x_re, double x_im, double matrix_radius, int auto_division,
                               const void* in, void* out) {
  S_REQ(s && in && out, "null argument");
  return guarded([&]() -> int {
    S_REQ(s->exp_taylor(x_re, x_im, matrix_radius, auto_division, in, out), "exp(xA)v is a Lanczos feature");
    return CMB_OK;
  });
}

int cmbs_host_tridiagonal_eigen(int64_t n, const double* alpha, const double* beta, double* w, double* z) {
  S_REQ(n >= 0 && (n == 0 || (alpha && w)), "bad argument");
  return guarded([&]() -> int {
    std::vector<double> ww, zz;
    bool ok = z ? detail::tridiagonal_eigensystem<double>(alpha, beta, int(n), ww, zz)
                : detail::tridiagonal_eigenvalues<double>(alpha, beta, int(n), ww);
    std::copy(ww.begin(), ww.end(), w);
    if (z) std::copy(zz.begin(), zz.end(), z);
    S_REQ(ok, "tridiagonal QR did not converge");
    return CMB_OK;
  });
}
int cmbs_host_symmetric_eigen(int64_t n, const double* a, double* w, double* z) {
  S_REQ(n >= 0 && (n == 0 || (a && w && z)), "bad argument");
  return guarded([&]() -> int {
    std::vector<double> aa(a, a + size_t(n) * size_t(n)), ww, zz;
    const bool ok = detail::symmetric_eigensystem<double>(int(n), aa, ww, zz);
    std::copy(ww.begin(), ww.end(), w);
    std::copy(zz.begin(), zz.end(), z);
    S_REQ(ok, "Jacobi iteration did not converge");
    return CMB_OK;
  });
}
int cmbs_host_hessenberg_eigen(int64_t n, const void* h, void* w, void* v) {
  S_REQ(n >= 0 && (n == 0 || (h && w)), "bad argument");
  return guarded([&]() -> int {
    using C = std::complex<double>;
    std::vector<C> ww, vv;
    bool ok = detail::hessenberg_eigen<double>(int(n), static_cast<const C*>(h), ww, v ? &vv : nullptr);
    std::copy(ww.begin(), ww.end(), static_cast<C*>(w));
    if (v) std::copy(vv.begin(), vv.end(), static_cast<C*>(v));
    S_REQ(ok, "Hessenberg QR did not converge");
    return CMB_OK;
  });
}
int cmbs_host_general_eigen(int64_t n, const void* a, void* w, void* v) {
  S_REQ(n >= 0 && (n == 0 || (a && w)), "bad argument");
  return guarded([&]() -> int {
    using C = std::complex<double>;
    std::vector<C> ww, vv;
    bool ok = detail::general_eigen<double>(int(n), static_cast<const C*>(a), ww, v ? &vv : nullptr);
    std::copy(ww.begin(), ww.end(), static_cast<C*>(w));
    if (v) std::copy(vv.begin(), vv.end(), static_cast<C*>(v));
    S_REQ(ok, "QR iteration did not converge");
    return CMB_OK;
  });
}

int cmb_host_alloc(size_t bytes, void** out) {
  S_REQ(out, "null argument");
  *out = nullptr;
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return CMB_ERR_NOMEM;
  }
  return CMB_OK;
}
int cmb_host_free(void* p) {
  if (p) cudaFreeHost(p);
  return CMB_OK;
}

}  // extern "C"
