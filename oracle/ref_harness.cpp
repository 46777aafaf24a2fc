// ref_harness.cpp — C-ABI around the UNMODIFIED reference solver classes.  TEST INFRASTRUCTURE ONLY.
//
// This translation unit includes /root/reference/include/cmpt/eigen_ex/{lanczos,arnoldi}.hpp exactly as they lie
// in the reference tree (-I/root/reference/include; nothing is copied into this repo) and compiles them against
// the stand-in Eigen of oracle/eigen_shim/ (the image has no Eigen).  The result, oracle/_ref/libref.so, is the
// reference's own CPU implementation of the hot path: LanczosBase / LanczosEigenSolver / LanczosExponentialSolver
// (lanczos.hpp:104-1164) and ArnoldiBase / ArnoldiEigenSolver (arnoldi.hpp:53-1027).  It is used
//   * to pin oracle/krylov_oracle.cpp + oracle/reference_solvers.py (tests/test_ref_pin.py),
//   * as the direct parity target of the CUDA path (tests/test_gpu_vs_ref.py),
//   * to generate tests/golden/ref_traces.npz (tests/golden/make_ref_golden.py),
//   * as bench.py's cpu_baseline / --impl reference leg (kind "reference").
// Only tests/, __graft_entry__.smoke() and those bench legs load it; the product never does.
//
// The operator is whatever the caller passes as a C callback apply(user, in, out) — the reference's
// MatMulFunction is built around it (lanczos.hpp:116,179-183).
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <random>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "cmpt/eigen_ex/arnoldi.hpp"
#include "cmpt/eigen_ex/lanczos.hpp"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

using cmpt::EigenEx::ArnoldiBase;
using cmpt::EigenEx::ArnoldiEigenSolver;
using cmpt::EigenEx::LanczosBase;
using cmpt::EigenEx::LanczosEigenSolver;
using cmpt::EigenEx::LanczosExponentialSolver;
using Index = Eigen::Index;
typedef void (*apply_fn)(void* user, const void* in, void* out);

template <class S>
using Vec = Eigen::Matrix<S, Eigen::Dynamic, 1>;

template <class S>
std::function<void(const S*, S*)> make_matmul(apply_fn f, void* user) {
  if (!f) return std::function<void(const S*, S*)>();
  return [f, user](const S* in, S* out) { f(user, in, out); };
}

template <class S>
Vec<S> to_vec(const void* p, int64_t n) {
  Vec<S> v(static_cast<Index>(n));
  if (n > 0) std::memcpy(static_cast<void*>(v.data()), p, sizeof(S) * static_cast<std::size_t>(n));
  return v;
}

template <class V>
void from_vec(const V& v, void* out) {
  using S = typename V::Scalar;
  if (v.size() > 0) std::memcpy(out, static_cast<const void*>(v.data()), sizeof(S) * static_cast<std::size_t>(v.size()));
}

int copy_string(const std::string& s, char* buf, int64_t cap) {
  if (cap <= 0) return static_cast<int>(s.size());
  const std::size_t m = std::min<std::size_t>(s.size(), static_cast<std::size_t>(cap - 1));
  std::memcpy(buf, s.data(), m);
  buf[m] = 0;
  return static_cast<int>(s.size());
}

// ---- step level: LanczosBase<S> ---------------------------------------------------------------------
template <class S>
struct LBase {
  LanczosBase<S> b;
  LBase() { b.clearLanczosSteps(); }  // the reference leaves iterations_ uninitialised until a clear
};

// ---- step level: ArnoldiBase<S> ---------------------------------------------------------------------
template <class S>
struct ABase {
  ArnoldiBase<S> b;
  ABase() { b.clearArnoldiSteps(); }
};

// ---- solver level ------------------------------------------------------------------------------------
template <class S>
struct LSolver {
  LanczosEigenSolver<S> es;
  std::string error;
};
// The reference's accessors for the convergence log and for the sorted Hessenberg eigenvectors do not compile
// (arnoldi.hpp:666,671), so the protected members are reached through a derived class; the reference class itself
// stays untouched.
struct ArnoldiAccess : ArnoldiEigenSolver<std::complex<double>> {
  const std::map<Index, std::vector<std::complex<double>>>& convergence_log() const { return this->convergenceLog_; }
  const ComplexMatrixType& sorted_eigenvectors_h() const { return this->eigenvectors_h_; }
};
struct ASolver {
  ArnoldiAccess es;
  std::string error;
};

template <class F>
int guarded(std::string& err, const F& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    err = e.what();
    return 1;
  } catch (...) {
    err = "unknown exception";
    return 2;
  }
}

}  // namespace

REF_API int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
REF_API void ref_set_num_threads(int t) {
#ifdef _OPENMP
  if (t > 0) omp_set_num_threads(t);
#else
  (void)t;
#endif
}

#define REF_COMMON(P, S)                                                                                          \
  /* random start vectors: LanczosBase::setInitialVector() (lanczos.hpp:214-218) and makeRandomVector */         \
  REF_API void ref_##P##_default_vector(int64_t n, void* out) {                                                   \
    LanczosBase<S> b;                                                                                             \
    b.setMatrixMultiplication([](const S*, S*) {}, static_cast<Index>(n));                                        \
    b.setInitialVector();                                                                                         \
    from_vec(b.initialVector(), out);                                                                             \
  }                                                                                                               \
  REF_API void ref_##P##_seeded_vector(uint32_t seed, int64_t n, void* out) {                                     \
    std::mt19937 g(seed);                                                                                         \
    from_vec(LanczosBase<S>::makeRandomVector(g, static_cast<Index>(n)), out);                                    \
  }                                                                                                               \
  REF_API void ref_##P##_arnoldi_default_vector(int64_t n, void* out) {                                           \
    ArnoldiBase<S> b;                                                                                             \
    b.setMatrixMultiplication([](const S*, S*) {}, static_cast<Index>(n));                                        \
    b.setInitialVector();                                                                                         \
    from_vec(b.initialVector(), out);                                                                             \
  }                                                                                                               \
  /* ---- LanczosBase ---- */                                                                                     \
  REF_API void* ref_##P##_lanczos_create() { return new LBase<S>(); }                                             \
  REF_API void ref_##P##_lanczos_destroy(void* h) { delete static_cast<LBase<S>*>(h); }                           \
  REF_API void ref_##P##_lanczos_set_op(void* h, apply_fn f, void* user, int64_t n) {                             \
    static_cast<LBase<S>*>(h)->b.setMatrixMultiplication(make_matmul<S>(f, user), static_cast<Index>(n));         \
  }                                                                                                               \
  REF_API void ref_##P##_lanczos_set_init(void* h, const void* v, int64_t n) {                                    \
    static_cast<LBase<S>*>(h)->b.setInitialVector(to_vec<S>(v, n));                                               \
  }                                                                                                               \
  REF_API void ref_##P##_lanczos_add_ortho(void* h, const void* v, int64_t n) {                                   \
    static_cast<LBase<S>*>(h)->b.refOrthogonalizingVectors().push_back(to_vec<S>(v, n));                          \
  }                                                                                                               \
  REF_API void ref_##P##_lanczos_set_params(void* h, double shift, int64_t interval, double threshold) {          \
    auto& b = static_cast<LBase<S>*>(h)->b;                                                                       \
    b.setEigenvalueShift(shift).setReorthogonalizeInterval(static_cast<Index>(interval)).setThreshold(threshold); \
  }                                                                                                               \
  REF_API void ref_##P##_lanczos_clear_steps(void* h) { static_cast<LBase<S>*>(h)->b.clearLanczosSteps(); }       \
  REF_API int ref_##P##_lanczos_step(void* h) { return static_cast<LBase<S>*>(h)->b.updateLanczosSteps() ? 1 : 0; } \
  REF_API int ref_##P##_lanczos_utmost(void* h) { return static_cast<LBase<S>*>(h)->b.lanczosStepIsUtmost() ? 1 : 0; } \
  REF_API int64_t ref_##P##_lanczos_iterations(void* h) { return static_cast<LBase<S>*>(h)->b.iterations(); }     \
  REF_API int64_t ref_##P##_lanczos_nvectors(void* h) {                                                           \
    return static_cast<int64_t>(static_cast<LBase<S>*>(h)->b.lanczosvectors().size());                            \
  }                                                                                                               \
  REF_API void ref_##P##_lanczos_get_vector(void* h, int64_t k, void* out) {                                      \
    from_vec(static_cast<LBase<S>*>(h)->b.lanczosvectors()[static_cast<std::size_t>(k)], out);                    \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lanczos_nalpha(void* h) {                                                             \
    return static_cast<int64_t>(static_cast<LBase<S>*>(h)->b.alpha().size());                                     \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lanczos_nbeta(void* h) {                                                              \
    return static_cast<int64_t>(static_cast<LBase<S>*>(h)->b.beta().size());                                      \
  }                                                                                                               \
  REF_API void ref_##P##_lanczos_get_alpha_beta(void* h, double* a, double* b) {                                  \
    auto& base = static_cast<LBase<S>*>(h)->b;                                                                    \
    std::copy(base.alpha().begin(), base.alpha().end(), a);                                                       \
    std::copy(base.beta().begin(), base.beta().end(), b);                                                         \
  }                                                                                                               \
  /* ---- ArnoldiBase ---- */                                                                                     \
  REF_API void* ref_##P##_arnoldi_create() { return new ABase<S>(); }                                             \
  REF_API void ref_##P##_arnoldi_destroy(void* h) { delete static_cast<ABase<S>*>(h); }                           \
  REF_API void ref_##P##_arnoldi_set_op(void* h, apply_fn f, void* user, int64_t n) {                             \
    static_cast<ABase<S>*>(h)->b.setMatrixMultiplication(make_matmul<S>(f, user), static_cast<Index>(n));         \
  }                                                                                                               \
  REF_API void ref_##P##_arnoldi_set_init(void* h, const void* v, int64_t n) {                                    \
    static_cast<ABase<S>*>(h)->b.setInitialVector(to_vec<S>(v, n));                                               \
  }                                                                                                               \
  REF_API void ref_##P##_arnoldi_add_ortho(void* h, const void* v, int64_t n) {                                   \
    static_cast<ABase<S>*>(h)->b.refOrthogonalizingVectors().push_back(to_vec<S>(v, n));                          \
  }                                                                                                               \
  REF_API void ref_##P##_arnoldi_set_params(void* h, const void* shift, double threshold) {                       \
    auto& b = static_cast<ABase<S>*>(h)->b;                                                                       \
    b.setEigenvalueShift(*static_cast<const S*>(shift)).setThreshold(threshold);                                  \
  }                                                                                                               \
  REF_API void ref_##P##_arnoldi_clear_steps(void* h) { static_cast<ABase<S>*>(h)->b.clearArnoldiSteps(); }       \
  REF_API int ref_##P##_arnoldi_step(void* h) { return static_cast<ABase<S>*>(h)->b.updateArnoldiSteps() ? 1 : 0; } \
  REF_API int ref_##P##_arnoldi_utmost(void* h) { return static_cast<ABase<S>*>(h)->b.arnoldiStepIsUtmost() ? 1 : 0; } \
  REF_API int64_t ref_##P##_arnoldi_iterations(void* h) { return static_cast<ABase<S>*>(h)->b.iterations(); }     \
  REF_API int64_t ref_##P##_arnoldi_nvectors(void* h) {                                                           \
    return static_cast<int64_t>(static_cast<ABase<S>*>(h)->b.arnoldivectors().size());                            \
  }                                                                                                               \
  REF_API void ref_##P##_arnoldi_get_vector(void* h, int64_t k, void* out) {                                      \
    from_vec(static_cast<ABase<S>*>(h)->b.arnoldivectors()[static_cast<std::size_t>(k)], out);                    \
  }                                                                                                               \
  REF_API double ref_##P##_arnoldi_residue(void* h) { return static_cast<ABase<S>*>(h)->b.residue(); }            \
  REF_API int64_t ref_##P##_arnoldi_hess_size(void* h) {                                                          \
    return static_cast<ABase<S>*>(h)->b.makeHessenbergMatrix().rows();                                            \
  }                                                                                                               \
  REF_API void ref_##P##_arnoldi_hessenberg(void* h, void* out) { /* column-major hs x hs */                      \
    from_vec(static_cast<ABase<S>*>(h)->b.makeHessenbergMatrix(), out);                                           \
  }                                                                                                               \
  REF_API int64_t ref_##P##_arnoldi_matrix_cols(void* h) {                                                        \
    return static_cast<ABase<S>*>(h)->b.makeArnoldiMatrix().cols();                                               \
  }                                                                                                               \
  /* ---- LanczosEigenSolver ---- */                                                                              \
  REF_API void* ref_##P##_lsolver_create() { return new LSolver<S>(); }                                           \
  REF_API void ref_##P##_lsolver_destroy(void* h) { delete static_cast<LSolver<S>*>(h); }                         \
  REF_API void ref_##P##_lsolver_set_op(void* h, apply_fn f, void* user, int64_t n) {                             \
    static_cast<LSolver<S>*>(h)->es.setMatrixMultiplication(make_matmul<S>(f, user), static_cast<Index>(n));      \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_set_init(void* h, const void* v, int64_t n) {                                    \
    static_cast<LSolver<S>*>(h)->es.setInitialVector(to_vec<S>(v, n));                                            \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_set_init_seeded(void* h, uint32_t seed) { /* sample_lanczos2.cpp:39,53 */        \
    auto& es = static_cast<LSolver<S>*>(h)->es;                                                                   \
    std::mt19937 g(seed);                                                                                         \
    es.setInitialVector(es.lanczosBase().makeRandomVector(g, es.matrixHeight()));                                 \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_get_init(void* h, void* out) {                                                   \
    from_vec(static_cast<LSolver<S>*>(h)->es.initialVector(), out);                                               \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lsolver_init_size(void* h) { return static_cast<LSolver<S>*>(h)->es.initialVector().size(); } \
  REF_API void ref_##P##_lsolver_clear_ortho(void* h) {                                                           \
    static_cast<LSolver<S>*>(h)->es.setOrthogonalizingVectors(std::vector<Vec<S>>());                             \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_add_ortho(void* h, const void* v, int64_t n) {                                   \
    static_cast<LSolver<S>*>(h)->es.refOrthogonalizingVectors().push_back(to_vec<S>(v, n));                       \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_set_params(void* h, double shift, int64_t interval, double threshold,            \
                                            double tolerance, int64_t min_it, int64_t max_it, int64_t max_eig,    \
                                            int vectors_on) {                                                     \
    auto& es = static_cast<LSolver<S>*>(h)->es;                                                                   \
    es.setEigenvalueShift(shift).setReorthogonalizeInterval(static_cast<Index>(interval)).setThreshold(threshold); \
    es.setTolerance(tolerance).setMinIterations(static_cast<Index>(min_it)).setMaxIterations(static_cast<Index>(max_it)); \
    es.setMaxEigenvalues(static_cast<Index>(max_eig)).setComputeEigenvectorsOn(vectors_on != 0);                  \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_set_indices(void* h, const int64_t* idx, int64_t count) {                        \
    std::vector<Index> v(idx, idx + count);                                                                       \
    static_cast<LSolver<S>*>(h)->es.setIndicesForConvergence(v);                                                  \
  }                                                                                                               \
  REF_API int ref_##P##_lsolver_compute(void* h) {                                                                \
    auto* s = static_cast<LSolver<S>*>(h);                                                                        \
    return guarded(s->error, [s] { s->es.compute(); });                                                           \
  }                                                                                                               \
  REF_API int ref_##P##_lsolver_continue(void* h) {                                                               \
    auto* s = static_cast<LSolver<S>*>(h);                                                                        \
    return guarded(s->error, [s] { s->es.continueToCompute(); });                                                 \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_clear(void* h) { static_cast<LSolver<S>*>(h)->es.clear(); }                      \
  REF_API int64_t ref_##P##_lsolver_iterations(void* h) { return static_cast<LSolver<S>*>(h)->es.iterations(); }  \
  REF_API int64_t ref_##P##_lsolver_nvectors(void* h) {                                                           \
    return static_cast<int64_t>(static_cast<LSolver<S>*>(h)->es.lanczosvectors().size());                         \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_get_vector(void* h, int64_t k, void* out) {                                      \
    from_vec(static_cast<LSolver<S>*>(h)->es.lanczosvectors()[static_cast<std::size_t>(k)], out);                 \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lsolver_nalpha(void* h) { return static_cast<int64_t>(static_cast<LSolver<S>*>(h)->es.alpha().size()); } \
  REF_API int64_t ref_##P##_lsolver_nbeta(void* h) { return static_cast<int64_t>(static_cast<LSolver<S>*>(h)->es.beta().size()); } \
  REF_API void ref_##P##_lsolver_get_alpha_beta(void* h, double* a, double* b) {                                  \
    auto& es = static_cast<LSolver<S>*>(h)->es;                                                                   \
    std::copy(es.alpha().begin(), es.alpha().end(), a);                                                           \
    std::copy(es.beta().begin(), es.beta().end(), b);                                                             \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lsolver_neig(void* h) { return static_cast<LSolver<S>*>(h)->es.eigenvalues().size(); } \
  REF_API void ref_##P##_lsolver_get_eigenvalues(void* h, double* out) {                                          \
    from_vec(static_cast<LSolver<S>*>(h)->es.eigenvalues(), out);                                                 \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_eigenvectors_shape(void* h, int64_t* rows, int64_t* cols) {                      \
    auto& m = static_cast<LSolver<S>*>(h)->es.eigenvectors();                                                     \
    *rows = m.rows();                                                                                             \
    *cols = m.cols();                                                                                             \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_get_eigenvectors(void* h, void* out) { /* column-major */                        \
    from_vec(static_cast<LSolver<S>*>(h)->es.eigenvectors(), out);                                                \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lsolver_tri_size(void* h) { return static_cast<LSolver<S>*>(h)->es.es_tri().eigenvalues().size(); } \
  REF_API void ref_##P##_lsolver_get_tri(void* h, double* theta, void* s_colmajor) {                              \
    auto& t = static_cast<LSolver<S>*>(h)->es.es_tri();                                                           \
    from_vec(t.eigenvalues(), theta);                                                                             \
    if (s_colmajor && t.eigenvalues().size() > 0) from_vec(t.eigenvectors(), s_colmajor);                         \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lsolver_nlog(void* h) { return static_cast<int64_t>(static_cast<LSolver<S>*>(h)->es.log().size()); } \
  REF_API int ref_##P##_lsolver_get_log(void* h, int64_t i, char* buf, int64_t cap) {                             \
    return copy_string(static_cast<LSolver<S>*>(h)->es.log()[static_cast<std::size_t>(i)], buf, cap);             \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lsolver_has_warn(void* h) { return static_cast<LSolver<S>*>(h)->es.hasWARN(); }       \
  REF_API int64_t ref_##P##_lsolver_has_error(void* h) { return static_cast<LSolver<S>*>(h)->es.hasERROR(); }     \
  REF_API int ref_##P##_lsolver_last_error(void* h, char* buf, int64_t cap) {                                     \
    return copy_string(static_cast<LSolver<S>*>(h)->error, buf, cap);                                             \
  }                                                                                                               \
  REF_API int64_t ref_##P##_lsolver_convlog_len(void* h, int64_t key) { /* -1: key absent */                      \
    auto& m = static_cast<LSolver<S>*>(h)->es.convergenceLog();                                                   \
    auto it = m.find(static_cast<Index>(key));                                                                    \
    return it == m.end() ? -1 : static_cast<int64_t>(it->second.size());                                          \
  }                                                                                                               \
  REF_API void ref_##P##_lsolver_get_convlog(void* h, int64_t key, double* out) {                                 \
    auto& m = static_cast<LSolver<S>*>(h)->es.convergenceLog();                                                   \
    auto it = m.find(static_cast<Index>(key));                                                                    \
    if (it != m.end()) std::copy(it->second.begin(), it->second.end(), out);                                      \
  }                                                                                                               \
  /* ---- LanczosExponentialSolver (lanczos.hpp:1004-1164) ---- */                                                \
  REF_API void ref_##P##_exp_with_eigens(const void* x, const double* eivals, int64_t nev, const void* eivecs,    \
                                         int64_t n, int64_t max_expand, const void* in, void* out) {              \
    using ES = LanczosExponentialSolver<S>;                                                                       \
    typename ES::RealVectorType w = to_vec<double>(eivals, nev);                                                  \
    typename ES::MatrixType y(static_cast<Index>(n), static_cast<Index>(nev));                                    \
    if (n * nev > 0) std::memcpy(static_cast<void*>(y.data()), eivecs, sizeof(S) * static_cast<std::size_t>(n * nev)); \
    Vec<S> vin = to_vec<S>(in, n), vout;                                                                          \
    ES::solveWithEigens(*static_cast<const S*>(x), w, y, static_cast<Index>(max_expand), vin, vout);              \
    from_vec(vout, out);                                                                                          \
  }                                                                                                               \
  REF_API void ref_##P##_exp_with_lanczos(const void* x, void* solver, void* out) {                               \
    Vec<S> vout;                                                                                                  \
    LanczosExponentialSolver<S>::solveWithLanczos(*static_cast<const S*>(x), static_cast<LSolver<S>*>(solver)->es, vout); \
    from_vec(vout, out);                                                                                          \
  }                                                                                                               \
  REF_API void ref_##P##_exp_taylor(int auto_division, const void* x, apply_fn f, void* user, int64_t n,          \
                                    double radius, const void* in, void* out, double error, int64_t max_expansion) { \
    using ES = LanczosExponentialSolver<S>;                                                                       \
    Vec<S> vin = to_vec<S>(in, n), vout;                                                                          \
    auto mm = make_matmul<S>(f, user);                                                                            \
    if (auto_division)                                                                                            \
      ES::solveWithTaylorAutoDivision(*static_cast<const S*>(x), mm, static_cast<Index>(n), radius, vin, vout, error, \
                                      static_cast<Index>(max_expansion));                                         \
    else                                                                                                          \
      ES::solveWithTaylorNoDivision(*static_cast<const S*>(x), mm, static_cast<Index>(n), radius, vin, vout, error, \
                                    static_cast<Index>(max_expansion));                                           \
    from_vec(vout, out);                                                                                          \
  }

REF_COMMON(d, double)
REF_COMMON(z, std::complex<double>)

// ---- ArnoldiEigenSolver<std::complex<double>> (the only Scalar the reference's class compiles for,
//      arnoldi.hpp:857,864) ------------------------------------------------------------------------------
using ZC = std::complex<double>;
REF_API void* ref_z_asolver_create() { return new ASolver(); }
REF_API void ref_z_asolver_destroy(void* h) { delete static_cast<ASolver*>(h); }
REF_API void ref_z_asolver_set_op(void* h, apply_fn f, void* user, int64_t n) {
  static_cast<ASolver*>(h)->es.setMatrixMultiplication(make_matmul<ZC>(f, user), static_cast<Index>(n));
}
REF_API void ref_z_asolver_set_init(void* h, const void* v, int64_t n) {
  static_cast<ASolver*>(h)->es.setInitialVector(to_vec<ZC>(v, n));
}
REF_API void ref_z_asolver_get_init(void* h, void* out) { from_vec(static_cast<ASolver*>(h)->es.initialVector(), out); }
REF_API int64_t ref_z_asolver_init_size(void* h) { return static_cast<ASolver*>(h)->es.initialVector().size(); }
REF_API void ref_z_asolver_clear_ortho(void* h) {
  static_cast<ASolver*>(h)->es.setOrthogonalizingVectors(std::vector<Vec<ZC>>());
}
REF_API void ref_z_asolver_add_ortho(void* h, const void* v, int64_t n) {
  static_cast<ASolver*>(h)->es.refOrthogonalizingVectors().push_back(to_vec<ZC>(v, n));
}
REF_API void ref_z_asolver_set_params(void* h, const void* shift, double threshold, double tolerance, int64_t min_it,
                                      int64_t max_it, int64_t max_eig, int vectors_on) {
  auto& es = static_cast<ASolver*>(h)->es;
  es.setEigenvalueShift(*static_cast<const ZC*>(shift)).setThreshold(threshold).setTolerance(tolerance);
  es.setMinIterations(static_cast<Index>(min_it)).setMaxIterations(static_cast<Index>(max_it));
  es.setMaxEigenvalues(static_cast<Index>(max_eig)).setComputeEigenvectorsOn(vectors_on != 0);
}
REF_API void ref_z_asolver_set_indices(void* h, const int64_t* idx, int64_t count) {
  std::vector<Index> v(idx, idx + count);
  static_cast<ASolver*>(h)->es.setIndicesForConvergence(v);
}
REF_API int ref_z_asolver_compute(void* h) {
  auto* s = static_cast<ASolver*>(h);
  return guarded(s->error, [s] { s->es.compute(); });
}
REF_API int ref_z_asolver_continue(void* h) {
  auto* s = static_cast<ASolver*>(h);
  return guarded(s->error, [s] { s->es.continueToCompute(); });
}
REF_API int64_t ref_z_asolver_iterations(void* h) { return static_cast<ASolver*>(h)->es.iterations(); }
REF_API int64_t ref_z_asolver_nvectors(void* h) { return static_cast<int64_t>(static_cast<ASolver*>(h)->es.arnoldivectors().size()); }
REF_API void ref_z_asolver_get_vector(void* h, int64_t k, void* out) {
  from_vec(static_cast<ASolver*>(h)->es.arnoldivectors()[static_cast<std::size_t>(k)], out);
}
REF_API double ref_z_asolver_residue(void* h) { return static_cast<ASolver*>(h)->es.arnoldiBase().residue(); }
REF_API int64_t ref_z_asolver_hess_size(void* h) { return static_cast<ASolver*>(h)->es.hessenbergMatrix().rows(); }
REF_API void ref_z_asolver_hessenberg(void* h, void* out) { from_vec(static_cast<ASolver*>(h)->es.hessenbergMatrix(), out); }
REF_API int64_t ref_z_asolver_neig(void* h) { return static_cast<ASolver*>(h)->es.eigenvalues().size(); }
REF_API void ref_z_asolver_get_eigenvalues(void* h, void* out) { from_vec(static_cast<ASolver*>(h)->es.eigenvalues(), out); }
REF_API void ref_z_asolver_eigenvectors_shape(void* h, int64_t* rows, int64_t* cols) {
  auto& m = static_cast<ASolver*>(h)->es.eigenvectors();
  *rows = m.rows();
  *cols = m.cols();
}
REF_API void ref_z_asolver_get_eigenvectors(void* h, void* out) { from_vec(static_cast<ASolver*>(h)->es.eigenvectors(), out); }
// eigen-decomposition of the last Hessenberg matrix as the dense solver returned it (unsorted); the reference's
// own accessor for the sorted copy, eigenvectors_h(), does not compile (arnoldi.hpp:666)
REF_API void ref_z_asolver_get_des(void* h, void* eivals, void* eivecs_colmajor) {
  auto& d = static_cast<ASolver*>(h)->es.des();
  from_vec(d.eigenvalues(), eivals);
  if (eivecs_colmajor) from_vec(d.eigenvectors(), eivecs_colmajor);
}
// eigenvectors of the Hessenberg matrix in the solver's order (descending modulus, arnoldi.hpp:813-822)
REF_API int64_t ref_z_asolver_yh_rows(void* h) { return static_cast<ASolver*>(h)->es.sorted_eigenvectors_h().rows(); }
REF_API void ref_z_asolver_get_yh(void* h, void* out) { from_vec(static_cast<ASolver*>(h)->es.sorted_eigenvectors_h(), out); }
REF_API int64_t ref_z_asolver_convlog_len(void* h, int64_t key) {
  auto& m = static_cast<ASolver*>(h)->es.convergence_log();
  auto it = m.find(static_cast<Index>(key));
  return it == m.end() ? -1 : static_cast<int64_t>(it->second.size());
}
REF_API void ref_z_asolver_get_convlog(void* h, int64_t key, void* out) {
  auto& m = static_cast<ASolver*>(h)->es.convergence_log();
  auto it = m.find(static_cast<Index>(key));
  if (it != m.end()) std::copy(it->second.begin(), it->second.end(), static_cast<std::complex<double>*>(out));
}
REF_API int64_t ref_z_asolver_nlog(void* h) { return static_cast<int64_t>(static_cast<ASolver*>(h)->es.log().size()); }
REF_API int ref_z_asolver_get_log(void* h, int64_t i, char* buf, int64_t cap) {
  return copy_string(static_cast<ASolver*>(h)->es.log()[static_cast<std::size_t>(i)], buf, cap);
}
REF_API int64_t ref_z_asolver_has_warn(void* h) { return static_cast<ASolver*>(h)->es.hasWARN(); }
REF_API int ref_z_asolver_last_error(void* h, char* buf, int64_t cap) {
  return copy_string(static_cast<ASolver*>(h)->error, buf, cap);
}
