// oracle/krylov_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement of the reference's
// Krylov basis builders (versmc/cmpt-eigenex, include/cmpt/eigen_ex/lanczos.hpp
// and arnoldi.hpp).  It is the checker for the CUDA path and the timed CPU
// baseline; nothing under cmpt_eigenex_b200/ or include/ may include, link or
// call it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs load the library built from it.
//
// PARITY STATUS: "parity unpinned" by the reference's own tests — the reference
// ships no assertions, fixtures or golden vectors, and it
// cannot be compiled here (Eigen3 absent).  The restatement is pinned instead on
// the analytic known answers of SURVEY.md §8(c) (tests/test_oracle_kat.py).
//
// What is restated here (operation order is the reference's):
//   orthogonalize()            lanczos.hpp:143-146, arnoldi.hpp:96-99
//   default start vector       lanczos.hpp:214-218, random.hpp:89-101, util.hpp:76-97,132-148
//   setInitialLanczosvector()  lanczos.hpp:299-323  (arnoldi.hpp:245-269)
//   lanczosStepIsUtmost()      lanczos.hpp:331-347
//   updateLanczosSteps()       lanczos.hpp:371-457
//   arnoldiStepIsUtmost()      arnoldi.hpp:277-288
//   updateArnoldiSteps()       arnoldi.hpp:312-392
//   makeHessenbergMatrix()     arnoldi.hpp:415-432
//   Ritz-vector assembly, normalisation, phase fix   lanczos.hpp:797-817, arnoldi.hpp:841-865
// The driver loops (mainCalculation_) and the dense m x m eigenproblems live in
// oracle/reference_solvers.py (LAPACK through scipy stands in for Eigen).
//
// Level-1 loops carry OpenMP pragmas so the same code can be timed with 1 thread
// (faithful to the reference, which has no threading) or with all host cores.

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using cplx = std::complex<double>;

inline double conj_(double a) { return a; }
inline cplx conj_(const cplx& a) { return std::conj(a); }
inline double real_(double a) { return a; }
inline double real_(const cplx& a) { return a.real(); }

// ---- level-1 kernels (the reference gets these from Eigen: .dot() conjugates
// ---- the left operand, .norm() is the 2-norm) -------------------------------
inline double dot(const double* a, const double* b, int64_t n) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
inline cplx dot(const cplx* a, const cplx* b, int64_t n) {
  double sr = 0.0, si = 0.0;
#pragma omp parallel for reduction(+ : sr, si) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const double ar = a[i].real(), ai = a[i].imag(), br = b[i].real(), bi = b[i].imag();
    sr += ar * br + ai * bi;  // conj(a)*b
    si += ar * bi - ai * br;
  }
  return cplx(sr, si);
}
template <class S>
inline double norm2(const S* a, int64_t n) {
  return std::sqrt(real_(dot(a, a, n)));
}
template <class S>
inline void axpy(S alpha, const S* x, S* y, int64_t n) {  // y += alpha x
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] += alpha * x[i];
}
template <class S>
inline void scal(S alpha, S* x, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] *= alpha;
}

// lanczos.hpp:143-146 / arnoldi.hpp:96-99:  t = <o|w> ; w -= t o
template <class S>
inline S orthogonalize(S* target, const S* ortho, int64_t n) {
  S t = dot(ortho, target, n);
  axpy<S>(-t, ortho, target, n);
  return t;
}

// ---- operators ----------------------------------------------------------------
template <class S>
struct Operator {
  virtual ~Operator() {}
  virtual void apply(const S* in, S* out) const = 0;
};

template <class S>
struct CsrOperator : Operator<S> {
  int64_t n;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<S> val;
  void apply(const S* in, S* out) const override {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
      S acc = S(0);
      for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) acc += val[p] * in[col[p]];
      out[r] = acc;
    }
  }
};

template <class S>
struct DenseOperator : Operator<S> {  // row-major n x n
  int64_t n;
  std::vector<S> a;
  void apply(const S* in, S* out) const override {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
      S acc = S(0);
      const S* row = a.data() + r * n;
      for (int64_t c = 0; c < n; ++c) acc += row[c] * in[c];
      out[r] = acc;
    }
  }
};

// Spin-1/2 Heisenberg chain, H = J * sum_i [SzSz + 1/2 (S+S- + S-S+)]_{i,i+1},
// basis = bit strings (bit i = spin i).  Defined by BASELINE.json cfg 4/5 and
// SURVEY.md §8(d); the reference has no such operator.
template <class S>
struct HeisenbergOperator : Operator<S> {
  int L;
  double J;
  int pbc;
  void apply(const S* in, S* out) const override {
    const int64_t dim = int64_t(1) << L;
    const int nb = (pbc && L > 2) ? L : L - 1;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < dim; ++s) {
      int aligned = 0;
      S off = S(0);
      for (int b = 0; b < nb; ++b) {
        const int i = b, j = (b + 1) % L;
        const int64_t bi = (s >> i) & 1, bj = (s >> j) & 1;
        if (bi == bj) {
          ++aligned;
        } else {
          off += in[s ^ ((int64_t(1) << i) | (int64_t(1) << j))];
        }
      }
      out[s] = S(J * 0.25 * double(aligned - (nb - aligned))) * in[s] + S(0.5 * J) * off;
    }
  }
};

template <class S>
struct CallbackOperator : Operator<S> {
  void (*fn)(const S*, S*, void*);
  void* user;
  void apply(const S* in, S* out) const override { fn(in, out, user); }
};

// ---- default start vector -------------------------------------------------------
// lanczos.hpp:214-218: std::mt19937 with the default seed; random.hpp:89-101 fills
// element by element; util.hpp:132-148 selects std::normal_distribution (real) or
// ComplexNormalDistribution (util.hpp:76-97: real part drawn first, then imaginary);
// then vec.normalize().
inline void draw(std::normal_distribution<double>& d, std::mt19937& g, double& out) { out = d(g); }
inline void draw(std::normal_distribution<double>& d, std::mt19937& g, cplx& out) {
  double re = d(g);
  double im = d(g);
  out = cplx(re, im);
}
template <class S>
void make_random_vector(std::mt19937& g, int64_t n, std::vector<S>& v) {
  std::normal_distribution<double> d;
  v.resize(n);
  for (int64_t i = 0; i < n; ++i) draw(d, g, v[i]);
  double nrm = norm2(v.data(), n);
  if (nrm > 0.0) scal<S>(S(1.0 / nrm), v.data(), n);  // Eigen normalize(): no-op on a zero vector
}

// ---- shared settings/state --------------------------------------------------------
template <class S>
struct KrylovCommon {
  int64_t n = 0;  // matrixHeight_
  Operator<S>* op = nullptr;
  std::vector<std::vector<S>> ortho;  // orthogonalizingVectors_
  std::vector<S> init;                // initialVector_
  double threshold = 1.0e-12;         // DefaultTolerance<double>, lanczos.hpp:62-83
  std::vector<std::vector<S>> basis;
  std::vector<S> v;
  int64_t iterations = 0;

  void set_default_initial_vector() {
    std::mt19937 g;  // default seed, lanczos.hpp:215 / arnoldi.hpp:162-166
    make_random_vector(g, n, init);
  }

  // lanczos.hpp:299-323 / arnoldi.hpp:245-269.  Returns false when no vector was made.
  bool set_initial_basis_vector() {
    if (int64_t(init.size()) != n) set_default_initial_vector();
    basis.resize(1);
    basis[0] = init;
    for (auto& o : ortho) orthogonalize(basis[0].data(), o.data(), n);
    double nrm = norm2(basis[0].data(), n);
    if (nrm < threshold) {
      basis.clear();
      return false;
    }
    scal<S>(S(1.0 / nrm), basis[0].data(), n);
    return true;
  }
};

// ---- Lanczos (lanczos.hpp:104-461) ---------------------------------------------------
template <class S>
struct LanczosO : KrylovCommon<S> {
  double shift = 0.0;    // eigenvalueShift_ (RealScalar)
  int64_t interval = 1;  // reorthogonalizeInterval_
  std::vector<double> alpha, beta;

  void clear_steps() {  // lanczos.hpp:277-283
    this->iterations = 0;
    this->basis.clear();
    alpha.clear();
    beta.clear();
    this->v.clear();
  }

  bool utmost() const {  // lanczos.hpp:331-347
    if (int64_t(this->basis.size()) == this->n) return true;
    if (!beta.empty()) return beta.back() <= this->threshold;
    return false;
  }

  bool step() {  // lanczos.hpp:371-457
    const int64_t n = this->n;
    auto& u = this->basis;
    auto& v = this->v;
    if (n <= 0 || this->op == nullptr) return false;
    if (u.empty()) {
      if (!this->set_initial_basis_vector()) return false;
      v.assign(n, S(0));
      this->op->apply(u[0].data(), v.data());
      if (shift != 0.0) axpy<S>(S(shift), u[0].data(), v.data(), n);
      alpha.push_back(real_(dot(u[0].data(), v.data(), n)));
      return true;  // iterations_ unchanged (lanczos.hpp:397)
    }
    const int64_t k = int64_t(u.size()) - 1;
    u.emplace_back(v);  // w = v
    S* w = u[k + 1].data();
    axpy<S>(S(-alpha[k]), u[k].data(), w, n);                   // lanczos.hpp:404
    if (k > 0) axpy<S>(S(-beta[k - 1]), u[k - 1].data(), w, n);  // lanczos.hpp:407
    if (interval > 0) {                                         // lanczos.hpp:411-426
      const int64_t kmod = (int64_t(u.size()) - 1) % interval;
      for (int64_t kk = kmod, nkk = int64_t(u.size()) - 1; kk < nkk; kk += interval)
        orthogonalize(w, u[kk].data(), n);
      if (kmod == 0)
        for (auto& o : this->ortho) orthogonalize(w, o.data(), n);
    }
    beta.push_back(norm2(w, n));  // lanczos.hpp:429
    if (beta[k] <= this->threshold) {  // lanczos.hpp:433-437: beta entry is kept
      u.pop_back();
      return false;
    }
    scal<S>(S(1.0 / beta[k]), w, n);
    this->op->apply(w, v.data());
    if (shift != 0.0) axpy<S>(S(shift), w, v.data(), n);
    alpha.push_back(real_(dot(w, v.data(), n)));
    ++this->iterations;
    return true;
  }
};

// ---- Arnoldi (arnoldi.hpp:53-438) ------------------------------------------------------
template <class S>
struct ArnoldiO : KrylovCommon<S> {
  S shift = S(0);  // eigenvalueShift_ is Scalar here (arnoldi.hpp:108)
  double residue = 0.0;
  std::vector<std::vector<S>> h;

  void clear_steps() {  // arnoldi.hpp:224-229
    this->iterations = 0;
    this->basis.clear();
    h.clear();
    this->v.clear();
  }

  bool utmost() const {  // arnoldi.hpp:277-288
    if (this->basis.empty()) return false;
    if (int64_t(this->basis.size()) == this->n) return true;
    return residue <= this->threshold;
  }

  bool step() {  // arnoldi.hpp:312-392
    const int64_t n = this->n;
    auto& q = this->basis;
    auto& v = this->v;
    if (n <= 0 || this->op == nullptr) return false;
    if (q.empty()) {
      if (!this->set_initial_basis_vector()) return false;
      v.assign(n, S(0));
      this->op->apply(q[0].data(), v.data());
      if (shift != S(0)) axpy<S>(shift, q[0].data(), v.data(), n);
      for (auto& o : this->ortho) orthogonalize(v.data(), o.data(), n);
      h.resize(1);
      h[0].resize(2);
      h[0][0] = orthogonalize(v.data(), q[0].data(), n);  // arnoldi.hpp:344-345
      h[0][1] = S(0);
      residue = norm2(v.data(), n);
      ++this->iterations;
      return true;
    }
    if (utmost()) return false;
    const int64_t k = int64_t(q.size());
    h[k - 1].resize(k + 1);
    h[k - 1][k] = S(residue);
    q.emplace_back(v);
    scal<S>(S(1.0) / h[k - 1][k], q[k].data(), n);
    this->op->apply(q[k].data(), v.data());
    if (shift != S(0)) axpy<S>(shift, q[k].data(), v.data(), n);
    for (auto& o : this->ortho) orthogonalize(v.data(), o.data(), n);
    h.resize(k + 1);
    h[k].resize(k + 2);
    for (int64_t i = 0; i <= k; ++i) h[k][i] = orthogonalize(v.data(), q[i].data(), n);  // :380-383
    h[k][k + 1] = S(0);
    residue = norm2(v.data(), n);
    ++this->iterations;
    return true;
  }

  int64_t hess_size() const { return std::min<int64_t>(int64_t(h.size()), this->n); }
  // arnoldi.hpp:415-432, column-major out[hs*hs]
  void hessenberg(S* out) const {
    const int64_t hs = hess_size();
    std::fill(out, out + hs * hs, S(0));
    for (int64_t c = 0; c < hs; ++c) {
      const int64_t nr = std::min<int64_t>(hs, int64_t(h[c].size()));
      for (int64_t r = 0; r < nr; ++r) out[c * hs + r] = h[c][r];
    }
  }
};

// Ritz-vector assembly.  lanczos.hpp:797-817: x = sum_m S(m,kk) u_m ; normalise ;
// multiply by 1/phase of the first element with |x_i| > 0.  arnoldi.hpp:841-865 does
// the same with complex coefficients (the Arnoldi variant below takes complex Y and
// writes complex output even for a real basis).
template <class SB, class SC>
void assemble(const std::vector<std::vector<SB>>& basis, int64_t n, const SC* coef, int64_t ldc,
              int64_t ncoef, int64_t nev, SC* out /* n x nev col-major */) {
  for (int64_t kk = 0; kk < nev; ++kk) {
    SC* x = out + kk * n;
    std::fill(x, x + n, SC(0));
    for (int64_t m = 0; m < ncoef; ++m) {
      const SC c = coef[kk * ldc + m];
      const SB* u = basis[m].data();
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) x[i] += c * SC(u[i]);
    }
    SC phase = SC(1.0);
    for (int64_t i = 0; i < n; ++i) {
      const double a = std::abs(x[i]);
      if (a > 0.0) {
        phase = x[i] / a;
        break;
      }
    }
    const double nrm = norm2(x, n);
    const SC f = (nrm > 0.0) ? SC(1.0) / phase * SC(1.0 / nrm) : SC(1.0) / phase;
    scal<SC>(f, x, n);
  }
}

template <class S>
Operator<S>* make_csr(int64_t n, const int64_t* rowptr, const int32_t* col, const S* val) {
  auto* o = new CsrOperator<S>();
  o->n = n;
  o->rowptr.assign(rowptr, rowptr + n + 1);
  o->col.assign(col, col + rowptr[n]);
  o->val.assign(val, val + rowptr[n]);
  return o;
}

}  // namespace

// ---- C interface (ctypes) ----------------------------------------------------------------
#define ORC_API extern "C" __attribute__((visibility("default")))

ORC_API int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_API void orc_set_num_threads(int t) {
#ifdef _OPENMP
  omp_set_num_threads(t);
#else
  (void)t;
#endif
}

#define ORC_INSTANTIATE(P, S)                                                                       \
  ORC_API void* orc_##P##_op_csr(int64_t n, const int64_t* rp, const int32_t* c, const S* v) {      \
    return make_csr<S>(n, rp, c, v);                                                                \
  }                                                                                                 \
  ORC_API void* orc_##P##_op_dense(int64_t n, const S* a) {                                         \
    auto* o = new DenseOperator<S>();                                                               \
    o->n = n;                                                                                       \
    o->a.assign(a, a + n * n);                                                                      \
    return static_cast<Operator<S>*>(o);                                                            \
  }                                                                                                 \
  ORC_API void* orc_##P##_op_heisenberg(int L, double J, int pbc) {                                 \
    auto* o = new HeisenbergOperator<S>();                                                          \
    o->L = L;                                                                                       \
    o->J = J;                                                                                       \
    o->pbc = pbc;                                                                                   \
    return static_cast<Operator<S>*>(o);                                                            \
  }                                                                                                 \
  ORC_API void* orc_##P##_op_callback(void (*fn)(const S*, S*, void*), void* user) {                \
    auto* o = new CallbackOperator<S>();                                                            \
    o->fn = fn;                                                                                     \
    o->user = user;                                                                                 \
    return static_cast<Operator<S>*>(o);                                                            \
  }                                                                                                 \
  ORC_API void orc_##P##_op_apply(void* op, const S* in, S* out) {                                  \
    static_cast<Operator<S>*>(op)->apply(in, out);                                                  \
  }                                                                                                 \
  ORC_API void orc_##P##_op_destroy(void* op) { delete static_cast<Operator<S>*>(op); }             \
  ORC_API void orc_##P##_default_vector(int64_t n, S* out) {                                        \
    std::mt19937 g;                                                                                 \
    std::vector<S> v;                                                                               \
    make_random_vector(g, n, v);                                                                    \
    std::copy(v.begin(), v.end(), out);                                                             \
  }                                                                                                 \
  ORC_API void orc_##P##_seeded_vector(uint32_t seed, int64_t n, S* out) {                          \
    std::mt19937 g(seed);                                                                           \
    std::vector<S> v;                                                                               \
    make_random_vector(g, n, v);                                                                    \
    std::copy(v.begin(), v.end(), out);                                                             \
  }                                                                                                 \
  /* ---- Lanczos ---- */                                                                           \
  ORC_API void* orc_##P##_lanczos_create() { return new LanczosO<S>(); }                            \
  ORC_API void orc_##P##_lanczos_destroy(void* h) { delete static_cast<LanczosO<S>*>(h); }          \
  ORC_API void orc_##P##_lanczos_set_op(void* h, void* op, int64_t n) {                             \
    auto* L = static_cast<LanczosO<S>*>(h);                                                         \
    L->op = static_cast<Operator<S>*>(op);                                                          \
    L->n = n;                                                                                       \
  }                                                                                                 \
  ORC_API void orc_##P##_lanczos_set_params(void* h, double shift, int64_t interval, double thr) {  \
    auto* L = static_cast<LanczosO<S>*>(h);                                                         \
    L->shift = shift;                                                                               \
    L->interval = interval;                                                                         \
    L->threshold = thr;                                                                             \
  }                                                                                                 \
  ORC_API void orc_##P##_lanczos_set_init(void* h, const S* v, int64_t n) {                         \
    static_cast<LanczosO<S>*>(h)->init.assign(v, v + n);                                            \
  }                                                                                                 \
  ORC_API void orc_##P##_lanczos_add_ortho(void* h, const S* v, int64_t n) {                        \
    static_cast<LanczosO<S>*>(h)->ortho.emplace_back(v, v + n);                                     \
  }                                                                                                 \
  ORC_API void orc_##P##_lanczos_clear_steps(void* h) { static_cast<LanczosO<S>*>(h)->clear_steps(); } \
  ORC_API int orc_##P##_lanczos_step(void* h) { return static_cast<LanczosO<S>*>(h)->step() ? 1 : 0; } \
  ORC_API int orc_##P##_lanczos_utmost(void* h) { return static_cast<LanczosO<S>*>(h)->utmost() ? 1 : 0; } \
  ORC_API int64_t orc_##P##_lanczos_iterations(void* h) { return static_cast<LanczosO<S>*>(h)->iterations; } \
  ORC_API int64_t orc_##P##_lanczos_nvectors(void* h) {                                             \
    return int64_t(static_cast<LanczosO<S>*>(h)->basis.size());                                     \
  }                                                                                                 \
  ORC_API int64_t orc_##P##_lanczos_nalpha(void* h) {                                               \
    return int64_t(static_cast<LanczosO<S>*>(h)->alpha.size());                                     \
  }                                                                                                 \
  ORC_API int64_t orc_##P##_lanczos_nbeta(void* h) {                                                \
    return int64_t(static_cast<LanczosO<S>*>(h)->beta.size());                                      \
  }                                                                                                 \
  ORC_API void orc_##P##_lanczos_get_alpha_beta(void* h, double* a, double* b) {                    \
    auto* L = static_cast<LanczosO<S>*>(h);                                                         \
    std::copy(L->alpha.begin(), L->alpha.end(), a);                                                 \
    std::copy(L->beta.begin(), L->beta.end(), b);                                                   \
  }                                                                                                 \
  ORC_API void orc_##P##_lanczos_get_vector(void* h, int64_t k, S* out) {                           \
    auto* L = static_cast<LanczosO<S>*>(h);                                                         \
    std::copy(L->basis[k].begin(), L->basis[k].end(), out);                                         \
  }                                                                                                 \
  ORC_API void orc_##P##_lanczos_assemble(void* h, const S* coef, int64_t ldc, int64_t ncoef,       \
                                          int64_t nev, S* out) {                                    \
    auto* L = static_cast<LanczosO<S>*>(h);                                                         \
    assemble<S, S>(L->basis, L->n, coef, ldc, ncoef, nev, out);                                     \
  }                                                                                                 \
  /* ---- Arnoldi ---- */                                                                           \
  ORC_API void* orc_##P##_arnoldi_create() { return new ArnoldiO<S>(); }                            \
  ORC_API void orc_##P##_arnoldi_destroy(void* h) { delete static_cast<ArnoldiO<S>*>(h); }          \
  ORC_API void orc_##P##_arnoldi_set_op(void* h, void* op, int64_t n) {                             \
    auto* A = static_cast<ArnoldiO<S>*>(h);                                                         \
    A->op = static_cast<Operator<S>*>(op);                                                          \
    A->n = n;                                                                                       \
  }                                                                                                 \
  ORC_API void orc_##P##_arnoldi_set_params(void* h, const S* shift, double thr) {                  \
    auto* A = static_cast<ArnoldiO<S>*>(h);                                                         \
    A->shift = *shift;                                                                              \
    A->threshold = thr;                                                                             \
  }                                                                                                 \
  ORC_API void orc_##P##_arnoldi_set_init(void* h, const S* v, int64_t n) {                         \
    static_cast<ArnoldiO<S>*>(h)->init.assign(v, v + n);                                            \
  }                                                                                                 \
  ORC_API void orc_##P##_arnoldi_add_ortho(void* h, const S* v, int64_t n) {                        \
    static_cast<ArnoldiO<S>*>(h)->ortho.emplace_back(v, v + n);                                     \
  }                                                                                                 \
  ORC_API void orc_##P##_arnoldi_clear_steps(void* h) { static_cast<ArnoldiO<S>*>(h)->clear_steps(); } \
  ORC_API int orc_##P##_arnoldi_step(void* h) { return static_cast<ArnoldiO<S>*>(h)->step() ? 1 : 0; } \
  ORC_API int orc_##P##_arnoldi_utmost(void* h) { return static_cast<ArnoldiO<S>*>(h)->utmost() ? 1 : 0; } \
  ORC_API int64_t orc_##P##_arnoldi_iterations(void* h) { return static_cast<ArnoldiO<S>*>(h)->iterations; } \
  ORC_API int64_t orc_##P##_arnoldi_nvectors(void* h) {                                             \
    return int64_t(static_cast<ArnoldiO<S>*>(h)->basis.size());                                     \
  }                                                                                                 \
  ORC_API double orc_##P##_arnoldi_residue(void* h) { return static_cast<ArnoldiO<S>*>(h)->residue; } \
  ORC_API int64_t orc_##P##_arnoldi_hess_size(void* h) { return static_cast<ArnoldiO<S>*>(h)->hess_size(); } \
  ORC_API void orc_##P##_arnoldi_hessenberg(void* h, S* out) { static_cast<ArnoldiO<S>*>(h)->hessenberg(out); } \
  ORC_API void orc_##P##_arnoldi_get_vector(void* h, int64_t k, S* out) {                           \
    auto* A = static_cast<ArnoldiO<S>*>(h);                                                         \
    std::copy(A->basis[k].begin(), A->basis[k].end(), out);                                         \
  }                                                                                                 \
  ORC_API void orc_##P##_arnoldi_assemble(void* h, const cplx* coef, int64_t ldc, int64_t ncoef,    \
                                          int64_t nev, cplx* out) {                                 \
    auto* A = static_cast<ArnoldiO<S>*>(h);                                                         \
    assemble<S, cplx>(A->basis, A->n, coef, ldc, ncoef, nev, out);                                  \
  }

ORC_INSTANTIATE(d, double)
ORC_INSTANTIATE(z, cplx)
