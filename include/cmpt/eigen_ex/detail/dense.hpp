// detail/dense.hpp — value types of the public API.
//
// The reference expresses its API in Eigen types (lanczos.hpp:473-482: VectorType, RealVectorType,
// MatrixType = Eigen::Matrix<Scalar,...>, Index = Eigen::Index).  When Eigen is on the include path the
// same types are used here, so existing user code keeps compiling.  When it is not (this build image has
// no Eigen), a minimal column-major Vector/Matrix with the member names the samples use
// (.data() .size() .rows() .cols() (i,j) [i] .col(j) .norm() .dot()) stands in.
#ifndef CMPT_EIGEN_EX_DETAIL_DENSE_HPP_
#define CMPT_EIGEN_EX_DETAIL_DENSE_HPP_

#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <ostream>
#include <vector>

#if !defined(CMPT_EIGENEX_NO_EIGEN) && defined(__has_include)
#if __has_include("Eigen/Core")
#define CMPT_EIGENEX_HAVE_EIGEN 1
#endif
#endif

#ifdef CMPT_EIGENEX_HAVE_EIGEN
#include "Eigen/Core"
#else
#include "cmpt_b200_solver.h"  // cmb_host_alloc / cmb_host_free for pinned result buffers
#endif

namespace cmpt {
namespace EigenEx {

template <class S>
struct RealOf {
  using type = S;
};
template <class R>
struct RealOf<std::complex<R>> {
  using type = R;
};

namespace detail {
inline double conj_(double a) { return a; }
inline float conj_(float a) { return a; }
template <class R>
inline std::complex<R> conj_(const std::complex<R>& a) {
  return std::conj(a);
}
inline double real_(double a) { return a; }
template <class R>
inline R real_(const std::complex<R>& a) {
  return a.real();
}
}  // namespace detail

#ifdef CMPT_EIGENEX_HAVE_EIGEN

using Index = Eigen::Index;
template <class S>
using Vector = Eigen::Matrix<S, Eigen::Dynamic, 1>;
template <class S>
using Matrix = Eigen::Matrix<S, Eigen::Dynamic, Eigen::Dynamic>;

#else

using Index = std::ptrdiff_t;

namespace detail {
// Contiguous storage with Eigen's resize semantics: growing does NOT initialise the new elements (a
// std::vector would zero-fill 670 MB for five Ritz vectors of cfg 2) and shrinking keeps the allocation.
// Large result buffers can be allocated as pinned host memory (cmb_host_alloc) so device->host copies of
// Ritz vectors run at full PCIe rate.
template <class S>
class Storage {
 public:
  Storage() {}
  explicit Storage(std::size_t n) { resize(n); }
  Storage(std::size_t n, const S& v) {
    resize(n);
    for (std::size_t i = 0; i < n; ++i) p_[i] = v;
  }
  Storage(const S* b, const S* e) {
    resize(static_cast<std::size_t>(e - b));
    if (n_) std::memcpy(static_cast<void*>(p_), b, n_ * sizeof(S));
  }
  Storage(const Storage& o) {
    resize(o.n_);
    if (n_) std::memcpy(static_cast<void*>(p_), o.p_, n_ * sizeof(S));
  }
  Storage(Storage&& o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_), pinned_(o.pinned_) {
    o.p_ = nullptr;
    o.n_ = o.cap_ = 0;
  }
  Storage& operator=(const Storage& o) {
    if (this != &o) {
      resize(o.n_);
      if (n_) std::memcpy(static_cast<void*>(p_), o.p_, n_ * sizeof(S));
    }
    return *this;
  }
  Storage& operator=(Storage&& o) noexcept {
    if (this != &o) {
      release();
      p_ = o.p_;
      n_ = o.n_;
      cap_ = o.cap_;
      pinned_ = o.pinned_;
      o.p_ = nullptr;
      o.n_ = o.cap_ = 0;
    }
    return *this;
  }
  ~Storage() { release(); }
  std::size_t size() const { return n_; }
  S* data() { return p_; }
  const S* data() const { return p_; }
  S& operator[](std::size_t i) { return p_[i]; }
  const S& operator[](std::size_t i) const { return p_[i]; }
  S* begin() { return p_; }
  S* end() { return p_ + n_; }
  const S* begin() const { return p_; }
  const S* end() const { return p_ + n_; }
  // new elements are left uninitialised; `pinned` requests page-locked memory for a (re)allocation
  void resize(std::size_t n, bool pinned = false) {
    if (n > cap_ || (pinned && !pinned_ && n > 0)) {
      S* q = nullptr;
      bool got_pinned = false;
      if (pinned) {
        void* v = nullptr;
        if (cmb_host_alloc(n * sizeof(S), &v) == CMB_OK) {
          q = static_cast<S*>(v);
          got_pinned = true;
        }
      }
      if (!q) q = static_cast<S*>(::operator new(n * sizeof(S)));
      if (n_) std::memcpy(static_cast<void*>(q), p_, (n_ < n ? n_ : n) * sizeof(S));
      release();
      p_ = q;
      cap_ = n;
      pinned_ = got_pinned;
    }
    n_ = n;
  }

 private:
  void release() {
    if (p_) {
      if (pinned_)
        cmb_host_free(p_);
      else
        ::operator delete(p_);
    }
    p_ = nullptr;
    n_ = cap_ = 0;
    pinned_ = false;
  }
  S* p_ = nullptr;
  std::size_t n_ = 0, cap_ = 0;
  bool pinned_ = false;
};
}  // namespace detail

template <class S>
class Vector {
 public:
  using Scalar = S;
  using RealScalar = typename RealOf<S>::type;
  Vector() {}
  explicit Vector(Index n) : d_(static_cast<std::size_t>(n)) {}
  Vector(Index n, const S& v) : d_(static_cast<std::size_t>(n), v) {}
  Vector(const S* p, Index n) : d_(p, p + n) {}
  static Vector Zero(Index n) { return Vector(n, S(0)); }
  Index size() const { return static_cast<Index>(d_.size()); }
  Index rows() const { return size(); }
  Index cols() const { return 1; }
  void resize(Index n) { d_.resize(static_cast<std::size_t>(n)); }
  /// resize into page-locked host memory (falls back to pageable): host->device copies run at PCIe rate
  void resizePinned(Index n) { d_.resize(static_cast<std::size_t>(n), true); }
  S* data() { return d_.data(); }
  const S* data() const { return d_.data(); }
  S& operator[](Index i) { return d_[static_cast<std::size_t>(i)]; }
  const S& operator[](Index i) const { return d_[static_cast<std::size_t>(i)]; }
  S& operator()(Index i) { return d_[static_cast<std::size_t>(i)]; }
  const S& operator()(Index i) const { return d_[static_cast<std::size_t>(i)]; }
  // <this|other>: conjugates *this, as Eigen's dot()
  S dot(const Vector& o) const {
    S s = S(0);
    for (std::size_t i = 0; i < d_.size(); ++i) s += detail::conj_(d_[i]) * o.d_[i];
    return s;
  }
  RealScalar squaredNorm() const {
    RealScalar s = 0;
    for (const S& x : d_) s += std::norm(x);
    return s;
  }
  RealScalar norm() const { return std::sqrt(squaredNorm()); }
  void normalize() {
    RealScalar n = norm();
    if (n > RealScalar(0))
      for (S& x : d_) x /= n;
  }
  Vector normalized() const {
    Vector r(*this);
    r.normalize();
    return r;
  }
  Vector& operator*=(const S& a) {
    for (S& x : d_) x *= a;
    return *this;
  }
  Vector& operator+=(const Vector& o) {
    for (std::size_t i = 0; i < d_.size(); ++i) d_[i] += o.d_[i];
    return *this;
  }
  Vector& operator-=(const Vector& o) {
    for (std::size_t i = 0; i < d_.size(); ++i) d_[i] -= o.d_[i];
    return *this;
  }

 private:
  detail::Storage<S> d_;
};

template <class S>
class Matrix {
 public:
  using Scalar = S;
  Matrix() : r_(0), c_(0) {}
  Matrix(Index r, Index c) : r_(r), c_(c), d_(static_cast<std::size_t>(r * c)) {}
  static Matrix Zero(Index r, Index c) {
    Matrix m(r, c);
    for (S& x : m.d_) x = S(0);
    return m;
  }
  static Matrix Identity(Index r, Index c) {
    Matrix m = Zero(r, c);
    for (Index i = 0; i < (r < c ? r : c); ++i) m(i, i) = S(1);
    return m;
  }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  void resize(Index r, Index c) {
    r_ = r;
    c_ = c;
    d_.resize(static_cast<std::size_t>(r * c));
  }
  /// resize into page-locked host memory (falls back to pageable): device->host copies run at PCIe rate
  void resizePinned(Index r, Index c) {
    r_ = r;
    c_ = c;
    d_.resize(static_cast<std::size_t>(r * c), true);
  }
  S* data() { return d_.data(); }
  const S* data() const { return d_.data(); }
  S& operator()(Index i, Index j) { return d_[static_cast<std::size_t>(j * r_ + i)]; }
  const S& operator()(Index i, Index j) const { return d_[static_cast<std::size_t>(j * r_ + i)]; }
  // column j as a vector (copy)
  Vector<S> col(Index j) const { return Vector<S>(d_.data() + j * r_, r_); }
  void setCol(Index j, const Vector<S>& v) {
    for (Index i = 0; i < r_; ++i) (*this)(i, j) = v[i];
  }

 private:
  Index r_, c_;
  detail::Storage<S> d_;
};

template <class S>
std::ostream& operator<<(std::ostream& os, const Vector<S>& v) {
  for (Index i = 0; i < v.size(); ++i) os << v[i] << (i + 1 < v.size() ? "\n" : "");
  return os;
}
template <class S>
std::ostream& operator<<(std::ostream& os, const Matrix<S>& m) {
  for (Index i = 0; i < m.rows(); ++i) {
    for (Index j = 0; j < m.cols(); ++j) os << m(i, j) << (j + 1 < m.cols() ? " " : "");
    if (i + 1 < m.rows()) os << "\n";
  }
  return os;
}

#endif  // CMPT_EIGENEX_HAVE_EIGEN

namespace detail {
// Copy n scalars into a vector that will be uploaded to the device (start vector): large vectors live in pinned
// memory, and an existing allocation of the right size is reused.
template <class S>
inline void assign_upload(Vector<S>& v, const S* src, Index n) {
#ifdef CMPT_EIGENEX_HAVE_EIGEN
  v.resize(n);
#else
  if (static_cast<std::size_t>(n) * sizeof(S) >= (std::size_t(1) << 20))
    v.resizePinned(n);
  else
    v.resize(n);
#endif
  if (n > 0) std::memcpy(static_cast<void*>(v.data()), src, static_cast<std::size_t>(n) * sizeof(S));
}
// (Re)size a matrix that will receive device results (Ritz vectors): contents are left uninitialised.
template <class S>
inline void resize_result(Matrix<S>& m, Index r, Index c) {
#ifdef CMPT_EIGENEX_HAVE_EIGEN
  m.resize(r, c);
#else
  if (static_cast<std::size_t>(r) * static_cast<std::size_t>(c) * sizeof(S) >= (std::size_t(1) << 20))
    m.resizePinned(r, c);
  else
    m.resize(r, c);
#endif
}
}  // namespace detail

}  // namespace EigenEx
}  // namespace cmpt

#endif
