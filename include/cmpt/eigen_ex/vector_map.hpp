// vector_map.hpp — composable linear maps (reference: vector_map.hpp:38-292).
//
// A VectorMap wraps a function void(const Scalar* in, Scalar* out) together with its input and output sizes and
// supports the reference's algebra: (f + g)(x) = f(x) + g(x), (f * g)(x) = f(g(x)), scalar multiples, unary minus,
// composition of a list (applied front to back) and construction from a dense matrix.  The result plugs into the
// solvers exactly as in the reference: es.setMatrixMultiplication(vm.function(), vm.sizeIn()) — the host-callback
// path of this build.  setFromDeviceOperator() (additive) wraps an operator that lives in HBM, so sums and products of
// device operators can be fed to the solvers as well (each application then costs the host round trip of the
// callback path; a single DeviceOperator should be handed to the solver directly).
//
// Device-resident algebra (additive): a map built with setFromDeviceOperator remembers the operator, and sums,
// differences, scalar multiples and products of such maps compose the operators ON THE DEVICE (DeviceOperator::linear /
// ::product, cmb_op_linear_create / cmb_op_product_create): hasDeviceOperator() is then true and
// es.setMatrixMultiplication(vm.deviceOperator()) runs every Krylov step without touching the host.  As soon as a
// host-only map takes part, the result falls back to the callback path.
//
// Independent implementation; the reference's quirks that are bugs are not reproduced: its composition size check
// compares every map with itself (vector_map.hpp:78-88) and operator*= has no return statement (:261-263).
#ifndef CMPT_EIGEN_EX_VECTOR_MAP_HPP_
#define CMPT_EIGEN_EX_VECTOR_MAP_HPP_

#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "detail/dense.hpp"
#include "device.hpp"

namespace cmpt {
namespace EigenEx {

class VectorMapException : public std::runtime_error {
 public:
  explicit VectorMapException(const std::string& m) : std::runtime_error(m) {}
};

template <class Scalar_>
class VectorMap {
 public:
  using Scalar = Scalar_;
  using RealScalar = typename RealOf<Scalar>::type;
  using Index = EigenEx::Index;
  using FunctionType = std::function<void(Scalar const*, Scalar*)>;
  using MatrixType = Matrix<Scalar>;
  using VectorType = Vector<Scalar>;

  VectorMap() : sizeIn_(0), sizeOut_(0) {}

  const FunctionType& function() const { return function_; }
  Index sizeIn() const { return sizeIn_; }
  Index sizeOut() const { return sizeOut_; }

  VectorMap& setFromFunction(const FunctionType& func, Index size_in, Index size_out) {
    function_ = func;
    sizeIn_ = size_in;
    sizeOut_ = size_out;
    device_ = DeviceOperator<Scalar>();  // an arbitrary host function has no device form
    return *this;
  }

  /// additive: true when the whole map is a composition of device operators
  bool hasDeviceOperator() const { return static_cast<bool>(device_); }
  /// additive: the map as one operator in HBM (empty when a host-only map takes part in it)
  const DeviceOperator<Scalar>& deviceOperator() const { return device_; }

  /// the maps act in list order: out = vmaps.back()( ... vmaps.front()(in) ); an empty list is the 0 x 0 identity
  VectorMap& setFromComposition(const std::vector<VectorMap>& vmaps) {
    for (std::size_t i = 0; i + 1 < vmaps.size(); ++i)
      if (vmaps[i].sizeOut() != vmaps[i + 1].sizeIn())
        throw VectorMapException("setFromComposition: sizeOut of a map differs from sizeIn of the next one");
    if (vmaps.empty()) return setFromFunction([](Scalar const*, Scalar*) {}, 0, 0);
    const std::vector<VectorMap> chain(vmaps);
    return setFromFunction(
        [chain](Scalar const* in, Scalar* out) {
          if (chain.size() == 1) {
            chain.front().function()(in, out);
            return;
          }
          std::vector<Scalar> a, b;  // ping-pong buffers for the intermediate vectors
          const Scalar* src = in;
          for (std::size_t i = 0; i < chain.size(); ++i) {
            const bool last = (i + 1 == chain.size());
            std::vector<Scalar>& dstbuf = (i % 2 == 0) ? a : b;
            Scalar* dst = out;
            if (!last) {
              dstbuf.assign(static_cast<std::size_t>(chain[i].sizeOut()), Scalar(0));
              dst = dstbuf.data();
            }
            chain[i].function()(src, dst);
            src = dst;
          }
        },
        vmaps.front().sizeIn(), vmaps.back().sizeOut());
  }

  /// f(x) = A x
  VectorMap& setFromMatrix(const MatrixType& mat) {
    const MatrixType A(mat);
    return setFromFunction(
        [A](Scalar const* in, Scalar* out) {
          const Index r = A.rows(), c = A.cols();
          for (Index i = 0; i < r; ++i) out[i] = Scalar(0);
          for (Index j = 0; j < c; ++j) {
            const Scalar xj = in[j];
            for (Index i = 0; i < r; ++i) out[i] += A(i, j) * xj;
          }
        },
        mat.cols(), mat.rows());
  }

  /// additive: f(x) = Op x for anything with apply(const Scalar*, Scalar*), height() — e.g. DeviceOperator<Scalar>
  template <class Op>
  VectorMap& setFromDeviceOperator(const Op& op) {
    const Op held(op);
    return setFromFunction([held](Scalar const* in, Scalar* out) { held.apply(in, out); }, held.height(), held.height());
  }
  /// the device operator is remembered, so that the algebra below can stay on the device
  VectorMap& setFromDeviceOperator(const DeviceOperator<Scalar>& op) {
    const DeviceOperator<Scalar> held(op);
    setFromFunction([held](Scalar const* in, Scalar* out) { held.apply(in, out); }, held.height(), held.height());
    device_ = held;
    return *this;
  }

  VectorType makeOperated(const VectorType& v_in) const {
    if (static_cast<Index>(v_in.size()) != sizeIn()) throw VectorMapException("makeOperated: v_in.size() != sizeIn()");
    VectorType v_out(sizeOut());
    function_(v_in.data(), v_out.data());
    return v_out;
  }

  /// appends x -> c x after the map; c == 0 replaces the map by the zero map (the inner map is no longer called)
  void scalarMultiple(const Scalar& c) {
    const Index so = sizeOut();
    const DeviceOperator<Scalar> dev = device_ ? DeviceOperator<Scalar>::linear({device_}, {c}) : DeviceOperator<Scalar>();
    scalarMultipleHost_(c);
    device_ = dev;
    (void)so;
  }

 protected:
  void scalarMultipleHost_(const Scalar& c) {
    const Index so = sizeOut();
    if (c == Scalar(0)) {
      setFromFunction(
          [so](Scalar const*, Scalar* out) {
            for (Index i = 0; i < so; ++i) out[i] = Scalar(0);
          },
          sizeIn(), so);
      return;
    }
    const FunctionType inner = function_;
    setFromFunction(
        [inner, c, so](Scalar const* in, Scalar* out) {
          inner(in, out);
          for (Index i = 0; i < so; ++i) out[i] = c * out[i];
        },
        sizeIn(), so);
  }

 public:
  VectorMap scalarMultipled(const Scalar& c) const {
    VectorMap vm(*this);
    vm.scalarMultiple(c);
    return vm;
  }

  VectorMap operator+() const { return *this; }
  VectorMap operator-() const { return scalarMultipled(Scalar(-1)); }

  VectorMap& operator+=(const VectorMap& other) {
    if (sizeIn() != other.sizeIn()) throw VectorMapException("operator+=: sizeIn() != other.sizeIn()");
    if (sizeOut() != other.sizeOut()) throw VectorMapException("operator+=: sizeOut() != other.sizeOut()");
    const FunctionType f = function_, g = other.function_;
    const Index so = sizeOut();
    function_ = [f, g, so](Scalar const* in, Scalar* out) {
      std::vector<Scalar> t(static_cast<std::size_t>(so));
      f(in, out);
      g(in, t.data());
      for (Index i = 0; i < so; ++i) out[i] += t[static_cast<std::size_t>(i)];
    };
    device_ = (device_ && other.device_) ? DeviceOperator<Scalar>::linear({device_, other.device_}, {Scalar(1), Scalar(1)})
                                         : DeviceOperator<Scalar>();
    return *this;
  }
  VectorMap& operator-=(const VectorMap& other) { return *this += (-other); }
  /// (*this) <- (*this) o other, i.e. other acts first
  VectorMap& operator*=(const VectorMap& other) {
    const VectorMap self(*this);
    const DeviceOperator<Scalar> dev = (self.device_ && other.device_) ? DeviceOperator<Scalar>::product(self.device_, other.device_)
                                                                        : DeviceOperator<Scalar>();
    setFromComposition({other, self});
    device_ = dev;
    return *this;
  }

 protected:
  FunctionType function_;
  Index sizeIn_;
  Index sizeOut_;
  DeviceOperator<Scalar> device_;  // non-empty: the whole map as one operator in HBM
};

template <class S>
VectorMap<S> operator+(const VectorMap<S>& a, const VectorMap<S>& b) {
  VectorMap<S> r(a);
  r += b;
  return r;
}
template <class S>
VectorMap<S> operator-(const VectorMap<S>& a, const VectorMap<S>& b) {
  VectorMap<S> r(a);
  r -= b;
  return r;
}
/// (a * b)(x) = a(b(x))
template <class S>
VectorMap<S> operator*(const VectorMap<S>& a, const VectorMap<S>& b) {
  VectorMap<S> r(a);
  r *= b;
  return r;
}

}  // namespace EigenEx
}  // namespace cmpt

#endif
