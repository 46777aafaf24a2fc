// detail/krylov_device.hpp — shared device plumbing of LanczosBase / ArnoldiBase: owns the cmb_krylov
// handle, wraps a legacy std::function operator into a callback cmb_op, uploads start / deflation vectors.
#ifndef CMPT_EIGEN_EX_DETAIL_KRYLOV_DEVICE_HPP_
#define CMPT_EIGEN_EX_DETAIL_KRYLOV_DEVICE_HPP_

#include <cstdint>
#include <exception>
#include <functional>
#include <memory>
#include <vector>

#include "../device.hpp"

namespace cmpt {
namespace EigenEx {
namespace detail {

template <class Scalar>
class KrylovDevice {
 public:
  using MatMulFunction = std::function<void(const Scalar*, Scalar*)>;
  using VectorType = Vector<Scalar>;

  ~KrylovDevice() { release(); }
  KrylovDevice() {}
  // the device state is not shared between copies of a solver: a copy starts without one
  KrylovDevice(const KrylovDevice&) {}
  KrylovDevice& operator=(const KrylovDevice&) {
    release();
    return *this;
  }

  void release() {
    if (k_) cmb_krylov_destroy(k_);
    k_ = nullptr;
    startVersion_ = 0;
    op_ = DeviceOperator<Scalar>();
    bridge_.reset();
  }

  cmb_krylov* handle() const { return k_; }
  cmb_op* op() const { return op_.get(); }
  bool ready() const { return k_ != nullptr; }

  // Bind the operator (device operator if given, else the host callback) and (re)create the Krylov state.
  void prepare(const DeviceOperator<Scalar>& devop, const MatMulFunction& fn, Index n, Index reserve) {
    if (devop) {
      if (!(op_ && op_.get() == devop.get())) {
        // a new operator of the same shape on the same context keeps the allocated Krylov state
        const bool same_shape = k_ && !bridge_ && ctxOf_ == devop.context()->get() && height_ == devop.height() &&
                                rows_ == devop.rows() && rowBegin_ == devop.rowBegin();
        if (!same_shape) release();
        op_ = devop;
        ctxOf_ = devop.context()->get();
        height_ = devop.height();
        rows_ = devop.rows();
        rowBegin_ = devop.rowBegin();
      }
    } else {
      if (!bridge_ || bridge_->n != n || !op_) {
        release();
        bridge_ = std::make_shared<Bridge>();
        bridge_->n = n;
        op_ = DeviceOperator<Scalar>::fromCallback(n, &KrylovDevice::trampoline, bridge_.get());
      }
      bridge_->fn = fn;  // the user may have replaced the function object
    }
    if (!k_) {
      // the operator carries the row range of this rank (all rows on a single-rank context)
      const Index rb = op_.rowBegin();
      check(cmb_krylov_create(op_.context()->get(), DTypeOf<Scalar>::value, op_.height(), rb, rb + op_.rows(),
                              reserve, &k_),
            "cmb_krylov_create");
    }
  }
  void setDeflation(const std::vector<VectorType>& vecs, Index n) {
    if (vecs.empty()) {
      check(cmb_krylov_set_deflation(k_, 0, nullptr, n), "cmb_krylov_set_deflation");
      return;
    }
    std::vector<Scalar> packed(static_cast<std::size_t>(n) * vecs.size());
    for (std::size_t j = 0; j < vecs.size(); ++j) {
      if (vecs[j].size() != n) throw LanczosException("orthogonalizingVectors: size differs from matrixHeight");
      std::copy(vecs[j].data(), vecs[j].data() + n, packed.begin() + j * static_cast<std::size_t>(n));
    }
    check(cmb_krylov_set_deflation(k_, static_cast<std::int64_t>(vecs.size()), packed.data(), n),
          "cmb_krylov_set_deflation");
  }

  /// Starts the chain from `init`.  `version` identifies its contents (the solver bumps it whenever the start vector is
  /// set): a state that already holds that version in HBM starts from its own copy, without a host-to-device transfer.
  void start(const VectorType& init, std::uint64_t version, double threshold, int* status) {
    if (version != 0 && version == startVersion_) {
      check(cmb_krylov_restart(k_, threshold, status), "cmb_krylov_restart");
      return;
    }
    startVersion_ = 0;
    check(cmb_krylov_start(k_, init.data(), threshold, status), "cmb_krylov_start");
    startVersion_ = version;
  }

  void rethrowCallbackError() {
    if (bridge_ && bridge_->err) {
      std::exception_ptr e = bridge_->err;
      bridge_->err = nullptr;
      std::rethrow_exception(e);
    }
  }

 private:
  struct Bridge {
    MatMulFunction fn;
    Index n = 0;
    std::exception_ptr err;
  };
  static void trampoline(const void* in, void* out, void* user) {
    Bridge* b = static_cast<Bridge*>(user);
    try {
      b->fn(static_cast<const Scalar*>(in), static_cast<Scalar*>(out));
    } catch (...) {
      // exceptions must not cross the C-ABI: remember it, hand back zeros, rethrow after the call returns
      b->err = std::current_exception();
      Scalar* o = static_cast<Scalar*>(out);
      for (Index i = 0; i < b->n; ++i) o[i] = Scalar(0);
    }
  }

  cmb_krylov* k_ = nullptr;
  std::uint64_t startVersion_ = 0;  // version of the start vector the device state holds (0: none)
  DeviceOperator<Scalar> op_;
  std::shared_ptr<Bridge> bridge_;
  cmb_ctx* ctxOf_ = nullptr;
  Index height_ = -1, rows_ = -1, rowBegin_ = -1;
};

}  // namespace detail
}  // namespace EigenEx
}  // namespace cmpt

#endif
