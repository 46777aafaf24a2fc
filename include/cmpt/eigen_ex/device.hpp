// device.hpp — C++ handles over the C-ABI (include/cmpt_b200.h): context and device-resident operators.
//
// These are the additive part of the drop-in API: next to the reference's
//   setMatrixMultiplication(std::function<void(const Scalar*, Scalar*)>, Index height)   (lanczos.hpp:179-188)
// the solvers accept a DeviceOperator<Scalar>, which keeps the matrix in HBM and lets every Krylov step
// run on the GPU without a host round trip.
#ifndef CMPT_EIGEN_EX_DEVICE_HPP_
#define CMPT_EIGEN_EX_DEVICE_HPP_

#include <complex>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "cmpt_b200.h"
#include "detail/dense.hpp"

namespace cmpt {
namespace EigenEx {

/// exception class used in lanczos (lanczos.hpp:90-95); ArnoldiException aliases it (arnoldi.hpp:45)
class LanczosException : public std::runtime_error {
 public:
  LanczosException(const char* _Message) : runtime_error(_Message) {}
  LanczosException(const std::string& m) : runtime_error(m) {}
};

namespace detail {
inline void check(int rc, const char* what) {
  if (rc != CMB_OK) throw LanczosException(std::string(what) + ": " + cmb_last_error());
}
template <class S>
struct DTypeOf;
template <>
struct DTypeOf<double> {
  static constexpr cmb_dtype value = CMB_F64;
};
template <>
struct DTypeOf<std::complex<double>> {
  static constexpr cmb_dtype value = CMB_C64;
};
}  // namespace detail

/// One GPU (one rank).  The default context uses device $CMPT_B200_DEVICE (default 0).
class DeviceContext {
 public:
  explicit DeviceContext(int device = 0) { detail::check(cmb_ctx_create(device, &ctx_), "cmb_ctx_create"); }
  /// row-partitioned run: one context per rank, nccl_id from cmb_nccl_unique_id() broadcast by the host program
  DeviceContext(int device, int rank, int nranks, const void* nccl_id) {
    detail::check(cmb_ctx_create_dist(device, rank, nranks, nccl_id, &ctx_), "cmb_ctx_create_dist");
  }
  ~DeviceContext() {
    if (owned_) cmb_ctx_destroy(ctx_);
  }
  /// non-owning view of a context created through the C-ABI
  static std::shared_ptr<DeviceContext> borrow(cmb_ctx* c) {
    std::shared_ptr<DeviceContext> p(new DeviceContext(c, false));
    return p;
  }
  DeviceContext(const DeviceContext&) = delete;
  DeviceContext& operator=(const DeviceContext&) = delete;
  cmb_ctx* get() const { return ctx_; }
  int rank() const { return cmb_ctx_rank(ctx_); }
  int nranks() const { return cmb_ctx_nranks(ctx_); }

  static std::shared_ptr<DeviceContext> defaultContext() {
    static std::shared_ptr<DeviceContext> ctx;
    if (!ctx) {
      const char* e = std::getenv("CMPT_B200_DEVICE");
      ctx = std::make_shared<DeviceContext>(e ? std::atoi(e) : 0);
    }
    return ctx;
  }

 private:
  DeviceContext(cmb_ctx* c, bool owned) : ctx_(c), owned_(owned) {}
  cmb_ctx* ctx_ = nullptr;
  bool owned_ = true;
};

/// Operator resident in HBM: the device replacement of the `matmul` callback.
template <class Scalar>
class DeviceOperator {
 public:
  using Index = EigenEx::Index;
  DeviceOperator() {}

  /// CSR rows [rowBegin,rowEnd) of an n x n matrix (global column indices); converted to SELL-32 on the device
  static DeviceOperator fromCSR(Index n, const std::int64_t* rowptr, const std::int32_t* col, const Scalar* val,
                                std::shared_ptr<DeviceContext> ctx = DeviceContext::defaultContext(),
                                Index rowBegin = 0, Index rowEnd = -1) {
    DeviceOperator op;
    op.ctx_ = ctx;
    if (rowEnd < 0) rowEnd = n;
    cmb_op* h = nullptr;
    detail::check(cmb_op_csr_create(ctx->get(), detail::DTypeOf<Scalar>::value, n, rowBegin, rowEnd, rowptr, col, val, &h),
                  "cmb_op_csr_create");
    op.reset(h);
    return op;
  }
  /// dense row-major n x n matrix
  static DeviceOperator fromDenseRowMajor(Index n, const Scalar* a,
                                          std::shared_ptr<DeviceContext> ctx = DeviceContext::defaultContext()) {
    DeviceOperator op;
    op.ctx_ = ctx;
    cmb_op* h = nullptr;
    detail::check(cmb_op_dense_create(ctx->get(), detail::DTypeOf<Scalar>::value, n, 0, n, a, &h), "cmb_op_dense_create");
    op.reset(h);
    return op;
  }
  /// dense column-major matrix (Eigen's default storage): transposed on the host while copying
  static DeviceOperator fromMatrix(const Matrix<Scalar>& m,
                                   std::shared_ptr<DeviceContext> ctx = DeviceContext::defaultContext()) {
    const Index n = m.rows();
    std::vector<Scalar> rm(static_cast<std::size_t>(n) * n);
    for (Index i = 0; i < n; ++i)
      for (Index j = 0; j < n; ++j) rm[static_cast<std::size_t>(i) * n + j] = m(i, j);
    return fromDenseRowMajor(n, rm.data(), ctx);
  }
  /// matrix-free spin-1/2 Heisenberg chain of L sites (2^L states)
  static DeviceOperator heisenbergChain(int L, double J = 1.0, bool pbc = true,
                                        std::shared_ptr<DeviceContext> ctx = DeviceContext::defaultContext()) {
    DeviceOperator op;
    op.ctx_ = ctx;
    cmb_op* h = nullptr;
    detail::check(cmb_op_heisenberg_create(ctx->get(), detail::DTypeOf<Scalar>::value, L, J, pbc ? 1 : 0, &h),
                  "cmb_op_heisenberg_create");
    op.reset(h);
    return op;
  }
  /// legacy: wrap a host callback (reference signature + user pointer)
  static DeviceOperator fromCallback(Index n, cmb_matmul_fn fn, void* user,
                                     std::shared_ptr<DeviceContext> ctx = DeviceContext::defaultContext()) {
    DeviceOperator op;
    op.ctx_ = ctx;
    cmb_op* h = nullptr;
    detail::check(cmb_op_callback_create(ctx->get(), detail::DTypeOf<Scalar>::value, n, fn, user, &h),
                  "cmb_op_callback_create");
    op.reset(h);
    return op;
  }

  /// sum_i coefs[i] * ops[i] as one operator that stays in HBM (device-resident counterpart of VectorMap's operator+,
  /// scalarMultiple: vector_map.hpp:207-245 of the reference); the terms are kept alive by the result
  static DeviceOperator linear(const std::vector<DeviceOperator>& ops, const std::vector<Scalar>& coefs) {
    if (ops.empty() || ops.size() != coefs.size()) throw LanczosException("DeviceOperator::linear: bad argument");
    DeviceOperator op;
    op.ctx_ = ops.front().ctx_;
    std::vector<cmb_op*> raw;
    for (const auto& o : ops) raw.push_back(o.get());
    cmb_op* h = nullptr;
    detail::check(cmb_op_linear_create(op.ctx_->get(), static_cast<std::int64_t>(raw.size()), raw.data(), coefs.data(), &h),
                  "cmb_op_linear_create");
    op.reset(h, ops);
    return op;
  }
  /// outer * inner (inner acts first) as one operator that stays in HBM (vector_map.hpp:247-266 of the reference)
  static DeviceOperator product(const DeviceOperator& outer, const DeviceOperator& inner) {
    if (!outer || !inner) throw LanczosException("DeviceOperator::product: empty operand");
    DeviceOperator op;
    op.ctx_ = inner.ctx_;
    cmb_op* h = nullptr;
    detail::check(cmb_op_product_create(op.ctx_->get(), outer.get(), inner.get(), &h), "cmb_op_product_create");
    op.reset(h, {outer, inner});
    return op;
  }

  /// non-owning view of an operator created through the C-ABI (the caller keeps it alive)
  static DeviceOperator borrow(cmb_op* h) {
    DeviceOperator op;
    op.ctx_ = DeviceContext::borrow(cmb_op_context(h));
    op.op_ = std::shared_ptr<cmb_op>(h, [](cmb_op*) {});
    return op;
  }

  explicit operator bool() const { return static_cast<bool>(op_); }
  cmb_op* get() const { return op_.get(); }
  const std::shared_ptr<DeviceContext>& context() const { return ctx_; }
  Index height() const { return op_ ? cmb_op_height(op_.get()) : 0; }
  Index rows() const { return op_ ? cmb_op_rows(op_.get()) : 0; }
  Index rowBegin() const { return op_ ? cmb_op_row_begin(op_.get()) : 0; }
  /// y = A x on host vectors (local slabs)
  void apply(const Scalar* x, Scalar* y) const { detail::check(cmb_op_apply_host(op_.get(), x, y), "cmb_op_apply_host"); }

 private:
  void reset(cmb_op* h) {
    op_ = std::shared_ptr<cmb_op>(h, [](cmb_op* p) { cmb_op_destroy(p); });
  }
  // a composition holds its children: they are released after the composed operator is destroyed
  void reset(cmb_op* h, const std::vector<DeviceOperator>& children) {
    auto keep = std::make_shared<std::vector<DeviceOperator>>(children);
    op_ = std::shared_ptr<cmb_op>(h, [keep](cmb_op* p) {
      cmb_op_destroy(p);
      keep->clear();
    });
  }
  std::shared_ptr<DeviceContext> ctx_;
  std::shared_ptr<cmb_op> op_;
};

}  // namespace EigenEx
}  // namespace cmpt

#endif
