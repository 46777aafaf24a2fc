// triplets_matrix.hpp — COO on-ramp to the device operators.
//
// The reference's TripletsMatrix (triplets_matrix.hpp) is the container its users build sparse operators in and
// hand to the solvers through makeMatMulFunction() (:373-380).  This header keeps the part of that class that
// feeds the Lanczos/Arnoldi path — triplet storage, shrink() (:238-283: sort column-major, merge duplicates, drop
// zeros / small entries), operate() / addOperatedVector() (:314-330), makeMatMulFunction(), Gershgorin discs and
// estimateEigenvalueRange() (:486-523) — and adds makeDeviceOperator(): COO -> CSR -> SELL-32 in HBM.
// The rest of the reference class (dense/sparse Eigen conversions, arithmetic operators, several members that do
// not compile when instantiated) is outside the hot path.
#ifndef CMPT_EIGEN_EX_TRIPLETS_MATRIX_HPP_
#define CMPT_EIGEN_EX_TRIPLETS_MATRIX_HPP_

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <utility>
#include <vector>

#include "device.hpp"

namespace cmpt {
namespace EigenEx {

/// (row, col, value) with the accessor names of Eigen::Triplet
template <class Scalar, class StorageIndex = int>
class Triplet {
 public:
  Triplet() : r_(0), c_(0), v_(0) {}
  Triplet(StorageIndex i, StorageIndex j, const Scalar& v = Scalar(0)) : r_(i), c_(j), v_(v) {}
  StorageIndex row() const { return r_; }
  StorageIndex col() const { return c_; }
  const Scalar& value() const { return v_; }

 private:
  StorageIndex r_, c_;
  Scalar v_;
};

template <class Scalar_>
class TripletsMatrix {
 public:
  using Scalar = Scalar_;
  using RealScalar = typename RealOf<Scalar>::type;
  using Index = int;
  using TripletType = EigenEx::Triplet<Scalar>;
  using Triplets = std::vector<TripletType>;
  using MatMulFunction = std::function<void(Scalar const*, Scalar*)>;

  TripletsMatrix() : rows_(0), cols_(0) {}
  TripletsMatrix(Index rows, Index cols) : rows_(rows), cols_(cols) {}
  TripletsMatrix(Index rows, Index cols, const Triplets& t) : rows_(rows), cols_(cols), triplets_(t) {}
  /// shape fitted to the largest row / column index
  explicit TripletsMatrix(const Triplets& t) : rows_(0), cols_(0), triplets_(t) { fitSize(); }

  Index rows() const { return rows_; }
  Index cols() const { return cols_; }
  const Triplets& triplets() const { return triplets_; }
  Triplets& ref_triplets() { return triplets_; }
  TripletsMatrix& setTriplets(const Triplets& t) {
    triplets_ = t;
    return *this;
  }
  TripletsMatrix& pushBack(Index r, Index c, const Scalar& v) {
    triplets_.push_back(TripletType(r, c, v));
    return *this;
  }
  TripletsMatrix& fitSize() {
    for (const auto& t : triplets_) {
      rows_ = std::max<Index>(rows_, t.row() + 1);
      cols_ = std::max<Index>(cols_, t.col() + 1);
    }
    return *this;
  }
  bool rangeIsInvalid() const {
    return !std::all_of(triplets_.begin(), triplets_.end(), [this](const TripletType& t) {
      return t.row() >= 0 && t.col() >= 0 && t.row() < rows_ && t.col() < cols_;
    });
  }

  /// column-major order, as the reference's default predicate (triplets_matrix.hpp:193-208)
  static bool less_than_for_sort_default(const TripletType& a, const TripletType& b) {
    if (a.col() != b.col()) return a.col() < b.col();
    return a.row() < b.row();
  }
  TripletsMatrix& sort() {
    std::stable_sort(triplets_.begin(), triplets_.end(), less_than_for_sort_default);
    return *this;
  }

  /// 1. sort, 2. add same elements, 3. erase zero terms and terms with |value| < threshold (:238-283)
  TripletsMatrix& shrink(RealScalar threshold = 0.0) {
    sort();
    Triplets merged;
    for (std::size_t i = 0; i < triplets_.size();) {
      std::size_t j = i;
      Scalar c = Scalar(0);
      while (j < triplets_.size() && triplets_[j].row() == triplets_[i].row() && triplets_[j].col() == triplets_[i].col())
        c += triplets_[j++].value();
      if (c != Scalar(0) && !(std::abs(c) < threshold)) merged.push_back(TripletType(triplets_[i].row(), triplets_[i].col(), c));
      i = j;
    }
    triplets_.swap(merged);
    return *this;
  }
  TripletsMatrix shrinked(RealScalar threshold = 0.0) const { return TripletsMatrix(*this).shrink(threshold); }

  /// out += A in  (:314-318)
  void addOperatedVector(Scalar const* in, Scalar* out) const {
    for (const auto& t : triplets_) out[t.row()] += in[t.col()] * t.value();
  }
  /// out = A in  (:324-329)
  void operate(Scalar const* in, Scalar* out) const {
    for (Index r = 0; r < rows_; ++r) out[r] = Scalar(0);
    addOperatedVector(in, out);
  }
  /// the function object the reference's users pass to setMatrixMultiplication (:373-380); copy-captures *this
  MatMulFunction makeMatMulFunction() const {
    TripletsMatrix m(*this);
    return MatMulFunction([m](Scalar const* in, Scalar* out) { m.operate(in, out); });
  }

  /// CSR arrays (duplicates merged, columns ascending within a row)
  void makeCSR(std::vector<std::int64_t>& rowptr, std::vector<std::int32_t>& col, std::vector<Scalar>& val) const {
    Triplets t = shrinked().triplets();
    std::stable_sort(t.begin(), t.end(), [](const TripletType& a, const TripletType& b) {
      if (a.row() != b.row()) return a.row() < b.row();
      return a.col() < b.col();
    });
    rowptr.assign(static_cast<std::size_t>(rows_) + 1, 0);
    col.resize(t.size());
    val.resize(t.size());
    for (std::size_t i = 0; i < t.size(); ++i) {
      rowptr[static_cast<std::size_t>(t[i].row()) + 1]++;
      col[i] = t[i].col();
      val[i] = t[i].value();
    }
    for (Index r = 0; r < rows_; ++r) rowptr[static_cast<std::size_t>(r) + 1] += rowptr[static_cast<std::size_t>(r)];
  }

  /// additive: the operator resident in HBM (COO -> CSR -> SELL-32 on the device); square matrices only
  DeviceOperator<Scalar> makeDeviceOperator(std::shared_ptr<DeviceContext> ctx = DeviceContext::defaultContext()) const {
    if (rows_ != cols_) throw LanczosException("makeDeviceOperator: the matrix must be square");
    if (rangeIsInvalid()) throw LanczosException("makeDeviceOperator: triplet index out of range");
    std::vector<std::int64_t> rowptr;
    std::vector<std::int32_t> col;
    std::vector<Scalar> val;
    makeCSR(rowptr, col, val);
    return DeviceOperator<Scalar>::fromCSR(rows_, rowptr.data(), col.data(), val.data(), ctx);
  }

  /// Gershgorin discs (centre, radius) per row (:486-505)
  std::vector<std::pair<Scalar, RealScalar>> makeGershgorinDiscs() const {
    std::vector<std::pair<Scalar, RealScalar>> discs(static_cast<std::size_t>(rows_),
                                                     std::pair<Scalar, RealScalar>(Scalar(0), RealScalar(0)));
    for (const auto& t : triplets_) {
      if (t.row() == t.col())
        discs[static_cast<std::size_t>(t.row())].first += t.value();
      else
        discs[static_cast<std::size_t>(t.row())].second += std::abs(t.value());
    }
    return discs;
  }
  /// lower and upper bound of the eigenvalues by Gershgorin's theorem (:510-523); useful to pick eigenvalueShift
  std::array<RealScalar, 2> estimateEigenvalueRange() const {
    RealScalar lo = std::numeric_limits<RealScalar>::max();
    RealScalar hi = std::numeric_limits<RealScalar>::lowest();
    for (const auto& d : makeGershgorinDiscs()) {
      lo = std::min(lo, std::real(d.first) - d.second);
      hi = std::max(hi, std::real(d.first) + d.second);
    }
    return std::array<RealScalar, 2>{lo, hi};
  }

 protected:
  Index rows_;
  Index cols_;
  Triplets triplets_;
};

}  // namespace EigenEx
}  // namespace cmpt

#endif
