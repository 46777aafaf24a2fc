// util.hpp — the pieces of the reference's util.hpp that the Lanczos/Arnoldi path uses:
// ComplexNormalDistribution (util.hpp:76-97), NormalDistributionGen (util.hpp:132-148),
// rowwiseShuffle / cwiseShuffle (util.hpp:654-696; NB rowwiseShuffle permutes COLUMNS).
// Everything else in the reference's util.hpp is outside the hot path.
#ifndef CMPT_EIGEN_EX_UTIL_HPP_
#define CMPT_EIGEN_EX_UTIL_HPP_

#include <complex>
#include <random>
#include <type_traits>
#include <vector>

#include "detail/dense.hpp"

namespace cmpt {
namespace EigenEx {

/// complex normal distribution: real part drawn first, then the imaginary part
template <class RealScalarType>
class ComplexNormalDistribution {
 public:
  using ComplexType = std::complex<RealScalarType>;
  using result_type = ComplexType;
  std::normal_distribution<RealScalarType> norm;
  ComplexNormalDistribution(RealScalarType mean = 0.0, RealScalarType stddev = 1.0) : norm(mean, stddev) {}
  template <class URBG>
  ComplexType operator()(URBG& g) {
    RealScalarType real = norm(g);
    RealScalarType imag = norm(g);
    return ComplexType(real, imag);
  }
  void reset() { norm.reset(); }
};

/// The normal distribution that draws a Scalar: std::normal_distribution for real scalars, the complex one above
/// otherwise (util.hpp:132-148 of the reference names the same two types).
template <class Scalar_>
struct NormalDistributionGen {
  using Scalar = Scalar_;
  using Real = typename RealOf<Scalar>::type;
  using Type = typename std::conditional<std::is_same<Scalar, Real>::value, std::normal_distribution<Real>,
                                         ComplexNormalDistribution<Real>>::type;
};

/// permutes the COLUMNS of db: new column c = old column shuffle[c]
template <class MatrixLike, class Indices>
void rowwiseShuffle(MatrixLike& db, const Indices& shuffle) {
  MatrixLike old = db;
  for (Index c = 0, nc = db.cols(); c < nc; ++c)
    for (Index r = 0, nr = db.rows(); r < nr; ++r) db(r, c) = old(r, static_cast<Index>(shuffle[c]));
}

/// element-wise shuffle in storage order: new element i = old element shuffle[i]
template <class VectorLike, class Indices>
void cwiseShuffle(VectorLike& db, const Indices& shuffle) {
  VectorLike old = db;
  for (Index i = 0, ni = db.size(); i < ni; ++i) db.data()[i] = old.data()[static_cast<Index>(shuffle[i])];
}

}  // namespace EigenEx
}  // namespace cmpt
#endif
