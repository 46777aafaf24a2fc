// random.hpp — VectorDistribution (reference random.hpp:73-104): the generator behind the default start
// vector (lanczos.hpp:124-135,214-218).  Elements are drawn in index order, then the vector is normalised.
#ifndef CMPT_EIGEN_EX_RANDOM_HPP_
#define CMPT_EIGEN_EX_RANDOM_HPP_

#include "util.hpp"

namespace cmpt {
namespace EigenEx {

template <class Distribution>
class VectorDistribution {
 public:
  using ScalarType = typename Distribution::result_type;
  using result_type = Vector<ScalarType>;
  Distribution dist;
  Index size;  // the reference stores an int (random.hpp:82); Index avoids the 2^31 limit
  bool normalize_on;
  VectorDistribution(const Distribution& dist_ = Distribution(), Index size_ = 0, bool normalize_on_ = true)
      : dist(dist_), size(size_), normalize_on(normalize_on_) {}
  template <class URBG>
  result_type operator()(URBG& g) {
    result_type vec;
    vec.resize(size);
    for (Index i = 0; i < size; ++i) vec[i] = dist(g);
    if (normalize_on) vec.normalize();
    return vec;
  }
};

}  // namespace EigenEx
}  // namespace cmpt
#endif
