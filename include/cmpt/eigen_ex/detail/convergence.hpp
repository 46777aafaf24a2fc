// detail/convergence.hpp — the small bookkeeping both eigen solvers share: which Ritz value an index names, whether the
// watched Ritz values have stopped moving, how many log lines carry a given tag.  Behaviour follows the reference's
// stop rule (lanczos.hpp:837-896, 903-922; arnoldi.hpp:938-996): relative change of every watched value between two
// consecutive trips, measured against the spread of the current Ritz values.
#ifndef CMPT_EIGEN_EX_DETAIL_CONVERGENCE_HPP_
#define CMPT_EIGEN_EX_DETAIL_CONVERGENCE_HPP_

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "dense.hpp"

namespace cmpt {
namespace EigenEx {
namespace detail {

/// Position of entry `i` in a list of `n` values.  Negative `i` counts back from the end (-1 is the last entry).
/// Returns -1 when the list has no such entry.
inline Index wrap_index(Index i, Index n) {
  if (i >= n || i < -n) return -1;
  return i < 0 ? n + i : i;
}

/// One trip of the driver: for every watched index that exists among `values`, append that value to its history.
template <class History, class Watched, class Values>
void record_trip(History& history, const Watched& watched, const Values& values, Index nvalues) {
  for (Index w : watched) {
    const Index at = wrap_index(w, nvalues);
    if (at >= 0) history[w].push_back(values[at]);
  }
}

/// True when every watched index has at least two recorded values and the last two differ by at most
/// `tolerance * spread`.
template <class History, class Watched, class Real>
bool histories_settled(const History& history, const Watched& watched, Real spread, Real tolerance) {
  for (Index w : watched) {
    const auto found = history.find(w);
    if (found == history.end()) return false;
    const auto& trail = found->second;
    const std::size_t len = trail.size();
    if (len < 2) return false;
    if (std::abs((trail[len - 1] - trail[len - 2]) / spread) > tolerance) return false;
  }
  return true;
}

/// Number of log lines that begin with `tag`.
inline Index count_tagged(const std::vector<std::string>& lines, const std::string& tag) {
  return static_cast<Index>(std::count_if(lines.begin(), lines.end(), [&tag](const std::string& line) {
    return line.compare(0, tag.size(), tag) == 0;
  }));
}

}  // namespace detail
}  // namespace EigenEx
}  // namespace cmpt

#endif
