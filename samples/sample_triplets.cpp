// COO on-ramp: a 1D Laplacian assembled from (deliberately split) triplets, merged with shrink(), spectral range
// from Gershgorin discs, then solved twice — through makeMatMulFunction() (the reference's route, host callback)
// and through makeDeviceOperator() (matrix resident in HBM).  Both must give the same eigenvalues.
#include <cmath>
#include <cstdio>
#include <iostream>

#include "cmpt/eigen_ex/lanczos.hpp"
#include "cmpt/eigen_ex/triplets_matrix.hpp"

int main() {
  using namespace cmpt::EigenEx;
  const int n = 300;
  TripletsMatrix<double> T(n, n);
  for (int i = 0; i < n; ++i) {
    T.pushBack(i, i, 1.25).pushBack(i, i, 0.75);  // duplicates: merged by shrink()
    if (i + 1 < n) {
      T.pushBack(i, i + 1, -1.0);
      T.pushBack(i + 1, i, -1.0);
    }
    T.pushBack(i, (i * 7) % n, 0.0);  // explicit zero: removed by shrink()
  }
  const std::size_t before = T.triplets().size();
  T.shrink();
  auto range = T.estimateEigenvalueRange();
  std::printf("triplets %zu -> %zu, gershgorin range [%.6f, %.6f]\n", before, T.triplets().size(), range[0], range[1]);

  Vector<double> x0(n);
  for (int i = 0; i < n; ++i) x0[i] = std::cos(0.3 * i) + 0.1;
  double ev[2][3];
  for (int mode = 0; mode < 2; ++mode) {
    LanczosEigenSolver<double> es;
    if (mode == 0)
      es.setMatrixMultiplication(T.makeMatMulFunction(), n);
    else
      es.setMatrixMultiplication(T.makeDeviceOperator());
    es.setInitialVector(x0).setMinIterations(120).setMaxIterations(120).setMaxEigenvalues(3);
    es.setEigenvalueShift(-range[1]);  // shift chosen from the Gershgorin bound
    es.compute();
    for (int k = 0; k < 3; ++k) ev[mode][k] = es.eigenvalues()[k];
    std::printf("%s: %.12f %.12f %.12f\n", mode == 0 ? "callback" : "device  ", ev[mode][0], ev[mode][1], ev[mode][2]);
  }
  const double pi = std::acos(-1.0);
  std::printf("exact lowest: %.12f\n", 2.0 - 2.0 * std::cos(pi / (n + 1)));
  double d = 0;
  for (int k = 0; k < 3; ++k) d = std::max(d, std::abs(ev[0][k] - ev[1][k]));
  std::printf("max |callback - device| = %.3e\n", d);
  return 0;
}
