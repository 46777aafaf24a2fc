// Smallest Lanczos example on the drop-in API: a 3x3 symmetric matrix through the LEGACY host callback
// (same call sequence as the reference's src/samples/sample_lanczos1.cpp).  Known answer: 2-sqrt(1.5), 2, 2+sqrt(1.5).
#include <cstdio>
#include <iostream>

#include "cmpt/eigen_ex/lanczos.hpp"

int main() {
  using namespace cmpt::EigenEx;
  const int n = 3;
  const double H[3][3] = {{1.0, 0.5, 0.0}, {0.5, 2.0, 0.5}, {0.0, 0.5, 3.0}};
  auto matmul = [&H, n](double const* in, double* out) {
    for (int i = 0; i < n; ++i) {
      out[i] = 0.0;
      for (int j = 0; j < n; ++j) out[i] += H[i][j] * in[j];
    }
  };
  LanczosEigenSolver<double> lanczos;
  lanczos.setMatrixMultiplication(matmul, n);
  lanczos.setTolerance(1.0e-5);
  lanczos.setMaxIterations(100);
  lanczos.compute();

  auto eivals = lanczos.eigenvalues();
  auto eivecs = lanczos.eigenvectors();
  std::printf("eigenvalues:");
  for (Index i = 0; i < eivals.size(); ++i) std::printf(" %.15f", eivals[i]);
  std::printf("\neigenvectors:\n");
  for (Index i = 0; i < eivecs.rows(); ++i) {
    for (Index j = 0; j < eivecs.cols(); ++j) std::printf(" %.12f", eivecs(i, j));
    std::printf("\n");
  }
  for (auto& log : lanczos.log()) std::cout << "log: " << log << std::endl;
  return 0;
}
