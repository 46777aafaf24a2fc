// ArnoldiEigenSolver on a dense complex matrix (the reference's sample_arnoldi.cpp shape: n = 50, m = 40, two
// leading eigenpairs) and to full Krylov dimension on a 4x4 (arnoldi_test.cpp).  Prints max|AP - PD|.
#include <complex>
#include <cstdio>
#include <iostream>
#include <random>
#include <vector>

#include "cmpt/eigen_ex/arnoldi.hpp"

using namespace cmpt::EigenEx;
using Scalar = std::complex<double>;

static double max_ap_pd(const std::vector<Scalar>& A, int n, const ArnoldiEigenSolver<Scalar>& es) {
  const auto& P = es.eigenvectors();
  const auto& D = es.eigenvalues();
  double worst = 0.0;
  for (Index c = 0; c < P.cols(); ++c)
    for (int r = 0; r < n; ++r) {
      Scalar ap = 0;
      for (int j = 0; j < n; ++j) ap += A[static_cast<std::size_t>(r) * n + j] * P(j, c);
      worst = std::max(worst, std::abs(ap - P(r, c) * D[c]));
    }
  return worst;
}

int main() {
  std::mt19937 g(12345);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  {
    const int n = 50, m = 40;
    std::vector<Scalar> A(static_cast<std::size_t>(n) * n);
    for (auto& a : A) a = Scalar(u(g), u(g));
    auto matmul = [n, A](Scalar const* in, Scalar* out) {
      for (int r = 0; r < n; ++r) {
        Scalar acc = 0;
        for (int j = 0; j < n; ++j) acc += A[static_cast<std::size_t>(r) * n + j] * in[j];
        out[r] = acc;
      }
    };
    ArnoldiEigenSolver<Scalar> es;
    es.setMatrixMultiplication(matmul, n);
    es.setMaxIterations(m);
    es.setMinIterations(m);
    es.setTolerance(1.0e-14);
    es.setMaxEigenvalues(2);
    es.compute();
    std::printf("m=40 eigenvalues: (%.10f,%.10f) (%.10f,%.10f)\n", es.eigenvalues()[0].real(), es.eigenvalues()[0].imag(),
                es.eigenvalues()[1].real(), es.eigenvalues()[1].imag());
    auto rr = es.ritzResiduals();
    std::printf("m=40 max|AP-PD| = %.3e  ritz residuals %.3e %.3e\n", max_ap_pd(A, n, es), rr[0], rr[1]);
    for (auto& s : es.log()) std::cout << "log: " << s << std::endl;
  }
  {
    const int n = 4;
    std::vector<Scalar> A(static_cast<std::size_t>(n) * n);
    for (auto& a : A) a = Scalar(u(g), u(g));
    ArnoldiEigenSolver<Scalar> aes;
    aes.setMatrixMultiplication(DeviceOperator<Scalar>::fromDenseRowMajor(n, A.data()));
    aes.setThreshold(1.0e-14);
    aes.setEigenvalueShift(0.0);
    aes.setMaxIterations(aes.unlimited);
    aes.setMinIterations(aes.unlimited);
    aes.setInitialVector();
    aes.setMaxEigenvalues(5);
    aes.setTolerance(1.0e-10);
    aes.compute();
    std::printf("full-krylov n=4: %ld eigenvalues, max|AP-PD| = %.3e\n", static_cast<long>(aes.eigenvalues().size()),
                max_ap_pd(A, n, aes));
    for (auto& s : aes.log()) std::cout << "log: " << s << std::endl;
  }
  return 0;
}
