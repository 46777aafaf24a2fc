// Device path end to end in C++: 2D Laplacian N x N as a CSR device operator, Lanczos m iterations with full
// reorthogonalisation, lowest 5 eigenpairs, then continueToCompute() with a larger budget.
// usage: sample_device_csr [N=256] [m=100]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cmpt/eigen_ex/lanczos.hpp"

int main(int argc, char** argv) {
  using namespace cmpt::EigenEx;
  const int N = argc > 1 ? std::atoi(argv[1]) : 256;
  const int m = argc > 2 ? std::atoi(argv[2]) : 100;
  const std::int64_t n = static_cast<std::int64_t>(N) * N;
  std::vector<std::int64_t> rowptr(n + 1, 0);
  std::vector<std::int32_t> col;
  std::vector<double> val;
  col.reserve(5 * n);
  val.reserve(5 * n);
  for (std::int64_t r = 0; r < n; ++r) {
    const std::int64_t i = r / N, j = r % N;
    auto put = [&](std::int64_t c, double v) {
      col.push_back(static_cast<std::int32_t>(c));
      val.push_back(v);
    };
    if (i > 0) put(r - N, -1.0);
    if (j > 0) put(r - 1, -1.0);
    put(r, 4.0);
    if (j < N - 1) put(r + 1, -1.0);
    if (i < N - 1) put(r + N, -1.0);
    rowptr[r + 1] = static_cast<std::int64_t>(col.size());
  }
  Vector<double> x0(n);
  for (std::int64_t i = 0; i < n; ++i) x0[i] = std::sin(0.37 * double(i) + 0.11) + 0.5;
  LanczosEigenSolver<double> es;
  es.setMatrixMultiplication(DeviceOperator<double>::fromCSR(n, rowptr.data(), col.data(), val.data()));
  es.setInitialVector(x0);
  es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5).setIndicesForConvergence({0, 1, 2, 3, 4});
  es.compute();
  const double pi = std::acos(-1.0);
  const double exact0 = 4.0 - 4.0 * std::cos(pi / (N + 1));
  std::printf("n=%ld m=%d iterations=%ld\n", static_cast<long>(n), m, static_cast<long>(es.iterations()));
  std::printf("lowest Ritz value %.12e (exact lowest eigenvalue %.12e)\n", es.eigenvalues()[0], exact0);
  auto rr = es.ritzResiduals();
  std::printf("ritz residuals: %.3e %.3e %.3e\n", rr[0], rr[1], rr[2]);
  std::printf("algorithmic GB moved: %.3f\n", es.lanczosBase().deviceBytes() / 1e9);
  es.setMinIterations(2 * m).setMaxIterations(2 * m);
  es.continueToCompute();
  std::printf("after continueToCompute: iterations=%ld lowest %.12e\n", static_cast<long>(es.iterations()),
              es.eigenvalues()[0]);
  for (auto& s : es.log()) std::printf("log: %s\n", s.c_str());
  return 0;
}
