// Every setter of LanczosEigenSolver on a complex Hermitian problem (the reference's sample_lanczos2.cpp
// settings): n = 200 chain with -i / +i off-diagonals, spectrum 2cos(k pi/201).  The operator is built as a
// device CSR operator; pass "callback" as argv[1] to run the same matrix through the legacy host callback.
#include <complex>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <random>
#include <vector>

#include "cmpt/eigen_ex/lanczos.hpp"

int main(int argc, char** argv) {
  using namespace cmpt::EigenEx;
  using Scalar = std::complex<double>;
  using Solver = LanczosEigenSolver<Scalar>;
  const int n = 200;
  std::vector<std::int64_t> rowptr(n + 1, 0);
  std::vector<std::int32_t> col;
  std::vector<Scalar> val;
  for (int i = 0; i < n; ++i) {
    if (i > 0) {
      col.push_back(i - 1);
      val.push_back(Scalar(0.0, +1.0));
    }
    if (i < n - 1) {
      col.push_back(i + 1);
      val.push_back(Scalar(0.0, -1.0));
    }
    rowptr[i + 1] = static_cast<std::int64_t>(col.size());
  }
  std::mt19937 random_engine(1);
  Solver es;
  if (argc > 1 && std::strcmp(argv[1], "callback") == 0) {
    es.setMatrixMultiplication(
        [&](Scalar const* in, Scalar* out) {
          for (int r = 0; r < n; ++r) {
            Scalar acc = 0;
            for (std::int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) acc += val[p] * in[col[p]];
            out[r] = acc;
          }
        },
        n);
  } else {
    es.setMatrixMultiplication(DeviceOperator<Scalar>::fromCSR(n, rowptr.data(), col.data(), val.data()));
  }
  es.setEigenvalueShift(0.0);
  es.setTolerance(1.0e-7);
  es.setThreshold(1.0e-14);
  es.setMinIterations(Solver::unlimited);
  es.setMaxIterations(1000);
  es.setComputeEigenvectorsOn(true);
  es.setIndicesForConvergence({0});
  es.setInitialVector(es.lanczosBase().makeRandomVector(random_engine, n));
  es.setMaxEigenvalues(10);
  es.setOrthogonalizingVectors({});
  es.setReorthogonalizeInterval(1);
  es.setReserveSize(128);

  es.compute();

  std::printf("matrix height : %ld\n", static_cast<long>(es.matrixHeight()));
  std::printf("iterations : %ld\n", static_cast<long>(es.iterations()));
  std::printf("subspace rank : %zu\n", es.lanczosvectors().size());
  std::printf("eigenvalues:");
  for (Index i = 0; i < es.eigenvalues().size(); ++i) std::printf(" %.12f", es.eigenvalues()[i]);
  std::printf("\nresiduals:");
  auto rr = es.ritzResiduals();
  for (Index i = 0; i < rr.size(); ++i) std::printf(" %.3e", rr[i]);
  std::printf("\n");
  for (auto& str : es.log()) std::cout << "log: " << str << std::endl;
  std::printf("convergence log entries for index 0: %zu\n", es.convergenceLog().at(0).size());
  return 0;
}
