// Thick-restart Lanczos (additive API, thick_restart.hpp) and the deflation workflow of the reference
// (setOrthogonalizingVectors, lanczos.hpp:167-176) on a 2D Laplacian that does not fit a small Krylov basis:
//  1. the 6 lowest eigenpairs of the Nx x Ny Dirichlet Laplacian with at most 32 basis vectors on the device,
//     checked against the closed form 4 - 2cos(i pi/(Nx+1)) - 2cos(j pi/(Ny+1)) and through ||A x - theta x||;
//  2. a square grid (exactly degenerate pairs): a Lanczos run sees one vector per eigenspace, so the second copy of
//     lambda(1,2) = lambda(2,1) is found by deflating the vectors already converged and running again.
// usage: sample_thick_restart [Nx Ny]
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cmpt/eigen_ex/thick_restart.hpp"
#include "cmpt/eigen_ex/triplets_matrix.hpp"

using namespace cmpt::EigenEx;

static TripletsMatrix<double> laplacian2d(int nx, int ny) {
  TripletsMatrix<double> T(nx * ny, nx * ny);
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      const int r = i * ny + j;
      T.pushBack(r, r, 4.0);
      if (i > 0) T.pushBack(r, r - ny, -1.0);
      if (i + 1 < nx) T.pushBack(r, r + ny, -1.0);
      if (j > 0) T.pushBack(r, r - 1, -1.0);
      if (j + 1 < ny) T.pushBack(r, r + 1, -1.0);
    }
  return T;
}

static std::vector<double> exact_spectrum(int nx, int ny) {
  const double pi = std::acos(-1.0);
  std::vector<double> e;
  for (int i = 1; i <= nx; ++i)
    for (int j = 1; j <= ny; ++j) e.push_back(4.0 - 2.0 * std::cos(i * pi / (nx + 1)) - 2.0 * std::cos(j * pi / (ny + 1)));
  std::sort(e.begin(), e.end());
  return e;
}

static Vector<double> start_vector(int n, unsigned long long seed = 0) {
  Vector<double> x0(n);
  unsigned long long s = 88172645463325252ull + 0x9E3779B97F4A7C15ull * seed;
  for (int i = 0; i < n; ++i) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    x0[i] = double(s >> 11) / 9007199254740992.0 - 0.5;
  }
  return x0;
}

static double residual(const TripletsMatrix<double>& T, const Matrix<double>& X, int col, double theta) {
  const int n = T.rows();
  std::vector<double> x(n), y(n);
  for (int i = 0; i < n; ++i) x[i] = X(i, col);
  T.operate(x.data(), y.data());
  double r = 0;
  for (int i = 0; i < n; ++i) r += (y[i] - theta * x[i]) * (y[i] - theta * x[i]);
  return std::sqrt(r);
}

int main(int argc, char** argv) {
  const int nx = argc > 2 ? std::atoi(argv[1]) : 120, ny = argc > 2 ? std::atoi(argv[2]) : 97;
  bool pass = true;
  {
    TripletsMatrix<double> T = laplacian2d(nx, ny);
    const std::vector<double> exact = exact_spectrum(nx, ny);
    ThickRestartLanczos<double> tr;
    tr.setMatrixMultiplication(T.makeDeviceOperator()).setInitialVector(start_vector(nx * ny));
    tr.setWanted(6).setMaxBasis(32).setTolerance(1e-10).setMaxRestarts(400);
    tr.compute();
    std::printf("rect %dx%d: %ld restarts, %ld operator applications, %ld/%d converged, basis <= %d vectors\n", nx, ny,
                long(tr.restarts()), long(tr.operatorApplications()), long(tr.converged()), 6, 32);
    for (const auto& l : tr.log()) std::printf("  log: %s\n", l.c_str());
    double dmax = 0, rmax = 0;
    for (int k = 0; k < 6; ++k) {
      const double r = residual(T, tr.eigenvectors(), k, tr.eigenvalues()[k]);
      std::printf("  theta[%d] = %.12f  exact %.12f  bound %.2e  ||Ax-theta x|| %.2e\n", k, tr.eigenvalues()[k], exact[k],
                  tr.residuals()[k], r);
      dmax = std::max(dmax, std::abs(tr.eigenvalues()[k] - exact[k]));
      rmax = std::max(rmax, r);
    }
    pass = pass && tr.converged() == 6 && tr.restarts() > 0 && dmax < 1e-9 && rmax < 1e-8;
    std::printf("rect: max |theta - exact| = %.2e, max residual = %.2e\n", dmax, rmax);
  }
  {
    // degenerate pair lambda(1,2) = lambda(2,1) on a square grid through deflation
    const int n1 = 64;
    TripletsMatrix<double> T = laplacian2d(n1, n1);
    const std::vector<double> exact = exact_spectrum(n1, n1);
    DeviceOperator<double> op = T.makeDeviceOperator();
    std::vector<Vector<double>> found;
    std::vector<double> vals;
    for (int round = 0; round < 2; ++round) {
      ThickRestartLanczos<double> tr;
      // a Krylov space holds ONE direction of a degenerate eigenspace (the projection of its start vector), so the
      // second round needs a different start vector: the same one, deflated, has no component left in that eigenspace
      tr.setMatrixMultiplication(op).setInitialVector(start_vector(n1 * n1, round));
      tr.setOrthogonalizingVectors(found);
      tr.setWanted(round == 0 ? 2 : 1).setMaxBasis(30).setTolerance(1e-11).setMaxRestarts(400);
      tr.compute();
      for (Index k = 0; k < Index(tr.eigenvalues().size()); ++k) {
        Vector<double> x(n1 * n1);
        for (int i = 0; i < n1 * n1; ++i) x[i] = tr.eigenvectors()(i, k);
        found.push_back(x);
        vals.push_back(tr.eigenvalues()[k]);
      }
      std::printf("square %dx%d round %d: %ld restarts, %ld/%ld converged\n", n1, n1, round, long(tr.restarts()),
                  long(tr.converged()), long(tr.wanted()));
    }
    // rounds: {lambda_11, one copy of lambda_12}, then the other copy of lambda_12 (not lambda_11, not the first copy)
    std::printf("  found %.12f %.12f | %.12f ; exact %.12f %.12f %.12f\n", vals[0], vals[1], vals[2], exact[0], exact[1],
                exact[2]);
    double dot = 0;
    for (int i = 0; i < n1 * n1; ++i) dot += found[1][i] * found[2][i];
    std::printf("  <x_12 | x_21> = %.2e\n", dot);
    pass = pass && std::abs(vals[0] - exact[0]) < 1e-9 && std::abs(vals[1] - exact[1]) < 1e-9 &&
           std::abs(vals[2] - exact[2]) < 1e-9 && std::abs(dot) < 1e-8;
  }
  {
    // complex Scalar: the Hermitian chain of the reference's sample_lanczos2.cpp (spectrum 2 cos(k pi/(n+1))), four
    // lowest pairs with a basis of 24 vectors
    using C = std::complex<double>;
    const int n = 200;
    std::vector<std::int64_t> rowptr(n + 1, 0);
    std::vector<std::int32_t> col;
    std::vector<C> val;
    for (int i = 0; i < n; ++i) {
      if (i > 0) col.push_back(i - 1), val.push_back(C(0.0, +1.0));
      if (i < n - 1) col.push_back(i + 1), val.push_back(C(0.0, -1.0));
      rowptr[i + 1] = static_cast<std::int64_t>(col.size());
    }
    Vector<C> x0(n);
    for (int i = 0; i < n; ++i) x0[i] = C(std::cos(0.37 * i) + 0.2, std::sin(0.11 * i * i));
    ThickRestartLanczos<C> tr;
    tr.setMatrixMultiplication(DeviceOperator<C>::fromCSR(n, rowptr.data(), col.data(), val.data()));
    tr.setInitialVector(x0).setWanted(4).setMaxBasis(24).setTolerance(1e-11).setMaxRestarts(500);
    tr.compute();
    const double pi = std::acos(-1.0);
    double dmax = 0;
    for (int k = 0; k < 4; ++k) dmax = std::max(dmax, std::abs(tr.eigenvalues()[k] - 2.0 * std::cos((n - k) * pi / (n + 1))));
    std::printf("complex chain n=%d: %ld restarts, %ld/4 converged, max |theta - exact| = %.2e\n", n, long(tr.restarts()),
                long(tr.converged()), dmax);
    pass = pass && tr.converged() == 4 && dmax < 1e-9;
  }
  std::printf("%s\n", pass ? "PASS" : "FAIL");
  return pass ? 0 : 1;
}
