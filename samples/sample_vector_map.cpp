// VectorMap algebra feeding the solver (reference: vector_map.hpp): H = T + 0.5 V is composed
// from two operators that live in HBM and handed to LanczosEigenSolver through function(), exactly as a reference
// user passes any callback.  Checked against the same H assembled as one TripletsMatrix and solved on the device.
#include <cmath>
#include <cstdio>

#include "cmpt/eigen_ex/lanczos.hpp"
#include "cmpt/eigen_ex/triplets_matrix.hpp"
#include "cmpt/eigen_ex/vector_map.hpp"

int main() {
  using namespace cmpt::EigenEx;
  const int n = 400;
  TripletsMatrix<double> T(n, n), V(n, n), H(n, n);
  for (int i = 0; i < n; ++i) {
    const double pot = 0.02 * (i - n / 2) * (i - n / 2) / double(n);
    T.pushBack(i, i, 2.0);
    V.pushBack(i, i, pot);
    H.pushBack(i, i, 2.0 + 0.5 * pot);
    if (i + 1 < n) {
      T.pushBack(i, i + 1, -1.0).pushBack(i + 1, i, -1.0);
      H.pushBack(i, i + 1, -1.0).pushBack(i + 1, i, -1.0);
    }
  }
  VectorMap<double> t, v;
  t.setFromDeviceOperator(T.makeDeviceOperator());
  v.setFromDeviceOperator(V.makeDeviceOperator());
  const VectorMap<double> h = t + v.scalarMultipled(0.5);

  Vector<double> x0(n);
  for (int i = 0; i < n; ++i) x0[i] = std::sin(0.05 * i) + 0.3;
  double ev[3][3];
  if (!h.hasDeviceOperator()) {
    std::printf("FAIL: the composed map lost its device operator\n");
    return 1;
  }
  for (int mode = 0; mode < 3; ++mode) {
    LanczosEigenSolver<double> es;
    if (mode == 0)
      es.setMatrixMultiplication(h.function(), h.sizeIn());
    else if (mode == 1)
      es.setMatrixMultiplication(H.makeDeviceOperator());
    else
      es.setMatrixMultiplication(h.deviceOperator());
    es.setInitialVector(x0).setMinIterations(150).setMaxIterations(150).setMaxEigenvalues(3);
    es.compute();
    for (int k = 0; k < 3; ++k) ev[mode][k] = es.eigenvalues()[k];
    std::printf("%s: %.12f %.12f %.12f\n", mode == 0 ? "vector map" : (mode == 1 ? "assembled " : "device map"), ev[mode][0],
                ev[mode][1], ev[mode][2]);
  }
  double d = 0, dd = 0;
  for (int k = 0; k < 3; ++k) d = std::max(d, std::abs(ev[0][k] - ev[1][k]));
  for (int k = 0; k < 3; ++k) dd = std::max(dd, std::abs(ev[2][k] - ev[1][k]));
  std::printf("max |vector map - assembled| = %.3e\n", d);
  std::printf("max |device map - assembled| = %.3e\n", dd);
  // product on the device: the lowest eigenvalue of T*T is the square of the lowest eigenvalue of T
  const VectorMap<double> tt = t * t;
  double e_t = 0, e_tt = 0;
  for (int mode = 0; mode < 2; ++mode) {
    LanczosEigenSolver<double> es;
    es.setMatrixMultiplication(mode == 0 ? t.deviceOperator() : tt.deviceOperator());
    es.setInitialVector(x0).setMinIterations(399).setMaxIterations(399).setMaxEigenvalues(1).setComputeEigenvectorsOn(false);
    es.compute();
    (mode == 0 ? e_t : e_tt) = es.eigenvalues()[0];
  }
  const double dp = std::abs(e_tt - e_t * e_t);
  std::printf("lowest of T*T on the device %.12e vs (lowest of T)^2 %.12e, difference %.3e\n", e_tt, e_t * e_t, dp);
  const bool ok = d < 1e-10 && dd < 1e-10 && dp < 1e-10 && tt.hasDeviceOperator();
  std::printf("%s\n", ok ? "PASS" : "FAIL");
  return ok ? 0 : 1;
}
