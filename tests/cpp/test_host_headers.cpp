// Host-only checks of the header layer (no GPU): start-vector generators (random.hpp / util.hpp), the column / element
// shuffles of util.hpp and the TripletsMatrix container.  Prints machine-readable lines that tests/test_cpp_samples.py
// compares with the CPU oracle (libstdc++ mt19937 + normal_distribution streams) and with numpy.
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "cmpt/eigen_ex/lanczos.hpp"
#include "cmpt/eigen_ex/triplets_matrix.hpp"
#include "cmpt/eigen_ex/util.hpp"

using namespace cmpt::EigenEx;

int main() {
  int bad = 0;
  // ---- start vectors: LanczosBase::makeRandomVector (lanczos.hpp:124-135) ----
  {
    std::mt19937 g(1);
    const Vector<double> x = LanczosBase<double>::makeRandomVector(g, 7);
    std::printf("random_d");
    for (Index i = 0; i < 7; ++i) std::printf(" %.17g", x[i]);
    std::printf("\n");
    std::mt19937 h(1);
    const Vector<std::complex<double>> z = LanczosBase<std::complex<double>>::makeRandomVector(h, 5);
    std::printf("random_z");
    for (Index i = 0; i < 5; ++i) std::printf(" %.17g %.17g", z[i].real(), z[i].imag());
    std::printf("\n");
    double nx = 0, nz = 0;
    for (Index i = 0; i < 7; ++i) nx += x[i] * x[i];
    for (Index i = 0; i < 5; ++i) nz += std::norm(z[i]);
    bad += !(std::abs(nx - 1.0) < 1e-14 && std::abs(nz - 1.0) < 1e-14);  // normalised
  }
  // ---- shuffles (util.hpp:654-696): rowwiseShuffle permutes COLUMNS ----
  {
    Matrix<double> m(2, 3);
    for (Index j = 0; j < 3; ++j)
      for (Index i = 0; i < 2; ++i) m(i, j) = 10.0 * double(i) + double(j);
    const std::vector<std::size_t> perm{2, 0, 1};
    rowwiseShuffle(m, perm);
    bad += !(m(0, 0) == 2.0 && m(1, 0) == 12.0 && m(0, 1) == 0.0 && m(0, 2) == 1.0 && m(1, 2) == 11.0);
    Vector<double> v(3);
    v[0] = 5, v[1] = 6, v[2] = 7;
    cwiseShuffle(v, perm);
    bad += !(v[0] == 7 && v[1] == 5 && v[2] == 6);
  }
  // ---- TripletsMatrix: shrink, operate, makeCSR, Gershgorin (triplets_matrix.hpp:238-283,314-330,486-523) ----
  {
    TripletsMatrix<double> T(4, 4);
    T.pushBack(0, 0, 1.0).pushBack(0, 0, 1.0).pushBack(1, 1, 3.0).pushBack(2, 2, -1.0).pushBack(3, 3, 0.5);
    T.pushBack(0, 1, -0.5).pushBack(1, 0, -0.5).pushBack(2, 3, 2.0).pushBack(3, 2, 2.0);
    T.pushBack(1, 2, 0.0).pushBack(3, 0, 1e-20).pushBack(0, 3, 4.0).pushBack(0, 3, -4.0);
    const std::size_t before = T.triplets().size();
    T.shrink(1e-15);
    bad += !(before == 13 && T.triplets().size() == 8);
    // column-major order after shrink
    for (std::size_t i = 1; i < T.triplets().size(); ++i)
      bad += TripletsMatrix<double>::less_than_for_sort_default(T.triplets()[i], T.triplets()[i - 1]);
    double x[4] = {1, 2, 3, 4}, y[4];
    T.operate(x, y);
    bad += !(y[0] == 2 * 1 - 0.5 * 2 && y[1] == -0.5 * 1 + 3 * 2 && y[2] == -3 + 2 * 4 && y[3] == 2 * 3 + 0.5 * 4);
    auto f = T.makeMatMulFunction();
    double y2[4];
    f(x, y2);
    for (int i = 0; i < 4; ++i) bad += !(y2[i] == y[i]);
    std::vector<std::int64_t> rp;
    std::vector<std::int32_t> col;
    std::vector<double> val;
    T.makeCSR(rp, col, val);
    bad += !(rp.size() == 5 && rp[4] == 8 && col[0] == 0 && col[1] == 1 && val[1] == -0.5 && col[7] == 3 && val[7] == 0.5);
    const auto range = T.estimateEigenvalueRange();
    bad += !(range[0] == -3.0 && range[1] == 3.5);  // rows: [2±0.5], [3±0.5], [-1±2], [0.5±2]
    TripletsMatrix<double> U(std::vector<Triplet<double>>{Triplet<double>(2, 5, 1.0)});
    bad += !(U.rows() == 3 && U.cols() == 6 && !U.rangeIsInvalid());
  }
  std::printf("%s\n", bad == 0 ? "PASS" : "FAIL");
  return bad == 0 ? 0 : 1;
}
