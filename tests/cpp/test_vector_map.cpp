// Host-only check of include/cmpt/eigen_ex/vector_map.hpp (no GPU needed): the algebra against dense matrices.
#include <cmath>
#include <complex>
#include <cstdio>

#include "cmpt/eigen_ex/vector_map.hpp"

using namespace cmpt::EigenEx;

template <class S>
static Matrix<S> mat(Index r, Index c, unsigned seed) {
  Matrix<S> m(r, c);
  unsigned long long s = 0x9E3779B97F4A7C15ull * (seed + 1);
  for (Index j = 0; j < c; ++j)
    for (Index i = 0; i < r; ++i) {
      s ^= s << 13, s ^= s >> 7, s ^= s << 17;
      m(i, j) = S(double(s >> 11) / 9007199254740992.0 - 0.5);
    }
  return m;
}

template <class S>
static Vector<S> mul(const Matrix<S>& A, const Vector<S>& x) {
  Vector<S> y(A.rows());
  for (Index i = 0; i < A.rows(); ++i) {
    S acc = S(0);
    for (Index j = 0; j < A.cols(); ++j) acc += A(i, j) * x[j];
    y[i] = acc;
  }
  return y;
}

template <class S>
static double diff(const Vector<S>& a, const Vector<S>& b) {
  double d = 0;
  for (Index i = 0; i < Index(a.size()); ++i) d = std::max(d, double(std::abs(a[i] - b[i])));
  return d;
}

template <class S>
static int run() {
  int bad = 0;
  const Matrix<S> A = mat<S>(5, 4, 1), B = mat<S>(5, 4, 2), C = mat<S>(3, 5, 3);
  Vector<S> x(4);
  for (Index i = 0; i < 4; ++i) x[i] = S(0.3 * double(i) - 0.4);
  VectorMap<S> a, b, c;
  a.setFromMatrix(A);
  b.setFromMatrix(B);
  c.setFromMatrix(C);
  bad += !(a.sizeIn() == 4 && a.sizeOut() == 5);
  const Vector<S> ax = mul(A, x), bx = mul(B, x);
  Vector<S> sum(5), dif(5);
  for (Index i = 0; i < 5; ++i) sum[i] = ax[i] + bx[i], dif[i] = ax[i] - bx[i];
  bad += diff((a + b).makeOperated(x), sum) > 1e-14;
  bad += diff((a - b).makeOperated(x), dif) > 1e-14;
  bad += diff((c * a).makeOperated(x), mul(C, ax)) > 1e-14;           // (c * a)(x) = c(a(x)): 4 -> 5 -> 3
  bad += !((c * a).sizeIn() == 4 && (c * a).sizeOut() == 3);
  VectorMap<S> comp;
  comp.setFromComposition({a, c});                                     // list order: a first
  bad += diff(comp.makeOperated(x), mul(C, ax)) > 1e-14;
  Vector<S> m2(5), zero(5);
  for (Index i = 0; i < 5; ++i) m2[i] = S(-2.5) * ax[i], zero[i] = S(0);
  bad += diff(a.scalarMultipled(S(-2.5)).makeOperated(x), m2) > 1e-14;
  bad += diff(a.scalarMultipled(S(0)).makeOperated(x), zero) > 0.0;
  Vector<S> neg(5);
  for (Index i = 0; i < 5; ++i) neg[i] = -ax[i];
  bad += diff((-a).makeOperated(x), neg) > 1e-15;
  int thrown = 0;
  try {
    (void)(a + c);
  } catch (const VectorMapException&) {
    ++thrown;
  }
  try {
    (void)(a * c);  // c: 5 -> 3, a: 4 -> 5: sizes do not chain
  } catch (const VectorMapException&) {
    ++thrown;
  }
  try {
    Vector<S> wrong(7);
    (void)a.makeOperated(wrong);
  } catch (const VectorMapException&) {
    ++thrown;
  }
  bad += thrown != 3;
  // a callback map (what a solver would call): 2 x + A^T(A x) built from pieces
  VectorMap<S> id2;
  id2.setFromFunction([](S const* in, S* out) { for (int i = 0; i < 4; ++i) out[i] = S(2) * in[i]; }, 4, 4);
  Matrix<S> At(4, 5);
  for (Index i = 0; i < 5; ++i)
    for (Index j = 0; j < 4; ++j) At(j, i) = A(i, j);
  VectorMap<S> at;
  at.setFromMatrix(At);
  VectorMap<S> h = id2 + at * a;
  Vector<S> want = mul(At, ax);
  for (Index i = 0; i < 4; ++i) want[i] += S(2) * x[i];
  bad += diff(h.makeOperated(x), want) > 1e-14;
  return bad;
}

int main() {
  const int bad = run<double>() + run<std::complex<double>>();
  std::printf("%s\n", bad == 0 ? "PASS" : "FAIL");
  return bad == 0 ? 0 : 1;
}
