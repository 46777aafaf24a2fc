"""GPU parity tests: the CUDA path (through the C-ABI / solver binding) against the CPU oracle on the same
seeded inputs.  Floating-point path: tolerances follow BASELINE.json's north star — converged/compared
eigenvalues within 1e-10 relative in double, eigenvectors up to sign/phase."""
import numpy as np
import pytest

import cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import capi
from cmpt_eigenex_b200 import synthetic as syn
from oracle import core
from oracle import reference_solvers as rs

pytestmark = pytest.mark.gpu

RTOL_EIG = 1e-10  # north star: eigenvalues within 1e-10 relative
ATOL_AB = 1e-11   # alpha / beta agreement (CGS2 vs the reference's MGS: ~1e-14)


@pytest.fixture(scope="module")
def ctx():
    c = pkg.Context(0)
    yield c
    c.close()


def _csr_dense(rp, c, v, n):
    A = np.zeros((n, n), dtype=v.dtype)
    for r in range(n):
        A[r, c[rp[r]:rp[r + 1]]] = v[rp[r]:rp[r + 1]]
    return A


# ------------------------------------------------------------------------------------------------
# operator apply
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["laplacian", "convdiff", "heisenberg", "hermitian_chain", "ragged"])
def test_csr_apply_matches_oracle(ctx, name):
    if name == "laplacian":
        rp, c, v = syn.laplacian2d_csr(37)
    elif name == "convdiff":
        rp, c, v = syn.convdiff3d_csr(11)
    elif name == "heisenberg":
        rp, c, v = syn.heisenberg_csr(11)
    elif name == "hermitian_chain":
        rp, c, v = syn.hermitian_chain_csr(333)
    else:  # ragged rows incl. empty ones and one long row; n not a multiple of 32
        rng = np.random.default_rng(5)
        n = 1000 + 7
        lens = rng.integers(0, 9, size=n)
        lens[13] = 300
        lens[500:520] = 0
        rp = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=rp[1:])
        c = rng.integers(0, n, size=rp[-1]).astype(np.int32)
        v = rng.normal(size=rp[-1])
    n = rp.size - 1
    x = syn.start_vector(n, seed=3, dtype=v.dtype)
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    y = op.apply(x)
    yr = core.Operator.csr(rp, c, v).apply(x)
    np.testing.assert_allclose(y, yr, rtol=0, atol=1e-14 * max(1.0, np.abs(yr).max()) * 8)
    assert op.bytes == syn.csr_bytes(n, rp[-1], v.dtype.itemsize)


# SELL-32-1024 (CMPT_B200_SELL_SORT=1): rows sorted by length inside 1024-row windows when the natural order would pad
# more than 5 %.  Opt-in: it removes the padding but scatters the gathers (slower on the Heisenberg matrices).
@pytest.mark.parametrize("case", ["heisenberg14", "heisenberg13_open", "ragged_real", "ragged_complex", "short_tail"])
def test_csr_rows_sorted_by_length_inside_windows(ctx, monkeypatch, case):
    rng = np.random.default_rng(17)
    if case == "heisenberg14":
        rp, c, v = syn.heisenberg_csr(14)
    elif case == "heisenberg13_open":
        rp, c, v = syn.heisenberg_csr(13, pbc=False)
    else:
        n = {"ragged_real": 5000 + 3, "ragged_complex": 2048 + 31, "short_tail": 1024 + 1}[case]
        lens = rng.integers(0, 40, size=n)
        lens[rng.integers(0, n, size=5)] = 200
        rp = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=rp[1:])
        c = rng.integers(0, n, size=rp[-1]).astype(np.int32)
        v = rng.normal(size=rp[-1])
        if case == "ragged_complex":
            v = v + 1j * rng.normal(size=rp[-1])
    n = rp.size - 1
    x = syn.start_vector(n, seed=4, dtype=v.dtype)
    yr = core.Operator.csr(rp, c, v).apply(x)
    monkeypatch.setenv("CMPT_B200_SELL_SORT", "1")
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    nnz, padded, srt = op.sell_stats()
    assert nnz == rp[-1] and srt
    y = op.apply(x)
    np.testing.assert_allclose(y, yr, rtol=0, atol=1e-14 * max(1.0, np.abs(yr).max()) * 8)
    a_sorted = None
    if case.startswith("heisenberg"):  # whole solver steps on the sorted operator
        es = pkg.LanczosEigenSolver()
        es.setMatrixMultiplication(op).setInitialVector(x).setMinIterations(8).setMaxIterations(8).setMaxEigenvalues(1)
        es.compute()
        a_sorted = es.alpha().copy()
        es.close()
    op.close()
    monkeypatch.delenv("CMPT_B200_SELL_SORT")
    op2 = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    nnz2, padded2, srt2 = op2.sell_stats()
    assert not srt2 and padded < padded2
    np.testing.assert_allclose(op2.apply(x), yr, rtol=0, atol=1e-14 * max(1.0, np.abs(yr).max()) * 8)
    if a_sorted is not None:  # the same Lanczos coefficients either way
        es = pkg.LanczosEigenSolver()
        es.setMatrixMultiplication(op2).setInitialVector(x).setMinIterations(8).setMaxIterations(8).setMaxEigenvalues(1)
        es.compute()
        np.testing.assert_allclose(es.alpha(), a_sorted, atol=1e-12)
        es.close()
    op2.close()
    if case.startswith("heisenberg"):
        assert padded <= 1.03 * nnz < padded2  # natural order pads ~20 %; the window sort leaves the odd slice per window


def test_csr_stencils_keep_the_natural_row_order(ctx):
    rp, c, v = syn.laplacian2d_csr(64)
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    nnz, padded, srt = op.sell_stats()
    assert not srt and padded <= 1.05 * nnz
    op.close()


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_dense_and_matrix_free_apply(ctx, dtype):
    rng = np.random.default_rng(1)
    # one warp per row (small n), several warps per row with 16-byte loads (even n) and with 8-byte loads (odd n)
    for n in (203, 1030, 1001, 2000):
        A = rng.normal(size=(n, n)).astype(dtype)
        if dtype == np.complex128:
            A = A + 1j * rng.normal(size=(n, n))
        x = syn.start_vector(n, seed=5, dtype=dtype)
        y = pkg.DeviceOperator.from_dense(ctx, A).apply(x)
        np.testing.assert_allclose(y, A @ x, atol=1e-13)
    L = 11
    xs = syn.start_vector(1 << L, seed=9, dtype=dtype)
    for pbc in (True, False):
        yh = pkg.DeviceOperator.heisenberg(ctx, L, 1.0, pbc, dtype=dtype).apply(xs)
        yo = core.Operator.heisenberg(L, 1.0, pbc, prefix="z" if dtype == np.complex128 else "d").apply(xs)
        np.testing.assert_allclose(yh, yo, atol=1e-14)


# ------------------------------------------------------------------------------------------------
# Lanczos
# ------------------------------------------------------------------------------------------------
def _oracle_lanczos(opr, x0, m, nev, prefix="d", **kw):
    ref = rs.LanczosEigenSolver(prefix)
    ref.set_matrix_multiplication(opr)
    ref.init = x0
    ref.min_iterations = ref.max_iterations = m
    ref.max_eigenvalues = nev
    for k, v in kw.items():
        setattr(ref, k, v)
    ref.compute()
    return ref


def _compare_lanczos(es, ref, check_vectors=True):
    a, b = es.alpha(), es.beta()
    ra, rb = ref.alpha_beta()
    assert es.iterations() == ref.iterations
    assert a.shape == ra.shape and b.shape == rb.shape
    np.testing.assert_allclose(a, ra, rtol=0, atol=ATOL_AB * max(1.0, np.abs(ra).max()))
    np.testing.assert_allclose(b, rb, rtol=0, atol=ATOL_AB * max(1.0, np.abs(ra).max()))
    ev, rev = es.eigenvalues(), ref.eigenvalues
    assert ev.shape == rev.shape
    scale = max(np.abs(ra).max(), 1e-300)
    # relative to the eigenvalue itself when it is not tiny compared with ||A||
    tol = RTOL_EIG * np.maximum(np.abs(rev), 1e-3 * scale)
    assert np.all(np.abs(ev - rev) <= tol), (ev, rev)
    if check_vectors and ev.size:
        X, R = es.eigenvectors(), ref.eigenvectors
        assert X.shape == R.shape
        ov = np.abs(np.sum(np.conj(R) * X, axis=0))
        assert np.all(np.abs(ov - 1) < 1e-8), ov
        # phase convention: first non-zero element real positive
        assert np.all(np.abs(X[0].imag) < 1e-12) and np.all(X[0].real > 0)
    assert es.log() == ref.log


def test_sample_lanczos1_kat(ctx):
    # src/samples/sample_lanczos1.cpp through the device dense operator
    H = np.array([[1.0, 0.5, 0.0], [0.5, 2.0, 0.5], [0.0, 0.5, 3.0]])
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(pkg.DeviceOperator.from_dense(ctx, H))
    es.setTolerance(1e-5).setMaxIterations(100)
    es.compute()
    want = np.array([2 - np.sqrt(1.5), 2.0, 2 + np.sqrt(1.5)])
    np.testing.assert_allclose(es.eigenvalues(), want, atol=1e-14)
    X = es.eigenvectors()
    np.testing.assert_allclose(H @ X, X * want, atol=1e-13)
    assert (X[0] > 0).all()
    assert "INFO      lanczos steps achieved full of Krylov subspace" in es.log()
    # default start vector = std::mt19937 default seed: identical to the oracle's restatement
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.dense(H))
    ref.tolerance = 1e-5
    ref.max_iterations = 100
    ref.compute()
    _compare_lanczos(es, ref)


@pytest.mark.parametrize("case", ["dense300", "laplacian48", "heisenberg12", "laplacian_chunked"])
def test_lanczos_fixed_m_matches_oracle(ctx, case):
    if case == "dense300":
        n, m = 300, 60
        A = syn.dense_symmetric(n, seed=1)
        op, opr = pkg.DeviceOperator.from_dense(ctx, A), core.Operator.dense(A)
    elif case == "laplacian48":
        N, m = 48, 100
        n = N * N
        rp, c, v = syn.laplacian2d_csr(N)
        op, opr = pkg.DeviceOperator.from_csr(ctx, rp, c, v), core.Operator.csr(rp, c, v)
    elif case == "heisenberg12":
        L, m = 12, 50
        n = 1 << L
        rp, c, v = syn.heisenberg_csr(L)
        op, opr = pkg.DeviceOperator.from_csr(ctx, rp, c, v), core.Operator.csr(rp, c, v)
    else:  # basis split over several column segments and > 128 columns (multi-chunk CGS2)
        N, m = 40, 150
        n = N * N
        rp, c, v = syn.laplacian2d_csr(N)
        op, opr = pkg.DeviceOperator.from_csr(ctx, rp, c, v), core.Operator.csr(rp, c, v)
    x0 = syn.start_vector(n, seed=7)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0)
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5).setIndicesForConvergence([0, 1, 2, 3, 4])
    if case == "laplacian_chunked":
        es.setReserveSize(24)
    es.compute()
    ref = _oracle_lanczos(opr, x0, m, 5, indices_for_convergence=[0, 1, 2, 3, 4])
    _compare_lanczos(es, ref)
    assert es.hasWARN() == 1
    for i in range(5):
        np.testing.assert_allclose(es.convergenceLog(i), ref.convergence_log[i], atol=1e-9)
    # Ritz residuals: ||A x - theta x|| equals |beta_m S(m,i)| up to rounding
    X = es.eigenvectors()
    for i in range(5):
        r = np.linalg.norm(op.apply(X[:, i]) - es.eigenvalues()[i] * X[:, i])
        assert abs(r - es.ritzResiduals()[i]) < 1e-9
    np.testing.assert_allclose(es.ritzResiduals(), ref.ritz_residuals(), atol=1e-9)
    # orthonormal basis
    V = np.stack([es.basisVector(k) for k in range(0, es.nvectors(), max(1, es.nvectors() // 12))], axis=1)
    np.testing.assert_allclose(V.T @ V, np.eye(V.shape[1]), atol=1e-13)


def test_repeated_compute_starts_from_the_device_copy_of_the_start_vector(ctx):
    # compute() again with the same start vector restarts from the copy kept in HBM (cmb_krylov_restart); setting a
    # new start vector — even one of the same size in the same storage — uploads again
    N, m = 24, 30
    rp, c, v = syn.laplacian2d_csr(N)
    x0, x1 = syn.start_vector(N * N, seed=7), syn.start_vector(N * N, seed=8)
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)

    def solve(es):
        es.compute()
        return es.alpha().copy(), es.beta().copy()

    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(2)
    a0, b0 = solve(es)
    a0b, b0b = solve(es)
    assert np.array_equal(a0, a0b) and np.array_equal(b0, b0b)
    es.setInitialVector(x1)
    a1, b1 = solve(es)
    fresh = pkg.LanczosEigenSolver()
    fresh.setMatrixMultiplication(op).setInitialVector(x1).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(2)
    a1f, b1f = solve(fresh)
    assert np.array_equal(a1, a1f) and np.array_equal(b1, b1f)
    assert not np.array_equal(a0, a1)
    ref = _oracle_lanczos(core.Operator.csr(rp, c, v), x1, m, 2)
    _compare_lanczos(es, ref, check_vectors=False)


_OVERLAP_PROBE = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import synthetic as syn
ctx = pkg.Context(0)
out = []
for N, m in ((24, 60), (96, 40)):
    rp, c, v = syn.laplacian2d_csr(N)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v)).setInitialVector(syn.start_vector(N * N, seed=7))
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(3)
    es.compute()
    out += [es.alpha(), es.beta(), es.eigenvectors().ravel()]
print(np.concatenate(out).tobytes().hex())
"""


def test_launch_overlap_does_not_change_a_single_bit():
    # programmatic dependent launch (an UPDATE pass fetches basis tiles while its predecessor drains) against plain
    # stream order: a kernel that touched data too early would show up as a difference (it did, once)
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env in ({"CMPT_B200_PDL": "0"}, {"CMPT_B200_PDL": "1", "CMPT_B200_PDL_APPLY": "1"}):
        p = subprocess.run([sys.executable, "-c", _OVERLAP_PROBE, root], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, **env))
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(p.stdout.strip().splitlines()[-1])
    assert len(outs[0]) > 1000 and outs[0] == outs[1]


def test_lanczos_complex_sample2(ctx):
    # src/samples/sample_lanczos2.cpp:19-59, exact settings, complex Scalar
    n = 200
    rp, c, v = syn.hermitian_chain_csr(n)
    x0 = core.seeded_vector(1, n, "z")  # makeRandomVector(std::mt19937(1), n)
    es = pkg.LanczosEigenSolver(np.complex128)
    es.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v))
    es.setEigenvalueShift(0.0).setTolerance(1e-7).setThreshold(1e-14)
    es.setMinIterations(es.unlimited).setMaxIterations(1000).setComputeEigenvectorsOn(True)
    es.setIndicesForConvergence([0]).setInitialVector(x0).setMaxEigenvalues(10)
    es.setOrthogonalizingVectors([]).setReorthogonalizeInterval(1).setReserveSize(128)
    es.compute()
    ref = rs.LanczosEigenSolver("z")
    ref.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    ref.tolerance, ref.threshold, ref.min_iterations, ref.max_iterations = 1e-7, 1e-14, -1, 1000
    ref.max_eigenvalues, ref.init = 10, x0
    ref.compute()
    assert abs(es.iterations() - ref.iterations) <= 1  # stop rule may flip by one trip
    if es.iterations() == ref.iterations:
        _compare_lanczos(es, ref)
    assert abs(es.eigenvalues()[0] - 2 * np.cos(200 * np.pi / 201)) < 1e-4
    # full Krylov space: exact spectrum 2cos(k pi/201)
    es2 = pkg.LanczosEigenSolver(np.complex128)
    es2.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v))
    es2.setThreshold(1e-14).setTolerance(0.0).setInitialVector(x0).setComputeEigenvectorsOn(False)
    es2.compute()
    assert es2.nvectors() == n
    np.testing.assert_allclose(es2.eigenvalues(), np.sort(2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))), atol=2e-13)


def test_lanczos_convergence_stop_heisenberg(ctx):
    # cfg 4 semantics at small L: ground state with the reference's stop rule (tolerance 1e-12, index 0)
    L = 12
    rp, c, v = syn.heisenberg_csr(L)
    x0 = syn.start_vector(1 << L, seed=7)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v)).setInitialVector(x0)
    es.setMaxIterations(200).setMaxEigenvalues(1)
    es.compute()
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    ref.init, ref.max_iterations, ref.max_eigenvalues = x0, 200, 1
    ref.compute()
    assert abs(es.iterations() - ref.iterations) <= 1
    assert abs(es.eigenvalues()[0] - syn.HEISENBERG_RING_E0[L]) < 2e-10
    assert abs(es.eigenvalues()[0] - ref.eigenvalues[0]) < RTOL_EIG * abs(ref.eigenvalues[0])
    assert "INFO      lanczos steps converged with tolerance" in es.log()
    # matrix-free operator gives the same run
    es2 = pkg.LanczosEigenSolver()
    es2.setMatrixMultiplication(pkg.DeviceOperator.heisenberg(ctx, L)).setInitialVector(x0)
    es2.setMaxIterations(200).setMaxEigenvalues(1)
    es2.compute()
    assert es2.iterations() == es.iterations()
    np.testing.assert_allclose(es2.alpha(), es.alpha(), atol=1e-12)


@pytest.mark.parametrize("interval", [0, 1, 2, 3, 7])
def test_lanczos_reorthogonalize_interval(ctx, interval):
    N, m = 24, 40
    n = N * N
    rp, c, v = syn.laplacian2d_csr(N)
    x0 = syn.start_vector(n, seed=7)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v)).setInitialVector(x0)
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(3).setReorthogonalizeInterval(interval)
    es.compute()
    ref = _oracle_lanczos(core.Operator.csr(rp, c, v), x0, m, 3, interval=interval)
    a, b = es.alpha(), es.beta()
    ra, rb = ref.alpha_beta()
    # without full reorthogonalisation rounding errors grow with the step count: compare the early steps
    # tightly and the Ritz values loosely
    k = 12 if interval != 1 else m
    np.testing.assert_allclose(a[:k], ra[:k], atol=1e-9)
    np.testing.assert_allclose(b[:k], rb[:k], atol=1e-9)
    assert abs(es.eigenvalues()[0] - ref.eigenvalues[0]) < (1e-10 if interval == 1 else 1e-6)


def test_lanczos_deflation_shift_and_continue(ctx):
    N = 20
    n = N * N
    rp, c, v = syn.laplacian2d_csr(N)
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    opr = core.Operator.csr(rp, c, v)
    x0 = syn.start_vector(n, seed=7)
    lam = syn.laplacian2d_eigenvalues(N, 4)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(80).setMaxIterations(80).setMaxEigenvalues(1)
    es.compute()
    g = es.eigenvectors()[:, 0].copy()
    assert abs(es.eigenvalues()[0] - lam[0]) < 1e-12
    # deflate the ground state + shift: lowest Ritz value becomes the first excited level
    es2 = pkg.LanczosEigenSolver()
    es2.setMatrixMultiplication(op).setInitialVector(x0).setOrthogonalizingVectors([g]).setEigenvalueShift(1.5)
    es2.setMinIterations(80).setMaxIterations(80).setMaxEigenvalues(1)
    es2.compute()
    ref2 = _oracle_lanczos(opr, x0, 80, 1, ortho=[g], shift=1.5)
    assert abs(es2.eigenvalues()[0] - lam[1]) < 1e-10
    assert abs(es2.eigenvalues()[0] - ref2.eigenvalues[0]) < 1e-10
    assert abs(g @ es2.eigenvectors()[:, 0]) < 1e-12
    np.testing.assert_allclose(es2.alpha(), ref2.alpha_beta()[0], atol=1e-10)
    # continueToCompute extends the same Krylov space (lanczos.hpp:701-712)
    es3 = pkg.LanczosEigenSolver()
    es3.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(30).setMaxIterations(30).setMaxEigenvalues(2)
    es3.compute()
    es3.setMinIterations(80).setMaxIterations(80)
    es3.continueToCompute()
    ref3 = _oracle_lanczos(opr, x0, 30, 2)
    ref3.min_iterations = ref3.max_iterations = 80
    ref3.continue_to_compute()
    assert es3.iterations() == 80 == ref3.iterations
    np.testing.assert_allclose(es3.alpha(), es.alpha(), atol=1e-12)
    np.testing.assert_allclose(es3.eigenvalues(), ref3.eigenvalues, atol=1e-10)
    assert es3.log() == ref3.log


def test_lanczos_edge_cases(ctx):
    H = np.diag([1.0, 2.0, 3.0, 4.0])
    op = pkg.DeviceOperator.from_dense(ctx, H)
    # zero start vector -> quiet failure (lanczos.hpp:316-318,748-752)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(np.zeros(4))
    es.compute()
    assert es.nvectors() == 0 and es.eigenvalues().size == 0
    assert "INFO      initial lanczosvector generation fail" in es.log()
    # start vector inside an invariant subspace -> breakdown after 2 vectors, beta kept (lanczos.hpp:433-437)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(np.array([1.0, 1.0, 0.0, 0.0]))
    es.compute()
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.dense(H))
    ref.init = np.array([1.0, 1.0, 0.0, 0.0])
    ref.compute()
    assert es.nvectors() == 2 == ref.base.nvectors
    assert es.beta().size == es.alpha().size == 2
    np.testing.assert_allclose(es.eigenvalues(), [1.0, 2.0], atol=1e-14)
    assert es.log() == ref.log
    # maxEigenvalues unlimited: every Ritz vector is assembled (n x (m+1))
    N = 12
    rp, c, v = syn.laplacian2d_csr(N)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v)).setInitialVector(syn.start_vector(N * N))
    es.setMinIterations(9).setMaxIterations(9)
    es.compute()
    X = es.eigenvectors()
    assert X.shape == (N * N, 10)
    np.testing.assert_allclose(X.T @ X, np.eye(10), atol=1e-12)
    # compute without eigenvectors
    es.setComputeEigenvectorsOn(False)
    es.compute()
    assert es.eigenvectors().size == 0 and es.eigenvalues().size == 10
    # wrong-size start vector is replaced by the default random one, with the INFO line
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(np.ones(3))
    es.compute()
    assert "INFO      in compute(), initial_vector is empty or invalid, then set at random" in es.log()
    np.testing.assert_allclose(es.eigenvalues(), [1, 2, 3, 4], atol=1e-13)


def test_legacy_callback_operator(ctx):
    # the reference's own plug-in: a host function (const Scalar*, Scalar*) (lanczos.hpp:116)
    N = 16
    n = N * N
    rp, c, v = syn.laplacian2d_csr(N)
    A = _csr_dense(rp, c, v, n)
    x0 = syn.start_vector(n, seed=7)
    calls = []

    def matmul(x):
        calls.append(1)
        return A @ x

    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(matmul, n).setInitialVector(x0).setMinIterations(30).setMaxIterations(30)
    es.setMaxEigenvalues(2).setEigenvalueShift(0.25)
    es.compute()
    ref = _oracle_lanczos(core.Operator.csr(rp, c, v), x0, 30, 2, shift=0.25)
    assert len(calls) == 31  # m+1 operator applies
    _compare_lanczos(es, ref)


# ------------------------------------------------------------------------------------------------
# Arnoldi
# ------------------------------------------------------------------------------------------------
def _sorted_close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    d = np.abs(a[:, None] - b[None, :])
    return d.min(axis=1).max() < tol and d.min(axis=0).max() < tol


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_arnoldi_convdiff_matches_oracle(ctx, dtype):
    M, m = 10, 50
    n = M ** 3
    rp, c, v = syn.convdiff3d_csr(M)
    v = v.astype(dtype)
    x0 = syn.start_vector(n, seed=7, dtype=dtype)
    es = pkg.ArnoldiEigenSolver(dtype)
    es.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v)).setInitialVector(x0)
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5).setIndicesForConvergence([0, 1, 2])
    es.compute()
    ref = rs.ArnoldiEigenSolver("z" if dtype == np.complex128 else "d")
    ref.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    ref.init = x0
    ref.min_iterations = ref.max_iterations = m
    ref.max_eigenvalues = 5
    ref.indices_for_convergence = [0, 1, 2]
    ref.compute()
    assert es.iterations() == m == ref.iterations
    # The Arnoldi recurrence on this non-normal operator with near-degenerate eigenvalues amplifies rounding
    # differences by ~10x every 4-5 steps (two CPU MGS runs with different summation order already differ by
    # 1e-4 in the late Hessenberg columns): compare the early columns and the stable leading Ritz value
    # tightly, the sensitive rest loosely, and check the exact identities on the GPU result itself.
    H, Hr = es.hessenbergMatrix(), ref.hessenberg
    np.testing.assert_allclose(H[:, :8], Hr[:, :8], atol=1e-11)
    np.testing.assert_allclose(H, Hr, atol=5e-3)
    assert abs(es.residue() - ref.base.residue) < 5e-3
    ev, rev = es.eigenvalues(), ref.eigenvalues
    assert abs(ev[0] - rev[0]) < RTOL_EIG * abs(rev[0])
    assert _sorted_close(ev, rev, 1e-5 * np.abs(rev).max())
    A = _csr_dense(rp, c, v, n)
    P, D = es.eigenvectors(), es.eigenvalues()
    res = np.linalg.norm(A @ P - P * D, axis=0)
    assert np.all(np.abs(res - es.ritzResiduals()) < 1e-9)  # ||A x - theta x|| = residue |Y(last,i)|
    np.testing.assert_allclose(np.linalg.norm(P, axis=0), 1.0, atol=1e-12)
    assert np.all(np.abs(P[0].imag) < 1e-12) and np.all(P[0].real > 0)
    Q = np.stack([es.basisVector(k) for k in range(m)], axis=1)
    np.testing.assert_allclose(Q.conj().T @ Q, np.eye(m), atol=1e-13)           # orthonormal basis
    np.testing.assert_allclose(Q.conj().T @ (A @ Q), H, atol=1e-12)             # H = Q^H A Q
    ov = np.abs(np.sum(np.conj(ref.eigenvectors[:, :1]) * P[:, :1], axis=0))
    assert np.all(np.abs(ov - 1) < 1e-7)
    assert es.log() == ref.log


def test_arnoldi_converged_eigenvalues_match_oracle(ctx):
    # non-symmetric operator with well separated dominant eigenvalues: the wanted Ritz values converge and
    # then agree with the oracle to the north-star tolerance (1e-10 relative)
    rng = np.random.default_rng(11)
    n, m = 2000, 50
    d = np.concatenate([[40.0, 35.0, 31.0, 28.0, 26.0], rng.uniform(0, 10, n - 5)])
    rows = np.repeat(np.arange(n), 4)
    cols = rng.integers(0, n, size=4 * n)
    vals = 0.05 * rng.normal(size=4 * n)
    import scipy.sparse as sp

    A = (sp.diags(d) + sp.csr_matrix((vals, (rows, cols)), shape=(n, n))).tocsr()
    A.sort_indices()
    rp, c, v = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data
    x0 = syn.start_vector(n, seed=7)
    es = pkg.ArnoldiEigenSolver(np.float64)
    es.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v)).setInitialVector(x0)
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5).setIndicesForConvergence([0, 1, 2, 3, 4])
    es.compute()
    ref = rs.ArnoldiEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    ref.init = x0
    ref.min_iterations = ref.max_iterations = m
    ref.max_eigenvalues = 5
    ref.indices_for_convergence = [0, 1, 2, 3, 4]
    ref.compute()
    ev, rev = es.eigenvalues(), ref.eigenvalues
    assert np.all(es.ritzResiduals() < 1e-9)
    assert np.all(np.abs(ev - rev) <= RTOL_EIG * np.abs(rev)), (ev, rev)
    P = es.eigenvectors()
    ov = np.abs(np.sum(np.conj(ref.eigenvectors) * P, axis=0))
    assert np.all(np.abs(ov - 1) < 1e-8)
    for i in range(5):
        np.testing.assert_allclose(es.convergenceLog(i)[-3:], np.array(ref.convergence_log[i][-3:]), atol=1e-9)


def test_arnoldi_sample_random_complex(ctx):
    # src/samples/sample_arnoldi.cpp: 50x50 complex, m = 40, two leading eigenpairs, AP - PD small
    rng = np.random.default_rng(0)
    n, m = 50, 40
    A = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
    es = pkg.ArnoldiEigenSolver(np.complex128)
    es.setMatrixMultiplication(pkg.DeviceOperator.from_dense(ctx, A))
    es.setMaxIterations(m).setMinIterations(m).setTolerance(1e-14).setMaxEigenvalues(2)
    es.compute()
    ref = rs.ArnoldiEigenSolver("z")
    ref.set_matrix_multiplication(core.Operator.dense(A))
    ref.min_iterations = ref.max_iterations = m
    ref.tolerance, ref.max_eigenvalues = 1e-14, 2
    ref.compute()
    assert _sorted_close(es.eigenvalues(), ref.eigenvalues, 1e-10 * np.abs(ref.eigenvalues).max())
    P, D = es.eigenvectors(), es.eigenvalues()
    res = np.linalg.norm(A @ P - P * D, axis=0)
    assert np.all(res < es.ritzResiduals() + 1e-10)
    # full Krylov dimension on a 4x4 (src/experiments/arnoldi/arnoldi_test.cpp:50-92)
    A4 = A[:4, :4].copy()
    es4 = pkg.ArnoldiEigenSolver(np.complex128)
    es4.setMatrixMultiplication(pkg.DeviceOperator.from_dense(ctx, A4)).setThreshold(1e-14)
    es4.setMaxIterations(es4.unlimited).setMinIterations(es4.unlimited).setMaxEigenvalues(5).setTolerance(1e-10)
    es4.setInitialVector()
    es4.compute()
    P, D = es4.eigenvectors(), es4.eigenvalues()
    assert D.size == 4 and np.abs(A4 @ P - P * D).max() < 1e-12
    assert _sorted_close(D, np.linalg.eigvals(A4), 1e-12)
    assert "INFO      arnoldi steps achieved full of Krylov subspace" in es4.log()


def test_arnoldi_shift_deflation_restart(ctx):
    M = 8
    n = M ** 3
    rp, c, v = syn.convdiff3d_csr(M)
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    x0 = syn.start_vector(n, seed=7)
    exact = syn.convdiff3d_eigenvalues(M, count=3)
    es = pkg.ArnoldiEigenSolver(np.float64)
    es.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(30).setMaxIterations(30).setMaxEigenvalues(2)
    es.setEigenvalueShift(2.0)
    es.compute()
    ref = rs.ArnoldiEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    ref.init, ref.shift = x0, 2.0
    ref.min_iterations = ref.max_iterations = 30
    ref.max_eigenvalues = 2
    ref.compute()
    assert abs(es.eigenvalues()[0] - ref.eigenvalues[0]) < 1e-6 * abs(ref.eigenvalues[0])  # unconverged: sensitive
    assert _sorted_close(es.eigenvalues(), ref.eigenvalues, 1e-3)
    # explicit restart (cfg 3): each cycle restarts from the leading Ritz vector; the leading Ritz value improves
    es2 = pkg.ArnoldiEigenSolver(np.float64)
    es2.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(20).setMaxIterations(20).setMaxEigenvalues(1)
    es2.compute()
    e1 = abs(es2.eigenvalues()[0] - exact[0])
    es2.computeWithRestarts(6)
    e6 = abs(es2.eigenvalues()[0] - exact[0])
    assert e6 < e1 and e6 < 1e-6


# ------------------------------------------------------------------------------------------------
# BASELINE sizes: size-independent properties (the oracle would take minutes here)
# ------------------------------------------------------------------------------------------------
def test_cfg2_full_size_properties(ctx):
    # cfg 2: 2D Laplacian 4096^2 (16.8M rows), Lanczos m=100 with full reorthogonalisation
    N, m = 4096, 100
    n = N * N
    rp, c, v = syn.laplacian2d_csr(N)
    assert rp[-1] == 83869696
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    del rp, c, v
    x0 = syn.start_vector(n, seed=7)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0)
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5).setIndicesForConvergence([0, 1, 2, 3, 4])
    es.compute()
    assert es.iterations() == m and es.nvectors() == m + 1
    ev = es.eigenvalues()
    assert np.all(np.diff(ev) > 0) and ev[0] > 0 and ev[-1] < 8
    X = es.eigenvectors(copy=False)
    rr = es.ritzResiduals()
    for i in range(5):
        xi = np.ascontiguousarray(X[:, i])
        r = np.linalg.norm(op.apply(xi) - ev[i] * xi)
        assert abs(np.linalg.norm(xi) - 1) < 1e-12
        assert abs(r - rr[i]) < 1e-8, (i, r, rr[i])  # ||A x - theta x|| = |beta_m S(m,i)|
    G = X.T @ X
    np.testing.assert_allclose(G, np.eye(5), atol=1e-11)
    # basis orthonormality on a sample of columns (V^T V - I <= 1e-13)
    cols = [0, 1, 37, 64, 99, 100]
    V = np.stack([es.basisVector(k) for k in cols], axis=1)
    np.testing.assert_allclose(V.T @ V, np.eye(len(cols)), atol=1e-13)
    # tridiagonal relation: A u_k = beta_{k-1} u_{k-1} + alpha_k u_k + beta_k u_{k+1}
    a, b = es.alpha(), es.beta()
    u36, u38 = es.basisVector(36), es.basisVector(38)
    resid = op.apply(V[:, 2]) - (b[36] * u36 + a[37] * V[:, 2] + b[37] * u38)
    assert np.linalg.norm(resid) < 1e-12
    # algorithmic byte count of SURVEY.md §8(d): 101 * B_spmv + sum_{c=1..100} (3c+7) n s  (+ first-step vectors)
    want = 100 * op.bytes + (3 * 5050 + 700) * n * 8.0
    assert abs(es.deviceBytes() - want) / want < 0.01


# ------------------------------------------------------------------------------------------------
# matrix-free Heisenberg apply beyond one shared-memory tile, open and periodic chains: 2 passes from L = 14 real / 13
# complex, 3 passes from L = 21 / 20.  With CMPT_B200_HEIS_SIBLINGS=1 a pass also serves the three bonds above its tile
# from the sibling tiles (from 16 / 15 local bits; 2 passes up to 27 bits) — an opt-in variant, same results.
# ------------------------------------------------------------------------------------------------
HEIS_MULTI_PASS = [(14, True, np.float64), (16, False, np.float64), (16, True, np.float64), (17, True, np.float64),
                   (19, False, np.float64), (20, True, np.float64), (21, True, np.float64), (22, False, np.float64),
                   (23, True, np.float64), (13, True, np.complex128), (15, True, np.complex128),
                   (16, False, np.complex128), (17, False, np.complex128), (20, True, np.complex128),
                   (21, True, np.complex128)]


@pytest.mark.parametrize("L,pbc,dtype", HEIS_MULTI_PASS)
def test_heisenberg_multi_pass_apply(ctx, L, pbc, dtype):
    x = syn.start_vector(1 << L, seed=31, dtype=dtype)
    op = pkg.DeviceOperator.heisenberg(ctx, L, 1.0, pbc, dtype=dtype)
    y = op.apply(x)
    yo = core.Operator.heisenberg(L, 1.0, pbc, prefix="z" if dtype == np.complex128 else "d").apply(x)
    np.testing.assert_allclose(y, yo, atol=1e-14)
    op.close()


@pytest.mark.parametrize("L,pbc,dtype", [(16, True, np.float64), (17, True, np.float64), (19, False, np.float64),
                                         (21, True, np.float64), (23, False, np.float64), (23, True, np.float64),
                                         (15, True, np.complex128), (16, False, np.complex128),
                                         (21, True, np.complex128)])
def test_heisenberg_multi_pass_apply_with_sibling_bonds(ctx, monkeypatch, L, pbc, dtype):
    monkeypatch.setenv("CMPT_B200_HEIS_SIBLINGS", "1")  # read when the operator plans its passes (first apply)
    x = syn.start_vector(1 << L, seed=32, dtype=dtype)
    op = pkg.DeviceOperator.heisenberg(ctx, L, 1.0, pbc, dtype=dtype)
    y = op.apply(x)
    yo = core.Operator.heisenberg(L, 1.0, pbc, prefix="z" if dtype == np.complex128 else "d").apply(x)
    np.testing.assert_allclose(y, yo, atol=1e-14)
    op.close()


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_heisenberg_sibling_passes_lanczos_matches_oracle(ctx, monkeypatch, dtype):
    # the u column, the 1/beta scaling and the alpha dot ride in the contiguous pass next to the sibling bonds: whole solver steps
    monkeypatch.setenv("CMPT_B200_HEIS_SIBLINGS", "1")
    L, m = 17, 12
    pre = "z" if dtype == np.complex128 else "d"
    x0 = syn.start_vector(1 << L, seed=5, dtype=dtype)
    es = pkg.LanczosEigenSolver(dtype)
    es.setMatrixMultiplication(pkg.DeviceOperator.heisenberg(ctx, L, dtype=dtype)).setInitialVector(x0)
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(3).setIndicesForConvergence([0, 1, 2])
    es.compute()
    ref = _oracle_lanczos(core.Operator.heisenberg(L, 1.0, True, prefix=pre), x0, m, 3, indices_for_convergence=[0, 1, 2],
                          prefix=pre)
    _compare_lanczos(es, ref)


# ------------------------------------------------------------------------------------------------
# row-partitioned matrix-free Heisenberg operator (cfg 5) with virtual ranks on one GPU
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L,nranks,pbc,dtype", [(10, 2, True, np.float64), (10, 2, False, np.float64),
                                                (11, 4, True, np.float64), (12, 8, True, np.float64),
                                                (12, 8, False, np.complex128), (9, 4, True, np.complex128),
                                                (13, 16, True, np.float64), (17, 4, True, np.float64),
                                                (23, 2, True, np.float64), (16, 4, False, np.complex128)])
def test_heisenberg_partitioned_virtual_ranks(ctx, L, nranks, pbc, dtype):
    import ctypes as C

    n = 1 << L
    x = syn.start_vector(n, seed=21, dtype=dtype)
    y = np.empty_like(x)
    capi.check(capi.lib().cmb_debug_heisenberg_virtual(ctx.h, capi.dtype_code(dtype), L, 1.0, int(pbc), nranks,
                                                       capi.ptr(x), capi.ptr(y)))
    yo = core.Operator.heisenberg(L, 1.0, pbc, prefix="z" if dtype == np.complex128 else "d").apply(x)
    np.testing.assert_allclose(y, yo, atol=1e-14)


# ------------------------------------------------------------------------------------------------
# LanczosExponentialSolver (lanczos.hpp:1002-1164)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,x", [(np.float64, -0.7), (np.float64, 0.3), (np.complex128, -0.4j), (np.complex128, -0.2 + 0.5j)])
def test_exponential_solver_matches_oracle_and_expm(ctx, dtype, x):
    import scipy.linalg as sla

    N = 9
    n = N * N
    rp, c, v = syn.laplacian2d_csr(N)
    if dtype == np.complex128:
        rp, c, v = syn.hermitian_chain_csr(n)
    A = _csr_dense(rp, c, v, n)
    x0 = 0.37 * syn.start_vector(n, seed=7, dtype=dtype)  # not normalised: the expansion uses the raw initial vector
    es = pkg.LanczosEigenSolver(dtype)
    es.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v)).setInitialVector(x0)
    es.setMinIterations(30).setMaxIterations(30)
    out = es.expSolveWithLanczos(x)
    ref = rs.LanczosEigenSolver("z" if dtype == np.complex128 else "d")
    ref.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    ref.init = x0
    ref.min_iterations = ref.max_iterations = 30
    want = rs.exp_solve_with_lanczos(x if dtype == np.complex128 else float(np.real(x)), ref)
    np.testing.assert_allclose(out, want, atol=1e-11)
    # truncated expansion (maxEigenvalues) follows the reference too
    es.setMaxEigenvalues(4)
    ref.max_eigenvalues = 4
    np.testing.assert_allclose(es.expSolveWithLanczos(x), rs.exp_solve_with_lanczos(
        x if dtype == np.complex128 else float(np.real(x)), ref), atol=1e-11)
    # no eigenvectors -> nothing is summed (lanczos.hpp:1037-1039)
    es.setComputeEigenvectorsOn(False)
    assert np.all(es.expSolveWithLanczos(x) == 0)
    # full Krylov space: the expansion is exp(xA) v exactly
    es2 = pkg.LanczosEigenSolver(dtype)
    es2.setMatrixMultiplication(pkg.DeviceOperator.from_csr(ctx, rp, c, v)).setInitialVector(x0)
    es2.setMinIterations(n).setMaxIterations(n).setThreshold(1e-13)
    full = es2.expSolveWithLanczos(x)
    np.testing.assert_allclose(full, sla.expm(x * A) @ x0, atol=1e-9)
    # Taylor variants through the device operator's host interface
    radius = np.abs(np.linalg.eigvalsh(A)).max()
    t1 = es2.expSolveWithTaylor(x, radius, x0, auto_division=True)
    np.testing.assert_allclose(t1, sla.expm(x * A) @ x0, atol=1e-10)
    t2 = es2.expSolveWithTaylor(x * 0.1, radius, x0, auto_division=False)
    np.testing.assert_allclose(t2, sla.expm(0.1 * x * A) @ x0, atol=1e-10)


def test_invalid_csr_is_rejected_on_the_device(ctx):
    """Bad column indices / a decreasing rowptr are caught during the SELL build (no out-of-bounds operator is left behind)."""
    rp, c, v = syn.laplacian2d_csr(10)
    bad = c.copy()
    bad[17] = 100  # n = 100: one past the end
    with pytest.raises(capi.CmbError) as e:
        pkg.DeviceOperator.from_csr(ctx, rp, bad, v)
    assert e.value.code == -1 and "column index" in str(e.value)
    bad[17] = -3
    with pytest.raises(capi.CmbError):
        pkg.DeviceOperator.from_csr(ctx, rp, bad, v)
    rp2 = rp.copy()
    rp2[5], rp2[6] = rp2[6], rp2[5] - 1  # decreasing
    with pytest.raises(capi.CmbError) as e:
        pkg.DeviceOperator.from_csr(ctx, rp2, c, v)
    assert "rowptr" in str(e.value)
    with pytest.raises(ValueError):
        pkg.DeviceOperator.from_csr(ctx, rp, c[:-3], v)
    # the context is still healthy
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    x = syn.start_vector(100, seed=1)
    np.testing.assert_allclose(op.apply(x), core.Operator.csr(rp, c, v).apply(x), atol=1e-14)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_device_resident_operator_algebra(ctx, dtype):
    """cmb_op_linear_create / cmb_op_product_create: sums, scalar multiples and products of operators in HBM applied on the
    device (the reference composes them on the host, vector_map.hpp:38-266), alone and under the Lanczos driver."""
    rng = np.random.default_rng(2)
    N = 18
    n = N * N
    rp, c, v = syn.laplacian2d_csr(N)
    A = _csr_dense(rp, c, v, n).astype(dtype)
    Bd = rng.normal(size=(n, n)).astype(dtype)
    if dtype == np.complex128:
        Bd = Bd + 1j * rng.normal(size=(n, n))
    Bd = 0.05 * (Bd + Bd.conj().T)  # Hermitian
    opA = pkg.DeviceOperator.from_csr(ctx, rp, c, v.astype(dtype))
    opB = pkg.DeviceOperator.from_dense(ctx, Bd)
    ca, cb = (2.0, -0.5) if dtype == np.float64 else (2.0 + 0.0j, -0.5 + 0.0j)
    lin = pkg.DeviceOperator.linear(ctx, [opA, opB], [ca, cb])
    prod = pkg.DeviceOperator.product(ctx, opA, opB)
    x = syn.start_vector(n, seed=4, dtype=dtype)
    np.testing.assert_allclose(lin.apply(x), ca * (A @ x) + cb * (Bd @ x), atol=1e-12)
    np.testing.assert_allclose(prod.apply(x), A @ (Bd @ x), atol=1e-12)
    if dtype == np.complex128:
        lz = pkg.DeviceOperator.linear(ctx, [opA, opB], [1.0 + 0.5j, 0.25 - 1.0j])
        np.testing.assert_allclose(lz.apply(x), (1.0 + 0.5j) * (A @ x) + (0.25 - 1.0j) * (Bd @ x), atol=1e-12)
        lz.close()
    # under the solver: the linear combination is Hermitian; compare with the assembled dense operator and the oracle
    H = ca * A + cb * Bd
    m = 40
    es = pkg.LanczosEigenSolver(dtype)
    es.setMatrixMultiplication(lin).setInitialVector(x).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(3).setEigenvalueShift(0.3)
    es.compute()
    ref = _oracle_lanczos(core.Operator.dense(H), x, m, 3, prefix="z" if dtype == np.complex128 else "d", shift=0.3)
    _compare_lanczos(es, ref)
    es.close()
    for o in (lin, prod, opA, opB):
        o.close()
