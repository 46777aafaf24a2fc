"""Pins the CPU oracle (oracle/) on the analytic known answers of SURVEY.md §8(c).

The reference ships no golden vectors or asserting tests ("parity unpinned"), so these
KATs are what anchors the restatement: the two Lanczos samples' matrices, the Laplacian /
convection-diffusion closed-form spectra, and Heisenberg-ring ground-state energies.
"""
import numpy as np
import pytest

from cmpt_eigenex_b200 import synthetic as syn
from oracle import core
from oracle import reference_solvers as rs

core.set_num_threads(1)  # tiny vectors: OpenMP fork/join would dominate


def test_sample_lanczos1_3x3():
    # src/samples/sample_lanczos1.cpp:13-32
    H = np.array([[1.0, 0.5, 0.0], [0.5, 2.0, 0.5], [0.0, 0.5, 3.0]])
    es = rs.LanczosEigenSolver("d")
    es.set_matrix_multiplication(core.Operator.dense(H))
    es.tolerance = 1.0e-5
    es.max_iterations = 100
    assert es.compute() == 0
    want = np.array([2 - np.sqrt(1.5), 2.0, 2 + np.sqrt(1.5)])
    np.testing.assert_allclose(es.eigenvalues, want, rtol=0, atol=1e-14)
    vecs = np.array([[0.90824829, -0.40824829, 0.09175171],
                     [0.40824829, 0.81649658, -0.40824829],
                     [0.09175171, 0.40824829, 0.90824829]]).T
    np.testing.assert_allclose(es.eigenvectors, vecs, atol=1e-8)
    assert (es.eigenvectors[0] > 0).all()  # phase fix: first component real positive
    assert es.base.nvectors == 3 and es.iterations == 2
    assert rs.HEAD_INFO + "lanczos steps achieved full of Krylov subspace" in es.log
    # compute() pushes "was called" and then clears the log (lanczos.hpp:719-721)
    assert not any("was called" in s for s in es.log)
    assert es.log[-1] == rs.HEAD_INFO + "EigenSolver<ScalarType>::compute(...) finish computing"
    # H x = theta x
    np.testing.assert_allclose(H @ es.eigenvectors, es.eigenvectors * es.eigenvalues, atol=1e-13)


def _sample2_solver():
    rp, c, v = syn.hermitian_chain_csr(200)
    es = rs.LanczosEigenSolver("z")
    es.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    es.threshold = 1.0e-14
    es.init = core.seeded_vector(1, 200, "z")
    return es


def test_sample_lanczos2_settings():
    # src/samples/sample_lanczos2.cpp:44-59 with its exact settings
    es = _sample2_solver()
    es.tolerance = 1.0e-7
    es.min_iterations = rs.UNLIMITED
    es.max_iterations = 1000
    es.max_eigenvalues = 10
    es.compute()
    assert es.eigenvalues.shape == (10,) and es.eigenvectors.shape == (200, 10)
    assert rs.HEAD_INFO + "lanczos steps converged with tolerance" in es.log
    # the stop rule only watches index 0: it is converged to ~tolerance*scale*few
    assert abs(es.eigenvalues[0] - 2 * np.cos(200 * np.pi / 201)) < 1e-4
    assert np.all(np.diff(es.eigenvalues) > 0)
    # first non-zero element of every eigenvector is real positive
    x0 = es.eigenvectors[0]
    assert np.all(np.abs(x0.imag) < 1e-14) and np.all(x0.real > 0)
    np.testing.assert_allclose(np.linalg.norm(es.eigenvectors, axis=0), 1.0, atol=1e-13)


def test_sample_lanczos2_full_krylov_spectrum():
    es = _sample2_solver()
    es.tolerance = 0.0
    es.compute()
    want = np.sort(2 * np.cos(np.arange(1, 201) * np.pi / 201))
    assert es.base.nvectors == 200
    np.testing.assert_allclose(es.eigenvalues, want, atol=2e-13)


def test_default_start_vector_stream():
    # std::mt19937 default seed (5489) + std::normal_distribution<double>, libstdc++ stream
    v = core.default_vector(5, "d")
    assert abs(np.linalg.norm(v) - 1) < 1e-15
    z = core.default_vector(3, "z")
    assert abs(np.linalg.norm(z) - 1) < 1e-15
    # same stream: complex draws (re, im) consume the values the real vector gets in order
    w = core.default_vector(6, "d")
    zz = w[0::2] + 1j * w[1::2]
    np.testing.assert_allclose(z, zz / np.linalg.norm(zz), atol=1e-15)


@pytest.mark.parametrize("N", [8, 20])
def test_laplacian2d_spectrum(N):
    rp, c, v = syn.laplacian2d_csr(N)
    n = N * N
    assert rp[-1] == 5 * n - 4 * N
    es = rs.LanczosEigenSolver("d")
    es.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    es.init = syn.start_vector(n, seed=7)
    es.tolerance = 0.0
    es.compute_eigenvectors_on = False
    es.compute()  # runs until breakdown: one Ritz value per distinct eigenvalue
    k = np.arange(1, N + 1)
    cc = 2 - 2 * np.cos(k * np.pi / (N + 1))
    exact = np.unique(np.round((cc[:, None] + cc[None, :]).ravel(), 12))
    assert abs(es.eigenvalues[0] - syn.laplacian2d_eigenvalues(N, 1)[0]) < 1e-12
    # extremal Ritz values at breakdown are the distinct eigenvalues (interior ones may still be
    # unconverged copies because the run only ends when beta <= threshold)
    np.testing.assert_allclose(es.eigenvalues[:3], exact[:3], atol=1e-10)
    np.testing.assert_allclose(es.eigenvalues[-3:], exact[-3:], atol=1e-10)
    # ends by breakdown, or earlier if the tracked Ritz value repeats bit-for-bit (tolerance 0)
    assert (rs.HEAD_INFO + "lanczos steps finished with threshold" in es.log
            or rs.HEAD_INFO + "lanczos steps converged with tolerance" in es.log)


def test_convdiff3d_spectrum_arnoldi():
    M = 5
    rp, c, v = syn.convdiff3d_csr(M)
    n = M ** 3
    A = np.zeros((n, n))
    for r in range(n):
        A[r, c[rp[r]:rp[r + 1]]] = v[rp[r]:rp[r + 1]]
    exact = syn.convdiff3d_eigenvalues(M, count=3)
    np.testing.assert_allclose(np.sort(np.linalg.eigvals(A).real)[::-1][:3], exact, atol=1e-10)
    for prefix in ("d", "z"):
        es = rs.ArnoldiEigenSolver(prefix)
        es.set_matrix_multiplication(core.Operator.csr(rp, c, v.astype(complex) if prefix == "z" else v))
        es.init = syn.start_vector(n, seed=7, dtype=complex if prefix == "z" else float)
        es.min_iterations = es.max_iterations = 60
        es.max_eigenvalues = 3
        es.compute()
        assert es.iterations == 60 and es.has_warn() == 1
        np.testing.assert_allclose(es.eigenvalues.real, exact, atol=1e-8)
        assert np.abs(es.eigenvalues.imag).max() < 1e-8
        # A P = P D within the Ritz residual
        P, D = es.eigenvectors, es.eigenvalues
        res = np.linalg.norm(A @ P - P * D, axis=0)
        assert np.all(res < es.ritz_residuals() + 1e-10)


def test_arnoldi_full_krylov_random_complex():
    # src/experiments/arnoldi/arnoldi_test.cpp:50-92 shape: run to full Krylov dimension, AP - PD ~ 0
    rng = np.random.default_rng(3)
    n = 6
    A = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
    es = rs.ArnoldiEigenSolver("z")
    es.set_matrix_multiplication(core.Operator.dense(A))
    es.threshold = 1e-14
    es.min_iterations = es.max_iterations = rs.UNLIMITED
    es.tolerance = 1e-10
    es.max_eigenvalues = 7
    es.compute()
    assert es.base.nvectors == n and len(es.eigenvalues) == n
    P, D = es.eigenvectors, es.eigenvalues
    assert np.abs(A @ P - P * D).max() < 1e-11
    assert np.all(np.diff(np.abs(D)) <= 1e-12)  # descending |lambda|
    want = np.linalg.eigvals(A)
    assert np.abs(np.sort_complex(D) - np.sort_complex(want)).max() < 1e-11


@pytest.mark.parametrize("L", [8, 12])
def test_heisenberg_ring_ground_state(L):
    n = 1 << L
    rp, c, v = syn.heisenberg_csr(L)
    es = rs.LanczosEigenSolver("d")
    es.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    es.init = syn.start_vector(n, seed=7)
    es.max_iterations = 200
    es.compute_eigenvectors_on = False
    es.compute()
    assert abs(es.eigenvalues[0] - syn.HEISENBERG_RING_E0[L]) < 2e-10
    if L == 12:
        assert abs(es.iterations - 30) <= 1  # stop rule fires at ~30 for L=12
    # matrix-free operator agrees with the explicit CSR
    x = syn.start_vector(n, seed=11)
    y1 = core.Operator.csr(rp, c, v).apply(x)
    y2 = core.Operator.heisenberg(L).apply(x)
    np.testing.assert_allclose(y1, y2, atol=1e-14)


def test_deflation_and_shift_and_interval():
    # orthogonalizingVectors_ (lanczos.hpp:312-314,421-425): deflating the ground state makes
    # the first excited state the lowest Ritz value; shift is added then removed (:390,:794).
    N = 10
    n = N * N
    rp, c, v = syn.laplacian2d_csr(N)
    op = core.Operator.csr(rp, c, v)
    es = rs.LanczosEigenSolver("d")
    es.set_matrix_multiplication(op)
    es.init = syn.start_vector(n, seed=7)
    es.min_iterations = es.max_iterations = 60
    es.max_eigenvalues = 1
    es.compute()
    g = es.eigenvectors[:, 0].copy()
    es2 = rs.LanczosEigenSolver("d")
    es2.set_matrix_multiplication(op)
    es2.init = syn.start_vector(n, seed=7)
    es2.ortho = [g]
    es2.shift = 1.5
    es2.min_iterations = es2.max_iterations = 60
    es2.max_eigenvalues = 1
    es2.compute()
    lam = syn.laplacian2d_eigenvalues(N, 3)
    assert abs(es.eigenvalues[0] - lam[0]) < 1e-12
    assert abs(es2.eigenvalues[0] - lam[1]) < 1e-10
    assert abs(g @ es2.eigenvectors[:, 0]) < 1e-12
    # interval 0 -> plain three-term recurrence: orthogonality degrades but alpha/beta start equal
    es3 = rs.LanczosEigenSolver("d")
    es3.set_matrix_multiplication(op)
    es3.init = syn.start_vector(n, seed=7)
    es3.interval = 0
    es3.min_iterations = es3.max_iterations = 10
    es3.compute()
    a3, b3 = es3.alpha_beta()
    es4 = rs.LanczosEigenSolver("d")
    es4.set_matrix_multiplication(op)
    es4.init = syn.start_vector(n, seed=7)
    es4.min_iterations = es4.max_iterations = 10
    es4.compute()
    a4, b4 = es4.alpha_beta()
    np.testing.assert_allclose(a3, a4, atol=1e-10)
    np.testing.assert_allclose(b3, b4, atol=1e-10)


def test_zero_start_vector_fails_quietly():
    # lanczos.hpp:316-318,748-752: norm below threshold -> no vectors, INFO log line, empty results
    H = np.eye(4)
    es = rs.LanczosEigenSolver("d")
    es.set_matrix_multiplication(core.Operator.dense(H))
    es.init = np.zeros(4)
    es.compute()
    assert es.base.nvectors == 0 and len(es.eigenvalues) == 0
    assert rs.HEAD_INFO + "initial lanczosvector generation fail" in es.log


def test_formal_index():
    assert [rs.formal_index(i, 4) for i in (-5, -4, -1, 0, 3, 4)] == [-1, 0, 3, 0, 3, -1]
