"""Thick-restart (Krylov-Schur) Arnoldi — include/cmpt/eigen_ex/arnoldi_restart.hpp through the solver binding.

The reference has no restarted solver; BASELINE cfg 3 asks for "Arnoldi (restarted, m=50)".
Checks: wanted eigenvalues against closed forms / dense LAPACK where the eigenproblem is well conditioned, Ritz residuals
||A x - theta x|| recomputed on the host, bounded basis, complex-conjugate pairs of a real operator kept together,
complex Scalar, and the same run row-partitioned over two virtual ranks.

Conditioning note: the convection-diffusion operator of cfg 3 (gamma = 0.3, 0.2, 0.1) is similar to a symmetric matrix
through a diagonal scaling of condition ((1+g)/(1-g))^(M/2) per dimension — 1e34 at M = 256 — so its exact eigenvalues
cannot be reproduced to 1e-10 in floating point at that size by ANY method (a residual of 1e-12 still leaves the Ritz
value anywhere in a large pseudo-spectrum).  The closed form is therefore checked at sizes / gammas where that
condition number is moderate, and by the residual everywhere.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pkg.Context(0)
    yield c
    c.close()


def _scipy(rp, c, v, n):
    return sp.csr_matrix((v, c, rp), shape=(n, n))


def _solve(ctx, rp, c, v, dtype=np.float64, wanted=5, basis=30, tol=1e-11, which=None, x0=None, restarts=500):
    n = rp.size - 1
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v.astype(dtype))
    es = pkg.ThickRestartArnoldi(dtype)
    es.setMatrixMultiplication(op).setInitialVector(syn.start_vector(n, seed=7, dtype=dtype) if x0 is None else x0)
    es.setWanted(wanted).setMaxBasis(basis).setTolerance(tol).setMaxRestarts(restarts)
    if which is not None:
        es.setWhich(which)
    es.compute()
    return es, op


def test_convection_diffusion_matches_closed_form(ctx):
    # mild convection: the similarity to a symmetric matrix has condition ~ 7, the closed form is attainable
    M, gamma = 24, (0.03, 0.02, 0.01)
    rp, c, v = syn.convdiff3d_csr(M, gamma=gamma)
    es, op = _solve(ctx, rp, c, v, basis=30)
    ev = es.eigenvalues()
    exact = syn.convdiff3d_eigenvalues(M, gamma=gamma, count=5)
    assert es.converged() == 5 and "converged" in es.log()[-1]
    assert es.restarts() >= 1 and es.nvectors() <= 30
    assert np.abs(ev.imag).max() < 1e-9
    np.testing.assert_allclose(np.sort(ev.real)[::-1], exact, rtol=1e-10)
    A = _scipy(rp, c, v, M ** 3)
    X = es.eigenvectors()
    res = np.linalg.norm(A @ X - X * ev, axis=0)
    np.testing.assert_allclose(np.linalg.norm(X, axis=0), 1.0, atol=1e-12)
    assert np.all(res < 1e-9) and np.all(np.abs(res - es.residuals()) < 1e-9)
    es.close()
    op.close()


def test_cfg3_operator_small_grid_against_dense_eig(ctx):
    # the cfg 3 operator itself (gamma = 0.3, 0.2, 0.1) at a size LAPACK can do and the conditioning allows
    M = 10
    rp, c, v = syn.convdiff3d_csr(M)
    es, op = _solve(ctx, rp, c, v, basis=40, which=pkg.ThickRestartArnoldi.LARGEST_REAL)
    A = _scipy(rp, c, v, M ** 3)
    w = np.linalg.eigvals(A.toarray())
    want = np.sort(w.real)[::-1][:5]
    np.testing.assert_allclose(np.sort(es.eigenvalues().real)[::-1], want, rtol=1e-8)
    np.testing.assert_allclose(want, syn.convdiff3d_eigenvalues(M, count=5), rtol=1e-8)
    X = es.eigenvectors()
    assert np.all(np.linalg.norm(A @ X - X * es.eigenvalues(), axis=0) < 1e-9)
    es.close()
    op.close()


def test_real_operator_with_complex_pairs_and_complex_scalar(ctx):
    rng = np.random.default_rng(5)
    n = 600
    # block-diagonal rotations with decreasing moduli plus a small non-normal perturbation: the dominant eigenvalues
    # come in complex-conjugate pairs
    mods = np.concatenate([[9.0, 8.0, 7.2], rng.uniform(0.5, 5.0, n // 2 - 3)])
    rows, cols, vals = [], [], []
    for b, r in enumerate(mods):
        th = 0.4 + 0.05 * b
        i = 2 * b
        for (di, dj, val) in ((0, 0, r * np.cos(th)), (0, 1, -r * np.sin(th)), (1, 0, r * np.sin(th)), (1, 1, r * np.cos(th))):
            rows.append(i + di)
            cols.append(i + dj)
            vals.append(val)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n)) + 0.02 * sp.random(n, n, density=0.01, random_state=3, format="csr")
    A = A.tocsr()
    A.sort_indices()
    rp, c, v = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data
    es, op = _solve(ctx, rp, c, v, wanted=4, basis=24, tol=1e-10)
    ev = es.eigenvalues()
    w = np.linalg.eigvals(A.toarray())
    want = w[np.argsort(-np.abs(w))][:4]
    assert es.converged() == 4
    for z in want:
        assert np.abs(ev - z).min() < 1e-8 * abs(z)
    assert abs(ev[0] - np.conj(ev[1])) < 1e-8 * abs(ev[0])  # the pair stays together
    X = es.eigenvectors()
    assert np.all(np.linalg.norm(A @ X - X * ev, axis=0) < 1e-8)
    es.close()
    op.close()
    # complex Scalar: a complex non-Hermitian operator
    d = np.concatenate([[12 + 3j, -11 + 1j, 10 - 2j, 9j + 4], rng.uniform(-3, 3, n - 4) + 1j * rng.uniform(-3, 3, n - 4)])
    B = (sp.diags(d) + 0.05 * (sp.random(n, n, density=0.01, random_state=8) + 1j * sp.random(n, n, density=0.01, random_state=9))).tocsr()
    B.sort_indices()
    rp, c, v = B.indptr.astype(np.int64), B.indices.astype(np.int32), B.data
    es, op = _solve(ctx, rp, c, v, dtype=np.complex128, wanted=4, basis=20, tol=1e-10)
    w = np.linalg.eigvals(B.toarray())
    want = w[np.argsort(-np.abs(w))][:4]
    np.testing.assert_allclose(es.eigenvalues(), want, rtol=1e-8)
    X = es.eigenvectors()
    assert np.all(np.linalg.norm(B @ X - X * es.eigenvalues(), axis=0) < 1e-8)
    es.close()
    op.close()


def test_thick_restart_arnoldi_on_two_virtual_ranks(ctx):
    import multirank_checks as mc

    M = 12
    n = M ** 3
    gamma = (0.03, 0.02, 0.01)
    full = syn.convdiff3d_csr(M, gamma=gamma)
    x0 = syn.start_vector(n, seed=7)
    single, op = _solve(ctx, *full, basis=30)
    want = single.eigenvalues().copy()
    single.close()
    op.close()

    def work(vctx, comm):
        r0, r1 = comm.row_range(n)
        opv = pkg.DeviceOperator.from_csr(vctx, *mc.shard_of(full, r0, r1), n_global=n, row_begin=r0)
        es = pkg.ThickRestartArnoldi(np.float64)
        es.setMatrixMultiplication(opv).setInitialVector(x0[r0:r1]).setWanted(5).setMaxBasis(30).setTolerance(1e-11).setMaxRestarts(500)
        es.compute()
        out = (es.eigenvalues(), es.restarts(), es.converged(), np.concatenate(comm.gather(es.eigenvectors())))
        es.close()
        opv.close()
        return out

    results, _ = pkg.run_virtual_ranks(2, work)
    A = _scipy(*full, n)
    for ev, restarts, conv, X in results:
        assert conv == 5 and restarts >= 1
        np.testing.assert_allclose(np.sort(ev.real)[::-1], np.sort(want.real)[::-1], rtol=1e-10)
        np.testing.assert_allclose(np.sort(ev.real)[::-1], syn.convdiff3d_eigenvalues(M, gamma=gamma, count=5), rtol=1e-10)
        assert np.all(np.linalg.norm(A @ X - X * ev, axis=0) < 1e-9)
