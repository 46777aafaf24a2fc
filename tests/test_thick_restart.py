"""cmb_lanczos_thick_restart through the C-ABI: after compressing the basis to [Ritz vectors, u_m] the basis stays
orthonormal, V^H A V has the arrowhead + tridiagonal structure the host assumes, and continuing the iteration
converges to the eigenvalues of the operator (numpy does the host-side algebra here)."""
import ctypes as C

import numpy as np
import pytest

import cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import capi, synthetic as syn

pytestmark = pytest.mark.gpu


def _run(lib, kh, oph, nsteps):
    a = np.zeros(nsteps + 1)
    b = np.zeros(nsteps + 1)
    done, status = C.c_int64(), C.c_int32()
    capi.check(lib.cmb_lanczos_run(kh, oph, 0.0, 1, 1e-12, nsteps, a.ctypes.data, b.ctypes.data, C.byref(done),
                                   C.byref(status)))
    return a, b, done.value


def _basis(lib, kh, n):
    nk = lib.cmb_krylov_ncols(kh)
    V = np.zeros((n, nk))
    col = np.zeros(n)
    for j in range(nk):
        capi.check(lib.cmb_krylov_get_col(kh, j, col.ctypes.data))
        V[:, j] = col
    return V


def test_thick_restart_structure_and_convergence():
    lib = capi.lib()
    n, m, keep = 1500, 24, 6
    A = syn.dense_symmetric(n, seed=3)
    ctx = pkg.Context(0)
    op = pkg.DeviceOperator.from_dense(ctx, A)
    kh = C.c_void_p()
    capi.check(lib.cmb_krylov_create(ctx.h, capi.CMB_F64, n, 0, n, 64, C.byref(kh)))
    x0 = syn.start_vector(n, seed=5)
    st = C.c_int32()
    capi.check(lib.cmb_krylov_start(kh, x0.ctypes.data, 1e-12, C.byref(st)))
    a, b, done = _run(lib, kh, op.h, m + 1)  # u_0..u_m
    assert done == m + 1
    alpha, beta = list(a[:m + 1]), list(b[:m])
    exact = np.linalg.eigvalsh(A)
    theta_hist = []
    for cycle in range(12):
        mm = len(alpha) - 1
        T = np.diag(alpha[:mm])
        k = keep if cycle > 0 else 0
        for i in range(mm - 1):
            j = k if i < k else i + 1
            T[i, j] = T[j, i] = beta[i]
        V = _basis(lib, kh, n)
        if cycle in (0, 1, 5):
            # the structure the host assumes is what the device holds
            assert np.abs(V.T @ V - np.eye(V.shape[1])).max() < 1e-12
            assert np.abs(V[:, :mm].T @ A @ V[:, :mm] - T).max() < 1e-10
        w, S = np.linalg.eigh(T)
        theta_hist.append(w[0])
        coef = np.asfortranarray(S[:, :keep])
        coupling = beta[mm - 1] * S[mm - 1, :keep]
        capi.check(lib.cmb_lanczos_thick_restart(kh, coef.ctypes.data_as(C.POINTER(C.c_double)), mm, mm, keep))
        assert lib.cmb_krylov_ncols(kh) == keep + 1
        alpha = list(w[:keep]) + [alpha[mm]]
        beta = list(coupling)
        a, b, done = _run(lib, kh, op.h, m - keep)
        assert done == m - keep
        alpha += list(a[:done])
        beta += list(b[:done])
    assert theta_hist[-1] <= theta_hist[0] + 1e-12  # Ritz values decrease monotonically towards the eigenvalue
    assert abs(theta_hist[-1] - exact[0]) < 1e-8 * max(1.0, abs(exact[0]))
    capi.check(lib.cmb_krylov_destroy(kh))
    op.close()
    ctx.close()


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_thick_restart_solver_binding(dtype):
    """ThickRestartLanczos through the solver binding: lowest pairs of a rectangular 2D Laplacian (real) and of the
    Hermitian chain (complex) with a basis that could never hold an unrestarted run."""
    ctx = pkg.Context(0)
    if dtype == np.float64:
        nx, ny = 60, 47
        n = nx * ny
        ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        rows, cols, vals = [], [], []
        for di, dj, v in ((0, 0, 4.0), (1, 0, -1.0), (-1, 0, -1.0), (0, 1, -1.0), (0, -1, -1.0)):
            ok = (ii + di >= 0) & (ii + di < nx) & (jj + dj >= 0) & (jj + dj < ny)
            rows.append((ii * ny + jj)[ok])
            cols.append(((ii + di) * ny + jj + dj)[ok])
            vals.append(np.full(ok.sum(), v))
        import scipy.sparse as sp

        A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
        A.sort_indices()
        rp, c, v = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data
        ex = np.sort((4 - 2 * np.cos(np.arange(1, nx + 1)[:, None] * np.pi / (nx + 1))
                      - 2 * np.cos(np.arange(1, ny + 1)[None, :] * np.pi / (ny + 1))).reshape(-1))
    else:
        n = 200
        rp, c, v = syn.hermitian_chain_csr(n)
        ex = np.sort(2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1)))
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    tr = pkg.ThickRestartLanczos(dtype)
    tr.setMatrixMultiplication(op).setInitialVector(syn.start_vector(n, seed=3, dtype=dtype))
    tr.setWanted(4).setMaxBasis(24).setTolerance(1e-11).setMaxRestarts(500)
    tr.compute()
    assert tr.converged() == 4 and tr.restarts() > 0
    assert tr.nvectors() <= 24
    np.testing.assert_allclose(tr.eigenvalues(), ex[:4], atol=1e-9)
    X = tr.eigenvectors()
    assert X.shape == (n, 4)
    import scipy.sparse as sp

    A = sp.csr_matrix((v, c, rp), shape=(n, n))
    res = np.linalg.norm(A @ X - X * tr.eigenvalues(), axis=0)
    assert res.max() < 1e-8 and np.all(np.abs(res - tr.residuals()) < 1e-8)
    assert any("thick-restart lanczos converged" in line for line in tr.log())
    with pytest.raises(capi.CmbError):
        tr.continueToCompute()
    tr.close()
    op.close()
    ctx.close()


def test_thick_restart_two_lowest_states_of_the_heisenberg_ring():
    """SURVEY.md Appendix E known answers at scale: ground state and first excited level of the L = 20 ring
    (matrix-free operator, 2^20 states) with at most 20 Lanczos vectors on the device."""
    L = 20
    ctx = pkg.Context(0)
    op = pkg.DeviceOperator.heisenberg(ctx, L)
    tr = pkg.ThickRestartLanczos(np.float64)
    tr.setMatrixMultiplication(op).setInitialVector(syn.start_vector(1 << L, seed=7))
    tr.setWanted(2).setMaxBasis(20).setTolerance(1e-9).setMaxRestarts(200).setComputeEigenvectorsOn(False)
    tr.compute()
    assert tr.converged() == 2 and tr.nvectors() <= 20
    e = tr.eigenvalues()
    assert abs(e[0] - (-8.904386529876)) < 1e-9 and abs(e[1] - (-8.686440986187)) < 1e-8
    tr.close()
    op.close()
    ctx.close()


@pytest.mark.parametrize("nranks", [2, 4])
def test_thick_restart_lanczos_row_partitioned_on_virtual_ranks(nranks):
    """The compress step (cmb_lanczos_thick_restart) and the restarted chain on a row-partitioned basis: every rank holds
    a slab of the kept Ritz vectors; the projected arrowhead matrix and the stop decision are identical on all ranks."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import multirank_checks as mc

    N = 40
    n = N * N
    full = syn.laplacian2d_csr(N)
    x0 = syn.start_vector(n, seed=3)
    ex = np.sort((4 - 2 * np.cos(np.arange(1, N + 1)[:, None] * np.pi / (N + 1))
                  - 2 * np.cos(np.arange(1, N + 1)[None, :] * np.pi / (N + 1))).reshape(-1))

    def work(ctx, comm):
        r0, r1 = comm.row_range(n)
        op = pkg.DeviceOperator.from_csr(ctx, *mc.shard_of(full, r0, r1), n_global=n, row_begin=r0)
        tr = pkg.ThickRestartLanczos(np.float64)
        tr.setMatrixMultiplication(op).setInitialVector(x0[r0:r1])
        tr.setWanted(3).setMaxBasis(20).setTolerance(1e-11).setMaxRestarts(500)
        tr.compute()
        out = (tr.eigenvalues(), tr.converged(), tr.restarts(), tr.nvectors(), np.concatenate(comm.gather(tr.eigenvectors())))
        tr.close()
        op.close()
        return out

    results, _ = pkg.run_virtual_ranks(nranks, work)
    import scipy.sparse as sp

    A = sp.csr_matrix((full[2], full[1], full[0]), shape=(n, n))
    for ev, conv, restarts, nvec, X in results:
        assert conv == 3 and restarts > 0 and nvec <= 20
        # the lowest level is simple, the next one is a degenerate pair of which a single Krylov sequence finds one member
        np.testing.assert_allclose(ev[:2], ex[:2], atol=1e-9)
        assert np.abs(ex - ev[2]).min() < 1e-9
        assert np.all(np.linalg.norm(A @ X - X * ev, axis=0) < 1e-8)
    assert all(np.array_equal(results[0][0], r[0]) for r in results[1:])
