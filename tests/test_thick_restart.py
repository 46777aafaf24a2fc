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
