"""CPU tests of the product's host-side pieces: the m x m Ritz solvers (detail/tridiag_eigen.hpp,
detail/hessenberg_eigen.hpp) against LAPACK, and the C-ABI export list.  No GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.linalg as sla

import cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [1, 2, 3, 10, 57, 101])
def test_tridiagonal_eigen_matches_lapack(n):
    rng = np.random.default_rng(n)
    a = rng.normal(size=n)
    b = rng.normal(size=max(n - 1, 0))
    w, z = pkg.host_tridiagonal_eigen(a, b)
    T = np.diag(a) + np.diag(b, 1) + np.diag(b, -1)
    wl = np.linalg.eigvalsh(T)
    np.testing.assert_allclose(w, wl, atol=1e-13 * max(1, np.abs(wl).max()))
    np.testing.assert_allclose(z.T @ z, np.eye(n), atol=1e-13)
    np.testing.assert_allclose(T @ z, z * w, atol=1e-12)
    w2 = pkg.host_tridiagonal_eigen(a, b, vectors=False)
    np.testing.assert_allclose(w2, wl, atol=1e-13 * max(1, np.abs(wl).max()))


def test_tridiagonal_eigen_lanczos_like_and_extra_beta():
    # Lanczos T of a Laplacian: graded, nearly-degenerate Ritz values; beta may carry one extra entry
    n = 80
    a = 4 + 0.01 * np.cos(np.arange(n))
    b = np.full(n, 2.0) * (1 + 1e-8 * np.arange(n))  # n entries: the last one must be ignored
    w, z = pkg.host_tridiagonal_eigen(a, b)
    wl = sla.eigh_tridiagonal(a, b[: n - 1], eigvals_only=True)
    np.testing.assert_allclose(w, wl, atol=1e-13)
    # zero off-diagonal (breakdown) splits the matrix
    b2 = b.copy()
    b2[30] = 0.0
    w3 = pkg.host_tridiagonal_eigen(a, b2, vectors=False)
    np.testing.assert_allclose(w3, sla.eigh_tridiagonal(a, b2[: n - 1], eigvals_only=True), atol=1e-13)


@pytest.mark.parametrize("n,cplx", [(1, True), (2, True), (5, False), (20, True), (50, False), (64, True)])
def test_hessenberg_eigen_matches_lapack(n, cplx):
    rng = np.random.default_rng(100 + n)
    h = rng.normal(size=(n, n))
    if cplx:
        h = h + 1j * rng.normal(size=(n, n))
    h = np.triu(h, -1)
    w, v = pkg.host_hessenberg_eigen(h)
    wl = np.linalg.eigvals(h)
    # compare as multisets
    d = np.abs(w[:, None] - wl[None, :])
    assert d.min(axis=1).max() < 1e-10 and d.min(axis=0).max() < 1e-10
    np.testing.assert_allclose(np.linalg.norm(v, axis=0), 1.0, atol=1e-12)
    res = np.linalg.norm(h @ v - v * w, axis=0)
    assert res.max() < 1e-9 * max(1.0, np.abs(h).max() * n)
    w2 = pkg.host_hessenberg_eigen(h, vectors=False)
    d2 = np.abs(w2[:, None] - wl[None, :])
    assert d2.min(axis=1).max() < 1e-10 and d2.min(axis=0).max() < 1e-10


def test_hessenberg_eigen_real_nonsymmetric_conjugate_pairs():
    # real Hessenberg with complex-conjugate eigenvalues embedded in complex arithmetic
    h = np.array([[0.0, -1.0, 0.3], [1.0, 0.0, 0.2], [0.0, 0.5, 2.0]])
    w, v = pkg.host_hessenberg_eigen(h)
    wl = np.linalg.eigvals(h)
    assert np.abs(np.sort_complex(w) - np.sort_complex(wl)).max() < 1e-12
    assert np.abs(h @ v - v * w).max() < 1e-12


@pytest.mark.parametrize("n,kind", [(1, "complex"), (2, "real"), (7, "complex"), (30, "real"), (45, "complex"), (30, "arrow")])
def test_general_eigen_matches_lapack(n, kind):
    """Householder reduction + Hessenberg QR (detail::general_eigen): what ThickRestartArnoldi solves after a restart,
    when the projected matrix is a full block with a spike row, and what the Eigen stand-in's ComplexEigenSolver runs."""
    rng = np.random.default_rng(500 + n)
    a = rng.normal(size=(n, n)).astype(np.complex128)
    if kind == "complex":
        a = a + 1j * rng.normal(size=(n, n))
    if kind == "arrow":  # kept Schur block (upper triangular) + spike row + fresh Hessenberg part
        k = 8
        a = np.triu(a, -1)
        a[:k, :k] = np.triu(a[:k, :k])
        a[k, :k] = rng.normal(size=k)
    w, v = pkg.host_general_eigen(a)
    wl = np.linalg.eigvals(a)
    d = np.abs(w[:, None] - wl[None, :])
    assert d.min(axis=1).max() < 1e-10 and d.min(axis=0).max() < 1e-10
    np.testing.assert_allclose(np.linalg.norm(v, axis=0), 1.0, atol=1e-12)
    assert np.linalg.norm(a @ v - v * w, axis=0).max() < 1e-9 * max(1.0, np.abs(a).max() * n)
    w2 = pkg.host_general_eigen(a, vectors=False)
    d2 = np.abs(w2[:, None] - wl[None, :])
    assert d2.min(axis=1).max() < 1e-10 and d2.min(axis=0).max() < 1e-10


def _declared_in_headers():
    names = set()
    for hdr in ("cmpt_b200.h", "cmpt_b200_solver.h", "cmpt_b200_debug.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"\b(cmbs?_[a-z0-9_]+)\s*\(", src))
    names.discard("cmb_matmul_fn")
    return names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "libcmpt_b200.so must be built by __graft_entry__.build()"
    L = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared_in_headers()
    assert len(declared) > 60
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    # the ctypes binding covers the same list
    assert set(capi.declared_symbols()) == declared


def test_no_silent_cpu_fallback_without_gpu():
    n = ctypes.c_int(-1)
    capi.check(capi.lib().cmb_device_count(ctypes.byref(n)))
    if n.value > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.CmbError) as e:
        pkg.Context(0)
    assert e.value.code == capi.CMB_ERR_NO_DEVICE
    assert b"no CPU fallback" in capi.lib().cmb_last_error()


def test_bench_and_smoke_fail_loudly_without_gpu():
    """The product path must not fall back to the oracle: without a GPU bench.py's own arm and smoke() exit non-zero
    with the library's "no CPU fallback" error (the oracle is only the checker / the reference arm)."""
    import os
    import subprocess
    import sys

    n = ctypes.c_int(-1)
    capi.check(capi.lib().cmb_device_count(ctypes.byref(n)))
    if n.value > 0:
        pytest.skip("a GPU is visible")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0", "--grid", "64"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr and not any(ln.startswith("{") for ln in p.stdout.splitlines())
    p = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, capture_output=True,
                       text=True, timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


@pytest.mark.parametrize("L,pbc,nranks", [(10, True, 2), (12, True, 4), (12, False, 4), (30, True, 8), (20, True, 16),
                                          (9, False, 1)])
def test_heisenberg_exchange_plan_is_consistent_between_partners(L, pbc, nranks):
    """Host-only: the slab exchange of the matrix-free Heisenberg chain is pairwise and symmetric (what lets a rank
    compute where its slab lands in the partner's receive buffer, csrc/heisenberg.cu setup_p2p), and it covers exactly
    the bonds that cross the rank bits."""
    import ctypes as C

    lib = capi.lib()

    def plan(rank):
        arr = [(C.c_int32 * 8)() for _ in range(5)]
        nrem = lib.cmb_heisenberg_plan(L, int(pbc), nranks, rank, *arr)
        assert nrem >= 0
        return [tuple(int(a[k]) for a in arr) for k in range(nrem)]

    p = nranks.bit_length() - 1
    plans = [plan(r) for r in range(nranks)]
    if nranks == 1:
        assert plans[0] == []
        return
    nb = L if (pbc and L > 2) else L - 1
    crossing = p + (1 if nb == L else 0)  # bonds (Ll-1,Ll) .. (L-2,L-1) and the periodic wrap bond
    for r, pl in enumerate(plans):
        assert len(pl) == crossing
        off = 0
        for k, (kind, partner, needed, offset, length) in enumerate(pl):
            assert 0 <= partner < nranks and partner != r
            assert offset == off
            off += length
            # the partner lists the same bond at the same position, points back at me, and agrees on the size
            kind2, partner2, needed2, _, length2 = plans[partner][k]
            assert (kind2, partner2, needed2, length2) == (kind, r, needed, length)
            if kind == 2:      # straddle bond: partner differs in rank bit 0, half slab
                assert partner == r ^ 1 and needed == 1 and length == 1
            elif kind == 3:    # rank-rank bond b: full slab iff my two rank bits differ
                b = k - 1
                differ = ((r >> b) ^ (r >> (b + 1))) & 1
                assert partner == r ^ (3 << b) and needed == differ and length == 2 * differ
            else:              # periodic wrap: partner differs in the top rank bit, packed half slab
                assert kind == 4 and partner == r ^ (1 << (p - 1)) and needed == 1 and length == 1
    # the heaviest ranks of cfg 5 (L = 30 on 8 GPUs) receive three slabs' worth per apply
    if (L, nranks) == (30, 8):
        assert max(sum(e[4] for e in pl) for pl in plans) == 6 and min(sum(e[4] for e in pl) for pl in plans) == 2


@pytest.mark.parametrize("n", [1, 2, 3, 7, 24, 61])
def test_symmetric_jacobi_matches_lapack(n):
    rng = np.random.default_rng(100 + n)
    a = rng.normal(size=(n, n))
    a = (a + a.T) / 2
    w, z = pkg.host_symmetric_eigen(a)
    np.testing.assert_allclose(w, np.linalg.eigvalsh(a), atol=1e-13 * max(1.0, np.abs(a).max()) * n)
    assert np.abs(z.T @ z - np.eye(n)).max() < 1e-13
    assert np.abs(a @ z - z * w).max() < 1e-12 * max(1.0, np.abs(a).max()) * n


def test_symmetric_jacobi_on_thick_restart_arrowhead():
    """The projected matrix after a thick restart: diag(theta) with an arrow at row/column k, tridiagonal tail."""
    k, m = 6, 20
    rng = np.random.default_rng(5)
    t = np.diag(np.sort(rng.uniform(0, 1, m)))
    t[:k, k] = t[k, :k] = 1e-3 * rng.normal(size=k)
    for i in range(k, m - 1):
        t[i, i + 1] = t[i + 1, i] = rng.uniform(0.1, 0.5)
    w, z = pkg.host_symmetric_eigen(t)
    np.testing.assert_allclose(w, np.linalg.eigvalsh(t), atol=1e-14)
    assert np.abs(t @ z - z * w).max() < 1e-14
    # degenerate and tiny couplings do not stall the sweeps
    t2 = np.diag([1.0, 1.0, 1.0, 2.0])
    t2[0, 3] = t2[3, 0] = 1e-200
    w2, _ = pkg.host_symmetric_eigen(t2)
    np.testing.assert_allclose(w2, [1, 1, 1, 2], atol=1e-15)
