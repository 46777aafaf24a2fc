"""Pins the CPU oracle on the reference ITSELF.

oracle/_ref/libref.so is /root/reference/include/cmpt/eigen_ex/{lanczos,arnoldi}.hpp compiled UNMODIFIED
against oracle/eigen_shim/ (recipe: oracle/Makefile target `ref`, harness: oracle/ref_harness.cpp).  Every test
here runs the restatement (oracle/krylov_oracle.cpp + oracle/reference_solvers.py) and the reference's own
classes through the same call sequence on the same operator routine and the same start vector and demands
agreement to rounding (alpha/beta/H ~1e-14, identical iteration counts, identical log strings, convergence logs
to ~1e-13).  The call sequences are those of the reference's samples (src/samples/sample_lanczos1.cpp,
sample_lanczos2.cpp, sample_arnoldi.cpp) plus seeded Laplacian / Heisenberg / convection-diffusion cases.

CPU only.  In the authoring container the library is rebuilt from /root/reference; on a box without the
reference tree the prebuilt library that travelled with the snapshot is used.
"""
import os
import subprocess

import numpy as np
import pytest

from cmpt_eigenex_b200 import synthetic as syn
from oracle import core, ref
from oracle import reference_solvers as rs

pytestmark = pytest.mark.skipif(not ref.available(), reason="neither oracle/_ref/libref.so nor /root/reference present")

core.set_num_threads(1)
TIGHT = 5e-14


@pytest.fixture(scope="module", autouse=True)
def _single_thread():
    ref.set_num_threads(1)
    yield


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


# ------------------------------------------------------------------------------------------------------
# start vectors: libstdc++ mt19937 + normal_distribution stream through the reference's own random.hpp
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", ["d", "z"])
def test_default_and_seeded_start_vectors(p):
    for n in (1, 5, 200, 4097):
        assert _rel(core.default_vector(n, p), ref.default_vector(n, p)) < 1e-15
        assert _rel(core.default_vector(n, p), ref.arnoldi_default_vector(n, p)) < 1e-15
        for seed in (1, 12345):
            assert _rel(core.seeded_vector(seed, n, p), ref.seeded_vector(seed, n, p)) < 1e-15
    assert abs(np.linalg.norm(ref.default_vector(300, p)) - 1) < 1e-14


# ------------------------------------------------------------------------------------------------------
# step level: LanczosBase
# ------------------------------------------------------------------------------------------------------
def _lanczos_cases():
    rp, c, v = syn.laplacian2d_csr(24)
    yield "laplacian24", core.Operator.csr(rp, c, v), "d"
    yield "dense300", core.Operator.dense(syn.dense_symmetric(300, seed=1)), "d"
    rp, c, v = syn.hermitian_chain_csr(200)
    yield "hermitian_chain200", core.Operator.csr(rp, c, v), "z"
    yield "heisenberg10", core.Operator.heisenberg(10, 1.0, True, "d"), "d"
    yield "heisenberg8z", core.Operator.heisenberg(8, 1.0, False, "z"), "z"


def _start(n, p, seed=7):
    x = syn.start_vector(n, seed=seed)
    if p == "z":
        x = x + 1j * syn.start_vector(n, seed=seed + 100)
    return x


@pytest.mark.parametrize("interval,shift,ndefl", [(1, 0.0, 0), (1, 0.75, 2), (0, 0.0, 0), (3, -0.5, 1), (2, 0.0, 0)])
def test_lanczos_base_steps_match_reference(interval, shift, ndefl):
    for name, op, p in _lanczos_cases():
        n = op.n
        a, b = core.LanczosBase(p), ref.LanczosBase(p)
        defl = []
        rng = np.random.default_rng(3)
        for _ in range(ndefl):  # orthonormal deflation vectors
            w = rng.standard_normal(n) + (1j * rng.standard_normal(n) if p == "z" else 0)
            for d in defl:
                w = w - np.vdot(d, w) * d
            defl.append(w / np.linalg.norm(w))
        for s in (a, b):
            s.set_op(op)
            s.set_params(shift, interval, 1e-12)
            s.set_init(_start(n, p))
            for d in defl:
                s.add_ortho(d)
        # the Heisenberg ring's spectrum is highly degenerate: a generic start vector spans a small Krylov space and
        # the recurrence runs into near-breakdown (tiny beta) soon after 20 steps, where rounding decides the digits
        steps = 18 if name.startswith("heisenberg") else 40
        for k in range(steps):
            ra, rb = a.step(), b.step()
            assert ra == rb, (name, k)
            assert a.utmost() == b.utmost() and a.iterations == b.iterations and a.nvectors == b.nvectors
        (aa, ab), (ba, bb) = a.alpha_beta(), b.alpha_beta()
        # without full reorthogonalisation rounding differences are amplified along the recurrence
        # (summation order differs: blocked sums in the stand-in Eigen, sequential sums in the restatement); with it
        # the leading steps agree to a few ulp and the rest to ~1e-13
        assert _rel(aa[:10], ba[:10]) < TIGHT and _rel(ab[:10], bb[:10]) < TIGHT, (name, _rel(aa[:10], ba[:10]))
        tol = 1e-12 if interval == 1 else 1e-9
        assert _rel(aa, ba) < tol and _rel(ab, bb) < tol, (name, _rel(aa, ba), _rel(ab, bb))
        # basis vectors: early ones to rounding; late ones are only compared where no beta has become small (a tiny
        # beta normalises a vector of rounding noise and the two summation orders then legitimately differ)
        last = a.nvectors - 1 if ab.min() > 1e-3 * ab.max() else 12
        for k in (0, 1, last):
            assert _rel(a.vector(k), b.vector(k)) < (1e-11 if interval == 1 else 1e-7), (name, k)


def test_lanczos_base_breakdown_and_quiet_failures():
    H = np.array([[1.0, 0.5, 0.0], [0.5, 2.0, 0.5], [0.0, 0.5, 3.0]])
    op = core.Operator.dense(H)
    a, b = core.LanczosBase("d"), ref.LanczosBase("d")
    for s in (a, b):
        s.set_op(op)
        s.set_init(np.array([1.0, 2.0, 3.0]))
    seq_a = [a.step() for _ in range(6)]
    seq_b = [b.step() for _ in range(6)]
    assert seq_a == seq_b
    (aa, ab), (ba, bb) = a.alpha_beta(), b.alpha_beta()
    # once the space is exhausted every further call appends another (tiny) beta: lanczos.hpp:429-436
    assert len(aa) == len(ba) == 3 and len(ab) == len(bb)
    assert _rel(aa, ba) < TIGHT and _rel(ab[:2], bb[:2]) < TIGHT
    assert a.utmost() and b.utmost()
    # zero start vector: the basis stays empty and the step reports failure (lanczos.hpp:315-318,382-384)
    a, b = core.LanczosBase("d"), ref.LanczosBase("d")
    for s in (a, b):
        s.set_op(op)
        s.set_init(np.zeros(3))
        assert s.step() is False and s.nvectors == 0
    # no operator / height 0 (lanczos.hpp:372-377)
    a, b = core.LanczosBase("d"), ref.LanczosBase("d")
    assert a.step() is False and b.step() is False


# ------------------------------------------------------------------------------------------------------
# solver level: LanczosEigenSolver with the samples' call sequences
# ------------------------------------------------------------------------------------------------------
def _same_solver_state(a, b, vec_tol=1e-10):
    assert a.iterations == b.iterations
    assert a.log == b.log
    (aa, ab), (ba, bb) = a.alpha_beta(), b.alpha_beta()
    assert _rel(aa, ba) < 1e-12 and _rel(ab, bb) < 1e-12, (_rel(aa, ba), _rel(ab, bb))
    assert _rel(aa[:10], ba[:10]) < TIGHT and _rel(ab[:10], bb[:10]) < TIGHT
    assert _rel(a.eigenvalues, b.eigenvalues) < 1e-12
    assert sorted(a.convergence_log) == sorted(b.convergence_log)
    for k in a.convergence_log:
        assert _rel(a.convergence_log[k], b.convergence_log[k]) < 1e-12, k
    assert a.eigenvectors.shape == b.eigenvectors.shape
    if a.eigenvectors.size:
        # both are phase-fixed the same way, so the vectors themselves must agree
        assert _rel(a.eigenvectors, b.eigenvectors) < vec_tol


def test_sample_lanczos1_sequence():
    # src/samples/sample_lanczos1.cpp:13-32 — default (random, fixed-seed) start vector
    H = np.array([[1.0, 0.5, 0.0], [0.5, 2.0, 0.5], [0.0, 0.5, 3.0]])
    out = []
    for mod in (rs, ref):
        es = mod.LanczosEigenSolver("d")
        es.set_matrix_multiplication(core.Operator.dense(H))
        es.tolerance = 1.0e-5
        es.max_iterations = 100
        assert es.compute() == 0
        out.append(es)
    _same_solver_state(*out)
    np.testing.assert_allclose(out[1].eigenvalues, [2 - np.sqrt(1.5), 2.0, 2 + np.sqrt(1.5)], atol=1e-14)
    assert out[1].log[0] == rs.HEAD_INFO + "in compute(), initial_vector is empty or invalid, then set at random"


@pytest.mark.parametrize("settings", ["sample", "fixed_m", "negative_indices"])
def test_sample_lanczos2_sequence(settings):
    # src/samples/sample_lanczos2.cpp:22-59
    rp, c, v = syn.hermitian_chain_csr(200)
    op = core.Operator.csr(rp, c, v)
    out = []
    for mod in (rs, ref):
        es = mod.LanczosEigenSolver("z")
        es.set_matrix_multiplication(op)
        es.shift = 0.0
        es.threshold = 1.0e-14
        es.compute_eigenvectors_on = True
        es.max_eigenvalues = 10
        es.interval = 1
        if settings == "sample":
            es.tolerance = 1.0e-7
            es.min_iterations = -1
            es.max_iterations = 1000
            es.indices_for_convergence = [0]
        elif settings == "fixed_m":
            es.min_iterations = es.max_iterations = 60
            es.indices_for_convergence = [0, 1, 2]
        else:
            es.tolerance = 1.0e-9
            es.max_iterations = 150
            es.indices_for_convergence = [0, -1, 3, 400]  # 400 is never in range: lanczos.hpp:857-859
        if mod is ref:
            es.set_init_seeded(1)
        else:
            es.init = core.seeded_vector(1, 200, "z")
        es.compute()
        out.append(es)
    _same_solver_state(*out)
    if settings == "sample":
        assert out[1].log[0] == rs.HEAD_INFO + "lanczos steps converged with tolerance"


def test_lanczos_solver_laplacian_deflation_shift_continue():
    rp, c, v = syn.laplacian2d_csr(20)
    op = core.Operator.csr(rp, c, v)
    n = op.n
    # first run: lowest pair, then deflate it and continue in two legs
    first = []
    for mod in (rs, ref):
        es = mod.LanczosEigenSolver("d")
        es.set_matrix_multiplication(op)
        es.init = _start(n, "d")
        es.min_iterations = es.max_iterations = 80
        es.max_eigenvalues = 2
        es.compute()
        first.append(es)
    _same_solver_state(*first)
    out = []
    for mod, prev in zip((rs, ref), first):
        es = mod.LanczosEigenSolver("d")
        es.set_matrix_multiplication(op)
        es.init = _start(n, "d", seed=11)
        es.ortho = [prev.eigenvectors[:, 0].copy()]
        es.shift = 0.3
        es.min_iterations = es.max_iterations = 30
        es.max_eigenvalues = 3
        es.indices_for_convergence = [0, 1]
        es.compute()
        es.min_iterations = es.max_iterations = 55  # continueToCompute from the current state: lanczos.hpp:701-712
        es.continue_to_compute()
        out.append(es)
    _same_solver_state(*out)
    assert out[1].log[0].endswith("continueToCompute(...) was called") or "continueToCompute" in " ".join(out[1].log)
    # the deflated run finds the second level first
    assert abs(out[1].eigenvalues[0] - syn.laplacian2d_eigenvalues(20, 2)[1]) < 1e-6


def test_lanczos_solver_quiet_failure_and_full_krylov():
    H = syn.dense_symmetric(12, seed=5)
    op = core.Operator.dense(H)
    out = []
    for mod in (rs, ref):
        es = mod.LanczosEigenSolver("d")
        es.set_matrix_multiplication(op)
        es.init = np.zeros(12)
        es.compute()
        out.append(es)
    assert out[0].log == out[1].log and out[1].log[0].endswith("initial lanczosvector generation fail")
    assert out[1].eigenvalues.size == 0 and out[0].eigenvalues.size == 0
    out = []
    for mod in (rs, ref):
        es = mod.LanczosEigenSolver("d")
        es.set_matrix_multiplication(op)
        es.init = _start(12, "d")
        es.tolerance = 0.0
        es.compute()
        out.append(es)
    _same_solver_state(*out, vec_tol=1e-8)
    np.testing.assert_allclose(out[1].eigenvalues, np.linalg.eigvalsh(H), atol=1e-12)


def test_heisenberg_ground_state_stop_rule():
    # the reference stop rule (tolerance 1e-12, index 0) fires at iterations() == 30 for L = 12
    op = core.Operator.heisenberg(12, 1.0, True, "d")
    out = []
    for mod in (rs, ref):
        es = mod.LanczosEigenSolver("d")
        es.set_matrix_multiplication(op)
        es.init = _start(op.n, "d")
        es.max_iterations = 200
        es.max_eigenvalues = 1
        es.compute_eigenvectors_on = False
        es.compute()
        out.append(es)
    _same_solver_state(*out)
    assert abs(out[1].eigenvalues[0] - (-5.387390917445)) < 1e-10


# ------------------------------------------------------------------------------------------------------
# exponential solver (lanczos.hpp:1004-1164)
# ------------------------------------------------------------------------------------------------------
def test_exponential_solver_matches_reference():
    rp, c, v = syn.hermitian_chain_csr(60)
    op = core.Operator.csr(rp, c, v)
    x = -0.4 + 0.3j
    vin = _start(60, "z")
    got = []
    for mod in (rs, ref):
        es = mod.LanczosEigenSolver("z")
        es.set_matrix_multiplication(op)
        es.init = vin
        es.min_iterations = es.max_iterations = 40
        if mod is ref:
            got.append(es.exp_with_lanczos(x))
        else:
            got.append(rs.exp_solve_with_lanczos(x, es))
    assert _rel(got[0], got[1]) < 1e-12
    A = np.zeros((60, 60), complex)
    for i in range(60):
        A[:, i] = op.apply(np.eye(60, dtype=complex)[i])
    w, y = np.linalg.eigh(A)
    a = rs.exp_solve_with_eigens(x, w, y, 60, vin)
    b = ref.exp_with_eigens(x, w, y, 60, vin, "z")
    assert _rel(a, b) < 1e-13
    # Taylor variants against the dense exponential
    import scipy.linalg as sla

    want = sla.expm(x * A) @ vin
    t = ref.exp_taylor(x, op, 2.0, vin)
    assert _rel(t, want) < 1e-12
    # AutoDivision applies every sub-step to `in` (lanczos.hpp:1147-1159): the result is exp(x/div A) in
    rad = abs(x * 2.0)
    div = int(rad + 1.0)
    t2 = ref.exp_taylor(x, op, 2.0, vin, auto_division=True)
    assert _rel(t2, sla.expm(x / div * A) @ vin) < 1e-12


# ------------------------------------------------------------------------------------------------------
# Arnoldi
# ------------------------------------------------------------------------------------------------------
def _arnoldi_ops():
    rng = np.random.default_rng(0)
    A = rng.uniform(-1, 1, (50, 50)) + 1j * rng.uniform(-1, 1, (50, 50))  # sample_arnoldi.cpp:26 (n = 50, random complex)
    yield "random50", core.Operator.dense(A), "z"
    rp, c, v = syn.convdiff3d_csr(8)
    yield "convdiff8", core.Operator.csr(rp, c, v.astype(complex)), "z"
    yield "convdiff8_real", core.Operator.csr(rp, c, v), "d"


@pytest.mark.parametrize("shift,ndefl", [(0.0, 0), (0.4, 1)])
def test_arnoldi_base_steps_match_reference(shift, ndefl):
    for name, op, p in _arnoldi_ops():
        n = op.n
        a, b = core.ArnoldiBase(p), ref.ArnoldiBase(p)
        defl = []
        rng = np.random.default_rng(4)
        for _ in range(ndefl):
            w = rng.standard_normal(n) + (1j * rng.standard_normal(n) if p == "z" else 0)
            defl.append(w / np.linalg.norm(w))
        for s in (a, b):
            s.set_op(op)
            s.set_params(shift, 1e-12)
            s.set_init(_start(n, p))
            for d in defl:
                s.add_ortho(d)
        for k in range(30 if name == "random50" else 16):
            assert a.step() == b.step()
            assert a.iterations == b.iterations and a.nvectors == b.nvectors and a.utmost() == b.utmost()
            assert abs(a.residue - b.residue) < 1e-12 * max(1.0, abs(b.residue)), (name, k)
        Ha, Hb = a.hessenberg(), b.hessenberg()
        # single-pass MGS on a non-normal operator: rounding differences grow along the columns, the leading block is tight
        assert _rel(Ha[:8, :8], Hb[:8, :8]) < TIGHT, (name, _rel(Ha[:8, :8], Hb[:8, :8]))
        assert _rel(Ha, Hb) < 1e-9, (name, _rel(Ha, Hb))
        assert _rel(a.vector(1), b.vector(1)) < 1e-12


def test_arnoldi_base_full_space_and_zero_start():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    op = core.Operator.dense(A)
    a, b = core.ArnoldiBase("z"), ref.ArnoldiBase("z")
    for s in (a, b):
        s.set_op(op)
        s.set_init(_start(4, "z"))
    assert [a.step() for _ in range(7)] == [b.step() for _ in range(7)]
    assert a.utmost() and b.utmost() and a.nvectors == b.nvectors == 4
    assert _rel(a.hessenberg(), b.hessenberg()) < 1e-12
    a, b = core.ArnoldiBase("z"), ref.ArnoldiBase("z")
    for s in (a, b):
        s.set_op(op)
        s.set_init(np.zeros(4, complex))
        assert s.step() is False and s.nvectors == 0 and s.utmost() is False


@pytest.mark.parametrize("case", ["sample", "converge", "full"])
def test_sample_arnoldi_sequence(case):
    # src/samples/sample_arnoldi.cpp:21-41: n = 50, m = 40, min = max = m, tolerance 1e-14, 2 eigenvalues
    rng = np.random.default_rng(0)
    n = 50 if case != "full" else 9
    A = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
    op = core.Operator.dense(A)
    out = []
    for mod in (rs, ref):
        es = mod.ArnoldiEigenSolver("z")
        es.set_matrix_multiplication(op)
        if case == "sample":
            es.min_iterations = es.max_iterations = 40
            es.tolerance = 1.0e-14
            es.max_eigenvalues = 2
        elif case == "converge":
            es.init = _start(n, "z")
            es.tolerance = 1.0e-9
            es.max_iterations = 49
            es.max_eigenvalues = 3
            es.indices_for_convergence = [0, 1, -1]
            es.shift = 0.25 - 0.5j
        else:  # arnoldi_test.cpp:56-76: unlimited iterations until the Krylov space is full
            es.threshold = 1.0e-14
            es.min_iterations = es.max_iterations = -1
            es.max_eigenvalues = 5
            es.tolerance = 1.0e-10
        es.compute()
        out.append(es)
    a, b = out
    assert a.iterations == b.iterations and a.log == b.log
    assert _rel(a.hessenberg[:8, :8], b.hessenberg[:8, :8]) < TIGHT
    # LAPACK (restatement) vs the host Hessenberg QR (reference build): eigenvalues of the same H to ~eps*cond
    assert _rel(a.eigenvalues, b.eigenvalues) < 1e-9
    assert sorted(a.convergence_log) == sorted(b.convergence_log)
    for k in a.convergence_log:
        assert len(a.convergence_log[k]) == len(b.convergence_log[k])
        assert _rel(a.convergence_log[k][-3:], b.convergence_log[k][-3:]) < 1e-8
    assert a.eigenvectors.shape == b.eigenvectors.shape
    for j in range(a.eigenvectors.shape[1]):
        assert abs(abs(np.vdot(a.eigenvectors[:, j], b.eigenvectors[:, j])) - 1) < 1e-7
        # A x = lambda x to the accuracy the Krylov space allows, identically for both
        ra = np.linalg.norm(A @ a.eigenvectors[:, j] - a.eigenvalues[j] * a.eigenvectors[:, j])
        rb = np.linalg.norm(A @ b.eigenvectors[:, j] - b.eigenvalues[j] * b.eigenvectors[:, j])
        assert abs(ra - rb) < 1e-6 * max(1.0, ra)
    if case == "full":
        np.testing.assert_allclose(sorted(b.eigenvalues, key=abs, reverse=True)[:5],
                                   sorted(np.linalg.eigvals(A), key=abs, reverse=True)[:5], atol=1e-10)


# ------------------------------------------------------------------------------------------------------
# the reference's sample programs, built as they are
# ------------------------------------------------------------------------------------------------------
def _run_sample(name):
    exe = os.path.join(os.path.dirname(ref._SO), name)
    if not os.path.exists(exe):
        pytest.skip(exe + " not built")
    return subprocess.run([exe], capture_output=True, text=True, check=True, timeout=120).stdout


def test_reference_sample_programs_run():
    ref.lib()
    out = _run_sample("sample_lanczos1")
    assert "0.775255" in out and "3.22474" in out
    assert "INFO      lanczos steps achieved full of Krylov subspace" in out
    out = _run_sample("sample_lanczos2")
    assert "iterations : 143" in out and "subspace rank : 144" in out
    assert "INFO      lanczos steps converged with tolerance" in out
    vals = [float(x) for x in out.split("eigen values :")[1].split("log :")[0].split()]
    # same settings through the harness
    rp, c, v = syn.hermitian_chain_csr(200)
    es = ref.LanczosEigenSolver("z")
    es.set_matrix_multiplication(core.Operator.csr(rp, c, v))
    es.tolerance, es.threshold, es.min_iterations, es.max_iterations, es.max_eigenvalues = 1e-7, 1e-14, -1, 1000, 10
    es.set_init_seeded(1)
    es.compute()
    assert es.iterations == 143
    np.testing.assert_allclose(vals, es.eigenvalues, atol=1e-5)  # printed with 6 significant digits
    out = _run_sample("sample_arnoldi")
    assert out.startswith("AP-PD")
