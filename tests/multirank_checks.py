"""Checks of the row-partitioned (multi-rank) code paths against the CPU checker, written once and run two ways:

  * real ranks   — one process per GPU under torch.distributed.run (tests/dist_worker.py, tests/test_dist.py),
  * virtual ranks — P host threads driving P virtual ranks on ONE GPU (tests/test_virtual_ranks.py,
                    cmpt_b200_debug.h): same kernels, mailboxes, halo flags and sequence numbers as real ranks.

`expected(...)` runs the checker (the restatement oracle/reference_solvers.py, or the reference itself through
oracle/ref.py) on the full problems once; `run_checks(...)` is what every rank executes.  comm must offer rank,
world, row_range(n), gather(obj) -> list over ranks, barrier().
"""
import numpy as np

from cmpt_eigenex_b200 import synthetic as syn


def _tridiag_csr(nb):
    rp = np.zeros(nb + 1, np.int64)
    cols, vals = [], []
    for r in range(nb):
        for c_, v_ in ((r - 1, -1.0), (r, 2.0 + 0.1 * r), (r + 1, -1.0)):
            if 0 <= c_ < nb:
                cols.append(c_)
                vals.append(v_)
        rp[r + 1] = len(cols)
    return rp, np.array(cols, np.int32), np.array(vals)


def _dense_to_csr(A):
    n = A.shape[0]
    rp = np.zeros(n + 1, np.int64)
    cols, vals = [], []
    for r in range(n):
        nz = np.nonzero(A[r])[0]
        cols += nz.tolist()
        vals += A[r, nz].tolist()
        rp[r + 1] = len(cols)
    return rp, np.array(cols, np.int32), np.array(vals)


def _csr_dense(full, n):
    rp, c, v = full
    A = np.zeros((n, n))
    for r in range(n):
        A[r, c[rp[r]:rp[r + 1]]] = v[rp[r]:rp[r + 1]]
    return A


def shard_of(full, r0, r1):
    rp, c, v = full
    return rp[r0:r1 + 1] - rp[r0], c[rp[r0]:rp[r1]], v[rp[r0]:rp[r1]]


def invariant_subspace_operator(n=96, k=12, stride=7):
    """Symmetric operator whose Krylov space from the returned start vector has dimension exactly k: a path graph over
    the k indices S = {0, stride, 2 stride, ...} (spread over all ranks) decoupled from a path over the other indices;
    the start vector lives on S.  Lanczos must break down at step k (beta_k = rounding noise <= threshold)."""
    S = [(i * stride) % n for i in range(k)]
    rest = [i for i in range(n) if i not in S]
    A = np.zeros((n, n))
    for chain, d0 in ((S, 1.0), (rest, 3.0)):
        for a, i in enumerate(chain):
            A[i, i] = d0 + 0.37 * a
            if a + 1 < len(chain):
                A[i, chain[a + 1]] = A[chain[a + 1], i] = -1.0 - 0.05 * a
    x0 = np.zeros(n)
    x0[S] = 1.0 + 0.1 * np.arange(k)
    return A, x0 / np.linalg.norm(x0), S


def problems():
    """name -> dict(full=(rp, c, v), n, ...) of the CSR problems (deterministic, small)."""
    out = {}
    out["laplacian"] = dict(full=syn.laplacian2d_csr(40), n=1600, m=60, kind="lanczos")
    out["heisenberg"] = dict(full=syn.heisenberg_csr(12), n=4096, m=40, kind="lanczos")
    out["convdiff_arnoldi"] = dict(full=syn.convdiff3d_csr(9), n=729, m=30, kind="arnoldi")
    out["exhaust"] = dict(full=_tridiag_csr(24), n=24, kind="exhaust")
    A, x0, _ = invariant_subspace_operator()
    out["breakdown_at_k"] = dict(full=_dense_to_csr(A), n=A.shape[0], x0=x0, k=12, kind="breakdown")
    out["deflation"] = dict(full=syn.laplacian2d_csr(16), n=256, kind="deflation")
    return out


def deflation_vectors(N=16):
    ii, jj = np.meshgrid(np.arange(1, N + 1), np.arange(1, N + 1), indexing="ij")
    defl = []
    for (p_, q_) in ((1, 1), (1, 2)):
        e = (np.sin(p_ * np.pi * ii / (N + 1)) * np.sin(q_ * np.pi * jj / (N + 1))).reshape(-1)
        defl.append(e / np.linalg.norm(e))
    lam = np.sort((4 - 2 * np.cos(ii * np.pi / (N + 1)) - 2 * np.cos(jj * np.pi / (N + 1))).reshape(-1))
    return defl, lam


def expected(rs, core, heisenberg_L=14):
    """Checker results on the full problems (rs: oracle.reference_solvers or oracle.ref)."""
    exp = {}
    P = problems()
    for name in ("laplacian", "heisenberg"):
        pr = P[name]
        x0 = syn.start_vector(pr["n"], seed=7)
        ref = rs.LanczosEigenSolver("d")
        ref.set_matrix_multiplication(core.Operator.csr(*pr["full"]))
        ref.init = x0
        ref.min_iterations = ref.max_iterations = pr["m"]
        ref.max_eigenvalues = 3
        ref.indices_for_convergence = [0, 1, 2]
        ref.compute()
        ra, rb = ref.alpha_beta()
        theta, S = _tri(ref)
        exp[name] = dict(alpha=ra, beta=rb, eigenvalues=np.array(ref.eigenvalues), eigenvectors=np.array(ref.eigenvectors),
                         log=list(ref.log), apply=core.Operator.csr(*pr["full"]).apply(x0),
                         residuals=np.abs(rb[-1] * S[len(ra) - 1, :3]) if len(rb) >= len(ra) else None)
    pr = P["convdiff_arnoldi"]
    x0 = syn.start_vector(pr["n"], seed=7)
    if hasattr(rs, "UNLIMITED"):  # the restatement runs real Arnoldi; the reference's class only compiles for complex Scalar
        ref = rs.ArnoldiEigenSolver("d")
        ref.set_matrix_multiplication(core.Operator.csr(*pr["full"]))
    else:
        ref = rs.ArnoldiEigenSolver("z")
        rp, c, v = pr["full"]
        ref.set_matrix_multiplication(core.Operator.csr(rp, c, v.astype(complex)))
    ref.init = x0
    ref.min_iterations = ref.max_iterations = pr["m"]
    ref.max_eigenvalues = 2
    ref.compute()
    exp["convdiff_arnoldi"] = dict(eigenvalues=np.array(ref.eigenvalues), hessenberg=np.real(np.array(ref.hessenberg)),
                                   apply=core.Operator.csr(*pr["full"]).apply(x0))
    pr = P["exhaust"]
    x0 = syn.start_vector(pr["n"], seed=11)
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(*pr["full"]))
    ref.init, ref.min_iterations, ref.max_iterations, ref.max_eigenvalues = x0, 40, 60, 4
    ref.compute()
    exp["exhaust"] = dict(log=list(ref.log), eigenvalues=np.array(ref.eigenvalues))
    pr = P["breakdown_at_k"]
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(*pr["full"]))
    ref.init, ref.min_iterations, ref.max_iterations, ref.max_eigenvalues = pr["x0"], 30, 40, 3
    ref.compute()
    ra, rb = ref.alpha_beta()
    exp["breakdown_at_k"] = dict(log=list(ref.log), eigenvalues=np.array(ref.eigenvalues), alpha=ra, beta=rb,
                                 iterations=ref.iterations)
    defl, lam = deflation_vectors()
    pr = P["deflation"]
    x0 = syn.start_vector(pr["n"], seed=13)
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(*pr["full"]))
    ref.init, ref.ortho = x0, defl
    ref.min_iterations = ref.max_iterations = 50
    ref.max_eigenvalues = 2
    ref.compute()
    ra, rb = ref.alpha_beta()
    if hasattr(rs, "UNLIMITED"):
        refa = rs.ArnoldiEigenSolver("d")
        refa.set_matrix_multiplication(core.Operator.csr(*pr["full"]))
        refa.init, refa.ortho = x0, defl
    else:
        refa = rs.ArnoldiEigenSolver("z")
        rp, c, v = pr["full"]
        refa.set_matrix_multiplication(core.Operator.csr(rp, c, v.astype(complex)))
        refa.init, refa.ortho = x0.astype(complex), [d.astype(complex) for d in defl]
    refa.min_iterations = refa.max_iterations = 20
    refa.max_eigenvalues = 1
    refa.compute()
    exp["deflation"] = dict(alpha=ra, beta=rb, arnoldi_hessenberg=np.real(np.array(refa.hessenberg)))
    n = 1 << heisenberg_L
    x0 = syn.start_vector(n, seed=7)
    opm = core.Operator.heisenberg(heisenberg_L)
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(opm)
    ref.init, ref.max_iterations, ref.max_eigenvalues = x0, 200, 1
    ref.compute()
    exp["heisenberg_mf"] = dict(L=heisenberg_L, apply=opm.apply(x0), iterations=ref.iterations, eigenvalues=np.array(ref.eigenvalues))
    return exp


def _tri(ref):
    if hasattr(ref, "tridiagonal_eigensystem"):
        return ref.tridiagonal_eigensystem()
    return ref._theta, ref._S


def gather_rows(comm, local):
    return np.concatenate(comm.gather(np.asarray(local)))


def run_checks(pkg, ctx, comm, exp, only=None):
    """Everything one rank does.  Returns a small summary dict.  only: optional subset of section names."""
    def want(name):
        return only is None or name in only

    P = problems()
    results = {}
    for name in ("laplacian", "heisenberg", "convdiff_arnoldi"):
        if not want(name):
            continue
        pr, ex = P[name], exp[name]
        n, m = pr["n"], pr["m"]
        r0, r1 = comm.row_range(n)
        x0 = syn.start_vector(n, seed=7)
        op = pkg.DeviceOperator.from_csr(ctx, *shard_of(pr["full"], r0, r1), n_global=n, row_begin=r0)
        y = op.apply(x0[r0:r1])  # operator apply: local slab of A x
        assert np.abs(y - ex["apply"][r0:r1]).max() < 1e-13, name
        if pr["kind"] == "lanczos":
            es = pkg.LanczosEigenSolver()
            es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1])
            es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(3).setIndicesForConvergence([0, 1, 2])
            es.compute()
            assert es.iterations() == m
            assert np.abs(es.alpha() - ex["alpha"]).max() < 1e-11 and np.abs(es.beta() - ex["beta"]).max() < 1e-11, name
            assert np.abs(es.eigenvalues() - ex["eigenvalues"]).max() < 1e-10 * np.abs(ex["alpha"]).max(), name
            X = es.eigenvectors()
            assert X.shape == (r1 - r0, 3)
            Xfull = gather_rows(comm, X)
            ov = np.abs(np.sum(ex["eigenvectors"] * Xfull, axis=0))
            assert np.abs(ov - 1).max() < 1e-8, (name, ov)
            assert np.all(Xfull[0] > 0)
            assert es.log() == ex["log"]
            if ex["residuals"] is not None:
                assert np.abs(es.ritzResiduals() - ex["residuals"]).max() < 1e-9
            results[name] = es.eigenvalues()
            es.close()
        else:
            es = pkg.ArnoldiEigenSolver(np.float64)
            es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1])
            es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(2)
            es.compute()
            assert abs(es.eigenvalues()[0] - ex["eigenvalues"][0]) < 1e-8 * abs(ex["eigenvalues"][0]), name
            H = es.hessenbergMatrix()
            assert np.abs(H[:, :6] - ex["hessenberg"][:, :6]).max() < 1e-11
            Pv = gather_rows(comm, es.eigenvectors())
            assert np.abs(np.linalg.norm(Pv, axis=0) - 1).max() < 1e-12
            A = _csr_dense(pr["full"], n)
            res = np.linalg.norm(A @ Pv - Pv * es.eigenvalues(), axis=0)
            assert np.all(np.abs(res - es.ritzResiduals()) < 1e-9)
            es.close()
        op.close()
    # Krylov space exhausted on a row-partitioned operator: every rank must take the same halting decision on the
    # device (beta^2 from the reduced coefficients) and the chain must stop with the reference's log lines
    if want("exhaust"):
        _check_exhaust(pkg, ctx, comm, P["exhaust"], exp["exhaust"])
    if want("breakdown_at_k"):
        _check_breakdown(pkg, ctx, comm, P["breakdown_at_k"], exp["breakdown_at_k"])
    if want("deflation"):
        _check_deflation(pkg, ctx, comm, P["deflation"], exp["deflation"])
    if want("one_directional"):
        _check_one_directional(pkg, ctx, comm)
    if want("heisenberg_mf"):
        results["heisenberg_mf_E0"] = _check_heisenberg_mf(pkg, ctx, comm, exp["heisenberg_mf"])
    comm.barrier()
    return {k: np.asarray(v).tolist() for k, v in results.items()}


def _check_exhaust(pkg, ctx, comm, pr, ex):
    nb = pr["n"]
    r0, r1 = comm.row_range(nb)
    x0 = syn.start_vector(nb, seed=11)
    op = pkg.DeviceOperator.from_csr(ctx, *shard_of(pr["full"], r0, r1), n_global=nb, row_begin=r0)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setMinIterations(40).setMaxIterations(60)
    es.setMaxEigenvalues(4)
    es.compute()
    assert es.log() == ex["log"], (es.log(), ex["log"])
    assert any("full of Krylov subspace" in line for line in es.log())
    assert es.alpha().size == nb
    Ad = _csr_dense(pr["full"], nb)
    assert np.abs(es.eigenvalues() - np.linalg.eigvalsh(Ad)[:4]).max() < 1e-12
    # back-to-back applies without any reduction in between: the receive buffers alternate correctly
    xs = x0[r0:r1].copy()
    xf = x0.copy()
    for _ in range(5):
        xs = op.apply(xs)
        xf = Ad @ xf
    assert np.abs(xs - xf[r0:r1]).max() < 1e-10 * np.abs(xf).max()
    es.close()
    op.close()


def _check_breakdown(pkg, ctx, comm, pr, ex):
    # exact breakdown at step k (invariant subspace of dimension k spread over the ranks): beta_k is rounding noise, the
    # chain halts on the device at the same step on every rank and the driver reports what the reference reports
    n = pr["n"]
    r0, r1 = comm.row_range(n)
    op = pkg.DeviceOperator.from_csr(ctx, *shard_of(pr["full"], r0, r1), n_global=n, row_begin=r0)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(pr["x0"][r0:r1]).setMinIterations(30).setMaxIterations(40)
    es.setMaxEigenvalues(3)
    es.compute()
    assert es.iterations() == ex["iterations"] == pr["k"] - 1, (es.iterations(), ex["iterations"])
    assert es.log() == ex["log"], (es.log(), ex["log"])
    a, b = es.alpha(), es.beta()
    assert a.size == ex["alpha"].size == pr["k"] and b.size == ex["beta"].size == pr["k"]
    assert np.abs(a - ex["alpha"]).max() < 1e-12
    assert np.abs(b[:-1] - ex["beta"][:-1]).max() < 1e-12 and b[-1] <= 1e-12 and ex["beta"][-1] <= 1e-12
    assert np.abs(es.eigenvalues() - ex["eigenvalues"]).max() < 1e-12
    es.close()
    op.close()


def _check_deflation(pkg, ctx, comm, pr, ex):
    # deflation vectors on a row-partitioned operator (they keep the norm's own reduction, see gram_schmidt2_mailed):
    # two exact eigenvectors of the 2D Laplacian are projected out, Lanczos and Arnoldi must agree with the checker
    n = pr["n"]
    r0, r1 = comm.row_range(n)
    defl, lam = deflation_vectors()
    x0 = syn.start_vector(n, seed=13)
    op = pkg.DeviceOperator.from_csr(ctx, *shard_of(pr["full"], r0, r1), n_global=n, row_begin=r0)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setOrthogonalizingVectors([d[r0:r1] for d in defl])
    es.setMinIterations(50).setMaxIterations(50).setMaxEigenvalues(2)
    es.compute()
    assert np.abs(es.alpha() - ex["alpha"]).max() < 1e-11 and np.abs(es.beta() - ex["beta"]).max() < 1e-11
    # lam[0] is gone; the level lam[1] = lam[2] is doubly degenerate and only one copy was deflated
    assert es.eigenvalues()[0] > lam[1] - 1e-9 and es.eigenvalues()[0] > lam[0] + 1e-3
    Xd = gather_rows(comm, es.eigenvectors())
    assert max(abs(d @ Xd[:, 0]) for d in defl) < 1e-10
    es.close()
    ea = pkg.ArnoldiEigenSolver(np.float64)
    ea.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setOrthogonalizingVectors([d[r0:r1] for d in defl])
    ea.setMinIterations(20).setMaxIterations(20).setMaxEigenvalues(1)
    ea.compute()
    assert np.abs(ea.hessenbergMatrix()[:, :8] - ex["arnoldi_hessenberg"][:, :8]).max() < 1e-10
    ea.close()
    op.close()


def _check_one_directional(pkg, ctx, comm):
    # one-directional coupling (rank q reads from rank q+1 only): ranks that receive nothing still follow the protocol
    nu = 64
    r0, r1 = comm.row_range(nu)
    Au = np.diag(1.0 + 0.05 * np.arange(nu)) + np.diag(0.3 * np.ones(nu - 20), 20)
    op = pkg.DeviceOperator.from_csr(ctx, *shard_of(_dense_to_csr(Au), r0, r1), n_global=nu, row_begin=r0)
    xs, xf = syn.start_vector(nu, seed=3)[r0:r1].copy(), syn.start_vector(nu, seed=3)
    for _ in range(6):
        xs = op.apply(xs)
        xf = Au @ xf
    assert np.abs(xs - xf[r0:r1]).max() < 1e-12 * max(1.0, np.abs(xf).max())
    op.close()


def _check_heisenberg_mf(pkg, ctx, comm, ex):
    # matrix-free Heisenberg ring, slabs exchanged through peer memory (cfg 5 at small L)
    Lm = ex["L"]
    n = 1 << Lm
    r0, r1 = comm.row_range(n)
    x0 = syn.start_vector(n, seed=7)
    op = pkg.DeviceOperator.heisenberg(ctx, Lm, 1.0, True)
    assert op.rows == r1 - r0 and op.height == n
    y = op.apply(x0[r0:r1])
    assert np.abs(y - ex["apply"][r0:r1]).max() < 1e-13
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setMaxIterations(200).setMaxEigenvalues(1)
    es.compute()
    assert abs(es.iterations() - ex["iterations"]) <= 1
    assert abs(es.eigenvalues()[0] - ex["eigenvalues"][0]) < 1e-10 * abs(ex["eigenvalues"][0])
    e0 = es.eigenvalues()
    es.close()
    op.close()
    return e0
