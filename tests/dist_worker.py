"""Worker of the multi-rank tests: launched by torch.distributed.run, one process per GPU (or, with
--cpu, one process per rank on the CPU for the host-side logic only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cmpt_eigenex_b200 as pkg  # noqa: E402
from cmpt_eigenex_b200 import capi, dist  # noqa: E402
from cmpt_eigenex_b200 import synthetic as syn  # noqa: E402


def gather_rows(local, n, td):
    import torch

    world = td.get_world_size()
    parts = [None] * world
    td.all_gather_object(parts, np.asarray(local))
    return np.concatenate(parts)


def cpu_main():
    """host-side logic at world_size 2 over gloo: partition, halo plans, id broadcast."""
    import ctypes as C

    import torch

    td = dist.init()
    rank, world = td.get_rank(), td.get_world_size()
    L = capi.lib()
    N = 12
    n = N * N
    r0, r1 = dist.row_range(n)
    assert (r0, r1) == (L.cmb_partition_begin(n, world, rank), L.cmb_partition_begin(n, world, rank + 1))
    rp, c, v = syn.laplacian2d_csr(N, r0, r1)
    col_local = np.empty_like(c)
    cnt = C.c_int64()
    per = np.zeros(world, np.int64)
    halo = np.empty(c.size, np.int32)
    capi.check(L.cmb_plan_halo(n, world, rank, c.size, capi.ptr(c), capi.ptr(col_local), C.byref(cnt), capi.ptr(per),
                               capi.ptr(halo), halo.size))
    halo = halo[: cnt.value]
    # what I need is exactly the set of remote columns, owned by the other rank
    remote = np.unique(c[(c < r0) | (c >= r1)])
    assert np.array_equal(halo, remote) and per[rank] == 0 and per.sum() == halo.size
    assert halo.size == N  # one grid line of the neighbour block
    own = (c >= r0) & (c < r1)
    assert np.array_equal(col_local[own], c[own] - r0)
    assert np.array_equal(halo[col_local[~own] - (r1 - r0)], c[~own])
    # the lists are consistent across ranks: every index I ask for is owned by the peer
    lists = [None] * world
    td.all_gather_object(lists, (r0, r1, halo.tolist()))
    for q, (q0, q1, need) in enumerate(lists):
        for g in need:
            owner = [k for k, (a, b, _) in enumerate(lists) if a <= g < b]
            assert owner and owner[0] != q
    # 128-byte token broadcast (the path the NCCL id takes)
    tok = np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, np.uint8)
    t = torch.from_numpy(tok)
    td.broadcast(t, src=0)
    assert np.array_equal(t.numpy(), np.arange(128, dtype=np.uint8))
    assert dist.all_max(float(rank)) == world - 1 and dist.all_sum(1.0) == world
    td.barrier()
    if rank == 0:
        print("DIST_CPU_OK")


def gpu_main():
    from oracle import core
    from oracle import reference_solvers as rs

    td = dist.init()
    rank, world = td.get_rank(), td.get_world_size()
    ctx = dist.make_context()
    core.set_num_threads(2)
    results = {}
    for name in ("laplacian", "heisenberg", "convdiff_arnoldi"):
        if name == "laplacian":
            N, m = 40, 60
            n = N * N
            full = syn.laplacian2d_csr(N)
            r0, r1 = dist.row_range(n)
            shard = syn.laplacian2d_csr(N, r0, r1)
        elif name == "heisenberg":
            Ls, m = 12, 40
            n = 1 << Ls
            full = syn.heisenberg_csr(Ls)
            r0, r1 = dist.row_range(n)
            shard = syn.heisenberg_csr(Ls, r0=r0, r1=r1)
        else:
            M, m = 9, 30
            n = M ** 3
            full = syn.convdiff3d_csr(M)
            r0, r1 = dist.row_range(n)
            shard = syn.convdiff3d_csr(M, r0=r0, r1=r1)
        x0 = syn.start_vector(n, seed=7)
        op = pkg.DeviceOperator.from_csr(ctx, *shard, n_global=n, row_begin=r0)
        # operator apply: local slab of A x
        y = op.apply(x0[r0:r1])
        yr = core.Operator.csr(*full).apply(x0)
        assert np.abs(y - yr[r0:r1]).max() < 1e-13, name
        if name != "convdiff_arnoldi":
            es = pkg.LanczosEigenSolver()
            es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1])
            es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(3).setIndicesForConvergence([0, 1, 2])
            es.compute()
            ref = rs.LanczosEigenSolver("d")
            ref.set_matrix_multiplication(core.Operator.csr(*full))
            ref.init = x0
            ref.min_iterations = ref.max_iterations = m
            ref.max_eigenvalues = 3
            ref.indices_for_convergence = [0, 1, 2]
            ref.compute()
            ra, rb = ref.alpha_beta()
            assert es.iterations() == m
            assert np.abs(es.alpha() - ra).max() < 1e-11 and np.abs(es.beta() - rb).max() < 1e-11, name
            assert np.abs(es.eigenvalues() - ref.eigenvalues).max() < 1e-10 * np.abs(ra).max(), name
            X = es.eigenvectors()
            assert X.shape == (r1 - r0, 3)
            Xfull = gather_rows(X, n, td)
            ov = np.abs(np.sum(ref.eigenvectors * Xfull, axis=0))
            assert np.abs(ov - 1).max() < 1e-8, (name, ov)
            assert np.all(Xfull[0] > 0)
            assert es.log() == ref.log
            rr = es.ritzResiduals()
            assert np.abs(rr - ref.ritz_residuals()).max() < 1e-9
            results[name] = es.eigenvalues()
            es.close()
        else:
            es = pkg.ArnoldiEigenSolver(np.float64)
            es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1])
            es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(2)
            es.compute()
            ref = rs.ArnoldiEigenSolver("d")
            ref.set_matrix_multiplication(core.Operator.csr(*full))
            ref.init = x0
            ref.min_iterations = ref.max_iterations = m
            ref.max_eigenvalues = 2
            ref.compute()
            assert abs(es.eigenvalues()[0] - ref.eigenvalues[0]) < 1e-8 * abs(ref.eigenvalues[0]), name
            H = es.hessenbergMatrix()
            assert np.abs(H[:, :6] - ref.hessenberg[:, :6]).max() < 1e-11
            P = gather_rows(es.eigenvectors(), n, td)
            assert np.abs(np.linalg.norm(P, axis=0) - 1).max() < 1e-12
            A = np.zeros((n, n))
            rp, c, v = full
            for r in range(n):
                A[r, c[rp[r]:rp[r + 1]]] = v[rp[r]:rp[r + 1]]
            res = np.linalg.norm(A @ P - P * es.eigenvalues(), axis=0)
            assert np.all(np.abs(res - es.ritzResiduals()) < 1e-9)
            es.close()
        op.close()
    # Krylov space exhausted on a row-partitioned operator: every rank must take the same halting decision on the
    # device (beta^2 from the reduced coefficients) and the chain must stop with the reference's log lines
    nb = 24
    r0, r1 = dist.row_range(nb)
    rp = np.zeros(nb + 1, np.int64)
    cols, vals = [], []
    for r in range(nb):
        for c_, v_ in ((r - 1, -1.0), (r, 2.0 + 0.1 * r), (r + 1, -1.0)):
            if 0 <= c_ < nb:
                cols.append(c_)
                vals.append(v_)
        rp[r + 1] = len(cols)
    cols, vals = np.array(cols, np.int32), np.array(vals)
    full = (rp, cols, vals)
    shard = (rp[r0:r1 + 1] - rp[r0], cols[rp[r0]:rp[r1]], vals[rp[r0]:rp[r1]])
    x0 = syn.start_vector(nb, seed=11)
    op = pkg.DeviceOperator.from_csr(ctx, *shard, n_global=nb, row_begin=r0)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setMinIterations(40).setMaxIterations(60)
    es.setMaxEigenvalues(4)
    es.compute()
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(*full))
    ref.init, ref.min_iterations, ref.max_iterations, ref.max_eigenvalues = x0, 40, 60, 4
    ref.compute()
    assert es.log() == ref.log, (es.log(), ref.log)
    assert any("full of Krylov subspace" in line for line in es.log())
    assert es.alpha().size == nb
    Ad = np.zeros((nb, nb))
    for r in range(nb):
        Ad[r, cols[rp[r]:rp[r + 1]]] = vals[rp[r]:rp[r + 1]]
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
    # deflation vectors on a row-partitioned operator (they keep the norm's own reduction, see gram_schmidt2_mailed):
    # two exact eigenvectors of the 2D Laplacian are projected out, Lanczos and Arnoldi must agree with the oracle
    N = 16
    n = N * N
    full = syn.laplacian2d_csr(N)
    r0, r1 = dist.row_range(n)
    shard = syn.laplacian2d_csr(N, r0, r1)
    ii, jj = np.meshgrid(np.arange(1, N + 1), np.arange(1, N + 1), indexing="ij")
    defl = []
    for (p_, q_) in ((1, 1), (1, 2)):
        e = (np.sin(p_ * np.pi * ii / (N + 1)) * np.sin(q_ * np.pi * jj / (N + 1))).reshape(-1)
        defl.append(e / np.linalg.norm(e))
    x0 = syn.start_vector(n, seed=13)
    op = pkg.DeviceOperator.from_csr(ctx, *shard, n_global=n, row_begin=r0)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setOrthogonalizingVectors([d[r0:r1] for d in defl])
    es.setMinIterations(50).setMaxIterations(50).setMaxEigenvalues(2)
    es.compute()
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.csr(*full))
    ref.init, ref.ortho = x0, defl
    ref.min_iterations = ref.max_iterations = 50
    ref.max_eigenvalues = 2
    ref.compute()
    ra, rb = ref.alpha_beta()
    assert np.abs(es.alpha() - ra).max() < 1e-11 and np.abs(es.beta() - rb).max() < 1e-11
    lam = np.sort((4 - 2 * np.cos(ii * np.pi / (N + 1)) - 2 * np.cos(jj * np.pi / (N + 1))).reshape(-1))
    # lam[0] is gone; the level lam[1] = lam[2] is doubly degenerate and only one copy was deflated
    assert es.eigenvalues()[0] > lam[1] - 1e-9 and es.eigenvalues()[0] > lam[0] + 1e-3
    Xd = gather_rows(es.eigenvectors(), n, td)
    assert max(abs(d @ Xd[:, 0]) for d in defl) < 1e-10
    es.close()
    ea = pkg.ArnoldiEigenSolver(np.float64)
    ea.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setOrthogonalizingVectors([d[r0:r1] for d in defl])
    ea.setMinIterations(20).setMaxIterations(20).setMaxEigenvalues(1)
    ea.compute()
    refa = rs.ArnoldiEigenSolver("d")
    refa.set_matrix_multiplication(core.Operator.csr(*full))
    refa.init, refa.ortho = x0, defl
    refa.min_iterations = refa.max_iterations = 20
    refa.max_eigenvalues = 1
    refa.compute()
    assert np.abs(ea.hessenbergMatrix()[:, :8] - refa.hessenberg[:, :8]).max() < 1e-10
    ea.close()
    op.close()
    # one-directional coupling (rank q reads from rank q+1 only): ranks that receive nothing still follow the protocol
    nu = 64
    r0, r1 = dist.row_range(nu)
    Au = np.diag(1.0 + 0.05 * np.arange(nu)) + np.diag(0.3 * np.ones(nu - 20), 20)
    rpu = np.zeros(r1 - r0 + 1, np.int64)
    cu, vu = [], []
    for r in range(r0, r1):
        nzc = np.nonzero(Au[r])[0]
        cu += nzc.tolist()
        vu += Au[r, nzc].tolist()
        rpu[r - r0 + 1] = len(cu)
    op = pkg.DeviceOperator.from_csr(ctx, rpu, np.array(cu, np.int32), np.array(vu), n_global=nu, row_begin=r0)
    xs, xf = syn.start_vector(nu, seed=3)[r0:r1].copy(), syn.start_vector(nu, seed=3)
    for _ in range(6):
        xs = op.apply(xs)
        xf = Au @ xf
    assert np.abs(xs - xf[r0:r1]).max() < 1e-12 * max(1.0, np.abs(xf).max())
    op.close()
    # matrix-free Heisenberg ring, slabs exchanged over NVLink (cfg 5 at small L)
    Lm = 14
    n = 1 << Lm
    r0, r1 = dist.row_range(n)
    x0 = syn.start_vector(n, seed=7)
    op = pkg.DeviceOperator.heisenberg(ctx, Lm, 1.0, True)
    assert op.rows == r1 - r0 and op.height == n
    y = op.apply(x0[r0:r1])
    yr = core.Operator.heisenberg(Lm).apply(x0)
    assert np.abs(y - yr[r0:r1]).max() < 1e-13
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setMaxIterations(200).setMaxEigenvalues(1)
    es.compute()
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.heisenberg(Lm))
    ref.init, ref.max_iterations, ref.max_eigenvalues = x0, 200, 1
    ref.compute()
    assert abs(es.iterations() - ref.iterations) <= 1
    assert abs(es.eigenvalues()[0] - ref.eigenvalues[0]) < 1e-10 * abs(ref.eigenvalues[0])
    results["heisenberg_mf_E0"] = es.eigenvalues()
    es.close()
    op.close()
    td.barrier()
    ctx.close()
    if rank == 0:
        print("DIST_GPU_OK", {k: v.tolist() for k, v in results.items()})


if __name__ == "__main__":
    if "--cpu" in sys.argv:
        cpu_main()
    else:
        gpu_main()
