"""Runs the BASELINE configs 1, 3, 4 and the single-GPU variant of 5 on one B200 (plus bounded CPU-oracle samples)
and prints one JSON line per config: the numbers behind the table in BASELINE.md §3 / DESIGN.md §6.

usage: run_configs.py [cfg1] [cfg3] [cfg4] [cfg5] [--no-cpu] [--L4 24] [--L5 26]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cmpt_eigenex_b200 as pkg  # noqa: E402
from cmpt_eigenex_b200 import synthetic as syn  # noqa: E402

args = sys.argv[1:]
want = [a for a in args if a.startswith("cfg")] or ["cfg1", "cfg3", "cfg4", "cfg5"]
do_cpu = "--no-cpu" not in args


def opt(name, default):
    return int(args[args.index(name) + 1]) if name in args else default


ctx = pkg.Context(0)


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ctx.sync()
    t = []
    for _ in range(reps):
        ctx.flush_l2()
        ctx.sync()
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        t.append(time.perf_counter() - t0)
    return min(t), float(np.mean(t))


def cpu_sample(make_solver, m_sample, threads):
    from oracle import core

    core.set_num_threads(threads)
    ref = make_solver()
    ref.min_iterations = ref.max_iterations = m_sample
    ref.compute_eigenvectors_on = False
    t0 = time.perf_counter()
    ref.compute()
    return m_sample / (time.perf_counter() - t0)


FAMS = ("cgs_dot", "cgs_update_dot", "cgs_update_norm", "spmv_sell", "heisenberg_mf", "gemv_dense", "vec_dot")


def families(fn):
    """per-family device time (ms) and launch count of one call of fn (CUDA events around every launch)"""
    ctx.sync()
    ctx.profile(True)
    fn()
    ctx.sync()
    out = {}
    for f in FAMS:
        ms, cnt = ctx.profile_get(f)
        if cnt:
            out[f] = [round(ms, 3), cnt]
    ctx.profile(False)
    return out


def emit(d):
    print(json.dumps(d), flush=True)


if "cfg1" in want:
    from oracle import core, reference_solvers as rs

    n, m = 2000, 100
    A = syn.dense_symmetric(n, seed=1)
    x0 = syn.start_vector(n, seed=7)
    op = pkg.DeviceOperator.from_dense(ctx, A)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5)
    es.setIndicesForConvergence([0, 1, 2, 3, 4])
    best, mean = timed(es.compute, reps=5)
    core.set_num_threads(1)
    ref = rs.LanczosEigenSolver("d")
    ref.set_matrix_multiplication(core.Operator.dense(A))
    ref.init, ref.max_eigenvalues, ref.indices_for_convergence = x0, 5, [0, 1, 2, 3, 4]
    ref.min_iterations = ref.max_iterations = m
    t0 = time.perf_counter()
    ref.compute()
    t_cpu1 = time.perf_counter() - t0
    rel = np.abs(es.eigenvalues() - ref.eigenvalues) / np.abs(ref.eigenvalues)
    emit({"cfg": 1, "what": "dense symmetric n=2000, Lanczos m=100 (with 5 Ritz vectors)", "gpu_it_per_s": m / best,
          "gpu_ms": best * 1e3, "cpu_1thread_it_per_s": m / t_cpu1, "max_rel_eig_diff_vs_oracle": float(rel.max()),
          "eigenvalues": es.eigenvalues().tolist(), "launches_per_solve": None})
    es.close()
    op.close()

if "cfg3" in want:
    from oracle import core, reference_solvers as rs

    M, m, cycles = 256, 50, 3
    n = M ** 3
    rp, c, v = syn.convdiff3d_csr(M)
    assert rp[-1] == 117047296
    x0 = syn.start_vector(n, seed=7)
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    es = pkg.ArnoldiEigenSolver(np.float64)
    es.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5)
    es.setIndicesForConvergence([0, 1, 2, 3, 4]).setComputeEigenvectorsOn(False).setReserveSize(m + 1)
    best, mean = timed(es.compute, reps=3)
    cyc_bytes = es.deviceBytes()
    # restarted run (explicit restart, needs the leading Ritz vector)
    es.setComputeEigenvectorsOn(True).setMaxEigenvalues(1)
    ctx.sync()
    t0 = time.perf_counter()
    es.computeWithRestarts(cycles)
    ctx.sync()
    t_restart = time.perf_counter() - t0
    lead = es.eigenvalues()[0]
    exact = syn.convdiff3d_eigenvalues(M, count=1)[0]
    out = {"cfg": 3, "what": "3D conv-diff 256^3 CSR, Arnoldi m=50 per cycle (real Scalar)", "gpu_it_per_s": m / best,
           "gpu_ms_per_cycle": best * 1e3, "algorithmic_GBps": cyc_bytes / best / 1e9,
           "restart_cycles": cycles, "restart_s": t_restart, "leading_ritz_after_restarts": [lead.real, lead.imag],
           "exact_leading": float(exact), "ritz_residual": float(es.ritzResiduals()[0])}
    if do_cpu:
        def mk():
            r = rs.ArnoldiEigenSolver("d")
            r.set_matrix_multiplication(core.Operator.csr(rp, c, v))
            r.init, r.max_eigenvalues = x0, 5
            return r
        out["cpu_1thread_it_per_s_first10"] = cpu_sample(mk, 10, 1)
        out["cpu_allthreads_it_per_s_first10"] = cpu_sample(mk, 10, os.cpu_count())
        out["cpu_threads"] = os.cpu_count()
    emit(out)
    es.close()
    op.close()
    del rp, c, v

if "cfg4" in want:
    from oracle import core, reference_solvers as rs

    L = opt("--L4", 24)
    n = 1 << L
    t0 = time.perf_counter()
    rp, c, v = syn.heisenberg_csr(L)
    t_gen = time.perf_counter() - t0
    x0 = syn.start_vector(n, seed=7)
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0).setMaxIterations(200).setMaxEigenvalues(1).setReserveSize(128)
    es.setComputeEigenvectorsOn(False)
    ctx.sync()
    t0 = time.perf_counter()
    es.compute()
    ctx.sync()
    t_conv = time.perf_counter() - t0
    e0, it_conv = float(es.eigenvalues()[0]), es.iterations()
    m = 100
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5).setIndicesForConvergence([0, 1, 2, 3, 4])
    best, mean = timed(es.compute, reps=3)
    out = {"cfg": 4, "what": "Heisenberg ring L=%d explicit CSR (nnz=%d)" % (L, int(rp[-1])), "csr_generation_s": t_gen,
           "converged_iterations": it_conv, "converged_s": t_conv, "E0": e0,
           "E0_reference": syn.HEISENBERG_RING_E0.get(L), "gpu_it_per_s_m100": m / best, "gpu_ms_m100": best * 1e3,
           "algorithmic_GBps": es.deviceBytes() / best / 1e9}
    if do_cpu:
        def mk():
            r = rs.LanczosEigenSolver("d")
            r.set_matrix_multiplication(core.Operator.csr(rp, c, v))
            r.init, r.max_eigenvalues = x0, 5
            return r
        out["cpu_1thread_it_per_s_first10"] = cpu_sample(mk, 10, 1)
        out["cpu_allthreads_it_per_s_first10"] = cpu_sample(mk, 10, os.cpu_count())
        out["cpu_threads"] = os.cpu_count()
    emit(out)
    es.close()
    op.close()
    del rp, c, v

if "cfg5" in want:
    L = opt("--L5", 26)
    n = 1 << L
    x0 = syn.start_vector(n, seed=7)
    op = pkg.DeviceOperator.heisenberg(ctx, L)
    es = pkg.LanczosEigenSolver()
    m = 40
    es.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(1)
    es.setComputeEigenvectorsOn(False).setReserveSize(m + 1)
    best, mean = timed(es.compute, reps=3)
    rr = es.ritzResiduals()
    fam5 = families(es.compute)
    emit({"cfg": 5, "families_ms": fam5, "what": "matrix-free Heisenberg ring L=%d on ONE GPU (cfg 5 is L=30 on 8), Lanczos m=40" % L,
          "gpu_it_per_s": m / best, "gpu_ms": best * 1e3, "algorithmic_GBps": es.deviceBytes() / best / 1e9,
          "lowest_ritz": float(es.eigenvalues()[0]), "E0_per_site": float(es.eigenvalues()[0]) / L,
          "ritz_residual": float(rr[0])})
    es.close()
    op.close()
