"""Row-partitioned runs of cfg 4 (Heisenberg ring as explicit CSR) and cfg 5 (matrix-free Heisenberg ring) under
torchrun: one JSON line per config from rank 0.
usage: torchrun ... run_dist_configs.py [cfg4] [cfg5] [--L4 24] [--L5 30] [--m5 40]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cmpt_eigenex_b200 as pkg  # noqa: E402
from cmpt_eigenex_b200 import dist, synthetic as syn  # noqa: E402

args = sys.argv[1:]
want = [a for a in args if a.startswith("cfg")] or ["cfg4", "cfg5"]


def opt(name, default):
    return int(args[args.index(name) + 1]) if name in args else default


td = dist.init()
rank, world = td.get_rank(), td.get_world_size()
ctx = dist.make_context()


def slab_start_vector(n, r0, r1, seed=7):
    """local slab of the normalised global start vector (norm over all ranks)"""
    x = np.empty(r1 - r0)
    chunk = 1 << 22
    for s in range(r0, r1, chunk):
        c = min(chunk, r1 - s)
        x[s - r0: s - r0 + c] = 2.0 * syn.uniform01(seed, s, c) - 1.0
    nrm2 = dist.all_sum(float(x @ x))
    return x / np.sqrt(nrm2)


def timed(fn, reps=2, warm=1):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        ctx.sync()
        dist.barrier()
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        best = min(best, dist.all_max(time.perf_counter() - t0))
    return best


if "cfg4" in want:
    L = opt("--L4", 24)
    n = 1 << L
    r0, r1 = dist.row_range(n)
    t0 = time.perf_counter()
    rp, c, v = syn.heisenberg_csr(L, r0=r0, r1=r1)
    t_gen = time.perf_counter() - t0
    x0 = slab_start_vector(n, r0, r1)
    t0 = time.perf_counter()
    op = pkg.DeviceOperator.from_csr(ctx, rp, c, v, n_global=n, row_begin=r0)
    ctx.sync()
    t_build = dist.all_max(time.perf_counter() - t0)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0).setMaxIterations(200).setMaxEigenvalues(1)
    es.setComputeEigenvectorsOn(False).setReserveSize(128)
    t_conv = timed(es.compute, reps=1, warm=1)
    e0, it_conv = float(es.eigenvalues()[0]), es.iterations()
    m = 100
    es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5).setIndicesForConvergence([0, 1, 2, 3, 4])
    best = timed(es.compute, reps=3)
    byts = dist.all_sum(es.deviceBytes())
    if rank == 0:
        print(json.dumps({"cfg": 4, "n_gpus": world, "what": "Heisenberg ring L=%d explicit CSR, row-partitioned" % L,
                          "csr_generation_s": t_gen, "operator_build_s": t_build, "converged_iterations": it_conv,
                          "converged_s": t_conv, "E0": e0, "E0_reference": syn.HEISENBERG_RING_E0.get(L),
                          "it_per_s_m100": m / best, "ms_m100": best * 1e3, "algorithmic_GBps_total": byts / best / 1e9}),
              flush=True)
    es.close()
    op.close()
    del rp, c, v

if "cfg5" in want:
    L, m = opt("--L5", 30), opt("--m5", 40)
    n = 1 << L
    r0, r1 = dist.row_range(n)
    x0 = slab_start_vector(n, r0, r1)
    op = pkg.DeviceOperator.heisenberg(ctx, L)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(1)
    es.setComputeEigenvectorsOn(False).setReserveSize(m + 1)
    best = timed(es.compute, reps=2)
    rr = es.ritzResiduals()
    byts = dist.all_sum(es.deviceBytes())
    ctx.sync()
    ctx.profile(True)
    es.compute()
    ctx.sync()
    fams = {}
    for f in ("cgs_dot", "cgs_update_dot", "cgs_update_norm", "heisenberg_mf", "halo_pack", "nccl_allreduce"):
        ms, cnt = ctx.profile_get(f)
        if cnt:
            fams[f] = [round(ms, 2), cnt]
    ctx.profile(False)
    if rank == 0:
        print(json.dumps({"cfg": 5, "families_ms_rank0": fams, "n_gpus": world, "what": "matrix-free Heisenberg ring L=%d, Lanczos m=%d" % (L, m),
                          "it_per_s": m / best, "ms": best * 1e3, "algorithmic_GBps_total": byts / best / 1e9,
                          "lowest_ritz": float(es.eigenvalues()[0]), "E0_per_site": float(es.eigenvalues()[0]) / L,
                          "ritz_residual": float(rr[0])}), flush=True)
    es.close()
    op.close()
dist.barrier()
ctx.close()
