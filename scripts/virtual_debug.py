"""Debug driver for virtual ranks: a short row-partitioned Lanczos run with progress prints.
usage: python scripts/virtual_debug.py [nranks] [m]"""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import cmpt_eigenex_b200 as pkg  # noqa: E402
import multirank_checks as mc  # noqa: E402
from cmpt_eigenex_b200 import synthetic as syn  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 40
n = N * N
full = syn.laplacian2d_csr(N)
x0 = syn.start_vector(n, seed=7)
g = pkg.VirtualGroup(0, P)
print("group", g.info(), flush=True)
g.close()


def work(ctx, comm):
    ctx.set_spin_timeout(float(os.environ.get("SPIN", "5")))
    r0, r1 = comm.row_range(n)
    op = pkg.DeviceOperator.from_csr(ctx, *mc.shard_of(full, r0, r1), n_global=n, row_begin=r0)
    print(comm.rank, "operator built", flush=True)
    y = op.apply(x0[r0:r1])
    print(comm.rank, "apply ok", float(np.abs(y).max()), flush=True)
    for steps in range(1, m + 1):
        es = pkg.LanczosEigenSolver()
        es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setMinIterations(steps).setMaxIterations(steps).setMaxEigenvalues(1)
        es.setComputeEigenvectorsOn(False)
        t = time.time()
        es.compute()
        print(comm.rank, "lanczos", steps, "ok", es.alpha()[:3], "%.3fs" % (time.time() - t), flush=True)
        es.close()
    op.close()
    return 0


print(pkg.run_virtual_ranks(P, work, timeout=120))
