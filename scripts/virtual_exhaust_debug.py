import ctypes as C
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import cmpt_eigenex_b200 as pkg  # noqa: E402
import multirank_checks as mc  # noqa: E402
from cmpt_eigenex_b200 import capi, synthetic as syn  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pr = mc.problems()["exhaust"]
nb = pr["n"]
x0 = syn.start_vector(nb, seed=11)


def xcount(op):
    c = C.c_longlong()
    capi.check(capi.lib().cmb_debug_op_exchange_count(op.h, C.byref(c)))
    return c.value


from oracle import core  # noqa: E402
from oracle import reference_solvers as rs  # noqa: E402

core.set_num_threads(2)
PRE = sys.argv[2].split(",") if len(sys.argv) > 2 else []
exp = mc.expected(rs, core) if PRE else None


def work(ctx, comm):
    ctx.set_spin_timeout(3.0)
    if PRE:
        mc.run_checks(pkg, ctx, comm, exp, only=PRE)
        print(comm.rank, "pre-sections done", flush=True)
    r0, r1 = comm.row_range(nb)
    op = pkg.DeviceOperator.from_csr(ctx, *mc.shard_of(pr["full"], r0, r1), n_global=nb, row_begin=r0)
    print(comm.rank, "exchanges after build", xcount(op), flush=True)
    es = pkg.LanczosEigenSolver()
    es.setMatrixMultiplication(op).setInitialVector(x0[r0:r1]).setMinIterations(40).setMaxIterations(60).setMaxEigenvalues(4)
    es.setComputeEigenvectorsOn(os.environ.get("VEC", "1") == "1")
    if os.environ.get("BARRIER0"):
        comm.barrier()
    es.compute()
    print(comm.rank, "after compute: exchanges", xcount(op), "iterations", es.iterations(), "nalpha", es.alpha().size, "beta tail", es.beta()[-2:], es.log(), flush=True)
    if os.environ.get("BARRIER1", "1") == "1":
        comm.barrier()
    xs = x0[r0:r1].copy()
    for i in range(3):
        xs = op.apply(xs)
        print(comm.rank, "apply", i, "exchanges", xcount(op), flush=True)
    es.close()
    op.close()


try:
    pkg.run_virtual_ranks(P, work, timeout=100)
    print("OK")
except Exception as e:
    print("ERROR", str(e)[:200])
