"""Tiny driver for profiling the matrix-free Heisenberg apply: L sites on one GPU, a few Lanczos steps.
usage: heis_probe.py [L] [m]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cmpt_eigenex_b200 as pkg  # noqa: E402
from cmpt_eigenex_b200 import synthetic as syn  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 26
m = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = pkg.Context(0)
op = pkg.DeviceOperator.heisenberg(ctx, L)
es = pkg.LanczosEigenSolver()
es.setMatrixMultiplication(op).setInitialVector(syn.start_vector(1 << L, seed=7))
es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(1).setComputeEigenvectorsOn(False).setReserveSize(m + 1)
es.compute()
ctx.sync()
print("lowest ritz", float(es.eigenvalues()[0]))
es.close()
op.close()
ctx.close()
