"""Times the pieces of the end-to-end path of bench.py (host buffers in, host results out)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import capi, synthetic as syn

N, m, nev = 4096, 100, 5
n = N * N
rp, c, v = syn.laplacian2d_csr(N)
prp, pc, pv = capi.PinnedBuffer(rp.shape, np.int64), capi.PinnedBuffer(c.shape, np.int32), capi.PinnedBuffer(v.shape, np.float64)
prp.array[:], pc.array[:], pv.array[:] = rp, c, v
px = capi.PinnedBuffer((n,), np.float64)
px.array[:] = syn.start_vector(n, seed=7)
ctx = pkg.Context(0)
es = pkg.LanczosEigenSolver()
es.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(nev).setReserveSize(m + 1)
for it in range(3):
    T = {}
    def tick(name, t0):
        ctx.sync(); T[name] = time.perf_counter() - t0
    t0 = time.perf_counter(); op = pkg.DeviceOperator.from_csr(ctx, prp.array, pc.array, pv.array); tick("op_create", t0)
    t0 = time.perf_counter(); es.setMatrixMultiplication(op).setInitialVector(px.array); tick("set_inputs", t0)
    es.setComputeEigenvectorsOn(False)
    t0 = time.perf_counter(); es.compute(); tick("compute_novec", t0)
    es.setComputeEigenvectorsOn(True)
    t0 = time.perf_counter(); es.compute(); tick("compute_vec", t0)
    t0 = time.perf_counter(); ev = es.eigenvalues(); X = es.eigenvectors(copy=False); s = float(X[0, 0]); tick("fetch", t0)
    t0 = time.perf_counter(); op.close(); tick("op_close", t0)
    print(it, {k: round(v * 1e3, 1) for k, v in T.items()})
