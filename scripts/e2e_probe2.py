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
es2 = pkg.LanczosEigenSolver()
es2.setInitialVector(px.array)
es2.setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(nev).setReserveSize(m + 1).setIndicesForConvergence(list(range(nev)))
for i in range(5):
    T = []
    ctx.sync(); t0 = time.perf_counter()
    op2 = pkg.DeviceOperator.from_csr(ctx, prp.array, pc.array, pv.array, n_global=n, row_begin=0); T.append(time.perf_counter() - t0)
    es2.setMatrixMultiplication(op2).setInitialVector(px.array); T.append(time.perf_counter() - t0)
    es2.compute(); T.append(time.perf_counter() - t0)
    ev = es2.eigenvalues(); X = es2.eigenvectors(copy=False); chk = float(X[0, 0]) + float(ev[0]); T.append(time.perf_counter() - t0)
    op2.close(); T.append(time.perf_counter() - t0)
    ctx.sync(); T.append(time.perf_counter() - t0)
    print(i, [round(x * 1e3, 1) for x in T])
