"""Times the phases of one Lanczos solve through the raw C-ABI (start vector upload, step chain)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cmpt_eigenex_b200 as pkg
from cmpt_eigenex_b200 import capi, synthetic as syn
from cmpt_eigenex_b200.capi import check, lib, ptr
N, m = 4096, 100
n = N * N
rp, c, v = syn.laplacian2d_csr(N)
ctx = pkg.Context(0)
op = pkg.DeviceOperator.from_csr(ctx, rp, c, v)
x0 = syn.start_vector(n, seed=7)
px = capi.PinnedBuffer((n,), np.float64); px.array[:] = x0
K = C.c_void_p()
check(lib().cmb_krylov_create(ctx.h, 0, n, 0, n, m + 1, C.byref(K)))
a = np.zeros(m + 2); b = np.zeros(m + 2); done = C.c_int64(); st = C.c_int()
for it in range(3):
    for name, buf in (("pageable", x0), ("pinned", px.array)):
        ctx.sync(); t0 = time.perf_counter()
        check(lib().cmb_krylov_start(K, ptr(buf), 1e-12, C.byref(st))); t1 = time.perf_counter()
        check(lib().cmb_lanczos_run(K, op.h, 0.0, 1, 1e-12, m + 1, ptr(a), ptr(b), C.byref(done), C.byref(st))); t2 = time.perf_counter()
        print(it, name, "start %.2f ms  run %.2f ms  (steps %d)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, done.value))
# host-side tridiagonal work as the solver class does it
t0 = time.perf_counter()
for j in range(1, m + 2):
    pkg.host_tridiagonal_eigen(a[:j], b[:max(j - 1, 0)], vectors=False)
t1 = time.perf_counter()
pkg.host_tridiagonal_eigen(a[:m + 1], b[:m], vectors=True)
t2 = time.perf_counter()
print("serial replay %.2f ms, final eigensystem %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
es = pkg.LanczosEigenSolver(); es.setMatrixMultiplication(op).setInitialVector(px.array).setMinIterations(m).setMaxIterations(m).setMaxEigenvalues(5).setComputeEigenvectorsOn(False).setReserveSize(m + 1).setIndicesForConvergence([0,1,2,3,4])
for it in range(3):
    ctx.sync(); t0 = time.perf_counter(); es.compute(); ctx.sync(); print("compute() wall %.2f ms" % ((time.perf_counter() - t0) * 1e3))
