"""Ground state and first excited level of the spin-1/2 Heisenberg ring with a bounded Lanczos basis
(ThickRestartLanczos on the matrix-free operator).  usage: thick_restart_heisenberg.py [L=24] [maxBasis=24]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cmpt_eigenex_b200 as pkg  # noqa: E402
from cmpt_eigenex_b200 import synthetic as syn  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 24
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ctx = pkg.Context(0)
op = pkg.DeviceOperator.heisenberg(ctx, L)
tr = pkg.ThickRestartLanczos(np.float64)
tr.setMatrixMultiplication(op).setInitialVector(syn.start_vector(1 << L, seed=7))
tr.setWanted(2).setMaxBasis(mb).setTolerance(1e-9).setMaxRestarts(300).setComputeEigenvectorsOn(False)
ctx.sync()
t0 = time.perf_counter()
tr.compute()
ctx.sync()
dt = time.perf_counter() - t0
print(json.dumps({"L": L, "max_basis": mb, "basis_GB": mb * (8 << L) / 1e9, "seconds": dt, "restarts": tr.restarts(),
                  "operator_applications": tr.operatorApplications(), "converged": tr.converged(),
                  "E0_E1": tr.eigenvalues().tolist(), "residual_bounds": tr.residuals().tolist(),
                  "reference_L24": [-10.670014516537, -10.487293480731]}))
